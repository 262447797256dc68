/* Minimal stand-in for lv2/time (TEST INFRASTRUCTURE ONLY): URI strings only. */
#ifndef LV2_TIME_H
#define LV2_TIME_H
#define LV2_TIME_URI "http://lv2plug.in/ns/ext/time"
#define LV2_TIME_PREFIX LV2_TIME_URI "#"
#define LV2_TIME__Position LV2_TIME_PREFIX "Position"
#define LV2_TIME__frame LV2_TIME_PREFIX "frame"
#define LV2_TIME__speed LV2_TIME_PREFIX "speed"
#define LV2_TIME__bar LV2_TIME_PREFIX "bar"
#define LV2_TIME__barBeat LV2_TIME_PREFIX "barBeat"
#define LV2_TIME__beatUnit LV2_TIME_PREFIX "beatUnit"
#define LV2_TIME__beatsPerBar LV2_TIME_PREFIX "beatsPerBar"
#define LV2_TIME__beatsPerMinute LV2_TIME_PREFIX "beatsPerMinute"
#endif
