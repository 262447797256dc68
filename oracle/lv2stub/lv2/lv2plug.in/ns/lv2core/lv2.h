/* Minimal stand-in for the LV2 core header (TEST INFRASTRUCTURE ONLY, see oracle/lv2stub/README).
 * The LV2 SDK is not installed in this image; these declarations restate the public LV2 C ABI from the LV2
 * specification, just enough for the reference's src/meters.cc to compile unmodified into oracle/_ref. */
#ifndef LV2_H_INCLUDED
#define LV2_H_INCLUDED
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef void* LV2_Handle;
typedef struct { const char* URI; void* data; } LV2_Feature;
typedef struct LV2_Descriptor {
    const char* URI;
    LV2_Handle (*instantiate) (const struct LV2_Descriptor* descriptor, double sample_rate, const char* bundle_path, const LV2_Feature* const* features);
    void (*connect_port) (LV2_Handle instance, uint32_t port, void* data_location);
    void (*activate) (LV2_Handle instance);
    void (*run) (LV2_Handle instance, uint32_t sample_count);
    void (*deactivate) (LV2_Handle instance);
    void (*cleanup) (LV2_Handle instance);
    const void* (*extension_data) (const char* uri);
} LV2_Descriptor;
#define LV2_SYMBOL_EXPORT __attribute__ ((visibility ("default")))
const LV2_Descriptor* lv2_descriptor (uint32_t index);
#ifdef __cplusplus
}
#endif
#endif
