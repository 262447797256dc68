/* lv2_stub.h — types-only stand-in for <lv2/lv2plug.in/ns/lv2core/lv2.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  The LV2 SDK is not installed in this image; the reference's
 * spectrum plugin (src/spectrumlv2.c) needs nothing from it but these four type names and
 * the LV2_Descriptor field order (the public LV2 core C ABI, restated from the LV2
 * specification — it is not part of /root/reference).
 */
#ifndef B200M_LV2_STUB_H
#define B200M_LV2_STUB_H
#include <stdint.h>
typedef void* LV2_Handle;
typedef struct { const char* URI; void* data; } LV2_Feature;
typedef struct LV2_Descriptor_ {
    const char* URI;
    LV2_Handle (*instantiate) (const struct LV2_Descriptor_*, double, const char*, const LV2_Feature* const*);
    void (*connect_port) (LV2_Handle, uint32_t, void*);
    void (*activate) (LV2_Handle);
    void (*run) (LV2_Handle, uint32_t);
    void (*deactivate) (LV2_Handle);
    void (*cleanup) (LV2_Handle);
    const void* (*extension_data) (const char*);
} LV2_Descriptor;
#endif
