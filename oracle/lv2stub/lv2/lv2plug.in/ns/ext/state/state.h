/* Minimal stand-in for lv2/state (TEST INFRASTRUCTURE ONLY). */
#ifndef LV2_STATE_H
#define LV2_STATE_H
#include <stddef.h>
#include <stdint.h>
#include "../../lv2core/lv2.h"
#define LV2_STATE_URI "http://lv2plug.in/ns/ext/state"
#define LV2_STATE__interface LV2_STATE_URI "#interface"
typedef void* LV2_State_Handle;
typedef enum { LV2_STATE_IS_POD = 1, LV2_STATE_IS_PORTABLE = 1 << 1, LV2_STATE_IS_NATIVE = 1 << 2 } LV2_State_Flags;
typedef enum { LV2_STATE_SUCCESS = 0, LV2_STATE_ERR_UNKNOWN = 1, LV2_STATE_ERR_BAD_TYPE = 2, LV2_STATE_ERR_BAD_FLAGS = 3,
               LV2_STATE_ERR_NO_FEATURE = 4, LV2_STATE_ERR_NO_PROPERTY = 5, LV2_STATE_ERR_NO_SPACE = 6 } LV2_State_Status;
typedef LV2_State_Status (*LV2_State_Store_Function) (LV2_State_Handle handle, uint32_t key, const void* value, size_t size, uint32_t type, uint32_t flags);
typedef const void* (*LV2_State_Retrieve_Function) (LV2_State_Handle handle, uint32_t key, size_t* size, uint32_t* type, uint32_t* flags);
typedef struct {
    LV2_State_Status (*save) (LV2_Handle instance, LV2_State_Store_Function store, LV2_State_Handle handle, uint32_t flags, const LV2_Feature* const* features);
    LV2_State_Status (*restore) (LV2_Handle instance, LV2_State_Retrieve_Function retrieve, LV2_State_Handle handle, uint32_t flags, const LV2_Feature* const* features);
} LV2_State_Interface;
#endif
