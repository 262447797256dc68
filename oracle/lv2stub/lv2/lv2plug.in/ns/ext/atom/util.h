/* Minimal stand-in for lv2/atom/util (TEST INFRASTRUCTURE ONLY). */
#ifndef LV2_ATOM_UTIL_H
#define LV2_ATOM_UTIL_H
#include <stdarg.h>
#include <stdbool.h>
#include <string.h>
#include "atom.h"
static inline uint32_t lv2_atom_pad_size (uint32_t size) { return (size + 7U) & (~7U); }
static inline uint32_t lv2_atom_total_size (const LV2_Atom* atom) { return (uint32_t)sizeof (LV2_Atom) + atom->size; }
static inline LV2_Atom_Event* lv2_atom_sequence_begin (const LV2_Atom_Sequence_Body* body) { return (LV2_Atom_Event*)(body + 1); }
static inline bool lv2_atom_sequence_is_end (const LV2_Atom_Sequence_Body* body, uint32_t size, const LV2_Atom_Event* i) { return (const uint8_t*)i >= ((const uint8_t*)body + size); }
static inline LV2_Atom_Event* lv2_atom_sequence_next (const LV2_Atom_Event* i) { return (LV2_Atom_Event*)((const uint8_t*)i + sizeof (LV2_Atom_Event) + lv2_atom_pad_size (i->body.size)); }
#define LV2_ATOM_SEQUENCE_FOREACH(seq, iter) \
    for (LV2_Atom_Event* iter = lv2_atom_sequence_begin (&(seq)->body); !lv2_atom_sequence_is_end (&(seq)->body, (seq)->atom.size, (iter)); (iter) = lv2_atom_sequence_next (iter))
static inline int lv2_atom_object_get (const LV2_Atom_Object* object, ...)
{
    /* key/value lookup over the object's properties */
    int matches = 0, n_queries = 0;
    va_list args;
    va_start (args, object);
    for (n_queries = 0; va_arg (args, uint32_t); ++n_queries) { if (!va_arg (args, const LV2_Atom**)) { va_end (args); return -1; } }
    va_end (args);
    const uint8_t* p = (const uint8_t*)LV2_ATOM_CONTENTS_CONST (LV2_Atom_Object, object);
    const uint8_t* end = (const uint8_t*)object + sizeof (LV2_Atom) + object->atom.size;
    while (p < end) {
        const LV2_Atom_Property_Body* prop = (const LV2_Atom_Property_Body*)p;
        va_start (args, object);
        for (int i = 0; i < n_queries; ++i) {
            uint32_t qkey = va_arg (args, uint32_t);
            const LV2_Atom** qval = va_arg (args, const LV2_Atom**);
            if (qkey == prop->key && !*qval) { *qval = &prop->value; if (++matches == n_queries) { va_end (args); return matches; } break; }
        }
        va_end (args);
        p += lv2_atom_pad_size ((uint32_t)sizeof (LV2_Atom_Property_Body) + prop->value.size);
    }
    return matches;
}
#endif
