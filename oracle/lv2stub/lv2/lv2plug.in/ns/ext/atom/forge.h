/* Minimal stand-in for lv2/atom/forge (TEST INFRASTRUCTURE ONLY): a small working atom forge (buffer sink only). */
#ifndef LV2_ATOM_FORGE_H
#define LV2_ATOM_FORGE_H
#include <assert.h>
#include "atom.h"
#include "util.h"
#include "../urid/urid.h"
typedef intptr_t LV2_Atom_Forge_Ref;
typedef struct LV2_Atom_Forge_Frame { struct LV2_Atom_Forge_Frame* parent; LV2_Atom_Forge_Ref ref; } LV2_Atom_Forge_Frame;
typedef struct {
    uint8_t* buf; uint32_t offset; uint32_t size; LV2_Atom_Forge_Frame* stack;
    LV2_URID Blank, Bool, Chunk, Double, Float, Int, Long, Literal, Object, Path, Property, Resource, Sequence, String, Tuple, URI, URID, Vector;
} LV2_Atom_Forge;
static inline void lv2_atom_forge_set_buffer (LV2_Atom_Forge* f, uint8_t* buf, size_t size) { f->buf = buf; f->size = (uint32_t)size; f->offset = 0; f->stack = 0; }
static inline void lv2_atom_forge_init (LV2_Atom_Forge* f, LV2_URID_Map* map)
{
    lv2_atom_forge_set_buffer (f, 0, 0);
    f->Blank = map->map (map->handle, LV2_ATOM__Blank); f->Bool = map->map (map->handle, LV2_ATOM__Bool); f->Chunk = map->map (map->handle, LV2_ATOM__Chunk);
    f->Double = map->map (map->handle, LV2_ATOM__Double); f->Float = map->map (map->handle, LV2_ATOM__Float); f->Int = map->map (map->handle, LV2_ATOM__Int);
    f->Long = map->map (map->handle, LV2_ATOM__Long); f->Literal = map->map (map->handle, LV2_ATOM__Literal); f->Object = map->map (map->handle, LV2_ATOM__Object);
    f->Path = map->map (map->handle, LV2_ATOM__Path); f->Property = map->map (map->handle, LV2_ATOM__Property); f->Resource = map->map (map->handle, LV2_ATOM__Resource);
    f->Sequence = map->map (map->handle, LV2_ATOM__Sequence); f->String = map->map (map->handle, LV2_ATOM__String); f->Tuple = map->map (map->handle, LV2_ATOM__Tuple);
    f->URI = map->map (map->handle, LV2_ATOM__URI); f->URID = map->map (map->handle, LV2_ATOM__URID); f->Vector = map->map (map->handle, LV2_ATOM__Vector);
}
static inline LV2_Atom* lv2_atom_forge_deref (LV2_Atom_Forge* f, LV2_Atom_Forge_Ref ref) { (void)f; return (LV2_Atom*)ref; }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_push (LV2_Atom_Forge* f, LV2_Atom_Forge_Frame* frame, LV2_Atom_Forge_Ref ref) { frame->parent = f->stack; frame->ref = ref; f->stack = frame; return ref; }
static inline void lv2_atom_forge_pop (LV2_Atom_Forge* f, LV2_Atom_Forge_Frame* frame) { (void)frame; f->stack = frame->parent; }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_raw (LV2_Atom_Forge* f, const void* data, uint32_t size)
{
    if (f->offset + size > f->size) return 0;
    uint8_t* mem = f->buf + f->offset;
    f->offset += size;
    if (data) memcpy (mem, data, size);
    for (LV2_Atom_Forge_Frame* fr = f->stack; fr; fr = fr->parent) lv2_atom_forge_deref (f, fr->ref)->size += size;
    return (LV2_Atom_Forge_Ref)mem;
}
static inline void lv2_atom_forge_pad (LV2_Atom_Forge* f, uint32_t written) { const uint64_t pad = 0; const uint32_t n = lv2_atom_pad_size (written) - written; if (n) lv2_atom_forge_raw (f, &pad, n); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_write (LV2_Atom_Forge* f, const void* data, uint32_t size) { LV2_Atom_Forge_Ref out = lv2_atom_forge_raw (f, data, size); if (out) lv2_atom_forge_pad (f, size); return out; }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_primitive (LV2_Atom_Forge* f, const LV2_Atom* a) { return lv2_atom_forge_write (f, a, (uint32_t)sizeof (LV2_Atom) + a->size); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_int (LV2_Atom_Forge* f, int32_t v) { const LV2_Atom_Int a = {{sizeof (v), f->Int}, v}; return lv2_atom_forge_primitive (f, &a.atom); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_long (LV2_Atom_Forge* f, int64_t v) { const LV2_Atom_Long a = {{sizeof (v), f->Long}, v}; return lv2_atom_forge_primitive (f, &a.atom); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_float (LV2_Atom_Forge* f, float v) { const LV2_Atom_Float a = {{sizeof (v), f->Float}, v}; return lv2_atom_forge_primitive (f, &a.atom); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_double (LV2_Atom_Forge* f, double v) { const LV2_Atom_Double a = {{sizeof (v), f->Double}, v}; return lv2_atom_forge_primitive (f, &a.atom); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_bool (LV2_Atom_Forge* f, bool v) { const LV2_Atom_Bool a = {{sizeof (int32_t), f->Bool}, v ? 1 : 0}; return lv2_atom_forge_primitive (f, &a.atom); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_urid (LV2_Atom_Forge* f, LV2_URID v) { const LV2_Atom_URID a = {{sizeof (v), f->URID}, v}; return lv2_atom_forge_primitive (f, &a.atom); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_vector (LV2_Atom_Forge* f, uint32_t child_size, uint32_t child_type, uint32_t n_elems, const void* elems)
{
    const LV2_Atom_Vector a = {{(uint32_t)sizeof (LV2_Atom_Vector_Body) + n_elems * child_size, f->Vector}, {child_size, child_type}};
    LV2_Atom_Forge_Ref out = lv2_atom_forge_write (f, &a, sizeof (a));
    if (out) lv2_atom_forge_write (f, elems, child_size * n_elems);
    return out;
}
static inline LV2_Atom_Forge_Ref lv2_atom_forge_object (LV2_Atom_Forge* f, LV2_Atom_Forge_Frame* frame, LV2_URID id, LV2_URID otype)
{
    const LV2_Atom_Object a = {{(uint32_t)sizeof (LV2_Atom_Object_Body), f->Object}, {id, otype}};
    return lv2_atom_forge_push (f, frame, lv2_atom_forge_write (f, &a, sizeof (a)));
}
static inline LV2_Atom_Forge_Ref lv2_atom_forge_blank (LV2_Atom_Forge* f, LV2_Atom_Forge_Frame* frame, uint32_t id, LV2_URID otype)
{
    const LV2_Atom_Object a = {{(uint32_t)sizeof (LV2_Atom_Object_Body), f->Blank}, {id, otype}};
    return lv2_atom_forge_push (f, frame, lv2_atom_forge_write (f, &a, sizeof (a)));
}
static inline LV2_Atom_Forge_Ref lv2_atom_forge_resource (LV2_Atom_Forge* f, LV2_Atom_Forge_Frame* frame, LV2_URID id, LV2_URID otype) { return lv2_atom_forge_object (f, frame, id, otype); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_key (LV2_Atom_Forge* f, LV2_URID key) { const uint32_t body[2] = {key, 0}; return lv2_atom_forge_write (f, body, sizeof (body)); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_property_head (LV2_Atom_Forge* f, LV2_URID key, LV2_URID context) { const uint32_t body[2] = {key, context}; return lv2_atom_forge_write (f, body, sizeof (body)); }
static inline LV2_Atom_Forge_Ref lv2_atom_forge_sequence_head (LV2_Atom_Forge* f, LV2_Atom_Forge_Frame* frame, uint32_t unit)
{
    const LV2_Atom_Sequence a = {{(uint32_t)sizeof (LV2_Atom_Sequence_Body), f->Sequence}, {unit, 0}};
    return lv2_atom_forge_push (f, frame, lv2_atom_forge_write (f, &a, sizeof (a)));
}
static inline LV2_Atom_Forge_Ref lv2_atom_forge_frame_time (LV2_Atom_Forge* f, int64_t frames) { return lv2_atom_forge_write (f, &frames, sizeof (frames)); }
#endif
