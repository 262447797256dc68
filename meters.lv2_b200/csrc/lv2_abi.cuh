// lv2_abi.cuh — the slice of the LV2 C ABI the façade plugins need (host side only, no device code).
//
// The LV2 SDK is not installed in this image, so the public, frozen struct layouts are restated here from the LV2
// specification (lv2core: LV2_Descriptor / LV2_Feature; atom: LV2_Atom and its Sequence / Event / Object / Property
// bodies, everything padded to 8 bytes; urid: LV2_URID_Map; state: LV2_State_Interface).  AtomWriter / AtomObject are
// this library's own small writer and reader for the two things the reference's plugins do with atoms: forge a
// sequence of "event at frame 0 holding an object of int/float/bool properties" into the host's notify buffer, and
// walk the objects of the control sequence (src/uris.h:279-318, src/ebulv2.cc:244-337).
#pragma once
#include <stdint.h>
#include <string.h>

extern "C" {
typedef void* LV2_Handle;
typedef struct { const char* URI; void* data; } LV2_Feature;
typedef struct LV2_Descriptor_s {
    const char* URI;
    LV2_Handle (*instantiate) (const struct LV2_Descriptor_s*, double, const char*, const LV2_Feature* const*);
    void (*connect_port) (LV2_Handle, uint32_t, void*);
    void (*activate) (LV2_Handle);
    void (*run) (LV2_Handle, uint32_t);
    void (*deactivate) (LV2_Handle);
    void (*cleanup) (LV2_Handle);
    const void* (*extension_data) (const char*);
} LV2_Descriptor;

typedef uint32_t LV2_URID;
typedef struct { void* handle; LV2_URID (*map) (void* handle, const char* uri); } LV2_URID_Map;

typedef uint32_t (*LV2_State_Store_Function) (void* handle, uint32_t key, const void* value, size_t size, uint32_t type, uint32_t flags);
typedef const void* (*LV2_State_Retrieve_Function) (void* handle, uint32_t key, size_t* size, uint32_t* type, uint32_t* flags);
typedef struct {
    uint32_t (*save) (LV2_Handle, LV2_State_Store_Function, void*, uint32_t, const LV2_Feature* const*);
    uint32_t (*restore) (LV2_Handle, LV2_State_Retrieve_Function, void*, uint32_t, const LV2_Feature* const*);
} LV2_State_Interface;
}

#define MTR_URI "http://gareus.org/oss/lv2/meters#"      /* src/uris.h:37 */
#define B200M_LV2_ATOM "http://lv2plug.in/ns/ext/atom#"
#define B200M_LV2_TIME "http://lv2plug.in/ns/ext/time#"
#define B200M_LV2_URID_MAP "http://lv2plug.in/ns/ext/urid#map"
#define B200M_LV2_STATE_INTERFACE "http://lv2plug.in/ns/ext/state#interface"

namespace b200m {

struct AtomHead { uint32_t size, type; };                      // LV2_Atom: size of the body that follows, type URID

inline uint32_t atom_pad (uint32_t n) { return (n + 7u) & ~7u; }

// Appends atoms to a host buffer.  Every open container's `size` field is kept current while children are appended
// (the reference reads notify->atom.size mid-way to ration its histogram messages, src/ebulv2.cc:433).
class AtomWriter {
public:
    LV2_URID t_sequence = 0, t_object = 0, t_int = 0, t_float = 0, t_bool = 0, t_long = 0, t_double = 0, t_vector = 0;

    void begin_sequence (void* buf, uint32_t capacity)
    {
        base_ = (uint8_t*)buf; cap_ = capacity; off_ = 0; depth_ = 0; ok_ = true;
        const uint32_t head[4] = {8u, t_sequence, 0u, 0u};     // atom {size = sizeof body, type}; body {unit = frames, pad}
        const uint32_t at = off_;
        if (put (head, sizeof (head))) open_[depth_++] = at;
    }
    // one event at frame 0 whose body is an object (id 1, as the reference forges them) of the given type
    void begin_event_object (LV2_URID otype)
    {
        const int64_t frames = 0;
        put (&frames, sizeof (frames));
        const uint32_t head[4] = {8u, t_object, 1u, otype};
        const uint32_t at = off_;
        if (put (head, sizeof (head)) && depth_ < 4) open_[depth_++] = at;
    }
    void end_object () { if (depth_ > 1) --depth_; }
    void prop_int (LV2_URID key, int32_t v) { prop (key, t_int, &v); }
    void prop_float (LV2_URID key, float v) { prop (key, t_float, &v); }
    void prop_bool (LV2_URID key, bool v) { const int32_t b = v ? 1 : 0; prop (key, t_bool, &b); }
    void prop_long (LV2_URID key, int64_t v) { prop8 (key, t_long, &v); }
    void prop_double (LV2_URID key, double v) { prop8 (key, t_double, &v); }
    // atom:Vector of n 4-byte elements (child_size 4, child_type Int or Float); the data is padded to 8 bytes like every atom body
    void prop_vector32 (LV2_URID key, LV2_URID child_type, const void* v, uint32_t n)
    {
        const uint32_t head[6] = {key, 0u, 8u + 4u * n, t_vector, 4u, child_type};
        if (!put (head, sizeof (head))) return;
        if (!put (v, 4u * (n & ~1u))) return;
        if (n & 1u) { uint32_t tail[2] = {0u, 0u}; memcpy (tail, (const uint8_t*)v + 4u * (n - 1), 4); put (tail, sizeof (tail)); }
    }
    void prop_vector_i32 (LV2_URID key, const int32_t* v, uint32_t n) { prop_vector32 (key, t_int, v, n); }
    void prop_vector_f32 (LV2_URID key, const float* v, uint32_t n) { prop_vector32 (key, t_float, v, n); }
    uint32_t sequence_size () const { return base_ ? ((const AtomHead*)base_)->size : 0; }
    bool ok () const { return ok_; }

private:
    uint8_t* base_ = nullptr; uint32_t cap_ = 0, off_ = 0; uint32_t open_[4]; int depth_ = 0; bool ok_ = true;

    bool put (const void* data, uint32_t n)                    // n is a multiple of 8 for everything written here
    {
        if (!base_ || off_ + n > cap_) { ok_ = false; return false; }
        memcpy (base_ + off_, data, n);
        off_ += n;
        for (int d = 0; d < depth_; ++d) ((AtomHead*)(base_ + open_[d]))->size += n;
        return true;
    }
    void prop8 (LV2_URID key, LV2_URID type, const void* v8)
    {
        uint32_t w[6] = {key, 0u, 8u, type, 0u, 0u};          // 8-byte body: no padding
        memcpy (&w[4], v8, 8);
        put (w, sizeof (w));
    }
    void prop (LV2_URID key, LV2_URID type, const void* v4)
    {
        uint32_t w[6] = {key, 0u, 4u, type, 0u, 0u};          // property {key, context}; atom {size 4, type}; 4-byte body + 4 pad
        memcpy (&w[4], v4, 4);
        put (w, sizeof (w));
    }
};

// Read-only view of one object atom of a control sequence.
struct AtomObject {
    const AtomHead* a = nullptr;                               // a->type is Object or Blank
    uint32_t otype () const { return a->size >= 8 ? ((const uint32_t*)(a + 1))[1] : 0u; }      // 0 = no type: a truncated object matches nothing
    // value atom of the first property with this key, or NULL
    const AtomHead* get (LV2_URID key) const
    {
        const uint8_t* body = (const uint8_t*)(a + 1);
        uint64_t off = 8;                                      // past {id, otype}; 64-bit: a hostile size field must not wrap the walk
        const uint64_t end = a->size;
        while (off + 16 <= end) {
            const uint32_t* p = (const uint32_t*)(body + off);
            const AtomHead* val = (const AtomHead*)(p + 2);
            if ((uint64_t)val->size > end - off - 16) return nullptr;      // the value runs past the object: malformed, matches nothing
            if (p[0] == key) return val;
            off += 8 + (((uint64_t)8 + val->size + 7u) & ~(uint64_t)7u);
        }
        return nullptr;
    }
};

// for (AtomEvents it (seq); it.valid (); it.next ()) { it.body () ... } over an atom:Sequence port buffer
class AtomEvents {
public:
    explicit AtomEvents (const void* seq) : seq_ ((const AtomHead*)seq), off_ (8) {}
    bool valid () const { return seq_ && off_ + 16 <= seq_->size && (uint64_t)body ()->size <= (uint64_t)seq_->size - off_ - 16; }
    const AtomHead* body () const { return (const AtomHead*)((const uint8_t*)(seq_ + 1) + off_ + 8); }
    void next () { off_ += 8 + (((uint64_t)8 + body ()->size + 7u) & ~(uint64_t)7u); }      // always advances by >= 16
private:
    const AtomHead* seq_; uint64_t off_;
};

}  // namespace b200m
