/* Minimal stand-in for lv2/atom (TEST INFRASTRUCTURE ONLY): the atom POD layouts of the LV2 specification. */
#ifndef LV2_ATOM_H
#define LV2_ATOM_H
#include <stddef.h>
#include <stdint.h>
#define LV2_ATOM_URI "http://lv2plug.in/ns/ext/atom"
#define LV2_ATOM_PREFIX LV2_ATOM_URI "#"
#define LV2_ATOM__Blank LV2_ATOM_PREFIX "Blank"
#define LV2_ATOM__Bool LV2_ATOM_PREFIX "Bool"
#define LV2_ATOM__Chunk LV2_ATOM_PREFIX "Chunk"
#define LV2_ATOM__Double LV2_ATOM_PREFIX "Double"
#define LV2_ATOM__Float LV2_ATOM_PREFIX "Float"
#define LV2_ATOM__Int LV2_ATOM_PREFIX "Int"
#define LV2_ATOM__Long LV2_ATOM_PREFIX "Long"
#define LV2_ATOM__Literal LV2_ATOM_PREFIX "Literal"
#define LV2_ATOM__Object LV2_ATOM_PREFIX "Object"
#define LV2_ATOM__Path LV2_ATOM_PREFIX "Path"
#define LV2_ATOM__Property LV2_ATOM_PREFIX "Property"
#define LV2_ATOM__Resource LV2_ATOM_PREFIX "Resource"
#define LV2_ATOM__Sequence LV2_ATOM_PREFIX "Sequence"
#define LV2_ATOM__String LV2_ATOM_PREFIX "String"
#define LV2_ATOM__Tuple LV2_ATOM_PREFIX "Tuple"
#define LV2_ATOM__URI LV2_ATOM_PREFIX "URI"
#define LV2_ATOM__URID LV2_ATOM_PREFIX "URID"
#define LV2_ATOM__Vector LV2_ATOM_PREFIX "Vector"
#define LV2_ATOM__eventTransfer LV2_ATOM_PREFIX "eventTransfer"
#define LV2_ATOM__atomTransfer LV2_ATOM_PREFIX "atomTransfer"
#define LV2_ATOM__beatTime LV2_ATOM_PREFIX "beatTime"
#define LV2_ATOM__frameTime LV2_ATOM_PREFIX "frameTime"
#define LV2_ATOM_CONTENTS(type, atom) ((void*)((uint8_t*)(atom) + sizeof (type)))
#define LV2_ATOM_CONTENTS_CONST(type, atom) ((const void*)((const uint8_t*)(atom) + sizeof (type)))
#define LV2_ATOM_BODY(atom) LV2_ATOM_CONTENTS (LV2_Atom, atom)
#define LV2_ATOM_BODY_CONST(atom) LV2_ATOM_CONTENTS_CONST (LV2_Atom, atom)
typedef struct { uint32_t size; uint32_t type; } LV2_Atom;
typedef struct { LV2_Atom atom; int32_t body; } LV2_Atom_Int;
typedef struct { LV2_Atom atom; int64_t body; } LV2_Atom_Long;
typedef struct { LV2_Atom atom; float body; } LV2_Atom_Float;
typedef struct { LV2_Atom atom; double body; } LV2_Atom_Double;
typedef LV2_Atom_Int LV2_Atom_Bool;
typedef struct { LV2_Atom atom; uint32_t body; } LV2_Atom_URID;
typedef struct { uint32_t child_size; uint32_t child_type; } LV2_Atom_Vector_Body;
typedef struct { LV2_Atom atom; LV2_Atom_Vector_Body body; } LV2_Atom_Vector;
typedef struct { uint32_t key; uint32_t context; LV2_Atom value; } LV2_Atom_Property_Body;
typedef struct { uint32_t id; uint32_t otype; } LV2_Atom_Object_Body;
typedef struct { LV2_Atom atom; LV2_Atom_Object_Body body; } LV2_Atom_Object;
typedef struct { union { int64_t frames; double beats; } time; LV2_Atom body; } LV2_Atom_Event;
typedef struct { uint32_t unit; uint32_t pad; } LV2_Atom_Sequence_Body;
typedef struct { LV2_Atom atom; LV2_Atom_Sequence_Body body; } LV2_Atom_Sequence;
#endif
