// ref_spectr_tu.cc — compiles the reference's spectrum plugin UNTOUCHED, by path.
//
// TEST INFRASTRUCTURE ONLY.  src/spectr.c and src/spectrumlv2.c are "static include" files
// of src/meters.cc (:672-681); this TU supplies what meters.cc would have supplied (LV2 type
// names, MTR_URI from src/uris.h:37, an extension_data symbol) and then drives the plugin's
// own LV2_Descriptor exactly as a host would: instantiate / connect_port / run / cleanup.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "lv2_stub.h"
#define MTR_URI "http://gareus.org/oss/lv2/meters#"
static const void* extension_data (const char*) { return 0; }

#include "src/spectr.c"
#include "src/spectrumlv2.c"

struct RefSpec { LV2_Handle h; float spd, rst, amp; };

extern "C" {
void* refspec_new (double rate, int nchan)
{
    const LV2_Descriptor* d = nchan == 2 ? &descriptorSpectrum2 : &descriptorSpectrum1;
    RefSpec* s = new RefSpec;
    s->h = d->instantiate (d, rate, "", 0);
    if (!s->h) { delete s; return 0; }
    s->spd = 1.0f; s->rst = -4.0f; s->amp = 0;   /* defaults: rst_h=-4, spd_h=1 (src/spectrumlv2.c:95-96) */
    d->connect_port (s->h, SA_SPEED, &s->spd);
    d->connect_port (s->h, SA_RESET, &s->rst);
    d->connect_port (s->h, SA_AMP, &s->amp);
    return s;
}
void refspec_free (void* p) { RefSpec* s = (RefSpec*)p; descriptorSpectrum2.cleanup (s->h); delete s; }
void refspec_run (void* p, const float* l, const float* r, uint32_t n, float speed, float reset, float* ports60)
{
    RefSpec* s = (RefSpec*)p; const LV2_Descriptor* d = &descriptorSpectrum2;
    s->spd = speed; s->rst = reset;
    for (uint32_t i = 0; i < 60; ++i) d->connect_port (s->h, i, &ports60[i]);
    /* in-place (in == out): the pass-through memcpy is skipped, src/spectrumlv2.c:251-256 */
    d->connect_port (s->h, SA_INPUT0, (void*)l);  d->connect_port (s->h, SA_OUTPUT0, (void*)l);
    d->connect_port (s->h, SA_INPUT1, (void*)r);  d->connect_port (s->h, SA_OUTPUT1, (void*)r);
    d->run (s->h, n);
}
void refspec_state (void* p, double* z, float* v, float* m)
{
    LV2spec* self = (LV2spec*)((RefSpec*)p)->h;
    for (int b = 0; b < FILTER_COUNT; ++b) {
        v[b] = self->val_f[b]; m[b] = self->max_f[b];
        for (int s = 0; s < MAXORDER; ++s) { z[(b * 6 + s) * 2] = self->flt[b].f[s].z[0]; z[(b * 6 + s) * 2 + 1] = self->flt[b].f[s].z[1]; }
    }
}
void refspec_coeffs (void* p, double* W)
{
    LV2spec* self = (LV2spec*)((RefSpec*)p)->h;
    for (int b = 0; b < FILTER_COUNT; ++b) for (int s = 0; s < MAXORDER; ++s) for (int k = 0; k < 6; ++k) W[(b * 6 + s) * 6 + k] = self->flt[b].f[s].W[k];
}
}
