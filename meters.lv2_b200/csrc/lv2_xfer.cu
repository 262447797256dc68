// lv2_xfer.cu — the "phasewheel" and "stereoscope" plugins (descriptors 23, 24 of the reference, src/meters.cc:769-770):
// xfer_run (src/xfer.c:180-276) forwards the raw stereo block to the GUI as one rawstereo object (two atom:Vector of
// float) while the UI is open, answers ui_on with a ui_state {samplerate} message, and -- phasewheel only -- runs the
// stereo correlation meter whose value goes to control port 6.  The correlation runs on the GPU (b200m_cor_*); the FFT
// analysis the reference's GUI performs on the forwarded audio is available GPU-side through b200m_pw_* (pw.cu).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "lv2_abi.cuh"

namespace {

using namespace b200m;

enum { SPR_CONTROL = 0, SPR_NOTIFY, SPR_INPUT0, SPR_OUTPUT0, SPR_INPUT1, SPR_OUTPUT1, SPR_PHASE, SPR_GAIN, SPR_RANGE };

struct XferPlugin {
    b200m_cor* cor = nullptr;                                  // phasewheel only (:93-95)
    float* stage = nullptr; size_t stage_cap = 0;
    AtomWriter out;
    const void* control = nullptr; void* notify = nullptr;
    float* input[2] = {nullptr, nullptr}; float* output[2] = {nullptr, nullptr}; float* p_phase = nullptr;
    double rate = 48000; bool ui_active = false, send_settings_to_ui = false, warned = false;
    LV2_URID atom_Blank = 0, atom_Object = 0, rawstereo = 0, audioleft = 0, audioright = 0, samplerate = 0, ui_on = 0, ui_off = 0, ui_state = 0;
};

LV2_Handle xfer_instantiate (const LV2_Descriptor* d, double rate, const char*, const LV2_Feature* const* features)
{
    const LV2_URID_Map* map = nullptr;
    for (int i = 0; features && features[i]; ++i) if (!strcmp (features[i]->URI, B200M_LV2_URID_MAP)) map = (const LV2_URID_Map*)features[i]->data;
    if (!map) { fprintf (stderr, "meters.lv2 error: Host does not support urid:map\n"); return nullptr; }
    const bool wheel = !strcmp (d->URI, MTR_URI "phasewheel");
    if (!wheel && strcmp (d->URI, MTR_URI "stereoscope")) return nullptr;
    XferPlugin* p = new (std::nothrow) XferPlugin;
    if (!p) return nullptr;
    if (wheel && b200m_cor_create (&p->cor, 0, 1, (int)rate, 2e3f, 0.3f)) { delete p; return nullptr; }
    p->rate = rate;
    auto M = [&] (const char* uri) { return map->map (map->handle, uri); };
    p->atom_Blank = M (B200M_LV2_ATOM "Blank"); p->atom_Object = M (B200M_LV2_ATOM "Object");
    p->out.t_sequence = M (B200M_LV2_ATOM "Sequence"); p->out.t_object = p->atom_Object; p->out.t_vector = M (B200M_LV2_ATOM "Vector");
    p->out.t_float = M (B200M_LV2_ATOM "Float"); p->out.t_int = M (B200M_LV2_ATOM "Int");
    p->rawstereo = M (MTR_URI "rawstereo"); p->audioleft = M (MTR_URI "audioleft"); p->audioright = M (MTR_URI "audioright");
    p->samplerate = M (MTR_URI "samplerate"); p->ui_on = M (MTR_URI "ui_on"); p->ui_off = M (MTR_URI "ui_off"); p->ui_state = M (MTR_URI "ui_state");
    if (p->cor && b200m_host_alloc ((void**)&p->stage, (size_t)2 * B200M_MAX_BLOCK * sizeof (float)) == 0) p->stage_cap = B200M_MAX_BLOCK;   // pinned staging for the largest cycle, allocated here so that run() never allocates (it stays lazy only as a fallback)
    return p;
}

void xfer_connect (LV2_Handle h, uint32_t port, void* data)
{
    XferPlugin* p = (XferPlugin*)h;
    switch (port) {
    case SPR_CONTROL: p->control = data; break;
    case SPR_NOTIFY: p->notify = data; break;
    case SPR_INPUT0: p->input[0] = (float*)data; break;
    case SPR_OUTPUT0: p->output[0] = (float*)data; break;
    case SPR_INPUT1: p->input[1] = (float*)data; break;
    case SPR_OUTPUT1: p->output[1] = (float*)data; break;
    case SPR_PHASE: p->p_phase = (float*)data; break;
    default: break;
    }
}

void xfer_run (LV2_Handle h, uint32_t n)
{
    XferPlugin* p = (XferPlugin*)h;
    // audio first: the reference forwards it at the end of xfer_run (src/xfer.c:262-275) and therefore DROPS it in a cycle whose
    // atom buffer is too small (:192-205); here a metering / messaging problem never costs audio
    for (int c = 0; c < 2; ++c) if (p->output[c] && p->input[c] && p->input[c] != p->output[c]) memcpy (p->output[c], p->input[c], sizeof (float) * n);
    if (!p->notify || !p->input[0] || !p->input[1]) return;
    const size_t size = (sizeof (float) * n + 64) * 2;
    const uint32_t capacity = ((const AtomHead*)p->notify)->size;
    if (capacity < size + 128) {                               // the whole cycle is skipped, as in the reference (:190-205)
        if (!p->warned) { fprintf (stderr, "meters.lv2 error: LV2 comm-buffersize is insufficient %u/%zu bytes.\n", capacity, size + 160); p->warned = true; }
        return;
    }
    p->out.begin_sequence (p->notify, capacity);
    if (p->send_settings_to_ui && p->ui_active) {
        p->send_settings_to_ui = false;
        p->out.begin_event_object (p->ui_state);
        p->out.prop_float (p->samplerate, (float)p->rate);
        p->out.end_object ();
    }
    if (p->control) {
        for (AtomEvents ev (p->control); ev.valid (); ev.next ()) {
            const AtomHead* a = ev.body ();
            if (a->type != p->atom_Blank && a->type != p->atom_Object) continue;
            AtomObject obj; obj.a = a;
            if (obj.otype () == p->ui_on) { p->ui_active = true; p->send_settings_to_ui = true; }
            else if (obj.otype () == p->ui_off) p->ui_active = false;
        }
    }
    if (p->cor && n >= 1 && n <= B200M_MAX_BLOCK) {            // stcor->process; *p_phase = stcor->read () (:248-251)
        if (n > p->stage_cap) {
            if (p->stage) b200m_host_free (p->stage);
            p->stage = nullptr; p->stage_cap = 0;
            const size_t cap = n < 1024 ? 1024 : B200M_MAX_BLOCK;
            if (b200m_host_alloc ((void**)&p->stage, 2 * cap * sizeof (float)) == 0) p->stage_cap = cap;
        }
        if (p->stage_cap) {
            memcpy (p->stage, p->input[0], n * sizeof (float)); memcpy (p->stage + p->stage_cap, p->input[1], n * sizeof (float));
            float v = 0;
            if (b200m_cor_process_host (p->cor, p->stage, p->stage_cap, n) == 0 && b200m_cor_results (p->cor, &v, nullptr) == 0 && p->p_phase) *p->p_phase = v;
        }
    }
    if (p->ui_active) {                                        // tx_rawstereo (:162-178)
        p->out.begin_event_object (p->rawstereo);
        p->out.prop_vector_f32 (p->audioleft, p->input[0], n);
        p->out.prop_vector_f32 (p->audioright, p->input[1], n);
        p->out.end_object ();
    }
}

void xfer_cleanup (LV2_Handle h)
{
    XferPlugin* p = (XferPlugin*)h;
    b200m_cor_destroy (p->cor);
    if (p->stage) b200m_host_free (p->stage);
    delete p;
}

const void* xfer_extension_data (const char*) { return nullptr; }

const LV2_Descriptor g_wheel = {MTR_URI "phasewheel", xfer_instantiate, xfer_connect, nullptr, xfer_run, nullptr, xfer_cleanup, xfer_extension_data};
const LV2_Descriptor g_scope = {MTR_URI "stereoscope", xfer_instantiate, xfer_connect, nullptr, xfer_run, nullptr, xfer_cleanup, xfer_extension_data};

}  // namespace

namespace b200m { const LV2_Descriptor* lv2_xfer_descriptor (uint32_t i) { return i == 0 ? &g_wheel : i == 1 ? &g_scope : nullptr; } }
