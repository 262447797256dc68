// lv2_shim.cu — per-instance LV2 façade over the batched engine: `lv2_descriptor()` with the reference's URIs,
// port indices and run() semantics, so that an LV2 host can load this library where it loaded meters.so.
//
// Covers the 28 pure control-port plugins (src/meters.cc:745-792 lists all 38 descriptors):
//   VU / BBC / EBU / DIN / NOR mono+stereo (run :298-331), BBCM6 (bbcm_run :552-589),
//   COR (cor_run :511-536), dBTPmono/stereo (dbtp_run :438-508), K12/K14/K20 mono/stereo (kmeter_run :333-418),
//   spectr30mono/stereo (spectrum_run, src/spectrumlv2.c:159-257), surround3..8 (sur_run, src/surmeter.c:115-147);
// the plugins with atom ports live in lv2_ebur128.cu (EBUr128), lv2_stats.cu (SigDistHist, bitmeter) and lv2_dr14.cu
// (dr14mono/stereo, TPnRMSmono/stereo).
// lv2_xfer.cu adds phasewheel and stereoscope (raw-audio forwarding to the GUI + correlation).  Not wrapped: goniometer,
// whose GUI reaches into the plugin's C struct through LV2 instance-access (ring buffer, mutex: src/goniometer.h) -- that
// struct layout is a private ABI between the reference's plugin and its own GUI.
// Each LV2 instance owns a bank of one instance; run() is synchronous (host buffers in, ports out), exactly the
// reference's calling convention (robtk/jackwrap.c:531-544).  LV2 core types are restated from the LV2
// specification (the SDK is not installed); the struct layout is the stable public C ABI.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "lv2_abi.cuh"

namespace b200m {
const LV2_Descriptor* lv2_ebur128_descriptor ();        // lv2_ebur128.cu
const LV2_Descriptor* lv2_sigdisthist_descriptor ();    // lv2_stats.cu
const LV2_Descriptor* lv2_bitmeter_descriptor ();
const LV2_Descriptor* lv2_dr14_descriptor (uint32_t i);   // lv2_dr14.cu: dr14mono, dr14stereo, TPnRMSmono, TPnRMSstereo
const LV2_Descriptor* lv2_xfer_descriptor (uint32_t i);   // lv2_xfer.cu: phasewheel, stereoscope
}

namespace {

enum Kind { K_COR, K_DBTP, K_KMETER, K_SPEC, K_NEEDLE, K_BBCM6, K_SUR };

struct Shim {
    Kind kind; uint32_t chn;
    b200m_cor* cor = nullptr; b200m_tpk* tpk = nullptr; b200m_spec* spec = nullptr; b200m_ppm* ppm = nullptr;
    float rlgain = 1.0f;                  // needle meters: reference-level gain (src/meters.cc:243,303-306)
    float* port[68] = {nullptr};          // raw port pointers, indexed as in the reference's enums
    float* stage = nullptr; size_t stage_cap = 0;   // pinned [chn][cap] planar staging
    float* stage2 = nullptr; size_t stage2_cap = 0; // surround meters: [8][cap] = the 4 correlation pairs
    float p_refl = -9999, peak_max[2] = {0, 0}, peak_hold = 0;   // src/meters.cc:245-251
};

// port enums: src/meters.cc:59-70 (MTR_*), src/spectrumlv2.c:35-44 (SA_*)
enum { MTR_REFLEVEL = 0, MTR_INPUT0, MTR_OUTPUT0, MTR_LEVEL0, MTR_INPUT1, MTR_OUTPUT1, MTR_LEVEL1, MTR_PEAK0, MTR_PEAK1, MTR_HOLD };
enum { SA_SPEED = 60, SA_RESET = 61, SA_AMP = 62, SA_STATE = 63, SA_INPUT0 = 64, SA_OUTPUT0 = 65, SA_INPUT1 = 66, SA_OUTPUT1 = 67 };

bool stage_in (Shim* s, const float* const* in, uint32_t n)
{
    if (n > s->stage_cap) {
        if (s->stage) b200m_host_free (s->stage);
        s->stage = nullptr; s->stage_cap = 0;
        const size_t cap = n < 1024 ? 1024 : B200M_MAX_BLOCK;
        if (b200m_host_alloc ((void**)&s->stage, (size_t)s->chn * cap * sizeof (float))) return false;
        s->stage_cap = cap;
    }
    for (uint32_t c = 0; c < s->chn; ++c) memcpy (s->stage + (size_t)c * s->stage_cap, in[c], n * sizeof (float));
    return true;
}

void pass_through (float* const* in, float* const* out, uint32_t chn, uint32_t n)
{
    for (uint32_t c = 0; c < chn; ++c) if (in[c] != out[c] && in[c] && out[c]) memcpy (out[c], in[c], sizeof (float) * n);
}

LV2_Handle shim_instantiate (const LV2_Descriptor* d, double rate, const char*, const LV2_Feature* const*)
{
    Shim* s = new (std::nothrow) Shim;
    if (!s) return nullptr;
    const char* u = d->URI + strlen (MTR_URI);
    int rc = -1;
    if (!strcmp (u, "COR")) { s->kind = K_COR; s->chn = 2; rc = b200m_cor_create (&s->cor, 0, 1, (int)rate, 2e3f, 0.3f); }          // :204-207
    else if (!strncmp (u, "dBTP", 4)) { s->kind = K_DBTP; s->chn = strstr (u, "stereo") ? 2 : 1; rc = b200m_tpk_create (&s->tpk, 0, s->chn, (float)rate, B200M_TPK_TRUEPEAK); }
    else if (u[0] == 'K') { s->kind = K_KMETER; s->chn = strstr (u, "stereo") ? 2 : 1; rc = b200m_tpk_create (&s->tpk, 0, s->chn, (float)rate, B200M_TPK_KMETER); }
    else if (!strcmp (u, "BBCM6")) { s->kind = K_BBCM6; s->chn = 2; rc = b200m_ppm_create (&s->ppm, 0, 1, (float)rate, B200M_PPM_MS); }     // :208-214
    else if (!strncmp (u, "VU", 2) || !strncmp (u, "BBC", 3) || !strncmp (u, "EBU", 3) || !strncmp (u, "DIN", 3) || !strncmp (u, "NOR", 3)) {
        // MTRDEF (src/meters.cc:172-190,215-219): VU -> Vumeterdsp, BBC/EBU -> Iec2ppmdsp, DIN/NOR -> Iec1ppmdsp
        s->kind = K_NEEDLE; s->chn = strstr (u, "stereo") ? 2 : 1;
        const int kind = !strncmp (u, "VU", 2) ? B200M_PPM_VU : (!strncmp (u, "DIN", 3) || !strncmp (u, "NOR", 3)) ? B200M_PPM_IEC1 : B200M_PPM_IEC2;
        rc = b200m_ppm_create (&s->ppm, 0, s->chn, (float)rate, kind);
    }
    else if (!strncmp (u, "surround", 8) && u[8] >= '3' && u[8] <= '8' && !u[9]) {          // src/surmeter.c:24-70
        s->kind = K_SUR; s->chn = (uint32_t)(u[8] - '0');
        rc = b200m_tpk_create (&s->tpk, 0, s->chn, (float)rate, B200M_TPK_KMETER);
        if (!rc) rc = b200m_cor_create (&s->cor, 0, 4, (int)rate, 2e3f, 0.3f);
        if (rc) { b200m_tpk_destroy (s->tpk); b200m_cor_destroy (s->cor); }
    }
    else if (!strncmp (u, "spectr30", 8)) { s->kind = K_SPEC; s->chn = strstr (u, "stereo") ? 2 : 1; rc = b200m_spec_create (&s->spec, 0, 1, s->chn, rate); }
    if (rc) { delete s; return nullptr; }                  // instantiate() -> NULL, as the reference does on failure
    return s;
}

void shim_connect (LV2_Handle h, uint32_t port, void* data)
{
    Shim* s = (Shim*)h;
    if (port < 68) s->port[port] = (float*)data;
}

void shim_cleanup (LV2_Handle h)
{
    Shim* s = (Shim*)h;
    b200m_cor_destroy (s->cor); b200m_tpk_destroy (s->tpk); b200m_spec_destroy (s->spec); b200m_ppm_destroy (s->ppm);
    if (s->stage) b200m_host_free (s->stage);
    if (s->stage2) b200m_host_free (s->stage2);
    delete s;
}

const void* shim_extension_data (const char*) { return nullptr; }

void run_cor (Shim* s, uint32_t off, uint32_t n)
{
    float* in[2] = {s->port[MTR_INPUT0] + off, s->port[MTR_INPUT1] + off};
    if (!stage_in (s, in, n)) return;
    float v = 0;
    if (b200m_cor_process_host (s->cor, s->stage, s->stage_cap, n) == 0 && b200m_cor_results (s->cor, &v, nullptr) == 0)
        *s->port[MTR_LEVEL0] = v;                          // *level[0] = cor->read() (:516-517)
}

// the "re-use port 0 to request/notify UI" handshake shared by dbtp_run (:444-463) and kmeter_run (:339-357)
bool refl_handshake (Shim* s, bool kmeter)
{
    bool reinit = false;
    const float r = *s->port[MTR_REFLEVEL];
    if (s->p_refl != r) {
        if (fabsf (r) < 3) {
            reinit = true;
            if (kmeter) s->peak_hold = 0; else { s->peak_max[0] = 0; s->peak_max[1] = 0; }
            b200m_tpk_reset (s->tpk, -1, nullptr);
        }
        if (kmeter) { if (fabsf (r) == 3) reinit = true; else s->p_refl = r; }
        else if (fabsf (r) != 3) s->p_refl = r;
    }
    if (!kmeter && fabsf (r) == 3) reinit = true;
    return reinit;
}

void run_tpk (Shim* s, uint32_t off, uint32_t n)
{
    const bool km = s->kind == K_KMETER;
    const bool reinit = refl_handshake (s, km);
    // a mono meter re-uses the second channel's port slots for its peak values: only chn audio pointers exist
    float* in[2] = {s->port[MTR_INPUT0] + off, s->chn == 2 ? s->port[MTR_INPUT1] + off : nullptr};
    if (!stage_in (s, in, n)) return;
    if (b200m_tpk_process_host (s->tpk, s->stage, s->stage_cap, n, B200M_TP_MODE_PROCESS)) return;
    if (reinit) {                                          // force parameter change (:381-389, :476-489); no read() in such a cycle
        b200m_tpk_result sync[2];
        b200m_tpk_results (s->tpk, sync, nullptr);          // stream sync only: run() must not return while the upload of `stage` is in flight
        if (km) { if (s->chn == 1) *s->port[MTR_OUTPUT1] = -1 - (rand () & 0xffff); else *s->port[MTR_HOLD] = -1 - (rand () & 0xffff); }
        else if (s->chn == 1) { *s->port[MTR_LEVEL0] = -500 - (rand () & 0xffff); *s->port[MTR_INPUT1] = -500 - (rand () & 0xffff); }
        else { for (int p : {MTR_LEVEL0, MTR_LEVEL1, MTR_PEAK0, MTR_PEAK1}) *s->port[p] = -500 - (rand () & 0xffff); }
        return;
    }
    b200m_tpk_result r[2];
    if (b200m_tpk_read_device (s->tpk, nullptr) || b200m_tpk_results (s->tpk, r, nullptr)) return;
    const float rlgain = 1.0f;                             // :243
    if (km) {                                              // :391-407
        if (s->chn == 1) {
            *s->port[MTR_LEVEL0] = rlgain * r[0].km_rms;
            *s->port[MTR_INPUT1] = rlgain * r[0].km_peak;
            if (*s->port[MTR_INPUT1] > s->peak_hold) s->peak_hold = *s->port[MTR_INPUT1];
            *s->port[MTR_OUTPUT1] = s->peak_hold;
        } else {
            for (int c = 0; c < 2; ++c) {
                *s->port[c ? MTR_LEVEL1 : MTR_LEVEL0] = rlgain * r[c].km_rms;
                float* pk = s->port[c ? MTR_PEAK1 : MTR_PEAK0];
                *pk = rlgain * r[c].km_peak;
                if (*pk > s->peak_hold) s->peak_hold = *pk;
            }
            *s->port[MTR_HOLD] = s->peak_hold;
        }
    } else {                                               // :491-507
        for (uint32_t c = 0; c < s->chn; ++c) {
            if (s->peak_max[c] < rlgain * r[c].tp_p) s->peak_max[c] = rlgain * r[c].tp_p;
            *s->port[c ? MTR_LEVEL1 : MTR_LEVEL0] = rlgain * r[c].tp_m;
        }
        if (s->chn == 1) *s->port[MTR_INPUT1] = s->peak_max[0];
        else { *s->port[MTR_PEAK0] = s->peak_max[0]; *s->port[MTR_PEAK1] = s->peak_max[1]; }
    }
}

// run() and bbcm_run() of the needle meters (src/meters.cc:298-331,552-589)
void run_needle (Shim* s, uint32_t off, uint32_t n)
{
    const float r = *s->port[MTR_REFLEVEL];
    if (s->p_refl != r) { s->p_refl = r; s->rlgain = powf (10.0f, 0.05f * (s->p_refl + 18.0)); }
    float* in[2] = {s->port[MTR_INPUT0] + off, s->chn == 2 ? s->port[MTR_INPUT1] + off : nullptr};
    if (s->kind == K_BBCM6) {
        const bool s20 = (*s->port[MTR_PEAK0] > 0.5) ? true : false;           // port 7
        b200m_ppm_set_gain (s->ppm, -6, s20 ? +14 : -6);
    }
    if (!stage_in (s, in, n)) return;
    float v[2] = {0, 0};
    if (b200m_ppm_process_host (s->ppm, s->stage, s->stage_cap, n) || b200m_ppm_read_device (s->ppm, nullptr) || b200m_ppm_results (s->ppm, v, nullptr)) return;
    *s->port[MTR_LEVEL0] = s->rlgain * v[0];
    if (s->chn == 2) *s->port[MTR_LEVEL1] = s->rlgain * v[1];
}

void run_spec (Shim* s, uint32_t off, uint32_t n)
{
    float* in[2] = {s->port[SA_INPUT0] + off, s->chn == 2 ? s->port[SA_INPUT1] + off : nullptr};
    if (!stage_in (s, in, n)) return;
    float ports[60];
    if (b200m_spec_process_host (s->spec, s->stage, s->stage_cap, n, *s->port[SA_SPEED], *s->port[SA_RESET])) return;
    if (b200m_spec_results (s->spec, ports, nullptr)) return;
    for (int i = 0; i < 30; ++i) {
        if (s->port[i]) *s->port[i] = ports[i];
        if (s->port[30 + i]) *s->port[30 + i] = ports[30 + i] <= -500.0f ? -500.0f - (rand () & 0xffff) : ports[30 + i];   // :243-246
    }
}

// sur_run (src/surmeter.c:115-147): 3 or 4 selectable-pair correlation meters + one K-meter per channel
void run_sur (Shim* s, uint32_t off, uint32_t n)
{
    float* in[8];
    for (uint32_t c = 0; c < s->chn; ++c) { in[c] = s->port[13 + 4 * c]; if (!in[c]) return; in[c] += off; }
    if (n > s->stage2_cap) {
        if (s->stage2) b200m_host_free (s->stage2);
        s->stage2 = nullptr; s->stage2_cap = 0;
        const size_t cap = n < 1024 ? 1024 : B200M_MAX_BLOCK;
        if (b200m_host_alloc ((void**)&s->stage2, 8 * cap * sizeof (float))) return;
        s->stage2_cap = cap;
    }
    const uint32_t cors = s->chn > 3 ? 4 : 3;
    for (uint32_t c = 0; c < 4; ++c) {
        float* a = s->stage2 + (size_t)(2 * c) * s->stage2_cap; float* b = a + s->stage2_cap;
        if (c < cors && s->port[1 + 3 * c] && s->port[2 + 3 * c]) {
            uint32_t in_a = (uint32_t)rintf (*s->port[1 + 3 * c]), in_b = (uint32_t)rintf (*s->port[2 + 3 * c]);
            if (in_a >= s->chn) in_a = s->chn - 1;
            if (in_b >= s->chn) in_b = s->chn - 1;
            memcpy (a, in[in_a], n * sizeof (float)); memcpy (b, in[in_b], n * sizeof (float));
        } else { memset (a, 0, n * sizeof (float)); memset (b, 0, n * sizeof (float)); }     // cor4[3] idles on a 3-channel meter
    }
    float cv[4] = {0, 0, 0, 0};
    if (b200m_cor_process_host (s->cor, s->stage2, s->stage2_cap, n) || b200m_cor_results (s->cor, cv, nullptr)) return;
    for (uint32_t c = 0; c < cors; ++c) if (s->port[3 + 3 * c]) *s->port[3 + 3 * c] = cv[c];
    if (!stage_in (s, in, n)) return;
    b200m_tpk_result r[8];
    if (b200m_tpk_process_host (s->tpk, s->stage, s->stage_cap, n, B200M_TP_MODE_PROCESS) || b200m_tpk_read_device (s->tpk, nullptr) ||
        b200m_tpk_results (s->tpk, r, nullptr)) return;
    for (uint32_t c = 0; c < s->chn; ++c) {
        if (s->port[15 + 4 * c]) *s->port[15 + 4 * c] = r[c].km_rms;         // Kmeterdsp::read (m, p): *level = m, *peak = p
        if (s->port[16 + 4 * c]) *s->port[16 + 4 * c] = r[c].km_peak;
    }
}

void shim_run (LV2_Handle h, uint32_t n)
{
    Shim* s = (Shim*)h;
    if (n == 0) return;
    // Audio is forwarded FIRST and unconditionally (every reference run() ends in the in -> out memcpy, src/meters.cc:326-330,
    // :415-417, :531-535; src/spectrumlv2.c:249-256; src/surmeter.c:143-146): a metering failure -- no memory, a CUDA error --
    // may cost a meter reading, never the audio.  run() itself never fails.
    {
        float* in[8] = {nullptr}; float* out[8] = {nullptr};
        if (s->kind == K_SUR) for (uint32_t c = 0; c < s->chn; ++c) { in[c] = s->port[13 + 4 * c]; out[c] = s->port[14 + 4 * c]; }
        else if (s->kind == K_SPEC) { in[0] = s->port[SA_INPUT0]; out[0] = s->port[SA_OUTPUT0]; if (s->chn == 2) { in[1] = s->port[SA_INPUT1]; out[1] = s->port[SA_OUTPUT1]; } }
        else { in[0] = s->port[MTR_INPUT0]; out[0] = s->port[MTR_OUTPUT0]; if (s->chn == 2) { in[1] = s->port[MTR_INPUT1]; out[1] = s->port[MTR_OUTPUT1]; } }
        pass_through (in, out, s->chn, n);
        for (uint32_t c = 0; c < s->chn; ++c) if (!in[c]) return;          // unconnected input: nothing to meter
    }
    // the engine's block limit is 8192 frames (B200M_MAX_BLOCK = the hosts' MAXPERIOD); the reference's needle / COR / K-meter /
    // spectrum plugins take any n, so longer cycles are metered in pieces of at most 8192 frames
    for (uint32_t off = 0; off < n; off += B200M_MAX_BLOCK) {
        const uint32_t k = n - off < B200M_MAX_BLOCK ? n - off : B200M_MAX_BLOCK;
        switch (s->kind) {
        case K_COR: run_cor (s, off, k); break;
        case K_DBTP: case K_KMETER: run_tpk (s, off, k); break;
        case K_SPEC: run_spec (s, off, k); break;
        case K_NEEDLE: case K_BBCM6: run_needle (s, off, k); break;
        case K_SUR: run_sur (s, off, k); break;
        }
    }
}

#define DESC(NAME) {MTR_URI NAME, shim_instantiate, shim_connect, nullptr, shim_run, nullptr, shim_cleanup, shim_extension_data}
const LV2_Descriptor g_desc[] = {
    DESC ("VUmono"), DESC ("VUstereo"), DESC ("BBCmono"), DESC ("BBCstereo"), DESC ("EBUmono"), DESC ("EBUstereo"),
    DESC ("DINmono"), DESC ("DINstereo"), DESC ("NORmono"), DESC ("NORstereo"), DESC ("BBCM6"),
    DESC ("COR"), DESC ("spectr30mono"), DESC ("dBTPmono"), DESC ("dBTPstereo"),
    DESC ("K12mono"), DESC ("K14mono"), DESC ("K20mono"), DESC ("K12stereo"), DESC ("K14stereo"), DESC ("K20stereo"),
    DESC ("spectr30stereo"),
    DESC ("surround8"), DESC ("surround7"), DESC ("surround6"), DESC ("surround5"), DESC ("surround4"), DESC ("surround3"),
};

}  // namespace

// The one symbol meters.so exports (src/meters.cc:739-792).  Hosts look plugins up by URI; the indices here are
// the covered subset in the reference's order.
extern "C" __attribute__ ((visibility ("default"))) const LV2_Descriptor* lv2_descriptor (uint32_t index)
{
    constexpr uint32_t n = sizeof (g_desc) / sizeof (g_desc[0]);
    if (index < n) return &g_desc[index];
    if (index == n) return b200m::lv2_ebur128_descriptor ();       // the atom-port plugins follow the control-port ones
    if (index == n + 1) return b200m::lv2_sigdisthist_descriptor ();
    if (index == n + 2) return b200m::lv2_bitmeter_descriptor ();
    if (index < n + 7) return b200m::lv2_dr14_descriptor (index - (n + 3));
    if (index < n + 9) return b200m::lv2_xfer_descriptor (index - (n + 7));
    return nullptr;
}
