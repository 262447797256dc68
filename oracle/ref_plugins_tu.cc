// ref_plugins_tu.cc — compiles the reference's whole plugin glue (src/meters.cc and every .c/.cc it #includes)
// UNMODIFIED, by path, against the stand-in LV2 headers in oracle/lv2stub, and drives individual plugins through
// their own LV2_Descriptor (instantiate / connect_port / run) exactly as a host would.
//
// TEST INFRASTRUCTURE ONLY (see oracle_api.h).  No audio arithmetic lives here: the hooks below only create
// instances, hand them buffers and copy internal result fields out (this TU sees LV2meter because it includes the
// reference source textually).  Used for the rows whose arithmetic sits in the glue files themselves:
// bit-meter float_stats (src/bitmeter.c:63-105), signal-distribution histogram (src/sigdistlv2.c:303-318).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "src/meters.cc"

namespace {

std::vector<std::string> g_uris;                         // urid:map feature: index + 1 = URID
LV2_URID map_uri (LV2_URID_Map_Handle, const char* uri)
{
    for (size_t i = 0; i < g_uris.size (); ++i) if (g_uris[i] == uri) return (LV2_URID)(i + 1);
    g_uris.push_back (uri);
    return (LV2_URID)g_uris.size ();
}
LV2_URID_Map g_map = {nullptr, map_uri};

struct RefPlug {
    const LV2_Descriptor* d; LV2_Handle h;
    LV2_Atom_Sequence control;                             // empty event sequence
    std::vector<uint64_t> notify;                          // 64 KiB, 8-byte aligned
};

const LV2_Descriptor* find_desc (const char* suffix)
{
    const std::string want = std::string (MTR_URI) + suffix;
    for (uint32_t i = 0;; ++i) { const LV2_Descriptor* d = lv2_descriptor (i); if (!d) return nullptr; if (want == d->URI) return d; }
}

}  // namespace

extern "C" {

void* refplug_new (const char* uri_suffix, double rate)
{
    const LV2_Descriptor* d = find_desc (uri_suffix);
    if (!d) return nullptr;
    const LV2_Feature fmap = {LV2_URID__map, &g_map};
    const LV2_Feature* feats[] = {&fmap, nullptr};
    RefPlug* p = new RefPlug;
    p->d = d; p->h = d->instantiate (d, rate, "", feats);
    if (!p->h) { delete p; return nullptr; }
    p->notify.assign (8192, 0);
    p->control.atom.size = sizeof (LV2_Atom_Sequence_Body); p->control.atom.type = map_uri (nullptr, LV2_ATOM__Sequence);
    p->control.body.unit = 0; p->control.body.pad = 0;
    d->connect_port (p->h, 0, &p->control);
    d->connect_port (p->h, 1, p->notify.data ());
    return p;
}
void refplug_free (void* vp) { RefPlug* p = (RefPlug*)vp; p->d->cleanup (p->h); delete p; }
void refplug_connect (void* vp, uint32_t port, void* data) { RefPlug* p = (RefPlug*)vp; p->d->connect_port (p->h, port, data); }
void refplug_run (void* vp, uint32_t n)
{
    RefPlug* p = (RefPlug*)vp;
    LV2_Atom_Sequence* s = (LV2_Atom_Sequence*)p->notify.data ();
    s->atom.size = (uint32_t)(p->notify.size () * 8 - sizeof (LV2_Atom));   // host convention: capacity of the output port
    s->atom.type = 0;
    p->d->run (p->h, n);
}

/* bit-meter (src/bitmeter.c): mode = 1 -> CTL_AVERAGE (cumulative), 0 -> windowed (cleared ~5x per second) */
void refbim_mode (void* vp, int average, int integrating) { LV2meter* m = (LV2meter*)((RefPlug*)vp)->h; m->bim_average = average; m->ebu_integrating = integrating; }
void refbim_snapshot (void* vp, int32_t* hist584, int32_t* cnt5, float* minmax2, int64_t* itime)
{
    LV2meter* m = (LV2meter*)((RefPlug*)vp)->h;
    memcpy (hist584, m->histS, BIM_LAST * sizeof (int32_t));
    cnt5[0] = m->bim_zero; cnt5[1] = m->bim_pos; cnt5[2] = m->bim_nan; cnt5[3] = m->bim_inf; cnt5[4] = m->bim_den;
    minmax2[0] = m->bim_min; minmax2[1] = m->bim_max; *itime = (int64_t)m->integration_time;
}
/* signal distribution histogram (src/sigdistlv2.c) */
void refsdh_integrate (void* vp, int on) { LV2meter* m = (LV2meter*)((RefPlug*)vp)->h; m->ebu_integrating = on; }
void refsdh_snapshot (void* vp, int32_t* hist361, int32_t* maxpeak2, double* avgtmpvar3, int64_t* itime)
{
    LV2meter* m = (LV2meter*)((RefPlug*)vp)->h;
    memcpy (hist361, m->histS, DIST_BIN * sizeof (int32_t));
    maxpeak2[0] = m->hist_maxS; maxpeak2[1] = m->hist_peakS;
    avgtmpvar3[0] = m->hist_avgS; avgtmpvar3[1] = m->hist_tmpS; avgtmpvar3[2] = m->hist_varS; *itime = (int64_t)m->integration_time;
}

/* EBUr128 (src/ebulv2.cc): the UI's CTL_START / dBTP-enable messages are applied directly to the instance, the audio
 * cycle then runs through the plugin's own ebur128_run; results are read back from the instance */
void refebu_ctl (void* vp, int integrate, int dbtp)
{
    LV2meter* m = (LV2meter*)((RefPlug*)vp)->h;
    if (integrate && !m->ebu_integrating) { m->ebu->integr_start (); m->ebu_integrating = true; }
    if (!integrate && m->ebu_integrating) { m->ebu->integr_pause (); m->ebu_integrating = false; }
    m->dbtp_enable = dbtp != 0;
}
void refebu_snapshot (void* vp, float* out10)
{
    LV2meter* m = (LV2meter*)((RefPlug*)vp)->h;
    out10[0] = m->ebu->loudness_M (); out10[1] = m->ebu->maxloudn_M (); out10[2] = m->ebu->loudness_S (); out10[3] = m->ebu->maxloudn_S ();
    out10[4] = m->ebu->integrated (); out10[5] = m->ebu->integ_thr (); out10[6] = m->ebu->range_min (); out10[7] = m->ebu->range_max ();
    out10[8] = m->ebu->range_thr (); out10[9] = m->tp_max;
}

/* goniometer (src/goniometer.h:113-169): offsets of the struct the reference GUI reaches through instance-access -- the same list,
 * in the same order, as b200m_lv2_gon_layout of the product (csrc/lv2_gon.cu) */
int refgon_layout (size_t* out, int n)
{
    const size_t v[] = {offsetof (LV2gm, rb), offsetof (LV2gm, ui_active), offsetof (LV2gm, rb_overrun), offsetof (LV2gm, s_sfact),
                        offsetof (LV2gm, s_linewidth), offsetof (LV2gm, input), offsetof (LV2gm, rate), offsetof (LV2gm, ntfy),
                        offsetof (LV2gm, msg_thread_lock), offsetof (LV2gm, map), sizeof (LV2gm)};
    const int m = (int)(sizeof (v) / sizeof (v[0]));
    for (int i = 0; i < n && i < m; ++i) out[i] = v[i];
    return m;
}

}  // extern "C"
