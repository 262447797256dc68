// tpk_internal.cuh — pieces of the true-peak / K-meter bank (tpk.cu) shared with the DR-14 bank (dr14.cu).
#pragma once
#include "common.cuh"

namespace b200m {

// DR-14 accumulation riding on the process() kernel (dr14_run's sample loop, src/dr14.c:401-416): per channel
// rms_sum += v * v; peak_cur = MAX (peak_cur, v); when the 3 s window closes inside this block (sample index `cut`,
// host-tracked: every instance shares the window clock) the sums are handed to the scoring kernel (dr14.cu) unless the
// whole instance was silent (dr14_calc_rms_score :287-297).  rms_sum == nullptr: off.
struct TpkDr {
    float *rms_sum, *peak_cur;              // running, per channel
    float *emit_rms, *emit_peak; int* emit_valid;      // the closed window's values, per channel; valid = 0 for a silent instance
    int cut, nch;                           // window closes after sample `cut` of this block (-1: not in this block); channels per instance
    double silent_thr;                      // 1e-9 * (float) n_sample_cnt
};

// internal hooks of the true-peak / K-meter bank for dr14.cu (hidden visibility)
void tpk_set_dr (b200m_tpk* h, const TpkDr* dr);                 // DR accumulation of the following process() calls (nullptr: off)
const b200m_tpk_result* tpk_device_results (b200m_tpk* h);      // device array filled by b200m_tpk_read_device

}  // namespace b200m
