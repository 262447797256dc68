// ebu.cu — EBU R128 loudness bank: kernels K1 (K-weighting + fragment power) and K2 (loudness,
// histograms, gating) and the b200m_ebu_* C ABI.
//
// Replaces LV2M::Ebu_r128_proc (ebumeter/ebu_r128_proc.{h,cc} of the reference) for N independent
// instances.  Nothing here is translated from the reference's control flow: the per-sample loop
// (detect_process, ebu_r128_proc.cc:302-337) becomes one thread per mono channel fed by a
// cp.async shared-memory tile pipeline; the 20 Hz bookkeeping (process/addfrags, :207-260) and
// the histogram statistics (Ebu_r128_hist, :66-150) become one warp per instance that walks the
// 751 bins with ballot/shuffle.  Arithmetic ORDER is the reference's, operation by operation
// (no FMA contraction, IEEE div/sqrt, glibc-exact log10f), because the results feed integer
// histogram bins that must be bit-exact.
#include <math.h>
#include <stdlib.h>
#include <cuda.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace b200m {

#ifndef B200M_EBU_TILE
#define B200M_EBU_TILE 64
#endif
#ifndef B200M_EBU_TMA_UNROLL
#define B200M_EBU_TMA_UNROLL 8
#endif
#ifndef B200M_EBU_STAGES
#define B200M_EBU_STAGES 3
#endif
constexpr int EBU_TILE   = B200M_EBU_TILE;   // samples per smem tile (64 or 128)
constexpr int EBU_ROWP   = EBU_TILE + 4;  // padded row pitch (floats): = 4 mod 32 -> LDS.128 conflict free
constexpr int EBU_STAGES = B200M_EBU_STAGES; // cp.async pipeline depth (2 tiles = 128 samples in flight per channel)
static_assert (EBU_TILE == 64 || EBU_TILE == 128, "tile geometry");
constexpr int EBU_WARPS  = 4;             // warps per CTA: one per SM sub-partition, each an independent 32-channel pipeline
constexpr int EBU_WARP_FLOATS = EBU_STAGES * 32 * EBU_ROWP;
constexpr int EBU_SMEM_BYTES = EBU_WARPS * EBU_WARP_FLOATS * 4;
constexpr int EBU_MAXCHUNK = 32;          // chunks (block/fragment edges) handled per K1 launch
constexpr int HIST_PITCH = 752;           // 751 bins padded to a 16-byte multiple

struct EbuCoef { float a0, a1, a2, b1, b2, c3, c4; };

struct EbuChunks {                        // bit31: chunk ends a 50 ms fragment
    int n;
    uint32_t v[EBU_MAXCHUNK];
};

// ---- K1: K-weighting recurrence + per-chunk power sums ------------------------------------
// One warp = 32 consecutive mono channels (lane = channel).  Tiles of [32 ch x 64 samples] are
// copied global->shared with cp.async (each row of the planar input is contiguous, so every
// 16-byte copy is fully coalesced), two tiles in flight behind the one being consumed; lane l
// walks row l with LDS.128, the next float4 always loaded one group ahead (the recurrence is a
// pure dependent chain: an exposed LDS latency costs as much as two samples).
// The kernel is bound by per-warp instruction issue, not HBM: a CTA therefore
// carries exactly one warp per SM sub-partition.
B200M_DEV void kw_step (float p, const EbuCoef& c, float& z1, float& z2, float& z3, float& z4, float& sj)
{
    // x = p - b1*z1 - b2*z2 + 1e-15f;  y = a0*x + a1*z1 + a2*z2 - c3*z3 - c4*z4   (:321-322)
    float x = __fsub_rn (p, __fmul_rn (c.b1, z1));
    x = __fsub_rn (x, __fmul_rn (c.b2, z2));
    x = __fadd_rn (x, 1e-15f);
    float y = __fadd_rn (__fmul_rn (c.a0, x), __fmul_rn (c.a1, z1));
    y = __fadd_rn (y, __fmul_rn (c.a2, z2));
    y = __fsub_rn (y, __fmul_rn (c.c3, z3));
    y = __fsub_rn (y, __fmul_rn (c.c4, z4));
    z2 = z1; z1 = x;
    z4 = __fadd_rn (z4, z3);
    z3 = __fadd_rn (z3, y);
    sj = __fadd_rn (sj, __fmul_rn (y, y));
}

// ---- staging policies: how a warp's [32 channels x 64 samples] tiles reach shared memory and how lane = channel reads them ----

// (A) cp.async into row-padded tiles (pitch 68 floats = 4 mod 32: conflict-free LDS.128).  Works for any alignment.
template <bool ALIGNED>
struct PaddedStage {
    static constexpr int UNROLL = 4;
    const float* in; size_t stride; float* tile; const float* src_base; int lane, k0, k_end, nfram, ntiles; bool full_warp;

    B200M_DEV void init (const float* in_, size_t stride_, float* smem_warp, int lane_, int k0_, int k_end_, int nfram_)
    {
        in = in_; stride = stride_; tile = smem_warp; lane = lane_; k0 = k0_; k_end = k_end_; nfram = nfram_;
        ntiles = (nfram + EBU_TILE - 1) / EBU_TILE;
        full_warp = k0 + 32 <= k_end;
        src_base = in + (size_t)min (k0 + lane / (EBU_TILE / 4), k_end - 1) * stride + (lane & (EBU_TILE / 4 - 1)) * 4;
    }
    B200M_DEV void issue (int t)
    {
        if (t < ntiles) {
            float* dst = tile + (t % EBU_STAGES) * (32 * EBU_ROWP);
            const int s0 = t * EBU_TILE;
            constexpr int LPR = EBU_TILE / 4;                    // lanes per row (16-byte pieces), rows per pass = 32 / LPR
            constexpr int RPP = 32 / LPR;
            if (ALIGNED) {
                const int c4 = (lane & (LPR - 1)) * 4;           // column of this lane's 16-byte piece
                const int left = (nfram - (s0 + c4)) * 4;        // bytes still inside the block
                const int nb = left >= 16 ? 16 : (left > 0 ? left : 0);
                if (full_warp) {
                    // rows RPP * i + lane / LPR: one base pointer per lane, a constant row step (all 32 channels exist)
                    const float* sp = nb ? src_base + s0 : in;
                    const size_t step = nb ? RPP * stride : 0;
                    float* d = dst + (lane / LPR) * EBU_ROWP + c4;
#pragma unroll
                    for (int i = 0; i < 32 / RPP; ++i) cp_async16 (d + i * RPP * EBU_ROWP, sp + i * step, nb);
                } else {
#pragma unroll 4
                    for (int i = 0; i < 32 / RPP; ++i) {
                        const int r = RPP * i + lane / LPR;
                        const int kr = min (k0 + r, k_end - 1);
                        const float* src = in + (size_t)kr * stride + s0 + c4;
                        cp_async16 (dst + r * EBU_ROWP + c4, nb ? src : in, nb);
                    }
                }
            } else {
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int kr = min (k0 + r, k_end - 1);
#pragma unroll
                    for (int h = 0; h < EBU_TILE / 32; ++h) {
                        const int c = lane + 32 * h;
                        const bool ok = (s0 + c) < nfram;
                        cp_async4 (dst + r * EBU_ROWP + c, ok ? in + (size_t)kr * stride + s0 + c : in, ok ? 4 : 0);
                    }
                }
            }
        }
        cp_async_commit ();
    }
    B200M_DEV void prologue () {
#pragma unroll
        for (int t = 0; t < EBU_STAGES - 1; ++t) issue (t);
    }
    B200M_DEV void acquire (int) { cp_async_wait<EBU_STAGES - 2> (); __syncwarp (); }
    // requested after tile t is consumed, not before: measured 178 vs 183 us per EBUr128 cycle (the earlier request competes
    // with the recurrence for issue slots), 64-sample tiles x 3 stages vs 128 x 2: +5 us per cycle for -1.5 us standalone
    B200M_DEV void release (int t) { __syncwarp (); issue (t + EBU_STAGES - 1); }
    B200M_DEV void drain () { cp_async_wait<0> (); }
    B200M_DEV const float* row (int t) const { return tile + (t % EBU_STAGES) * (32 * EBU_ROWP) + lane * EBU_ROWP; }
    B200M_DEV float4 ld4 (int t, int q) const { return reinterpret_cast<const float4*> (row (t))[q]; }
    B200M_DEV float4 ld4_dyn (int t, int q) const { return ld4 (t, q); }
    B200M_DEV float ld (int t, int e) const { return row (t)[e]; }
};

// (B) TMA: one elected lane asks the copy engine for the warp's tile as two [32 rows x 32 floats] boxes (128-byte rows, 128B
// swizzle) completing on an mbarrier -- no per-lane address arithmetic, no zero-fill logic (out-of-range rows / columns
// read as 0).  The 128B swizzle stores 16-byte chunk c of row r at chunk c ^ (r & 7), so the eight lanes of an LDS.128
// phase (rows r..r+7, same logical chunk) hit eight different bank groups: conflict-free without padding.
B200M_DEV uint32_t smem_u32 (const void* p) { return (uint32_t)__cvta_generic_to_shared (p); }

struct TmaStage {
    static constexpr int UNROLL = B200M_EBU_TMA_UNROLL;       // 8 float4 = one 128-byte row segment: the swizzled offsets are then compile-time
    static constexpr int STAGE_BYTES = 32 * EBU_TILE * 4;        // 8 KB: two 4 KB boxes
    const CUtensorMap* tmap; uint8_t* tile; uint32_t bar; int lane, k0, ntiles; uint32_t off[4];

    B200M_DEV void init (const CUtensorMap* tm, uint8_t* smem_warp, uint64_t* bars, int lane_, int k0_, int nfram)
    {
        tmap = tm; tile = smem_warp; bar = smem_u32 (bars); lane = lane_; k0 = k0_;
        ntiles = (nfram + EBU_TILE - 1) / EBU_TILE;
#pragma unroll
        for (int j = 0; j < 4; ++j) off[j] = (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4));
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < EBU_STAGES; ++s) asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar + 8 * s), "r"(1));
            asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp ();
    }
    B200M_DEV void issue (int t)
    {
        if (t < ntiles && lane == 0) {
            const int s = t % EBU_STAGES;
            const uint32_t dst = smem_u32 (tile + s * STAGE_BYTES), b = bar + 8 * s;
            asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");        // the slot's previous readers (generic proxy) are done
            asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(STAGE_BYTES) : "memory");
#pragma unroll
            for (int h = 0; h < EBU_TILE / 32; ++h)
                asm volatile ("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                              :: "r"(dst + 4096 * h), "l"(tmap), "r"(t * EBU_TILE + 32 * h), "r"(k0), "r"(b) : "memory");
        }
    }
    B200M_DEV void prologue () {
#pragma unroll
        for (int t = 0; t < EBU_STAGES - 1; ++t) issue (t);
    }
    B200M_DEV void acquire (int t)
    {
        const uint32_t b = bar + 8 * (t % EBU_STAGES), parity = (uint32_t)(t / EBU_STAGES) & 1u;
        uint32_t ok = 0;
        while (!ok) asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    }
    B200M_DEV void release (int t) { __syncwarp (); issue (t + EBU_STAGES - 1); }
    B200M_DEV void drain () {}
    // float4 group q (0..15) of this lane's row: box q >> 3, logical chunk q & 7 = (q & 3) | (q & 4); (j | 4) ^ m = (j ^ m) ^ 4
    B200M_DEV float4 ld4 (int t, int q) const
    {
        const uint8_t* p = tile + (t % EBU_STAGES) * STAGE_BYTES + (q >> 3) * 4096 + (off[q & 3] ^ ((uint32_t)(q & 4) << 4));
        return *reinterpret_cast<const float4*> (p);
    }
    // run-time q (slow path): the offset is computed, not looked up (off[] stays in registers)
    B200M_DEV float4 ld4_dyn (int t, int q) const
    {
        const uint8_t* p = tile + (t % EBU_STAGES) * STAGE_BYTES + (q >> 3) * 4096 + lane * 128 + (((q & 7) ^ (lane & 7)) << 4);
        return *reinterpret_cast<const float4*> (p);
    }
    B200M_DEV float ld (int t, int e) const
    {
        const uint8_t* p = tile + (t % EBU_STAGES) * STAGE_BYTES + (e >> 5) * 4096 + lane * 128 + ((((e >> 2) & 7) ^ (lane & 7)) << 4) + (e & 3) * 4;
        return *reinterpret_cast<const float*> (p);
    }
};

// The recurrence over one block for the 32 channels of a warp; `sg` supplies the tiles.
template <int NCHAN, class Stage>
B200M_DEV void kw_warp (Stage& sg, int lane, int k, bool live, int nchans, int nfram, const EbuCoef& cf, const EbuChunks& ck, float fragm_f,
                        float* __restrict__ zst, float* __restrict__ frpwr, float* __restrict__ fragpw, int n_inst)
{
    const int ntiles = (nfram + EBU_TILE - 1) / EBU_TILE;
    float z1 = zst[0 * (size_t)nchans + k], z2 = zst[1 * (size_t)nchans + k];
    float z3 = zst[2 * (size_t)nchans + k], z4 = zst[3 * (size_t)nchans + k];
    const int inst = k / NCHAN;
    float fp = frpwr[inst];
    float sj = 0.0f;
    int ci = 0, nfr = 0;
    int cend = (int)(ck.v[0] & 0x7fffffffu);           // end position (exclusive) of the current chunk
    bool cfrag = (ck.v[0] >> 31) != 0;

    // end of one detect_process() call (:324-335): state scrub, channel sum, _frpwr +=, fragment hand-over (:217-221)
    auto chunk_end = [&] () {
        z1 = scrub (z1); z2 = scrub (z2); z3 = scrub (z3); z4 = scrub (z4);
        float si;
        if (NCHAN == 1) si = __fmul_rn (2.0f, sj);
        else if (NCHAN == 2) si = __fadd_rn (sj, __shfl_xor_sync (0xffffffffu, sj, 1));   // 1.0f*sjL + 1.0f*sjR
        else {
            // si = sum_i _chan_gain[i] * sj_i in channel order, gains 1 1 1 1.41 1.41 (:29,328-329); the instance's lanes are contiguous
            const int lead = lane - lane % NCHAN;
            si = __fmul_rn (1.0f, __shfl_sync (0xffffffffu, sj, lead));
#pragma unroll
            for (int c = 1; c < NCHAN; ++c) si = __fadd_rn (si, __fmul_rn (c >= 3 ? 1.41f : 1.0f, __shfl_sync (0xffffffffu, sj, (lead + c) & 31)));
        }
        fp = __fadd_rn (fp, si);
        if (cfrag) {
            if (live && (k % NCHAN) == 0) fragpw[(size_t)nfr * n_inst + inst] = __fdiv_rn (fp, fragm_f);
            fp = 1e-30f;
            ++nfr;
        }
        sj = 0.0f;
        ++ci;
        if (ci < ck.n) { cend = (int)(ck.v[ci] & 0x7fffffffu); cfrag = (ck.v[ci] >> 31) != 0; }
        else cend = 0x7fffffff;
    };

    sg.prologue ();
    for (int t = 0; t < ntiles; ++t) {
        sg.acquire (t);
        int a = t * EBU_TILE;
        const int b = min (a + EBU_TILE, nfram);
        if (b - a == EBU_TILE && cend >= b) {
            // fast path: a whole tile inside one chunk; float4 groups with a one-group register prefetch
            float4 cur = sg.ld4 (t, 0);
#pragma unroll Stage::UNROLL                     // cp.async staging, measured: unroll 4 26.7 us/block, 2: 29.1, 8: 27.9
            for (int q = 0; q < EBU_TILE / 4; ++q) {
                const float4 nxt = sg.ld4 (t, (q + 1) & (EBU_TILE / 4 - 1));
                kw_step (cur.x, cf, z1, z2, z3, z4, sj);
                kw_step (cur.y, cf, z1, z2, z3, z4, sj);
                kw_step (cur.z, cf, z1, z2, z3, z4, sj);
                kw_step (cur.w, cf, z1, z2, z3, z4, sj);
                cur = nxt;
            }
            a = b;
            if (a == cend) chunk_end ();
        } else {
            while (a < b) {
                const int e = min (b, cend);
                int j = a;
                // scalar head up to a 4-aligned position, vector body, scalar tail
                for (; j < e && (j & 3); ++j) kw_step (sg.ld (t, j - t * EBU_TILE), cf, z1, z2, z3, z4, sj);
                for (; j + 4 <= e; j += 4) {
                    const float4 v = sg.ld4_dyn (t, (j - t * EBU_TILE) >> 2);
                    kw_step (v.x, cf, z1, z2, z3, z4, sj);
                    kw_step (v.y, cf, z1, z2, z3, z4, sj);
                    kw_step (v.z, cf, z1, z2, z3, z4, sj);
                    kw_step (v.w, cf, z1, z2, z3, z4, sj);
                }
                for (; j < e; ++j) kw_step (sg.ld (t, j - t * EBU_TILE), cf, z1, z2, z3, z4, sj);
                a = e;
                if (a == cend) chunk_end ();
            }
        }
        sg.release (t);
    }
    sg.drain ();
    if (live) {
        zst[0 * (size_t)nchans + k] = z1; zst[1 * (size_t)nchans + k] = z2;
        zst[2 * (size_t)nchans + k] = z3; zst[3 * (size_t)nchans + k] = z4;
        if ((k % NCHAN) == 0) frpwr[inst] = fp;
    }
}

template <int NCHAN, bool ALIGNED>
__global__ void __launch_bounds__ (EBU_WARPS * 32)
ebu_kweight_frag (const float* __restrict__ in, size_t stride, int nchans, int k_first, int k_end, int nfram, EbuCoef cf, EbuChunks ck,
                  float fragm_f, float* __restrict__ zst, float* __restrict__ frpwr, float* __restrict__ fragpw, int n_inst, int pdl_trigger)
{
    // channels [k_first, k_end) of the bank's nchans (a slice: *_run_host overlaps the copy of slice s+1 with slice s)
    extern __shared__ __align__ (16) float ebu_smem[];
    // programmatic dependent launch: a kernel launched behind this one with the programmatic-serialization attribute (the
    // true-peak kernel of the EBUr128 cycle, r128.cu) may start as soon as every CTA of this grid is running
    if (pdl_trigger) asm volatile ("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int CPW = (32 / NCHAN) * NCHAN;          // channels per warp: whole instances only (30 lanes for 3- and 5-channel banks)
    const int k0 = k_first + (blockIdx.x * EBU_WARPS + warp) * CPW;
    if (k0 >= k_end) return;                           // warp-uniform; warps never synchronise with each other
    PaddedStage<ALIGNED> sg;
    sg.init (in, stride, ebu_smem + warp * EBU_WARP_FLOATS, lane, k0, k_end, nfram);
    kw_warp<NCHAN> (sg, lane, min (k0 + lane, k_end - 1) /* tail lanes shadow the last channel (no stores) */, lane < CPW && (k0 + lane) < k_end,
                    nchans, nfram, cf, ck, fragm_f, zst, frpwr, fragpw, n_inst);
}

// ---- K1 split over two warps per 32 channels -------------------------------------------------------------------------------
// The recurrence is two biquad-like stages in series (:321-322): stage 1  x = p - b1 z1 - b2 z2 + 1e-15  feeds stage 2
// y = a0 x + a1 z1 + a2 z2 - c3 z3 - c4 z4;  z4 += z3;  z3 += y;  sj += y y.  Each stage is a 16-cycle dependent chain per sample, and
// one warp per SM sub-partition (all a 16384-channel bank offers) cannot hide either behind the other: 36 cycles per sample measured.
// Here warp A runs stage 1 and hands the x stream to warp B (same sub-partition: warps w and w + 4 of the CTA) through a
// double-buffered shared-memory tile; the two chains then interleave on one scheduler.  Every channel sees exactly the same
// operations in the same order as in kw_warp, so the results stay bit-identical.  Named barriers (bar.arrive / bar.sync on 64
// threads) hand the tiles over: a waiting warp is parked by the hardware and takes no issue slots from its partner.
// MEASURED (round 2, profiles/r2_ncu_ebu_kweight_split.txt): 27.2 us under ncu against 22.8 us for the one-warp kernel, 28.6 vs 26.2 us
// live -- slower, so it is opt-in (B200M_EBU_SPLIT=1) and kept for the record.  The premise was wrong: ncu's stall breakdown of
// the one-warp kernel shows 21.5 issue cycles + 6.5 dependency-wait cycles + 8 other per sample, i.e. the warp is limited by the
// ~21.5 instructions it must ISSUE per sample on its scheduler, not by the 16-cycle chains; a second warp on the SAME scheduler
// adds hand-over instructions and barrier waits (0.54 cycles per instruction) without adding issue slots, and every scheduler of
// the 128 SMs in use already hosts a warp.  Only more channels per scheduler help (0.65 of HBM at 32768 instances).
constexpr int EBU_SPLIT_PAIRS = 4;
constexpr int EBU_SPLIT_PAIR_FLOATS = (EBU_STAGES + 2) * 32 * EBU_ROWP;
constexpr int EBU_SPLIT_SMEM = EBU_SPLIT_PAIRS * EBU_SPLIT_PAIR_FLOATS * 4;

B200M_DEV void bar_sync64 (int id) { asm volatile ("bar.sync %0, 64;" :: "r"(id) : "memory"); }
B200M_DEV void bar_arrive64 (int id) { asm volatile ("bar.arrive %0, 64;" :: "r"(id) : "memory"); }

B200M_DEV float kw_stage1 (float p, const EbuCoef& c, float& z1, float& z2)
{
    float x = __fsub_rn (p, __fmul_rn (c.b1, z1));
    x = __fsub_rn (x, __fmul_rn (c.b2, z2));
    x = __fadd_rn (x, 1e-15f);
    z2 = z1; z1 = x;
    return x;
}
B200M_DEV void kw_stage2 (float x, const EbuCoef& c, float& z1, float& z2, float& z3, float& z4, float& sj)
{
    float y = __fadd_rn (__fmul_rn (c.a0, x), __fmul_rn (c.a1, z1));       // z1, z2: the two x values before this one
    y = __fadd_rn (y, __fmul_rn (c.a2, z2));
    y = __fsub_rn (y, __fmul_rn (c.c3, z3));
    y = __fsub_rn (y, __fmul_rn (c.c4, z4));
    z2 = z1; z1 = x;
    z4 = __fadd_rn (z4, z3);
    z3 = __fadd_rn (z3, y);
    sj = __fadd_rn (sj, __fmul_rn (y, y));
}

template <int NCHAN, bool ALIGNED>
__global__ void __launch_bounds__ (2 * EBU_SPLIT_PAIRS * 32)
ebu_kweight_split (const float* __restrict__ in, size_t stride, int nchans, int k_first, int k_end, int nfram, EbuCoef cf, EbuChunks ck,
                   float fragm_f, float* __restrict__ zst, float* __restrict__ frpwr, float* __restrict__ fragpw, int n_inst, int pdl_trigger)
{
    extern __shared__ __align__ (16) float ebu_smem[];
    if (pdl_trigger) asm volatile ("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp & (EBU_SPLIT_PAIRS - 1), role = warp / EBU_SPLIT_PAIRS;      // warps w and w + 4 share a sub-partition
    const int k0 = k_first + (blockIdx.x * EBU_SPLIT_PAIRS + pair) * 32;
    if (k0 >= k_end) return;                           // both warps of the pair leave together
    float* pair_smem = ebu_smem + pair * EBU_SPLIT_PAIR_FLOATS;
    float* xt = pair_smem + EBU_STAGES * 32 * EBU_ROWP;   // two x tiles [32][EBU_ROWP]
    const int id_full = pair * 4, id_empty = pair * 4 + 2;
    const int k = min (k0 + lane, k_end - 1);
    const bool live = (k0 + lane) < k_end;
    const int ntiles = (nfram + EBU_TILE - 1) / EBU_TILE;
    float z1 = zst[0 * (size_t)nchans + k], z2 = zst[1 * (size_t)nchans + k];
    int ci = 0;
    int cend = (int)(ck.v[0] & 0x7fffffffu);

    if (role == 0) {
        // ---- warp A: stage 1, input tiles by cp.async, x tiles out
        PaddedStage<ALIGNED> sg;
        sg.init (in, stride, pair_smem, lane, k0, k_end, nfram);
        sg.prologue ();
        for (int t = 0; t < ntiles; ++t) {
            sg.acquire (t);
            if (t >= 2) bar_sync64 (id_empty + (t & 1));              // warp B is done with the x tile written two tiles ago
            float* xrow = xt + (t & 1) * (32 * EBU_ROWP) + lane * EBU_ROWP;
            int a = t * EBU_TILE;
            const int b = min (a + EBU_TILE, nfram);
            if (b - a == EBU_TILE && cend >= b) {
                float4 cur = sg.ld4 (t, 0);
#pragma unroll 4
                for (int q = 0; q < EBU_TILE / 4; ++q) {
                    const float4 nxt = sg.ld4 (t, (q + 1) & (EBU_TILE / 4 - 1));
                    float4 o;
                    o.x = kw_stage1 (cur.x, cf, z1, z2); o.y = kw_stage1 (cur.y, cf, z1, z2);
                    o.z = kw_stage1 (cur.z, cf, z1, z2); o.w = kw_stage1 (cur.w, cf, z1, z2);
                    reinterpret_cast<float4*> (xrow)[q] = o;
                    cur = nxt;
                }
                a = b;
                if (a == cend) { z1 = scrub (z1); z2 = scrub (z2); ++ci; cend = ci < ck.n ? (int)(ck.v[ci] & 0x7fffffffu) : 0x7fffffff; }
            } else {
                while (a < b) {
                    const int e = min (b, cend);
                    for (int j = a; j < e; ++j) xrow[j - t * EBU_TILE] = kw_stage1 (sg.ld (t, j - t * EBU_TILE), cf, z1, z2);
                    a = e;
                    if (a == cend) { z1 = scrub (z1); z2 = scrub (z2); ++ci; cend = ci < ck.n ? (int)(ck.v[ci] & 0x7fffffffu) : 0x7fffffff; }
                }
            }
            bar_arrive64 (id_full + (t & 1));                        // x tile t is complete
            sg.release (t);
        }
        sg.drain ();
        if (live) { zst[0 * (size_t)nchans + k] = z1; zst[1 * (size_t)nchans + k] = z2; }
    } else {
        // ---- warp B: stage 2, power sums, fragment hand-over (the chunk_end of kw_warp)
        float z3 = zst[2 * (size_t)nchans + k], z4 = zst[3 * (size_t)nchans + k];
        const int inst = k / NCHAN;
        float fp = frpwr[inst];
        float sj = 0.0f;
        int nfr = 0;
        bool cfrag = (ck.v[0] >> 31) != 0;
        auto chunk_end = [&] () {
            z1 = scrub (z1); z2 = scrub (z2); z3 = scrub (z3); z4 = scrub (z4);
            float si;
            if (NCHAN == 1) si = __fmul_rn (2.0f, sj);
            else si = __fadd_rn (sj, __shfl_xor_sync (0xffffffffu, sj, 1));
            fp = __fadd_rn (fp, si);
            if (cfrag) {
                if (live && (k % NCHAN) == 0) fragpw[(size_t)nfr * n_inst + inst] = __fdiv_rn (fp, fragm_f);
                fp = 1e-30f;
                ++nfr;
            }
            sj = 0.0f;
            ++ci;
            if (ci < ck.n) { cend = (int)(ck.v[ci] & 0x7fffffffu); cfrag = (ck.v[ci] >> 31) != 0; }
            else cend = 0x7fffffff;
        };
        for (int t = 0; t < ntiles; ++t) {
            bar_sync64 (id_full + (t & 1));
            const float* xrow = xt + (t & 1) * (32 * EBU_ROWP) + lane * EBU_ROWP;
            int a = t * EBU_TILE;
            const int b = min (a + EBU_TILE, nfram);
            if (b - a == EBU_TILE && cend >= b) {
                float4 cur = reinterpret_cast<const float4*> (xrow)[0];
#pragma unroll 4
                for (int q = 0; q < EBU_TILE / 4; ++q) {
                    const float4 nxt = reinterpret_cast<const float4*> (xrow)[(q + 1) & (EBU_TILE / 4 - 1)];
                    kw_stage2 (cur.x, cf, z1, z2, z3, z4, sj); kw_stage2 (cur.y, cf, z1, z2, z3, z4, sj);
                    kw_stage2 (cur.z, cf, z1, z2, z3, z4, sj); kw_stage2 (cur.w, cf, z1, z2, z3, z4, sj);
                    cur = nxt;
                }
                a = b;
                if (a == cend) chunk_end ();
            } else {
                while (a < b) {
                    const int e = min (b, cend);
                    for (int j = a; j < e; ++j) kw_stage2 (xrow[j - t * EBU_TILE], cf, z1, z2, z3, z4, sj);
                    a = e;
                    if (a == cend) chunk_end ();
                }
            }
            if (t + 2 < ntiles) bar_arrive64 (id_empty + (t & 1));       // x tile t may be overwritten (by tile t + 2)
        }
        if (live) {
            zst[2 * (size_t)nchans + k] = z3; zst[3 * (size_t)nchans + k] = z4;
            if ((k % NCHAN) == 0) frpwr[inst] = fp;
        }
    }
}

// the same kernel fed by TMA (16-byte aligned input with a 16-byte multiple row pitch: every bank-sized call in practice)
constexpr int EBU_TMA_SMEM = EBU_WARPS * EBU_STAGES * TmaStage::STAGE_BYTES + EBU_WARPS * EBU_STAGES * 8 + 1024;
template <int NCHAN>
__global__ void __launch_bounds__ (EBU_WARPS * 32)
ebu_kweight_tma (const __grid_constant__ CUtensorMap tmap, int nchans, int k_first, int k_end, int nfram, EbuCoef cf, EbuChunks ck,
                 float fragm_f, float* __restrict__ zst, float* __restrict__ frpwr, float* __restrict__ fragpw, int n_inst, int pdl_trigger)
{
    extern __shared__ uint8_t ebu_smem_raw[];
    if (pdl_trigger) asm volatile ("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = k_first + (blockIdx.x * EBU_WARPS + warp) * 32;
    if (k0 >= k_end) return;
    uint8_t* base = ebu_smem_raw + ((1024u - (smem_u32 (ebu_smem_raw) & 1023u)) & 1023u);     // 128B-swizzled boxes want 1 KB alignment (pointer arithmetic
                                                                                          // on the __shared__ symbol keeps the accesses LDS, not generic LD)
    uint64_t* bars = (uint64_t*)(base + EBU_WARPS * EBU_STAGES * TmaStage::STAGE_BYTES) + warp * EBU_STAGES;
    TmaStage sg;
    sg.init (&tmap, base + warp * EBU_STAGES * TmaStage::STAGE_BYTES, bars, lane, k0, nfram);
    // rows >= k_end read as zeros (or as the neighbouring slice's channels): those lanes never store
    kw_warp<NCHAN> (sg, lane, min (k0 + lane, k_end - 1), (k0 + lane) < k_end, nchans, nfram, cf, ck, fragm_f, zst, frpwr, fragpw, n_inst);
}

// ---- K2: per-fragment loudness, histograms, gated integration ------------------------------
struct EbuCtl { int div1, div2, integr, calc; };

// Ebu_r128_hist::integrate (:82-102).  `c[t]` holds bin t*32+lane.  Only non-zero bins change the
// running float sum, so the warp walks them in bin order (ballot) and applies the "/= 10 after
// every bin = 99 mod 100" steps in between: identical rounding sequence, ~#non-zero-bins steps.
B200M_DEV float hist_integrate (const int (&c)[24], int i0, const float* bp, int lane)
{
    float s = 0.0f; int n = 0;
    int next_div = i0 - (i0 % 100) + 99;
#pragma unroll
    for (int t = 0; t < 24; ++t) {
        const int bin_l = t * 32 + lane;
        unsigned m = __ballot_sync (0xffffffffu, c[t] != 0 && bin_l >= i0 && bin_l <= 750);
        while (m) {
            const int l = __ffs (m) - 1; m &= m - 1;
            const int kk = __shfl_sync (0xffffffffu, c[t], l);
            const int bin = t * 32 + l;
            while (next_div < bin) { s = __fdiv_rn (s, 10.0f); next_div += 100; }
            s = __fadd_rn (s, __fmul_rn ((float)kk, bp[bin % 100]));
            n += kk;
        }
    }
    while (next_div <= 750) { s = __fdiv_rn (s, 10.0f); next_div += 100; }
    return __fdiv_rn (s, (float)n);
}

B200M_DEV void hist_load (const int* row, int (&c)[24], int lane)
{
#pragma unroll
    for (int t = 0; t < 24; ++t) { const int b = t * 32 + lane; c[t] = (b <= 750) ? row[b] : 0; }
}

// Ebu_r128_hist::calc_integ (:105-125)
B200M_DEV void hist_calc_integ (const int* row, int count, const float* bp, int lane, float& vi, float& th)
{
    if (count < 50) { vi = -200.0f; return; }
    int c[24]; hist_load (row, c, lane);
    float s = hist_integrate (c, 0, bp, lane);
    const float lg = log10f_glibc (s);
    th = __fsub_rn (__fmul_rn (10.0f, lg), 10.0f);
    int k = (int)floorf (__fadd_rn (__fmul_rn (100.0f, lg), 0.5f)) + 600;
    if (k < 0) k = 0;
    s = hist_integrate (c, k, bp, lane);
    vi = __fmul_rn (10.0f, log10f_glibc (s));
}

// Ebu_r128_hist::calc_range (:128-150)
B200M_DEV void hist_calc_range (const int* row, int count, const float* bp, int lane, float& v0, float& v1, float& th)
{
    if (count < 20) { v0 = -200.0f; v1 = -200.0f; return; }
    int c[24]; hist_load (row, c, lane);
    float s = hist_integrate (c, 0, bp, lane);
    const float lg = log10f_glibc (s);
    th = __fsub_rn (__fmul_rn (10.0f, lg), 20.0f);
    // floorf (100 * log10f (s) + 0.5): the 0.5 literal is a double in the reference (:141)
    int k = (int)floorf ((float)((double)__fmul_rn (100.0f, lg) + 0.5)) + 500;
    if (k < 0) k = 0;
    int n = 0;
#pragma unroll
    for (int t = 0; t < 24; ++t) { const int b = t * 32 + lane; if (b >= k && b <= 750) n += c[t]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync (0xffffffffu, n, o);
    const float a = __fmul_rn (0.10f, (float)n), b95 = __fmul_rn (0.95f, (float)n);
    // for (i = k, s = 0; s < a; i++) s += histc[i];
    int i = k; s = 0.0f; bool done = !(s < a);
#pragma unroll
    for (int t = 0; t < 24; ++t) {
        const int bin_l = t * 32 + lane;
        unsigned m = done ? 0u : __ballot_sync (0xffffffffu, c[t] != 0 && bin_l >= k && bin_l <= 750);
        while (m && !done) {
            const int l = __ffs (m) - 1; m &= m - 1;
            s = __fadd_rn (s, (float)__shfl_sync (0xffffffffu, c[t], l));
            if (!(s < a)) { i = t * 32 + l + 1; done = true; }
        }
    }
    // for (j = 750, s = n; s > b; j--) s -= histc[j];
    int j = 750; s = (float)n; done = !(s > b95);
#pragma unroll
    for (int t = 23; t >= 0; --t) {
        const int bin_l = t * 32 + lane;
        unsigned m = done ? 0u : __ballot_sync (0xffffffffu, c[t] != 0 && bin_l <= 750);
        while (m && !done) {
            const int l = 31 - __clz (m); m &= ~(1u << l);
            s = __fsub_rn (s, (float)__shfl_sync (0xffffffffu, c[t], l));
            if (!(s > b95)) { j = t * 32 + l - 1; done = true; }
        }
    }
    v0 = __fdiv_rn ((float)(i - 701), 10.0f);
    v1 = __fdiv_rn ((float)(j - 699), 10.0f);
}

// K2a: one THREAD per instance, one completed 50 ms fragment (process :217-243 minus the gated statistics).
// ring layout [64][n_inst] so that lane = instance accesses coalesce; each thread parks its 64-slot ring column
// in shared memory (column private to the thread: no barrier needed) for the two ordered sums.
constexpr int K2A_THREADS = 128;

__global__ void __launch_bounds__ (K2A_THREADS)
ebu_fragment_kernel (int n_inst, int frag, int wrind, const float* __restrict__ fragpw, float* __restrict__ ring,
                     EbuCtl* __restrict__ ctl, b200m_ebu_result* __restrict__ res, int* __restrict__ histM,
                     int* __restrict__ histS, int* __restrict__ cnt)
{
    __shared__ float sring[64][K2A_THREADS];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * K2A_THREADS + tid;
    if (i >= n_inst) return;
    // all 64 ring slots in flight at once (one round of memory latency, not 64): cp.async straight into the column
#pragma unroll
    for (int w = 0; w < 64; ++w) cp_async4 (&sring[w][tid], ring + (size_t)w * n_inst + i, 4);
    cp_async_commit ();
    EbuCtl c = ctl[i];
    b200m_ebu_result r = res[i];
    const float p = fragpw[(size_t)frag * n_inst + i];
    cp_async_wait<0> ();
    sring[wrind][tid] = p;                                 // _power[_wrind++] = _frpwr / _fragm (:218)
    ring[(size_t)wrind * n_inst + i] = p;
    const int wr = (wrind + 1) & 63;
    r.frag_power = p;
    // addfrags (8), addfrags (60) (:251-260): sequential sums, oldest fragment first
    float s8 = 0.0f, s60 = 0.0f;
    { const int k = (wr - 8) & 63;
#pragma unroll
      for (int q = 0; q < 8; ++q)  s8  = __fadd_rn (s8,  sring[(q + k) & 63][tid]); }
    { const int k = (wr - 60) & 63;
#pragma unroll 10
      for (int q = 0; q < 60; ++q) s60 = __fadd_rn (s60, sring[(q + k) & 63][tid]); }
    float lm = __fadd_rn (-0.6976f, __fmul_rn (10.0f, log10f_glibc (__fdiv_rn (s8, 8.0f))));
    float ls = __fadd_rn (-0.6976f, __fmul_rn (10.0f, log10f_glibc (__fdiv_rn (s60, 60.0f))));
    if (!finitef_ (lm) || lm < -200.0f) lm = -200.0f;      // :224-225
    if (!finitef_ (ls) || ls < -200.0f) ls = -200.0f;
    r.loudness_M = lm; r.loudness_S = ls;
    if (lm > r.maxloudn_M) r.maxloudn_M = lm;
    if (ls > r.maxloudn_S) r.maxloudn_S = ls;
    if (c.integr) {                                         // :228-242
        if (++c.div1 == 2) {                                // Ebu_r128_hist::addpoint (:66-79)
            c.div1 = 0;
            int k = (int)floorf (__fadd_rn (__fmul_rn (10.0f, lm), 700.5f));
            if (k >= 0) {
                if (k > 750) { k = 750; cnt[(size_t)i * 4 + 2]++; }
                histM[(size_t)i * HIST_PITCH + k]++;
                r.hist_M_count = ++cnt[(size_t)i * 4 + 0];
            }
        }
        if (++c.div2 == 10) {
            c.div2 = 0;
            int k = (int)floorf (__fadd_rn (__fmul_rn (10.0f, ls), 700.5f));
            if (k >= 0) {
                if (k > 750) { k = 750; cnt[(size_t)i * 4 + 3]++; }
                histS[(size_t)i * HIST_PITCH + k]++;
                r.hist_S_count = ++cnt[(size_t)i * 4 + 1];
            }
            c.calc = 1;                                     // calc_integ + calc_range follow in K2b
        }
    }
    ctl[i] = c; res[i] = r;
}

// K2b: one WARP per instance, only launched when the host's phase book-keeping says that some instance
// completed its 10th fragment: calc_integ on hist_M, calc_range on hist_S (:240-241).
constexpr int K2B_WARPS = 4;

__global__ void __launch_bounds__ (K2B_WARPS * 32)
ebu_gate_kernel (int n_inst, EbuCtl* __restrict__ ctl, b200m_ebu_result* __restrict__ res, const int* histM,
                 const int* histS, const int* __restrict__ cnt, const float* __restrict__ bin_power)
{
    __shared__ float sbp[100];
    for (int q = threadIdx.x; q < 100; q += blockDim.x) sbp[q] = bin_power[q];
    __syncthreads ();
    const int lane = threadIdx.x & 31;
    const int inst = blockIdx.x * K2B_WARPS + (threadIdx.x >> 5);
    if (inst >= n_inst) return;
    if (!ctl[inst].calc) return;
    float vi = res[inst].integrated, th = res[inst].integ_thr;
    float v0 = res[inst].range_min, v1 = res[inst].range_max, rt = res[inst].range_thr;
    hist_calc_integ (histM + (size_t)inst * HIST_PITCH, cnt[(size_t)inst * 4 + 0], sbp, lane, vi, th);
    hist_calc_range (histS + (size_t)inst * HIST_PITCH, cnt[(size_t)inst * 4 + 1], sbp, lane, v0, v1, rt);
    if (lane == 0) {
        res[inst].integrated = vi; res[inst].integ_thr = th;
        res[inst].range_min = v0; res[inst].range_max = v1; res[inst].range_thr = rt;
        ctl[inst].calc = 0;
    }
}

// reset / integration control, one thread per instance
__global__ void ebu_ctl_kernel (int n_inst, int inst_sel, int cmd, int nchan, float* zst, float* frpwr, float* ring,
                                EbuCtl* ctl, b200m_ebu_result* res, int* histM, int* histS, int* cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst || (inst_sel >= 0 && i != inst_sel)) return;
    if (cmd == 0) { ctl[i].integr = 0; return; }            // integr_pause
    if (cmd == 1) { ctl[i].integr = 1; return; }            // integr_start
    // cmd 2: integr_reset (:193-204); cmd 3: reset (:176-190) = integr off + filter/ring clear + integr_reset
    for (int b = 0; b < HIST_PITCH; ++b) { histM[(size_t)i * HIST_PITCH + b] = 0; histS[(size_t)i * HIST_PITCH + b] = 0; }
    for (int b = 0; b < 4; ++b) cnt[(size_t)i * 4 + b] = 0;
    b200m_ebu_result r = res[i];
    r.maxloudn_M = r.maxloudn_S = r.integrated = r.integ_thr = -200.0f;
    r.range_min = r.range_max = r.range_thr = -200.0f;
    r.hist_M_count = r.hist_S_count = 0;
    ctl[i].div1 = ctl[i].div2 = 0; ctl[i].calc = 0;
    if (cmd == 3) {
        ctl[i].integr = 0;
        frpwr[i] = 1e-30f;
        r.loudness_M = r.loudness_S = -200.0f; r.frag_power = 0.0f;
        for (int b = 0; b < 64; ++b) ring[(size_t)b * n_inst + i] = 0.0f;
        const size_t nch = (size_t)n_inst * nchan;
        for (int c = 0; c < nchan; ++c) for (int z = 0; z < 4; ++z) zst[z * nch + (size_t)i * nchan + c] = 0.0f;
    }
    res[i] = r;
}

// whole-mix histogram sum: grid-stride atomics into one int32[B200M_MIX_WORDS] vector
__global__ void ebu_mix_reduce_kernel (int n_inst, const int* __restrict__ histM, const int* __restrict__ histS,
                                       const int* __restrict__ cnt, int* __restrict__ out)
{
    // blockIdx.x: bin column group (coalescing is across bins); blockIdx.y: slice of the instances.  Integer partial sums are
    // merged with atomicAdd into the zeroed output: order-independent, hence exact.
    const int col = blockIdx.x * blockDim.x + threadIdx.x;    // 0..1507
    if (col >= B200M_MIX_WORDS) return;
    const int per = (n_inst + gridDim.y - 1) / gridDim.y, i0 = blockIdx.y * per, i1 = min (n_inst, i0 + per);
    const int* src; size_t pitch; int c;
    if (col < 752) { src = histM; pitch = HIST_PITCH; c = col; }
    else if (col < 1504) { src = histS; pitch = HIST_PITCH; c = col - 752; }
    else { src = cnt; pitch = 4; c = col - 1504; }
    int acc = 0;
#pragma unroll 8
    for (int i = i0; i < i1; ++i) acc += src[(size_t)i * pitch + c];
    if (acc) atomicAdd (out + col, acc);
}

__global__ void ebu_mix_finish_kernel (const int* __restrict__ mix, const float* __restrict__ bin_power, float* out5)
{
    __shared__ float sbp[100];
    for (int i = threadIdx.x; i < 100; i += blockDim.x) sbp[i] = bin_power[i];
    __syncthreads ();
    const int lane = threadIdx.x;
    float vi = -200.0f, th = -200.0f, v0 = -200.0f, v1 = -200.0f, rt = -200.0f;
    hist_calc_integ (mix, mix[1504], sbp, lane, vi, th);
    hist_calc_range (mix + 752, mix[1505], sbp, lane, v0, v1, rt);
    if (lane == 0) { out5[0] = vi; out5[1] = th; out5[2] = v0; out5[3] = v1; out5[4] = rt; }
}

}  // namespace b200m

using namespace b200m;

// ---------------------------------------------------------------------------- host side
// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*TmaEncodeFn) (CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encoder ()
{
    static TmaEncodeFn fn = [] () -> TmaEncodeFn {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (TmaEncodeFn)p;
    }();
    return fn;
}
// [rows x cols] float32 view of the planar input for K1: boxes of 32 rows x 32 floats, 128B swizzle, zeros outside
static bool tma_input_map (CUtensorMap* tm, const float* base, size_t stride, uint32_t rows, uint32_t cols)
{
    const cuuint64_t gdim[2] = {cols, rows}; const cuuint64_t gstr[1] = {(cuuint64_t)stride * sizeof (float)};
    const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    return tma_encoder () (tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*> (base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct b200m_ebu {
    int device; uint32_t n_inst, nchan; float fsamp; int fragm;
    int frcnt, wrind;                    // shared 50 ms fragment clock (host-tracked, see b200meters.h)
    EbuCoef cf;
    float *d_z = nullptr, *d_frpwr = nullptr, *d_fragpw = nullptr, *d_ring = nullptr, *d_binpow = nullptr, *d_out5 = nullptr;
    EbuCtl* d_ctl = nullptr; b200m_ebu_result* d_res = nullptr;
    int *d_histM = nullptr, *d_histS = nullptr, *d_cnt = nullptr;
    cudaStream_t own = nullptr; HostStage stage; bool last_host = false;
    bool use_tma = false;                // K1 tiles by TMA (opt-in, 16-byte aligned input only) instead of cp.async
    bool split = false;                  // K1 as two warps per 32 channels (ebu_kweight_split), opt-in with B200M_EBU_SPLIT=1: measured slower, see the kernel
    // Host mirror of every instance's S-histogram period (_div2, :234-241), kept in O(1) per fragment: an
    // integrating instance has div2 = (G - base) mod 10 where G counts fragments; cnt10[r] = number of integrating
    // instances with base = r.  The gated-statistics kernel (K2b) is launched only for fragments where some
    // instance wraps, i.e. cnt10[G] > 0.
    std::vector<uint8_t> integ, base, frozen; int cnt10[10]; int G = 0;
    void phase_reset () { integ.assign (n_inst, 0); base.assign (n_inst, 0); frozen.assign (n_inst, 0); for (int& c : cnt10) c = 0; G = 0; }
    void phase_ctl (int32_t inst, int cmd) {
        for (uint32_t i = 0; i < n_inst; ++i) {
            if (inst >= 0 && (uint32_t)inst != i) continue;
            if (cmd == 0 && integ[i]) { frozen[i] = (uint8_t)((G - base[i] + 10) % 10); cnt10[base[i]]--; integ[i] = 0; }
            else if (cmd == 1 && !integ[i]) { base[i] = (uint8_t)((G - frozen[i] + 10) % 10); cnt10[base[i]]++; integ[i] = 1; }
            else if (cmd == 2) { if (integ[i]) { cnt10[base[i]]--; base[i] = (uint8_t)G; cnt10[base[i]]++; } else frozen[i] = 0; }
        }
    }
    bool phase_tick () { G = (G + 1) % 10; return cnt10[G] > 0; }     // one fragment completed: does anyone wrap?
};

// Host-side coefficient design; restates Ebu_r128_proc::detect_init (ebu_r128_proc.cc:263-293).
// Same literals and the same float expression types (tan on a float argument is the float
// overload in C++), evaluated with the host libm, so every coefficient is bitwise the oracle's.
static void ebu_design (float fsamp, EbuCoef& k)
{
    const float rt = 1 / tanf (4712.3890f / fsamp);
    const float wa = rt / 1.12201f, wb = rt * 1.12201f;
    const float u = 1.4085f + 210.0f / fsamp;
    const float pa = u * wa, pb = wa * wa, pc = u * wb, pd = wb * wb;
    const float den = 1 + pa + pb;
    k.a0 = (1 + pc + pd) / den;
    k.a1 = (2 - 2 * pd) / den;
    k.a2 = (1 - pc + pd) / den;
    k.b1 = (2 - 2 * pb) / den;
    k.b2 = (1 - pa + pb) / den;
    const float q = 48.0f / fsamp;
    float ha = 4.9886075f * q, hb = 6.2298014f * q * q;
    const float hden = 1 + ha + hb;
    ha *= 2 / hden; hb *= 4 / hden;
    k.c3 = ha + hb; k.c4 = hb;
    const float g = 1.004995f / hden;
    k.a0 *= g; k.a1 *= g; k.a2 *= g;
}

static int ebu_ctl (b200m_ebu* h, int32_t inst, int cmd, void* stream)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    if (inst >= (int32_t)h->n_inst) return set_err (B200M_E_INVAL, "instance %d out of range", inst);
    DeviceGuard g (h->device);
    if (cmd == 3) h->phase_reset (); else h->phase_ctl (inst, cmd);
    // after process_host the bank runs on its own stream: a control on any other stream would race with the kernels in flight
    ebu_ctl_kernel<<<(h->n_inst + 127) / 128, 128, 0, h->last_host ? h->own : (cudaStream_t)stream>>> (
        (int)h->n_inst, inst, cmd, (int)h->nchan, h->d_z, h->d_frpwr, h->d_ring, h->d_ctl, h->d_res, h->d_histM, h->d_histS, h->d_cnt);
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

extern "C" {

int b200m_design_ebu (float fsamp, float o[7])
{
    if (!o || !(fsamp >= 1000.0f)) return set_err (B200M_E_INVAL, "bad argument");
    EbuCoef k; ebu_design (fsamp, k);
    o[0] = k.a0; o[1] = k.a1; o[2] = k.a2; o[3] = k.b1; o[4] = k.b2; o[5] = k.c3; o[6] = k.c4;
    return 0;
}

int b200m_ebu_create (b200m_ebu** out, int device, uint32_t n_inst, uint32_t nchan, float fsamp)
{
    if (!out) return set_err (B200M_E_INVAL, "NULL out pointer");
    *out = nullptr;
    if (n_inst == 0 || !(fsamp >= 1000.0f)) return set_err (B200M_E_INVAL, "bad n_inst/fsamp");
    if (nchan < 1 || nchan > 5) return set_err (B200M_E_INVAL, "nchan %u outside 1..5", nchan);
    if (b200m_device_count () <= 0) return set_err (B200M_E_NODEVICE, "no CUDA device: b200meters has no CPU path");
    DeviceGuard g (device);
    if (!g.ok) return set_err (B200M_E_NODEVICE, "cannot select CUDA device %d", device);
    b200m_ebu* h = new (std::nothrow) b200m_ebu;
    if (!h) return set_err (B200M_E_NOMEM, "host allocation failed");
    h->device = device; h->n_inst = n_inst; h->nchan = nchan; h->fsamp = fsamp;
    h->fragm = (int)fsamp / 20;                     // :170
    h->frcnt = h->fragm; h->wrind = 0;
    ebu_design (fsamp, h->cf);
    const size_t nch = (size_t)n_inst * nchan;
    float bp[100];
    for (int i = 0; i < 100; ++i) bp[i] = powf (10.0f, i / 100.0f);   // Ebu_r128_hist::initstat (:54-63)
    cudaError_t e = cudaSuccess;
    auto A = [&] (void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc (p, bytes); if (e == cudaSuccess) e = cudaMemset (*p, 0, bytes); } };
    A ((void**)&h->d_z, 4 * nch * sizeof (float));
    A ((void**)&h->d_frpwr, n_inst * sizeof (float));
    A ((void**)&h->d_fragpw, (size_t)EBU_MAXCHUNK * n_inst * sizeof (float));
    A ((void**)&h->d_ring, (size_t)64 * n_inst * sizeof (float));
    A ((void**)&h->d_binpow, 100 * sizeof (float));
    A ((void**)&h->d_out5, 8 * sizeof (float));
    A ((void**)&h->d_ctl, n_inst * sizeof (EbuCtl));
    A ((void**)&h->d_res, n_inst * sizeof (b200m_ebu_result));
    A ((void**)&h->d_histM, (size_t)HIST_PITCH * n_inst * sizeof (int));
    A ((void**)&h->d_histS, (size_t)HIST_PITCH * n_inst * sizeof (int));
    A ((void**)&h->d_cnt, (size_t)4 * n_inst * sizeof (int));
    if (e == cudaSuccess) e = cudaMemcpy (h->d_binpow, bp, sizeof (bp), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags (&h->own, cudaStreamNonBlocking);
    // K1 carries 102 KB of dynamic shared memory per CTA
    // ... and asks for the largest shared-memory carveout: with the default the driver configures the SM for just what this kernel
    // needs (132 KB), which leaves room for ONE CTA of the true-peak kernel that is meant to share the SM with it (r128.cu)
#define EBU_ATTR(NC, AL) if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_frag<NC, AL>, cudaFuncAttributeMaxDynamicSharedMemorySize, EBU_SMEM_BYTES); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_frag<NC, AL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
    EBU_ATTR (1, true); EBU_ATTR (1, false); EBU_ATTR (2, true); EBU_ATTR (2, false);
    EBU_ATTR (3, true); EBU_ATTR (3, false); EBU_ATTR (4, true); EBU_ATTR (4, false); EBU_ATTR (5, true); EBU_ATTR (5, false);
#undef EBU_ATTR
    if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, EBU_TMA_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, EBU_TMA_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_tma<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_tma<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // TMA staging is bit-identical and removes ~120 address instructions per tile, but measured no faster standalone (26.0 vs
    // 25.8 us per block) and 3 % slower inside the EBUr128 cycle (the mbarrier try_wait spin takes issue slots from the
    // co-running true-peak kernel, a scoreboard wait does not): opt-in with B200M_EBU_TMA=1
#define EBU_SATTR(NC, AL) if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_split<NC, AL>, cudaFuncAttributeMaxDynamicSharedMemorySize, EBU_SPLIT_SMEM); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute (ebu_kweight_split<NC, AL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
    EBU_SATTR (1, true); EBU_SATTR (1, false); EBU_SATTR (2, true); EBU_SATTR (2, false);
#undef EBU_SATTR
    if (const char* v = getenv ("B200M_EBU_SPLIT")) h->split = atoi (v) != 0;
    h->use_tma = false;
    if (const char* v = getenv ("B200M_EBU_TMA")) h->use_tma = atoi (v) != 0 && tma_encoder () != nullptr;
    if (e != cudaSuccess) { int rc = cuda_fail (e, "ebu_create allocations", __FILE__, __LINE__); b200m_ebu_destroy (h); return rc; }
    h->phase_reset ();
    int rc = b200m_ebu_reset (h, -1, nullptr);       // constructor + init() end in reset() (:153-173)
    if (rc == 0 && cudaDeviceSynchronize () != cudaSuccess) rc = set_err (B200M_E_CUDA, "reset kernel failed");
    if (rc) { b200m_ebu_destroy (h); return rc; }
    *out = h;
    return 0;
}

int b200m_ebu_destroy (b200m_ebu* h)
{
    if (!h) return 0;
    DeviceGuard g (h->device);
    cudaDeviceSynchronize ();
    cudaFree (h->d_z); cudaFree (h->d_frpwr); cudaFree (h->d_fragpw); cudaFree (h->d_ring); cudaFree (h->d_binpow);
    cudaFree (h->d_out5); cudaFree (h->d_ctl); cudaFree (h->d_res); cudaFree (h->d_histM); cudaFree (h->d_histS); cudaFree (h->d_cnt);
    h->stage.release ();
    if (h->own) cudaStreamDestroy (h->own);
    delete h;
    return 0;
}

int b200m_ebu_reset (b200m_ebu* h, int32_t inst, void* stream)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    if (inst != -1) return set_err (B200M_E_UNSUPPORTED, "reset() restarts the shared fragment clock: only inst = -1");
    h->frcnt = h->fragm; h->wrind = 0;
    return ebu_ctl (h, -1, 3, stream);
}
int b200m_ebu_clear (b200m_ebu* h, int32_t inst, void* stream)
{
    // what reset() does to ONE instance -- integration off, filter states, 64-fragment ring, loudness values, histograms -- without
    // restarting the bank's shared 50 ms fragment clock
    if (!h || inst < 0 || inst >= (int32_t)h->n_inst) return set_err (B200M_E_INVAL, "bad argument");
    DeviceGuard g (h->device);
    h->phase_ctl (inst, 0); h->phase_ctl (inst, 2);
    ebu_ctl_kernel<<<(h->n_inst + 127) / 128, 128, 0, h->last_host ? h->own : (cudaStream_t)stream>>> (
        (int)h->n_inst, inst, 3, (int)h->nchan, h->d_z, h->d_frpwr, h->d_ring, h->d_ctl, h->d_res, h->d_histM, h->d_histS, h->d_cnt);
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaGetLastError ());
    return 0;
}
int b200m_ebu_integr_start (b200m_ebu* h, int32_t inst, void* stream) { return ebu_ctl (h, inst, 1, stream); }
int b200m_ebu_integr_pause (b200m_ebu* h, int32_t inst, void* stream) { return ebu_ctl (h, inst, 0, stream); }
int b200m_ebu_integr_reset (b200m_ebu* h, int32_t inst, void* stream) { return ebu_ctl (h, inst, 2, stream); }

// One Ebu_r128_proc::process call for every instance.  `nsl` instance slices [bounds[s], bounds[s+1]) are launched
// separately, slice s after event ready[s] (the host->device copy of its rows) when `ready` is given; the
// fragment/gating kernels run once, after the last slice.
int ebu_process_sliced (b200m_ebu* h, const float* d_in, size_t stride, uint32_t nfram, cudaStream_t st,
                        int nsl, const uint32_t* bounds, cudaEvent_t* ready, int (*after_k1) (void*), void* after_arg)
{
    const int nch = (int)(h->n_inst * h->nchan);
    const bool aligned = ((uintptr_t)d_in % 16 == 0) && (stride % 4 == 0);
    uint32_t done = 0;
    while (done < nfram) {
        // cut [done, nfram) at fragment edges, at most EBU_MAXCHUNK pieces per launch (:212-216)
        EbuChunks ck; ck.n = 0;
        int nfrag = 0; uint32_t pos = 0;
        while (done + pos < nfram && ck.n < EBU_MAXCHUNK) {
            const uint32_t rem = nfram - done - pos;
            const uint32_t k = (uint32_t)h->frcnt < rem ? (uint32_t)h->frcnt : rem;
            pos += k; h->frcnt -= (int)k;
            uint32_t v = pos;
            if (h->frcnt == 0) { v |= 0x80000000u; h->frcnt = h->fragm; ++nfrag; }
            ck.v[ck.n++] = v;
        }
        const float* src = d_in + done;
        const bool al = aligned && (done % 4 == 0);
        CUtensorMap tmap;
        const bool tma = al && h->use_tma && h->nchan <= 2 && tma_input_map (&tmap, src, stride, (uint32_t)nch, pos);
        for (int sl = 0; sl < nsl; ++sl) {
            const int kf = (int)(bounds[sl] * h->nchan), ke = (int)(bounds[sl + 1] * h->nchan);
            if (ke <= kf) continue;
            if (ready && done == 0) B200M_CUDA (cudaStreamWaitEvent (st, ready[sl], 0));
            const int cpw = (32 / (int)h->nchan) * (int)h->nchan;          // a warp takes whole instances only
            const int nwarps = (ke - kf + cpw - 1) / cpw;
            dim3 grid ((nwarps + EBU_WARPS - 1) / EBU_WARPS), blk (EBU_WARPS * 32);
#define EBU_K1(NC, AL) ebu_kweight_frag<NC, AL><<<grid, blk, EBU_SMEM_BYTES, st>>> (src, stride, nch, kf, ke, (int)pos, h->cf, ck, (float)h->fragm, h->d_z, h->d_frpwr, h->d_fragpw, (int)h->n_inst, (after_k1 && done == 0) ? 1 : 0)
#define EBU_K1T(NC) ebu_kweight_tma<NC><<<grid, blk, EBU_TMA_SMEM, st>>> (tmap, nch, kf, ke, (int)pos, h->cf, ck, (float)h->fragm, h->d_z, h->d_frpwr, h->d_fragpw, (int)h->n_inst, (after_k1 && done == 0) ? 1 : 0)
            if (h->nchan > 2) {                                  // surround banks (3..5 channels): the one-warp kernel, lanes grouped per instance
                switch (h->nchan * 2 + (al ? 1 : 0)) {
                case 7: EBU_K1 (3, true); break; case 6: EBU_K1 (3, false); break;
                case 9: EBU_K1 (4, true); break; case 8: EBU_K1 (4, false); break;
                case 11: EBU_K1 (5, true); break; default: EBU_K1 (5, false); break;
                }
            }
            else if (h->split && !tma) {
                dim3 sgrid ((nwarps + EBU_SPLIT_PAIRS - 1) / EBU_SPLIT_PAIRS), sblk (2 * EBU_SPLIT_PAIRS * 32);
#define EBU_K1S(NC, AL) ebu_kweight_split<NC, AL><<<sgrid, sblk, EBU_SPLIT_SMEM, st>>> (src, stride, nch, kf, ke, (int)pos, h->cf, ck, (float)h->fragm, h->d_z, h->d_frpwr, h->d_fragpw, (int)h->n_inst, (after_k1 && done == 0) ? 1 : 0)
                if (h->nchan == 1) { if (al) EBU_K1S (1, true); else EBU_K1S (1, false); }
                else               { if (al) EBU_K1S (2, true); else EBU_K1S (2, false); }
#undef EBU_K1S
            }
            else if (tma) { if (h->nchan == 1) EBU_K1T (1); else EBU_K1T (2); }
            else if (h->nchan == 1) { if (al) EBU_K1 (1, true); else EBU_K1 (1, false); }
            else               { if (al) EBU_K1 (2, true); else EBU_K1 (2, false); }
#undef EBU_K1
#undef EBU_K1T
            B200M_LAUNCHED (1);
        }
        if (after_k1 && done == 0) { if (int rc = after_k1 (after_arg)) return rc; }      // work to enqueue right behind the first K1
        for (int f = 0; f < nfrag; ++f) {                    // fragments complete in order; each may trigger gating
            ebu_fragment_kernel<<<(h->n_inst + K2A_THREADS - 1) / K2A_THREADS, K2A_THREADS, 0, st>>> (
                (int)h->n_inst, f, h->wrind, h->d_fragpw, h->d_ring, h->d_ctl, h->d_res, h->d_histM, h->d_histS, h->d_cnt);
            B200M_LAUNCHED (1);
            h->wrind = (h->wrind + 1) & 63;
            if (h->phase_tick ()) {
                ebu_gate_kernel<<<(h->n_inst + K2B_WARPS - 1) / K2B_WARPS, K2B_WARPS * 32, 0, st>>> (
                    (int)h->n_inst, h->d_ctl, h->d_res, h->d_histM, h->d_histS, h->d_cnt, h->d_binpow);
                B200M_LAUNCHED (1);
            }
        }
        done += pos;
    }
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

static int ebu_process (b200m_ebu* h, const float* d_in, size_t stride, uint32_t nfram, cudaStream_t st)
{
    const uint32_t bounds[2] = {0, h->n_inst};
    return ebu_process_sliced (h, d_in, stride, nfram, st, 1, bounds, nullptr, nullptr, nullptr);
}

int b200m_ebu_process_device (b200m_ebu* h, const float* d_in, size_t stride, uint32_t nfram, void* stream)
{
    if (int rc = check_block_args (h, d_in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    h->last_host = false;
    return ebu_process (h, d_in, stride, nfram, (cudaStream_t)stream);
}

int b200m_ebu_process_host (b200m_ebu* h, const float* in, size_t stride, uint32_t nfram)
{
    if (int rc = check_block_args (h, in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    B200M_ENTER_HOST_PATH (h);
    const size_t nch = (size_t)h->n_inst * h->nchan;
    if (h->stage.ensure (nch, nfram)) return set_err (B200M_E_NOMEM, "staging buffer allocation failed");
    B200M_CUDA (cudaMemcpy2DAsync (h->stage.d, h->stage.cap * sizeof (float), in, stride * sizeof (float),
                                   (size_t)nfram * sizeof (float), nch, cudaMemcpyHostToDevice, h->own));
    h->last_host = true;
    return ebu_process (h, h->stage.d, h->stage.cap, nfram, h->own);
}

static cudaStream_t ebu_stream (b200m_ebu* h, void* stream) { return h->last_host ? h->own : (cudaStream_t)stream; }

int b200m_ebu_results (b200m_ebu* h, b200m_ebu_result* out, void* stream)
{
    if (!h || !out) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    B200M_CUDA (cudaMemcpyAsync (out, h->d_res, h->n_inst * sizeof (b200m_ebu_result), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

int b200m_ebu_histogram (b200m_ebu* h, uint32_t inst, int32_t* hist_M, int32_t* hist_S, void* stream)
{
    if (!h || !hist_M || !hist_S || inst >= h->n_inst) return set_err (B200M_E_INVAL, "bad argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    B200M_CUDA (cudaMemcpyAsync (hist_M, h->d_histM + (size_t)inst * HIST_PITCH, 751 * sizeof (int), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaMemcpyAsync (hist_S, h->d_histS + (size_t)inst * HIST_PITCH, 751 * sizeof (int), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

// ---- snapshot / restore (the reference never saves DSP state; a batch engine integrating for hours should) ----
// blob = header, the host-side clocks and phase book-keeping, then every device state array in a fixed order
namespace {
struct EbuSnapHead { uint32_t magic, n_inst, nchan; float fsamp; int32_t frcnt, wrind, G, cnt10[10]; };
constexpr uint32_t EBU_SNAP_MAGIC = 0x42453031u;              // "BE01"
struct EbuSeg { void* p; size_t bytes; };
int ebu_segments (b200m_ebu* h, EbuSeg* seg)
{
    const size_t n = h->n_inst, nch = (size_t)h->n_inst * h->nchan;
    int k = 0;
    seg[k++] = {h->d_z, 4 * nch * sizeof (float)}; seg[k++] = {h->d_frpwr, n * sizeof (float)}; seg[k++] = {h->d_ring, 64 * n * sizeof (float)};
    seg[k++] = {h->d_ctl, n * sizeof (EbuCtl)}; seg[k++] = {h->d_res, n * sizeof (b200m_ebu_result)};
    seg[k++] = {h->d_histM, (size_t)HIST_PITCH * n * sizeof (int)}; seg[k++] = {h->d_histS, (size_t)HIST_PITCH * n * sizeof (int)};
    seg[k++] = {h->d_cnt, 4 * n * sizeof (int)};
    return k;
}
}

size_t b200m_ebu_snapshot_size (b200m_ebu* h)
{
    if (!h) return 0;
    EbuSeg seg[8]; const int k = ebu_segments (h, seg);
    size_t b = sizeof (EbuSnapHead) + 3 * (size_t)h->n_inst;
    b = (b + 15) & ~size_t (15);
    for (int i = 0; i < k; ++i) b += (seg[i].bytes + 15) & ~size_t (15);
    return b;
}

int b200m_ebu_snapshot (b200m_ebu* h, void* buf, size_t bytes, void* stream)
{
    if (!h || !buf || bytes < b200m_ebu_snapshot_size (h)) return set_err (B200M_E_INVAL, "bad argument / buffer too small");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    uint8_t* o = (uint8_t*)buf;
    EbuSnapHead hd = {EBU_SNAP_MAGIC, h->n_inst, h->nchan, h->fsamp, h->frcnt, h->wrind, h->G, {0}};
    memcpy (hd.cnt10, h->cnt10, sizeof (hd.cnt10));
    memcpy (o, &hd, sizeof (hd)); o += sizeof (hd);
    memcpy (o, h->integ.data (), h->n_inst); o += h->n_inst; memcpy (o, h->base.data (), h->n_inst); o += h->n_inst; memcpy (o, h->frozen.data (), h->n_inst); o += h->n_inst;
    o = (uint8_t*)buf + ((sizeof (hd) + 3 * (size_t)h->n_inst + 15) & ~size_t (15));
    EbuSeg seg[8]; const int k = ebu_segments (h, seg);
    for (int i = 0; i < k; ++i) { B200M_CUDA (cudaMemcpyAsync (o, seg[i].p, seg[i].bytes, cudaMemcpyDeviceToHost, st)); o += (seg[i].bytes + 15) & ~size_t (15); }
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

int b200m_ebu_restore (b200m_ebu* h, const void* buf, size_t bytes, void* stream)
{
    if (!h || !buf || bytes < b200m_ebu_snapshot_size (h)) return set_err (B200M_E_INVAL, "bad argument / buffer too small");
    EbuSnapHead hd; memcpy (&hd, buf, sizeof (hd));
    if (hd.magic != EBU_SNAP_MAGIC || hd.n_inst != h->n_inst || hd.nchan != h->nchan || hd.fsamp != h->fsamp)
        return set_err (B200M_E_INVAL, "snapshot does not match this bank (instances / channels / sample rate)");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    const uint8_t* o = (const uint8_t*)buf + sizeof (hd);
    h->frcnt = hd.frcnt; h->wrind = hd.wrind; h->G = hd.G; memcpy (h->cnt10, hd.cnt10, sizeof (hd.cnt10));
    memcpy (h->integ.data (), o, h->n_inst); o += h->n_inst; memcpy (h->base.data (), o, h->n_inst); o += h->n_inst; memcpy (h->frozen.data (), o, h->n_inst);
    o = (const uint8_t*)buf + ((sizeof (hd) + 3 * (size_t)h->n_inst + 15) & ~size_t (15));
    EbuSeg seg[8]; const int k = ebu_segments (h, seg);
    for (int i = 0; i < k; ++i) { B200M_CUDA (cudaMemcpyAsync (seg[i].p, o, seg[i].bytes, cudaMemcpyHostToDevice, st)); o += (seg[i].bytes + 15) & ~size_t (15); }
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

int b200m_ebu_coeffs (const b200m_ebu* h, float o[7])
{
    if (!h || !o) return set_err (B200M_E_INVAL, "NULL argument");
    o[0] = h->cf.a0; o[1] = h->cf.a1; o[2] = h->cf.a2; o[3] = h->cf.b1; o[4] = h->cf.b2; o[5] = h->cf.c3; o[6] = h->cf.c4;
    return 0;
}

int b200m_ebu_state (b200m_ebu* h, uint32_t inst, float* z, float* power64, float* frpwr, int32_t c4[4], void* stream)
{
    if (!h || !z || !power64 || !frpwr || !c4 || inst >= h->n_inst) return set_err (B200M_E_INVAL, "bad argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    const size_t nch = (size_t)h->n_inst * h->nchan;
    for (uint32_t c = 0; c < h->nchan; ++c)
        for (int q = 0; q < 4; ++q)
            B200M_CUDA (cudaMemcpyAsync (z + 4 * c + q, h->d_z + q * nch + (size_t)inst * h->nchan + c, sizeof (float), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaMemcpy2DAsync (power64, sizeof (float), h->d_ring + inst, (size_t)h->n_inst * sizeof (float), sizeof (float), 64, cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaMemcpyAsync (frpwr, h->d_frpwr + inst, sizeof (float), cudaMemcpyDeviceToHost, st));
    EbuCtl c;
    B200M_CUDA (cudaMemcpyAsync (&c, h->d_ctl + inst, sizeof (c), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaStreamSynchronize (st));
    c4[0] = h->frcnt; c4[1] = h->wrind; c4[2] = c.div1; c4[3] = c.div2;
    return 0;
}

int b200m_ebu_mix_reduce (b200m_ebu* h, int32_t* d_out, void* stream)
{
    if (!h || !d_out) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    B200M_CUDA (cudaMemsetAsync (d_out, 0, B200M_MIX_WORDS * sizeof (int32_t), st));
    const unsigned slices = (unsigned)std::min<uint32_t> (h->n_inst, 128u);      // ~24 x 128 CTAs: enough loads in flight to stream the histograms
    ebu_mix_reduce_kernel<<<dim3 ((B200M_MIX_WORDS + 63) / 64, slices), 64, 0, st>>> ((int)h->n_inst, h->d_histM, h->d_histS, h->d_cnt, d_out);
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

int b200m_ebu_mix_finish (b200m_ebu* h, const int32_t* d_mix, float out5[5], void* stream)
{
    if (!h || !d_mix || !out5) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ebu_stream (h, stream);
    ebu_mix_finish_kernel<<<1, 32, 0, st>>> (d_mix, h->d_binpow, h->d_out5);
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaMemcpyAsync (out5, h->d_out5, 5 * sizeof (float), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

}  // extern "C"
