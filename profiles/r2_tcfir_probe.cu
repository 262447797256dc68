// r2_tcfir_probe.cu -- standalone probe (not part of the library): the 4x true-peak FIR as a Toeplitz GEMM on the tcgen05 tensor cores.
//
// Formulation.  Rows are 16-sample blocks of one channel (128 rows = 8 channels x 256 samples), A_row[k] = x[16 tb - 48 + k], k < 64;
// columns are (output position j < 16, phase 1..3), B[k][(j, ph)] = h_ph[j + 48 - k] (zero outside the 48 taps): 25 % of the MACs
// multiply zeros.  fp32 accuracy comes from the 3xTF32 split x = hi + lo (hi = top 11 significand bits): D = A_hi B_hi + A_hi B_lo +
// A_lo B_hi, as two instructions per K step: A_hi x [B_hi | B_lo] (N = 96) and A_lo x B_hi (N = 48), kind::tf32, M = 128, K = 8,
// accumulators in TMEM.  Phase 0 of the filter is the input delayed by 24 samples and is taken from the window directly.
// v1 (tcfir_kernel): A im2col'd into shared memory (K-major, no swizzle, 16-byte chunks), all warps in lock step.
// v5 (tcfir2_kernel; v2 - v4 were steps towards it): A written to TMEM by four builder warps (tcgen05.st), B in shared memory, a dedicated
//     MMA warp, four epilogue warps (tcgen05.ld + max + atomicMax), a loader thread filling eight input stages by cp.async.bulk; every
//     hand-over is an mbarrier; A and D double-buffered in TMEM (512 columns).  The library's tpmax_tc_kernel is this kernel plus the
//     bank's history, the group book-keeping and the EBUr128 epilogue.
//
// MEASURED on B200 (round 2; 16384 channels x 1024 frames, the headline's bank):
//   * numerics: channel maxima within 6.7e-7 relative of a float64 FIR (the contract's tolerance is 1.15e-5), ragged block lengths
//     and channel counts included -- descriptors, instruction descriptor and TMEM layout as written here are right;
//   * time: v1 203 us, v2 (A in TMEM) 87 us, v4 (8 producer warps, 4 input stages) 81-85 us = the CUDA-core tpmax_kernel (81 us);
//     v5 52.8 us;
//   * ablation of v4: no MMA 78 us, no tcgen05.st 76, no tcgen05.ld 86, NO INPUT LOADS 54, only the producers' ALU work and barriers
//     45: too few bytes in flight (three tiles of bulk copies) and all stages of a tile in the same warps; v5 fixes both;
//   * ablation of v5 (mode bits of argv[5]): no input loads 49.4, no MMAs 47.3, no TMEM stores / loads / epilogue arithmetic 43.2,
//     everything stubbed 28.8;
//   * r2_mma_bench.cu: a tcgen05.mma with K = 8 costs >= ~110 cycles whatever N <= 128 (1890 MAC/clk at N = 256 with A in TMEM), so
//     small-N Toeplitz tiles waste the tensor pipe as well.
// Shipped as tpmax_tc_kernel (csrc/tpk.cu).  Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tcfir r2_tcfir_probe.cu
// Run: ./tcfir <channels> <frames> <compare every output 0|1> <kernel 1|2> <ablation mode bits>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf ("CUDA error %s at %s:%d\n", cudaGetErrorString (e_), __FILE__, __LINE__); exit (1); } } while (0)

constexpr int ROWS = 128, NCOL = 48, KCH = 16;           // 128 rows = 8 channels x 16 blocks of 16 samples; K = 64 = 16 chunks of 4
constexpr int A_LBO = 2048 + 32;                          // bytes between K chunks of A (padded: conflict-free staging stores)
constexpr int A_BYTES = KCH * A_LBO;                      // one of {hi, lo}
constexpr int B_LBO = NCOL * 16;                          // 768
constexpr int B_BYTES = KCH * B_LBO;                      // 12288 per {hi, lo}
constexpr int SBO = 128;
constexpr int SMEM_BYTES = 2 * 2 * A_BYTES + 2 * B_BYTES + 64;
constexpr int TILE_S = 256;                               // samples per tile and channel

__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return (uint32_t)__cvta_generic_to_shared (p); }
__device__ __forceinline__ uint64_t make_desc (uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32 (uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                  :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16 (uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float (r[i]);
}

// in: [n_chan][stride] with 48 samples of history BEFORE `in` (in[-48..-1] valid); nfram multiple of 4.
__global__ void __launch_bounds__ (128, 1)
tcfir_kernel (const float* __restrict__ in, size_t stride, int n_chan, int nfram, const float* __restrict__ bcanon, unsigned* __restrict__ out_max, float* __restrict__ dbg)
{
    extern __shared__ __align__ (128) uint8_t smem[];
    uint8_t* sA = smem;                                      // [buf 2][hi, lo][A_BYTES]
    uint8_t* sB = smem + 4 * A_BYTES;                        // [hi, lo][B_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*> (smem + 4 * A_BYTES + 2 * B_BYTES);   // [2] MMA-done barriers
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (nfram + TILE_S - 1) / TILE_S;
    const int ngroups = (n_chan + 7) / 8;
    const int ntiles = ngroups * nchunks;

    for (int i = tid; i < 2 * B_BYTES / 16; i += 128) reinterpret_cast<float4*> (sB)[i] = reinterpret_cast<const float4*> (bcanon)[i];
    if (tid == 0) {
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bars[0])));
        asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bars[1])));
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32 (&s_tmem)), "n"(128) : "memory");
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 48, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

    auto stage = [&] (int tile, int buf) {
        const int grp = tile / nchunks, chunk = tile - grp * nchunks;
        const int c0 = grp * 8, s0 = chunk * TILE_S;
        uint8_t* ah = sA + (size_t)buf * 2 * A_BYTES; uint8_t* al = ah + A_BYTES;
        for (int idx = tid; idx < 8 * 76; idx += 128) {
            const int c8 = idx / 76, q = idx - 76 * c8;
            const int ch = min (c0 + c8, n_chan - 1);
            const int pos = s0 - 48 + 4 * q;                  // block-relative position of the float4's first sample
            float4 v = make_float4 (0.f, 0.f, 0.f, 0.f);
            if (pos < nfram) v = *reinterpret_cast<const float4*> (in + (size_t)ch * stride + pos);      // nfram % 4 == 0 here
            float4 h, l;
            h.x = __uint_as_float (__float_as_uint (v.x) & 0xffffe000u); l.x = v.x - h.x;
            h.y = __uint_as_float (__float_as_uint (v.y) & 0xffffe000u); l.y = v.y - h.y;
            h.z = __uint_as_float (__float_as_uint (v.z) & 0xffffe000u); l.z = v.z - h.z;
            h.w = __uint_as_float (__float_as_uint (v.w) & 0xffffe000u); l.w = v.w - h.w;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int c = (q & 3) + 4 * m, tb = (q >> 2) - m;
                if (tb >= 0 && tb < 16) {
                    const int off = c * A_LBO + (c8 * 16 + tb) * 16;
                    *reinterpret_cast<float4*> (ah + off) = h;
                    *reinterpret_cast<float4*> (al + off) = l;
                }
            }
        }
    };
    auto issue = [&] (int buf) {
        // one thread: 8 K steps x {hi hi, hi lo, lo hi}
        const uint32_t ah = smem_u32 (sA + (size_t)buf * 2 * A_BYTES), al = ah + A_BYTES;
        const uint32_t bh = smem_u32 (sB), bl = bh + B_BYTES;
        const uint32_t d = tmem + (uint32_t)buf * 64u;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const uint64_t dah = make_desc (ah + 2 * s * A_LBO, A_LBO, SBO), dal = make_desc (al + 2 * s * A_LBO, A_LBO, SBO);
            const uint64_t dbh = make_desc (bh + 2 * s * B_LBO, B_LBO, SBO), dbl = make_desc (bl + 2 * s * B_LBO, B_LBO, SBO);
            mma_tf32 (d, dal, dbh, idesc, s > 0 ? 1u : 0u);
            mma_tf32 (d, dah, dbl, idesc, 1u);
            mma_tf32 (d, dah, dbh, idesc, 1u);
        }
        asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bars[buf])) : "memory");
    };
    auto epilogue = [&] (int tile, int buf, uint32_t parity) {
        const int grp = tile / nchunks, chunk = tile - grp * nchunks;
        const int c0 = grp * 8, s0 = chunk * TILE_S;
        mbar_wait (smem_u32 (&bars[buf]), parity);
        asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int r = tid, c8 = r >> 4, tb = r & 15;
        const int vj = min (16, max (0, nfram - (s0 + 16 * tb)));      // valid output positions of this row
        float mx = 0.0f;
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)buf * 64u;
#pragma unroll
        for (int ph = 0; ph < 3; ++ph) {
            float v[16];
            tmem_ld16 (taddr + 16 * ph, v);
            if (dbg) for (int j = 0; j < 16; ++j) dbg[((size_t)tile * 128 + r) * 48 + ph * 16 + j] = v[j];
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < vj) mx = fmaxf (mx, fabsf (v[j]));
        }
        // phase 0 is the input delayed by 24 samples: A_row[k'] for k' = 24 + j (hi + lo = x exactly)
        const uint8_t* ah = sA + (size_t)buf * 2 * A_BYTES; const uint8_t* al = ah + A_BYTES;
#pragma unroll
        for (int cc = 6; cc < 10; ++cc) {
            const float4 h = *reinterpret_cast<const float4*> (ah + cc * A_LBO + r * 16), l = *reinterpret_cast<const float4*> (al + cc * A_LBO + r * 16);
            const int j0 = 4 * (cc - 6);
            if (j0 + 0 < vj) mx = fmaxf (mx, fabsf (h.x + l.x));
            if (j0 + 1 < vj) mx = fmaxf (mx, fabsf (h.y + l.y));
            if (j0 + 2 < vj) mx = fmaxf (mx, fabsf (h.z + l.z));
            if (j0 + 3 < vj) mx = fmaxf (mx, fabsf (h.w + l.w));
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) mx = fmaxf (mx, __shfl_xor_sync (0xffffffffu, mx, o));
        if (tb == 0 && c0 + c8 < n_chan && mx > 0.0f) atomicMax (out_max + c0 + c8, __float_as_uint (mx));
        asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    int it = 0;
    int prev_tile = -1;
    long long t_stage = 0, t_sync = 0, t_issue = 0, t_epi = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        // buffer `buf` (smem A and TMEM accumulator) was last used by iteration it - 2, whose epilogue ran in iteration it - 1
        long long t0 = clock64 ();
        stage (tile, buf);
        asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
        long long t1 = clock64 ();
        __syncthreads ();
        long long t2 = clock64 ();
        if (tid == 0) { asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory"); issue (buf); }
        long long t3 = clock64 ();
        if (prev_tile >= 0) epilogue (prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
        long long t4 = clock64 ();
        t_stage += t1 - t0; t_sync += t2 - t1; t_issue += t3 - t2; t_epi += t4 - t3;
        prev_tile = tile;
    }
    if (dbg == nullptr && blockIdx.x == 3 && (tid == 0 || tid == 64) && out_max[0] == 0xffffffffu) printf ("never\n");
    if (blockIdx.x == 3 && (tid == 0 || tid == 64) && ntiles > 4000) printf ("tid %d: %d tiles, cycles per tile: stage %lld sync %lld issue %lld epilogue(+mma wait) %lld\n", tid, it, t_stage / it, t_sync / it, t_issue / it, t_epi / it);
    if (prev_tile >= 0) epilogue (prev_tile, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));
    __syncthreads ();
    if (warp == 0) asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(128) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------------
// v3: A operand in TMEM (tcgen05.st by the producer warps), B = [B_hi | B_lo] (N = 96) in shared memory, a dedicated MMA warp,
// input tiles by bulk copies (one thread) completing on an mbarrier.
constexpr int XPITCH = 308;                                // floats per channel row of the input tile: 48 + 256 + 4; = 20 mod 32
constexpr int B2_LBO = 96 * 16;                            // [B_hi | B_lo]: 96 rows per K chunk
constexpr int B2_BYTES = KCH * B2_LBO;                     // 24576
__device__ __forceinline__ void mma_tf32_ts (uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                  :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st32 (uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile ("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                  "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                  :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                     "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                     "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                     "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait (uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld8_nowait (uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
constexpr int XSTAGES = 8;
constexpr int V4_SMEM = B2_BYTES + XSTAGES * 8 * XPITCH * 4 + 4 * 128 * 4 + 256;
constexpr int NTHREADS = 320;                              // warps 0-3 build A, 4-7 epilogue, 8 MMA issue, 9 input loads

// v5: every stage of a tile on its own warps.  mbarriers: afull[2] A[b] written (128), done[2] MMAs of the tile complete (commit),
// dempty[2] D[b] read out (128), xfull[s] input stage landed (tx), xempty[s] input stage read (128).
__global__ void __launch_bounds__ (NTHREADS, 1)
tcfir2_kernel (const float* __restrict__ in, size_t stride, int n_chan, int nfram, const float* __restrict__ bcanon, unsigned* __restrict__ out_max, float* __restrict__ dbg, int mode)
{
    extern __shared__ __align__ (128) uint8_t smem[];
    uint8_t* sB = smem;                                                        // [B_hi | B_lo], K-major canonical, 96 rows
    float* xbuf = reinterpret_cast<float*> (smem + B2_BYTES);                  // [XSTAGES][8][XPITCH]
    float* p0buf = reinterpret_cast<float*> (smem + B2_BYTES + XSTAGES * 8 * XPITCH * 4);       // [4][128]
    uint64_t* bars = reinterpret_cast<uint64_t*> (smem + B2_BYTES + XSTAGES * 8 * XPITCH * 4 + 4 * 128 * 4);
    uint64_t* afull = bars; uint64_t* done = bars + 2; uint64_t* dempty = bars + 4; uint64_t* xfull = bars + 6; uint64_t* xempty = bars + 6 + XSTAGES;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (nfram + TILE_S - 1) / TILE_S;
    const int ntiles = ((n_chan + 7) / 8) * nchunks;
    const int n_it = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (int i = tid; i < B2_BYTES / 16; i += NTHREADS) reinterpret_cast<float4*> (sB)[i] = reinterpret_cast<const float4*> (bcanon)[i];
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile ("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(smem_u32 (&afull[i])));
            asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&done[i])));
            asm volatile ("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(smem_u32 (&dempty[i])));
        }
        for (int i = 0; i < XSTAGES; ++i) {
            asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&xfull[i])));
            asm volatile ("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(smem_u32 (&xempty[i])));
        }
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32 (&s_tmem)), "n"(512) : "memory");
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");             // B (generic stores) -> tensor core (async proxy)
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc96 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
    const uint32_t idesc48 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
    // TMEM columns: A[b] hi at 128 b, lo at 128 b + 64; D[b] at 256 + 128 b: [0,48) hi hi + lo hi, [48,96) hi lo
    // lane -> (channel, 16-sample block) of the row it owns, chosen so that an LDS.128 phase of the builders is conflict-free
    const int wq = warp & 3;
    const int c8 = (wq & 1) * 4 + (lane & 3);
    const int tb = ((wq >> 1) * 4 + (lane >> 3)) * 2 + ((lane >> 2) & 1);
    const int row = 32 * wq + lane;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;

    if (warp == 9) {
        // ---------------- input loads: one thread, XSTAGES tiles ahead
        if (lane == 0 && !(mode & 8))
            for (int it = 0; it < n_it; ++it) {
                const int st = it % XSTAGES;
                if (it >= XSTAGES) mbar_wait (smem_u32 (&xempty[st]), (uint32_t)((it / XSTAGES - 1) & 1));
                const int tile = blockIdx.x + it * gridDim.x;
                const int grp = tile / nchunks, chunk = tile - grp * nchunks;
                const int c0 = grp * 8, s0 = chunk * TILE_S;
                const uint32_t xb = smem_u32 (xbuf + (size_t)st * 8 * XPITCH), bar = smem_u32 (&xfull[st]);
                const uint32_t bytes = (uint32_t)min (304, nfram - (s0 - 48)) * 4u;
                asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(8u * bytes) : "memory");
                for (int cc = 0; cc < 8; ++cc) {
                    const float* src = in + (size_t)min (c0 + cc, n_chan - 1) * stride + (s0 - 48);
                    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                  :: "r"(xb + (uint32_t)(cc * XPITCH * 4)), "l"(src), "r"(bytes), "r"(bar) : "memory");
                }
            }
    } else if (warp == 8) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t bb = smem_u32 (sB);
            for (int it = 0; it < n_it; ++it) {
                const int b = it & 1;
                mbar_wait (smem_u32 (&afull[b]), (uint32_t)((it >> 1) & 1));
                if (it >= 2) mbar_wait (smem_u32 (&dempty[b]), (uint32_t)(((it - 2) >> 1) & 1));
                asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + 256u + 128u * b, ah = tmem + 128u * b, al = ah + 64u;
                if (!(mode & 1)) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const uint64_t db = make_desc (bb + 2 * s * B2_LBO, B2_LBO, SBO);
                        mma_tf32_ts (d, ah + 8 * s, db, idesc96, s > 0 ? 1u : 0u);      // A_hi x [B_hi | B_lo]
                        mma_tf32_ts (d, al + 8 * s, db, idesc48, 1u);                    // A_lo x B_hi
                    }
                }
                asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&done[b])) : "memory");
            }
        }
    } else if (warp < 4) {
        // ---------------- builders: this row's 64-sample window -> {hi, lo} -> TMEM lane
        for (int it = 0; it < n_it; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int chunk = tile % nchunks, s0 = chunk * TILE_S;
            const int b = it & 1, st = it % XSTAGES;
            if (!(mode & 8)) mbar_wait (smem_u32 (&xfull[st]), (uint32_t)((it / XSTAGES) & 1));
            const float* xw = xbuf + (size_t)st * 8 * XPITCH + c8 * XPITCH + 16 * tb;
            float4 v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = *reinterpret_cast<const float4*> (xw + 4 * c);
            if (!(mode & 8)) asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (&xempty[st])) : "memory");
            const int vj = min (16, max (0, nfram - (s0 + 16 * tb)));
            const int nin = nfram - (s0 - 48) - 16 * tb;          // window elements k < nin lie inside the block; the rest read as 0
            if (nin < 64) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (4 * c + 0 >= nin) v[c].x = 0.0f;
                    if (4 * c + 1 >= nin) v[c].y = 0.0f;
                    if (4 * c + 2 >= nin) v[c].z = 0.0f;
                    if (4 * c + 3 >= nin) v[c].w = 0.0f;
                }
            }
            float p0 = 0.0f;                                       // phase 0 = the input delayed by 24 samples: elements 24 + j
#pragma unroll
            for (int c = 6; c < 10; ++c) {
                const int j0 = 4 * (c - 6);
                if (j0 + 0 < vj) p0 = fmaxf (p0, fabsf (v[c].x));
                if (j0 + 1 < vj) p0 = fmaxf (p0, fabsf (v[c].y));
                if (j0 + 2 < vj) p0 = fmaxf (p0, fabsf (v[c].z));
                if (j0 + 3 < vj) p0 = fmaxf (p0, fabsf (v[c].w));
            }
            if (it >= 2) mbar_wait (smem_u32 (&done[b]), (uint32_t)(((it - 2) >> 1) & 1));       // the MMAs of tile it - 2 have read A[b]
            asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
            p0buf[(it & 3) * 128 + row] = p0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 q = v[8 * half + c];
                    const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t h = __float_as_uint (vv[e]) & 0xffffe000u;
                        hi[4 * c + e] = h; lo[4 * c + e] = __float_as_uint (vv[e] - __uint_as_float (h));
                    }
                }
                if (!(mode & 2)) {
                    tmem_st32 (tmem + lane_base + 128u * b + 32u * half, hi);
                    tmem_st32 (tmem + lane_base + 128u * b + 64u + 32u * half, lo);
                }
            }
            asm volatile ("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (&afull[b])) : "memory");
        }
    } else {
        // ---------------- epilogue warps 4..7: D[b] -> registers -> maxima
        for (int it = 0; it < n_it; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int grp = tile / nchunks, chunk = tile - grp * nchunks;
            const int c0 = grp * 8, s0 = chunk * TILE_S;
            const int b = it & 1;
            mbar_wait (smem_u32 (&done[b]), (uint32_t)((it >> 1) & 1));
            asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + lane_base + 256u + 128u * b;
            uint32_t u[3][16], w[3][16];
            if (mode & 4) {
#pragma unroll
                for (int ph = 0; ph < 3; ++ph) for (int j = 0; j < 16; ++j) { u[ph][j] = 0x3f000000u + tile + j; w[ph][j] = ph; }
            } else {
#pragma unroll
            for (int ph = 0; ph < 3; ++ph) { tmem_ld16_nowait (taddr + 16 * ph, u[ph]); tmem_ld16_nowait (taddr + 48 + 16 * ph, w[ph]); }
            asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            float mx = p0buf[(it & 3) * 128 + row];
            asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (&dempty[b])) : "memory");
            const int vj = min (16, max (0, nfram - (s0 + 16 * tb)));
            if (mode & 16) { mx = fmaxf (mx, __uint_as_float (u[0][0] ^ w[2][15])); } else
#pragma unroll
            for (int ph = 0; ph < 3; ++ph)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float val = __uint_as_float (u[ph][j]) + __uint_as_float (w[ph][j]);
                    if (dbg) dbg[((size_t)tile * 128 + c8 * 16 + tb) * 48 + ph * 16 + j] = val;
                    if (j < vj) mx = fmaxf (mx, fabsf (val));
                }
            mx = fmaxf (mx, __shfl_xor_sync (0xffffffffu, mx, 4));
            mx = fmaxf (mx, __shfl_xor_sync (0xffffffffu, mx, 8));
            mx = fmaxf (mx, __shfl_xor_sync (0xffffffffu, mx, 16));
            if (lane < 4 && c0 + c8 < n_chan && mx > 0.0f) atomicMax (out_max + c0 + c8, __float_as_uint (mx));
        }
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    if (warp == 0) asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
}

static void zita_table (float* tab, unsigned hl, unsigned np, double fr)
{
    for (unsigned j = 0; j <= np; ++j) {
        double t = (double)j / (double)np;
        for (unsigned i = 0; i < hl; ++i) {
            double xs = fabs (t * fr), sc = 1.0;
            if (!(xs < 1e-6)) { xs *= M_PI; sc = sin (xs) / xs; }
            double xw = fabs (t / hl), wn = 0.0;
            if (!(xw >= 1.0)) { xw *= M_PI; wn = 0.384 + 0.500 * cos (xw) + 0.116 * cos (2 * xw); }
            tab[j * hl + (hl - i - 1)] = (float)(fr * sc * wn);
            t += 1;
        }
    }
}

int main (int argc, char** argv)
{
    const int n_chan = argc > 1 ? atoi (argv[1]) : 16, nfram = argc > 2 ? atoi (argv[2]) : 512, want_dbg = argc > 3 ? atoi (argv[3]) : 1, ver = argc > 4 ? atoi (argv[4]) : 2, mode = argc > 5 ? atoi (argv[5]) : 0;
    float tab[120]; zita_table (tab, 24, 4, 1.0);
    auto hph = [&] (int ph, int d) -> float { return d >= 24 ? tab[24 * ph + 47 - d] : tab[24 * (4 - ph) + d]; };
    // canonical B: element (n, k') at (k'/4) * B_LBO + (n/8) * 128 + (n%8) * 16 + (k'%4) * 4; n = (ph-1)*16 + j
    std::vector<float> bc (2 * B_BYTES / 4, 0.0f);
    for (int n = 0; n < NCOL; ++n) for (int k = 0; k < 64; ++k) {
        const int ph = n / 16 + 1, j = n % 16, d = j + 48 - k;
        const float c = (d >= 0 && d <= 47) ? hph (ph, d) : 0.0f;
        uint32_t u; memcpy (&u, &c, 4); u &= 0xffffe000u; float hi; memcpy (&hi, &u, 4);
        const size_t off = ((size_t)(k / 4) * B_LBO + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4) / 4;
        bc[off] = hi; bc[B_BYTES / 4 + off] = c - hi;
    }
    std::vector<float> bc2 (B2_BYTES / 4, 0.0f);
    for (int n = 0; n < NCOL; ++n) for (int k = 0; k < 64; ++k) {
        const size_t o1 = ((size_t)(k / 4) * B_LBO + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4) / 4;
        const size_t o2 = ((size_t)(k / 4) * B2_LBO + n * 16 + (k % 4) * 4) / 4;
        bc2[o2] = bc[o1]; bc2[o2 + 48 * 4] = bc[B_BYTES / 4 + o1];
    }
    const size_t stride = (size_t)nfram + 64;               // 48 history + pad, multiple of 4 when nfram is
    std::vector<float> x ((size_t)n_chan * stride);
    srand (7);
    for (auto& v : x) v = ((rand () / (float)RAND_MAX) * 2.0f - 1.0f) * 0.5f;
    float *d_x, *d_b, *d_dbg = nullptr; unsigned* d_max;
    CK (cudaMalloc (&d_x, x.size () * 4)); CK (cudaMemcpy (d_x, x.data (), x.size () * 4, cudaMemcpyHostToDevice));
    CK (cudaMalloc (&d_b, bc.size () * 4)); CK (cudaMemcpy (d_b, bc.data (), bc.size () * 4, cudaMemcpyHostToDevice));
    float* d_b2; CK (cudaMalloc (&d_b2, bc2.size () * 4)); CK (cudaMemcpy (d_b2, bc2.data (), bc2.size () * 4, cudaMemcpyHostToDevice));
    CK (cudaMalloc (&d_max, n_chan * 4)); CK (cudaMemset (d_max, 0, n_chan * 4));
    const int nchunks = (nfram + TILE_S - 1) / TILE_S, ntiles = ((n_chan + 7) / 8) * nchunks;
    if (want_dbg) { CK (cudaMalloc (&d_dbg, (size_t)ntiles * 128 * 48 * 4)); CK (cudaMemset (d_dbg, 0, (size_t)ntiles * 128 * 48 * 4)); }
    CK (cudaFuncSetAttribute (tcfir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CK (cudaFuncSetAttribute (tcfir2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V4_SMEM));
    int dev = 0, sms = 0; CK (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = ntiles < sms ? ntiles : sms;
    printf ("n_chan %d nfram %d tiles %d grid %d smem %d\n", n_chan, nfram, ntiles, grid, SMEM_BYTES);
    if (ver == 2) tcfir2_kernel<<<grid, NTHREADS, V4_SMEM>>> (d_x + 48, stride, n_chan, nfram, d_b2, d_max, d_dbg, mode); else tcfir_kernel<<<grid, 128, SMEM_BYTES>>> (d_x + 48, stride, n_chan, nfram, d_b, d_max, d_dbg);
    CK (cudaGetLastError ()); CK (cudaDeviceSynchronize ());
    std::vector<unsigned> gm (n_chan); CK (cudaMemcpy (gm.data (), d_max, n_chan * 4, cudaMemcpyDeviceToHost));
    // CPU reference (double)
    double worst = 0, worst_d = 0;
    const int ncheck = n_chan < 64 ? n_chan : 64;
    std::vector<float> dbg; if (want_dbg) { dbg.resize ((size_t)ntiles * 128 * 48); CK (cudaMemcpy (dbg.data (), d_dbg, dbg.size () * 4, cudaMemcpyDeviceToHost)); }
    for (int ch = 0; ch < ncheck; ++ch) {
        const float* xr = x.data () + (size_t)ch * stride + 48;
        double m = 0;
        for (int n = 0; n < nfram; ++n) {
            double v0 = fabs ((double)xr[n - 24]); if (v0 > m) m = v0;
            for (int ph = 1; ph < 4; ++ph) {
                double acc = 0;
                for (int d = 0; d < 48; ++d) acc += (double)hph (ph, d) * (double)xr[n - d];
                if (fabs (acc) > m) m = fabs (acc);
                if (want_dbg) {
                    const int tile = (ch / 8) * nchunks + n / TILE_S, r = (ch % 8) * 16 + (n % TILE_S) / 16, col = (ph - 1) * 16 + n % 16;
                    const double e = fabs ((double)dbg[((size_t)tile * 128 + r) * 48 + col] - acc);
                    if (e > worst_d) worst_d = e;
                }
            }
        }
        float g; memcpy (&g, &gm[ch], 4);
        const double e = fabs ((double)g - m) / m;
        if (e > worst) worst = e;
        if (ch < 3) printf ("ch %d gpu %.8f ref %.8f\n", ch, g, m);
    }
    printf ("worst relative error of the channel maxima %.3g; worst abs error of single outputs %.3g\n", worst, worst_d);
    if (!want_dbg) {
        cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
        for (int i = 0; i < 3; ++i) if (ver == 2) tcfir2_kernel<<<grid, NTHREADS, V4_SMEM>>> (d_x + 48, stride, n_chan, nfram, d_b2, d_max, nullptr, mode); else tcfir_kernel<<<grid, 128, SMEM_BYTES>>> (d_x + 48, stride, n_chan, nfram, d_b, d_max, nullptr);
        cudaEventRecord (e0);
        for (int i = 0; i < 20; ++i) if (ver == 2) tcfir2_kernel<<<grid, NTHREADS, V4_SMEM>>> (d_x + 48, stride, n_chan, nfram, d_b2, d_max, nullptr, mode); else tcfir_kernel<<<grid, 128, SMEM_BYTES>>> (d_x + 48, stride, n_chan, nfram, d_b, d_max, nullptr);
        cudaEventRecord (e1); CK (cudaDeviceSynchronize ());
        float ms; cudaEventElapsedTime (&ms, e0, e1);
        printf ("%.2f us per launch, %.1f G samples/s\n", ms * 1000 / 20, (double)n_chan * nfram / (ms / 20 * 1e-3) / 1e9);
    }
    return 0;
}
