// ppm.cu — needle-meter ballistics bank: VU, IEC type I / II peak programme meters, BBC M/S PPM.
//
// Replaces, for N meters at once, LV2M::Vumeterdsp (jmeters/vumeterdsp.cc:45-93), LV2M::Iec1ppmdsp / Iec2ppmdsp
// (jmeters/iec1ppmdsp.cc:47-99, iec2ppmdsp.cc:47-99) and LV2M::Msppmdsp (jmeters/msppmdsp.cc:50-143) as driven by
// run() and bbcm_run() (src/meters.cc:298-331,552-589).  SURVEY.md §8(f) rank 3: the same recurrence family as the
// true-peak ballistics, without the oversampler.  One lane per meter (M/S: one lane per stereo pair, both meters),
// [32 rows x 64 samples] cp.async tiles per warp, four independent warps per CTA; operation order is the reference's,
// unfused, so every state word is bit-identical.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace b200m {

constexpr int PPM_T = 64, PPM_P = PPM_T + 4, PPM_STAGES = 3, PPM_WARPS = 4;
constexpr int PPM_PLANE = 32 * PPM_P;                      // floats of one [32 x 64] plane

struct PpmParams { float w1, w2, w3, mv_m, mv_s; };

// KIND 0: VU   1: IEC I/II PPM   3: M/S PPM (two planes: L and R rows of the pair)
template <int KIND, bool ALIGNED>
__global__ void __launch_bounds__ (PPM_WARPS * 32)
ppm_kernel (const float* __restrict__ in, size_t stride, int n_units, int nfram, PpmParams pr,
            float* __restrict__ z1s, float* __restrict__ z2s, float* __restrict__ ms, int* __restrict__ ress)
{
    constexpr int PLANES = KIND == 3 ? 2 : 1;
    extern __shared__ __align__ (16) float ppm_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u0 = (blockIdx.x * PPM_WARPS + warp) * 32;    // first unit (meter, or stereo pair) of this warp
    if (u0 >= n_units) return;
    float* ring = ppm_smem + warp * (PPM_STAGES * PLANES * PPM_PLANE);
    const int u = min (u0 + lane, n_units - 1);
    const bool live = (u0 + lane) < n_units;
    const int nproc = (nfram / 4) * 4;                     // "n /= 4": the last n mod 4 samples are ignored
    const int ntiles = (nproc + PPM_T - 1) / PPM_T;

    auto issue = [&] (int t) {
        if (t < ntiles) {
            float* dst = ring + (t % PPM_STAGES) * (PLANES * PPM_PLANE);
            const int s0 = t * PPM_T;
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl) {
                if (ALIGNED) {
                    const int c4 = (lane & 15) * 4;
                    const int left = (nproc - (s0 + c4)) * 4;
                    const int nb = left >= 16 ? 16 : (left > 0 ? left : 0);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = 2 * i + (lane >> 4);
                        const int ur = min (u0 + r, n_units - 1);
                        const float* src = in + (size_t)(PLANES * ur + pl) * stride + s0 + c4;
                        cp_async16 (dst + pl * PPM_PLANE + r * PPM_P + c4, nb ? src : in, nb);
                    }
                } else {
#pragma unroll 4
                    for (int r = 0; r < 32; ++r) {
                        const int ur = min (u0 + r, n_units - 1);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = lane + 32 * h;
                            const bool ok = (s0 + c) < nproc;
                            cp_async4 (dst + pl * PPM_PLANE + r * PPM_P + c, ok ? in + (size_t)(PLANES * ur + pl) * stride + s0 + c : in, ok ? 4 : 0);
                        }
                    }
                }
            }
        }
        cp_async_commit ();
    };

    // meter state: for M/S the pair's two meters sit at 2u (mid) and 2u + 1 (side)
    constexpr int NM = KIND == 3 ? 2 : 1;
    float z1[NM], z2[NM], m[NM];
#pragma unroll
    for (int q = 0; q < NM; ++q) {
        const int idx = NM * u + q;
        const float a = z1s[idx], b = z2s[idx];
        if (KIND == 0) { z1[q] = a > 20 ? 20 : (a < -20 ? -20 : a); z2[q] = b > 20 ? 20 : (b < -20 ? -20 : b); }   // vumeterdsp.cc:49-50
        else           { z1[q] = a > 20 ? 20 : (a < 0 ? 0 : a);     z2[q] = b > 20 ? 20 : (b < 0 ? 0 : b); }         // iec1ppmdsp.cc:51-52
        m[q] = ress[idx] ? 0.0f : ms[idx];
    }
    const float w4 = __fmul_rn (4.0f, pr.w1);

    auto group = [&] (const float4 a, const float4 b) {
        const float xa[4] = {a.x, a.y, a.z, a.w}, xb[4] = {b.x, b.y, b.z, b.w};
        if (KIND == 0) {                                   // vumeterdsp.cc:57-68
            const float t2 = __fmul_rn (z2[0], 0.5f);      // z2 / 2 (exact either way)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float t1 = __fsub_rn (fabsf (xa[i]), t2);
                z1[0] = __fadd_rn (z1[0], __fmul_rn (pr.w1, __fsub_rn (t1, z1[0])));
            }
            z2[0] = __fadd_rn (z2[0], __fmul_rn (w4, __fsub_rn (z1[0], z2[0])));
            if (z2[0] > m[0]) m[0] = z2[0];
        } else {
#pragma unroll
            for (int q = 0; q < NM; ++q) { z1[q] = __fmul_rn (z1[q], pr.w3); z2[q] = __fmul_rn (z2[q], pr.w3); }   // iec1ppmdsp.cc:59-60
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float t[NM];
                if (KIND == 1) t[0] = fabsf (xa[i]);
                else { t[0] = __fmul_rn (pr.mv_m, fabsf (__fadd_rn (xa[i], xb[i])));             // msppmdsp.cc:63,97
                       t[NM - 1] = __fmul_rn (pr.mv_s, fabsf (__fsub_rn (xa[i], xb[i]))); }
#pragma unroll
                for (int q = 0; q < NM; ++q) {
                    if (t[q] > z1[q]) z1[q] = __fadd_rn (z1[q], __fmul_rn (pr.w1, __fsub_rn (t[q], z1[q])));
                    if (t[q] > z2[q]) z2[q] = __fadd_rn (z2[q], __fmul_rn (pr.w2, __fsub_rn (t[q], z2[q])));
                }
            }
#pragma unroll
            for (int q = 0; q < NM; ++q) { const float s = __fadd_rn (z1[q], z2[q]); if (s > m[q]) m[q] = s; }
        }
    };

#pragma unroll
    for (int t = 0; t < PPM_STAGES - 1; ++t) issue (t);
    for (int t = 0; t < ntiles; ++t) {
        cp_async_wait<PPM_STAGES - 2> ();
        __syncwarp ();
        const float* base = ring + (t % PPM_STAGES) * (PLANES * PPM_PLANE) + lane * PPM_P;
        const float4* pa = reinterpret_cast<const float4*> (base);
        const float4* pb = reinterpret_cast<const float4*> (base + (PLANES - 1) * PPM_PLANE);
        const int ng = (min (PPM_T, nproc - t * PPM_T)) / 4;
        for (int q = 0; q < ng; ++q) group (pa[q], pb[q]);
        __syncwarp ();
        issue (t + PPM_STAGES - 1);
    }
    cp_async_wait<0> ();
    if (live) {
#pragma unroll
        for (int q = 0; q < NM; ++q) {
            const int idx = NM * u + q;
            if (KIND == 0) {                               // vumeterdsp.cc:70-72
                if (!finitef_ (z1[q])) { z1s[idx] = 0; m[q] = INFINITY; } else z1s[idx] = z1[q];
                if (!finitef_ (z2[q])) { z2s[idx] = 0; m[q] = INFINITY; } else z2s[idx] = __fadd_rn (z2[q], 1e-10f);
            } else {                                       // iec1ppmdsp.cc:77-78
                z1s[idx] = __fadd_rn (z1[q], 1e-10f); z2s[idx] = __fadd_rn (z2[q], 1e-10f);
            }
            ms[idx] = m[q]; ress[idx] = 0;
        }
    }
}

__global__ void ppm_read_kernel (int nm, float g, const float* __restrict__ ms, int* __restrict__ ress, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    out[i] = __fmul_rn (g, ms[i]);                          // read(): _res = true; return _g * _m
    ress[i] = 1;
}
__global__ void ppm_init_kernel (int nm, int* ress) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < nm) ress[i] = 1; }

}  // namespace b200m

using namespace b200m;

struct b200m_ppm {
    int device, kind; uint32_t n_units, n_meters; float fsamp, g, db_m, db_s;
    PpmParams pr;
    float *d_z1 = nullptr, *d_z2 = nullptr, *d_m = nullptr, *d_out = nullptr; int* d_res = nullptr;
    cudaStream_t own = nullptr; HostStage stage; bool last_host = false;
};

static void ppm_design (int kind, float fs, float w[4])
{
    if (kind == B200M_PPM_VU)        { w[0] = 11.1f / fs; w[1] = 0; w[2] = 0; w[3] = 1.5f * 1.571f; }                                  // vumeterdsp.cc:89-93
    else if (kind == B200M_PPM_IEC1) { w[0] = 450.0f / fs; w[1] = 1300.0f / fs; w[2] = 1.0f - 5.4f / fs; w[3] = 0.5108f; }            // iec1ppmdsp.cc:93-99
    else                             { w[0] = 200.0f / fs; w[1] = 860.0f / fs;  w[2] = 1.0f - 4.0f / fs; w[3] = 0.5141f; }            // iec2ppmdsp.cc:93-99, msppmdsp.cc:127-133
}

static cudaStream_t ppm_stream (b200m_ppm* h, void* stream) { return h->last_host ? h->own : (cudaStream_t)stream; }

static int ppm_process (b200m_ppm* h, const float* d_in, size_t stride, uint32_t nfram, cudaStream_t st)
{
    const bool al = ((uintptr_t)d_in % 16 == 0) && (stride % 4 == 0);
    const int planes = h->kind == B200M_PPM_MS ? 2 : 1;
    const size_t smem = (size_t)PPM_WARPS * PPM_STAGES * planes * PPM_PLANE * sizeof (float);
    const int nwarps = (h->n_units + 31) / 32;
    dim3 grid ((nwarps + PPM_WARPS - 1) / PPM_WARPS), blk (PPM_WARPS * 32);
#define PPM_GO(K, A) ppm_kernel<K, A><<<grid, blk, smem, st>>> (d_in, stride, (int)h->n_units, (int)nfram, h->pr, h->d_z1, h->d_z2, h->d_m, h->d_res)
    if (h->kind == B200M_PPM_VU) { if (al) PPM_GO (0, true); else PPM_GO (0, false); }
    else if (h->kind == B200M_PPM_MS) { if (al) PPM_GO (3, true); else PPM_GO (3, false); }
    else { if (al) PPM_GO (1, true); else PPM_GO (1, false); }
#undef PPM_GO
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

extern "C" {

int b200m_design_ppm (int kind, float fsamp, float w[4])
{
    if (!w || kind < 0 || kind > 3 || !(fsamp >= 1000.0f)) return set_err (B200M_E_INVAL, "bad argument");
    ppm_design (kind, fsamp, w);
    return 0;
}

int b200m_ppm_create (b200m_ppm** out, int device, uint32_t n_units, float fsamp, int kind)
{
    if (!out) return set_err (B200M_E_INVAL, "NULL out pointer");
    *out = nullptr;
    if (n_units == 0 || !(fsamp >= 1000.0f) || kind < 0 || kind > 3) return set_err (B200M_E_INVAL, "bad n/fsamp/kind");
    if (b200m_device_count () <= 0) return set_err (B200M_E_NODEVICE, "no CUDA device: b200meters has no CPU path");
    DeviceGuard g (device);
    if (!g.ok) return set_err (B200M_E_NODEVICE, "cannot select CUDA device %d", device);
    b200m_ppm* h = new (std::nothrow) b200m_ppm;
    if (!h) return set_err (B200M_E_NOMEM, "host allocation failed");
    h->device = device; h->kind = kind; h->n_units = n_units; h->n_meters = kind == B200M_PPM_MS ? 2 * n_units : n_units; h->fsamp = fsamp;
    float w[4]; ppm_design (kind, fsamp, w);
    h->pr.w1 = w[0]; h->pr.w2 = w[1]; h->pr.w3 = w[2]; h->g = w[3];
    h->db_m = h->db_s = 0; h->pr.mv_m = h->pr.mv_s = 1.0f;
    if (kind == B200M_PPM_MS) b200m_ppm_set_gain (h, -6, -6);          // new Msppmdsp (-6) x2, src/meters.cc:210-212
    cudaError_t e = cudaSuccess;
    auto A = [&] (void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc (p, bytes); if (e == cudaSuccess) e = cudaMemset (*p, 0, bytes); } };
    const size_t nm = h->n_meters;
    A ((void**)&h->d_z1, nm * 4); A ((void**)&h->d_z2, nm * 4); A ((void**)&h->d_m, nm * 4); A ((void**)&h->d_out, nm * 4); A ((void**)&h->d_res, nm * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags (&h->own, cudaStreamNonBlocking);
    const int smem_max = PPM_WARPS * PPM_STAGES * 2 * PPM_PLANE * (int)sizeof (float);
#define PPM_ATTR(K, A) if (e == cudaSuccess) e = cudaFuncSetAttribute (ppm_kernel<K, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)
    PPM_ATTR (0, true); PPM_ATTR (0, false); PPM_ATTR (1, true); PPM_ATTR (1, false); PPM_ATTR (3, true); PPM_ATTR (3, false);
#undef PPM_ATTR
    if (e == cudaSuccess) {
        ppm_init_kernel<<<(unsigned)((nm + 255) / 256), 256>>> ((int)nm, h->d_res);    // constructors: _res (true)
        B200M_LAUNCHED (1);
        e = cudaDeviceSynchronize ();
    }
    if (e != cudaSuccess) { int rc = cuda_fail (e, "ppm_create", __FILE__, __LINE__); b200m_ppm_destroy (h); return rc; }
    *out = h;
    return 0;
}

int b200m_ppm_destroy (b200m_ppm* h)
{
    if (!h) return 0;
    DeviceGuard g (h->device);
    cudaDeviceSynchronize ();
    cudaFree (h->d_z1); cudaFree (h->d_z2); cudaFree (h->d_m); cudaFree (h->d_out); cudaFree (h->d_res); h->stage.release ();
    if (h->own) cudaStreamDestroy (h->own);
    delete h;
    return 0;
}

int b200m_ppm_set_gain (b200m_ppm* h, float db_m, float db_s)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    if (h->kind != B200M_PPM_MS) return set_err (B200M_E_INVAL, "set_gain applies to the M/S PPM bank");
    // Msppmdsp::set_gain (msppmdsp.cc:135-143): _mv = powf (10, .05 * db), skipped when db is unchanged
    if (h->db_m != db_m) { h->db_m = db_m; h->pr.mv_m = powf (10, .05 * db_m); }
    if (h->db_s != db_s) { h->db_s = db_s; h->pr.mv_s = powf (10, .05 * db_s); }
    return 0;
}

int b200m_ppm_process_device (b200m_ppm* h, const float* d_in, size_t stride, uint32_t nfram, void* stream)
{
    if (int rc = check_block_args (h, d_in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    h->last_host = false;
    return ppm_process (h, d_in, stride, nfram, (cudaStream_t)stream);
}

int b200m_ppm_process_host (b200m_ppm* h, const float* in, size_t stride, uint32_t nfram)
{
    if (int rc = check_block_args (h, in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    B200M_ENTER_HOST_PATH (h);
    const size_t rows = h->kind == B200M_PPM_MS ? (size_t)2 * h->n_units : h->n_units;
    if (h->stage.ensure (rows, nfram)) return set_err (B200M_E_NOMEM, "staging buffer allocation failed");
    B200M_CUDA (cudaMemcpy2DAsync (h->stage.d, h->stage.cap * sizeof (float), in, stride * sizeof (float),
                                   (size_t)nfram * sizeof (float), rows, cudaMemcpyHostToDevice, h->own));
    h->last_host = true;
    return ppm_process (h, h->stage.d, h->stage.cap, nfram, h->own);
}

int b200m_ppm_read_device (b200m_ppm* h, void* stream)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    DeviceGuard g (h->device);
    ppm_read_kernel<<<(h->n_meters + 255) / 256, 256, 0, ppm_stream (h, stream)>>> ((int)h->n_meters, h->g, h->d_m, h->d_res, h->d_out);
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

int b200m_ppm_results (b200m_ppm* h, float* out, void* stream)
{
    if (!h || !out) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ppm_stream (h, stream);
    B200M_CUDA (cudaMemcpyAsync (out, h->d_out, h->n_meters * sizeof (float), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

int b200m_ppm_state (b200m_ppm* h, float* state4, void* stream)
{
    if (!h || !state4) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = ppm_stream (h, stream);
    const size_t nm = h->n_meters;
    float* tmp = (float*)malloc (4 * nm * sizeof (float));
    if (!tmp) return set_err (B200M_E_NOMEM, "host allocation failed");
    const void* src[4] = {h->d_z1, h->d_z2, h->d_m, h->d_res};
    cudaError_t e = cudaSuccess;
    for (int q = 0; q < 4 && e == cudaSuccess; ++q) e = cudaMemcpyAsync (tmp + q * nm, src[q], nm * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize (st);
    if (e != cudaSuccess) { free (tmp); return cuda_fail (e, "ppm_state", __FILE__, __LINE__); }
    for (size_t i = 0; i < nm; ++i) {
        state4[4 * i] = tmp[i]; state4[4 * i + 1] = tmp[nm + i]; state4[4 * i + 2] = tmp[2 * nm + i];
        state4[4 * i + 3] = (float)((const int*)(tmp + 3 * nm))[i];
    }
    free (tmp);
    return 0;
}

}  // extern "C"
