// cor.cu — stereo phase-correlation bank.
//
// Replaces LV2M::Stcorrdsp (jmeters/stcorrdsp.cc:47-93) for N stereo pairs: five coupled one-pole
// recurrences per pair, strictly serial in time.  One lane owns one pair; a warp stages
// [32 pairs x 2 channels x 32 samples] tiles in shared memory with a 4-deep cp.async pipeline
// (separate L and R planes, row pitch 36 floats so that LDS.128 by 32 lanes is conflict free).
// Operation order is the reference's, without FMA contraction (bit-identical state).
#include <math.h>
#include "common.cuh"

namespace b200m {

constexpr int COR_T = 32, COR_P = COR_T + 4, COR_STAGES = 4;

__global__ void __launch_bounds__ (32)
cor_kernel (const float* __restrict__ in, size_t stride, int n_inst, int nfram, int aligned, float w1, float w2,
            float* __restrict__ st /* [5][n_inst] */, float* __restrict__ res)
{
    __shared__ __align__ (16) float tile[COR_STAGES][2][32 * COR_P];
    const int lane = threadIdx.x;
    const int i0 = blockIdx.x * 32;
    const int inst = min (i0 + lane, n_inst - 1);
    const bool live = (i0 + lane) < n_inst;
    const int ntiles = (nfram + COR_T - 1) / COR_T;

    auto issue = [&] (int t) {
        if (t < ntiles) {
            const int s0 = t * COR_T;
            float* d0 = tile[t % COR_STAGES][0];
            if (aligned) {
                const int c4 = (lane & 7) * 4;
                const int left = (nfram - (s0 + c4)) * 4;
                const int nb = left >= 16 ? 16 : (left > 0 ? left : 0);
#pragma unroll
                for (int i = 0; i < 16; ++i) {                 // 64 rows, 4 rows per instruction
                    const int row = 4 * i + (lane >> 3);       // 0..63 : pair = row>>1, channel = row&1
                    const int pr = min (i0 + (row >> 1), n_inst - 1);
                    const float* src = in + (size_t)(2 * pr + (row & 1)) * stride + s0 + c4;
                    cp_async16 (d0 + (row & 1) * (32 * COR_P) + (row >> 1) * COR_P + c4, nb ? src : in, nb);
                }
            } else {
#pragma unroll 4
                for (int row = 0; row < 64; ++row) {
                    const int pr = min (i0 + (row >> 1), n_inst - 1);
                    const bool ok = (s0 + lane) < nfram;
                    cp_async4 (d0 + (row & 1) * (32 * COR_P) + (row >> 1) * COR_P + lane,
                               ok ? in + (size_t)(2 * pr + (row & 1)) * stride + s0 + lane : in, ok ? 4 : 0);
                }
            }
        }
        cp_async_commit ();
    };

    float zl = st[0 * (size_t)n_inst + inst], zr = st[1 * (size_t)n_inst + inst];
    float zlr = st[2 * (size_t)n_inst + inst], zll = st[3 * (size_t)n_inst + inst], zrr = st[4 * (size_t)n_inst + inst];

    auto step = [&] (float l, float r) {                         // stcorrdsp.cc:58-62
        zl = __fadd_rn (zl, __fadd_rn (__fmul_rn (w1, __fsub_rn (l, zl)), 1e-20f));
        zr = __fadd_rn (zr, __fadd_rn (__fmul_rn (w1, __fsub_rn (r, zr)), 1e-20f));
        zlr = __fadd_rn (zlr, __fmul_rn (w2, __fsub_rn (__fmul_rn (zl, zr), zlr)));
        zll = __fadd_rn (zll, __fmul_rn (w2, __fsub_rn (__fmul_rn (zl, zl), zll)));
        zrr = __fadd_rn (zrr, __fmul_rn (w2, __fsub_rn (__fmul_rn (zr, zr), zrr)));
    };

#pragma unroll
    for (int t = 0; t < COR_STAGES - 1; ++t) issue (t);
    for (int t = 0; t < ntiles; ++t) {
        cp_async_wait<COR_STAGES - 2> ();
        __syncwarp ();
        const float* pl = tile[t % COR_STAGES][0] + lane * COR_P;
        const float* pr = tile[t % COR_STAGES][1] + lane * COR_P;
        const int len = min (COR_T, nfram - t * COR_T);
        int j = 0;
        for (; j + 4 <= len; j += 4) {
            const float4 a = *reinterpret_cast<const float4*> (pl + j);
            const float4 b = *reinterpret_cast<const float4*> (pr + j);
            step (a.x, b.x); step (a.y, b.y); step (a.z, b.z); step (a.w, b.w);
        }
        for (; j < len; ++j) step (pl[j], pr[j]);
        __syncwarp ();
        issue (t + COR_STAGES - 1);
    }
    cp_async_wait<0> ();
    // end of process(): non-finite scrub, anti-denormal bias on the three products (:65-75)
    zl = scrub (zl); zr = scrub (zr); zlr = scrub (zlr); zll = scrub (zll); zrr = scrub (zrr);
    zlr = __fadd_rn (zlr, 1e-10f); zll = __fadd_rn (zll, 1e-10f); zrr = __fadd_rn (zrr, 1e-10f);
    if (live) {
        st[0 * (size_t)n_inst + inst] = zl;  st[1 * (size_t)n_inst + inst] = zr;
        st[2 * (size_t)n_inst + inst] = zlr; st[3 * (size_t)n_inst + inst] = zll; st[4 * (size_t)n_inst + inst] = zrr;
        // Stcorrdsp::read (:79-82)
        res[inst] = __fdiv_rn (zlr, __fsqrt_rn (__fadd_rn (__fmul_rn (zll, zrr), 1e-10f)));
    }
}

}  // namespace b200m

using namespace b200m;

struct b200m_cor {
    int device; uint32_t n_inst; float w1, w2;
    float *d_st = nullptr, *d_res = nullptr;
    cudaStream_t own = nullptr; HostStage stage; bool last_host = false;
};

static cudaStream_t cor_stream (b200m_cor* h, void* stream) { return h->last_host ? h->own : (cudaStream_t)stream; }

static int cor_process (b200m_cor* h, const float* d_in, size_t stride, uint32_t nfram, cudaStream_t st)
{
    const int aligned = ((uintptr_t)d_in % 16 == 0) && (stride % 4 == 0);
    cor_kernel<<<(h->n_inst + 31) / 32, 32, 0, st>>> (d_in, stride, (int)h->n_inst, (int)nfram, aligned, h->w1, h->w2, h->d_st, h->d_res);
    B200M_LAUNCHED (1);
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

extern "C" {

int b200m_design_cor (int fsamp, float flp, float tcf, float w[2])
{
    if (!w || fsamp < 1000) return set_err (B200M_E_INVAL, "bad argument");
    w[0] = 6.28f * flp / fsamp;                    // Stcorrdsp::init (stcorrdsp.cc:85-93), int fsamp
    w[1] = 1 / (tcf * fsamp);
    return 0;
}

int b200m_cor_create (b200m_cor** out, int device, uint32_t n_inst, int fsamp, float flp, float tcf)
{
    if (!out) return set_err (B200M_E_INVAL, "NULL out pointer");
    *out = nullptr;
    if (n_inst == 0 || fsamp < 1000) return set_err (B200M_E_INVAL, "bad n_inst/fsamp");
    if (b200m_device_count () <= 0) return set_err (B200M_E_NODEVICE, "no CUDA device: b200meters has no CPU path");
    DeviceGuard g (device);
    if (!g.ok) return set_err (B200M_E_NODEVICE, "cannot select CUDA device %d", device);
    b200m_cor* h = new (std::nothrow) b200m_cor;
    if (!h) return set_err (B200M_E_NOMEM, "host allocation failed");
    h->device = device; h->n_inst = n_inst;
    { float w[2]; b200m_design_cor (fsamp, flp, tcf, w); h->w1 = w[0]; h->w2 = w[1]; }
    cudaError_t e = cudaMalloc ((void**)&h->d_st, (size_t)5 * n_inst * sizeof (float));
    if (e == cudaSuccess) e = cudaMemset (h->d_st, 0, (size_t)5 * n_inst * sizeof (float));
    if (e == cudaSuccess) e = cudaMalloc ((void**)&h->d_res, n_inst * sizeof (float));
    if (e == cudaSuccess) e = cudaMemset (h->d_res, 0, n_inst * sizeof (float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags (&h->own, cudaStreamNonBlocking);
    if (e != cudaSuccess) { int rc = cuda_fail (e, "cor_create", __FILE__, __LINE__); b200m_cor_destroy (h); return rc; }
    *out = h;
    return 0;
}

int b200m_cor_destroy (b200m_cor* h)
{
    if (!h) return 0;
    DeviceGuard g (h->device);
    cudaDeviceSynchronize ();
    cudaFree (h->d_st); cudaFree (h->d_res); h->stage.release ();
    if (h->own) cudaStreamDestroy (h->own);
    delete h;
    return 0;
}

int b200m_cor_process_device (b200m_cor* h, const float* d_in, size_t stride, uint32_t nfram, void* stream)
{
    if (int rc = check_block_args (h, d_in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    h->last_host = false;
    return cor_process (h, d_in, stride, nfram, (cudaStream_t)stream);
}

int b200m_cor_process_host (b200m_cor* h, const float* in, size_t stride, uint32_t nfram)
{
    if (int rc = check_block_args (h, in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    B200M_ENTER_HOST_PATH (h);
    if (h->stage.ensure ((size_t)2 * h->n_inst, nfram)) return set_err (B200M_E_NOMEM, "staging buffer allocation failed");
    B200M_CUDA (cudaMemcpy2DAsync (h->stage.d, h->stage.cap * sizeof (float), in, stride * sizeof (float),
                                   (size_t)nfram * sizeof (float), (size_t)2 * h->n_inst, cudaMemcpyHostToDevice, h->own));
    h->last_host = true;
    return cor_process (h, h->stage.d, h->stage.cap, nfram, h->own);
}

int b200m_cor_results (b200m_cor* h, float* out, void* stream)
{
    if (!h || !out) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = cor_stream (h, stream);
    B200M_CUDA (cudaMemcpyAsync (out, h->d_res, h->n_inst * sizeof (float), cudaMemcpyDeviceToHost, st));
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

int b200m_cor_state (b200m_cor* h, float* state5, void* stream)
{
    if (!h || !state5) return set_err (B200M_E_INVAL, "NULL argument");
    DeviceGuard g (h->device);
    cudaStream_t st = cor_stream (h, stream);
    const size_t n = h->n_inst;
    float* tmp = (float*)malloc (5 * n * sizeof (float));
    if (!tmp) return set_err (B200M_E_NOMEM, "host allocation failed");
    cudaError_t e = cudaMemcpyAsync (tmp, h->d_st, 5 * n * sizeof (float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize (st);
    if (e != cudaSuccess) { free (tmp); return cuda_fail (e, "cor_state", __FILE__, __LINE__); }
    for (size_t i = 0; i < n; ++i) for (int q = 0; q < 5; ++q) state5[5 * i + q] = tmp[q * n + i];
    free (tmp);
    return 0;
}

int b200m_cor_coeffs (const b200m_cor* h, float w[2])
{
    if (!h || !w) return set_err (B200M_E_INVAL, "NULL argument");
    w[0] = h->w1; w[1] = h->w2;
    return 0;
}

}  // extern "C"
