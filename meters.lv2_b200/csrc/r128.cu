// r128.cu — the EBUr128 plugin's audio cycle for N stereo instances: EBU R128 loudness + optional dBTP.
//
// Mirrors what ebur128_run does between its atom parsing and atom forging (src/ebulv2.cc:341-367):
//   ebu->process (n, {inL, inR});  if (dbtp_enable) { mtr[0]->process_max (inL); mtr[1]->process_max (inR); }
//   lm/mm/ls/ms/il/rn/rx getters;  tp = coef_to_db (max (mtr[0]->read (), mtr[1]->read ()));  tp_max = max (tp_max, tp)
// It composes the EBU bank (ebu.cu) and the true-peak bank (tpk.cu) over ONE host->device copy of the block.
#include <math.h>
#include "common.cuh"

namespace b200m {

// coef_to_db (src/ebulv2.cc:227-230) and the tp_max hold (:360-367); one thread per instance
__global__ void r128_tp_kernel (int n_inst, const float* __restrict__ tp_m, int* __restrict__ tp_res, float* __restrict__ tp_max)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst) return;
    const float a = tp_m[2 * i], b = tp_m[2 * i + 1];           // TruePeakdsp::read(): returns _m, sets _res
    tp_res[2 * i] = 1; tp_res[2 * i + 1] = 1;
    const float v = a > b ? a : b;
    const float tp = (v == 0) ? -INFINITY : __double2float_rn (__dmul_rn (20.0, (double)log10f_glibc (v)));
    if (tp > tp_max[i]) tp_max[i] = tp;
}
__global__ void r128_fill_kernel (int n, float* p, float v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// accessors implemented in tpk.cu (internal linkage across the library, hidden visibility)
void tpk_raw_pointers (b200m_tpk* h, float** tp_m, int** tp_res);

}  // namespace b200m

using namespace b200m;

struct b200m_r128 {
    int device; uint32_t n_inst; int dbtp;
    b200m_ebu* ebu = nullptr; b200m_tpk* tpk = nullptr;
    float* d_tpmax = nullptr;
    cudaStream_t own = nullptr; HostStage stage; bool last_host = false;
};

static int r128_run (b200m_r128* h, const float* d_in, size_t stride, uint32_t nfram, cudaStream_t st)
{
    if (int rc = b200m_ebu_process_device (h->ebu, d_in, stride, nfram, st)) return rc;
    if (h->dbtp) {
        if (int rc = b200m_tpk_process_device (h->tpk, d_in, stride, nfram, B200M_TP_MODE_MAX, st)) return rc;
        float* tp_m; int* tp_res;
        tpk_raw_pointers (h->tpk, &tp_m, &tp_res);
        r128_tp_kernel<<<(h->n_inst + 255) / 256, 256, 0, st>>> ((int)h->n_inst, tp_m, tp_res, h->d_tpmax);
        B200M_LAUNCHED (1);
    } else {
        r128_fill_kernel<<<(h->n_inst + 255) / 256, 256, 0, st>>> ((int)h->n_inst, h->d_tpmax, -INFINITY);   // :365-366
        B200M_LAUNCHED (1);
    }
    B200M_CUDA (cudaGetLastError ());
    return 0;
}

extern "C" {

int b200m_r128_create (b200m_r128** out, int device, uint32_t n_inst, float fsamp, int dbtp_enable)
{
    if (!out) return set_err (B200M_E_INVAL, "NULL out pointer");
    *out = nullptr;
    b200m_r128* h = new (std::nothrow) b200m_r128;
    if (!h) return set_err (B200M_E_NOMEM, "host allocation failed");
    h->device = device; h->n_inst = n_inst; h->dbtp = dbtp_enable ? 1 : 0;
    int rc = b200m_ebu_create (&h->ebu, device, n_inst, 2, fsamp);                 // ebu->init (2, rate), src/ebulv2.cc:190
    if (!rc) rc = b200m_tpk_create (&h->tpk, device, 2 * n_inst, fsamp, B200M_TPK_TRUEPEAK);   // 2 x TruePeakdsp, :192-196
    if (!rc) {
        DeviceGuard g (device);
        cudaError_t e = cudaMalloc ((void**)&h->d_tpmax, n_inst * sizeof (float));
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags (&h->own, cudaStreamNonBlocking);
        if (e == cudaSuccess) {
            r128_fill_kernel<<<(n_inst + 255) / 256, 256>>> ((int)n_inst, h->d_tpmax, -INFINITY);
            B200M_LAUNCHED (1);
            e = cudaDeviceSynchronize ();
        }
        if (e != cudaSuccess) rc = cuda_fail (e, "r128_create", __FILE__, __LINE__);
    }
    if (rc) { b200m_r128_destroy (h); return rc; }
    *out = h;
    return 0;
}

int b200m_r128_destroy (b200m_r128* h)
{
    if (!h) return 0;
    b200m_ebu_destroy (h->ebu); b200m_tpk_destroy (h->tpk);
    DeviceGuard g (h->device);
    cudaFree (h->d_tpmax); h->stage.release ();
    if (h->own) cudaStreamDestroy (h->own);
    delete h;
    return 0;
}

int b200m_r128_control (b200m_r128* h, int32_t inst, int cmd, void* stream)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    void* st = h->last_host ? (void*)h->own : stream;
    switch (cmd) {
    case B200M_R128_START: return b200m_ebu_integr_start (h->ebu, inst, st);
    case B200M_R128_PAUSE: return b200m_ebu_integr_pause (h->ebu, inst, st);
    case B200M_R128_RESET: return b200m_ebu_integr_reset (h->ebu, inst, st);
    default: return set_err (B200M_E_INVAL, "unknown control %d", cmd);
    }
}

int b200m_r128_run_device (b200m_r128* h, const float* d_in, size_t stride, uint32_t nfram, void* stream)
{
    if (int rc = check_block_args (h, d_in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    h->last_host = false;
    return r128_run (h, d_in, stride, nfram, (cudaStream_t)stream);
}

int b200m_r128_run_host (b200m_r128* h, const float* in, size_t stride, uint32_t nfram)
{
    if (int rc = check_block_args (h, in, stride, nfram)) return rc;
    DeviceGuard g (h->device);
    const size_t nch = (size_t)2 * h->n_inst;
    if (h->stage.ensure (nch, nfram)) return set_err (B200M_E_NOMEM, "staging buffer allocation failed");
    B200M_CUDA (cudaMemcpy2DAsync (h->stage.d, h->stage.cap * sizeof (float), in, stride * sizeof (float),
                                   (size_t)nfram * sizeof (float), nch, cudaMemcpyHostToDevice, h->own));
    h->last_host = true;
    return r128_run (h, h->stage.d, h->stage.cap, nfram, h->own);
}

int b200m_r128_results (b200m_r128* h, b200m_ebu_result* ebu_out, float* tp_max_db, void* stream)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    DeviceGuard g (h->device);
    cudaStream_t st = h->last_host ? h->own : (cudaStream_t)stream;
    if (tp_max_db) B200M_CUDA (cudaMemcpyAsync (tp_max_db, h->d_tpmax, h->n_inst * sizeof (float), cudaMemcpyDeviceToHost, st));
    if (ebu_out) return b200m_ebu_results (h->ebu, ebu_out, st);
    B200M_CUDA (cudaStreamSynchronize (st));
    return 0;
}

b200m_ebu* b200m_r128_ebu (b200m_r128* h) { return h ? h->ebu : nullptr; }
b200m_tpk* b200m_r128_tpk (b200m_r128* h) { return h ? h->tpk : nullptr; }

}  // extern "C"
