// host_common.cu — error reporting, launch accounting and pinned-memory helpers of the C ABI.
#include <stdarg.h>
#include "common.cuh"

namespace b200m {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int set_err (int code, const char* fmt, ...)
{
    va_list ap; va_start (ap, fmt);
    vsnprintf (g_err, sizeof (g_err), fmt, ap);
    va_end (ap);
    return code;
}

int cuda_fail (cudaError_t e, const char* what, const char* file, int line)
{
    const char* base = strrchr (file, '/');
    snprintf (g_err, sizeof (g_err), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString (e), what, base ? base + 1 : file, line);
    cudaGetLastError ();   // clear the sticky-free error state
    return (e == cudaErrorMemoryAllocation) ? B200M_E_NOMEM
         : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? B200M_E_NODEVICE : B200M_E_CUDA;
}

int check_block_args (const void* h, const void* in, size_t stride, uint32_t nfram)
{
    if (!h) return set_err (B200M_E_INVAL, "NULL handle");
    if (!in) return set_err (B200M_E_INVAL, "NULL input");
    if (nfram == 0 || nfram > B200M_MAX_BLOCK) return set_err (B200M_E_INVAL, "nfram %u outside 1..%u", nfram, B200M_MAX_BLOCK);
    if (stride < nfram) return set_err (B200M_E_INVAL, "stride %zu < nfram %u", stride, nfram);
    return 0;
}

}  // namespace b200m

using namespace b200m;

extern "C" {

int b200m_abi_version (void) { return B200M_ABI_VERSION; }
const char* b200m_last_error (void) { return g_err; }

int b200m_device_count (void)
{
    int n = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
    return n;
}

int b200m_host_alloc (void** p, size_t bytes)
{
    if (!p) return set_err (B200M_E_INVAL, "NULL out pointer");
    B200M_CUDA (cudaHostAlloc (p, bytes, cudaHostAllocDefault));
    return 0;
}

int b200m_host_free (void* p)
{
    if (!p) return 0;
    B200M_CUDA (cudaFreeHost (p));
    return 0;
}

uint64_t b200m_launch_count (void) { return g_launches.load (); }

}
