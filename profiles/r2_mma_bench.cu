// micro-benchmark: cycles per tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N; A from TMEM or from shared memory
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf ("CUDA error %s at %s:%d\n", cudaGetErrorString (e_), __FILE__, __LINE__); exit (1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return (uint32_t)__cvta_generic_to_shared (p); }
__device__ __forceinline__ uint64_t make_desc (uint32_t saddr, uint32_t lbo, uint32_t sbo)
{ return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46); }
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity)
{ uint32_t ok = 0; while (!ok) asm volatile ("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); }

__global__ void __launch_bounds__ (128, 1) bench (int N, int reps, int a_in_tmem, int kind16, long long* out)
{
    extern __shared__ __align__ (128) uint8_t smem[];      // zeros: A [128 x 8] and B [N x 8] tiles
    __shared__ uint32_t s_tmem; __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*> (smem)[i] = 0.0f;
    if (tid == 0) { asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32 (&bar))); asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32 (&s_tmem)), "n"(512) : "memory");
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t fmt = kind16 ? 1u : 2u;              // bf16 : tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc (smem_u32 (smem), 2048, 128), db = make_desc (smem_u32 (smem) + 8192, (uint32_t)N * 16, 128);
        const uint32_t d = tmem + 256, a = tmem;
        long long t0 = clock64 ();
        for (int i = 0; i < reps; ++i) {
            if (kind16) {
                if (a_in_tmem) asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" :: "r"(d), "r"(a), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                else asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
            } else {
                if (a_in_tmem) asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" :: "r"(d), "r"(a), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
                else asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
            }
        }
        long long t1 = clock64 ();
        asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (&bar)) : "memory");
        mbar_wait (smem_u32 (&bar), 0);
        long long t2 = clock64 ();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    if (warp == 0) asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
}
int main ()
{
    long long* d; CK (cudaMalloc (&d, 16)); long long h[2];
    CK (cudaFuncSetAttribute (bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    const int Ns[] = {16, 48, 64, 96, 128, 192, 256};
    for (int k16 = 0; k16 < 2; ++k16) for (int at = 0; at < 2; ++at) for (int N : Ns) {
        const int reps = 512;
        bench<<<1, 128, 48 * 1024>>> (N, reps, at, k16, d); CK (cudaDeviceSynchronize ());
        bench<<<1, 128, 48 * 1024>>> (N, reps, at, k16, d); CK (cudaDeviceSynchronize ());
        CK (cudaMemcpy (h, d, 16, cudaMemcpyDeviceToHost));
        const int K = k16 ? 16 : 8;
        printf ("%s A-%s N %3d: issue %.1f clk/mma, complete %.1f clk/mma = %.0f MAC/clk\n", k16 ? "bf16 K16" : "tf32 K8 ", at ? "tmem" : "smem", N, (double)h[0] / reps, (double)h[1] / reps, 128.0 * N * K / ((double)h[1] / reps));
    }
    return 0;
}
