// micro-probe: cycles per input sample of the PPM ballistics chain, one warp per SM-resident block
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void k (const float4* in, float* out, long long* cyc, int n, float w3, float wf0, float wf1)
{
    __shared__ float4 tile[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) tile[i] = in[i];
    __syncwarp ();
    const int lane = threadIdx.x;
    const float wf = lane >> 4 ? wf1 : wf0;
    float z = 0.001f * lane, m = 0, p = 0;
    const long long t0 = clock64 ();
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
        const float4 v4 = tile[j & 1023];
        z = __fmul_rn (z, w3);
        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = vv[i];
            if (V == 0) { if (v > z) z = __fadd_rn (z, __fmul_rn (wf, __fsub_rn (v, z))); }
            else z = __fadd_rn (z, __fmul_rn (wf, fmaxf (__fsub_rn (v, z), 0.0f)));
            if (V != 3) p = fmaxf (p, v);
        }
        if (V != 2 && V != 3) { const float tt = __fadd_rn (z, __shfl_xor_sync (0xffffffffu, z, 16)); if (tt > m) m = tt; }
    }
    const long long t1 = clock64 ();
    out[blockIdx.x * 32 + lane] = z + m + p;
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}
int main ()
{
    float4* in; float* out; long long* cyc;
    cudaMalloc (&in, 1024 * 16); cudaMalloc (&out, 4096 * 4); cudaMallocManaged (&cyc, 1024 * 8);
    float4 h[1024]; for (int i = 0; i < 1024; ++i) h[i] = make_float4 (0.1f + 0.001f * (i % 97), 0.2f, 0.05f + 0.002f * (i % 31), 0.3f);
    cudaMemcpy (in, h, sizeof (h), cudaMemcpyHostToDevice);
    const int n = 8192;
    for (int blocks : {1, 148, 148 * 7}) {
#define RUN(V) k<V><<<blocks, 32>>> (in, out, cyc, n, 0.99996f, 0.0208f, 0.0896f); cudaDeviceSynchronize (); printf ("blocks %4d variant %d: %.1f cycles per sample\n", blocks, V, (double)cyc[0] / n);
        RUN (0) RUN (1) RUN (2) RUN (3)
    }
    return 0;
}
