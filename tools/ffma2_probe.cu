// ffma2_probe.cu -- issue rate of the packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against its operand pattern, per SM sub-partition.
// Decides what the depthwise stencils (dwconv_smem.cu) can reach: acc(pair) += w(pair) * x(pair) reads three 64-bit operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ffma2_probe tools/ffma2_probe.cu && tools/_bin/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ void fma2(u64 &acc, u64 w, u64 x) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(x)); }

// MODE 0: one weight pair reused by 16 accumulators (inner r loop of the stencil)   MODE 1: a different weight pair per FMA
// MODE 2: scalar FFMA, 32 accumulators (same FMA count)                               MODE 3: weight pair = {s, s} broadcast, reused
// MODE 4: 8 accumulators only (dependent-issue distance 8)                            MODE 5: MODE 0 with x pairs also rotating (16 x, 4 w)
template <int MODE>
__global__ void probe(float *out, long long *cyc, int iters)
{
    u64 acc[16], x[16], w[16];
    float facc[32];
    for (int i = 0; i < 16; ++i) {
        acc[i] = 0;
        x[i] = ((u64)__float_as_uint(1.f + i + threadIdx.x) << 32) | __float_as_uint(0.5f * i);
        w[i] = ((u64)__float_as_uint(1e-3f * (i + 1)) << 32) | __float_as_uint(2e-3f * (i + 1));
    }
    for (int i = 0; i < 32; ++i) facc[i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 16; ++i) fma2(acc[i], w[k], x[(i + k) & 15]);
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 16; ++i) fma2(acc[i], w[(i + k) & 15], x[(i + 2 * k) & 15]);
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 32; ++i) facc[i] = fmaf(__uint_as_float((unsigned)w[k]), __uint_as_float((unsigned)x[(i + k) & 15]), facc[i]);
        } else if (MODE == 4) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) fma2(acc[i], w[k], x[(i + k) & 15]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 16; ++i) fma2(acc[i], w[k], x[(i * 3 + k) & 15]);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    for (int i = 0; i < 32; ++i) s += facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps_per_smsp, float *out, long long *cyc)
{
    const int iters = 4096, threads = warps_per_smsp * 4 * 32;
    probe<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    probe<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i];
    mean /= 148;
    const double fma_instr_per_warp = (double)iters * 64 * (MODE == 2 ? 2 : 1);   // FFMA2 (or FFMA) instructions per warp
    const double per_smsp = fma_instr_per_warp * warps_per_smsp;
    printf("%-58s warps/SMSP %d : %.2f cycles per %s per SMSP -> %.1f FMA/clk/SM (%s)\n", name, warps_per_smsp, mean / per_smsp,
           MODE == 2 ? "FFMA" : "FFMA2", (MODE == 2 ? 32.0 : 64.0) * 4 / (mean / per_smsp), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int w : {1, 2, 4}) {
        if (w == 1) { run<0>("FFMA2 w pair reused over 16 acc", 1, out, cyc); run<1>("FFMA2 distinct w pair per FMA", 1, out, cyc); run<2>("FFMA scalar, 32 acc", 1, out, cyc); run<3>("FFMA2 w = {s,s}", 1, out, cyc); run<4>("FFMA2 8 acc", 1, out, cyc); run<5>("FFMA2 x stride 3", 1, out, cyc); }
        if (w == 2) { run<0>("FFMA2 w pair reused over 16 acc", 2, out, cyc); run<1>("FFMA2 distinct w pair per FMA", 2, out, cyc); run<2>("FFMA scalar, 32 acc", 2, out, cyc); run<4>("FFMA2 8 acc", 2, out, cyc); }
        if (w == 4) { run<0>("FFMA2 w pair reused over 16 acc", 4, out, cyc); run<1>("FFMA2 distinct w pair per FMA", 4, out, cyc); run<2>("FFMA scalar, 32 acc", 4, out, cyc); run<4>("FFMA2 8 acc", 4, out, cyc); }
    }
    return 0;
}
