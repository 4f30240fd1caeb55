// l1_probe.cu -- micro-benchmark: how many cycles does one SM need per 128-byte line for L1-hit global loads of
// different widths, against the same lines served from shared memory?  Decides the gather path of the deformable conv.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/l1_probe tools/l1_probe.cu && tools/_bin/l1_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int THREADS = 512, ITERS = 2048, NOFF = 8;
constexpr int BUF_LINES = 512;  // 64 KB per SM: L1 resident

// mode 0: LDG.128, 8 lanes per line (4 lines / instr)     mode 1: LDG.32, 32 lanes one line (1 line / instr)
// mode 2: LDG.64, 16 lanes per line (2 lines / instr)     mode 3: LDS.128, 8 lanes per line   mode 4: LDS.32 one line
// mode 5: LDG.128 all 32 lanes within ONE 512-byte span (4 consecutive lines)
// mode 6: LDS.128 warp-uniform address (broadcast)        mode 7: LDS.128, 4 distinct 16-byte addresses (one per 8 lanes)
// mode 8: LDG.256 (ld.global.nc.v8.f32), 4 lanes per line (8 random lines / instr)
// mode 9: LDG.256, 8 lanes per 256-byte pair of ADJACENT lines (4 random pairs / instr: the x0 / x1 corners of a sample
//         when the volume is laid out [chunk][d][h][w][32 ch])
// mode 10: LDG.128, 16 lanes per adjacent line pair (2 random pairs / instr)
template <int MODE>
__global__ void __launch_bounds__(THREADS) probe(const float *__restrict__ g, long long *cycles, float *sink)
{
    extern __shared__ float4 sm4[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *base = g + (size_t)blockIdx.x * BUF_LINES * 32;
    float *sm = reinterpret_cast<float *>(sm4);
    for (int i = tid; i < BUF_LINES * 32; i += THREADS) sm[i] = base[i];
    __syncthreads();
    // per-thread line offsets (floats), pseudo random, fixed in registers
    int off[NOFF];
    unsigned h = (warp * 977u + 13u) * 2654435761u;
#pragma unroll
    for (int i = 0; i < NOFF; ++i) {
        h = h * 1664525u + 1013904223u;
        unsigned grp = MODE == 0 || MODE == 3 ? lane >> 3 : (MODE == 2 ? lane >> 4 : 0);
        unsigned line = ((h >> 8) + grp * 37u) % BUF_LINES;
        if (MODE == 5) line = ((h >> 8) % (BUF_LINES / 4)) * 4 + (lane >> 3);
        if (MODE == 6) line = (h >> 8) % BUF_LINES;
        if (MODE == 7) line = ((h >> 8) + (lane >> 3) * 37u) % BUF_LINES;
        int within = MODE == 0 || MODE == 3 || MODE == 5 ? (lane & 7) * 4 : (MODE == 2 ? (lane & 15) * 2 : lane);
        if (MODE == 6 || MODE == 7) within = 4 * ((h >> 4) & 7);
        if (MODE == 8) { line = ((h >> 8) + (lane >> 2) * 37u) % BUF_LINES; within = (lane & 3) * 8; }
        if (MODE == 9) { line = (((h >> 8) + (lane >> 3) * 37u) % (BUF_LINES / 2)) * 2; within = (lane & 7) * 8; }
        if (MODE == 10) { line = (((h >> 8) + (lane >> 4) * 37u) % (BUF_LINES / 2)) * 2; within = (lane & 15) * 4; }
        off[i] = line * 32 + within;
    }
    // warm L1
    float acc = 0.f;
    for (int i = tid; i < BUF_LINES * 32; i += THREADS) acc += __ldg(base + i);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NOFF; ++i) {
            const int o = (off[i] + it * 32 * (MODE == 9 || MODE == 10 ? 6 : 5)) & (BUF_LINES * 32 - 1);
            if (MODE == 8 || MODE == 9) {
                float a, b, c, d, e, f, g2, h2;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g2), "=f"(h2) : "l"(base + o));
                acc += ((a + b) + (c + d)) + ((e + f) + (g2 + h2));
            } else if (MODE == 0 || MODE == 5 || MODE == 10) { const float4 v = __ldg(reinterpret_cast<const float4 *>(base + o)); acc += (v.x + v.y) + (v.z + v.w); }
            else if (MODE == 1) acc += __ldg(base + o);
            else if (MODE == 2) { const float2 v = __ldg(reinterpret_cast<const float2 *>(base + o)); acc += v.x + v.y; }
            else if (MODE == 3 || MODE == 6 || MODE == 7) { const float4 v = *reinterpret_cast<const float4 *>(sm + o); acc += (v.x + v.y) + (v.z + v.w); }
            else acc += sm[o];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * THREADS + tid] = acc;
}

template <int MODE>
void run(const char *name, int lines_per_instr, const float *g, long long *cyc, float *sink, int nsm)
{
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, BUF_LINES * 128);
    probe<MODE><<<nsm, THREADS, BUF_LINES * 128>>>(g, cyc, sink);
    cudaDeviceSynchronize();
    probe<MODE><<<nsm, THREADS, BUF_LINES * 128>>>(g, cyc, sink);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < nsm; ++i) mean += (double)h[i];
    mean /= nsm;
    const double instr = (double)(THREADS / 32) * ITERS * NOFF;
    printf("%-44s %8.2f cyc/warp-instr  %6.2f cyc/line  %6.1f B/clk/SM  (%s)\n", name, mean / instr, mean / (instr * lines_per_instr),
           instr * lines_per_instr * 128.0 / mean, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float *g, *sink;
    long long *cyc;
    cudaMalloc(&g, (size_t)nsm * BUF_LINES * 128);
    cudaMemset(g, 0, (size_t)nsm * BUF_LINES * 128);
    cudaMalloc(&sink, (size_t)nsm * THREADS * 4);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    run<0>("LDG.128  8 lanes/line, 4 random lines", 4, g, cyc, sink, nsm);
    run<5>("LDG.128  4 consecutive lines", 4, g, cyc, sink, nsm);
    run<10>("LDG.128 16 lanes / adjacent line pair, 2 pairs", 4, g, cyc, sink, nsm);
    run<8>("LDG.256  4 lanes/line, 8 random lines", 8, g, cyc, sink, nsm);
    run<9>("LDG.256  8 lanes / adjacent line pair, 4 pairs", 8, g, cyc, sink, nsm);
    run<2>("LDG.64  16 lanes/line, 2 random lines", 2, g, cyc, sink, nsm);
    run<1>("LDG.32  32 lanes, 1 line", 1, g, cyc, sink, nsm);
    run<3>("LDS.128  8 lanes/line, 4 random lines", 4, g, cyc, sink, nsm);
    run<4>("LDS.32  32 lanes, 1 line", 1, g, cyc, sink, nsm);
    run<6>("LDS.128 warp-uniform address", 1, g, cyc, sink, nsm);
    run<7>("LDS.128 4 addresses, 8-lane broadcast", 1, g, cyc, sink, nsm);
    return 0;
}
