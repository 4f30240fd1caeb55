// dwconv_smem.cu -- depthwise K^3 stencil with dilation L (5^3/L=1 and 7^3/L=3 of LKA3d_deform,
// transformerblock.py:637-638) on CUDA cores, fp32, shared-memory plane streaming.
//
// A dilated conv with dilation L only couples voxels of the same residue mod L on every axis, so the volume is
// processed as L^3 independent sub-lattices ("phases"), each a dense K^3 conv with halo (K-1)/2.
// One CTA = (batch b, 32-channel chunk, phase, lattice tile TD x TH x TW).  Input planes of the tile
// (TH+K-1) x (TW+K-1) voxels x 32 channels (128 B per voxel) stream through a cp.async double buffer; every
// thread owns 4 channels (float4) x R outputs along w x TD outputs along d in registers, so each input row loaded
// from shared memory feeds up to TD*K*R FMAs and each weight vector (broadcast across the warp) R FMAs.
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dlka {
namespace {

// tile shapes per stencil (lattice voxels): TD x TH x TW outputs per CTA, R outputs along w per thread.
// 5^3:    D,H,W = 64,128,128 divide evenly.  7^3 dil 3: the sub-lattices of the headline volume are 22 x 43 x 43, so
// 15 x 22 tiles (3 x 2 per plane) waste 7 % of the lanes where 16 x 16 wasted 20 %.
#ifndef DLKA_DS5_TD
#define DLKA_DS5_TD 4
#define DLKA_DS5_TH 16
#define DLKA_DS5_TW 16
#define DLKA_DS5_R 8
#endif
#ifndef DLKA_DS7_TD
#define DLKA_DS7_TD 2
#define DLKA_DS7_TH 15
#define DLKA_DS7_TW 22
#define DLKA_DS7_R 11
#endif
#ifndef DLKA_DS_NBUF
#define DLKA_DS_NBUF 2   // plane buffers: 2 = cp.async double buffer (1 CTA/SM); 1 = single buffer, 2 CTAs/SM overlap each other
#define DLKA_DS_MINB 1
#endif
constexpr int DS_NBUF = DLKA_DS_NBUF;
#ifndef DLKA_DS_BULK
#define DLKA_DS_BULK 0   // 1: planes + weights by cp.async.bulk (one 128-byte TMA copy per voxel): correct but measured slower (5.9 vs 5.0 ms)
#endif
constexpr bool DS_BULK = DLKA_DS_BULK != 0 && DS_NBUF == 2;
constexpr int DS_CCH = 32;                                     // channels per CTA

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int K, int L, int DS_TD, int DS_TH, int DS_TW, int DS_R>
__global__ void __launch_bounds__((DS_CCH / 4) * (DS_TW / DS_R) * DS_TH, DLKA_DS_MINB) dwconv_smem_kernel(const float *__restrict__ x, const float *__restrict__ wp,
                                                                    const float *__restrict__ bias, float *__restrict__ y,
                                                                    int C, int D, int H, int W, int tiles_d, int tiles_h, int tiles_w)
{
    constexpr int DS_THREADS = (DS_CCH / 4) * (DS_TW / DS_R) * DS_TH;
    static_assert(DS_TW % DS_R == 0 && DS_THREADS <= 1024, "tile shape");
    constexpr int P = (K - 1) / 2;
    constexpr int PH = DS_TH + K - 1, PW = DS_TW + K - 1;   // plane extent (lattice voxels)
    constexpr int PLANE_F4 = PH * PW * (DS_CCH / 4);         // float4 elements per plane
    constexpr int NPLANES = DS_TD + K - 1;
    extern __shared__ __align__(128) float4 smem4[];
    float4 *sW = smem4;                                      // [K^3][8] float4 : weights of this channel chunk
    float4 *sP = sW + K * K * K * (DS_CCH / 4);              // [NBUF][PH][PW][8] float4
    int *sOff = reinterpret_cast<int *>(sP + DS_NBUF * PLANE_F4);  // [PLANE_F4] in-plane source offsets (elements) or -1

    const int tid = threadIdx.x;
    constexpr int WRUNS = DS_TW / DS_R;
    const int q = tid & 7, wr = (tid >> 3) % WRUNS, hl = (tid >> 3) / WRUNS;
    // CTA decomposition: x = tile, y = phase * nchunks + chunk, z = batch
    int bid = blockIdx.x;
    const int tw = bid % tiles_w; bid /= tiles_w;
    const int th = bid % tiles_h; bid /= tiles_h;
    const int td = bid;
    const int nchunks = C / DS_CCH;
    const int chunk = blockIdx.y % nchunks, phase = blockIdx.y / nchunks;
    const int pw_ = phase % L, ph_ = (phase / L) % L, pd_ = phase / (L * L);
    const int b = blockIdx.z;
    const int c0 = chunk * DS_CCH;
    // lattice tile origin and this thread's outputs
    const int zd0 = td * DS_TD, zh0 = th * DS_TH, zw0 = tw * DS_TW;
    const float *xb = x + (i64)b * D * H * W * C + c0;

    uint64_t *bars = reinterpret_cast<uint64_t *>(sOff + PH * PW * 8);   // [0..1] plane buffers, [2] weights
    const uint32_t bar0 = ptx::smem_u32(bars);
    if (DS_BULK && tid == 0) {
        for (int i = 0; i < 3; ++i) ptx::mbar_init(bar0 + 8u * i, DS_THREADS);
        ptx::fence_barrier_init();
    }
    // in-plane source offsets are the same for every plane of the tile: computed once (the plane loop only adds d).
    // bulk mode: one entry per voxel (element offset of its 32-channel / 128-byte chunk); cp.async mode: one per float4.
    if (DS_BULK) {
        for (int v = tid; v < PH * PW; v += DS_THREADS) {
            const int ww = v % PW, hh = v / PW;
            const int zh = zh0 - P + hh, zw = zw0 - P + ww;
            const int hr = ph_ + L * zh, wrr = pw_ + L * zw;
            sOff[v] = (zh >= 0 && zw >= 0 && hr < H && wrr < W) ? (hr * W + wrr) * C : -1;
        }
    } else {
        for (int i = tid; i < PLANE_F4; i += DS_THREADS) {
            const int qq = i & 7, v = i >> 3;
            const int ww = v % PW, hh = v / PW;
            const int zh = zh0 - P + hh, zw = zw0 - P + ww;
            const int hr = ph_ + L * zh, wrr = pw_ + L * zw;
            sOff[i] = (zh >= 0 && zw >= 0 && hr < H && wrr < W) ? (hr * W + wrr) * C + qq * 4 : -1;
        }
    }
    __syncthreads();

    // weights of the chunk -> smem ([tap][C] packed layout in global: 128 contiguous bytes per tap)
    if (DS_BULK) {
        int nb = 0;
        for (int tap = tid; tap < K * K * K; tap += DS_THREADS) ++nb;
        ptx::mbar_arrive_expect_tx(bar0 + 16u, (uint32_t)nb * 128u);
        for (int tap = tid; tap < K * K * K; tap += DS_THREADS)
            ptx::bulk_g2s(ptx::smem_u32(sW + tap * 8), wp + (i64)tap * C + c0, 128u, bar0 + 16u);
    } else {
        for (int i = tid; i < K * K * K * (DS_CCH / 4); i += DS_THREADS) {
            const int tap = i >> 3, qq = i & 7;
            cp_async16(sW + i, wp + (i64)tap * C + c0 + qq * 4, true);   // lands with the first plane (same group)
        }
    }

    auto load_plane = [&](int s, int buf) {
        // lattice d index of plane s, real coordinate
        const int zd = zd0 - P + s;
        const int dr = pd_ + L * zd;
        const bool dok = zd >= 0 && dr < D;
        float4 *dst = sP + buf * PLANE_F4;
        const float *pb = xb + (i64)(dok ? dr : 0) * H * W * C;
        if (DS_BULK) {
            // zero-fill the out-of-volume voxels (generic stores, released by the arrive below), then one 128-byte
            // bulk copy per in-volume voxel; every thread arrives once with the bytes it is about to request
            uint32_t bytes = 0;
            for (int v = tid; v < PH * PW; v += DS_THREADS) {
                if (dok && sOff[v] >= 0) {
                    bytes += 128u;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) dst[v * 8 + e] = f4zero();
                }
            }
            ptx::fence_proxy_async();   // earlier generic reads of this buffer are ordered before the async writes
            ptx::mbar_arrive_expect_tx(bar0 + 8u * buf, bytes);
            for (int v = tid; v < PH * PW; v += DS_THREADS) {
                const int o = sOff[v];
                if (dok && o >= 0) ptx::bulk_g2s(ptx::smem_u32(dst + v * 8), pb + o, 128u, bar0 + 8u * buf);
            }
            return;
        }
#pragma unroll 4
        for (int i = tid; i < PLANE_F4; i += DS_THREADS) {
            const int o = sOff[i];
            const bool ok = dok && o >= 0;
            cp_async16(dst + i, pb + (ok ? o : 0), ok);
        }
        cp_async_commit();
    };

    float4 acc[DS_TD][DS_R];
    {
        const float4 bv = bias ? ldg4(bias + c0 + q * 4) : f4zero();
#pragma unroll
        for (int t = 0; t < DS_TD; ++t)
#pragma unroll
            for (int r = 0; r < DS_R; ++r) acc[t][r] = bv;
    }

    if (DS_NBUF == 2) load_plane(0, 0);
#pragma unroll 1
    for (int s = 0; s < NPLANES; ++s) {
        if (DS_BULK) {
            if (s + 1 < NPLANES) load_plane(s + 1, (s + 1) & 1);
            if (s == 0) ptx::mbar_wait(bar0 + 16u, 0);
            ptx::mbar_wait(bar0 + 8u * (s & 1), (s >> 1) & 1);
        } else if (DS_NBUF == 2) {
            if (s + 1 < NPLANES) {
                load_plane(s + 1, (s + 1) & 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
        } else {
            load_plane(s, 0);
            cp_async_wait<0>();
        }
        if (!DS_BULK) __syncthreads();
        const float4 *pl = sP + (DS_NBUF == 2 ? (s & 1) : 0) * PLANE_F4;
        // plane s contributes to output t with depth tap i = s - t
#pragma unroll
        for (int j = 0; j < K; ++j) {
            float4 in[DS_R + K - 1];
            const float4 *row = pl + ((hl + j) * PW + wr * DS_R) * 8 + q;
#pragma unroll
            for (int e = 0; e < DS_R + K - 1; ++e) in[e] = row[e * 8];
#pragma unroll
            for (int t = 0; t < DS_TD; ++t) {
                const int i = s - t;
                if (i < 0 || i >= K) continue;  // uniform across the CTA
                const float4 *wrow = sW + ((i * K + j) * K) * 8 + q;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float4 wv = wrow[k * 8];
#pragma unroll
                    for (int r = 0; r < DS_R; ++r) fma4v(acc[t][r], wv, in[r + k]);
                }
            }
        }
        __syncthreads();  // plane buffer (s & 1) is refilled two iterations later
    }

    // store: real coordinates of this thread's outputs
    const int hr = ph_ + L * (zh0 + hl);
    if (hr < H) {
#pragma unroll
        for (int t = 0; t < DS_TD; ++t) {
            const int dr = pd_ + L * (zd0 + t);
            if (dr >= D) continue;
#pragma unroll
            for (int r = 0; r < DS_R; ++r) {
                const int wrr = pw_ + L * (zw0 + wr * DS_R + r);
                if (wrr < W)
                    *reinterpret_cast<float4 *>(y + ((((i64)b * D + dr) * H + hr) * W + wrr) * C + c0 + q * 4) = acc[t][r];
            }
        }
    }
}

template <int K, int L, int DS_TD, int DS_TH, int DS_TW, int DS_R>
int launch_ds(const float *x, const float *wp, const float *bias, float *y, int B, int C, int D, int H, int W, cudaStream_t st)
{
    constexpr int PH = DS_TH + K - 1, PW = DS_TW + K - 1;
    const size_t smem = ((size_t)K * K * K * 8 + DS_NBUF * (size_t)PH * PW * 8) * sizeof(float4) + (size_t)PH * PW * 8 * sizeof(int) + 64;
    constexpr int DS_THREADS = (DS_CCH / 4) * (DS_TW / DS_R) * DS_TH;
    auto kern = dwconv_smem_kernel<K, L, DS_TD, DS_TH, DS_TW, DS_R>;
    static thread_local bool configured = false;
    if (!configured) {
        DLKA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // lattice extents of the largest phase
    const int ld = (int)cdiv(D, L), lh = (int)cdiv(H, L), lw = (int)cdiv(W, L);
    const int tiles_d = (int)cdiv(ld, DS_TD), tiles_h = (int)cdiv(lh, DS_TH), tiles_w = (int)cdiv(lw, DS_TW);
    dim3 grid((unsigned)(tiles_d * tiles_h * tiles_w), (unsigned)(L * L * L * (C / DS_CCH)), (unsigned)B);
    if (grid.y > 65535u || grid.z > 65535u) return DLKA_ERR_UNSUPPORTED;
    DLKA_LAUNCH(K == 5 ? "dwconv3d_smem_k5" : "dwconv3d_smem_k7d3", st,
                (kern<<<grid, DS_THREADS, smem, st>>>(x, wp, bias, y, C, D, H, W, tiles_d, tiles_h, tiles_w)));
    return DLKA_OK;
}

}  // namespace

bool dwconv_smem_supported(int C, int kd, int kh, int kw, int dil)
{
    if (C % DS_CCH != 0) return false;
    return (kd == 5 && kh == 5 && kw == 5 && dil == 1) || (kd == 7 && kh == 7 && kw == 7 && dil == 3);
}

// wp: packed [taps][C] weights (pack_dw layout)
int dwconv_smem(const float *x, const float *wp, const float *bias, float *y, int B, int C, int D, int H, int W, int k, int dil,
                cudaStream_t st)
{
    if (k == 5 && dil == 1) return launch_ds<5, 1, DLKA_DS5_TD, DLKA_DS5_TH, DLKA_DS5_TW, DLKA_DS5_R>(x, wp, bias, y, B, C, D, H, W, st);
    if (k == 7 && dil == 3) return launch_ds<7, 3, DLKA_DS7_TD, DLKA_DS7_TH, DLKA_DS7_TW, DLKA_DS7_R>(x, wp, bias, y, B, C, D, H, W, st);
    return DLKA_ERR_UNSUPPORTED;
}

}  // namespace dlka
