// dwconv.cu -- depthwise large-kernel stencils (5^3 pad 2; 7^3 dilation 3 pad 9) and the depthwise
// deformable convolution, channels-last fp32 on CUDA cores.
//
// Regular depthwise conv (LKA3d_deform.conv0 / conv_spatial, transformerblock.py:637-638):
//   lanes run over 4-channel chunks (coalesced 16 B per lane), each thread keeps R outputs along W
//   spaced by the dilation so that every loaded input feeds up to K FMAs from registers.
#include "kernels.cuh"

namespace dlka {
namespace {

// [C][taps] -> [taps][C]
__global__ void pack_dw_weight_kernel(const float *__restrict__ w, float *__restrict__ wp, int C, int taps)
{
    const int total = C * taps;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % C, t = i / C;
        wp[i] = w[(i64)c * taps + t];
    }
}

template <int KS, int DIL, int R>
__global__ void __launch_bounds__(256) dwconv3d_cl_kernel(const float *__restrict__ x, const float *__restrict__ wp,
                                                          const float *__restrict__ bias, float *__restrict__ y, int C,
                                                          int D, int H, int W, int kd_eff)
{
    // kd_eff: depth taps actually present (KS for 3D, 1 when D axis is degenerate 2D data with kd=1)
    constexpr int PAD = DIL * (KS - 1) / 2;
    const int c4 = threadIdx.x * 4;
    const int tw = blockIdx.x * blockDim.y + threadIdx.y;
    const int wblk = tw / DIL, phase = tw % DIL;
    const int w0 = wblk * DIL * R + phase;
    const int h = blockIdx.y;
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (w0 >= W) return;

    float4 acc[R];
    const float4 bv = bias ? ldg4(bias + c4) : f4zero();
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = bv;

    const int pad_d = kd_eff == 1 ? 0 : PAD;
    for (int i = 0; i < kd_eff; ++i) {
        const int din = d - pad_d + DIL * i;
        if ((unsigned)din >= (unsigned)D) continue;
        for (int j = 0; j < KS; ++j) {
            const int hin = h - PAD + DIL * j;
            if ((unsigned)hin >= (unsigned)H) continue;
            const float *row = x + ((((i64)b * D + din) * H + hin) * W) * (i64)C + c4;
            const float *wrow = wp + (i64)((i * KS + j) * KS) * C + c4;
            float4 in[R + KS - 1];
#pragma unroll
            for (int q = 0; q < R + KS - 1; ++q) {
                const int win = w0 - PAD + DIL * q;
                in[q] = ((unsigned)win < (unsigned)W) ? ldg4(row + (i64)win * C) : f4zero();
            }
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const float4 wv = ldg4(wrow + (i64)kk * C);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc[r].x = fmaf(wv.x, in[r + kk].x, acc[r].x);
                    acc[r].y = fmaf(wv.y, in[r + kk].y, acc[r].y);
                    acc[r].z = fmaf(wv.z, in[r + kk].z, acc[r].z);
                    acc[r].w = fmaf(wv.w, in[r + kk].w, acc[r].w);
                }
            }
        }
    }
    float *orow = y + ((((i64)b * D + d) * H + h) * W) * (i64)C + c4;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int wo = w0 + DIL * r;
        if (wo < W) *reinterpret_cast<float4 *>(orow + (i64)wo * C) = acc[r];
    }
}

// any odd kernel (kd, kh, kw) with per-axis dilation, "same" extent: thread = (voxel, 4-channel chunk).  Fallback for the
// stencil shapes / channel counts the shared-memory kernel has no instance for.
__global__ void __launch_bounds__(256) dwconv3d_generic_kernel(const float *__restrict__ x, const float *__restrict__ wp,
                                                               const float *__restrict__ bias, float *__restrict__ y, int B, int C,
                                                               int D, int H, int W, int kd, int kh, int kw, int dd, int dh, int dw, i64 ych,
                                                               int yldv)
{
    // output element (voxel, channel c) at y + (c / 32) * ych + voxel * yldv + c % 32: channels-last (ych = 32, yldv = C) or the
    // chunk-major gather layout of deform_ps.cu (ych = voxels * 32, yldv = 32)
    const int C4 = C / 4;
    const i64 total = (i64)B * D * H * W * C4;
    const int pd = dd * (kd - 1) / 2, ph = dh * (kh - 1) / 2, pw = dw * (kw - 1) / 2;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        i64 v = i / C4;
        const i64 voxel = v;
        const int w = (int)(v % W); v /= W;
        const int h = (int)(v % H); v /= H;
        const int d = (int)(v % D);
        const int b = (int)(v / D);
        float4 acc = bias ? ldg4(bias + c) : f4zero();
        for (int ii = 0; ii < kd; ++ii) {
            const int din = d - pd + dd * ii;
            if ((unsigned)din >= (unsigned)D) continue;
            for (int jj = 0; jj < kh; ++jj) {
                const int hin = h - ph + dh * jj;
                if ((unsigned)hin >= (unsigned)H) continue;
                for (int kk = 0; kk < kw; ++kk) {
                    const int win = w - pw + dw * kk;
                    if ((unsigned)win >= (unsigned)W) continue;
                    fma4v(acc, ldg4(wp + (i64)((ii * kh + jj) * kw + kk) * C + c),
                          ldg4(x + ((((i64)b * D + din) * H + hin) * W + win) * C + c));
                }
            }
        }
        *reinterpret_cast<float4 *>(y + (i64)(c >> 5) * ych + voxel * yldv + (c & 31)) = acc;
    }
}

// depthwise deformable conv: thread = (row, 4-channel chunk)
template <int NDIM>
__global__ void __launch_bounds__(256) deform_dwconv_cl_kernel(const float *__restrict__ x, const float *__restrict__ off,
                                                               const float *__restrict__ mask, const float *__restrict__ wp,
                                                               const float *__restrict__ bias, float *__restrict__ y,
                                                               const ConvGeo g, i64 M)
{
    const int C4 = g.C / 4;
    const i64 total = M * C4;
    const int cpg = g.C / g.dg;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C4) * 4;
        const i64 m = idx / C4;
        const int wo = (int)(m % g.Wo);
        i64 t = m / g.Wo;
        const int ho = (int)(t % g.Ho);
        t /= g.Ho;
        const int d_o = (int)(t % g.Do), b = (int)(t / g.Do);
        const int dgi = c / cpg;
        const float *vol = x + (i64)b * g.D * g.H * g.W * g.C + c;
        const float *o = off + m * (i64)(g.dg * NDIM * g.K) + (i64)dgi * g.K * NDIM;
        float4 acc = bias ? ldg4(bias + c) : f4zero();
        int tap = 0;
        for (int ii = 0; ii < g.kd; ++ii)
            for (int jj = 0; jj < g.kh; ++jj)
                for (int kk = 0; kk < g.kw; ++kk, ++tap) {
                    float4 v;
                    if (NDIM == 3) {
                        const float pd = sample_pos(d_o, g.sd, g.pd, ii, g.dd, __ldg(o + tap * 3 + 0));
                        const float ph = sample_pos(ho, g.sh, g.ph, jj, g.dh, __ldg(o + tap * 3 + 1));
                        const float pw = sample_pos(wo, g.sw, g.pw, kk, g.dw, __ldg(o + tap * 3 + 2));
                        v = trilinear4(vol, make_sample3(pd, ph, pw, g.D, g.H, g.W), g.H, g.W, g.C);
                    } else {
                        const float ph = sample_pos(ho, g.sh, g.ph, jj, g.dh, __ldg(o + tap * 2 + 0));
                        const float pw = sample_pos(wo, g.sw, g.pw, kk, g.dw, __ldg(o + tap * 2 + 1));
                        v = bilinear4(vol, make_sample2(ph, pw, g.H, g.W), g.W, g.C);
                        if (mask) {
                            const float mk = __ldg(mask + m * (i64)(g.dg * g.K) + dgi * g.K + tap);
                            v.x *= mk; v.y *= mk; v.z *= mk; v.w *= mk;
                        }
                    }
                    const float4 wv = ldg4(wp + (i64)tap * g.C + c);
                    acc.x = fmaf(wv.x, v.x, acc.x); acc.y = fmaf(wv.y, v.y, acc.y);
                    acc.z = fmaf(wv.z, v.z, acc.z); acc.w = fmaf(wv.w, v.w, acc.w);
                }
        *reinterpret_cast<float4 *>(y + m * (i64)g.C + c) = acc;
    }
}

#ifndef DLKA_DW2D_UNROLL
#define DLKA_DW2D_UNROLL 7
#endif
constexpr int DW2D_UNROLL = DLKA_DW2D_UNROLL;   // taps in flight per thread, 4 corner loads each (2 -> 7: 560 -> 512 us at 24x96x56x56)
// 2D depthwise deformable conv with SHARED sampling parameters: in the kernel above every thread of a pixel (C/4 of them) re-derives
// the same sampling position, validity and bilinear weights for every tap -- more instructions than the gather and the blend
// themselves.  Here a block owns PB consecutive pixels: phase 1 computes one 32-byte record per (pixel, offset group, tap)
// {4 clamped corner offsets, 4 corner weights with validity and the DCNv2 mask folded in} into shared memory (rules of
// torchvision's bilinear_interpolate through make_sample2), phase 2 runs thread = (pixel, 4 channels) over the taps with two
// broadcast LDS.128, four LDG.128 and the blend.
__global__ void __launch_bounds__(256) deform_dwconv2d_shared_kernel(const float *__restrict__ x, const float *__restrict__ off,
                                                                     const float *__restrict__ mask, const float *__restrict__ wp,
                                                                     const float *__restrict__ bias, float *__restrict__ y,
                                                                     const ConvGeo g, i64 M, int PB)
{
    extern __shared__ __align__(16) uint8_t dsm[];
    const int K = g.K, dg = g.dg, C = g.C, C4 = C / 4, cpg = C / dg;
    const int nrec = PB * dg * K;
    int4 *sO = reinterpret_cast<int4 *>(dsm);
    float4 *sW = reinterpret_cast<float4 *>(dsm) + nrec;
    const i64 ngroups = (M + PB - 1) / PB;
    for (i64 grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const i64 m0 = grp * PB;
        for (int e = threadIdx.x; e < nrec; e += blockDim.x) {
            const int tap = e % K, dgi = (e / K) % dg, p = e / (K * dg);
            const i64 m = m0 + p;
            int4 o = make_int4(0, 0, 0, 0);
            float4 w = f4zero();
            if (m < M) {
                const int wo = (int)(m % g.Wo), ho = (int)((m / g.Wo) % g.Ho);
                const int kk = tap % g.kw, jj = tap / g.kw;
                const float *op = off + m * (i64)(dg * 2 * K) + ((i64)dgi * K + tap) * 2;
                const float ph = sample_pos(ho, g.sh, g.ph, jj, g.dh, __ldg(op));
                const float pw = sample_pos(wo, g.sw, g.pw, kk, g.dw, __ldg(op + 1));
                const Sample2 sm = make_sample2(ph, pw, g.H, g.W);
                if (sm.mask & 1) {
                    const float lh = sm.l[0], lw = sm.l[1], hh = 1.f - lh, hw = 1.f - lw;
                    const float mk = mask ? __ldg(mask + m * (i64)(dg * K) + dgi * K + tap) : 1.f;
                    w.x = (sm.mask & (1 << 1)) ? hh * hw * mk : 0.f;
                    w.y = (sm.mask & (1 << 2)) ? hh * lw * mk : 0.f;
                    w.z = (sm.mask & (1 << 3)) ? lh * hw * mk : 0.f;
                    w.w = (sm.mask & (1 << 4)) ? lh * lw * mk : 0.f;
                    const int h0 = max(sm.lo[0], 0), h1 = min(sm.lo[0] + 1, g.H - 1), w0 = max(sm.lo[1], 0), w1 = min(sm.lo[1] + 1, g.W - 1);
                    o.x = (h0 * g.W + w0) * C; o.y = (h0 * g.W + w1) * C; o.z = (h1 * g.W + w0) * C; o.w = (h1 * g.W + w1) * C;
                }
            }
            sO[e] = o;
            sW[e] = w;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < PB * C4; item += blockDim.x) {
            const int p = item / C4, c = (item % C4) * 4;
            const i64 m = m0 + p;
            if (m >= M) continue;
            const int b = (int)(m / ((i64)g.Ho * g.Wo));
            const float *img = x + (i64)b * g.H * g.W * C + c;
            const int r0 = (p * dg + c / cpg) * K;
            float4 acc = bias ? ldg4(bias + c) : f4zero();
#pragma unroll DW2D_UNROLL
            for (int tap = 0; tap < K; ++tap) {
                const int4 o = sO[r0 + tap];
                const float4 w = sW[r0 + tap];
                const float4 v0 = ldg4(img + o.x), v1 = ldg4(img + o.y), v2 = ldg4(img + o.z), v3 = ldg4(img + o.w);
                float4 v = f4zero();
                fma4(v, w.x, v0); fma4(v, w.y, v1); fma4(v, w.z, v2); fma4(v, w.w, v3);
                const float4 wv = ldg4(wp + (i64)tap * C + c);
                acc.x = fmaf(wv.x, v.x, acc.x); acc.y = fmaf(wv.y, v.y, acc.y);
                acc.z = fmaf(wv.z, v.z, acc.z); acc.w = fmaf(wv.w, v.w, acc.w);
            }
            *reinterpret_cast<float4 *>(y + m * (i64)C + c) = acc;
        }
        __syncthreads();   // the records are rewritten by the next pixel group
    }
}

int pack_dw(const float *w, float *wp, int C, int taps, cudaStream_t st)
{
    if (pack_skipped()) return DLKA_OK;   // prepacked weights: see PackSkipScope
    DLKA_LAUNCH("pack_dw_weight", st, pack_dw_weight_kernel<<<(int)cdiv((i64)C * taps, 256), 256, 0, st>>>(w, wp, C, taps));
    return DLKA_OK;
}

template <int KS, int DIL, int R>
int launch_dw(const float *x, const float *wp, const float *bias, float *y, int B, int C, int D, int H, int W, int kd,
              cudaStream_t st)
{
    const int cx = C / 4;
    int ty = 256 / cx;
    if (ty < 1) ty = 1;
    const int wthreads = (int)cdiv(W, DIL * R) * DIL;
    if (ty > wthreads) ty = wthreads;
    dim3 block(cx, ty), grid((unsigned)cdiv(wthreads, ty), (unsigned)H, (unsigned)(B * D));
    if (grid.y > 65535u || grid.z > 65535u) return DLKA_ERR_UNSUPPORTED;
    DLKA_LAUNCH(KS == 5 ? "dwconv3d_k5" : "dwconv3d_k7d3", st,
                (dwconv3d_cl_kernel<KS, DIL, R><<<grid, block, 0, st>>>(x, wp, bias, y, C, D, H, W, kd)));
    return DLKA_OK;
}

}  // namespace

int dwconv_cl(const float *x, const float *w, const float *bias, float *y, int B, int C, int D, int H, int W, int kd,
              int kh, int kw, int dd, int dil, float *w_packed, cudaStream_t st, bool chunk_major_out)
{
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    if (chunk_major_out && C % 32 != 0) return DLKA_ERR_UNSUPPORTED;
    if (kd < 1 || kh < 1 || kw < 1 || !(kd & 1) || !(kh & 1) || !(kw & 1) || dd < 1 || dil < 1) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(pack_dw(w, w_packed, C, kd * kh * kw, st));
    const i64 voxels = (i64)B * D * H * W;
    // small volumes (the deep stages of the 3D nets, 4^3 .. 32^3): the plane-streaming kernel's cost is set by its tile shape, not by
    // the volume (measured ~140 us for 7^3-dil-3 and ~34 us for 5^3 on ANY volume below one wave of tiles), while the
    // thread-per-voxel kernel scales with the voxel count
    const bool small = voxels * C <= ((i64)1 << 20);
    if (!small && dwconv_smem_supported(C, kd, kh, kw, dd, dil, dil)) return dwconv_smem(x, w_packed, bias, y, B, C, D, H, W, kd, kh, dd, dil, st, chunk_major_out);
    if (!small && !chunk_major_out && C / 4 <= 256 && kh == kw && dd == dil && (kd == kh || kd == 1)) {
        if (kh == 5 && dil == 1) return launch_dw<5, 1, 4>(x, w_packed, bias, y, B, C, D, H, W, kd, st);
        if (kh == 7 && dil == 3) return launch_dw<7, 3, 4>(x, w_packed, bias, y, B, C, D, H, W, kd, st);
    }
    const i64 total = voxels * (C / 4);
    if (total <= 0) return DLKA_OK;
    const int blocks = (int)(cdiv(total, 256) < 148 * 32 ? cdiv(total, 256) : 148 * 32);
    const i64 ych = chunk_major_out ? voxels * 32 : 32;
    DLKA_LAUNCH("dwconv3d_generic", st,
                dwconv3d_generic_kernel<<<blocks, 256, 0, st>>>(x, w_packed, bias, y, B, C, D, H, W, kd, kh, kw, dd, dil, dil, ych,
                                                                 chunk_major_out ? 32 : C));
    return DLKA_OK;
}

int deform_dwconv_cl(const float *x, const float *off, const float *mask, const float *w, const float *bias, float *y,
                     const ConvGeo &g, float *w_packed, cudaStream_t st)
{
    if (g.C % 4 != 0 || (g.C / g.dg) % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(pack_dw(w, w_packed, g.C, g.K, st));
    const i64 M = (i64)g.B * g.Do * g.Ho * g.Wo;
    const i64 total = M * (g.C / 4);
    if (total <= 0) return DLKA_OK;
    const int blocks = (int)(cdiv(total, 256) < 148 * 32 ? cdiv(total, 256) : 148 * 32);
    if (g.ndim == 3) {
        DLKA_LAUNCH("deform_dwconv3d", st, deform_dwconv_cl_kernel<3><<<blocks, 256, 0, st>>>(x, off, mask, w_packed, bias, y, g, M));
        return DLKA_OK;
    }
    // shared-parameter kernel: pixels per block chosen so that (PB * C/4) fills whole rounds of 256 threads where possible and the
    // grid still has ~4 blocks per SM; records must fit the (opt-in) shared memory and image offsets 32 bits
    const int C4 = g.C / 4;
    int PB = 16;
    if ((16 * C4) % 256 != 0 && (32 * C4) % 256 == 0 && M / 32 >= 4 * 148) PB = 32;
    while (PB > 4 && M / PB < 4 * 148) PB >>= 1;
    const size_t smem = (size_t)PB * g.dg * g.K * 32;
    if (smem <= 96 * 1024 && (i64)g.H * g.W * g.C < ((i64)1 << 31)) {
        static SmemOptIn optin;
        DLKA_TRY(optin.ensure(deform_dwconv2d_shared_kernel, smem));
        const i64 ngroups = cdiv(M, (i64)PB);
        const int nb = (int)(ngroups < 148 * 16 ? ngroups : 148 * 16);
        DLKA_LAUNCH("deform_dwconv2d", st,
                    deform_dwconv2d_shared_kernel<<<nb, 256, smem, st>>>(x, off, mask, w_packed, bias, y, g, M, PB));
        return DLKA_OK;
    }
    DLKA_LAUNCH("deform_dwconv2d", st, deform_dwconv_cl_kernel<2><<<blocks, 256, 0, st>>>(x, off, mask, w_packed, bias, y, g, M));
    return DLKA_OK;
}

}  // namespace dlka
