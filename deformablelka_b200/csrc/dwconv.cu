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
    if (g.ndim == 3)
        DLKA_LAUNCH("deform_dwconv3d", st, deform_dwconv_cl_kernel<3><<<blocks, 256, 0, st>>>(x, off, mask, w_packed, bias, y, g, M));
    else
        DLKA_LAUNCH("deform_dwconv2d", st, deform_dwconv_cl_kernel<2><<<blocks, 256, 0, st>>>(x, off, mask, w_packed, bias, y, g, M));
    return DLKA_OK;
}

}  // namespace dlka
