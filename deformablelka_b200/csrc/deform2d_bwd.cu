// deform2d_bwd.cu -- backward of the 2D deformable convolution (SURVEY.md 8f row N2, 2D half): the gradients that
// torch.ops.torchvision.deform_conv2d's autograd produces (torchvision/ops/deform_conv.py:92-107; the published kernels
// deformable_col2im / deformable_col2im_coord / the weight GEMM) for input, weight, offset, mask and bias -- every
// configuration of the forward: weight groups (depthwise = the D-LKA 2D hot path, 2D/deformable_LKA/deformable_LKA.py:18-25),
// offset groups, DCNv2 mask.  No columns buffer is materialised (the reference's is C*K x B*Ho*Wo floats).
//
// With col[b,c,t,p] = mask * bilinear(x[b,c], p + tap t + offset) and out = W . col (+ bias):
//   gcol[b,c,t,p]   = sum_{co in group(c)} W[co, c', t] * gout[b, co, p]
//   grad_input      : gcol * mask scattered to the 4 corners with their bilinear weights
//   grad_offset     = sum_{c in offset group} gcol * mask * d val / d(y | x)
//   grad_mask       = sum_{c in offset group} gcol * val
//   grad_weight     = sum_{b,p} gout[b,co,p] * col[b,c,t,p]
//   grad_bias       = sum_{b,p} gout[b,co,p]
// One thread owns one channel of one pixel and walks the taps: consecutive lanes are consecutive channels of channels-last
// tensors, so the four corner reads and the four corner reductions (red.global.add) are coalesced; offset / mask gradients
// are reduced over the warp when its 32 channels share an offset group; a depthwise weight gradient is accumulated in
// registers over the block's pixels and reduced once per (channel, tap).  fp32 throughout.
#include "kernels.cuh"

namespace dlka {
namespace {

constexpr int B2_PIX = 64;   // pixels per block (each thread-row walks B2_PIX / blockDim.y of them)

// KH > 0: depthwise (Co == C == groups) with a compile-time KH x KW kernel: taps unrolled, weight gradient in registers.
// KH == 0: any configuration, runtime kernel size, weight gradient by atomics.
template <int KH, int KW>
__global__ void __launch_bounds__(256) deform2d_bwd_kernel(const float *__restrict__ x, const float *__restrict__ w /* [Co][Cg][K] */,
                                                           const float *__restrict__ off, const float *__restrict__ mask,
                                                           const float *__restrict__ gout, float *__restrict__ gx, float *__restrict__ gw,
                                                           float *__restrict__ goff, float *__restrict__ gmask, const ConvGeo g, i64 M)
{
    const int c = blockIdx.y * 32 + threadIdx.x;
    const bool cv = c < g.C;
    const int K = g.K, Cg = g.C / g.groups, Cog = g.Co / g.groups, cpo = g.C / g.dg;
    const int grp = cv ? c / Cg : 0, cin = cv ? c % Cg : 0, og = cv ? c / cpo : 0;
    // the whole warp shares one offset group -> one shuffle reduction + one atomic per (pixel, tap) instead of 32 atomics
    const bool warp_uniform_og = (cpo % 32) == 0 || g.dg == 1;
    constexpr bool DW = KH > 0;
    constexpr int KMAX = DW ? KH * KW : 1;
    float accw[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) accw[t] = 0.f;
    const int kh = DW ? KH : g.kh, kw = DW ? KW : g.kw;
    const i64 m0 = (i64)blockIdx.x * B2_PIX;
    for (int pp = threadIdx.y; pp < B2_PIX; pp += blockDim.y) {
        const i64 m = m0 + pp;
        if (m >= M) break;   // uniform across the warp (a warp = one threadIdx.y)
        const int wo = (int)(m % g.Wo);
        const i64 q = m / g.Wo;
        const int ho = (int)(q % g.Ho), b = (int)(q / g.Ho);
        const float *img = x + (i64)b * g.H * g.W * g.C + (cv ? c : 0);
        float *gimg = gx + (i64)b * g.H * g.W * g.C + (cv ? c : 0);
        const float *orow = off + m * (i64)(g.dg * 2 * K) + (i64)og * 2 * K;
        const float *mrow = mask ? mask + m * (i64)(g.dg * K) + (i64)og * K : nullptr;
        const float *grow = gout + m * (i64)g.Co;
#pragma unroll
        for (int jj = 0; jj < kh; ++jj)
#pragma unroll
            for (int kk = 0; kk < kw; ++kk) {
                const int t = jj * kw + kk;
                float gcol = 0.f, val = 0.f, dy = 0.f, dx = 0.f, mk = 1.f;
                float cw[4] = {0.f, 0.f, 0.f, 0.f};
                i64 co_[4] = {0, 0, 0, 0};
                if (cv) {
                    const float ph = sample_pos(ho, g.sh, g.ph, jj, g.dh, __ldg(orow + 2 * t));
                    const float pw = sample_pos(wo, g.sw, g.pw, kk, g.dw, __ldg(orow + 2 * t + 1));
                    if (mrow) mk = __ldg(mrow + t);
                    const Sample2 s = make_sample2(ph, pw, g.H, g.W);
                    if (DW) {
                        gcol = __ldg(w + (i64)c * K + t) * __ldg(grow + c);
                    } else {
                        for (int j = 0; j < Cog; ++j) {
                            const int co = grp * Cog + j;
                            gcol = fmaf(__ldg(w + ((i64)co * Cg + cin) * K + t), __ldg(grow + co), gcol);
                        }
                    }
                    if (s.mask & 1) {
                        const float lh = s.l[0], lw = s.l[1], hh = 1.f - lh, hw = 1.f - lw;
                        const i64 sH = (i64)g.W * g.C;
                        const i64 p00 = (i64)s.lo[0] * sH + (i64)s.lo[1] * g.C;
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
                        if (s.mask & 2) { v[0] = __ldg(img + p00); cw[0] = hh * hw; co_[0] = p00; }
                        if (s.mask & 4) { v[1] = __ldg(img + p00 + g.C); cw[1] = hh * lw; co_[1] = p00 + g.C; }
                        if (s.mask & 8) { v[2] = __ldg(img + p00 + sH); cw[2] = lh * hw; co_[2] = p00 + sH; }
                        if (s.mask & 16) { v[3] = __ldg(img + p00 + sH + g.C); cw[3] = lh * lw; co_[3] = p00 + sH + g.C; }
                        val = cw[0] * v[0] + cw[1] * v[1] + cw[2] * v[2] + cw[3] * v[3];
                        // d val / d y and d x: derivative of the corner weights (floor() has zero derivative), dropped corners drop out
                        dy = -hw * v[0] - lw * v[1] + hw * v[2] + lw * v[3];
                        dx = -hh * v[0] + hh * v[1] - lh * v[2] + lh * v[3];
                    }
                    const float gv = gcol * mk;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        if (cw[k4] != 0.f) atomicAdd(gimg + co_[k4], gv * cw[k4]);
                    if (DW) {
                        accw[DW ? t : 0] += __ldg(grow + c) * (mk * val);
                    } else {
                        const float colv = mk * val;
                        for (int j = 0; j < Cog; ++j) {
                            const int co = grp * Cog + j;
                            atomicAdd(gw + ((i64)co * Cg + cin) * K + t, __ldg(grow + co) * colv);
                        }
                    }
                }
                float gy = gcol * mk * dy, gxx = gcol * mk * dx, gm = gcol * val;
                float *gorow = goff + m * (i64)(g.dg * 2 * K) + (i64)og * 2 * K + 2 * t;
                if (warp_uniform_og) {
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        gy += __shfl_xor_sync(0xffffffffu, gy, sft);
                        gxx += __shfl_xor_sync(0xffffffffu, gxx, sft);
                        gm += __shfl_xor_sync(0xffffffffu, gm, sft);
                    }
                    if (threadIdx.x == 0) {
                        atomicAdd(gorow, gy);
                        atomicAdd(gorow + 1, gxx);
                        if (gmask) atomicAdd(gmask + m * (i64)(g.dg * K) + (i64)og * K + t, gm);
                    }
                } else if (cv) {
                    atomicAdd(gorow, gy);
                    atomicAdd(gorow + 1, gxx);
                    if (gmask) atomicAdd(gmask + m * (i64)(g.dg * K) + (i64)og * K + t, gm);
                }
            }
    }
    if (DW && cv) {
        // depthwise: one reduction per (channel, tap) and thread row (the rows of a block hold different pixels of the same channel)
#pragma unroll
        for (int t = 0; t < KMAX; ++t)
            if (accw[t] != 0.f) atomicAdd(gw + (i64)c * K + t, accw[t]);
    }
}

// grad_bias[co] = sum over rows of gout[m][co]
__global__ void colsum_kernel(const float *__restrict__ gout, float *__restrict__ gb, i64 M, int Co)
{
    const int co = blockIdx.y * 32 + threadIdx.x;
    float acc = 0.f;
    if (co < Co)
        for (i64 m = (i64)blockIdx.x * blockDim.y + threadIdx.y; m < M; m += (i64)gridDim.x * blockDim.y) acc += gout[m * Co + co];
    __shared__ float red[8][33];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && co < Co) {
        float s = 0.f;
        for (int i = 0; i < (int)blockDim.y; ++i) s += red[i][threadIdx.x];
        atomicAdd(gb + co, s);
    }
}

}  // namespace

// All tensors channels-last: x / gx [B][H][W][C], gout [M][Co], off / goff [M][dg*2*K], mask / gmask [M][dg*K] (or null),
// w / gw in the reference's [Co][C/g][kh*kw] layout, gb [Co] or null.  gx / gw / goff / gmask / gb are zeroed here.
int deform2d_backward_cl(const ConvGeo &g, const float *x, const float *w, const float *off, const float *mask, const float *gout,
                         float *gx, float *gw, float *goff, float *gmask, float *gb, cudaStream_t st)
{
    const i64 M = (i64)g.B * g.Ho * g.Wo;
    const int K = g.K;
    DLKA_CUDA_TRY(cudaMemsetAsync(gx, 0, (size_t)g.B * g.H * g.W * g.C * sizeof(float), st));
    DLKA_CUDA_TRY(cudaMemsetAsync(gw, 0, (size_t)g.Co * (g.C / g.groups) * K * sizeof(float), st));
    DLKA_CUDA_TRY(cudaMemsetAsync(goff, 0, (size_t)M * g.dg * 2 * K * sizeof(float), st));
    if (gmask) DLKA_CUDA_TRY(cudaMemsetAsync(gmask, 0, (size_t)M * g.dg * K * sizeof(float), st));
    if (gb) DLKA_CUDA_TRY(cudaMemsetAsync(gb, 0, (size_t)g.Co * sizeof(float), st));
    if (M <= 0) return DLKA_OK;
    const bool dwise = g.groups == g.C && g.Co == g.C;
    dim3 block(32, 8), grid((unsigned)cdiv(M, B2_PIX), (unsigned)cdiv(g.C, 32));
    if (dwise && g.kh == 5 && g.kw == 5)        // conv0 of the 2D block
        DLKA_LAUNCH("deform2d_bwd_dw", st, (deform2d_bwd_kernel<5, 5><<<grid, block, 0, st>>>(x, w, off, mask, gout, gx, gw, goff, gmask, g, M)));
    else if (dwise && g.kh == 7 && g.kw == 7)   // conv_spatial
        DLKA_LAUNCH("deform2d_bwd_dw", st, (deform2d_bwd_kernel<7, 7><<<grid, block, 0, st>>>(x, w, off, mask, gout, gx, gw, goff, gmask, g, M)));
    else if (dwise && g.kh == 3 && g.kw == 3)
        DLKA_LAUNCH("deform2d_bwd_dw", st, (deform2d_bwd_kernel<3, 3><<<grid, block, 0, st>>>(x, w, off, mask, gout, gx, gw, goff, gmask, g, M)));
    else
        DLKA_LAUNCH("deform2d_bwd", st, (deform2d_bwd_kernel<0, 0><<<grid, block, 0, st>>>(x, w, off, mask, gout, gx, gw, goff, gmask, g, M)));
    if (gb) {
        dim3 gb_grid((unsigned)(cdiv(M, 8 * 64) < 1024 ? cdiv(M, 8 * 64) : 1024), (unsigned)cdiv(g.Co, 32));
        DLKA_LAUNCH("colsum", st, (colsum_kernel<<<gb_grid, dim3(32, 8), 0, st>>>(gout, gb, M, g.Co)));
    }
    return DLKA_OK;
}

}  // namespace dlka
