/*
 * dlka_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in plain C, of the deformable-convolution arithmetic on the
 * D-LKA hot path of xmindflow/deformableLKA.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file.
 *
 * What is restated (citations are reference file:line under /root/reference):
 *   3D  deformable im2col           3D/dcn/src/cuda/deform_im2col_cuda.cuh:192-265
 *       trilinear sampler           3D/dcn/src/cuda/deform_im2col_cuda.cuh:26-72
 *       output extent / grouping    3D/dcn/src/cuda/deform_conv_cuda.cu:78-123
 *   2D  torchvision.ops.deform_conv2d (third-party dependency, pinned
 *       torchvision==0.12.0 in 2D/requirements.txt:69; source not vendored in
 *       the reference).  Its published algorithm (bilinear im2col + grouped GEMM)
 *       is restated here and pinned against the installed torchvision CPU op in
 *       tests/test_oracle.py.
 *
 * Parity pins (tests/test_oracle.py, tests/golden/): K1 zero offsets == stock
 * conv, K2 3D with D=1 == torchvision deform_conv2d, K3 fresh DeformConvPack ==
 * nn.Conv3d, golden vectors produced by importing the unmodified reference 2D
 * module (tests/golden/make_golden.py).
 *
 * All flattened indices are int64 (the reference overflows int32 at the headline
 * shape).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ---- sampling position (cuh:224-226,245-247): integer base first, one fp32 add ---- */
static inline float sample_pos(int out_idx, int stride, int pad, int tap, int dil, float delta)
{
    const int base = out_idx * stride - pad + tap * dil; /* int arithmetic */
    volatile float p = (float)base + delta;               /* single fp32 add, no contraction */
    return p;
}

/* ---- trilinear sampler, restating dmcn_im2col_bilinear (cuh:26-72) ---- */
static inline float trilinear(const float *vol, int D, int H, int W, float d, float h, float w)
{
    const int d_low = (int)floorf(d), h_low = (int)floorf(h), w_low = (int)floorf(w);
    const int d_high = d_low + 1, h_high = h_low + 1, w_high = w_low + 1;
    const float ld = d - d_low, lh = h - h_low, lw = w - w_low;
    const float hd = 1 - ld, hh = 1 - lh, hw = 1 - lw;
    const i64 HW = (i64)H * W;
    float v1 = 0, v2 = 0, v3 = 0, v4 = 0, v5 = 0, v6 = 0, v7 = 0, v8 = 0;
    if (d_low >= 0 && h_low >= 0 && w_low >= 0) v1 = vol[d_low * HW + (i64)h_low * W + w_low];
    if (d_low >= 0 && h_low >= 0 && w_high <= W - 1) v2 = vol[d_low * HW + (i64)h_low * W + w_high];
    if (d_low >= 0 && h_high <= H - 1 && w_low >= 0) v3 = vol[d_low * HW + (i64)h_high * W + w_low];
    if (d_low >= 0 && h_high <= H - 1 && w_high <= W - 1) v4 = vol[d_low * HW + (i64)h_high * W + w_high];
    if (d_high <= D - 1 && h_low >= 0 && w_low >= 0) v5 = vol[d_high * HW + (i64)h_low * W + w_low];
    if (d_high <= D - 1 && h_low >= 0 && w_high <= W - 1) v6 = vol[d_high * HW + (i64)h_low * W + w_high];
    if (d_high <= D - 1 && h_high <= H - 1 && w_low >= 0) v7 = vol[d_high * HW + (i64)h_high * W + w_low];
    if (d_high <= D - 1 && h_high <= H - 1 && w_high <= W - 1) v8 = vol[d_high * HW + (i64)h_high * W + w_high];
    const float w1 = hd * hh * hw, w2 = hd * hh * lw, w3 = hd * lh * hw, w4 = hd * lh * lw;
    const float w5 = ld * hh * hw, w6 = ld * hh * lw, w7 = ld * lh * hw, w8 = ld * lh * lw;
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4 + w5 * v5 + w6 * v6 + w7 * v7 + w8 * v8;
}

static inline int out_extent(int in, int pad, int dil, int k, int stride)
{
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1; /* cu:78-80 */
}

/*
 * 3D deformable im2col for a chunk of output voxels [v0, v1) of sample b.
 *   input  [B, C, D, H, W]                      (NCDHW, contiguous)
 *   offset [B, dg*3*K, Do, Ho, Wo]              channel 3t+{0,1,2} = (dd, dh, dw) of tap t
 *   cols   [(C*K), (v1-v0)]  row index = c*K + t (cuh:220,228,259-260)
 */
void oracle_deform_im2col3d(const float *input, const float *offset, float *cols,
                            int B, int C, int D, int H, int W,
                            int kd, int kh, int kw, int sd, int sh, int sw,
                            int pd, int ph, int pw, int dd, int dh, int dw,
                            int deformable_group, int b, i64 v0, i64 v1)
{
    (void)B;
    const int Do = out_extent(D, pd, dd, kd, sd), Ho = out_extent(H, ph, dh, kh, sh), Wo = out_extent(W, pw, dw, kw, sw);
    const int K = kd * kh * kw;
    const i64 Vo = (i64)Do * Ho * Wo, Vi = (i64)D * H * W, n = v1 - v0;
    const int cpg = C / deformable_group; /* cuh:222 */
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c) {
        const int dgi = c / cpg;
        const float *vol = input + ((i64)b * C + c) * Vi;
        const float *off = offset + ((i64)b * deformable_group + dgi) * 3 * K * Vo;
        for (i64 v = v0; v < v1; ++v) {
            const int wo = (int)(v % Wo), ho = (int)((v / Wo) % Ho), d_o = (int)(v / Wo / Ho);
            for (int i = 0; i < kd; ++i)
                for (int j = 0; j < kh; ++j)
                    for (int k = 0; k < kw; ++k) {
                        const int t = (i * kh + j) * kw + k;
                        const float od = off[(i64)(3 * t + 0) * Vo + v];
                        const float oh = off[(i64)(3 * t + 1) * Vo + v];
                        const float ow = off[(i64)(3 * t + 2) * Vo + v];
                        const float pd_ = sample_pos(d_o, sd, pd, i, dd, od);
                        const float ph_ = sample_pos(ho, sh, ph, j, dh, oh);
                        const float pw_ = sample_pos(wo, sw, pw, k, dw, ow);
                        float val = 0.f;
                        if (pd_ > -1 && ph_ > -1 && pw_ > -1 && pd_ < D && ph_ < H && pw_ < W) /* cuh:248 */
                            val = trilinear(vol, D, H, W, pd_, ph_, pw_);
                        cols[((i64)c * K + t) * n + (v - v0)] = val;
                    }
        }
    }
}

/*
 * Integer planes of the sampler for parity check K4 (bit-exact): for every
 * (b, voxel, tap): low[3] = floor(p) per axis and a 9-bit mask:
 * bit0 = whole-sample valid (cuh:248), bits1..8 = corner v1..v8 used (cuh:43-65).
 *   low  [B, Vo, K, 3] int32 ; mask [B, Vo, K] int32   (deformable group 0 .. dg-1 -> extra leading dim dg)
 */
void oracle_sample_indices3d(const float *offset, int32_t *low, int32_t *mask,
                             int B, int D, int H, int W,
                             int kd, int kh, int kw, int sd, int sh, int sw,
                             int pd, int ph, int pw, int dd, int dh, int dw, int deformable_group)
{
    const int Do = out_extent(D, pd, dd, kd, sd), Ho = out_extent(H, ph, dh, kh, sh), Wo = out_extent(W, pw, dw, kw, sw);
    const int K = kd * kh * kw;
    const i64 Vo = (i64)Do * Ho * Wo;
#pragma omp parallel for schedule(static)
    for (i64 bg = 0; bg < (i64)B * deformable_group; ++bg) {
        const float *off = offset + bg * 3 * K * Vo;
        for (i64 v = 0; v < Vo; ++v) {
            const int wo = (int)(v % Wo), ho = (int)((v / Wo) % Ho), d_o = (int)(v / Wo / Ho);
            for (int t = 0; t < K; ++t) {
                const int k = t % kw, j = (t / kw) % kh, i = t / kw / kh;
                const float p0 = sample_pos(d_o, sd, pd, i, dd, off[(i64)(3 * t + 0) * Vo + v]);
                const float p1 = sample_pos(ho, sh, ph, j, dh, off[(i64)(3 * t + 1) * Vo + v]);
                const float p2 = sample_pos(wo, sw, pw, k, dw, off[(i64)(3 * t + 2) * Vo + v]);
                const int l0 = (int)floorf(p0), l1 = (int)floorf(p1), l2 = (int)floorf(p2);
                int m = 0;
                if (p0 > -1 && p1 > -1 && p2 > -1 && p0 < D && p1 < H && p2 < W) {
                    m = 1;
                    const int dl = l0 >= 0, hl = l1 >= 0, wl = l2 >= 0;
                    const int dh_ = l0 + 1 <= D - 1, hh_ = l1 + 1 <= H - 1, wh_ = l2 + 1 <= W - 1;
                    m |= (dl && hl && wl) << 1;
                    m |= (dl && hl && wh_) << 2;
                    m |= (dl && hh_ && wl) << 3;
                    m |= (dl && hh_ && wh_) << 4;
                    m |= (dh_ && hl && wl) << 5;
                    m |= (dh_ && hl && wh_) << 6;
                    m |= (dh_ && hh_ && wl) << 7;
                    m |= (dh_ && hh_ && wh_) << 8;
                }
                const i64 o = (bg * Vo + v) * K + t;
                low[o * 3 + 0] = l0; low[o * 3 + 1] = l1; low[o * 3 + 2] = l2;
                mask[o] = m;
            }
        }
    }
}

/*
 * Full 3D deformable convolution forward in C (im2col chunk + naive GEMM),
 * restating deform_conv_cuda_forward (cu:78-123).  Used for small shapes; the
 * Python oracle uses oracle_deform_im2col3d + torch.addmm for big ones (the
 * reference itself calls at::addmm at cu:117).
 *   weight [Co, C/g, kd, kh, kw], bias [Co], output [B, Co, Do, Ho, Wo]
 */
void oracle_deform_conv3d_forward(const float *input, const float *offset, const float *weight,
                                  const float *bias, float *output,
                                  int B, int C, int D, int H, int W, int Co,
                                  int kd, int kh, int kw, int sd, int sh, int sw,
                                  int pd, int ph, int pw, int dd, int dh, int dw,
                                  int group, int deformable_group)
{
    const int Do = out_extent(D, pd, dd, kd, sd), Ho = out_extent(H, ph, dh, kh, sh), Wo = out_extent(W, pw, dw, kw, sw);
    const int K = kd * kh * kw;
    const i64 Vo = (i64)Do * Ho * Wo;
    const i64 chunk = 4096;
    float *cols = (float *)malloc(sizeof(float) * (size_t)C * K * chunk);
    const int cg = C / group, og = Co / group;
    for (int b = 0; b < B; ++b)
        for (i64 v0 = 0; v0 < Vo; v0 += chunk) {
            const i64 v1 = v0 + chunk < Vo ? v0 + chunk : Vo, n = v1 - v0;
            oracle_deform_im2col3d(input, offset, cols, B, C, D, H, W, kd, kh, kw, sd, sh, sw,
                                   pd, ph, pw, dd, dh, dw, deformable_group, b, v0, v1);
#pragma omp parallel for schedule(static)
            for (int o = 0; o < Co; ++o) {
                const int g = o / og;
                float *out = output + ((i64)b * Co + o) * Vo + v0;
                for (i64 v = 0; v < n; ++v) out[v] = bias ? bias[o] : 0.f;
                for (int r = 0; r < cg * K; ++r) { /* K index = c_in_group*K + t (cu:85,111,116) */
                    const float wv = weight[(i64)o * cg * K + r];
                    const float *col = cols + ((i64)g * cg * K + r) * n;
                    for (i64 v = 0; v < n; ++v) out[v] += wv * col[v];
                }
            }
        }
    free(cols);
}

/* ---- 2D: torchvision deform_conv2d (published algorithm), mask optional ---- */
static inline float bilinear(const float *img, int H, int W, float h, float w)
{
    if (h <= -1 || H <= h || w <= -1 || W <= w) return 0.f;
    const int h_low = (int)floorf(h), w_low = (int)floorf(w);
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
    float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0) v1 = img[(i64)h_low * W + w_low];
    if (h_low >= 0 && w_high <= W - 1) v2 = img[(i64)h_low * W + w_high];
    if (h_high <= H - 1 && w_low >= 0) v3 = img[(i64)h_high * W + w_low];
    if (h_high <= H - 1 && w_high <= W - 1) v4 = img[(i64)h_high * W + w_high];
    const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

/*
 *   input [B,C,H,W], offset [B, og*2*K, Ho, Wo] (channel 2t = dy, 2t+1 = dx),
 *   mask [B, og*K, Ho, Wo] or NULL, weight [Co, C/g, kh, kw], bias [Co] or NULL,
 *   output [B, Co, Ho, Wo]
 */
void oracle_deform_conv2d_forward(const float *input, const float *offset, const float *mask,
                                  const float *weight, const float *bias, float *output,
                                  int B, int C, int H, int W, int Co, int kh, int kw,
                                  int sh, int sw, int ph, int pw, int dh, int dw,
                                  int n_weight_grps, int n_offset_grps)
{
    const int Ho = out_extent(H, ph, dh, kh, sh), Wo = out_extent(W, pw, dw, kw, sw);
    const int K = kh * kw;
    const i64 P = (i64)Ho * Wo;
    const int cg = C / n_weight_grps, og = Co / n_weight_grps, cpo = C / n_offset_grps;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int o = 0; o < Co; ++o) {
            const int g = o / og;
            float *out = output + ((i64)b * Co + o) * P;
            for (i64 p = 0; p < P; ++p) {
                const int ox = (int)(p % Wo), oy = (int)(p / Wo);
                float acc = 0.f;
                for (int ci = 0; ci < cg; ++ci) {
                    const int c = g * cg + ci, ogi = c / cpo;
                    const float *img = input + ((i64)b * C + c) * H * W;
                    const float *off = offset + ((i64)b * n_offset_grps + ogi) * 2 * K * P;
                    const float *msk = mask ? mask + ((i64)b * n_offset_grps + ogi) * K * P : NULL;
                    for (int t = 0; t < K; ++t) {
                        const int i = t / kw, j = t % kw;
                        const float y = sample_pos(oy, sh, ph, i, dh, off[(i64)(2 * t) * P + p]);
                        const float x = sample_pos(ox, sw, pw, j, dw, off[(i64)(2 * t + 1) * P + p]);
                        float val = bilinear(img, H, W, y, x);
                        if (msk) val *= msk[(i64)t * P + p];
                        acc += weight[((i64)o * cg + ci) * K + t] * val;
                    }
                }
                out[p] = acc + (bias ? bias[o] : 0.f);
            }
        }
}

/* 2D integer planes (K4): low [B*og, P, K, 2] ; mask bit0 valid, bits1..4 corners v1..v4 */
void oracle_sample_indices2d(const float *offset, int32_t *low, int32_t *maskbits,
                             int B, int H, int W, int kh, int kw, int sh, int sw,
                             int ph, int pw, int dh, int dw, int n_offset_grps)
{
    const int Ho = out_extent(H, ph, dh, kh, sh), Wo = out_extent(W, pw, dw, kw, sw);
    const int K = kh * kw;
    const i64 P = (i64)Ho * Wo;
#pragma omp parallel for schedule(static)
    for (i64 bg = 0; bg < (i64)B * n_offset_grps; ++bg) {
        const float *off = offset + bg * 2 * K * P;
        for (i64 p = 0; p < P; ++p) {
            const int ox = (int)(p % Wo), oy = (int)(p / Wo);
            for (int t = 0; t < K; ++t) {
                const int i = t / kw, j = t % kw;
                const float y = sample_pos(oy, sh, ph, i, dh, off[(i64)(2 * t) * P + p]);
                const float x = sample_pos(ox, sw, pw, j, dw, off[(i64)(2 * t + 1) * P + p]);
                const int l0 = (int)floorf(y), l1 = (int)floorf(x);
                int m = 0;
                if (!(y <= -1 || H <= y || x <= -1 || W <= x)) {
                    m = 1;
                    const int hl = l0 >= 0, wl = l1 >= 0, hh_ = l0 + 1 <= H - 1, wh_ = l1 + 1 <= W - 1;
                    m |= (hl && wl) << 1; m |= (hl && wh_) << 2; m |= (hh_ && wl) << 3; m |= (hh_ && wh_) << 4;
                }
                const i64 o = (bg * P + p) * K + t;
                low[o * 2 + 0] = l0; low[o * 2 + 1] = l1;
                maskbits[o] = m;
            }
        }
    }
}
