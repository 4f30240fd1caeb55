// norm_mlp.cu -- the thin layers AROUND the D-LKA attention block, channels-last fp32:
//   * LayerNorm over C per token            (nn.LayerNorm(dim): 2D/networks/MaxViT_deform_LKA.py:151,156;
//                                            3D transformerblock.py:607, optional pos_embed add :622-623)
//   * layer-scale residual  out = x + s*y   (MaxViT_deform_LKA.py:173-175,183-185; transformerblock.py:624 gamma)
//   * depthwise 3x3 conv + bias + GELU      (Mlp.dwconv = DWConvLKA, MaxViT_deform_LKA.py:18-27,44-46)
// All three are HBM-bound streaming kernels; the contractions in between run on the tcgen05 dense kernel.
#include "kernels.cuh"

namespace dlka {
namespace {

// one warp per token row; C <= 32 * 4 * LN_MAXV
constexpr int LN_MAXV = 4;

// shuffle (P > 0): the input rows are [B*H*W][P*P] groups of C channels (output of PatchExpand's Linear) and group (p1, p2)
// of token (b, h, w) is written to token (b, h*P + p1, w*P + p2) -- einops "b h w (p1 p2 c) -> b (h p1) (w p2) c"
// (2D/networks/MaxViT_deform_LKA.py:510,540) fused into the LayerNorm that follows it.
__global__ void __launch_bounds__(256) layernorm_cl_kernel(const float *__restrict__ x, const float *__restrict__ pos,
                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                           float *__restrict__ y, i64 M, int C, i64 pos_rows, float eps, int P, int H,
                                                           int W)
{
    const int lane = threadIdx.x & 31;
    const i64 row = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    i64 orow = row;
    if (P > 0) {
        const int g = (int)(row % (P * P));
        i64 t = row / (P * P);
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const i64 b = t / H;
        orow = (b * (H * P) + h * P + g / P) * (i64)(W * P) + w * P + g % P;
    }
    const float *xr = x + row * C;
    const float *pr = pos ? pos + (row % pos_rows) * C : nullptr;
    float4 v[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        v[i] = f4zero();
        if (c < C) {
            v[i] = ldg4(xr + c);
            if (pr) { const float4 p = ldg4(pr + c); v[i].x += p.x; v[i].y += p.y; v[i].z += p.z; v[i].w += p.w; }
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            const float a = v[i].x - mean, b = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
            sq += (a * a + b * b) + (d * d + e * e);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);  // biased variance, as torch.nn.LayerNorm
    float *yr = y + orow * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < C) {
            const float4 g = gamma ? ldg4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 b = beta ? ldg4(beta + c) : f4zero();
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
            *reinterpret_cast<float4 *>(yr + c) = o;
        }
    }
}

// out = x (+ pos) + scale[c] * y
__global__ void __launch_bounds__(256) scale_residual_kernel(const float *__restrict__ x, const float *__restrict__ pos,
                                                             const float *__restrict__ scale, const float *__restrict__ y,
                                                             float *__restrict__ out, i64 M, int C, i64 pos_rows)
{
    const int C4 = C / 4;
    const i64 total = M * C4;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const i64 row = i / C4;
        float4 a = ldg4(x + row * C + c);
        if (pos) { const float4 p = ldg4(pos + (row % pos_rows) * C + c); a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w; }
        const float4 s = ldg4(scale + c), b = ldg4(y + row * C + c);
        a.x = fmaf(s.x, b.x, a.x); a.y = fmaf(s.y, b.y, a.y); a.z = fmaf(s.z, b.z, a.z); a.w = fmaf(s.w, b.w, a.w);
        *reinterpret_cast<float4 *>(out + row * C + c) = a;
    }
}

// depthwise 3x3, pad 1, stride 1, + bias, optional exact GELU; w packed [9][C]
__global__ void __launch_bounds__(256) dwconv2d3_cl_kernel(const float *__restrict__ x, const float *__restrict__ wp,
                                                           const float *__restrict__ bias, float *__restrict__ y, int B, int C,
                                                           int H, int W, int gelu)
{
    const int C4 = C / 4;
    const i64 total = (i64)B * H * W * C4;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        i64 p = i / C4;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        float4 acc = bias ? ldg4(bias + c) : f4zero();
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int hh = h + j - 1;
            if ((unsigned)hh >= (unsigned)H) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ww = w + k - 1;
                if ((unsigned)ww >= (unsigned)W) continue;
                fma4v(acc, ldg4(wp + (j * 3 + k) * C + c), ldg4(x + (((i64)b * H + hh) * W + ww) * C + c));
            }
        }
        if (gelu) { acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w); }
        *reinterpret_cast<float4 *>(y + (((i64)b * H + h) * W + w) * C + c) = acc;
    }
}

// y = act(scale[c] * y + shift[c] (+ e)); act: 1 LeakyReLU, 2 (+e) then LeakyReLU   (fallback path of conv3_bn_act_cl)
__global__ void __launch_bounds__(256) affine_act_kernel(float *__restrict__ y, const float *__restrict__ scale,
                                                         const float *__restrict__ shift, const float *__restrict__ e, i64 M, int C,
                                                         int act, float slope)
{
    const int C4 = C / 4;
    const i64 total = M * C4;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const i64 row = i / C4;
        float4 v = *reinterpret_cast<float4 *>(y + row * C + c);
        const float4 s = scale ? ldg4(scale + c) : make_float4(1.f, 1.f, 1.f, 1.f), t = shift ? ldg4(shift + c) : f4zero();
        v.x = fmaf(s.x, v.x, t.x); v.y = fmaf(s.y, v.y, t.y); v.z = fmaf(s.z, v.z, t.z); v.w = fmaf(s.w, v.w, t.w);
        if (act == 2) { const float4 r = ldg4(e + row * C + c); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        if (act) {
            v.x = v.x > 0.f ? v.x : slope * v.x; v.y = v.y > 0.f ? v.y : slope * v.y;
            v.z = v.z > 0.f ? v.z : slope * v.z; v.w = v.w > 0.f ? v.w : slope * v.w;
        }
        *reinterpret_cast<float4 *>(y + row * C + c) = v;
    }
}

__global__ void pack_dw9_kernel(const float *__restrict__ w, float *__restrict__ wp, int C)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 9 * C; i += gridDim.x * blockDim.x) wp[i] = w[(i % C) * 9 + i / C];
}

}  // namespace

int layernorm_cl(const float *x, const float *pos, i64 pos_rows, const float *gamma, const float *beta, float *y, i64 M, int C,
                 float eps, cudaStream_t st)
{
    if (C % 4 != 0 || C > 32 * 4 * LN_MAXV) return DLKA_ERR_UNSUPPORTED;
    if (M <= 0) return DLKA_OK;
    const int rows_per_block = 8;
    DLKA_LAUNCH("layernorm_cl", st,
                layernorm_cl_kernel<<<(unsigned)cdiv(M, rows_per_block), rows_per_block * 32, 0, st>>>(x, pos, gamma, beta, y, M, C,
                                                                                                       pos_rows > 0 ? pos_rows : 1, eps, 0, 1, 1));
    return DLKA_OK;
}

int layernorm_shuffle_cl(const float *x, const float *gamma, const float *beta, float *y, int B, int H, int W, int P, int C, float eps,
                         cudaStream_t st)
{
    if (C % 4 != 0 || C > 32 * 4 * LN_MAXV || P < 1) return DLKA_ERR_UNSUPPORTED;
    const i64 M = (i64)B * H * W * P * P;
    if (M <= 0) return DLKA_OK;
    const int rows_per_block = 8;
    DLKA_LAUNCH("layernorm_shuffle_cl", st,
                layernorm_cl_kernel<<<(unsigned)cdiv(M, rows_per_block), rows_per_block * 32, 0, st>>>(x, nullptr, gamma, beta, y, M, C, 1,
                                                                                                       eps, P, H, W));
    return DLKA_OK;
}

int scale_residual_cl(const float *x, const float *pos, i64 pos_rows, const float *scale, const float *y, float *out, i64 M, int C,
                      cudaStream_t st)
{
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    if (M <= 0) return DLKA_OK;
    const i64 total = M * (C / 4);
    const int blocks = (int)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
    DLKA_LAUNCH("scale_residual", st,
                scale_residual_kernel<<<blocks, 256, 0, st>>>(x, pos, scale, y, out, M, C, pos_rows > 0 ? pos_rows : 1));
    return DLKA_OK;
}

int affine_act_cl(float *y, const float *scale, const float *shift, const float *e, i64 M, int C, int act, float slope, cudaStream_t st)
{
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    if (M <= 0) return DLKA_OK;
    const i64 total = M * (C / 4);
    const int blocks = (int)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
    DLKA_LAUNCH("affine_act", st, affine_act_kernel<<<blocks, 256, 0, st>>>(y, scale, shift, e, M, C, act, slope));
    return DLKA_OK;
}

int dwconv2d3_cl(const float *x, const float *w, const float *bias, float *y, int B, int C, int H, int W, int gelu, float *w_packed,
                 cudaStream_t st)
{
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    if (!pack_skipped()) DLKA_LAUNCH("pack_dw9", st, pack_dw9_kernel<<<(int)cdiv(9 * C, 256), 256, 0, st>>>(w, w_packed, C));
    const i64 total = (i64)B * H * W * (C / 4);
    if (total <= 0) return DLKA_OK;
    const int blocks = (int)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
    DLKA_LAUNCH("dwconv2d3_gelu", st, dwconv2d3_cl_kernel<<<blocks, 256, 0, st>>>(x, w_packed, bias, y, B, C, H, W, gelu));
    return DLKA_OK;
}

}  // namespace dlka
