// blocks_api.cu -- C-ABI entry points for the enclosing transformer blocks, tokens in / tokens out.
//   2D  deformableLKABlock.forward                       2D/networks/MaxViT_deform_LKA.py:165-189
//   3D  TransformerBlock_3D_single_deform_LKA.forward    transformerblock.py:617-624 (pos-embed, LayerNorm, gamma residual;
//       the UnetResBlock / conv8 tail at :626-628 is row N3 and not part of this entry)
// Tokens [B, N, C] are channels-last data, so neither block needs a layout change.
#include "kernels.cuh"

using namespace dlka;

namespace {

struct Lka2dBlockPlan {
    float *ln, *attn, *x1, *h1, *h2, *f2, *wp_fc1, *wp_fc2, *wp_dw;
    void *attn_ws;
    size_t attn_ws_bytes;
};

bool plan_block(Arena &ar, int B, int C, int H, int W, int hidden, Lka2dBlockPlan &p)
{
    const size_t M = (size_t)B * H * W;
    p.ln = ar.take<float>(M * C);
    p.attn = ar.take<float>(M * C);
    p.x1 = ar.take<float>(M * C);
    p.h1 = ar.take<float>(M * hidden);
    p.h2 = ar.take<float>(M * hidden);
    p.f2 = ar.take<float>(M * C);
    p.wp_fc1 = ar.take<float>(dense_scratch_floats(hidden, C));
    p.wp_fc2 = ar.take<float>(dense_scratch_floats(C, hidden));
    p.wp_dw = ar.take<float>((size_t)9 * hidden);
    p.attn_ws_bytes = dlka_deformable_lka_attention2d_workspace_bytes(B, C, H, W);
    p.attn_ws = ar.take<char>(p.attn_ws_bytes);
    return ar.ok();
}

}  // namespace

extern "C" {

size_t dlka_deformable_lka_block2d_workspace_bytes(int B, int C, int H, int W, int hidden)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || hidden <= 0) return 0;
    Arena ar(nullptr, 0);
    Lka2dBlockPlan p;
    plan_block(ar, B, C, H, W, hidden, p);
    return ar.off + 256;
}

int dlka_deformable_lka_block2d_forward(const dlkaLkaBlock2dParams *P, const float *x, float *y, int B, int C, int H, int W,
                                        int math, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!P || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (!P->norm1_weight || !P->norm1_bias || !P->norm2_weight || !P->norm2_bias || !P->layer_scale_1 || !P->layer_scale_2 ||
        !P->fc1_weight || !P->fc1_bias || !P->dw_weight || !P->dw_bias || !P->fc2_weight || !P->fc2_bias)
        return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || P->hidden <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0 || P->hidden % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(device_ok());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    Lka2dBlockPlan p;
    if (!plan_block(ar, B, C, H, W, P->hidden, p)) return DLKA_ERR_WORKSPACE;
    const i64 M = (i64)B * H * W;
    const int hid = P->hidden;
    // x1 = x + layer_scale_1 * attn(norm1(x))                                  (MaxViT_deform_LKA.py:168-175)
    DLKA_TRY(layernorm_cl(x, nullptr, 0, P->norm1_weight, P->norm1_bias, p.ln, M, C, P->eps1, st));
    DLKA_TRY(attention2d_cl(&P->attn, p.ln, p.attn, B, C, H, W, math, p.attn_ws, p.attn_ws_bytes, st));
    DLKA_TRY(scale_residual_cl(x, nullptr, 0, P->layer_scale_1, p.attn, p.x1, M, C, st));
    // y = x1 + layer_scale_2 * fc2(GELU(dwconv3x3(fc1(norm2(x1)))))             (:179-185, Mlp :44-52)
    DLKA_TRY(layernorm_cl(p.x1, nullptr, 0, P->norm2_weight, P->norm2_bias, p.ln, M, C, P->eps2, st));
    DLKA_TRY(dense_cl(p.ln, C, M, C, hid, P->fc1_weight, P->fc1_bias, EPI_NONE, nullptr, 0, p.h1, hid, math, p.wp_fc1, st));
    DLKA_TRY(dwconv2d3_cl(p.h1, P->dw_weight, P->dw_bias, p.h2, B, hid, H, W, 1, p.wp_dw, st));
    DLKA_TRY(dense_cl(p.h2, hid, M, hid, C, P->fc2_weight, P->fc2_bias, EPI_NONE, nullptr, 0, p.f2, C, math, p.wp_fc2, st));
    DLKA_TRY(scale_residual_cl(p.x1, nullptr, 0, P->layer_scale_2, p.f2, y, M, C, st));
    return DLKA_OK;
}

size_t dlka_lka_transformer3d_prenorm_workspace_bytes(int B, int C, int D1, int D2, int D3)
{
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return 0;
    const size_t M = (size_t)B * D1 * D2 * D3;
    return dlka_lka_attention3d_deform_workspace_bytes(B, C, D1, D2, D3) + 2 * (M * C * sizeof(float) + 256) + 256;
}

// y = (x + pos) + gamma * LKA_Attention3d_deform(LayerNorm(x + pos))   (transformerblock.py:620-624)
int dlka_lka_transformer3d_prenorm_forward(const dlkaBlock3dParams *attn, const float *norm_weight, const float *norm_bias, float eps,
                                           const float *gamma, const float *pos_embed, const float *x, float *y, int B, int C,
                                           int D1, int D2, int D3, int math, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!attn || !norm_weight || !norm_bias || !gamma || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(device_ok());
    cudaStream_t st = (cudaStream_t)stream;
    const i64 N = (i64)D1 * D2 * D3, M = (i64)B * N;
    Arena ar(workspace, workspace_bytes);
    float *ln = ar.take<float>((size_t)M * C);
    float *a = ar.take<float>((size_t)M * C);
    const size_t aws_bytes = dlka_lka_attention3d_deform_workspace_bytes(B, C, D1, D2, D3);
    void *aws = ar.take<char>(aws_bytes);
    if (!ar.ok()) return DLKA_ERR_WORKSPACE;
    DLKA_TRY(layernorm_cl(x, pos_embed, N, norm_weight, norm_bias, ln, M, C, eps, st));
    DLKA_TRY(dlka_lka_attention3d_deform_forward(attn, ln, a, B, C, D1, D2, D3, math, aws, aws_bytes, stream));
    DLKA_TRY(scale_residual_cl(x, pos_embed, N, gamma, a, y, M, C, st));
    return DLKA_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Whole 3D transformer block on tokens (rows N1 + N3): TransformerBlock_3D_single_deform_LKA.forward
// (transformerblock.py:617-630), inference mode:
//   a  = x' + gamma * Attn(LayerNorm(x')),  x' = x + pos_embed
//   r  = LeakyReLU(BN1(conv1(a)));  r = LeakyReLU(BN2(conv2(r)) + a)        UnetResBlock (dynunet_block.py:65-80)
//   y  = a + conv8(r)                                                        Dropout3d is the identity in eval mode
// BatchNorm is passed already folded to per-channel (scale, shift) = (w/sqrt(var+eps), b - mean*scale).
// ---------------------------------------------------------------------------------------------------------
extern "C" {

size_t dlka_lka_transformer3d_block_workspace_bytes(int B, int C, int D1, int D2, int D3)
{
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return 0;
    const size_t M = (size_t)B * D1 * D2 * D3;
    return dlka_lka_transformer3d_prenorm_workspace_bytes(B, C, D1, D2, D3) + 3 * (M * C * sizeof(float) + 256) +
           (2 * conv3_scratch_floats(C) + dense_scratch_floats(C, C)) * sizeof(float) + 4 * 256;
}

int dlka_lka_transformer3d_block_forward(const dlkaTransformer3dParams *P, const float *x, float *y, int B, int C, int D1, int D2,
                                         int D3, int math, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!P || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (!P->norm_weight || !P->norm_bias || !P->gamma || !P->conv1_weight || !P->bn1_scale || !P->bn1_shift || !P->conv2_weight ||
        !P->bn2_scale || !P->bn2_shift || !P->conv8_weight || !P->conv8_bias)
        return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(device_ok());
    cudaStream_t st = (cudaStream_t)stream;
    const i64 M = (i64)B * D1 * D2 * D3;
    Arena ar(workspace, workspace_bytes);
    float *a = ar.take<float>((size_t)M * C), *r1 = ar.take<float>((size_t)M * C), *r2 = ar.take<float>((size_t)M * C);
    float *wp1 = ar.take<float>(conv3_scratch_floats(C)), *wp2 = ar.take<float>(conv3_scratch_floats(C));
    float *wp8 = ar.take<float>(dense_scratch_floats(C, C));
    const size_t pws_bytes = dlka_lka_transformer3d_prenorm_workspace_bytes(B, C, D1, D2, D3);
    void *pws = ar.take<char>(pws_bytes);
    if (!ar.ok()) return DLKA_ERR_WORKSPACE;
    DLKA_TRY(dlka_lka_transformer3d_prenorm_forward(&P->attn, P->norm_weight, P->norm_bias, P->eps, P->gamma, P->pos_embed, x, a, B, C,
                                                    D1, D2, D3, math, pws, pws_bytes, stream));
    DLKA_TRY(conv3_bn_act_cl(a, P->conv1_weight, P->bn1_scale, P->bn1_shift, 1, P->lrelu_slope, nullptr, r1, B, C, D1, D2, D3, math,
                             wp1, st));
    DLKA_TRY(conv3_bn_act_cl(r1, P->conv2_weight, P->bn2_scale, P->bn2_shift, 2, P->lrelu_slope, a, r2, B, C, D1, D2, D3, math, wp2,
                             st));
    DLKA_TRY(dense_cl(r2, C, M, C, C, P->conv8_weight, P->conv8_bias, EPI_ADD, a, C, y, C, math, wp8, st));
    return DLKA_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Rest of row N3 (2D decoder, 2D/networks/MaxViT_deform_LKA.py:488-620): the token Linear of MyDecoderLayer (x1_linear, with
// the skip tensor added in the epilogue, :604-607) and PatchExpand / FinalPatchExpand_X4 (Linear without bias -> pixel shuffle
// -> LayerNorm, :488-545) as one call each.  The Linear runs on the tcgen05 dense kernel, the shuffle is an index remap
// inside the LayerNorm kernel: no rearranged copy is ever materialised.
// ---------------------------------------------------------------------------------------------------------
extern "C" {

size_t dlka_linear_tokens_workspace_bytes(int K, int N)
{
    if (K <= 0 || N <= 0) return 0;
    return dense_scratch_floats(N, K) * sizeof(float) + 512;
}

// y[M,N] = x[M,K] @ weight[N,K]^T (+ bias) (+ add[M,N])
int dlka_linear_tokens_forward(const float *x, const float *weight, const float *bias, const float *add, float *y, long long M, int K,
                               int N, int math, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!x || !weight || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (M <= 0 || K <= 0 || N <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (K % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(device_ok());
    Arena ar(workspace, workspace_bytes);
    float *wp = ar.take<float>(dense_scratch_floats(N, K));
    if (!ar.ok()) return DLKA_ERR_WORKSPACE;
    return dense_cl(x, K, (i64)M, K, N, weight, bias, add ? EPI_ADD : EPI_NONE, add, N, y, N, math, wp, (cudaStream_t)stream);
}

size_t dlka_patch_expand2d_workspace_bytes(int B, int H, int W, int dim, int scale)
{
    if (B <= 0 || H <= 0 || W <= 0 || dim <= 0 || (scale != 2 && scale != 4)) return 0;
    const size_t M = (size_t)B * H * W, F = scale == 2 ? 2 * (size_t)dim : 16 * (size_t)dim;
    return M * F * sizeof(float) + dense_scratch_floats((int)F, dim) * sizeof(float) + 1024;
}

// scale 2: PatchExpand.forward          x [B, H*W, dim] -> y [B, 4*H*W, dim/2]     (expand_weight [2*dim, dim], norm over dim/2)
// scale 4: FinalPatchExpand_X4.forward  x [B, H*W, dim] -> y [B, 16*H*W, dim]      (expand_weight [16*dim, dim], norm over dim)
int dlka_patch_expand2d_forward(const float *x, const float *expand_weight, const float *norm_weight, const float *norm_bias, float eps,
                                float *y, int B, int H, int W, int dim, int scale, int math, void *workspace, size_t workspace_bytes,
                                void *stream)
{
    if (!x || !expand_weight || !norm_weight || !norm_bias || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || H <= 0 || W <= 0 || dim <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (scale != 2 && scale != 4) return DLKA_ERR_UNSUPPORTED;
    const int cg = scale == 2 ? dim / 2 : dim, F = scale * scale * cg;
    if (dim % 8 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(device_ok());
    cudaStream_t st = (cudaStream_t)stream;
    const i64 M = (i64)B * H * W;
    Arena ar(workspace, workspace_bytes);
    float *e = ar.take<float>((size_t)M * F);
    float *wp = ar.take<float>(dense_scratch_floats(F, dim));
    if (!ar.ok()) return DLKA_ERR_WORKSPACE;
    DLKA_TRY(dense_cl(x, dim, M, dim, F, expand_weight, nullptr, EPI_NONE, nullptr, 0, e, F, math, wp, st));
    return layernorm_shuffle_cl(e, norm_weight, norm_bias, y, B, H, W, scale, cg, eps, st);
}

}  // extern "C"
