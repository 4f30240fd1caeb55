// kernels.cuh -- internal launcher interface between api.cu and the kernel translation units.
// Internal activation layout is channels-last fp32: X[b][d][h][w][c] (2D: d = 0).
#pragma once
#include "common.cuh"

namespace dlka {

// ---------------- implicit GEMM (SIMT fp32 and tensor-core variants share the argument block) ----
enum IgemmMode { IGEMM_DENSE = 0, IGEMM_CONV = 1, IGEMM_DEFORM = 2 };
enum EpiMode { EPI_NONE = 0, EPI_GELU = 1, EPI_MUL = 2, EPI_ADD = 3 };

struct IgemmArgs {
    int mode;          // IgemmMode
    int epi;           // EpiMode
    ConvGeo geo;       // input/output geometry (dense: only C, Co, groups=1 matter)
    i64 M;             // output rows = B*Do*Ho*Wo
    int Ktot;          // taps * C/groups
    int Npad;          // padded columns of Wp per group
    const float *X;    // input, channels-last; dense: row stride ldX
    int ldX;
    i64 xch;           // 0: X is channels-last [voxel][C]; else X is CHUNK-MAJOR [C/32][voxel][32] with xch floats between chunks
    const float *Off;  // deformable: offsets [M][ldOff], first dg*ndim*K columns used
    int ldOff;         // row stride of Off (0 = dg*ndim*K)
    const float *Mask; // deformable 2D (DCNv2) modulation [M][dg*K] or null
    const float *Wp;   // packed weights [groups][Ktot][Npad]
    const float *bias; // [Co] or null
    const float *E;    // epilogue operand (gate / residual) [M][ldE] or null
    int ldE;
    float *Y;          // output [M][ldY]
    int ldY;
    int ybrick;        // conv_tiled, 3D only: write Y BRICK-MAJOR [brick = (b, d/4, h/4, w/8)][Co][128 rows] (the offsets of deform_ps.cu)
    // split-K (tcgen05 dense path only): grid.z slices of `ksplit_steps` K steps each write partial products to
    // Y + z * ysplit_stride (bias / epilogue operand applied by slice 0 only); 0 = no split
    int ksplit_steps;
    i64 ysplit_stride;
    // conv_tiled: optional scratch for K-split partial outputs (small problems: few CTAs, long K loop); the split and the
    // reduction happen inside conv_tiled_ex when the scratch is large enough
    float *split_scratch;
    i64 split_scratch_floats;
};

int igemm_simt_npad(int n_per_group);
int pack_weight(const float *w, float *wp, int Co, int Cg, int taps, int groups, int Npad, cudaStream_t st);
int igemm_simt(const IgemmArgs &a, cudaStream_t st);

// tensor-core (tcgen05, bf16 hi/lo split) variant -- mma_tc.cu
bool tc_supported(const IgemmArgs &a);
// persistent streaming variant for the HBM-bound 1x1 projections (dense_stream.cu): whole raw tiles in flight through
// cp.async.bulk, weights resident in shared memory; weights packed with deform3d_ps_pack(w, bp, Co, C, 1)
bool dense_stream_supported(const IgemmArgs &a);
int dense_stream(const IgemmArgs &a, const void *w_packed, cudaStream_t st);
int tc_kc(int C);
int tc_nt(int Co);
size_t tc_packed_weight_bytes(int Co, int C, int taps);
int tc_pack_weight(const float *w, void *bp, int Co, int C, int taps, cudaStream_t st);
int igemm_tc(const IgemmArgs &a, const void *bp, cudaStream_t st);

// 3D deformable conv, brick tiles + chunk-major K (L1-resident gather) -- deform_tc.cu
// optional fused epilogue chain: conv1 (1x1) * U  [-> proj_2 (1x1) + R]; weights in tc_pack_weight layout
struct DeformChain {
    int stages;           // 1 or 2
    const void *W1p; const float *b1; const float *U; int ldU;
    const void *W2p; const float *b2; const float *R; int ldR;
};
bool deform3d_tc_supported(const IgemmArgs &a);
bool deform3d_chain_supported(const IgemmArgs &a);
int deform3d_tc(const IgemmArgs &a, const float *w, void *bp, const DeformChain *chain, cudaStream_t st);

// persistent variant with a chunk-major gather source, double-buffered accumulators and the 1x1 chain fed from tensor memory
// (deform_ps.cu).  Weights (main and chain) are packed once with deform3d_ps_pack; X / xch as in IgemmArgs::xch.
bool deform3d_ps_supported(const IgemmArgs &a, int chain_stages);
size_t deform3d_ps_packed_bytes(int Co, int C, int taps);
int deform3d_ps_pack(const float *w, void *bp, int Co, int C, int taps, cudaStream_t st);
int deform3d_ps(const IgemmArgs &a, i64 xch, const void *bp, const DeformChain *chain, int off_mode, cudaStream_t st);
size_t deform3d_ps_offset_floats(int B, int D, int H, int W, int cols);   // brick-major offset buffer (off_mode 1)

// zero-copy tiled regular conv on tcgen05 (stride 1, groups 1, no epilogue operand) -- conv_tc.cu
bool conv_tiled_supported(const IgemmArgs &a);
size_t conv_tiled_packed_bytes(int Co, int C, int taps);
int conv_tiled(const IgemmArgs &a, const float *w, void *bp, cudaStream_t st);
int conv_tiled_ex(const IgemmArgs &a, const float *w, const float *wscale, int act, float slope, const float *E, int ldE, void *bp,
                  cudaStream_t st);

// ---------------- layout ----------------
// [B][C][S] -> [B][S][C]  and back (S = spatial size)
int transpose_cs_to_sc(const float *in, float *out, int B, int C, i64 S, cudaStream_t st);
int transpose_sc_to_cs(const float *in, float *out, int B, int C, i64 S, cudaStream_t st);
// [B][C][S] -> chunk-major [C/32][B][S][32] (C % 32 == 0)
int transpose_cs_to_chunk(const float *in, float *out, int B, int C, i64 S, cudaStream_t st);

// ---------------- depthwise (regular) conv, channels-last, "same" output extent, stride 1 ----------
// w: PyTorch layout [C][1][kd][kh][kw]; bias [C] or null.  pad = dil*(k-1)/2 on each axis.
// dd / dil: dilation along d / along h and w (the ACDC variant of the block is anisotropic, acdc/transformerblock.py:214-236)
// chunk_major_out: write y as [C/32][B][D][H][W][32] (only with the shared-memory variant; DLKA_ERR_UNSUPPORTED otherwise)
int dwconv_cl(const float *x, const float *w, const float *bias, float *y, int B, int C, int D, int H, int W, int kd,
              int kh, int kw, int dd, int dil, float *w_packed /*[K][C]*/, cudaStream_t st, bool chunk_major_out = false);

// shared-memory plane-streaming variant (C % 32 == 0; the five stencil shapes of the synapse / acdc blocks) -- dwconv_smem.cu
bool dwconv_smem_supported(int C, int kd, int kh, int kw, int dd, int dh, int dw);
int dwconv_smem(const float *x, const float *w_packed, const float *bias, float *y, int B, int C, int D, int H, int W, int kd, int k,
                int dd, int dil, cudaStream_t st, bool chunk_major_out = false);

// ---------------- depthwise deformable conv (groups == C == Co), channels-last ----------------------
// w: [C][1][taps] PyTorch layout; Off [M][dg*ndim*K]; Mask optional; bias optional.
int deform_dwconv_cl(const float *x, const float *off, const float *mask, const float *w, const float *bias, float *y,
                     const ConvGeo &g, float *w_packed /*[K][C]*/, cudaStream_t st);

// ---------------- layers around the attention block (row N1) -- norm_mlp.cu ----------------
int layernorm_cl(const float *x, const float *pos, i64 pos_rows, const float *gamma, const float *beta, float *y, i64 M, int C,
                 float eps, cudaStream_t st);
// LayerNorm over C of [B*H*W][P*P][C] rows, written pixel-shuffled to tokens [B][(H*P)*(W*P)][C] (PatchExpand, MaxViT_deform_LKA.py:488-545)
int layernorm_shuffle_cl(const float *x, const float *gamma, const float *beta, float *y, int B, int H, int W, int P, int C, float eps,
                         cudaStream_t st);
int scale_residual_cl(const float *x, const float *pos, i64 pos_rows, const float *scale, const float *y, float *out, i64 M, int C,
                      cudaStream_t st);
int dwconv2d3_cl(const float *x, const float *w, const float *bias, float *y, int B, int C, int H, int W, int gelu, float *w_packed,
                 cudaStream_t st);
int affine_act_cl(float *y, const float *scale, const float *shift, const float *e, i64 M, int C, int act, float slope,
                  cudaStream_t st);

// ---------------- internal block-level helpers exported by api.cu for blocks_api.cu ----------------
int attention2d_cl(const dlkaBlock2dParams *params, const float *x_cl, float *y_cl, int B, int C, int H, int W, int math,
                   void *workspace, size_t workspace_bytes, cudaStream_t st);
// split-K variant: partial[z][M][Co] for z < cdiv(Ci / KC, ksplit_steps); returns the number of slices in *nsplit (tcgen05 path only)
int dense_splitk_cl(const float *x, int ldX, i64 M, int Ci, int Co, const float *w, float *partial, int ksplit_steps, int *nsplit,
                    float *wscratch, cudaStream_t st);
int dense_cl(const float *x, int ldX, i64 M, int Ci, int Co, const float *w, const float *bias, int epi, const float *E, int ldE,
             float *y, int ldY, int math, float *wscratch, cudaStream_t st);
size_t dense_scratch_floats(int Co, int Ci);

// ---------------- backward of the 3D deformable conv (deform_bwd.cu), channels-last, groups = dg = 1 ----------------
int reduce_partials(const float *partial, float *out, i64 n, int nsplit, cudaStream_t st);   // out = sum_z partial[z]
int deform3d_bwd_chunk_rows(i64 M);   // rows per streamed chunk (multiple of 64)
int deform3d_backward_cl(const ConvGeo &g, const float *x, const float *off, const float *w, const float *gout, float *gin, float *goff,
                         float *gw, float *gb, float *wt, float *gwt, float *colbuf, float *colT, float *gchunk, float *gchunkT,
                         float *partial /* [Mc/512 + 1][K*C*Co] split-K partial products */, float *wscratch, int math, cudaStream_t st);

// ---------------- backward of the 2D deformable conv (deform2d_bwd.cu), channels-last, every forward configuration ----------------
int deform2d_backward_cl(const ConvGeo &g, const float *x, const float *w, const float *off, const float *mask, const float *gout,
                         float *gx, float *gw, float *goff, float *gmask, float *gb, cudaStream_t st);

// dense 3x3x3 conv C->C (stride 1, pad 1) on channels-last tokens with folded per-channel scale/shift and
// LeakyReLU (+ residual): the two convolutions of UnetResBlock (row N3).  SIMT fp32 fallback when math != bf16x3.
int conv3_bn_act_cl(const float *x, const float *w, const float *scale, const float *shift, int act, float slope, const float *E,
                    float *y, int B, int C, int D1, int D2, int D3, int math, float *wscratch, cudaStream_t st);
size_t conv3_scratch_floats(int C);
int device_ok();

// ---------------- sampler integer planes (parity K4) ----------------
int sample_indices(const float *off_cf, int32_t *low, int32_t *mask, const ConvGeo &g, cudaStream_t st);

}  // namespace dlka
