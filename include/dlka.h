/*
 * dlka.h -- C ABI of libdlka_b200.so: B200 (sm_100a) Deformable-LKA forward operators.
 *
 * Drop-in boundary for the D-LKA hot path of xmindflow/deformableLKA.  Every entry point
 * names the reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All tensors are fp32, contiguous.
 *   - pointers are DEVICE pointers unless the function name ends in `_host`.
 *   - the library allocates nothing on the device: the caller passes a workspace whose size
 *     is returned by the matching `*_workspace_bytes` query (0 is a legal answer).
 *   - `stream` is a cudaStream_t passed as void* (the reference launches on the caller's
 *     current stream: 3D/dcn/src/cuda/deform_conv_cuda.cu:97).
 *   - return value: DLKA_OK (0) or a negative dlkaStatus; `dlka_status_string` gives the text the
 *     Python host raises as RuntimeError (the reference raises c10::Error -> RuntimeError,
 *     3D/dcn/src/cuda/deform_conv_cuda.cu:41-76).
 *   - there is NO CPU path: without a CUDA device every compute entry point returns
 *     DLKA_ERR_NO_DEVICE (the reference's 3D op is CUDA-only too: 3D/dcn/src/deform_conv.h:46).
 *   - 64-bit addressing throughout (the reference's int32 index math overflows at
 *     (2,96,64,128,128): 3D/dcn/src/cuda/deform_im2col_cuda.cuh:228).
 */
#ifndef DLKA_H_
#define DLKA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLKA_VERSION 100 /* major*1000 + minor*10 */

#if defined(__GNUC__)
#define DLKA_API __attribute__((visibility("default")))
#else
#define DLKA_API
#endif

typedef enum dlkaStatus {
    DLKA_OK = 0,
    DLKA_ERR_INVALID_ARGUMENT = -1, /* bad shape / null pointer / size mismatch            */
    DLKA_ERR_UNSUPPORTED = -2,      /* legal in the reference, not implemented here (loud) */
    DLKA_ERR_WORKSPACE = -3,        /* workspace missing or too small                      */
    DLKA_ERR_NO_DEVICE = -4,        /* no CUDA device / not sm_100                         */
    DLKA_ERR_CUDA = -5              /* a CUDA runtime call or launch failed                */
} dlkaStatus;

/* Arithmetic used for the channel contractions (1x1 projections, offset conv, deformable GEMM).
 * Sampling positions / interpolation weights / depthwise stencils are always fp32.           */
typedef enum dlkaMath {
    DLKA_MATH_FP32_SIMT = 0, /* fp32 FMA on CUDA cores (validation path, slow)                       */
    DLKA_MATH_BF16X3 = 1     /* tcgen05 tensor cores, bf16 hi/lo split (3 MMAs), fp32 accumulate     */
} dlkaMath;

DLKA_API int dlka_version(void);
DLKA_API const char *dlka_status_string(int status);
/* last CUDA error text seen by this thread (empty string if none) */
DLKA_API const char *dlka_last_cuda_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
DLKA_API uint64_t dlka_launch_count(void);
/* Optional per-kernel timing: while enabled every kernel launch is bracketed by CUDA events on its
 * stream; dlka_profile_summary synchronises them and writes "name launches total_ms\n" lines.  */
DLKA_API int dlka_profile_enable(int on);
DLKA_API int dlka_profile_summary(char *buf, size_t buf_bytes);

/* ------------------------------------------------------------------------------------------
 * Operator: 3D deformable convolution forward.
 * Replaces  D3D.deform_conv_forward  (3D/dcn/src/vision.cpp:4-7, 3D/dcn/src/deform_conv.h:10-47,
 *           3D/dcn/src/cuda/deform_conv_cuda.cu:18-126) as called by DeformConvFunction.forward
 *           (3D/dcn/functions/deform_conv_func.py:17-36).
 *   input  [B, C, D, H, W]            weight [Co, C/group, kd, kh, kw]     bias [Co] (required)
 *   offset [B, dg*3*K, Do, Ho, Wo]    channel 3t+{0,1,2} = (dd,dh,dw) of tap t=(i*kh+j)*kw+k
 *   output [B, Co, Do, Ho, Wo]        Do = (D + 2*pd - (dild*(kd-1)+1))/sd + 1  (cu:78-80)
 * `im2col_step` is accepted for signature parity; only its divisibility rule is enforced
 * (batch % min(batch, im2col_step) == 0, cu:61-63): no im2col buffer exists here.
 * ------------------------------------------------------------------------------------------ */
DLKA_API size_t dlka_deform_conv3d_workspace_bytes(int B, int C, int D, int H, int W, int Co,
                                          int kd, int kh, int kw, int sd, int sh, int sw,
                                          int pd, int ph, int pw, int dild, int dilh, int dilw,
                                          int group, int deformable_group);
DLKA_API int dlka_deform_conv3d_forward(const float *input, const float *weight, const float *bias,
                               const float *offset, float *output,
                               int B, int C, int D, int H, int W, int Co,
                               int kd, int kh, int kw, int sd, int sh, int sw,
                               int pd, int ph, int pw, int dild, int dilh, int dilw,
                               int group, int deformable_group, int im2col_step, int math,
                               void *workspace, size_t workspace_bytes, void *stream);

/* Backward of the same operator (row N2).  Replaces D3D.deform_conv_backward
 * (3D/dcn/src/vision.cpp:6, deform_conv.h:49-86, cuda/deform_conv_cuda.cu:128-285): grad_input [B,C,D,H,W],
 * grad_offset [B,3*K,Do,Ho,Wo], grad_weight [Co,C,kd,kh,kw], grad_bias [Co], all fully overwritten.
 * groups == deformable groups == 1 and C % 4 == 0 (the D-LKA configuration); anything else is DLKA_ERR_UNSUPPORTED.
 * The gradients are those of the forward definition; the reference's pad_h/pad_w index defect in
 * deformable_col2im_coord (deform_im2col_cuda.cuh:448) is not reproduced.                                          */
DLKA_API size_t dlka_deform_conv3d_backward_workspace_bytes(int B, int C, int D, int H, int W, int Co,
                                                   int kd, int kh, int kw, int sd, int sh, int sw,
                                                   int pd, int ph, int pw, int dild, int dilh, int dilw,
                                                   int group, int deformable_group);
DLKA_API int dlka_deform_conv3d_backward(const float *input, const float *weight, const float *offset, const float *grad_output,
                                float *grad_input, float *grad_offset, float *grad_weight, float *grad_bias,
                                int B, int C, int D, int H, int W, int Co, int kd, int kh, int kw, int sd, int sh, int sw,
                                int pd, int ph, int pw, int dild, int dilh, int dilw, int group, int deformable_group,
                                int im2col_step, int math, void *workspace, size_t workspace_bytes, void *stream);

/* Integer planes of the 3D sampler for bit-exact parity checks:
 *   low [B*dg, Vo, K, 3] int32 = floor(p) per axis,  mask [B*dg, Vo, K] int32:
 *   bit0 = sample valid (cuh:248), bits 1..8 = corner v1..v8 read (cuh:43-65).
 * Uses the same device function as the convolution kernels.                                  */
DLKA_API int dlka_deform_conv3d_sample_indices(const float *offset, int32_t *low, int32_t *mask,
                                      int B, int D, int H, int W, int kd, int kh, int kw,
                                      int sd, int sh, int sw, int pd, int ph, int pw,
                                      int dild, int dilh, int dilw, int deformable_group, void *stream);

/* ------------------------------------------------------------------------------------------
 * Operator: 3D DeformConvPack forward = conv_offset (regular conv, dilation 1) + deformable conv.
 * Replaces  DeformConvPack.forward  (3D/d_lka_former/network_architecture/synapse/deform_conv.py:93-105;
 *           identical 3D/dcn/modules/deform_conv.py:88-100).
 *   offset_weight [dg*3*K, C, kd, kh, kw], offset_bias [dg*3*K]; other arguments as the 3D operator.
 * ------------------------------------------------------------------------------------------ */
DLKA_API size_t dlka_deform_conv_pack3d_workspace_bytes(int B, int C, int D, int H, int W, int Co,
                                               int kd, int kh, int kw, int sd, int sh, int sw,
                                               int pd, int ph, int pw, int dild, int dilh, int dilw,
                                               int group, int deformable_group);
DLKA_API int dlka_deform_conv_pack3d_forward(const float *input, const float *offset_weight, const float *offset_bias,
                                    const float *weight, const float *bias, float *output,
                                    int B, int C, int D, int H, int W, int Co,
                                    int kd, int kh, int kw, int sd, int sh, int sw,
                                    int pd, int ph, int pw, int dild, int dilh, int dilw,
                                    int group, int deformable_group, int im2col_step, int math,
                                    void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Operator: 2D deformable convolution forward (DCNv1 when mask == NULL, DCNv2 otherwise).
 * Replaces  torch.ops.torchvision.deform_conv2d  as called from
 *           2D/deformable_LKA/deformable_LKA.py:18-25,29 (torchvision/ops/deform_conv.py:92-107).
 *   input [B,C,H,W]  weight [Co, C/n_weight_grps, kh, kw]  offset [B, n_offset_grps*2*K, Ho, Wo]
 *   (channel 2t = dy, 2t+1 = dx, t = i*kw+j)   mask [B, n_offset_grps*K, Ho, Wo] or NULL
 *   bias [Co] or NULL    output [B, Co, Ho, Wo]
 * ------------------------------------------------------------------------------------------ */
DLKA_API size_t dlka_deform_conv2d_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw,
                                          int sh, int sw, int ph, int pw, int dilh, int dilw,
                                          int n_weight_grps, int n_offset_grps);
DLKA_API int dlka_deform_conv2d_forward(const float *input, const float *weight, const float *offset,
                               const float *mask, const float *bias, float *output,
                               int B, int C, int H, int W, int Co, int kh, int kw,
                               int sh, int sw, int ph, int pw, int dilh, int dilw,
                               int n_weight_grps, int n_offset_grps, int math,
                               void *workspace, size_t workspace_bytes, void *stream);
/* Backward of the 2D operator: the gradients torchvision's autograd returns for deform_conv2d
 * (torchvision/ops/deform_conv.py:92-107; published kernels deformable_col2im / deformable_col2im_coord + the weight GEMM):
 * grad_input [B,C,H,W], grad_weight [Co,C/g,kh,kw], grad_offset like offset, grad_mask like mask (both NULL without a mask),
 * grad_bias [Co] or NULL.  Every forward configuration (weight groups incl. depthwise, offset groups, DCNv2 mask); fp32.   */
DLKA_API size_t dlka_deform_conv2d_backward_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw,
                                                   int sh, int sw, int ph, int pw, int dilh, int dilw,
                                                   int n_weight_grps, int n_offset_grps, int has_mask);
DLKA_API int dlka_deform_conv2d_backward(const float *input, const float *weight, const float *offset, const float *mask,
                                const float *grad_output, float *grad_input, float *grad_weight, float *grad_offset,
                                float *grad_mask, float *grad_bias, int B, int C, int H, int W, int Co, int kh, int kw,
                                int sh, int sw, int ph, int pw, int dilh, int dilw, int n_weight_grps, int n_offset_grps,
                                void *workspace, size_t workspace_bytes, void *stream);
DLKA_API int dlka_deform_conv2d_sample_indices(const float *offset, int32_t *low, int32_t *mask,
                                      int B, int H, int W, int kh, int kw, int sh, int sw,
                                      int ph, int pw, int dilh, int dilw, int n_offset_grps, void *stream);

/* ------------------------------------------------------------------------------------------
 * Operator: 2D DeformConv wrapper forward = offset_net (regular conv with the SAME kernel / stride /
 * padding / dilation) + deformable conv without mask.
 * Replaces  DeformConv.forward  (2D/deformable_LKA/deformable_LKA.py:27-30).
 *   offset_weight [2*K, C, kh, kw], offset_bias [2*K]; weight [Co, C/groups, kh, kw]; bias [Co] or NULL.
 * ------------------------------------------------------------------------------------------ */
DLKA_API size_t dlka_deform_conv_pack2d_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw,
                                               int sh, int sw, int ph, int pw, int dilh, int dilw, int groups);
DLKA_API int dlka_deform_conv_pack2d_forward(const float *input, const float *offset_weight, const float *offset_bias,
                                    const float *weight, const float *bias, float *output,
                                    int B, int C, int H, int W, int Co, int kh, int kw,
                                    int sh, int sw, int ph, int pw, int dilh, int dilw, int groups, int math,
                                    void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Block: 3D D-LKA.  Parameters in the reference's state_dict layout (device pointers).
 * Replaces  LKA3d_deform.forward            (3D/d_lka_former/network_architecture/synapse/transformerblock.py:644-652)
 *      and  LKA_Attention3d_deform.forward  (same file :664-673).
 * The attention parameters (proj_1 / proj_2) are ignored by dlka_lka3d_deform_forward.
 * ------------------------------------------------------------------------------------------ */
/* Depthwise stencil shapes of the block.  NULL in dlkaBlock3dParams.dw_geom selects the synapse network's
 * (transformerblock.py:637-638: 5^3 dil 1, 7^3 dil 3).  The ACDC network uses the same block with shapes chosen per
 * channel count (3D/d_lka_former/network_architecture/acdc/transformerblock.py:214-236):
 *   dim 32, 64: conv0 5^3;  conv_spatial (5,7,7) dil (3,3,3)      dim 128: conv0 5^3;  conv_spatial (3,5,5) dil (1,3,3)
 *   dim 256:    conv0 3^3;  conv_spatial 3^3 dil 1
 * Axes are (D1, D2, D3) as nn.Conv3d sees them; kernels are odd, padding is dil*(k-1)/2 ("same"), and the library
 * requires k[1]==k[2], dil[1]==dil[2] (true for every shape the reference uses).                                  */
typedef struct dlkaDwGeom3d {
    int conv0_k[3], conv0_dil[3];
    int conv_spatial_k[3], conv_spatial_dil[3];
} dlkaDwGeom3d;

typedef struct dlkaBlock3dParams {
    const float *proj_1_weight, *proj_1_bias;             /* [C,C,1,1,1], [C]   transformerblock.py:659 */
    const float *conv0_weight, *conv0_bias;               /* [C,1,5,5,5], [C]   :637 (pad 2)            */
    const float *conv_spatial_weight, *conv_spatial_bias; /* [C,1,7,7,7], [C]   :638 (pad 9, dil 3)     */
    const float *conv_offset_weight, *conv_offset_bias;   /* [81,C,3,3,3], [81] synapse/deform_conv.py:80-85 */
    const float *deform_weight, *deform_bias;             /* [C,C,3,3,3], [C]   synapse/deform_conv.py:37-39 */
    const float *conv1_weight, *conv1_bias;               /* [C,C,1,1,1], [C]   :641                    */
    const float *proj_2_weight, *proj_2_bias;             /* [C,C,1,1,1], [C]   :662                    */
    const dlkaDwGeom3d *dw_geom;                          /* host pointer; NULL = synapse shapes above  */
} dlkaBlock3dParams;

/* x, y: [B, C, D1, D2, D3] (NCDHW as nn.Conv3d sees it). */
DLKA_API size_t dlka_lka3d_deform_workspace_bytes(int B, int C, int D1, int D2, int D3);
DLKA_API int dlka_lka3d_deform_forward(const dlkaBlock3dParams *params, const float *x, float *y,
                              int B, int C, int D1, int D2, int D3, int math,
                              void *workspace, size_t workspace_bytes, void *stream);

/* x, y: tokens [B, N, C] with N = D1*D2*D3 (the reference calls the three axes H, W, D; token
 * n = (i1*D2 + i2)*D3 + i3, i.e. channels-last over the Conv3d volume [D1,D2,D3]).            */
DLKA_API size_t dlka_lka_attention3d_deform_workspace_bytes(int B, int C, int D1, int D2, int D3);
DLKA_API int dlka_lka_attention3d_deform_forward(const dlkaBlock3dParams *params, const float *x, float *y,
                                        int B, int C, int D1, int D2, int D3, int math,
                                        void *workspace, size_t workspace_bytes, void *stream);
/* Prepacked weights.  The reference keeps its weights in the layout its GEMM consumes (deform_conv_cuda.cu:85,111-117); here the
 * packed bf16 hi/lo operand tiles of a block live in a caller-owned buffer of *_packed_bytes(C) bytes that persists across
 * calls.  packed_valid = 0: pack the weights of `params` into `packed`, then run;  packed_valid = 1: the caller asserts that
 * `packed` was filled by an earlier call with the SAME parameter values, shape (B, C, D1, D2, D3), math and device -- no packing
 * kernel is launched.  (The Python host keeps one buffer per module and shape, keyed on the parameters' data pointers and
 * version counters: ops.lka_attention3d_deform_forward(..., cache=).)                                                      */
DLKA_API size_t dlka_lka_attention3d_deform_packed_bytes(int C);
DLKA_API int dlka_lka_attention3d_deform_forward_packed(const dlkaBlock3dParams *params, const float *x, float *y,
                                               int B, int C, int D1, int D2, int D3, int math,
                                               void *packed, size_t packed_bytes, int packed_valid,
                                               void *workspace, size_t workspace_bytes, void *stream);
/* Same call with HOST buffers (x_host, y_host pinned or pageable): H2D, compute, D2H on `stream`,
 * then stream-synchronised.  `params` still holds device pointers; `dev_scratch` must hold
 * 2*B*N*C floats in addition to the workspace (x and y staging).                               */
DLKA_API int dlka_lka_attention3d_deform_forward_host(const dlkaBlock3dParams *params, const float *x_host,
                                             float *y_host, int B, int C, int D1, int D2, int D3, int math,
                                             void *dev_scratch, size_t dev_scratch_bytes,
                                             void *workspace, size_t workspace_bytes, void *stream);

/* Streaming variant: a pipeline context keeps `depth` steps in flight, so the H2D copy of step k+1 and the
 * D2H copy of step k-1 overlap the compute of step k.  `dev_scratch` must hold depth * 2 * B*N*C floats.
 * The async call returns without host synchronisation; x_host / y_host must stay valid until
 * dlka_host_pipe_wait().  dlka_host_pipe_join() orders `stream` after every D2H copy enqueued so far.   */
typedef struct dlkaHostPipe dlkaHostPipe;
DLKA_API int dlka_host_pipe_create(dlkaHostPipe **pipe, int depth);
DLKA_API int dlka_host_pipe_destroy(dlkaHostPipe *pipe);
DLKA_API int dlka_host_pipe_wait(dlkaHostPipe *pipe);
DLKA_API int dlka_host_pipe_join(dlkaHostPipe *pipe, void *stream);
DLKA_API int dlka_lka_attention3d_deform_forward_host_async(dlkaHostPipe *pipe, const dlkaBlock3dParams *params,
                                                   const float *x_host, float *y_host, int B, int C, int D1, int D2,
                                                   int D3, int math, void *dev_scratch, size_t dev_scratch_bytes,
                                                   void *workspace, size_t workspace_bytes, void *stream);
/* non-blocking lower bound of the number of submitted steps whose last D2H copy has finished (their host buffers may be
 * released); negative = dlkaStatus */
DLKA_API long long dlka_host_pipe_completed(dlkaHostPipe *pipe);

/* Host-side placement for the `_host` entries (the reference's only multi-GPU mode is nn.DataParallel out of un-placed
 * host tensors, 2D/trainer_MaxViT_deform_LKA.py:60-66; at 805 MB per direction per step the copy rate is the end-to-end rate).
 * dlka_host_numa_node: NUMA node of CUDA device `device` from sysfs, -1 if unknown.
 * dlka_host_bind_thread: pin the calling thread to that node's cores and prefer it for page allocation; returns the node or -1.
 * dlka_host_alloc: page-locked buffer, policy 0 = no placement, 1 = on the device's node, 2 = interleaved over all nodes;
 *                  mmap + mbind before first touch + cudaHostRegister(portable).  Release with dlka_host_free.           */
DLKA_API int dlka_host_numa_node(int device);
DLKA_API int dlka_host_bind_thread(int device);
DLKA_API int dlka_host_alloc(void **ptr, size_t bytes, int device, int policy);
DLKA_API int dlka_host_free(void *ptr);

/* ------------------------------------------------------------------------------------------
 * Block: 2D D-LKA.
 * Replaces  deformable_LKA.forward            (2D/deformable_LKA/deformable_LKA.py:98-104)
 *      and  deformable_LKA_Attention.forward  (2D/deformable_LKA/deformable_LKA.py:133-140).
 * ------------------------------------------------------------------------------------------ */
typedef struct dlkaBlock2dParams {
    const float *proj_1_weight, *proj_1_bias;                   /* [C,C,1,1], [C]        :127 */
    const float *conv0_offset_weight, *conv0_offset_bias;       /* [50,C,5,5], [50]      :10-16,93 */
    const float *conv0_deform_weight;                           /* [C,1,5,5] (no bias)   :18-25 */
    const float *conv_spatial_offset_weight, *conv_spatial_offset_bias; /* [98,C,7,7], [98] (dil 3, pad 9) :94 */
    const float *conv_spatial_deform_weight;                    /* [C,1,7,7]                  */
    const float *conv1_weight, *conv1_bias;                     /* [C,C,1,1], [C]        :95  */
    const float *proj_2_weight, *proj_2_bias;                   /* [C,C,1,1], [C]        :130 */
} dlkaBlock2dParams;

DLKA_API size_t dlka_deformable_lka2d_workspace_bytes(int B, int C, int H, int W);
DLKA_API int dlka_deformable_lka2d_forward(const dlkaBlock2dParams *params, const float *x, float *y,
                                  int B, int C, int H, int W, int math,
                                  void *workspace, size_t workspace_bytes, void *stream);
DLKA_API size_t dlka_deformable_lka_attention2d_workspace_bytes(int B, int C, int H, int W);
/* prepacked-weights variant: see dlka_lka_attention3d_deform_forward_packed */
DLKA_API size_t dlka_deformable_lka_attention2d_packed_bytes(int C);
DLKA_API int dlka_deformable_lka_attention2d_forward_packed(const dlkaBlock2dParams *params, const float *x, float *y,
                                                   int B, int C, int H, int W, int math,
                                                   void *packed, size_t packed_bytes, int packed_valid,
                                                   void *workspace, size_t workspace_bytes, void *stream);
DLKA_API int dlka_deformable_lka_attention2d_forward(const dlkaBlock2dParams *params, const float *x, float *y,
                                            int B, int C, int H, int W, int math,
                                            void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Enclosing transformer blocks.  Tokens [B, N, C] in, tokens out (channels-last data).
 *
 * 2D: replaces  deformableLKABlock.forward  (2D/networks/MaxViT_deform_LKA.py:165-189), eval mode
 *     (drop = drop_path = 0, linear = False):
 *       x1 = x  + layer_scale_1 * Attn(LayerNorm1(x))
 *       y  = x1 + layer_scale_2 * fc2(GELU(dwconv3x3(fc1(LayerNorm2(x1)))))
 * ------------------------------------------------------------------------------------------ */
typedef struct dlkaLkaBlock2dParams {
    const float *norm1_weight, *norm1_bias;   /* [C]                 MaxViT_deform_LKA.py:151 */
    dlkaBlock2dParams attn;                   /* deformable_LKA_Attention(dim)           :152 */
    const float *layer_scale_1;               /* [C]                                     :160 */
    const float *norm2_weight, *norm2_bias;   /* [C]                                     :156 */
    const float *fc1_weight, *fc1_bias;       /* [hidden,C,1,1], [hidden]   Mlp.fc1      :34  */
    const float *dw_weight, *dw_bias;         /* [hidden,1,3,3], [hidden]   DWConvLKA    :21  */
    const float *fc2_weight, *fc2_bias;       /* [C,hidden,1,1], [C]        Mlp.fc2      :37  */
    const float *layer_scale_2;               /* [C]                                     :161 */
    float eps1, eps2;                         /* LayerNorm eps (1e-5)                          */
    int hidden;                               /* int(dim * mlp_ratio)                          */
} dlkaLkaBlock2dParams;

DLKA_API size_t dlka_deformable_lka_block2d_workspace_bytes(int B, int C, int H, int W, int hidden);
DLKA_API int dlka_deformable_lka_block2d_forward(const dlkaLkaBlock2dParams *params, const float *x, float *y,
                                        int B, int C, int H, int W, int math,
                                        void *workspace, size_t workspace_bytes, void *stream);

/* 3D: replaces the attention half of  TransformerBlock_3D_single_deform_LKA.forward
 *     (3D/d_lka_former/network_architecture/synapse/transformerblock.py:620-624):
 *       x' = x + pos_embed (pos_embed [N, C] or NULL);  y = x' + gamma * LKA_Attention3d_deform(LayerNorm(x'))
 *     The UnetResBlock / conv8 tail (:626-628) is row N3 and is not part of this entry.        */
DLKA_API size_t dlka_lka_transformer3d_prenorm_workspace_bytes(int B, int C, int D1, int D2, int D3);
DLKA_API int dlka_lka_transformer3d_prenorm_forward(const dlkaBlock3dParams *attn, const float *norm_weight,
                                           const float *norm_bias, float eps, const float *gamma, const float *pos_embed,
                                           const float *x, float *y, int B, int C, int D1, int D2, int D3, int math,
                                           void *workspace, size_t workspace_bytes, void *stream);

/* Whole 3D transformer block on tokens (rows N1 + N3), inference mode: replaces
 * TransformerBlock_3D_single_deform_LKA.forward (transformerblock.py:617-630) between its two layout reshapes:
 *   a = x' + gamma * Attn(LayerNorm(x'));  r = LeakyReLU(BN1(conv1 a));  r = LeakyReLU(BN2(conv2 r) + a);  y = a + conv8(r)
 * conv1/conv2: UnetResBlock 3x3x3 convs without bias (dynunet_block.py:44-53, 65-80); BatchNorm3d (running statistics)
 * is passed folded: scale = weight / sqrt(running_var + eps), shift = bias - running_mean * scale.                  */
typedef struct dlkaTransformer3dParams {
    dlkaBlock3dParams attn;                    /* epa_block                                transformerblock.py:609 */
    const float *norm_weight, *norm_bias;      /* LayerNorm(hidden_size)                                      :607 */
    const float *gamma;                        /* [C]                                                         :608 */
    const float *pos_embed;                    /* [N, C] or NULL                                              :613-615 */
    const float *conv1_weight;                 /* conv51.conv1.conv.weight [C,C,3,3,3]                        :610 */
    const float *bn1_scale, *bn1_shift;        /* folded conv51.norm1                                              */
    const float *conv2_weight;                 /* conv51.conv2.conv.weight [C,C,3,3,3]                             */
    const float *bn2_scale, *bn2_shift;        /* folded conv51.norm2                                              */
    const float *conv8_weight, *conv8_bias;    /* conv8[1]: Conv3d(C, C, 1)                                   :611 */
    float eps;                                 /* LayerNorm eps                                                    */
    float lrelu_slope;                         /* 0.01 (dynunet_block.py:39)                                       */
} dlkaTransformer3dParams;

DLKA_API size_t dlka_lka_transformer3d_block_workspace_bytes(int B, int C, int D1, int D2, int D3);
DLKA_API int dlka_lka_transformer3d_block_forward(const dlkaTransformer3dParams *params, const float *x, float *y,
                                         int B, int C, int D1, int D2, int D3, int math,
                                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * 2D decoder glue (rest of row N3): 2D/networks/MaxViT_deform_LKA.py
 *   dlka_linear_tokens_forward     MyDecoderLayer.x1_linear + the skip add        (:604-607)   y = x W^T + b (+ add)
 *   dlka_patch_expand2d_forward    PatchExpand.forward (scale 2, :488-513) / FinalPatchExpand_X4.forward (scale 4, :516-545):
 *                                  Linear without bias -> "b h w (p1 p2 c) -> b (h p1) (w p2) c" -> LayerNorm, no rearranged copy
 * ------------------------------------------------------------------------------------------ */
DLKA_API size_t dlka_linear_tokens_workspace_bytes(int K, int N);
DLKA_API int dlka_linear_tokens_forward(const float *x, const float *weight, const float *bias, const float *add, float *y,
                               long long M, int K, int N, int math, void *workspace, size_t workspace_bytes, void *stream);
DLKA_API size_t dlka_patch_expand2d_workspace_bytes(int B, int H, int W, int dim, int scale);
DLKA_API int dlka_patch_expand2d_forward(const float *x, const float *expand_weight, const float *norm_weight, const float *norm_bias,
                                float eps, float *y, int B, int H, int W, int dim, int scale, int math,
                                void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DLKA_H_ */
