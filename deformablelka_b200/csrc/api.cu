// api.cu -- the C ABI of libdlka_b200.so (see include/dlka.h for the contract and the reference
// interfaces each entry point replaces).
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "kernels.cuh"

namespace dlka {

thread_local char g_last_cuda_error[256] = {0};
static std::atomic<uint64_t> g_launches{0};

void note_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static thread_local bool g_pack_skip = false;
bool pack_skipped() { return g_pack_skip; }
PackSkipScope::PackSkipScope(bool skip) : prev(g_pack_skip) { g_pack_skip = skip; }
PackSkipScope::~PackSkipScope() { g_pack_skip = prev; }

// ---- optional per-kernel event timing ---------------------------------------------------
struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
};
static std::atomic<int> g_profiling{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;

KernelScope::KernelScope(const char *n, cudaStream_t s) : name(n), st(s), a(nullptr), b(nullptr), active(false)
{
    if (!g_profiling.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    active = cudaEventRecord(a, st) == cudaSuccess;
}

KernelScope::~KernelScope()
{
    if (!active) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back({name, a, b});
}

int record_cuda_error(cudaError_t e, const char *what)
{
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? DLKA_ERR_NO_DEVICE : DLKA_ERR_CUDA;
}

namespace {

// The CURRENT device must be an sm_100 part.  Only a positive answer is cached, per device ordinal (a transient
// cudaGetDevice failure must not poison the thread, and one process may drive several devices).
int check_device()
{
    static std::atomic<int> ok[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return record_cuda_error(e, "cudaGetDevice");
    const bool cacheable = dev >= 0 && dev < 64;
    if (cacheable && ok[dev].load(std::memory_order_relaxed)) return DLKA_OK;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return record_cuda_error(e, "cudaDeviceGetAttribute");
    if (major != 10) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "device compute capability %d.x is not sm_100", major);
        return DLKA_ERR_NO_DEVICE;
    }
    if (cacheable) ok[dev].store(1, std::memory_order_relaxed);
    return DLKA_OK;
}

IgemmArgs dense_args(const float *X, int ldX, i64 M, int Ci, int Co, const float *Wp, int Npad, const float *bias, int epi,
                     const float *E, int ldE, float *Y, int ldY)
{
    IgemmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = IGEMM_DENSE; a.epi = epi;
    a.geo = make_geo(1, Ci, 1, 1, 1, Co, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 3);
    a.M = M; a.Ktot = Ci; a.Npad = Npad; a.X = X; a.ldX = ldX; a.Wp = Wp; a.bias = bias; a.E = E; a.ldE = ldE;
    a.Y = Y; a.ldY = ldY;
    return a;
}

IgemmArgs conv_args(int mode, const ConvGeo &g, const float *X, const float *Off, const float *Mask, const float *Wp, int Npad,
                    const float *bias, int epi, const float *E, int ldE, float *Y, int ldY)
{
    IgemmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.epi = epi; a.geo = g;
    a.M = (i64)g.B * g.Do * g.Ho * g.Wo; a.Ktot = g.K * (g.C / g.groups); a.Npad = Npad;
    a.X = X; a.ldX = g.C; a.Off = Off; a.Mask = Mask; a.Wp = Wp; a.bias = bias; a.E = E; a.ldE = ldE; a.Y = Y; a.ldY = ldY;
    return a;
}

// One channel contraction: packs the weight (reference layout [Co][C/g][taps]) into `wscratch` and runs the
// tensor-core kernel (DLKA_MATH_BF16X3, when the shape qualifies) or the exact fp32 SIMT kernel.
size_t contraction_scratch_floats(int Co, int C, int taps, int groups)
{
    const size_t simt = (size_t)groups * taps * (C / groups) * igemm_simt_npad(Co / groups);
    const size_t tc = (groups == 1 && tc_kc(C) != 0) ? tc_packed_weight_bytes(Co, C, taps) / sizeof(float) : 0;
    return (simt > tc ? simt : tc) + 64;
}

int contraction(IgemmArgs a, const float *w, int math, float *wscratch, cudaStream_t st)
{
    const ConvGeo &g = a.geo;
    if (math == DLKA_MATH_BF16X3 && a.mode == IGEMM_CONV && conv_tiled_supported(a)) return conv_tiled(a, w, wscratch, st);
    if (math == DLKA_MATH_BF16X3 && a.mode == IGEMM_DEFORM && deform3d_tc_supported(a)) return deform3d_tc(a, w, wscratch, nullptr, st);
    // large-M 1x1 projections: the persistent streaming kernel (HBM-bound; at least two 128-row tiles per SM)
    if (math == DLKA_MATH_BF16X3 && a.mode == IGEMM_DENSE && a.M >= 2 * 148 * 128 && dense_stream_supported(a)) {
        DLKA_TRY(deform3d_ps_pack(w, wscratch, g.Co, g.C, 1, st));
        return dense_stream(a, wscratch, st);
    }
    if (math == DLKA_MATH_BF16X3 && tc_supported(a)) {
        DLKA_TRY(tc_pack_weight(w, wscratch, g.Co, g.C, g.K, st));
        return igemm_tc(a, wscratch, st);
    }
    a.Npad = igemm_simt_npad(g.Co / g.groups);
    a.Wp = wscratch;
    DLKA_TRY(pack_weight(w, wscratch, g.Co, g.C / g.groups, g.K, g.groups, a.Npad, st));
    return igemm_simt(a, st);
}

bool bad_geo(const ConvGeo &g)
{
    return g.B <= 0 || g.C <= 0 || g.Co <= 0 || g.D <= 0 || g.H <= 0 || g.W <= 0 || g.kd <= 0 || g.kh <= 0 || g.kw <= 0 ||
           g.sd <= 0 || g.sh <= 0 || g.sw <= 0 || g.pd < 0 || g.ph < 0 || g.pw < 0 || g.dd <= 0 || g.dh <= 0 || g.dw <= 0 ||
           g.groups <= 0 || g.dg <= 0 || g.C % g.groups || g.Co % g.groups || g.C % g.dg || g.Do <= 0 || g.Ho <= 0 || g.Wo <= 0;
}

// DLKA_DEFORM_PS=0 selects the round-1 kernel (deform_tc.cu) for A/B measurements; both are product code paths with the same tests
bool deform_ps_enabled()
{
    static const bool on = [] {
        const char *e = getenv("DLKA_DEFORM_PS");
        return !(e && e[0] == '0');
    }();
    return on;
}

bool is_depthwise(const ConvGeo &g) { return g.groups == g.C && g.Co == g.C && g.C > 1; }

// ---- workspace plans -------------------------------------------------------------------
struct DeformOpPlan {
    float *x_cl, *off_cl, *mask_cl, *y_cl, *wp;
    int Npad;
};

bool plan_deform_op(Arena &ar, const ConvGeo &g, bool has_mask, DeformOpPlan &p)
{
    const i64 Vi = (i64)g.D * g.H * g.W, M = (i64)g.B * g.Do * g.Ho * g.Wo;
    p.x_cl = ar.take<float>((size_t)g.B * Vi * g.C);
    p.off_cl = ar.take<float>((size_t)M * g.dg * g.ndim * g.K);
    p.mask_cl = has_mask ? ar.take<float>((size_t)M * g.dg * g.K) : nullptr;
    p.y_cl = ar.take<float>((size_t)M * g.Co);
    p.Npad = 0;
    if (is_depthwise(g))
        p.wp = ar.take<float>((size_t)g.K * g.C);
    else
        p.wp = ar.take<float>(contraction_scratch_floats(g.Co, g.C, g.K, g.groups));
    return ar.ok();
}

int run_deform_op(const ConvGeo &g, const float *input, const float *weight, const float *bias, const float *offset,
                  const float *mask, float *output, int math, void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    Arena ar(workspace, workspace_bytes);
    DeformOpPlan p;
    if (!plan_deform_op(ar, g, mask != nullptr, p)) return DLKA_ERR_WORKSPACE;
    const i64 Vi = (i64)g.D * g.H * g.W, Vo = (i64)g.Do * g.Ho * g.Wo;
    {
        // 3D operator on the persistent kernel (deform_ps.cu): the NCDHW input is transposed straight into the chunk-major
        // gather layout (same bytes as the channels-last copy it replaces)
        IgemmArgs a = conv_args(IGEMM_DEFORM, g, p.x_cl, p.off_cl, nullptr, nullptr, 0, bias, EPI_NONE, nullptr, 0, p.y_cl, g.Co);
        if (math == DLKA_MATH_BF16X3 && !mask && deform_ps_enabled() && deform3d_ps_supported(a, 0)) {
            DLKA_TRY(transpose_cs_to_chunk(input, p.x_cl, g.B, g.C, Vi, st));
            DLKA_TRY(deform3d_ps_pack(weight, p.wp, g.Co, g.C, g.K, st));
            a.Off = offset;   // read in place: a warp's 32 brick rows are 4 runs of 8 consecutive voxels of the NCDHW tensor
            DLKA_TRY(deform3d_ps(a, (i64)g.B * Vi * 32, p.wp, nullptr, 2, st));
            DLKA_TRY(transpose_sc_to_cs(p.y_cl, output, g.B, g.Co, Vo, st));
            return DLKA_OK;
        }
    }
    DLKA_TRY(transpose_cs_to_sc(input, p.x_cl, g.B, g.C, Vi, st));
    DLKA_TRY(transpose_cs_to_sc(offset, p.off_cl, g.B, g.dg * g.ndim * g.K, Vo, st));
    if (mask) DLKA_TRY(transpose_cs_to_sc(mask, p.mask_cl, g.B, g.dg * g.K, Vo, st));
    if (is_depthwise(g)) {
        DLKA_TRY(deform_dwconv_cl(p.x_cl, p.off_cl, p.mask_cl, weight, bias, p.y_cl, g, p.wp, st));
    } else {
        if ((g.C / g.groups) % 4 != 0 || (g.C / g.dg) % 4 != 0) return DLKA_ERR_UNSUPPORTED;
        IgemmArgs a = conv_args(IGEMM_DEFORM, g, p.x_cl, p.off_cl, p.mask_cl, nullptr, 0, bias, EPI_NONE, nullptr, 0, p.y_cl, g.Co);
        DLKA_TRY(contraction(a, weight, math, p.wp, st));
    }
    DLKA_TRY(transpose_sc_to_cs(p.y_cl, output, g.B, g.Co, Vo, st));
    return DLKA_OK;
}

// ---- 3D block ---------------------------------------------------------------------------
constexpr int OFF3D_LD = 84;  // row stride of the [M][81] offset buffer (16-byte aligned rows)
constexpr size_t SPLIT_SCRATCH_CAP = (size_t)4 << 20;   // floats: upper bound of a K-split scratch buffer
constexpr size_t SPLIT3D_ROWS = 2048;   // volumes up to this many voxels run the C > 96 deformable conv K-split over channel chunks
struct Block3dPlan {
    float *t1, *t2, *t3, *off, *split, *off_split;
    i64 off_split_floats;
    float *wp_proj1, *wp_off, *wp_dcn, *wp_conv1, *wp_proj2, *wp_dw5, *wp_dw7;
    int np_c, np_off;
};

// wpk: arena of the packed weights (the caller's persistent buffer of *_forward_packed), or null = scratch in the workspace
bool plan_block3d(Arena &ar, int B, int C, int D1, int D2, int D3, Block3dPlan &p, Arena *wpk = nullptr)
{
    Arena &wa = wpk ? *wpk : ar;
    const size_t M = (size_t)B * D1 * D2 * D3;
    p.np_c = p.np_off = 0;
    p.t1 = ar.take<float>(M * C);
    p.t2 = ar.take<float>(M * C);
    p.t3 = ar.take<float>(M * C);
    {   // offsets: [M][84] rows (round-1 kernels) or brick-major bricks x 81 x 128 (deform_ps.cu); ragged volumes pad the bricks
        const size_t brick = deform3d_ps_offset_floats(B, D1, D2, D3, 81);
        const size_t offsz = M * OFF3D_LD > brick ? M * OFF3D_LD : brick;
        p.off = ar.take<float>(offsz);
        // K-split scratch of the offset conv on small volumes (conv_tiled_ex decides whether it splits)
        p.off_split_floats = M <= 16384 ? (i64)(16 * offsz < SPLIT_SCRATCH_CAP ? 16 * offsz : SPLIT_SCRATCH_CAP) : 0;
        p.off_split = p.off_split_floats ? ar.take<float>((size_t)p.off_split_floats) : nullptr;
    }
    // K-split partial sums of the deformable conv on small volumes (<= SPLIT3D_ROWS rows, one slice per 32-channel chunk)
    p.split = (M <= SPLIT3D_ROWS && C % 32 == 0 && C / 32 >= 2) ? ar.take<float>(M * C * (size_t)(C / 32)) : nullptr;
    p.wp_proj1 = wa.take<float>(contraction_scratch_floats(C, C, 1, 1));
    p.wp_conv1 = wa.take<float>(contraction_scratch_floats(C, C, 1, 1));
    p.wp_proj2 = wa.take<float>(contraction_scratch_floats(C, C, 1, 1));
    p.wp_off = wa.take<float>(contraction_scratch_floats(81, C, 27, 1));
    p.wp_dcn = wa.take<float>(contraction_scratch_floats(C, C, 27, 1));
    p.wp_dw5 = wa.take<float>((size_t)125 * C);
    p.wp_dw7 = wa.take<float>((size_t)343 * C);
    return ar.ok() && wa.ok();
}

// packed depthwise weights live in fixed workspace slots sized for 5^3 and 7^3 taps
bool bad_dw_geom(const dlkaDwGeom3d &G)
{
    auto bad = [](const int *k, const int *dil, int max_taps) {
        for (int i = 0; i < 3; ++i)
            if (k[i] < 1 || !(k[i] & 1) || dil[i] < 1) return true;
        return k[1] != k[2] || dil[1] != dil[2] || k[0] * k[1] * k[2] > max_taps;
    };
    return bad(G.conv0_k, G.conv0_dil, 125) || bad(G.conv_spatial_k, G.conv_spatial_dil, 343);
}

// u (channels-last, = GELU(proj_1 x) or x itself) -> gate = u * conv1(deform(dw7(dw5(u)))) into p.t3
// returns DLKA_OK with the gate in p.t3, or 1 when (fuse_proj2) proj_2 + shortcut were fused and y_final is complete
int run_lka3d_core(const dlkaBlock3dParams &P, const float *u, Block3dPlan &p, int B, int C, int D1, int D2, int D3,
                   int math, cudaStream_t st, bool fuse_proj2 = false, const float *resid = nullptr, float *y_final = nullptr)
{
    const i64 M = (i64)B * D1 * D2 * D3;
    static const dlkaDwGeom3d synapse = {{5, 5, 5}, {1, 1, 1}, {7, 7, 7}, {3, 3, 3}};   // transformerblock.py:637-638
    const dlkaDwGeom3d &G = P.dw_geom ? *P.dw_geom : synapse;
    if (bad_dw_geom(G)) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(dwconv_cl(u, P.conv0_weight, P.conv0_bias, p.t2, B, C, D1, D2, D3, G.conv0_k[0], G.conv0_k[1], G.conv0_k[2],
                       G.conv0_dil[0], G.conv0_dil[1], p.wp_dw5, st));
    // deformable 3x3x3 conv C -> C, groups 1, dg 1 (transformerblock.py:639); conv_offset: Conv3d(C -> 81, k3, stride 1, pad 1)
    // (synapse/deform_conv.py:80-85)
    const ConvGeo go = make_geo(B, C, D1, D2, D3, 81, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3);
    const ConvGeo gd = make_geo(B, C, D1, D2, D3, C, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3);
    IgemmArgs ad = conv_args(IGEMM_DEFORM, gd, p.t3, p.off, nullptr, nullptr, 0, P.deform_bias, EPI_NONE, nullptr, 0, p.t2, C);
    ad.ldOff = OFF3D_LD;
    IgemmArgs ao = conv_args(IGEMM_CONV, go, p.t3, nullptr, nullptr, nullptr, 0, P.conv_offset_bias, EPI_NONE, nullptr, 0, p.off,
                             OFF3D_LD);
    ao.split_scratch = p.off_split; ao.split_scratch_floats = p.off_split_floats;
    // persistent deformable kernel (deform_ps.cu): the large-kernel stencil writes its output CHUNK-MAJOR
    // ([C/32][B][D][H][W][32]), which the offset conv and the deformable gather then read -- no extra pass over the tensor
    const i64 xch = (i64)B * D1 * D2 * D3 * 32;
    IgemmArgs ao_cm = ao;
    ao_cm.xch = xch;
    ao_cm.ybrick = 1;
    const bool use_ps = math == DLKA_MATH_BF16X3 && deform_ps_enabled() && deform3d_ps_supported(ad, fuse_proj2 ? 2 : 1) &&
                        conv_tiled_supported(ao_cm);   // (both stencil kernels can write chunk-major when C % 32 == 0)
    DLKA_TRY(dwconv_cl(p.t2, P.conv_spatial_weight, P.conv_spatial_bias, p.t3, B, C, D1, D2, D3, G.conv_spatial_k[0],
                       G.conv_spatial_k[1], G.conv_spatial_k[2], G.conv_spatial_dil[0], G.conv_spatial_dil[1], p.wp_dw7, st, use_ps));
    if (use_ps) {
        DLKA_TRY(conv_tiled(ao_cm, P.conv_offset_weight, p.wp_off, st));
        DeformChain ch;
        memset(&ch, 0, sizeof(ch));
        ch.stages = fuse_proj2 ? 2 : 1;
        DLKA_TRY(deform3d_ps_pack(P.deform_weight, p.wp_dcn, C, C, 27, st));
        DLKA_TRY(deform3d_ps_pack(P.conv1_weight, p.wp_conv1, C, C, 1, st));
        ch.W1p = p.wp_conv1; ch.b1 = P.conv1_bias; ch.U = u; ch.ldU = C;
        if (fuse_proj2) {
            DLKA_TRY(deform3d_ps_pack(P.proj_2_weight, p.wp_proj2, C, C, 1, st));
            ch.W2p = p.wp_proj2; ch.b2 = P.proj_2_bias; ch.R = resid; ch.ldR = C;
            ad.Y = y_final;
        } else {
            ad.Y = p.t2;  // must not alias the gather source t3
        }
        DLKA_TRY(deform3d_ps(ad, xch, p.wp_dcn, &ch, 1, st));
        if (!fuse_proj2) {
            float *tmp = p.t2; p.t2 = p.t3; p.t3 = tmp;  // callers find the gate in p.t3
        }
        return fuse_proj2 ? 1 : DLKA_OK;
    }
    DLKA_TRY(contraction(ao, P.conv_offset_weight, math, p.wp_off, st));
    if (math == DLKA_MATH_BF16X3 && deform3d_chain_supported(ad)) {
        // deformable conv + conv1 + gate (+ proj_2 + shortcut) in ONE kernel: the 1x1 GEMMs run on the accumulator tile
        DeformChain ch;
        memset(&ch, 0, sizeof(ch));
        ch.stages = fuse_proj2 ? 2 : 1;
        DLKA_TRY(tc_pack_weight(P.conv1_weight, p.wp_conv1, C, C, 1, st));
        ch.W1p = p.wp_conv1; ch.b1 = P.conv1_bias; ch.U = u; ch.ldU = C;
        if (fuse_proj2) {
            DLKA_TRY(tc_pack_weight(P.proj_2_weight, p.wp_proj2, C, C, 1, st));
            ch.W2p = p.wp_proj2; ch.b2 = P.proj_2_bias; ch.R = resid; ch.ldR = C;
            ad.Y = y_final;
        } else {
            ad.Y = p.t2;  // must not alias the gather source t3
        }
        DLKA_TRY(deform3d_tc(ad, P.deform_weight, p.wp_dcn, &ch, st));
        if (!fuse_proj2) {
            float *tmp = p.t2; p.t2 = p.t3; p.t3 = tmp;  // callers find the gate in p.t3
        }
        return fuse_proj2 ? 1 : DLKA_OK;  // 1: the whole tail (proj_2 + shortcut) is already in y_final
    }
    if (math == DLKA_MATH_BF16X3 && p.split && deform3d_tc_supported(ad)) {
        // small volume, C > 96 (the deep stages of the 3D net: 128 x 8^3, 256 x 4^3): a handful of 128-row tiles with a K loop of
        // 27 * C/32 steps each -- one slice per channel chunk fills 4 .. 16x more SMs; the slices are summed by reduce_partials
        IgemmArgs as = ad;
        as.Y = p.split; as.ksplit_steps = 1; as.ysplit_stride = M * C;
        DLKA_TRY(deform3d_tc(as, P.deform_weight, p.wp_dcn, nullptr, st));
        DLKA_TRY(reduce_partials(p.split, p.t2, M * C, C / 32, st));
    } else {
        DLKA_TRY(contraction(ad, P.deform_weight, math, p.wp_dcn, st));
    }
    // conv1 (1x1x1) then gate with u
    IgemmArgs a1 = dense_args(p.t2, C, M, C, C, nullptr, 0, P.conv1_bias, EPI_MUL, u, C, p.t3, C);
    DLKA_TRY(contraction(a1, P.conv1_weight, math, p.wp_conv1, st));
    return DLKA_OK;
}

bool null_params3d(const dlkaBlock3dParams *P, bool attention)
{
    if (!P) return true;
    if (!P->conv0_weight || !P->conv0_bias || !P->conv_spatial_weight || !P->conv_spatial_bias || !P->conv_offset_weight ||
        !P->conv_offset_bias || !P->deform_weight || !P->deform_bias || !P->conv1_weight || !P->conv1_bias)
        return true;
    if (attention && (!P->proj_1_weight || !P->proj_1_bias || !P->proj_2_weight || !P->proj_2_bias)) return true;
    return false;
}

// ---- 2D block ---------------------------------------------------------------------------
struct Block2dPlan {
    float *x_cl, *t1, *t2, *t3, *off, *off_split;
    i64 off_split_floats;
    float *wp_proj1, *wp_off0, *wp_off1, *wp_conv1, *wp_proj2, *wp_dw0, *wp_dw1;
    int np_c, np_off0, np_off1;
};

bool plan_block2d(Arena &ar, int B, int C, int H, int W, Block2dPlan &p, Arena *wpk = nullptr)
{
    Arena &wa = wpk ? *wpk : ar;
    const size_t M = (size_t)B * H * W;
    p.np_c = p.np_off0 = p.np_off1 = 0;
    p.x_cl = ar.take<float>(M * C);
    p.t1 = ar.take<float>(M * C);
    p.t2 = ar.take<float>(M * C);
    p.t3 = ar.take<float>(M * C);
    p.off = ar.take<float>(M * 98);
    p.off_split_floats = M <= 32768 ? (i64)(8 * M * 98 < SPLIT_SCRATCH_CAP ? 8 * M * 98 : SPLIT_SCRATCH_CAP) : 0;
    p.off_split = p.off_split_floats ? ar.take<float>((size_t)p.off_split_floats) : nullptr;
    p.wp_proj1 = wa.take<float>(contraction_scratch_floats(C, C, 1, 1));
    p.wp_conv1 = wa.take<float>(contraction_scratch_floats(C, C, 1, 1));
    p.wp_proj2 = wa.take<float>(contraction_scratch_floats(C, C, 1, 1));
    p.wp_off0 = wa.take<float>(contraction_scratch_floats(50, C, 25, 1));
    p.wp_off1 = wa.take<float>(contraction_scratch_floats(98, C, 49, 1));
    p.wp_dw0 = wa.take<float>((size_t)25 * C);
    p.wp_dw1 = wa.take<float>((size_t)49 * C);
    return ar.ok() && wa.ok();
}

// u channels-last -> u * conv1(conv_spatial(conv0(u))) into p.t2
int run_lka2d_core(const dlkaBlock2dParams &P, const float *u, Block2dPlan &p, int B, int C, int H, int W, int math,
                   cudaStream_t st)
{
    const i64 M = (i64)B * H * W;
    // conv0: offset_net Conv2d(C->50, k5, pad 2) + depthwise deformable k5 (deformable_LKA.py:93)
    ConvGeo g0 = make_geo(B, C, 1, H, W, 50, 1, 5, 5, 1, 1, 1, 0, 2, 2, 1, 1, 1, 1, 1, 2);
    IgemmArgs a0 = conv_args(IGEMM_CONV, g0, u, nullptr, nullptr, nullptr, 0, P.conv0_offset_bias, EPI_NONE, nullptr, 0, p.off, 50);
    a0.split_scratch = p.off_split; a0.split_scratch_floats = p.off_split_floats;
    DLKA_TRY(contraction(a0, P.conv0_offset_weight, math, p.wp_off0, st));
    ConvGeo d0 = make_geo(B, C, 1, H, W, C, 1, 5, 5, 1, 1, 1, 0, 2, 2, 1, 1, 1, C, 1, 2);
    DLKA_TRY(deform_dwconv_cl(u, p.off, nullptr, P.conv0_deform_weight, nullptr, p.t2, d0, p.wp_dw0, st));
    // conv_spatial: offset_net Conv2d(C->98, k7, dil 3, pad 9) + depthwise deformable k7 dil 3 (:94)
    ConvGeo g1 = make_geo(B, C, 1, H, W, 98, 1, 7, 7, 1, 1, 1, 0, 9, 9, 1, 3, 3, 1, 1, 2);
    IgemmArgs a1 = conv_args(IGEMM_CONV, g1, p.t2, nullptr, nullptr, nullptr, 0, P.conv_spatial_offset_bias, EPI_NONE, nullptr, 0,
                             p.off, 98);
    a1.split_scratch = p.off_split; a1.split_scratch_floats = p.off_split_floats;
    DLKA_TRY(contraction(a1, P.conv_spatial_offset_weight, math, p.wp_off1, st));
    ConvGeo d1 = make_geo(B, C, 1, H, W, C, 1, 7, 7, 1, 1, 1, 0, 9, 9, 1, 3, 3, C, 1, 2);
    DLKA_TRY(deform_dwconv_cl(p.t2, p.off, nullptr, P.conv_spatial_deform_weight, nullptr, p.t3, d1, p.wp_dw1, st));
    // conv1 1x1 and the gate
    IgemmArgs ac = dense_args(p.t3, C, M, C, C, nullptr, 0, P.conv1_bias, EPI_MUL, u, C, p.t2, C);
    DLKA_TRY(contraction(ac, P.conv1_weight, math, p.wp_conv1, st));
    return DLKA_OK;
}

bool null_params2d(const dlkaBlock2dParams *P, bool attention)
{
    if (!P) return true;
    if (!P->conv0_offset_weight || !P->conv0_offset_bias || !P->conv0_deform_weight || !P->conv_spatial_offset_weight ||
        !P->conv_spatial_offset_bias || !P->conv_spatial_deform_weight || !P->conv1_weight || !P->conv1_bias)
        return true;
    if (attention && (!P->proj_1_weight || !P->proj_1_bias || !P->proj_2_weight || !P->proj_2_bias)) return true;
    return false;
}

// attention block on channels-last data with an existing plan: y_cl = proj_2(gate(GELU(proj_1 x))) + x   (y_cl != x_cl)
int attention2d_cl_planned(const dlkaBlock2dParams &P, const float *x_cl, float *y_cl, Block2dPlan &p, int B, int C, int H, int W,
                           int math, cudaStream_t st)
{
    const i64 M = (i64)B * H * W;
    IgemmArgs a1 = dense_args(x_cl, C, M, C, C, nullptr, 0, P.proj_1_bias, EPI_GELU, nullptr, 0, p.t1, C);
    DLKA_TRY(contraction(a1, P.proj_1_weight, math, p.wp_proj1, st));
    DLKA_TRY(run_lka2d_core(P, p.t1, p, B, C, H, W, math, st));  // gate -> t2
    IgemmArgs a2 = dense_args(p.t2, C, M, C, C, nullptr, 0, P.proj_2_bias, EPI_ADD, x_cl, C, y_cl, C);
    DLKA_TRY(contraction(a2, P.proj_2_weight, math, p.wp_proj2, st));
    return DLKA_OK;
}

}  // namespace

// ---- internal entry points used by blocks_api.cu (channels-last tokens in, tokens out) ----
int attention2d_cl(const dlkaBlock2dParams *params, const float *x_cl, float *y_cl, int B, int C, int H, int W, int math,
                   void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    if (null_params2d(params, true) || !x_cl || !y_cl || x_cl == y_cl) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(check_device());
    Arena ar(workspace, workspace_bytes);
    Block2dPlan p;
    if (!plan_block2d(ar, B, C, H, W, p)) return DLKA_ERR_WORKSPACE;
    return attention2d_cl_planned(*params, x_cl, y_cl, p, B, C, H, W, math, st);
}

int dense_cl(const float *x, int ldX, i64 M, int Ci, int Co, const float *w, const float *bias, int epi, const float *E, int ldE,
             float *y, int ldY, int math, float *wscratch, cudaStream_t st)
{
    IgemmArgs a = dense_args(x, ldX, M, Ci, Co, nullptr, 0, bias, epi, E, ldE, y, ldY);
    return contraction(a, w, math, wscratch, st);
}

int dense_splitk_cl(const float *x, int ldX, i64 M, int Ci, int Co, const float *w, float *partial, int ksplit_steps, int *nsplit,
                    float *wscratch, cudaStream_t st)
{
    IgemmArgs a = dense_args(x, ldX, M, Ci, Co, nullptr, 0, nullptr, EPI_NONE, nullptr, 0, partial, Co);
    if (!tc_supported(a) || ksplit_steps <= 0) return DLKA_ERR_UNSUPPORTED;
    a.ksplit_steps = ksplit_steps;
    a.ysplit_stride = M * (i64)Co;
    *nsplit = (int)cdiv(Ci / tc_kc(Ci), ksplit_steps);
    DLKA_TRY(tc_pack_weight(w, wscratch, Co, Ci, 1, st));
    return igemm_tc(a, wscratch, st);
}

size_t dense_scratch_floats(int Co, int Ci) { return contraction_scratch_floats(Co, Ci, 1, 1); }

size_t conv3_scratch_floats(int C) { return contraction_scratch_floats(C, C, 27, 1); }

int conv3_bn_act_cl(const float *x, const float *w, const float *scale, const float *shift, int act, float slope, const float *E,
                    float *y, int B, int C, int D1, int D2, int D3, int math, float *wscratch, cudaStream_t st)
{
    const ConvGeo g = make_geo(B, C, D1, D2, D3, C, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3);
    IgemmArgs a = conv_args(IGEMM_CONV, g, x, nullptr, nullptr, nullptr, 0, shift, EPI_NONE, nullptr, 0, y, C);
    if (math == DLKA_MATH_BF16X3 && conv_tiled_supported(a))
        return conv_tiled_ex(a, w, scale, act, slope, E, C, wscratch, st);   // scale folded into the weights, shift = bias
    a.bias = nullptr;
    DLKA_TRY(contraction(a, w, math, wscratch, st));
    return affine_act_cl(y, scale, shift, E, a.M, C, act, slope, st);
}

int device_ok() { return check_device(); }

}  // namespace dlka

using namespace dlka;

extern "C" {

int dlka_version(void) { return DLKA_VERSION; }

const char *dlka_status_string(int status)
{
    switch (status) {
    case DLKA_OK: return "ok";
    case DLKA_ERR_INVALID_ARGUMENT: return "invalid argument (shape / pointer / size mismatch)";
    case DLKA_ERR_UNSUPPORTED: return "configuration not supported by libdlka_b200";
    case DLKA_ERR_WORKSPACE: return "workspace missing or too small";
    case DLKA_ERR_NO_DEVICE: return "no sm_100 CUDA device (libdlka_b200 has no CPU path)";
    case DLKA_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
    }
}

const char *dlka_last_cuda_error(void) { return g_last_cuda_error; }
uint64_t dlka_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dlka_profile_enable(int on)
{
    g_profiling.store(on ? 1 : 0);
    return DLKA_OK;
}

int dlka_profile_summary(char *buf, size_t buf_bytes)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto &r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto &e = agg[r.name];
            e.first += 1;
            e.second += ms;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
    std::string out;
    char line[160];
    for (auto &kv : agg) {
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (!buf || buf_bytes == 0) return DLKA_ERR_INVALID_ARGUMENT;
    snprintf(buf, buf_bytes, "%s", out.c_str());
    return DLKA_OK;
}

// ---------------------------------------------------------------------------- 3D operator
size_t dlka_deform_conv3d_workspace_bytes(int B, int C, int D, int H, int W, int Co, int kd, int kh, int kw, int sd, int sh,
                                          int sw, int pd, int ph, int pw, int dild, int dilh, int dilw, int group,
                                          int deformable_group)
{
    ConvGeo g = make_geo(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, group, deformable_group, 3);
    if (bad_geo(g)) return 0;
    Arena ar(nullptr, 0);
    DeformOpPlan p;
    plan_deform_op(ar, g, false, p);
    return ar.off + 256;
}

int dlka_deform_conv3d_forward(const float *input, const float *weight, const float *bias, const float *offset, float *output,
                               int B, int C, int D, int H, int W, int Co, int kd, int kh, int kw, int sd, int sh, int sw, int pd,
                               int ph, int pw, int dild, int dilh, int dilw, int group, int deformable_group, int im2col_step,
                               int math, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!input || !weight || !bias || !offset || !output) return DLKA_ERR_INVALID_ARGUMENT;
    if (group <= 0 || deformable_group <= 0 || im2col_step <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, group, deformable_group, 3);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    const int step = B < im2col_step ? B : im2col_step;  // deform_conv_cuda.cu:61-63
    if (B % step != 0) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_TRY(check_device());
    return run_deform_op(g, input, weight, bias, offset, nullptr, output, math, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------- 3D operator, backward (row N2)
namespace {
struct DeformBwdPlan {
    float *x_cl, *off_cl, *gout_cl, *gin_cl, *goff_cl, *wt, *gwt, *colbuf, *colT, *gchunk, *gchunkT, *partial, *wscratch;
    int Mc;
};

bool plan_deform_bwd(Arena &ar, const ConvGeo &g, DeformBwdPlan &p)
{
    const size_t Vi = (size_t)g.D * g.H * g.W, M = (size_t)g.B * g.Do * g.Ho * g.Wo, KC = (size_t)g.K * g.C;
    p.Mc = deform3d_bwd_chunk_rows((i64)M);
    p.x_cl = ar.take<float>(g.B * Vi * g.C);
    p.gin_cl = ar.take<float>(g.B * Vi * g.C);
    p.off_cl = ar.take<float>(M * 3 * g.K * g.dg);
    p.goff_cl = ar.take<float>(M * 3 * g.K * g.dg);
    p.gout_cl = ar.take<float>(M * g.Co);
    p.wt = ar.take<float>(KC * g.Co);
    p.gwt = ar.take<float>(KC * g.Co);
    p.colbuf = ar.take<float>((size_t)p.Mc * KC);
    p.colT = ar.take<float>((size_t)p.Mc * KC);
    p.gchunk = ar.take<float>((size_t)p.Mc * g.Co);
    p.gchunkT = ar.take<float>((size_t)p.Mc * g.Co);
    p.partial = ar.take<float>(((size_t)p.Mc / 256 + 1) * KC * g.Co);   // split-K slices of >= 8 K steps of >= 32 rows
    const size_t s1 = dense_scratch_floats((int)KC, g.Co), s2 = dense_scratch_floats(g.Co, p.Mc);
    p.wscratch = ar.take<float>(s1 > s2 ? s1 : s2);
    return ar.ok();
}
}  // namespace

size_t dlka_deform_conv3d_backward_workspace_bytes(int B, int C, int D, int H, int W, int Co, int kd, int kh, int kw, int sd, int sh,
                                                   int sw, int pd, int ph, int pw, int dild, int dilh, int dilw, int group,
                                                   int deformable_group)
{
    ConvGeo g = make_geo(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, group, deformable_group, 3);
    if (bad_geo(g)) return 0;
    Arena ar(nullptr, 0);
    DeformBwdPlan p;
    plan_deform_bwd(ar, g, p);
    return ar.off + 256;
}

// D3D.deform_conv_backward (3D/dcn/src/vision.cpp:6, cuda/deform_conv_cuda.cu:128-285): all tensors in the reference's layouts
// (NCDHW, offset [B, 3*K, Do, Ho, Wo], weight [Co, C, kd, kh, kw]); the four gradients are fully overwritten.
int dlka_deform_conv3d_backward(const float *input, const float *weight, const float *offset, const float *grad_output,
                                float *grad_input, float *grad_offset, float *grad_weight, float *grad_bias, int B, int C, int D, int H,
                                int W, int Co, int kd, int kh, int kw, int sd, int sh, int sw, int pd, int ph, int pw, int dild, int dilh,
                                int dilw, int group, int deformable_group, int im2col_step, int math, void *workspace,
                                size_t workspace_bytes, void *stream)
{
    if (!input || !weight || !offset || !grad_output || !grad_input || !grad_offset || !grad_weight || !grad_bias)
        return DLKA_ERR_INVALID_ARGUMENT;
    if (group <= 0 || deformable_group <= 0 || im2col_step <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, group, deformable_group, 3);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    const int step = B < im2col_step ? B : im2col_step;  // deform_conv_cuda.cu:176-178
    if (B % step != 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % group != 0 || Co % group != 0 || C % deformable_group != 0) return DLKA_ERR_INVALID_ARGUMENT;   // deform_conv_cuda.cu:160-166
    if ((C / deformable_group) % 4 != 0) return DLKA_ERR_UNSUPPORTED;   // float4 channel vectors per deformable group
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    DeformBwdPlan p;
    if (!plan_deform_bwd(ar, g, p)) return DLKA_ERR_WORKSPACE;
    const i64 Vi = (i64)g.D * g.H * g.W, Vo = (i64)g.Do * g.Ho * g.Wo;
    DLKA_TRY(transpose_cs_to_sc(input, p.x_cl, g.B, g.C, Vi, st));
    DLKA_TRY(transpose_cs_to_sc(offset, p.off_cl, g.B, 3 * g.K * g.dg, Vo, st));
    DLKA_TRY(transpose_cs_to_sc(grad_output, p.gout_cl, g.B, g.Co, Vo, st));
    DLKA_CUDA_TRY(cudaMemsetAsync(p.gin_cl, 0, (size_t)g.B * Vi * g.C * sizeof(float), st));
    DLKA_TRY(deform3d_backward_cl(g, p.x_cl, p.off_cl, weight, p.gout_cl, p.gin_cl, p.goff_cl, grad_weight, grad_bias, p.wt, p.gwt,
                                  p.colbuf, p.colT, p.gchunk, p.gchunkT, p.partial, p.wscratch, math, st));
    DLKA_TRY(transpose_sc_to_cs(p.gin_cl, grad_input, g.B, g.C, Vi, st));
    DLKA_TRY(transpose_sc_to_cs(p.goff_cl, grad_offset, g.B, 3 * g.K * g.dg, Vo, st));
    return DLKA_OK;
}

int dlka_deform_conv3d_sample_indices(const float *offset, int32_t *low, int32_t *mask, int B, int D, int H, int W, int kd, int kh,
                                      int kw, int sd, int sh, int sw, int pd, int ph, int pw, int dild, int dilh, int dilw,
                                      int deformable_group, void *stream)
{
    if (!offset || !low || !mask) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, deformable_group, D, H, W, deformable_group, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, 1,
                         deformable_group, 3);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_TRY(check_device());
    return sample_indices(offset, low, mask, g, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------- 2D operator
size_t dlka_deform_conv2d_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                                          int dilh, int dilw, int n_weight_grps, int n_offset_grps)
{
    ConvGeo g = make_geo(B, C, 1, H, W, Co, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, n_weight_grps, n_offset_grps, 2);
    if (bad_geo(g)) return 0;
    Arena ar(nullptr, 0);
    DeformOpPlan p;
    plan_deform_op(ar, g, true, p);
    return ar.off + 256;
}

int dlka_deform_conv2d_forward(const float *input, const float *weight, const float *offset, const float *mask, const float *bias,
                               float *output, int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                               int dilh, int dilw, int n_weight_grps, int n_offset_grps, int math, void *workspace,
                               size_t workspace_bytes, void *stream)
{
    if (!input || !weight || !offset || !output) return DLKA_ERR_INVALID_ARGUMENT;
    if (n_weight_grps <= 0 || n_offset_grps <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, C, 1, H, W, Co, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, n_weight_grps, n_offset_grps, 2);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_TRY(check_device());
    return run_deform_op(g, input, weight, bias, offset, mask, output, math, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- backward of the 2D operator (row N2, 2D half): what torchvision's deform_conv2d autograd returns ----
namespace {
struct Deform2dBwdPlan {
    float *x_cl, *off_cl, *mask_cl, *gout_cl, *gx_cl, *goff_cl, *gmask_cl;
};
bool plan_deform2d_bwd(Arena &ar, const ConvGeo &g, bool has_mask, Deform2dBwdPlan &p)
{
    const size_t Vi = (size_t)g.B * g.H * g.W, M = (size_t)g.B * g.Ho * g.Wo;
    p.x_cl = ar.take<float>(Vi * g.C);
    p.gx_cl = ar.take<float>(Vi * g.C);
    p.off_cl = ar.take<float>(M * g.dg * 2 * g.K);
    p.goff_cl = ar.take<float>(M * g.dg * 2 * g.K);
    p.mask_cl = has_mask ? ar.take<float>(M * g.dg * g.K) : nullptr;
    p.gmask_cl = has_mask ? ar.take<float>(M * g.dg * g.K) : nullptr;
    p.gout_cl = ar.take<float>(M * g.Co);
    return ar.ok();
}
}  // namespace

size_t dlka_deform_conv2d_backward_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                                                   int dilh, int dilw, int n_weight_grps, int n_offset_grps, int has_mask)
{
    ConvGeo g = make_geo(B, C, 1, H, W, Co, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, n_weight_grps, n_offset_grps, 2);
    if (bad_geo(g)) return 0;
    Arena ar(nullptr, 0);
    Deform2dBwdPlan p;
    plan_deform2d_bwd(ar, g, has_mask != 0, p);
    return ar.off + 256;
}

int dlka_deform_conv2d_backward(const float *input, const float *weight, const float *offset, const float *mask,
                                const float *grad_output, float *grad_input, float *grad_weight, float *grad_offset, float *grad_mask,
                                float *grad_bias, int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                                int dilh, int dilw, int n_weight_grps, int n_offset_grps, void *workspace, size_t workspace_bytes,
                                void *stream)
{
    if (!input || !weight || !offset || !grad_output || !grad_input || !grad_weight || !grad_offset) return DLKA_ERR_INVALID_ARGUMENT;
    if ((mask == nullptr) != (grad_mask == nullptr)) return DLKA_ERR_INVALID_ARGUMENT;
    if (n_weight_grps <= 0 || n_offset_grps <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, C, 1, H, W, Co, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, n_weight_grps, n_offset_grps, 2);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    Deform2dBwdPlan p;
    if (!plan_deform2d_bwd(ar, g, mask != nullptr, p)) return DLKA_ERR_WORKSPACE;
    const i64 Vi = (i64)g.H * g.W, Vo = (i64)g.Ho * g.Wo;
    DLKA_TRY(transpose_cs_to_sc(input, p.x_cl, g.B, g.C, Vi, st));
    DLKA_TRY(transpose_cs_to_sc(offset, p.off_cl, g.B, g.dg * 2 * g.K, Vo, st));
    if (mask) DLKA_TRY(transpose_cs_to_sc(mask, p.mask_cl, g.B, g.dg * g.K, Vo, st));
    DLKA_TRY(transpose_cs_to_sc(grad_output, p.gout_cl, g.B, g.Co, Vo, st));
    DLKA_TRY(deform2d_backward_cl(g, p.x_cl, weight, p.off_cl, p.mask_cl, p.gout_cl, p.gx_cl, grad_weight, p.goff_cl, p.gmask_cl, grad_bias, st));
    DLKA_TRY(transpose_sc_to_cs(p.gx_cl, grad_input, g.B, g.C, Vi, st));
    DLKA_TRY(transpose_sc_to_cs(p.goff_cl, grad_offset, g.B, g.dg * 2 * g.K, Vo, st));
    if (mask) DLKA_TRY(transpose_sc_to_cs(p.gmask_cl, grad_mask, g.B, g.dg * g.K, Vo, st));
    return DLKA_OK;
}

int dlka_deform_conv2d_sample_indices(const float *offset, int32_t *low, int32_t *mask, int B, int H, int W, int kh, int kw, int sh,
                                      int sw, int ph, int pw, int dilh, int dilw, int n_offset_grps, void *stream)
{
    if (!offset || !low || !mask) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, n_offset_grps, 1, H, W, n_offset_grps, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, 1, n_offset_grps, 2);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_TRY(check_device());
    return sample_indices(offset, low, mask, g, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------- pack operators
namespace {
struct PackPlan {
    float *x_cl, *off_cl, *y_cl, *wp_off, *wp;
    int np_off, np;
};

bool plan_pack(Arena &ar, const ConvGeo &g, PackPlan &p)
{
    const i64 Vi = (i64)g.D * g.H * g.W, M = (i64)g.B * g.Do * g.Ho * g.Wo;
    const int noff = g.dg * g.ndim * g.K;
    p.x_cl = ar.take<float>((size_t)g.B * Vi * g.C);
    p.off_cl = ar.take<float>((size_t)M * noff);
    p.y_cl = ar.take<float>((size_t)M * g.Co);
    p.np_off = p.np = 0;
    p.wp_off = ar.take<float>(contraction_scratch_floats(noff, g.C, g.K, 1));
    if (is_depthwise(g))
        p.wp = ar.take<float>((size_t)g.K * g.C);
    else
        p.wp = ar.take<float>(contraction_scratch_floats(g.Co, g.C, g.K, g.groups));
    return ar.ok();
}

// g: geometry of the deformable conv; go: geometry of the offset conv (same taps, its own dilation)
int run_pack(const ConvGeo &g, const ConvGeo &go, const float *input, const float *offset_weight, const float *offset_bias,
             const float *weight, const float *bias, float *output, int math, void *workspace, size_t workspace_bytes,
             cudaStream_t st)
{
    if (go.Do != g.Do || go.Ho != g.Ho || go.Wo != g.Wo) return DLKA_ERR_INVALID_ARGUMENT;
    if (g.C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    Arena ar(workspace, workspace_bytes);
    PackPlan p;
    if (!plan_pack(ar, g, p)) return DLKA_ERR_WORKSPACE;
    const i64 Vi = (i64)g.D * g.H * g.W, Vo = (i64)g.Do * g.Ho * g.Wo;
    const int noff = g.dg * g.ndim * g.K;
    DLKA_TRY(transpose_cs_to_sc(input, p.x_cl, g.B, g.C, Vi, st));
    IgemmArgs ao = conv_args(IGEMM_CONV, go, p.x_cl, nullptr, nullptr, nullptr, 0, offset_bias, EPI_NONE, nullptr, 0, p.off_cl, noff);
    DLKA_TRY(contraction(ao, offset_weight, math, p.wp_off, st));
    if (is_depthwise(g)) {
        DLKA_TRY(deform_dwconv_cl(p.x_cl, p.off_cl, nullptr, weight, bias, p.y_cl, g, p.wp, st));
    } else {
        if ((g.C / g.groups) % 4 != 0 || (g.C / g.dg) % 4 != 0) return DLKA_ERR_UNSUPPORTED;
        IgemmArgs a = conv_args(IGEMM_DEFORM, g, p.x_cl, p.off_cl, nullptr, nullptr, 0, bias, EPI_NONE, nullptr, 0, p.y_cl, g.Co);
        DLKA_TRY(contraction(a, weight, math, p.wp, st));
    }
    DLKA_TRY(transpose_sc_to_cs(p.y_cl, output, g.B, g.Co, Vo, st));
    return DLKA_OK;
}
}  // namespace

size_t dlka_deform_conv_pack3d_workspace_bytes(int B, int C, int D, int H, int W, int Co, int kd, int kh, int kw, int sd, int sh,
                                               int sw, int pd, int ph, int pw, int dild, int dilh, int dilw, int group,
                                               int deformable_group)
{
    ConvGeo g = make_geo(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, group, deformable_group, 3);
    if (bad_geo(g)) return 0;
    Arena ar(nullptr, 0);
    PackPlan p;
    plan_pack(ar, g, p);
    return ar.off + 256;
}

int dlka_deform_conv_pack3d_forward(const float *input, const float *offset_weight, const float *offset_bias, const float *weight,
                                    const float *bias, float *output, int B, int C, int D, int H, int W, int Co, int kd, int kh,
                                    int kw, int sd, int sh, int sw, int pd, int ph, int pw, int dild, int dilh, int dilw, int group,
                                    int deformable_group, int im2col_step, int math, void *workspace, size_t workspace_bytes,
                                    void *stream)
{
    if (!input || !offset_weight || !offset_bias || !weight || !bias || !output) return DLKA_ERR_INVALID_ARGUMENT;
    if (group <= 0 || deformable_group <= 0 || im2col_step <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, C, D, H, W, Co, kd, kh, kw, sd, sh, sw, pd, ph, pw, dild, dilh, dilw, group, deformable_group, 3);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    const int step = B < im2col_step ? B : im2col_step;
    if (B % step != 0) return DLKA_ERR_INVALID_ARGUMENT;
    // conv_offset ignores the dilation (synapse/deform_conv.py:80-85)
    ConvGeo go = make_geo(B, C, D, H, W, deformable_group * 3 * g.K, kd, kh, kw, sd, sh, sw, pd, ph, pw, 1, 1, 1, 1, 1, 3);
    if (bad_geo(go)) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_TRY(check_device());
    return run_pack(g, go, input, offset_weight, offset_bias, weight, bias, output, math, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

size_t dlka_deform_conv_pack2d_workspace_bytes(int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                                               int dilh, int dilw, int groups)
{
    ConvGeo g = make_geo(B, C, 1, H, W, Co, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, groups, 1, 2);
    if (bad_geo(g)) return 0;
    Arena ar(nullptr, 0);
    PackPlan p;
    plan_pack(ar, g, p);
    return ar.off + 256;
}

int dlka_deform_conv_pack2d_forward(const float *input, const float *offset_weight, const float *offset_bias, const float *weight,
                                    const float *bias, float *output, int B, int C, int H, int W, int Co, int kh, int kw, int sh,
                                    int sw, int ph, int pw, int dilh, int dilw, int groups, int math, void *workspace,
                                    size_t workspace_bytes, void *stream)
{
    if (!input || !offset_weight || !offset_bias || !weight || !output) return DLKA_ERR_INVALID_ARGUMENT;
    if (groups <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo g = make_geo(B, C, 1, H, W, Co, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, groups, 1, 2);
    if (bad_geo(g)) return DLKA_ERR_INVALID_ARGUMENT;
    ConvGeo go = make_geo(B, C, 1, H, W, 2 * g.K, 1, kh, kw, 1, sh, sw, 0, ph, pw, 1, dilh, dilw, 1, 1, 2);
    DLKA_TRY(check_device());
    return run_pack(g, go, input, offset_weight, offset_bias, weight, bias, output, math, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------- 3D block
size_t dlka_lka3d_deform_workspace_bytes(int B, int C, int D1, int D2, int D3)
{
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return 0;
    Arena ar(nullptr, 0);
    Block3dPlan p;
    plan_block3d(ar, B, C, D1, D2, D3, p);
    return ar.off + 256;
}

int dlka_lka3d_deform_forward(const dlkaBlock3dParams *params, const float *x, float *y, int B, int C, int D1, int D2, int D3,
                              int math, void *workspace, size_t workspace_bytes, void *stream)
{
    if (null_params3d(params, false) || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    Block3dPlan p;
    if (!plan_block3d(ar, B, C, D1, D2, D3, p)) return DLKA_ERR_WORKSPACE;
    const i64 S = (i64)D1 * D2 * D3;
    DLKA_TRY(transpose_cs_to_sc(x, p.t1, B, C, S, st));              // u = x (channels-last)
    DLKA_TRY(run_lka3d_core(*params, p.t1, p, B, C, D1, D2, D3, math, st));  // gate -> t3
    DLKA_TRY(transpose_sc_to_cs(p.t3, y, B, C, S, st));
    return DLKA_OK;
}

size_t dlka_lka_attention3d_deform_workspace_bytes(int B, int C, int D1, int D2, int D3)
{
    return dlka_lka3d_deform_workspace_bytes(B, C, D1, D2, D3);
}

namespace {
int attention3d_impl(const dlkaBlock3dParams *params, const float *x, float *y, int B, int C, int D1, int D2, int D3, int math,
                     void *packed, size_t packed_bytes, int packed_valid, void *workspace, size_t workspace_bytes, void *stream)
{
    if (null_params3d(params, true) || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes), pk(packed, packed_bytes);
    Block3dPlan p;
    if (!plan_block3d(ar, B, C, D1, D2, D3, p, packed ? &pk : nullptr)) return DLKA_ERR_WORKSPACE;
    PackSkipScope skip(packed != nullptr && packed_valid != 0);
    const i64 M = (i64)B * D1 * D2 * D3;
    // tokens [B,N,C] are already channels-last over the Conv3d volume (transformerblock.py:665)
    IgemmArgs a1 = dense_args(x, C, M, C, C, nullptr, 0, params->proj_1_bias, EPI_GELU, nullptr, 0, p.t1, C);
    DLKA_TRY(contraction(a1, params->proj_1_weight, math, p.wp_proj1, st));
    const int rc = run_lka3d_core(*params, p.t1, p, B, C, D1, D2, D3, math, st, true, x, y);  // gate -> t3, or everything -> y
    if (rc == 1) return DLKA_OK;
    if (rc != DLKA_OK) return rc;
    IgemmArgs a2 = dense_args(p.t3, C, M, C, C, nullptr, 0, params->proj_2_bias, EPI_ADD, x, C, y, C);
    DLKA_TRY(contraction(a2, params->proj_2_weight, math, p.wp_proj2, st));
    return DLKA_OK;
}
}  // namespace

int dlka_lka_attention3d_deform_forward(const dlkaBlock3dParams *params, const float *x, float *y, int B, int C, int D1, int D2,
                                        int D3, int math, void *workspace, size_t workspace_bytes, void *stream)
{
    return attention3d_impl(params, x, y, B, C, D1, D2, D3, math, nullptr, 0, 0, workspace, workspace_bytes, stream);
}

size_t dlka_lka_attention3d_deform_packed_bytes(int C)
{
    if (C <= 0) return 0;
    Arena ar(nullptr, 0), pk(nullptr, 0);
    Block3dPlan p;
    plan_block3d(ar, 1, C, 1, 1, 1, p, &pk);
    return pk.off + 256;
}

int dlka_lka_attention3d_deform_forward_packed(const dlkaBlock3dParams *params, const float *x, float *y, int B, int C, int D1,
                                               int D2, int D3, int math, void *packed, size_t packed_bytes, int packed_valid,
                                               void *workspace, size_t workspace_bytes, void *stream)
{
    if (!packed) return DLKA_ERR_INVALID_ARGUMENT;
    return attention3d_impl(params, x, y, B, C, D1, D2, D3, math, packed, packed_bytes, packed_valid, workspace, workspace_bytes, stream);
}

int dlka_lka_attention3d_deform_forward_host(const dlkaBlock3dParams *params, const float *x_host, float *y_host, int B, int C,
                                             int D1, int D2, int D3, int math, void *dev_scratch, size_t dev_scratch_bytes,
                                             void *workspace, size_t workspace_bytes, void *stream)
{
    if (!x_host || !y_host || !dev_scratch) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    const size_t n = (size_t)B * D1 * D2 * D3 * C;
    if (dev_scratch_bytes < 2 * n * sizeof(float)) return DLKA_ERR_WORKSPACE;
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    float *xd = (float *)dev_scratch, *yd = xd + n;
    // Per-sample software pipeline over three streams: H2D of sample b+1 and D2H of sample b-1 overlap the compute of
    // sample b (samples are independent: SURVEY.md 8e).  Compute stays on the caller's stream.
    // the helper streams live for this (synchronous) call only: nothing is bound to whichever device was current first
    cudaStream_t s_in = nullptr, s_out = nullptr;
    DLKA_CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    if (cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess) {
        const cudaError_t ce = cudaGetLastError();
        cudaStreamDestroy(s_in);
        return record_cuda_error(ce, "cudaStreamCreateWithFlags");
    }
    const size_t n1 = n / B;
    std::vector<cudaEvent_t> ev(2 * (size_t)B + 2, nullptr);
    int rc = DLKA_OK;
    for (auto &e : ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { e = nullptr; rc = DLKA_ERR_CUDA; break; }
    do {
        if (rc != DLKA_OK) break;
        cudaEvent_t e_start = ev[2 * B], e_done = ev[2 * B + 1];
        if (cudaEventRecord(e_start, st) != cudaSuccess || cudaStreamWaitEvent(s_in, e_start, 0) != cudaSuccess ||
            cudaStreamWaitEvent(s_out, e_start, 0) != cudaSuccess) { rc = DLKA_ERR_CUDA; break; }
        for (int b = 0; b < B && rc == DLKA_OK; ++b) {
            if (cudaMemcpyAsync(xd + b * n1, x_host + b * n1, n1 * sizeof(float), cudaMemcpyHostToDevice, s_in) != cudaSuccess ||
                cudaEventRecord(ev[b], s_in) != cudaSuccess) rc = DLKA_ERR_CUDA;
        }
        for (int b = 0; b < B && rc == DLKA_OK; ++b) {
            if (cudaStreamWaitEvent(st, ev[b], 0) != cudaSuccess) { rc = DLKA_ERR_CUDA; break; }
            rc = dlka_lka_attention3d_deform_forward(params, xd + b * n1, yd + b * n1, 1, C, D1, D2, D3, math, workspace,
                                                     workspace_bytes, stream);
            if (rc != DLKA_OK) break;
            if (cudaEventRecord(ev[B + b], st) != cudaSuccess || cudaStreamWaitEvent(s_out, ev[B + b], 0) != cudaSuccess ||
                cudaMemcpyAsync(y_host + b * n1, yd + b * n1, n1 * sizeof(float), cudaMemcpyDeviceToHost, s_out) != cudaSuccess)
                rc = DLKA_ERR_CUDA;
        }
        if (rc != DLKA_OK) break;
        if (cudaEventRecord(e_done, s_out) != cudaSuccess || cudaStreamWaitEvent(st, e_done, 0) != cudaSuccess) rc = DLKA_ERR_CUDA;
    } while (0);
    const cudaError_t first = rc == DLKA_ERR_CUDA ? cudaGetLastError() : cudaSuccess;
    const cudaError_t se = cudaStreamSynchronize(st);
    cudaStreamSynchronize(s_in);
    cudaStreamSynchronize(s_out);
    for (auto &e : ev)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(s_in);
    cudaStreamDestroy(s_out);
    if (rc == DLKA_ERR_CUDA && first != cudaSuccess) return record_cuda_error(first, "dlka_lka_attention3d_deform_forward_host");
    if (rc == DLKA_ERR_CUDA) return record_cuda_error(cudaGetLastError(), "dlka_lka_attention3d_deform_forward_host");
    if (rc != DLKA_OK) return rc;
    if (se != cudaSuccess) return record_cuda_error(se, "cudaStreamSynchronize");
    return DLKA_OK;
}

// ---------------------------------------------------------------------------- 2D block
size_t dlka_deformable_lka2d_workspace_bytes(int B, int C, int H, int W)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    Arena ar(nullptr, 0);
    Block2dPlan p;
    plan_block2d(ar, B, C, H, W, p);
    return ar.off + 256;
}

int dlka_deformable_lka2d_forward(const dlkaBlock2dParams *params, const float *x, float *y, int B, int C, int H, int W, int math,
                                  void *workspace, size_t workspace_bytes, void *stream)
{
    if (null_params2d(params, false) || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    Block2dPlan p;
    if (!plan_block2d(ar, B, C, H, W, p)) return DLKA_ERR_WORKSPACE;
    DLKA_TRY(transpose_cs_to_sc(x, p.x_cl, B, C, (i64)H * W, st));
    DLKA_TRY(run_lka2d_core(*params, p.x_cl, p, B, C, H, W, math, st));  // gate -> t2
    DLKA_TRY(transpose_sc_to_cs(p.t2, y, B, C, (i64)H * W, st));
    return DLKA_OK;
}

size_t dlka_deformable_lka_attention2d_workspace_bytes(int B, int C, int H, int W)
{
    return dlka_deformable_lka2d_workspace_bytes(B, C, H, W);
}

namespace {
int attention2d_impl(const dlkaBlock2dParams *params, const float *x, float *y, int B, int C, int H, int W, int math, void *packed,
                     size_t packed_bytes, int packed_valid, void *workspace, size_t workspace_bytes, void *stream)
{
    if (null_params2d(params, true) || !x || !y) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    if (C % 4 != 0) return DLKA_ERR_UNSUPPORTED;
    DLKA_TRY(check_device());
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes), pk(packed, packed_bytes);
    Block2dPlan p;
    if (!plan_block2d(ar, B, C, H, W, p, packed ? &pk : nullptr)) return DLKA_ERR_WORKSPACE;
    PackSkipScope skip(packed != nullptr && packed_valid != 0);
    DLKA_TRY(transpose_cs_to_sc(x, p.x_cl, B, C, (i64)H * W, st));
    DLKA_TRY(attention2d_cl_planned(*params, p.x_cl, p.t3, p, B, C, H, W, math, st));
    DLKA_TRY(transpose_sc_to_cs(p.t3, y, B, C, (i64)H * W, st));
    return DLKA_OK;
}
}  // namespace

int dlka_deformable_lka_attention2d_forward(const dlkaBlock2dParams *params, const float *x, float *y, int B, int C, int H, int W,
                                            int math, void *workspace, size_t workspace_bytes, void *stream)
{
    return attention2d_impl(params, x, y, B, C, H, W, math, nullptr, 0, 0, workspace, workspace_bytes, stream);
}

size_t dlka_deformable_lka_attention2d_packed_bytes(int C)
{
    if (C <= 0) return 0;
    Arena ar(nullptr, 0), pk(nullptr, 0);
    Block2dPlan p;
    plan_block2d(ar, 1, C, 1, 1, p, &pk);
    return pk.off + 256;
}

int dlka_deformable_lka_attention2d_forward_packed(const dlkaBlock2dParams *params, const float *x, float *y, int B, int C, int H,
                                                   int W, int math, void *packed, size_t packed_bytes, int packed_valid,
                                                   void *workspace, size_t workspace_bytes, void *stream)
{
    if (!packed) return DLKA_ERR_INVALID_ARGUMENT;
    return attention2d_impl(params, x, y, B, C, H, W, math, packed, packed_bytes, packed_valid, workspace, workspace_bytes, stream);
}

}  // extern "C"
