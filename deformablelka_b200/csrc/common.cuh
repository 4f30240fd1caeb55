// common.cuh -- shared host/device helpers for libdlka_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "../../include/dlka.h"

typedef long long i64;

// ---------------------------------------------------------------------------------------
// host-side bookkeeping
// ---------------------------------------------------------------------------------------
namespace dlka {

extern thread_local char g_last_cuda_error[256];
void note_launch(int n = 1);
int record_cuda_error(cudaError_t e, const char *what);

#define DLKA_CUDA_TRY(expr)                                                   \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return dlka::record_cuda_error(_e, #expr);     \
    } while (0)

#define DLKA_CHECK_LAUNCH(name)                                               \
    do {                                                                      \
        dlka::note_launch();                                                  \
        cudaError_t _e = cudaGetLastError();                                  \
        if (_e != cudaSuccess) return dlka::record_cuda_error(_e, name);      \
    } while (0)

// Optional per-kernel CUDA-event timing (dlka_profile_enable): start/stop events on the launch stream.
struct KernelScope {
    const char *name;
    cudaStream_t st;
    cudaEvent_t a, b;
    bool active;
    KernelScope(const char *name, cudaStream_t st);
    ~KernelScope();
};

#define DLKA_LAUNCH(name, st, ...)                                            \
    do {                                                                      \
        {                                                                     \
            dlka::KernelScope _ks(name, st);                                  \
            __VA_ARGS__;                                                      \
        }                                                                     \
        DLKA_CHECK_LAUNCH(name);                                              \
    } while (0)

#define DLKA_TRY(expr)                                                        \
    do {                                                                      \
        int _s = (expr);                                                      \
        if (_s != DLKA_OK) return _s;                                         \
    } while (0)

// Opt-in to more than 48 KB of dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is kept per function AND
// per device (context), so the cache is indexed by the current device ordinal: one process driving several GPUs
// (nn.DataParallel, the reference 2D trainers' multi-GPU mode) configures each of them.  Monotonic and mutex-protected, so
// a second host thread (the autograd engine's) can never lower the limit under a launch that needs more.
struct SmemOptIn {
    static constexpr int MAX_DEVICES = 64;
    std::atomic<size_t> configured[MAX_DEVICES];
    std::mutex lock;
    template <typename Kern>
    int ensure(Kern kern, size_t smem)
    {
        int dev = 0;
        DLKA_CUDA_TRY(cudaGetDevice(&dev));
        if (dev < 0 || dev >= MAX_DEVICES) {   // beyond the cache: set it every time (cheap)
            DLKA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            return DLKA_OK;
        }
        if (smem > configured[dev].load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> guard(lock);
            if (smem > configured[dev].load(std::memory_order_relaxed)) {
                DLKA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured[dev].store(smem, std::memory_order_release);
            }
        }
        return DLKA_OK;
    }
};

// Prepacked weights (dlka_*_forward_packed): while a PackSkipScope is alive on this HOST thread, the weight-packing launchers
// (tc_pack_weight, pack_weight, pack_dw, deform3d_ps_pack, the packers inside conv_tiled_ex / deform3d_tc, pack_dw9) return
// without launching -- the caller has promised that the packed buffers already hold these weights.  Host-side, per thread,
// scoped to one entry-point call: no device state.
bool pack_skipped();
struct PackSkipScope {
    bool prev;
    explicit PackSkipScope(bool skip);
    ~PackSkipScope();
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline i64 cdiv(i64 a, i64 b) { return (a + b - 1) / b; }

// Bump allocator over the caller's workspace.
struct Arena {
    char *base;
    size_t cap, off;
    Arena(void *p, size_t bytes) : base((char *)p), cap(bytes), off(0) {}
    template <typename T>
    T *take(size_t n)
    {
        off = align_up(off, 256);
        T *r = (T *)(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap && (base != nullptr || off == 0); }
};

// Geometry of a (deformable) convolution over a channels-last volume.  2D uses D = kd = 1.
struct ConvGeo {
    int B, C, D, H, W;     // input extent
    int Co, Do, Ho, Wo;    // output extent
    int kd, kh, kw, sd, sh, sw, pd, ph, pw, dd, dh, dw;
    int K;                 // taps = kd*kh*kw
    int groups, dg;        // weight groups, deformable (offset) groups
    int ndim;              // 2 or 3: number of offset components per tap
};

static inline int out_extent(int in, int pad, int dil, int k, int stride)
{
    return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;  // deform_conv_cuda.cu:78-80
}

static inline ConvGeo make_geo(int B, int C, int D, int H, int W, int Co, int kd, int kh, int kw, int sd, int sh,
                               int sw, int pd, int ph, int pw, int dd, int dh, int dw, int groups, int dg, int ndim)
{
    ConvGeo g;
    g.B = B; g.C = C; g.D = D; g.H = H; g.W = W; g.Co = Co;
    g.kd = kd; g.kh = kh; g.kw = kw; g.sd = sd; g.sh = sh; g.sw = sw;
    g.pd = pd; g.ph = ph; g.pw = pw; g.dd = dd; g.dh = dh; g.dw = dw;
    g.Do = out_extent(D, pd, dd, kd, sd); g.Ho = out_extent(H, ph, dh, kh, sh); g.Wo = out_extent(W, pw, dw, kw, sw);
    g.K = kd * kh * kw; g.groups = groups; g.dg = dg; g.ndim = ndim;
    return g;
}

}  // namespace dlka

// ---------------------------------------------------------------------------------------
// device: the sampler.  Restates the arithmetic of dmcn_im2col_bilinear and its caller
// (3D/dcn/src/cuda/deform_im2col_cuda.cuh:26-72, :224-226, :245-248) for channels-last data.
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace dlka {

// p = float(o*stride - pad + tap*dil) + delta : integer base first, then ONE fp32 add.
__device__ __forceinline__ float sample_pos(int o, int stride, int pad, int tap, int dil, float delta)
{
    return __fadd_rn((float)(o * stride - pad + tap * dil), delta);
}

struct Sample3 {
    int lo[3];     // floor(p) per axis (d, h, w)
    float l[3];    // p - floor(p)
    int mask;      // bit0: sample valid; bits 1..8: corner v1..v8 is read
};

__device__ __forceinline__ Sample3 make_sample3(float pd, float ph, float pw, int D, int H, int W)
{
    Sample3 s;
    s.lo[0] = (int)floorf(pd); s.lo[1] = (int)floorf(ph); s.lo[2] = (int)floorf(pw);
    s.l[0] = pd - (float)s.lo[0]; s.l[1] = ph - (float)s.lo[1]; s.l[2] = pw - (float)s.lo[2];
    int m = 0;
    if (pd > -1.f && ph > -1.f && pw > -1.f && pd < (float)D && ph < (float)H && pw < (float)W) {
        const int dl = s.lo[0] >= 0, hl = s.lo[1] >= 0, wl = s.lo[2] >= 0;
        const int dh = s.lo[0] + 1 <= D - 1, hh = s.lo[1] + 1 <= H - 1, wh = s.lo[2] + 1 <= W - 1;
        m = 1 | ((dl & hl & wl) << 1) | ((dl & hl & wh) << 2) | ((dl & hh & wl) << 3) | ((dl & hh & wh) << 4) |
            ((dh & hl & wl) << 5) | ((dh & hl & wh) << 6) | ((dh & hh & wl) << 7) | ((dh & hh & wh) << 8);
    }
    s.mask = m;
    return s;
}

// 2D sampler (torchvision deform_conv2d bilinear_interpolate): same rules on (h, w).
struct Sample2 {
    int lo[2];
    float l[2];
    int mask;      // bit0 valid; bits 1..4 corners v1..v4
};

__device__ __forceinline__ Sample2 make_sample2(float ph, float pw, int H, int W)
{
    Sample2 s;
    s.lo[0] = (int)floorf(ph); s.lo[1] = (int)floorf(pw);
    s.l[0] = ph - (float)s.lo[0]; s.l[1] = pw - (float)s.lo[1];
    int m = 0;
    if (ph > -1.f && pw > -1.f && ph < (float)H && pw < (float)W) {
        const int hl = s.lo[0] >= 0, wl = s.lo[1] >= 0, hh = s.lo[0] + 1 <= H - 1, wh = s.lo[1] + 1 <= W - 1;
        m = 1 | ((hl & wl) << 1) | ((hl & wh) << 2) | ((hh & wl) << 3) | ((hh & wh) << 4);
    }
    s.mask = m;
    return s;
}

__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &v);
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// Trilinear blend of 4 consecutive channels at a sample; `vol` points at channel c of sample b
// (channels-last volume [D,H,W,C]).  Weight products follow cuh:67-68 (hd*hh*hw ...).
__device__ __forceinline__ float4 trilinear4(const float *__restrict__ vol, const Sample3 &s, int H, int W, int C)
{
    float4 acc = f4zero();
    if (!(s.mask & 1)) return acc;
    const float ld = s.l[0], lh = s.l[1], lw = s.l[2];
    const float hd = 1.f - ld, hh = 1.f - lh, hw = 1.f - lw;
    const i64 sH = (i64)W * C, sD = (i64)H * sH;
    const float *p000 = vol + (i64)s.lo[0] * sD + (i64)s.lo[1] * sH + (i64)s.lo[2] * C;
    if (s.mask & (1 << 1)) fma4(acc, hd * hh * hw, ldg4(p000));
    if (s.mask & (1 << 2)) fma4(acc, hd * hh * lw, ldg4(p000 + C));
    if (s.mask & (1 << 3)) fma4(acc, hd * lh * hw, ldg4(p000 + sH));
    if (s.mask & (1 << 4)) fma4(acc, hd * lh * lw, ldg4(p000 + sH + C));
    if (s.mask & (1 << 5)) fma4(acc, ld * hh * hw, ldg4(p000 + sD));
    if (s.mask & (1 << 6)) fma4(acc, ld * hh * lw, ldg4(p000 + sD + C));
    if (s.mask & (1 << 7)) fma4(acc, ld * lh * hw, ldg4(p000 + sD + sH));
    if (s.mask & (1 << 8)) fma4(acc, ld * lh * lw, ldg4(p000 + sD + sH + C));
    return acc;
}

__device__ __forceinline__ float4 bilinear4(const float *__restrict__ img, const Sample2 &s, int W, int C)
{
    float4 acc = f4zero();
    if (!(s.mask & 1)) return acc;
    const float lh = s.l[0], lw = s.l[1], hh = 1.f - lh, hw = 1.f - lw;
    const i64 sH = (i64)W * C;
    const float *p00 = img + (i64)s.lo[0] * sH + (i64)s.lo[1] * C;
    if (s.mask & (1 << 1)) fma4(acc, hh * hw, ldg4(p00));
    if (s.mask & (1 << 2)) fma4(acc, hh * lw, ldg4(p00 + C));
    if (s.mask & (1 << 3)) fma4(acc, lh * hw, ldg4(p00 + sH));
    if (s.mask & (1 << 4)) fma4(acc, lh * lw, ldg4(p00 + sH + C));
    return acc;
}

// packed 2 x fp32 FMA (Blackwell FFMA2): acc += w * x element-wise on a float4 (2 instructions instead of 4)
__device__ __forceinline__ void fma4v(float4 &acc, const float4 &w, const float4 &x)
{
    unsigned long long a0, a1, w0, w1, x0, x1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(acc.z), "f"(acc.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(w0) : "f"(w.x), "f"(w.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(w1) : "f"(w.z), "f"(w.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "f"(x.z), "f"(x.w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(w0), "l"(x0));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(w1), "l"(x1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a1));
}
__device__ __forceinline__ void fma2v(float2 &acc, const float2 &w, const float2 &x)
{
    unsigned long long a0, w0, x0;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(w0) : "f"(w.x), "f"(w.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(w0), "l"(x0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a0));
}
// acc += s * x with a scalar weight (broadcast into both halves)
__device__ __forceinline__ void fma4s(float4 &acc, float s, const float4 &x)
{
    unsigned long long a0, a1, sw, x0, x1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(acc.z), "f"(acc.w));
    asm("mov.b64 %0, {%1, %1};" : "=l"(sw) : "f"(s));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "f"(x.z), "f"(x.w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(sw), "l"(x0));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(sw), "l"(x1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a1));
}

__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &v) { fma4s(acc, w, v); }

// libm erf: used by the fp32 validation path (igemm_simt.cu)
__device__ __forceinline__ float gelu_erf_libm(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// exact-erf GELU  x * Phi(x), Phi(x) = 0.5 erfc(-x / sqrt 2), branch-free (15 instructions instead of erff's ~45 with
// divergent ranges: the proj_1 epilogue evaluates it 2e8 times per call).  erfc(z), z >= 0, by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7 absolute), evaluated on the erfc side for x < 0 so the left tail keeps its relative accuracy.
__device__ __forceinline__ float gelu_erf(float x)
{
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float y = 0.5f * p * t * __expf(-z * z);   // 0.5 erfc(z)
    return x * (x < 0.f ? y : 1.f - y);
}

}  // namespace dlka
#endif
