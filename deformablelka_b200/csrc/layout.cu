// layout.cu -- channel-first <-> channels-last transposes and the sampler integer planes.
#include "kernels.cuh"

namespace dlka {
namespace {

// in [B][C][S] -> out [B][S][C] with S on grid.x (S can be ~1e6, C is small)
__global__ void transpose_cs_sc_kernel(const float *__restrict__ in, float *__restrict__ out, int C, i64 S)
{
    __shared__ float tile[32][33];
    const i64 b = blockIdx.z;
    const i64 s0 = (i64)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float *src = in + b * C * S;
    float *dst = out + b * C * S;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const i64 s = s0 + threadIdx.x;
        if (c < C && s < S) tile[i][threadIdx.x] = src[(i64)c * S + s];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const i64 s = s0 + i;
        const int c = c0 + threadIdx.x;
        if (c < C && s < S) dst[s * C + c] = tile[threadIdx.x][i];
    }
}

// in [B][C][S] -> out CHUNK-MAJOR [C/32][B][S][32] (the gather layout of deform_ps.cu); C % 32 == 0, blockIdx.y = chunk
__global__ void transpose_cs_chunk_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int C, i64 S)
{
    __shared__ float tile[32][33];
    const i64 b = blockIdx.z;
    const i64 s0 = (i64)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float *src = in + b * C * S;
    float *dst = out + ((i64)blockIdx.y * B + b) * S * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const i64 s = s0 + threadIdx.x;
        if (s < S) tile[i][threadIdx.x] = src[(i64)(c0 + i) * S + s];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const i64 s = s0 + i;
        if (s < S) dst[s * 32 + threadIdx.x] = tile[threadIdx.x][i];
    }
}

__global__ void transpose_sc_cs_kernel(const float *__restrict__ in, float *__restrict__ out, int C, i64 S)
{
    __shared__ float tile[32][33];
    const i64 b = blockIdx.z;
    const i64 s0 = (i64)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float *src = in + b * C * S;
    float *dst = out + b * C * S;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const i64 s = s0 + i;
        const int c = c0 + threadIdx.x;
        if (c < C && s < S) tile[i][threadIdx.x] = src[s * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const i64 s = s0 + threadIdx.x;
        if (c < C && s < S) dst[(i64)c * S + s] = tile[threadIdx.x][i];
    }
}

__global__ void sample_indices_kernel(const float *__restrict__ off, int32_t *__restrict__ low, int32_t *__restrict__ mask,
                                      const ConvGeo g)
{
    const i64 Vo = (i64)g.Do * g.Ho * g.Wo;
    const i64 total = (i64)g.B * g.dg * Vo * g.K;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const int t = (int)(idx % g.K);
        const i64 v = (idx / g.K) % Vo;
        const i64 bg = idx / ((i64)g.K * Vo);
        const int wo = (int)(v % g.Wo), ho = (int)((v / g.Wo) % g.Ho), d_o = (int)(v / ((i64)g.Wo * g.Ho));
        const int kk = t % g.kw, jj = (t / g.kw) % g.kh, ii = t / (g.kw * g.kh);
        const float *o = off + bg * g.ndim * g.K * Vo + v;
        if (g.ndim == 3) {
            const float pd = sample_pos(d_o, g.sd, g.pd, ii, g.dd, o[(i64)(3 * t + 0) * Vo]);
            const float ph = sample_pos(ho, g.sh, g.ph, jj, g.dh, o[(i64)(3 * t + 1) * Vo]);
            const float pw = sample_pos(wo, g.sw, g.pw, kk, g.dw, o[(i64)(3 * t + 2) * Vo]);
            const Sample3 s = make_sample3(pd, ph, pw, g.D, g.H, g.W);
            low[idx * 3 + 0] = s.lo[0]; low[idx * 3 + 1] = s.lo[1]; low[idx * 3 + 2] = s.lo[2];
            mask[idx] = s.mask;
        } else {
            const float ph = sample_pos(ho, g.sh, g.ph, jj, g.dh, o[(i64)(2 * t + 0) * Vo]);
            const float pw = sample_pos(wo, g.sw, g.pw, kk, g.dw, o[(i64)(2 * t + 1) * Vo]);
            const Sample2 s = make_sample2(ph, pw, g.H, g.W);
            low[idx * 2 + 0] = s.lo[0]; low[idx * 2 + 1] = s.lo[1];
            mask[idx] = s.mask;
        }
    }
}

}  // namespace

int transpose_cs_to_sc(const float *in, float *out, int B, int C, i64 S, cudaStream_t st)
{
    if (B <= 0 || C <= 0 || S <= 0) return DLKA_OK;
    dim3 block(32, 8), grid((unsigned)cdiv(S, 32), (unsigned)cdiv(C, 32), (unsigned)B);
    DLKA_LAUNCH("transpose_cs_sc", st, transpose_cs_sc_kernel<<<grid, block, 0, st>>>(in, out, C, S));
    return DLKA_OK;
}

int transpose_cs_to_chunk(const float *in, float *out, int B, int C, i64 S, cudaStream_t st)
{
    if (B <= 0 || C <= 0 || S <= 0) return DLKA_OK;
    if (C % 32 != 0) return DLKA_ERR_UNSUPPORTED;
    dim3 block(32, 8), grid((unsigned)cdiv(S, 32), (unsigned)(C / 32), (unsigned)B);
    DLKA_LAUNCH("transpose_cs_chunk", st, transpose_cs_chunk_kernel<<<grid, block, 0, st>>>(in, out, B, C, S));
    return DLKA_OK;
}

int transpose_sc_to_cs(const float *in, float *out, int B, int C, i64 S, cudaStream_t st)
{
    if (B <= 0 || C <= 0 || S <= 0) return DLKA_OK;
    dim3 block(32, 8), grid((unsigned)cdiv(S, 32), (unsigned)cdiv(C, 32), (unsigned)B);
    DLKA_LAUNCH("transpose_sc_cs", st, transpose_sc_cs_kernel<<<grid, block, 0, st>>>(in, out, C, S));
    return DLKA_OK;
}

int sample_indices(const float *off_cf, int32_t *low, int32_t *mask, const ConvGeo &g, cudaStream_t st)
{
    const i64 total = (i64)g.B * g.dg * g.Do * g.Ho * g.Wo * g.K;
    if (total <= 0) return DLKA_OK;
    const int blocks = (int)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
    DLKA_LAUNCH("sample_indices", st, sample_indices_kernel<<<blocks, 256, 0, st>>>(off_cf, low, mask, g));
    return DLKA_OK;
}

}  // namespace dlka
