// host_pipe.cu -- streaming host-buffer entry for the 3D attention block: a small pipeline context that keeps
// `depth` steps in flight so that the H2D copy of step k+1 and the D2H copy of step k-1 overlap the compute of
// step k (per sample inside a step as well).  Samples / steps are independent, so this is plain
// stream plumbing: three streams + events, no extra kernels, the compute path is dlka_lka_attention3d_deform_forward.
#include <vector>

#include "kernels.cuh"

struct dlkaHostPipe {
    int depth;
    cudaStream_t s_in, s_out;
    unsigned long long step;
    // per slot: events for "inputs landed" / "compute done" per sample, and "all outputs copied back"
    std::vector<std::vector<cudaEvent_t>> ev_in, ev_comp;
    std::vector<cudaEvent_t> ev_out;
    std::vector<char> used;
    cudaEvent_t ev_start;
};

namespace {
int ensure_events(std::vector<cudaEvent_t> &v, size_t n)
{
    while (v.size() < n) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return DLKA_ERR_CUDA;
        v.push_back(e);
    }
    return DLKA_OK;
}
}  // namespace

extern "C" {

int dlka_host_pipe_create(dlkaHostPipe **pipe, int depth)
{
    if (!pipe || depth < 1 || depth > 8) return DLKA_ERR_INVALID_ARGUMENT;
    dlkaHostPipe *p = new dlkaHostPipe();
    p->depth = depth;
    p->step = 0;
    if (cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming) != cudaSuccess) {
        delete p;
        return dlka::record_cuda_error(cudaGetLastError(), "dlka_host_pipe_create");
    }
    p->ev_in.resize(depth);
    p->ev_comp.resize(depth);
    p->ev_out.resize(depth);
    p->used.assign(depth, 0);
    for (int s = 0; s < depth; ++s)
        if (cudaEventCreateWithFlags(&p->ev_out[s], cudaEventDisableTiming) != cudaSuccess) return DLKA_ERR_CUDA;
    *pipe = p;
    return DLKA_OK;
}

int dlka_host_pipe_wait(dlkaHostPipe *p)
{
    if (!p) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_CUDA_TRY(cudaStreamSynchronize(p->s_in));
    DLKA_CUDA_TRY(cudaStreamSynchronize(p->s_out));
    return DLKA_OK;
}

int dlka_host_pipe_destroy(dlkaHostPipe *p)
{
    if (!p) return DLKA_OK;
    cudaStreamSynchronize(p->s_in);
    cudaStreamSynchronize(p->s_out);
    for (auto &v : p->ev_in) for (auto e : v) cudaEventDestroy(e);
    for (auto &v : p->ev_comp) for (auto e : v) cudaEventDestroy(e);
    for (auto e : p->ev_out) cudaEventDestroy(e);
    cudaEventDestroy(p->ev_start);
    cudaStreamDestroy(p->s_in);
    cudaStreamDestroy(p->s_out);
    delete p;
    return DLKA_OK;
}

// Enqueue one step and return without host synchronisation.  dev_scratch holds depth * 2 * B*N*C floats.
// After the call, `stream` is ordered after this step's last D2H copy (so the caller can record an event on it).
int dlka_lka_attention3d_deform_forward_host_async(dlkaHostPipe *p, const dlkaBlock3dParams *params, const float *x_host,
                                                   float *y_host, int B, int C, int D1, int D2, int D3, int math,
                                                   void *dev_scratch, size_t dev_scratch_bytes, void *workspace,
                                                   size_t workspace_bytes, void *stream)
{
    if (!p || !x_host || !y_host || !dev_scratch) return DLKA_ERR_INVALID_ARGUMENT;
    if (B <= 0 || C <= 0 || D1 <= 0 || D2 <= 0 || D3 <= 0) return DLKA_ERR_INVALID_ARGUMENT;
    const size_t n = (size_t)B * D1 * D2 * D3 * C, n1 = n / B;
    if (dev_scratch_bytes < (size_t)p->depth * 2 * n * sizeof(float)) return DLKA_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int s = (int)(p->step % p->depth);
    float *xd = (float *)dev_scratch + (size_t)s * 2 * n, *yd = xd + n;
    DLKA_TRY(ensure_events(p->ev_in[s], B));
    DLKA_TRY(ensure_events(p->ev_comp[s], B));
    if (p->step == 0) {  // order the helper streams after whatever the caller queued before the first step
        DLKA_CUDA_TRY(cudaEventRecord(p->ev_start, st));
        DLKA_CUDA_TRY(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
        DLKA_CUDA_TRY(cudaStreamWaitEvent(p->s_out, p->ev_start, 0));
    }
    if (p->used[s]) {
        // slot reuse: the previous occupant's compute must have consumed xd[s]; its outputs must have left yd[s]
        DLKA_CUDA_TRY(cudaStreamWaitEvent(p->s_in, p->ev_comp[s][B - 1], 0));
        DLKA_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_out[s], 0));
    }
    for (int b = 0; b < B; ++b) {
        DLKA_CUDA_TRY(cudaMemcpyAsync(xd + b * n1, x_host + b * n1, n1 * sizeof(float), cudaMemcpyHostToDevice, p->s_in));
        DLKA_CUDA_TRY(cudaEventRecord(p->ev_in[s][b], p->s_in));
    }
    for (int b = 0; b < B; ++b) {
        DLKA_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_in[s][b], 0));
        DLKA_TRY(dlka_lka_attention3d_deform_forward(params, xd + b * n1, yd + b * n1, 1, C, D1, D2, D3, math, workspace,
                                                     workspace_bytes, stream));
        DLKA_CUDA_TRY(cudaEventRecord(p->ev_comp[s][b], st));
        DLKA_CUDA_TRY(cudaStreamWaitEvent(p->s_out, p->ev_comp[s][b], 0));
        DLKA_CUDA_TRY(cudaMemcpyAsync(y_host + b * n1, yd + b * n1, n1 * sizeof(float), cudaMemcpyDeviceToHost, p->s_out));
    }
    DLKA_CUDA_TRY(cudaEventRecord(p->ev_out[s], p->s_out));
    p->used[s] = 1;
    p->step++;
    return DLKA_OK;
}

// Make `stream` wait for every D2H copy enqueued so far (so an event recorded on it afterwards marks "results on host").
int dlka_host_pipe_join(dlkaHostPipe *p, void *stream)
{
    if (!p) return DLKA_ERR_INVALID_ARGUMENT;
    for (int s = 0; s < p->depth; ++s)
        if (p->used[s]) DLKA_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, p->ev_out[s], 0));
    return DLKA_OK;
}

}  // extern "C"
