// host_pipe.cu -- host-buffer plumbing for the 3D attention block:
//   * a streaming pipeline context that keeps `depth` steps in flight so that the H2D copy of step k+1 and the D2H copy of
//     step k-1 overlap the compute of step k (per sample inside a step as well).  Samples / steps are independent
//     (SURVEY.md 8e), so this is plain stream plumbing: three streams + events, no extra kernels; the compute path is
//     dlka_lka_attention3d_deform_forward;
//   * NUMA-placed pinned host buffers (dlka_host_alloc / dlka_host_free) and dlka_host_bind_thread: at 805 MB in and 805 MB
//     out per step the end-to-end rate is the PCIe + host-DRAM rate, and on the 8-GPU box (GPU0-3 on socket 0, GPU4-7 on
//     socket 1) un-placed pinned pages made every rank's DMA cross the socket link (SCALE_r01: e2e efficiency 0.47 at 8).
#include <ctype.h>
#include <errno.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"

struct dlkaHostPipe {
    int depth = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    unsigned long long step = 0;        // steps submitted
    unsigned long long completed = 0;   // lower bound of the steps whose last D2H copy has finished (advanced by polling)
    // per slot: "sample b landed" / "sample b computed" events, "slot inputs consumed", "slot outputs copied back"
    std::vector<std::vector<cudaEvent_t>> ev_in, ev_comp;
    std::vector<cudaEvent_t> ev_consumed, ev_out;
    std::vector<char> used;
    cudaEvent_t ev_start = nullptr;
};

namespace {

int ensure_events(std::vector<cudaEvent_t> &v, size_t n)
{
    while (v.size() < n) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
            return dlka::record_cuda_error(cudaGetLastError(), "cudaEventCreateWithFlags");
        v.push_back(e);
    }
    return DLKA_OK;
}

void pipe_release(dlkaHostPipe *p)
{
    if (!p) return;
    if (p->s_in) cudaStreamSynchronize(p->s_in);
    if (p->s_out) cudaStreamSynchronize(p->s_out);
    for (auto &v : p->ev_in) for (auto e : v) cudaEventDestroy(e);
    for (auto &v : p->ev_comp) for (auto e : v) cudaEventDestroy(e);
    for (auto e : p->ev_consumed) if (e) cudaEventDestroy(e);
    for (auto e : p->ev_out) if (e) cudaEventDestroy(e);
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

// ---------------------------------------------------------------------------------------------------- NUMA placement
#ifndef MPOL_PREFERRED
#define MPOL_DEFAULT 0
#define MPOL_PREFERRED 1
#define MPOL_BIND 2
#define MPOL_INTERLEAVE 3
#endif

bool read_text(const std::string &path, std::string &out)
{
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096];
    const size_t n = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    buf[n] = 0;
    out = buf;
    return true;
}

int numa_node_count()
{
    int n = 0;
    std::string s;
    while (n < 64 && read_text("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist", s)) ++n;
    return n;
}

// NUMA node of a CUDA device from sysfs (the PCI function's numa_node), -1 when unknown / single-node
int device_numa_node(int device)
{
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    std::string s;
    if (!read_text(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node", s)) return -1;
    const int node = atoi(s.c_str());
    return node >= 0 && node < numa_node_count() ? node : -1;
}

// "0-31,64-95" -> cpu_set_t
bool parse_cpulist(const std::string &s, cpu_set_t *set)
{
    CPU_ZERO(set);
    const char *p = s.c_str();
    bool any = false;
    while (*p) {
        while (*p == ',' || *p == ' ' || *p == '\n') ++p;
        if (!*p) break;
        char *end;
        long a = strtol(p, &end, 10), b = a;
        if (end == p) return false;
        p = end;
        if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); any = true; }
    }
    return any;
}

std::mutex g_alloc_mu;
std::map<void *, size_t> g_allocs;   // dlka_host_alloc'ed blocks: mmap length by pointer

}  // namespace

extern "C" {

int dlka_host_numa_node(int device)
{
    return device_numa_node(device);
}

// Pin the CALLING host thread to the cores of `device`'s NUMA node and prefer that node for its future page
// allocations.  Returns the node, or -1 if the topology is unknown (nothing changed).
int dlka_host_bind_thread(int device)
{
    const int node = device_numa_node(device);
    if (node < 0) return -1;
    std::string cpus;
    cpu_set_t set;
    if (read_text("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", cpus) && parse_cpulist(cpus, &set))
        sched_setaffinity(0, sizeof(set), &set);   // best effort: a cgroup cpuset may refuse
    unsigned long mask[1] = {1ul << node};
    syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, 64ul);
    return node;
}

// Page-locked host buffer placed for `device`:  policy 0 = no placement (kernel default), 1 = on the device's NUMA node
// (default), 2 = interleaved over all nodes (when one socket's DRAM cannot feed its GPUs).  The block is mmap'ed, bound
// with mbind BEFORE first touch, touched, and registered with cudaHostRegister (portable).  Free with dlka_host_free.
int dlka_host_alloc(void **ptr, size_t bytes, int device, int policy)
{
    // policy | 16: WRITE-COMBINED input buffer (cudaHostAllocWriteCombined: the DMA reads of the H2D copy do not snoop the CPU
    // caches; the host must only WRITE such a buffer -- CPU reads from it are uncached and very slow)
    const bool wc = (policy & 16) != 0;
    policy &= 15;
    if (!ptr || bytes == 0 || policy < 0 || policy > 2) return DLKA_ERR_INVALID_ARGUMENT;
    *ptr = nullptr;
    if (wc) {
        // the driver allocates the pages from the calling thread's memory policy: bind it for the duration of the call
        const int node = device_numa_node(device), nodes = numa_node_count();
        unsigned long mask[1] = {policy == 1 && node >= 0 ? 1ul << node : (nodes >= 64 ? ~0ul : (1ul << (nodes > 0 ? nodes : 1)) - 1)};
        if (policy == 1 && node >= 0) syscall(SYS_set_mempolicy, MPOL_BIND, mask, 64ul);
        else if (policy == 2 && nodes > 1) syscall(SYS_set_mempolicy, MPOL_INTERLEAVE, mask, 64ul);
        void *q = nullptr;
        const cudaError_t e = cudaHostAlloc(&q, bytes, cudaHostAllocWriteCombined | cudaHostAllocPortable);
        if (policy == 1 && node >= 0) { unsigned long pm[1] = {1ul << node}; syscall(SYS_set_mempolicy, MPOL_PREFERRED, pm, 64ul); }
        else syscall(SYS_set_mempolicy, MPOL_DEFAULT, nullptr, 0ul);
        if (e != cudaSuccess) return dlka::record_cuda_error(e, "cudaHostAlloc");
        {
            std::lock_guard<std::mutex> lk(g_alloc_mu);
            g_allocs[q] = 0;   // length 0 marks a cudaHostAlloc block
        }
        *ptr = q;
        return DLKA_OK;
    }
    const size_t len = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        snprintf(dlka::g_last_cuda_error, sizeof(dlka::g_last_cuda_error), "mmap(%zu): %s", len, strerror(errno));
        return DLKA_ERR_CUDA;
    }
    const int nodes = numa_node_count(), node = device_numa_node(device);
    if (policy == 1 && node >= 0) {
        unsigned long mask[1] = {1ul << node};
        syscall(SYS_mbind, p, len, MPOL_PREFERRED, mask, 64ul, 0u);
    } else if (policy == 2 && nodes > 1) {
        unsigned long mask[1] = {nodes >= 64 ? ~0ul : (1ul << nodes) - 1};
        syscall(SYS_mbind, p, len, MPOL_INTERLEAVE, mask, 64ul, 0u);
    }
    madvise(p, len, MADV_HUGEPAGE);
    memset(p, 0, len);   // first touch under the policy above
    const cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        munmap(p, len);
        return dlka::record_cuda_error(e, "cudaHostRegister");
    }
    {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        g_allocs[p] = len;
    }
    *ptr = p;
    return DLKA_OK;
}

int dlka_host_free(void *ptr)
{
    if (!ptr) return DLKA_OK;
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        auto it = g_allocs.find(ptr);
        if (it == g_allocs.end()) return DLKA_ERR_INVALID_ARGUMENT;
        len = it->second;
        g_allocs.erase(it);
    }
    if (len == 0) {   // cudaHostAlloc (write-combined) block
        const cudaError_t e = cudaFreeHost(ptr);
        return e == cudaSuccess ? DLKA_OK : dlka::record_cuda_error(e, "cudaFreeHost");
    }
    const cudaError_t e = cudaHostUnregister(ptr);
    munmap(ptr, len);
    return e == cudaSuccess ? DLKA_OK : dlka::record_cuda_error(e, "cudaHostUnregister");
}

int dlka_host_pipe_create(dlkaHostPipe **pipe, int depth)
{
    if (!pipe || depth < 1 || depth > 8) return DLKA_ERR_INVALID_ARGUMENT;
    *pipe = nullptr;
    dlkaHostPipe *p = new dlkaHostPipe();
    p->depth = depth;
    p->ev_in.resize(depth);
    p->ev_comp.resize(depth);
    p->ev_consumed.assign(depth, nullptr);
    p->ev_out.assign(depth, nullptr);
    p->used.assign(depth, 0);
    bool ok = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int s = 0; ok && s < depth; ++s)
        ok = cudaEventCreateWithFlags(&p->ev_out[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&p->ev_consumed[s], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        const int rc = dlka::record_cuda_error(cudaGetLastError(), "dlka_host_pipe_create");
        pipe_release(p);   // streams / events created so far
        return rc;
    }
    *pipe = p;
    return DLKA_OK;
}

int dlka_host_pipe_wait(dlkaHostPipe *p)
{
    if (!p) return DLKA_ERR_INVALID_ARGUMENT;
    DLKA_CUDA_TRY(cudaStreamSynchronize(p->s_in));
    DLKA_CUDA_TRY(cudaStreamSynchronize(p->s_out));
    p->completed = p->step;
    return DLKA_OK;
}

int dlka_host_pipe_destroy(dlkaHostPipe *p)
{
    pipe_release(p);
    return DLKA_OK;
}

// Enqueue one step and return without host synchronisation.  dev_scratch holds depth * 2 * B*N*C floats (a caller that
// changes B*N*C between steps must size it for the largest step and must not move it while steps are in flight).
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
        // slot reuse: the previous occupant's LAST compute (whatever its batch size was) must have consumed xd[s], and its
        // outputs must have left yd[s]
        DLKA_CUDA_TRY(cudaStreamWaitEvent(p->s_in, p->ev_consumed[s], 0));
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
    DLKA_CUDA_TRY(cudaEventRecord(p->ev_consumed[s], st));
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

// Non-blocking: a lower bound of the number of submitted steps whose last D2H copy has finished (the host tensors of those steps
// may be released).  Steps finish in submission order (one output stream); a slot's event is re-recorded when the slot is reused,
// so the count only advances over slots whose CURRENT event has completed -- conservative, never ahead of the truth.
long long dlka_host_pipe_completed(dlkaHostPipe *p)
{
    if (!p) return DLKA_ERR_INVALID_ARGUMENT;
    while (p->completed < p->step) {
        const int s = (int)(p->completed % p->depth);
        const cudaError_t e = cudaEventQuery(p->ev_out[s]);
        if (e == cudaErrorNotReady) { cudaGetLastError(); break; }
        if (e != cudaSuccess) return dlka::record_cuda_error(e, "cudaEventQuery");
        // ev_out[s] now belongs to the newest step that used slot s: that step and every earlier one are done
        unsigned long long newest = p->step - 1 - ((p->step - 1 - p->completed) % p->depth);   // newest step index with slot s
        p->completed = newest + 1;
    }
    return (long long)p->completed;
}

}  // extern "C"
