// mempool.cu — process-wide caching of the library's device buffers, page-locked host buffers and streams.
//
// Why: the reference's own benchmark suite (benchmark/bench_spread.jl, bench_grad.jl, bench_maxloc.jl,
// bench_disentangle.jl) calls the one-shot API — omega(bvectors, M, U), omega_grad(...), max_localize(model) — on
// 12-band, 64-k-point problems whose arithmetic is tens of microseconds on a B200.  Each such call creates a context,
// evaluates and destroys it; measured at that size (profiles/r4_summary.md) the context cost 6.5 ms, of which
// cudaGetDeviceProperties 2.4 ms, cudaMallocHost 1-2.4 ms, 21 cudaMalloc 1.0 ms, 30 cudaFree + cudaFreeHost 1.5 ms —
// against 0.06 ms for the evaluation itself and 0.44 ms for the CPU path.  The driver calls are the cost, not the
// memory: so freed blocks are kept (per device, by size) and handed out again, like every long-running CUDA
// application does (torch's caching allocator, cudaMallocAsync pools); blocks stay whole cudaMalloc allocations, so
// peer access and CUDA IPC see exactly what they saw before.
//
// wb_dev_free keeps cudaFree's semantics: it drains the block's device before the block can be handed out again.
// wb_dev_free_drained is for callers that have drained the device themselves (wb200_destroy: once for ~30 blocks).
//
//   WB200_POOL=0            every call goes straight to the driver (cudaMalloc / cudaFree)
//   WB200_POOL_MAX_BLOCK_MB blocks above this size are never kept (default 64)
//   WB200_POOL_MAX_MB       bytes kept per device (default 1024)
//   WB200_POOL_POISON=1     a recycled block is filled with 0xFF (NaN) before it is handed out: reads of memory
//                           the library never wrote show up in the parity tests instead of depending on what a
//                           fresh cudaMalloc happens to contain
#include "common.cuh"

#include <array>
#include <map>
#include <mutex>
#include <unordered_map>

namespace {

struct Pool {
    std::mutex mu;
    bool enabled = true, poison = false;
    size_t max_block = (size_t)64 << 20, max_total = (size_t)1024 << 20;
    struct Live { size_t bytes; int device; };
    std::unordered_map<void*, Live> live;                       // blocks handed out (device and pinned)
    std::map<int, std::multimap<size_t, void*>> free_dev;       // device -> size -> block
    std::map<int, size_t> cached_dev;                           // bytes kept per device
    std::multimap<size_t, void*> free_pin;
    size_t cached_pin = 0;
    std::map<int, std::vector<cudaStream_t>> streams;           // idle non-blocking streams per device
    long long hits = 0, misses = 0;
    Pool() {
        const char* e = getenv("WB200_POOL");
        enabled = !(e && e[0] == '0');
        e = getenv("WB200_POOL_POISON");
        poison = e && e[0] == '1';
        if ((e = getenv("WB200_POOL_MAX_BLOCK_MB"))) max_block = (size_t)atoll(e) << 20;
        if ((e = getenv("WB200_POOL_MAX_MB"))) max_total = (size_t)atoll(e) << 20;
    }
};
Pool& pool() {
    static Pool* p = new Pool();   // never destroyed: contexts may be closed from atexit handlers / finalisers
    return *p;
}
inline size_t round_up(size_t b) { return b == 0 ? 512 : (b + 511) & ~(size_t)511; }

// a kept block serves a request when it wastes at most 1/8 of itself (+ one page)
void* take(std::multimap<size_t, void*>& fl, size_t want, size_t* got) {
    auto it = fl.lower_bound(want);
    if (it == fl.end() || it->first > want + want / 8 + 4096) return nullptr;
    void* p = it->second;
    *got = it->first;
    fl.erase(it);
    return p;
}

void trim_device(Pool& P, int device) {
    auto& fl = P.free_dev[device];
    for (auto& kv : fl) cudaFree(kv.second);
    fl.clear();
    P.cached_dev[device] = 0;
}

}  // namespace

cudaError_t wb_dev_alloc_raw(void** out, size_t bytes) {
    Pool& P = pool();
    *out = nullptr;
    if (!P.enabled) return cudaMalloc(out, bytes);
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    const size_t want = round_up(bytes);
    {
        std::lock_guard<std::mutex> lk(P.mu);
        size_t got = 0;
        if (void* p = take(P.free_dev[device], want, &got)) {
            P.cached_dev[device] -= got;
            P.live[p] = {got, device};
            P.hits++;
            *out = p;
            if (P.poison) {   // (cudaMemset runs on the legacy stream, which non-blocking streams do not wait for)
                const cudaError_t pe = cudaMemset(p, 0xFF, got);
                return pe != cudaSuccess ? pe : cudaDeviceSynchronize();
            }
            return cudaSuccess;
        }
        P.misses++;
    }
    e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {   // give the kept blocks back to the driver and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(P.mu);
            trim_device(P, device);
        }
        e = cudaMalloc(out, want);
    }
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[*out] = {want, device};
    return cudaSuccess;
}

void wb_dev_free(void* p) {
    if (!p) return;
    Pool& P = pool();
    if (P.enabled) {
        int owner = -1, cur = 0;
        {
            std::lock_guard<std::mutex> lk(P.mu);
            auto it = P.live.find(p);
            if (it != P.live.end()) owner = it->second.device;
        }
        if (owner >= 0 && cudaGetDevice(&cur) == cudaSuccess) {
            if (cur != owner) cudaSetDevice(owner);
            cudaDeviceSynchronize();
            if (cur != owner) cudaSetDevice(cur);
        }
    }
    wb_dev_free_drained(p);
}

void wb_dev_free_drained(void* p) {
    if (!p) return;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it != P.live.end()) {
            const Pool::Live b = it->second;
            P.live.erase(it);
            if (P.enabled && b.bytes <= P.max_block && P.cached_dev[b.device] + b.bytes <= P.max_total) {
                P.free_dev[b.device].emplace(b.bytes, p);
                P.cached_dev[b.device] += b.bytes;
                return;
            }
        }
    }
    cudaFree(p);   // too large to keep, pool full or disabled, or a block this allocator never saw
}

cudaError_t wb_pinned_alloc_raw(void** out, size_t bytes) {
    Pool& P = pool();
    *out = nullptr;
    if (!P.enabled) return cudaMallocHost(out, bytes);
    const size_t want = round_up(bytes);
    {
        std::lock_guard<std::mutex> lk(P.mu);
        size_t got = 0;
        if (void* p = take(P.free_pin, want, &got)) {
            P.cached_pin -= got;
            P.live[p] = {got, -1};
            *out = p;
            if (P.poison) memset(p, 0xFF, got);
            return cudaSuccess;
        }
    }
    const cudaError_t e = cudaMallocHost(out, want);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[*out] = {want, -1};
    return cudaSuccess;
}

void wb_pinned_free(void* p) {
    if (!p) return;
    Pool& P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it != P.live.end()) {
            const size_t bytes = it->second.bytes;
            P.live.erase(it);
            if (P.enabled && bytes <= ((size_t)4 << 20) && P.cached_pin + bytes <= ((size_t)64 << 20)) {
                P.free_pin.emplace(bytes, p);
                P.cached_pin += bytes;
                return;
            }
        }
    }
    cudaFreeHost(p);
}

// non-blocking streams: creation and destruction are driver round trips too (0.07 + 0.06 ms measured)
cudaError_t wb_stream_acquire(cudaStream_t* s) {
    Pool& P = pool();
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    if (P.enabled) {
        std::lock_guard<std::mutex> lk(P.mu);
        auto& v = P.streams[device];
        if (!v.empty()) { *s = v.back(); v.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}
// the stream must be idle (the caller has synchronised it) and belong to the current device
void wb_stream_release(cudaStream_t s) {
    if (!s) return;
    Pool& P = pool();
    int device = 0;
    if (P.enabled && cudaGetDevice(&device) == cudaSuccess) {
        std::lock_guard<std::mutex> lk(P.mu);
        auto& v = P.streams[device];
        if (v.size() < 16) { v.push_back(s); return; }
    }
    cudaStreamDestroy(s);
}

// compute capability and SM count without cudaGetDeviceProperties (2.4 ms per call on this driver)
int wb_device_info(int device, int* major, int* minor, int* num_sms) {
    static std::mutex mu;
    static std::map<int, std::array<int, 3>> seen;
    std::lock_guard<std::mutex> lk(mu);
    auto it = seen.find(device);
    if (it == seen.end()) {
        std::array<int, 3> v{0, 0, 0};
        if (cudaDeviceGetAttribute(&v[0], cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
            cudaDeviceGetAttribute(&v[1], cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess ||
            cudaDeviceGetAttribute(&v[2], cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
            cudaGetLastError();
            return WB200_ERR_CUDA;
        }
        it = seen.emplace(device, v).first;
    }
    *major = it->second[0]; *minor = it->second[1]; *num_sms = it->second[2];
    return WB200_OK;
}

// give every kept block back to the driver (e.g. before another library needs the memory)
extern "C" void wb200_trim_pools(void) {
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& d : P.free_dev) {
        cudaSetDevice(d.first);
        trim_device(P, d.first);
    }
    for (auto& kv : P.free_pin) cudaFreeHost(kv.second);
    P.free_pin.clear();
    P.cached_pin = 0;
    for (auto& d : P.streams) {
        cudaSetDevice(d.first);
        for (cudaStream_t s : d.second) cudaStreamDestroy(s);
        d.second.clear();
    }
    cudaSetDevice(cur);
}

// bytes kept on `device` (< 0: page-locked host bytes), and the hit / miss counters of the device allocator
extern "C" long long wb200_pool_stats(int device, long long* hits, long long* misses) {
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    if (hits) *hits = P.hits;
    if (misses) *misses = P.misses;
    if (device < 0) return (long long)P.cached_pin;
    auto it = P.cached_dev.find(device);
    return it == P.cached_dev.end() ? 0 : (long long)it->second;
}
