// Minimal runtime layer: device memory, streams and kernel launch.
//
// Every kernel in this library is written as a functor whose operator() takes a BlockCtx and runs
// "phases": DR_STRIDE_LOOP / DR_THREAD_LOOP bodies separated by DR_BLOCK_SYNC(), with all
// cross-thread state in shared memory.  Compiled by nvcc (the product) a phase loop is the usual
// threadIdx-strided loop and the sync is __syncthreads().  Compiled by g++ with
// -DDR_HOST_EMULATION (tests/host only, never loaded by the dot_ring_b200 package) the same bodies
// run block after block, thread after thread, on the CPU, which lets the CPU test-suite exercise
// the exact kernel logic (indexing, phase structure, transcripts) without a GPU.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dot_ring_b200.h"
#include "fp.cuh"

#if !defined(DR_HOST_EMULATION)
#include <cuda_runtime.h>
#endif

namespace dr {

struct BlockCtx {
    uint32_t bx, by, bz;
    uint32_t gx, gy, gz;
    uint32_t nthreads;
    uint8_t* smem;
};

#if defined(__CUDA_ARCH__)
#define DR_THREAD_LOOP(t, ctx) for (uint32_t t = threadIdx.x, _dr_once = 1; _dr_once; _dr_once = 0)
#define DR_STRIDE_LOOP(i, n, ctx) for (uint32_t i = threadIdx.x; i < (uint32_t)(n); i += blockDim.x)
#define DR_BLOCK_SYNC() __syncthreads()
#else
#define DR_THREAD_LOOP(t, ctx) for (uint32_t t = 0; t < (ctx).nthreads; t++)
#define DR_STRIDE_LOOP(i, n, ctx) for (uint32_t i = 0; i < (uint32_t)(n); i++)
#define DR_BLOCK_SYNC() ((void)0)
#endif

// atomics: the emulation build runs the threads of a block one after another, so a plain update is exact there
#if defined(__CUDA_ARCH__)
DR_D uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
#else
inline uint32_t atomic_add_u32(uint32_t* p, uint32_t v) {
    uint32_t old = *p;
    *p = old + v;
    return old;
}
#endif

struct Dim3 {
    uint32_t x, y, z;
    Dim3(uint32_t x_ = 1, uint32_t y_ = 1, uint32_t z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#if !defined(DR_HOST_EMULATION)
// ------------------------------------------------------------------ CUDA build
typedef cudaStream_t Stream;

inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(DR_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define DR_CUDA(x) ::dr::cuda_check((x), #x)

template <class Body, class... Args>
__global__ void kernel_entry(Body body, Args... args) {
    extern __shared__ __align__(16) uint8_t dr_smem[];
    BlockCtx ctx{blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y, gridDim.z, blockDim.x, dr_smem};
    body(ctx, args...);
}
// Same, with an occupancy contract: the body is always launched with <= THREADS threads and wants BLOCKS CTAs per SM.
template <int THREADS, int BLOCKS, class Body, class... Args>
__global__ void __launch_bounds__(THREADS, BLOCKS) kernel_entry_lb(Body body, Args... args) {
    extern __shared__ __align__(16) uint8_t dr_smem[];
    BlockCtx ctx{blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y, gridDim.z, blockDim.x, dr_smem};
    body(ctx, args...);
}

// launch counter (bench.py reports it as gpu_launches)
inline std::atomic<uint64_t>& launch_counter() {  // several contexts (one per GPU, one thread each) launch concurrently
    static std::atomic<uint64_t> c{0};
    return c;
}

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per (kernel, device) and only ever raised: two host
// threads (two contexts) launching the same kernel with different sizes must not lower it under each other's launch.
template <class Kern>
inline void allow_dynamic_smem(Kern kern, size_t smem) {
    if (smem <= 48 * 1024) return;
    static std::atomic<size_t> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<size_t>& cur = configured[dev & 63];
    size_t seen = cur.load();
    while (smem > seen) {
        static std::mutex m;
        std::lock_guard<std::mutex> lock(m);
        seen = cur.load();
        if (smem <= seen) break;
        DR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur.store(smem);
        seen = smem;
    }
}

template <class Body, class... Args>
inline void launch(Stream s, Dim3 grid, uint32_t threads, size_t smem, Body body, Args... args) {
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    auto kern = kernel_entry<Body, Args...>;
    allow_dynamic_smem(kern, smem);
    kern<<<dim3(grid.x, grid.y, grid.z), threads, smem, s>>>(body, args...);
    DR_CUDA(cudaGetLastError());
    launch_counter()++;
}

template <int THREADS, int BLOCKS, class Body, class... Args>
inline void launch_lb(Stream s, Dim3 grid, uint32_t threads, size_t smem, Body body, Args... args) {
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    if (threads > (uint32_t)THREADS) throw Error(DR_ESTATE, "launch exceeds the kernel's launch bounds");
    auto kern = kernel_entry_lb<THREADS, BLOCKS, Body, Args...>;
    allow_dynamic_smem(kern, smem);
    kern<<<dim3(grid.x, grid.y, grid.z), threads, smem, s>>>(body, args...);
    DR_CUDA(cudaGetLastError());
    launch_counter()++;
}

// Device allocations go through a small caching layer: every C-ABI call allocates its scratch with DevBuf and frees it on
// return, and cudaMalloc / cudaFree of 100 MB-class buffers cost milliseconds each (and, measured on B200, occasional
// stalls of hundreds of ms).  Freed blocks are kept per device (up to DR_DEV_CACHE_CAP bytes) and reused for requests of a
// similar size; on an out-of-memory error the cache is emptied and the allocation retried.
// A block may be freed while kernels that use it are still queued (a scratch buffer that grows in the middle of a call, a
// temporary that goes out of scope before the call's final synchronisation).  Reuse on the SAME stream is ordered behind those
// kernels; another context on the same device (a second stream, possibly another host thread) must not get the block before
// they have run.  So every cached block remembers the stream it was freed on and an event recorded there at that moment:
// the freeing stream may take it back at once, any other stream only after the event has completed.
inline Stream& current_stream() {  // set by Ctx::activate(): the stream the calling thread's API call works on
    static thread_local Stream s = nullptr;
    return s;
}
struct CachedBlock {
    void* p;
    Stream owner;
    cudaEvent_t freed;  // null: nothing was pending
};
struct DevCache {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, CachedBlock> free_blocks;  // (device, size) -> block
    std::map<void*, std::pair<int, size_t>> live;                     // block -> (device, size)
    std::vector<cudaEvent_t> spare_events;
    size_t cached = 0;
    size_t cap = (size_t)24 << 30;
};
inline DevCache& dev_cache() {
    static DevCache c;
    return c;
}
inline void dev_cache_trim() {
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lock(c.m);
    for (auto& kv : c.free_blocks) {
        if (kv.second.freed) {
            cudaEventSynchronize(kv.second.freed);
            c.spare_events.push_back(kv.second.freed);
        }
        cudaFree(kv.second.p);
    }
    c.free_blocks.clear();
    c.cached = 0;
}
inline void* dev_alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    // size classes: 4 KiB granules below 1 MiB, 1 MiB granules above
    const size_t gran = bytes < ((size_t)1 << 20) ? 4096 : ((size_t)1 << 20);
    const size_t size = (bytes + gran - 1) / gran * gran;
    int dev = 0;
    cudaGetDevice(&dev);
    DevCache& c = dev_cache();
    {
        std::lock_guard<std::mutex> lock(c.m);
        for (auto it = c.free_blocks.lower_bound({dev, size}); it != c.free_blocks.end() && it->first.first == dev && it->first.second <= size + size / 4; ++it) {
            CachedBlock& b = it->second;
            const bool usable = b.owner == current_stream() || !b.freed || cudaEventQuery(b.freed) == cudaSuccess;
            if (!usable) {
                cudaGetLastError();  // cudaErrorNotReady is not an error
                continue;
            }
            void* p = b.p;
            if (b.freed) c.spare_events.push_back(b.freed);
            c.live[p] = it->first;
            c.cached -= it->first.second;
            c.free_blocks.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, size);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dev_cache_trim();
        e = cudaMalloc(&p, size);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();  // the failed allocation must not surface again at the next launch check
        throw Error(DR_ENOMEM, std::string("cudaMalloc(") + std::to_string(size) + "): " + cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lock(c.m);
    c.live[p] = {dev, size};
    return p;
}
inline void dev_free(void* p) {
    if (!p) return;
    // A buffer released while an exception unwinds the call may still be read by kernels already queued on the stream (or on
    // the context's side stream): drain the device before the block can be handed out again.
    if (std::uncaught_exceptions() > 0) cudaDeviceSynchronize();
    DevCache& c = dev_cache();
    std::unique_lock<std::mutex> lock(c.m);
    auto it = c.live.find(p);
    if (it == c.live.end()) {
        lock.unlock();
        cudaFree(p);
        return;
    }
    auto key = it->second;
    c.live.erase(it);
    // very large blocks (window tables) go back to the driver; the rest is kept for the next call
    if (key.second > ((size_t)4 << 30) || c.cached + key.second > c.cap) {
        lock.unlock();
        cudaFree(p);  // synchronises with the device's outstanding work by itself
        return;
    }
    CachedBlock b{p, current_stream(), nullptr};
    if (!c.spare_events.empty()) {
        b.freed = c.spare_events.back();
        c.spare_events.pop_back();
    } else if (cudaEventCreateWithFlags(&b.freed, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        b.freed = nullptr;
    }
    if (b.freed) {
        if (cudaEventRecord(b.freed, b.owner) != cudaSuccess) {  // e.g. the stream is already destroyed (context teardown): fall back to a full drain
            cudaGetLastError();
            cudaDeviceSynchronize();
            c.spare_events.push_back(b.freed);
            b.freed = nullptr;
        }
    } else {
        cudaDeviceSynchronize();
    }
    c.free_blocks.insert({key, b});
    c.cached += key.second;
}
inline void h2d(Stream s, void* dst, const void* src, size_t bytes) {
    if (bytes) DR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
}
inline void d2h(Stream s, void* dst, const void* src, size_t bytes) {
    if (bytes) DR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
}
inline void d2d(Stream s, void* dst, const void* src, size_t bytes) {
    if (bytes) DR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
}
inline void dev_zero(Stream s, void* dst, size_t bytes) {
    if (bytes) DR_CUDA(cudaMemsetAsync(dst, 0, bytes, s));
}
inline void stream_sync(Stream s) { DR_CUDA(cudaStreamSynchronize(s)); }
inline void* host_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    DR_CUDA(cudaMallocHost(&p, bytes ? bytes : 16));
    return p;
}
inline void host_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}
#else
// ------------------------------------------------------------------ host emulation (tests only)
typedef int Stream;
inline std::atomic<uint64_t>& launch_counter() {  // several contexts (one per GPU, one thread each) launch concurrently
    static std::atomic<uint64_t> c{0};
    return c;
}
template <class Body, class... Args>
inline void launch(Stream, Dim3 grid, uint32_t threads, size_t smem, Body body, Args... args) {
    std::vector<uint8_t> sm(smem + 16);
    for (uint32_t z = 0; z < grid.z; z++)
        for (uint32_t y = 0; y < grid.y; y++)
            for (uint32_t x = 0; x < grid.x; x++) {
                BlockCtx ctx{x, y, z, grid.x, grid.y, grid.z, threads, sm.data()};
                body(ctx, args...);
            }
    launch_counter()++;
}
template <int THREADS, int BLOCKS, class Body, class... Args>
inline void launch_lb(Stream s, Dim3 grid, uint32_t threads, size_t smem, Body body, Args... args) {
    launch(s, grid, threads, smem, body, args...);
}
inline void* dev_alloc(size_t bytes) {
    void* p = calloc(bytes ? bytes : 16, 1);
    if (!p) throw Error(DR_ENOMEM, "calloc failed");
    return p;
}
inline void dev_free(void* p) { free(p); }
inline void h2d(Stream, void* dst, const void* src, size_t bytes) {
    if (bytes) memcpy(dst, src, bytes);
}
inline void d2h(Stream, void* dst, const void* src, size_t bytes) {
    if (bytes) memcpy(dst, src, bytes);
}
inline void d2d(Stream, void* dst, const void* src, size_t bytes) {
    if (bytes) memmove(dst, src, bytes);
}
inline void dev_zero(Stream, void* dst, size_t bytes) { memset(dst, 0, bytes); }
inline void stream_sync(Stream) {}
inline void* host_alloc_pinned(size_t bytes) { return malloc(bytes ? bytes : 16); }
inline void host_free_pinned(void* p) { free(p); }
#endif

// RAII device buffer
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        p = (T*)dev_alloc(count * sizeof(T));
        n = count;
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    void release() {
        dev_free(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace dr
