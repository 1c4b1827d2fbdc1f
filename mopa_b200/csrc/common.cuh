// Shared helpers for libmopa_scn (sm_100a). No torch, no third-party headers: CUDA runtime only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <string>
#include <utility>

namespace mopa {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(const char *file, int line, const std::string &msg) {
    g_last_error = std::string(file) + ":" + std::to_string(line) + ": " + msg;
    return 1;
}

#define MOPA_FAIL(msg) return ::mopa::fail(__FILE__, __LINE__, (msg))
#define MOPA_CHECK(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) return ::mopa::fail(__FILE__, __LINE__, (msg)); \
    } while (0)
#define MOPA_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return ::mopa::fail(__FILE__, __LINE__, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define MOPA_TRY(expr)            \
    do {                          \
        int r__ = (expr);         \
        if (r__ != 0) return r__; \
    } while (0)
// after every kernel launch: count it and surface launch-configuration errors
#define MOPA_LAUNCHED()                                                                   \
    do {                                                                                  \
        ::mopa::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess)                                                           \
            return ::mopa::fail(__FILE__, __LINE__, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
    } while (0)

// ---- programmatic dependent launch. The network is a chain of ~180 dependent kernels per step on one stream; between two
// plain launches the GPU drains, then starts the next grid (a few microseconds each). A kernel launched through
// launch_pdl may begin while its predecessor in the stream is still running: its blocks run their prologue (barrier
// initialisation, TMEM allocation, loads of data that is older than the predecessor), then pdl_wait() blocks until the
// predecessor grid has completed and its writes are visible. pdl_trigger() lets the NEXT kernel in the stream start
// launching. Rules kept by every kernel that uses them: pdl_wait() before the first access to anything the predecessor
// may write, pdl_trigger() after it (so at most one dependent grid is ever waiting), both executed by all threads.
// Launched without the attribute (the default; MOPA_SCN_PDL=1 sets it) both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("MOPA_SCN_PDL");  // off by default: measured no gain (profiles/r02_overlap.txt): the gaps of the
        return e && e[0] == '1';                  // main stream are already filled by the d_weight and geometry streams
    }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

constexpr int kNumSMs = 148;  // B200; used where a work plan must be reproducible without a device (d_weight items)

// SM count of the CURRENT device (cached per device index); falls back to kNumSMs if the query fails
inline int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMs;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// Function attributes (cudaFuncSetAttribute) are per DEVICE: `fn` runs once per device index, under a lock, so a second
// GPU used by the same process (Metadata_new(device) / Program_new(..., device)) gets configured too.
template <class F>
inline int once_per_device(std::atomic<uint64_t> &done, F &&fn) {
    int dev = 0;
    MOPA_CUDA(cudaGetDevice(&dev));
    MOPA_CHECK(dev >= 0 && dev < 64, "device index out of range");
    const uint64_t bit = 1ull << dev;
    if (done.load(std::memory_order_acquire) & bit) return 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (done.load(std::memory_order_acquire) & bit) return 0;
    MOPA_TRY(fn());
    done.fetch_or(bit, std::memory_order_release);
    return 0;
}

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events on the launching stream around the main kernel
// of each op; off by default and free when off. tag = 10 * class + op, class 1 conv forward, 2 conv d_input,
// 3 conv d_weight, 4 BatchNorm forward, 5 BatchNorm backward, 6 input/output layer; op 1 subm, 2 conv, 3 deconv, 0 n/a.
struct Gather;
int prof_begin(int tag, const Gather *gt, int c_in, int c_out, int64_t rows, cudaStream_t s);  // -1 when profiling is off
void prof_end(int idx, cudaStream_t s);

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- site keys: 16 bits per axis, batch in the top 16 (valid for spatial_size <= 65536, batch < 65535) ----
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ inline uint64_t pack_key(uint32_t x, uint32_t y, uint32_t z, uint32_t b) {
    return ((uint64_t)b << 48) | ((uint64_t)x << 32) | ((uint64_t)y << 16) | (uint64_t)z;
}
__host__ __device__ inline void unpack_key(uint64_t k, int &x, int &y, int &z, int &b) {
    z = (int)(k & 0xFFFF);
    y = (int)((k >> 16) & 0xFFFF);
    x = (int)((k >> 32) & 0xFFFF);
    b = (int)(k >> 48);
}
// key of the stride-2 parent: every axis >> 1, batch unchanged
__host__ __device__ inline uint64_t parent_key(uint64_t k) {
    return ((k >> 1) & 0x00007FFF7FFF7FFFull) | (k & 0xFFFF000000000000ull);
}
__host__ __device__ inline uint32_t hash_key(uint64_t h) {  // splitmix64 finaliser
    h ^= h >> 30;
    h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 27;
    h *= 0x94d049bb133111ebull;
    h ^= h >> 31;
    return (uint32_t)h;
}

}  // namespace mopa
