// Shared helpers for libmopa_scn (sm_100a). No torch, no third-party headers: CUDA runtime only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

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
