// Empirical decode of the shared-memory layout tcgen05.mma kind::tf32 reads for an MN-major A operand: a single 1.0 is
// walked over the first 16 KB of the A region; B is an identity (K-major, known-good), so D[m][k] = A[m][k] tells which
// (m, k) that word is. One table per (layout type, LBO, SBO).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace mopa;

__device__ __forceinline__ uint64_t desc_make(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

constexpr int kWords = 4096;  // 16 KB probed

__global__ void probe(uint32_t idesc, uint32_t layout, uint32_t lbo, uint32_t sbo, int *map /*[kWords]*/) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    float *sA = reinterpret_cast<float *>(smem), *sB = reinterpret_cast<float *>(smem + 32768);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += blockDim.x) sA[i] = 0.f;
    for (int i = tid; i < 2048; i += blockDim.x) sB[i] = 0.f;
    __syncthreads();
    if (tid < 8) {  // B[n][k] = (n == k), K-major SW128: row n at n*128, chunk (k/4) ^ (n&7)
        const int n = tid, k = tid;
        sB[(n * 128 + (((k / 4) ^ (n & 7)) * 16) + (k % 4) * 4) / 4] = 1.f;
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tptr, 32);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = *tptr;
    uint32_t ph = 0;
    for (int w = 0; w < kWords; ++w) {
        if (tid == 0) sA[w] = 1.f;
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                umma_tf32(tb, desc_make(smem_u32(sA), lbo, sbo, layout), desc_make(smem_u32(sB), 16, 1024, 2), idesc, 0u);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after_sync();
        float v[16];
        tmem_ld16(tb + ((uint32_t)(32 * warp) << 16), v);
        for (int e = 0; e < 8; ++e)
            if (v[e] != 0.f) map[w] = tid * 8 + e;  // m * 8 + k
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) sA[w] = 0.f;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 32);
}

int main() {
    int *dmap;
    cudaMalloc(&dmap, kWords * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    struct Cfg { uint32_t layout, lbo, sbo; } cfgs[] = {{2, 1024, 4096}, {2, 4096, 1024}, {1, 1024, 4096}, {0, 1024, 4096}, {0, 128, 4096}, {4, 1024, 4096}, {6, 1024, 4096}};
    std::vector<int> map(kWords);
    for (auto &c : cfgs) {
        cudaMemset(dmap, 0xff, kWords * 4);
        probe<<<1, 128, 68 * 1024>>>(idesc, c.layout, c.lbo, c.sbo, dmap);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(map.data(), dmap, kWords * 4, cudaMemcpyDeviceToHost);
        int hits = 0;
        for (int w = 0; w < kWords; ++w) hits += map[w] >= 0;
        printf("== layout %u LBO %u SBO %u : %s, %d of %d words are read (expect 1024)\n", c.layout, c.lbo, c.sbo, cudaGetErrorString(e), hits, kWords);
        // byte offset of (m, k) for a few m
        std::vector<int> off(128 * 8, -1);
        for (int w = 0; w < kWords; ++w)
            if (map[w] >= 0) off[map[w]] = w * 4;
        const int ms[] = {0, 1, 2, 3, 4, 5, 7, 8, 12, 16, 28, 31, 32, 33, 64, 96, 127};
        for (int m : ms) {
            printf("  m=%3d:", m);
            for (int k = 0; k < 8; ++k) printf(" %6d", off[m * 8 + k]);
            printf("\n");
        }
    }
    return 0;
}
