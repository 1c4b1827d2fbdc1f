// Probe of tcgen05.mma's disable-output-lane operand (kind::tf32, cta_group::1, M = 128): are masked rows of D really
// left untouched, and do garbage / NaN A rows behind masked lanes stay harmless?  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I mopa_b200/csrc -I include scratch/umma_mask_probe.cu -o scratch/bin/umma_mask_probe
#include <cmath>
#include <cstdio>
#include <vector>

#include "ptx.cuh"
using namespace mopa;

__global__ void probe(const float *a_img, const float *b_img, int n, uint4 keep, float *out) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    float *sA = reinterpret_cast<float *>(smem), *sB = reinterpret_cast<float *>(smem + 16384);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 49152);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4096; i += blockDim.x) sA[i] = a_img[i];
    for (int i = tid; i < n * 32; i += blockDim.x) sB[i] = b_img[i];
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tptr, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = *tptr;
    // accumulators start at zero (tcgen05.st), as in k_conv_tc
    for (int q = 0; q < n / 16; ++q) tmem_st16_zero(tb + ((uint32_t)(32 * warp) << 16) + 16 * q);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0) {
        if (elect_one()) {
            const uint64_t hi = umma_desc_sw128(0);
            const uint64_t ad = hi | (uint64_t)((smem_u32(sA) & 0x3FFFFu) >> 4), bd = hi | (uint64_t)((smem_u32(sB) & 0x3FFFFu) >> 4);
            for (int rep = 0; rep < 2; ++rep)  // two accumulating passes: kept rows end at 2 * A.B
                for (int j = 0; j < 4; ++j)
                    umma_tf32_masked(tb, ad + 2 * j, bd + 2 * j, umma_idesc_tf32(n), ~keep.x, ~keep.y, ~keep.z, ~keep.w);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    for (int q = 0; q < n / 16; ++q) {
        float v[16];
        tmem_ld16(tb + ((uint32_t)(32 * warp) << 16) + 16 * q, v);
        for (int e = 0; e < 16; ++e) out[(size_t)tid * n + 16 * q + e] = v[e];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
    const int M = 128, K = 32;
    int total_bad = 0;
    for (int N : {16, 32, 96, 112}) {
        std::vector<float> A(M * K), B(N * K), ia(4096, 0.f), ib(N * 32, 0.f), out(M * N);
        srand(N);
        for (auto &x : A) x = (float)(rand() % 7 - 3);
        for (auto &x : B) x = (float)(rand() % 5 - 2);
        const uint32_t keep[4] = {0x0000ffffu, 0xa5a5a5a5u, 0u, 0x80000001u};
        auto kept = [&](int r) { return (keep[r / 32] >> (r % 32)) & 1u; };
        for (int r = 0; r < M; ++r)
            for (int k = 0; k < K; ++k)  // K-major SWIZZLE_128B image; masked rows hold NaN on purpose
                ia[(r * 128 + (((k / 4) ^ (r & 7)) * 16) + (k % 4) * 4) / 4] = kept(r) ? A[r * K + k] : NAN;
        for (int r = 0; r < N; ++r)
            for (int k = 0; k < K; ++k) ib[(r * 128 + (((k / 4) ^ (r & 7)) * 16) + (k % 4) * 4) / 4] = B[r * K + k];
        float *da, *db, *dout;
        cudaMalloc(&da, 16384); cudaMalloc(&db, N * 128); cudaMalloc(&dout, M * N * 4);
        cudaMemcpy(da, ia.data(), 16384, cudaMemcpyHostToDevice);
        cudaMemcpy(db, ib.data(), N * 128, cudaMemcpyHostToDevice);
        cudaMemset(dout, 0xff, M * N * 4);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        probe<<<1, 128, 52 * 1024>>>(da, db, N, make_uint4(keep[0], keep[1], keep[2], keep[3]), dout);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        int bad_kept = 0, bad_masked = 0;
        for (int r = 0; r < M; ++r)
            for (int c = 0; c < N; ++c) {
                float want = 0.f;
                if (kept(r)) { for (int k = 0; k < K; ++k) want += A[r * K + k] * B[c * K + k]; want *= 2.f; }
                const float got = out[r * N + c];
                if (!(got == want)) (kept(r) ? bad_kept : bad_masked)++;
            }
        printf("mask probe N=%3d: %s  wrong kept entries %d, wrong masked entries (must stay 0, A rows are NaN) %d\n", N,
               cudaGetErrorString(e), bad_kept, bad_masked);
        total_bad += bad_kept + bad_masked + (e != cudaSuccess);
        cudaFree(da); cudaFree(db); cudaFree(dout);
    }
    printf(total_bad ? "MASK PROBE FAILED\n" : "MASK PROBE OK\n");
    return total_bad != 0;
}
