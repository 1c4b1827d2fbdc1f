// Probe of tcgen05.mma kind::tf32 with MN-major shared-memory operands (SWIZZLE_128B): which smem layout / descriptor
// fields does the hardware expect? Build: nvcc -gencode arch=compute_100a,code=sm_100a -I mopa_b200/csrc -I include
// scratch/umma_mn_probe.cu -o gpurun_out/umma_probe ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace mopa;

__device__ __forceinline__ uint64_t desc_make(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

// variant bits: 1 = A MN-major, 2 = B MN-major
// a_img / b_img: byte images of the operand tiles as they should sit in shared memory (host builds them)
__global__ void probe(const float *a_img, const float *b_img, int a_bytes, int b_bytes, uint32_t idesc, uint32_t a_lbo,
                      uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int n, int nmma, uint32_t a_step, uint32_t b_step,
                      float *out, uint32_t a_layout, uint32_t b_layout) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    float *sA = reinterpret_cast<float *>(smem), *sB = reinterpret_cast<float *>(smem + 32768);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_bytes / 4; i += blockDim.x) sA[i] = a_img[i];
    for (int i = tid; i < b_bytes / 4; i += blockDim.x) sB[i] = b_img[i];
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tptr, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = *tptr;
    if (warp == 0) {
        if (elect_one()) {
            for (int j = 0; j < nmma; ++j)
                umma_tf32(tb, desc_make(smem_u32(sA) + j * a_step, a_lbo, a_sbo, a_layout), desc_make(smem_u32(sB) + j * b_step, b_lbo, b_sbo, b_layout),
                          idesc, j > 0 ? 1u : 0u);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    if (warp < 4) {
        for (int q = 0; q < n / 16; ++q) {
            float v[16];
            tmem_ld16(tb + ((uint32_t)(32 * warp) << 16) + 16 * q, v);
            for (int e = 0; e < 16; ++e) out[(size_t)tid * n + 16 * q + e] = v[e];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

static uint32_t idesc_of(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}


// MN-major tf32 = SWIZZLE_128B_BASE32B (layout type 1), decoded with umma_mn_probe2: element (mn, k) at
// (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + ((((mn%32)/8) ^ (k%4)) * 32) + (mn%8)*4. Rule-major packing used by k_dw_tc:
// LBO = 512 (atoms of one 4-rule group contiguous), SBO = n_atoms*512.
static void run(int N, int n_ma_used) {
    const int M = 128, K = 32;  // four MMAs of K = 8 (one k_dw_tc stage)
    const int n_na = (N + 31) / 32;
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(N);
    for (auto &x : A) x = (float)(rand() % 7 - 3);
    for (auto &x : B) x = (float)(rand() % 5 - 2);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
            D[m * N + n] = s;
        }
    auto img = [&](const std::vector<float> &X, int rows, int n_atoms, std::vector<float> &im) {
        im.assign(8192, 0.f);
        for (int mn = 0; mn < rows; ++mn)
            for (int k = 0; k < K; ++k) {
                if (mn / 32 >= n_atoms) continue;  // rows beyond the stored atoms: whatever aliases them (garbage rows of D)
                uint32_t off = (k / 4) * (n_atoms * 512) + (mn / 32) * 512 + (k % 4) * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4;
                im[off / 4] = X[mn * K + k];
            }
    };
    float *da, *db, *dout;
    cudaMalloc(&da, 32768); cudaMalloc(&db, 32768); cudaMalloc(&dout, M * N * 4);
    std::vector<float> ia, ib, out(M * N);
    img(A, M, n_ma_used, ia);
    img(B, N, n_na, ib);
    cudaMemcpy(da, ia.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemcpy(db, ib.data(), 32768, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, M * N * 4);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    probe<<<1, 128, 68 * 1024>>>(da, db, 32768, 32768, idesc, 512, n_ma_used * 512, 512, n_na * 512, N, K / 8, n_ma_used * 1024, n_na * 1024, dout, 1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 32 * n_ma_used; ++m)
        for (int n = 0; n < N; ++n) bad += out[m * N + n] != D[m * N + n];
    printf("BASE32B rule-major: N=%3d, A atoms stored %d: %s, mismatches in the first %d rows: %d\n", N, n_ma_used, cudaGetErrorString(e), 32 * n_ma_used, bad);
    cudaFree(da); cudaFree(db); cudaFree(dout);
}

int main() {
    const int M = 128, N = 32, K = 16;  // two MMAs of K = 8
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(1);
    for (auto &x : A) x = (float)(rand() % 7 - 3);
    for (auto &x : B) x = (float)(rand() % 5 - 2);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
            D[m * N + n] = s;
        }
    float *da, *db, *dout;
    cudaMalloc(&da, 32768); cudaMalloc(&db, 32768); cudaMalloc(&dout, M * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    // K-major SW128 image: element (r, k): byte r*128 + (((k/4) ^ (r&7)) * 16) + (k%4)*4  (rows 128 B, 8-row groups 1024 B)
    auto img_k = [&](const std::vector<float> &X, int rows, std::vector<float> &img) {
        img.assign(8192, 0.f);
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < K; ++k) img[(r * 128 + (((k / 4) ^ (r & 7)) * 16) + (k % 4) * 4) / 4] = X[r * K + k];
    };
    // MN-major SW128 image, parametrised: element (mn, k): atom a = mn/32 at a*atom_stride; K group g = k/8 at g*kg_stride;
    // inside: (k%8)*128 + ((((mn%32)/4) ^ (swz ? k%8 : 0)) * 16) + (mn%4)*4
    auto img_mn = [&](const std::vector<float> &X, int rows, uint32_t atom_stride, uint32_t kg_stride, bool swz, std::vector<float> &img) {
        img.assign(8192, 0.f);
        for (int mn = 0; mn < rows; ++mn)
            for (int k = 0; k < K; ++k) {
                uint32_t off = (mn / 32) * atom_stride + (k / 8) * kg_stride + (k % 8) * 128 + ((((mn % 32) / 4) ^ (swz ? (k % 8) : 0)) * 16) + (mn % 4) * 4;
                img[off / 4] = X[mn * K + k];
            }
    };
    run(16, 4); run(32, 4); run(48, 4); run(112, 4); run(16, 1); run(64, 2); run(96, 3); run(256, 3);
    struct Var { const char *name; bool a_mn, b_mn; uint32_t atom_stride_a, kg_a, atom_stride_b, kg_b; bool swz; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; };
    // A: 4 atoms (M = 128), B: 1 atom (N = 32). K group stride: A 4096, B 1024 (atoms of one K group contiguous)
    Var vars[] = {
        {"K-major both (harness check)", false, false, 0, 0, 0, 0, true, 16, 1024, 16, 1024},
        {"MN both: LBO=atom stride, SBO=K-group stride", true, true, 1024, 4096, 1024, 1024, true, 1024, 4096, 1024, 1024},
        {"MN both: LBO=K-group stride, SBO=atom stride", true, true, 1024, 4096, 1024, 1024, true, 4096, 1024, 1024, 1024},
        {"MN both, no swizzle in image", true, true, 1024, 4096, 1024, 1024, false, 1024, 4096, 1024, 1024},
        {"A MN / B K", true, false, 1024, 4096, 0, 0, true, 1024, 4096, 16, 1024},
        {"A K / B MN", false, true, 0, 0, 1024, 1024, true, 16, 1024, 1024, 1024},
    };
    std::vector<float> ia, ib, out(M * N);
    for (auto &v : vars) {
        uint32_t a_step, b_step;
        if (v.a_mn) { img_mn(A, M, v.atom_stride_a, v.kg_a, v.swz, ia); a_step = v.kg_a; } else { img_k(A, M, ia); a_step = 32; }
        if (v.b_mn) { img_mn(B, N, v.atom_stride_b, v.kg_b, v.swz, ib); b_step = v.kg_b; } else { img_k(B, N, ib); b_step = 32; }
        cudaMemcpy(da, ia.data(), 32768, cudaMemcpyHostToDevice);
        cudaMemcpy(db, ib.data(), 32768, cudaMemcpyHostToDevice);
        cudaMemset(dout, 0xff, M * N * 4);
        probe<<<1, 128, 68 * 1024>>>(da, db, 32768, 32768, idesc_of(N, v.a_mn, v.b_mn), v.a_lbo, v.a_sbo, v.b_lbo, v.b_sbo, N, K / 8,
                                     a_step, b_step, dout, 2, 2);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0, zeros = 0;
        for (int i = 0; i < M * N; ++i) { bad += out[i] != D[i]; zeros += out[i] == 0.f; }
        printf("%-48s: %s  mismatches %d / %d  zeros %d   D[0][0..3] = %g %g %g %g (want %g %g %g %g)  D[40][5]=%g (want %g)\n", v.name,
               cudaGetErrorString(e), bad, M * N, zeros, out[0], out[1], out[2], out[3], D[0], D[1], D[2], D[3], out[40 * N + 5], D[40 * N + 5]);
    }
    return 0;
}
