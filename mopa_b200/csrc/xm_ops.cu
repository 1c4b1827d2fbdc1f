// Cross-modal operators on either side of the UNetSCN path, sm_100a. C ABI: include/mopa_xm.h.
//   N2  2D -> 3D feature lifting + the two segmentation heads: replaces the per-sample Python loop and nn.Linear calls of
//       Net2DSeg.forward (/root/reference/mopa/models/xmuda_arch.py:62-77); cross-modal KL loss
//       (/root/reference/mopa/train/train_xmuda_mopa.py:389-398, 440-445).
//   N4  SAM mask-consistency loss (/root/reference/mopa/common/utils/loss.py:241-283) as one segmented reduction.
// All of it is HBM-bound streaming / gather work on small tensors; nothing here is GEMM-shaped enough for tensor cores
// (the heads are 64 x 5..10 products per point, done from a shared-memory tile).
#include <math.h>

#include "common.cuh"
#include "mopa_xm.h"

namespace mopa {

// Invalid indices are reported like torch's CUDA indexing reports them: asynchronously. The kernels set a sticky flag in
// mapped pinned host memory (no stream synchronisation on the hot path); the next mopa_xm_* call on this process, or an
// explicit mopa_xm_checkAsyncError(stream), turns it into an error return.
static int *xm_err_flag() {
    static int *flag = [] {
        int *p = nullptr;
        if (cudaHostAlloc((void **)&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return (int *)nullptr;
        *p = 0;
        return p;
    }();
    return flag;
}
static int xm_take_async_error() {
    int *f = xm_err_flag();
    MOPA_CHECK(f != nullptr, "cannot allocate the pinned error flag");
    const int v = *(volatile int *)f;
    if (v == 0) return 0;
    *(volatile int *)f = 0;
    MOPA_FAIL(v == 1 ? "PixelGatherHeads (an earlier call): an image index lies outside the feature map"
                     : "MaskConsLoss (an earlier call): a mask id is >= max_ids");
}
__device__ __forceinline__ void xm_raise(int *flag, int code) {
    *(volatile int *)flag = code;
    __threadfence_system();
}

// ================================================================================================ N2: gather + heads
constexpr int kPgThreads = 128;    // 4 warps; a warp handles 32 points per pass
constexpr int kPgMaxC = 128;       // feature channels (64 in MoPA: UNetResNet34 decoder width)
constexpr int kPgMaxK = 32;        // classes (5 / 10 / 11 in MoPA's configs)
constexpr int kPgMaxBatch = 64;

struct PgOffsets {
    int64_t off[kPgMaxBatch + 1];
};

__device__ __forceinline__ int pg_sample_of(const PgOffsets &so, int batch, int64_t n) {
    int b = 0;
    while (b + 1 < batch && n >= so.off[b + 1]) ++b;  // B <= 64, offsets in constant-bank parameter space
    return b;
}

// smem: w1 [K][C], w2 [K][C], then per warp a tile [32][C + 1] of gathered features
__global__ void __launch_bounds__(kPgThreads)
    k_pixel_gather_heads_fwd(const float *__restrict__ x, int batch, int C, int H, int W, const int64_t *__restrict__ idx,
                             const __grid_constant__ PgOffsets so, int64_t n, const float *__restrict__ w1,
                             const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2, int K,
                             float *__restrict__ feats, float *__restrict__ logit, float *__restrict__ logit2,
                             int *__restrict__ err) {
    extern __shared__ float pg_smem[];
    float *sw1 = pg_smem, *sw2 = sw1 + K * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = sw2 + K * C + (size_t)warp * 32 * (C + 1);
    for (int i = threadIdx.x; i < K * C; i += kPgThreads) {
        sw1[i] = w1[i];
        sw2[i] = w2 ? w2[i] : 0.f;
    }
    __syncthreads();
    const int64_t plane = (int64_t)H * W;
    const int64_t n_pass = (n + 31) / 32;
    for (int64_t pass = (int64_t)blockIdx.x * 4 + warp; pass < n_pass; pass += (int64_t)gridDim.x * 4) {
        const int64_t p_mine = pass * 32 + lane;  // lane p resolves point p: sample, pixel
        int64_t base = -1;
        if (p_mine < n) {
            int64_t r = idx[2 * p_mine], c = idx[2 * p_mine + 1];
            if (r < 0) r += H;  // torch advanced indexing wraps negative indices
            if (c < 0) c += W;
            if (r < 0 || r >= H || c < 0 || c >= W) {
                xm_raise(err, 1);
            } else {
                base = (int64_t)pg_sample_of(so, batch, p_mine) * C * plane + r * W + c;
            }
        }
        // gather: all lanes walk the 32 points, lane = channel (+32, +64, ...): 32 x C / 32 independent loads in flight
        for (int p = 0; p < 32; ++p) {
            const int64_t bp = __shfl_sync(0xffffffffu, base, p);
            const int64_t np = pass * 32 + p;
            for (int c = lane; c < C; c += 32) {
                const float v = bp >= 0 ? __ldg(x + bp + (int64_t)c * plane) : 0.f;
                tile[p * (C + 1) + c] = v;
                if (np < n) feats[np * C + c] = v;  // coalesced along c
            }
        }
        __syncwarp();
        if (p_mine < n) {  // heads: lane p owns point p; its tile row is conflict-free (row pitch C + 1)
            const float *f = tile + lane * (C + 1);
            for (int j = 0; j < K; ++j) {
                float a1 = b1 ? b1[j] : 0.f, a2 = (w2 && b2) ? b2[j] : 0.f;
                for (int c = 0; c < C; ++c) {
                    const float v = f[c];
                    a1 = fmaf(v, sw1[j * C + c], a1);
                    a2 = fmaf(v, sw2[j * C + c], a2);
                }
                logit[p_mine * K + j] = a1;
                if (logit2) logit2[p_mine * K + j] = a2;
            }
        }
        __syncwarp();
    }
}

// backward: one block = 128 threads walks tiles of 32 points.
//   d_x[pixel(n)][c] += d_feats[n][c] + sum_j d_logit[n][j] w1[j][c] + d_logit2[n][j] w2[j][c]     (fp32 atomics)
//   per-block partial d_w / d_b in registers (thread = channel c, loops the classes), written to the workspace; the last
//   block to finish sums the block partials in block order (deterministic).
__global__ void __launch_bounds__(kPgThreads)
    k_pixel_gather_heads_bwd(const float *__restrict__ feats, const int64_t *__restrict__ idx, const __grid_constant__ PgOffsets so,
                             int batch, int C, int H, int W, int64_t n, const float *__restrict__ w1,
                             const float *__restrict__ w2, int K, const float *__restrict__ d_feats,
                             const float *__restrict__ d_logit, const float *__restrict__ d_logit2, float *__restrict__ d_x,
                             float *__restrict__ d_w1, float *__restrict__ d_b1, float *__restrict__ d_w2,
                             float *__restrict__ d_b2, float *__restrict__ ws, unsigned int *__restrict__ counter) {
    extern __shared__ float pg_smem[];
    float *sw1 = pg_smem, *sw2 = sw1 + K * C;          // [K][C]
    float *sdl1 = sw2 + K * C, *sdl2 = sdl1 + 32 * K;  // [32][K] output gradients of the current tile
    float *sf = sdl2 + 32 * K;                         // [32][C] features of the current tile
    float *red = sf + 32 * C;                          // [nsub][2][K][C] d_w partials of the point subsets
    __shared__ int64_t sbase[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    for (int i = tid; i < K * C; i += kPgThreads) {
        sw1[i] = w1[i];
        sw2[i] = w2 ? w2[i] : 0.f;
    }
    const int64_t plane = (int64_t)H * W;
    const int nsub = kPgThreads / C > 0 ? kPgThreads / C : 1;  // point subsets that share a channel column
    const int c_mine = tid % C, sub = tid / C;
    const bool active = tid < nsub * C;
    float acc1[kPgMaxK], acc2[kPgMaxK];  // d_w partials of channel c_mine (and, for c_mine == 0, d_b in accb)
#pragma unroll
    for (int j = 0; j < kPgMaxK; ++j) acc1[j] = acc2[j] = 0.f;
    float accb1 = 0.f, accb2 = 0.f;  // thread (sub 0, c = j) accumulates d_b[j] for j < K
    const int64_t n_tile = (n + 31) / 32;
    for (int64_t t = blockIdx.x; t < n_tile; t += gridDim.x) {
        __syncthreads();
        const int64_t n0 = t * 32;
        const int np = (int)min((int64_t)32, n - n0);
        for (int i = tid; i < 32 * K; i += kPgThreads) {
            const int p = i / K, j = i - p * K;
            sdl1[i] = (d_logit && p < np) ? d_logit[(n0 + p) * K + j] : 0.f;
            sdl2[i] = (d_logit2 && p < np) ? d_logit2[(n0 + p) * K + j] : 0.f;
        }
        for (int i = tid; i < 32 * C; i += kPgThreads) {
            const int p = i / C;
            sf[i] = p < np ? feats[n0 * C + i] : 0.f;
        }
        if (tid < 32) {
            int64_t b = -1;
            if (tid < np) {
                int64_t r = idx[2 * (n0 + tid)], c = idx[2 * (n0 + tid) + 1];
                if (r < 0) r += H;
                if (c < 0) c += W;
                if (r >= 0 && r < H && c >= 0 && c < W) b = (int64_t)pg_sample_of(so, batch, n0 + tid) * C * plane + r * W + c;
            }
            sbase[tid] = b;
        }
        __syncthreads();
        // scatter of the feature gradient: thread -> (point, channel) pairs, channel fastest
        if (d_x) {
            for (int i = tid; i < np * C; i += kPgThreads) {
                const int p = i / C, c = i - p * C;
                if (sbase[p] < 0) continue;
                float g = d_feats ? d_feats[(n0 + p) * C + c] : 0.f;
                for (int j = 0; j < K; ++j) g = fmaf(sdl1[p * K + j], sw1[j * C + c], fmaf(sdl2[p * K + j], sw2[j * C + c], g));
                atomicAdd(d_x + sbase[p] + (int64_t)c * plane, g);
            }
        }
        if (active) {
            for (int p = sub; p < np; p += nsub) {
                const float f = sf[p * C + c_mine];
#pragma unroll
                for (int j = 0; j < kPgMaxK; ++j)
                    if (j < K) {
                        acc1[j] = fmaf(sdl1[p * K + j], f, acc1[j]);
                        acc2[j] = fmaf(sdl2[p * K + j], f, acc2[j]);
                    }
            }
        }
        if (tid < K)
            for (int p = 0; p < np; ++p) {
                accb1 += sdl1[p * K + tid];
                accb2 += sdl2[p * K + tid];
            }
    }
    // block partials -> workspace [block][2][K][C + 1] (column C = bias); subsets combined through shared memory
    __syncthreads();
    if (active) {
#pragma unroll
        for (int j = 0; j < kPgMaxK; ++j)
            if (j < K) {
                red[((sub * 2 + 0) * K + j) * C + c_mine] = acc1[j];
                red[((sub * 2 + 1) * K + j) * C + c_mine] = acc2[j];
            }
    }
    __syncthreads();
    float *mine = ws + (size_t)blockIdx.x * 2 * K * (C + 1);
    for (int i = tid; i < 2 * K * C; i += kPgThreads) {
        const int h = i / (K * C), r = i - h * K * C, j = r / C, c = r - j * C;
        float s = 0.f;
        for (int u = 0; u < nsub; ++u) s += red[((u * 2 + h) * K + j) * C + c];
        mine[(h * K + j) * (C + 1) + c] = s;
    }
    if (tid < K) {
        mine[(0 * K + tid) * (C + 1) + C] = accb1;
        mine[(1 * K + tid) * (C + 1) + C] = accb2;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = tid; i < 2 * K * (C + 1); i += kPgThreads) {
        float s = 0.f;
        for (unsigned b = 0; b < gridDim.x; ++b) s += ws[(size_t)b * 2 * K * (C + 1) + i];
        const int h = i / (K * (C + 1)), r = i - h * K * (C + 1), j = r / (C + 1), c = r - j * (C + 1);
        float *dw = h ? d_w2 : d_w1, *db = h ? d_b2 : d_b1;
        if (c < C) { if (dw) dw[j * C + c] = s; }
        else if (db) db[j] = s;
    }
    if (tid == 0) *counter = 0;  // ready for the next call
}

// ================================================================================================ N2: KL loss
// one thread per row; block partial sums (fp64) -> workspace; the last block adds them in block order and writes the mean
__global__ void __launch_bounds__(256) k_kl_div(const float *__restrict__ student, const float *__restrict__ teacher, int64_t n,
                                               int K, float *__restrict__ loss_out, float *__restrict__ d_student,
                                               float grad_scale, double *__restrict__ partial, unsigned int *__restrict__ counter) {
    __shared__ double red[8];
    __shared__ bool is_last;
    double local = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *s = student + i * K, *t = teacher + i * K;
        float ms = -INFINITY, mt = -INFINITY;
        for (int j = 0; j < K; ++j) { ms = fmaxf(ms, s[j]); mt = fmaxf(mt, t[j]); }
        float zs = 0.f, zt = 0.f;
        for (int j = 0; j < K; ++j) { zs += expf(s[j] - ms); zt += expf(t[j] - mt); }
        const float ls = logf(zs), lt = logf(zt);
        float row = 0.f;
        for (int j = 0; j < K; ++j) {
            const float logp = s[j] - ms - ls, logq = t[j] - mt - lt, q = expf(logq);
            if (q > 0.f) row += q * (logq - logp);  // F.kl_div(reduction='none'): target * (log target - input), 0 at target 0
            if (d_student) d_student[i * K + j] = grad_scale * (expf(logp) - q) / (float)n;
        }
        local += (double)row;
    }
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        partial[blockIdx.x] = s;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) s += partial[b];
        *loss_out = (float)(s / (double)n);
    }
}

// ================================================================================================ N4: mask consistency
// stats layout (doubles): per (image b, id m): [sum x_c (C) | sum over c of x_c^2 | pixel count]; after the finalise kernel the
// first C entries hold the means. Then, per image, the number of valid ids.
__host__ __device__ inline size_t mc_stride(int C) { return (size_t)C + 2; }

__global__ void __launch_bounds__(256) k_mask_stats(const float *__restrict__ probs, const int32_t *__restrict__ masks,
                                                   int64_t pixels, int C, int M, double *__restrict__ stats,
                                                   int *__restrict__ err) {
    extern __shared__ double mc_acc[];  // [M][C + 2]
    const int b = blockIdx.y;
    const size_t stride = mc_stride(C);
    for (int i = threadIdx.x; i < M * (int)stride; i += blockDim.x) mc_acc[i] = 0.0;
    __syncthreads();
    const float *pb = probs + (int64_t)b * pixels * C;
    const int32_t *mb = masks + (int64_t)b * pixels;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (int64_t)gridDim.x * blockDim.x) {
        const int id = mb[p];
        if (id < 0) continue;
        if (id >= M) { xm_raise(err, 2); continue; }
        double *a = mc_acc + (size_t)id * stride;
        double sq = 0.0;
        for (int c = 0; c < C; ++c) {
            const double v = (double)pb[p * C + c];
            atomicAdd(a + c, v);
            sq += v * v;
        }
        atomicAdd(a + C, sq);
        atomicAdd(a + C + 1, 1.0);
    }
    __syncthreads();
    double *g = stats + (size_t)b * M * stride;
    for (int i = threadIdx.x; i < M * (int)stride; i += blockDim.x)
        if (mc_acc[i] != 0.0) atomicAdd(g + i, mc_acc[i]);
}

// one block: per image, per id: mean, mse (+ entropy of the mean); image loss = mean over its ids; loss = mean over images
__global__ void __launch_bounds__(256) k_mask_finalise(double *__restrict__ stats, int batch, int C, int M, int min_entropy,
                                                      float entropy_norm, float *__restrict__ loss_out) {
    __shared__ double red_l[8], red_c[8];
    const size_t stride = mc_stride(C);
    double total = 0.0;
    for (int b = 0; b < batch; ++b) {
        double lsum = 0.0, cnt = 0.0;
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            double *a = stats + ((size_t)b * M + m) * stride;
            const double n = a[C + 1];
            if (n <= 0.0) continue;
            double mu2 = 0.0, ent = 0.0;
            for (int c = 0; c < C; ++c) {
                const double mu = a[c] / n;
                a[c] = mu;
                mu2 += mu * mu;
                ent -= mu * log2(mu + 1e-30);
            }
            double l = (a[C] - n * mu2) / (n * C);  // mean over n x C of (x - mean)^2
            if (l < 0.0) l = 0.0;
            if (min_entropy) l += ent / (double)entropy_norm;
            lsum += l;
            cnt += 1.0;
        }
        for (int d = 16; d > 0; d >>= 1) {
            lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { red_l[threadIdx.x >> 5] = lsum; red_c[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double l = 0.0, c = 0.0;
            for (int w = 0; w < 8; ++w) { l += red_l[w]; c += red_c[w]; }
            stats[(size_t)batch * M * stride + b] = c;  // valid ids of image b (0: the image contributes 0)
            if (c > 0.0) total += l / c;
        }
    }
    if (threadIdx.x == 0) *loss_out = (float)(total / (double)batch);
}

__global__ void __launch_bounds__(256) k_mask_bwd(const float *__restrict__ probs, const int32_t *__restrict__ masks, int batch,
                                                 int64_t pixels, int C, int M, int min_entropy, float entropy_norm,
                                                 const double *__restrict__ stats, float grad_scale, float *__restrict__ d_probs) {
    const size_t stride = mc_stride(C);
    const int64_t total = (int64_t)batch * pixels;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(p / pixels);
        const int id = masks[p];
        float *d = d_probs + p * C;
        if (id < 0 || id >= M) {
            for (int c = 0; c < C; ++c) d[c] = 0.f;
            continue;
        }
        const double *a = stats + ((size_t)b * M + id) * stride;
        const double n = a[C + 1], ids = stats[(size_t)batch * M * stride + b];
        const double g = (double)grad_scale / ((double)batch * ids);
        for (int c = 0; c < C; ++c) {
            const double mu = a[c];
            double v = 2.0 * ((double)probs[p * C + c] - mu) / (n * C);
            if (min_entropy) v -= (log2(mu + 1e-30) + mu / ((mu + 1e-30) * 0.6931471805599453)) / ((double)entropy_norm * n);
            d[c] = (float)(g * v);
        }
    }
}

static int pg_fill_offsets(PgOffsets &so, const int64_t *host, int batch, int64_t n) {
    MOPA_CHECK(batch >= 1 && batch <= kPgMaxBatch, "PixelGatherHeads: batch must be in [1, 64]");
    for (int b = 0; b <= batch; ++b) so.off[b] = host[b];
    MOPA_CHECK(host[0] == 0 && host[batch] == n, "PixelGatherHeads: sample offsets must run from 0 to n");
    for (int b = 0; b < batch; ++b) MOPA_CHECK(host[b] <= host[b + 1], "PixelGatherHeads: sample offsets must not decrease");
    return 0;
}

}  // namespace mopa

using namespace mopa;

extern "C" {

int mopa_xm_PixelGatherHeads_updateOutput(const float *x, int batch, int channels, int height, int width,
                                          const int64_t *img_indices, const int64_t *sample_offsets_host, int64_t n,
                                          const float *w1, const float *b1, const float *w2, const float *b2, int classes,
                                          float *feats, float *logit, float *logit2, void *stream) {
    MOPA_CHECK(x && img_indices && sample_offsets_host && w1 && feats && logit, "PixelGatherHeads: null argument");
    MOPA_CHECK(channels >= 1 && channels <= kPgMaxC && classes >= 1 && classes <= kPgMaxK,
               "PixelGatherHeads: channels must be in [1, 128], classes in [1, 32]");
    MOPA_CHECK((w2 != nullptr) == (logit2 != nullptr), "PixelGatherHeads: w2 and logit2 go together");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    PgOffsets so;
    MOPA_TRY(pg_fill_offsets(so, sample_offsets_host, batch, n));
    MOPA_TRY(xm_take_async_error());
    int *err = xm_err_flag();
    const size_t smem = ((size_t)2 * classes * channels + (size_t)4 * 32 * (channels + 1)) * 4;
    static std::atomic<uint64_t> configured{0};
    MOPA_TRY(once_per_device(configured, [] {
        MOPA_CUDA(cudaFuncSetAttribute(k_pixel_gather_heads_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        MOPA_CUDA(cudaFuncSetAttribute(k_pixel_gather_heads_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        return 0;
    }));
    int64_t blocks = ceil_div(n, 128);
    if (blocks > 4 * (int64_t)num_sms()) blocks = 4 * (int64_t)num_sms();
    k_pixel_gather_heads_fwd<<<(unsigned)blocks, kPgThreads, smem, s>>>(x, batch, channels, height, width, img_indices, so, n, w1,
                                                                       b1, w2, b2, classes, feats, logit, logit2, err);
    MOPA_LAUNCHED();
    return 0;
}

static int pg_bwd_blocks() { return 2 * num_sms(); }

size_t mopa_xm_pixelGatherWorkspaceBytes(int channels, int classes) {
    return (size_t)2 * kNumSMs * 4 * 2 * classes * (channels + 1) * 4 + 256;  // sized for up to 4x the B200's SM count
}

int mopa_xm_PixelGatherHeads_backward(const float *feats, const int64_t *img_indices, const int64_t *sample_offsets_host,
                                      int batch, int channels, int height, int width, int64_t n, const float *w1,
                                      const float *w2, int classes, const float *d_feats, const float *d_logit,
                                      const float *d_logit2, float *d_x, float *d_w1, float *d_b1, float *d_w2, float *d_b2,
                                      void *workspace, size_t workspace_bytes, void *stream) {
    MOPA_CHECK(feats && img_indices && sample_offsets_host && w1 && workspace, "PixelGatherHeads_backward: null argument");
    MOPA_CHECK(channels >= 1 && channels <= kPgMaxC && classes >= 1 && classes <= kPgMaxK,
               "PixelGatherHeads: channels must be in [1, 128], classes in [1, 32]");
    const int nsub = kPgThreads / channels > 0 ? kPgThreads / channels : 1;
    cudaStream_t s = (cudaStream_t)stream;
    PgOffsets so;
    MOPA_TRY(pg_fill_offsets(so, sample_offsets_host, batch, n));
    int64_t blocks = ceil_div(n > 0 ? n : 1, 32);
    if (blocks > pg_bwd_blocks()) blocks = pg_bwd_blocks();
    const size_t need = (size_t)blocks * 2 * classes * (channels + 1) * 4 + 256;
    MOPA_CHECK(workspace_bytes >= need, "PixelGatherHeads_backward: workspace too small");
    // the arrival counter lives in the last 256 bytes of the workspace; the kernel leaves it at zero
    unsigned int *counter = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(workspace) + workspace_bytes - 256);
    MOPA_CUDA(cudaMemsetAsync(counter, 0, 4, s));
    const size_t smem = ((size_t)2 * classes * channels + (size_t)2 * 32 * classes + (size_t)32 * channels +
                         (size_t)nsub * 2 * classes * channels) * 4;
    static std::atomic<uint64_t> configured{0};
    MOPA_TRY(once_per_device(configured, [] {
        MOPA_CUDA(cudaFuncSetAttribute(k_pixel_gather_heads_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        return 0;
    }));
    k_pixel_gather_heads_bwd<<<(unsigned)blocks, kPgThreads, smem, s>>>(feats, img_indices, so, batch, channels, height, width, n,
                                                                       w1, w2, classes, d_feats, d_logit, d_logit2, d_x, d_w1,
                                                                       d_b1, d_w2, d_b2, reinterpret_cast<float *>(workspace),
                                                                       counter);
    MOPA_LAUNCHED();
    return 0;
}

int mopa_xm_KLDivLoss_updateOutput(const float *student, const float *teacher, int64_t n, int classes, float *loss_out,
                                   float *d_student, float grad_scale, void *stream) {
    MOPA_CHECK(student && teacher && loss_out, "KLDivLoss: null argument");
    MOPA_CHECK(n > 0 && classes >= 1 && classes <= 1024, "KLDivLoss: needs n > 0 rows and 1..1024 classes");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t blocks = ceil_div(n, 256);
    if (blocks > 2 * (int64_t)num_sms()) blocks = 2 * (int64_t)num_sms();
    double *scratch;  // block partials + arrival counter (stream-ordered pool allocation, not a result tensor)
    MOPA_CUDA(cudaMallocAsync((void **)&scratch, (size_t)(blocks + 1) * 8, s));
    unsigned int *counter = reinterpret_cast<unsigned int *>(scratch + blocks);
    MOPA_CUDA(cudaMemsetAsync(counter, 0, 8, s));
    k_kl_div<<<(unsigned)blocks, 256, 0, s>>>(student, teacher, n, classes, loss_out, d_student, grad_scale, scratch, counter);
    MOPA_LAUNCHED();
    MOPA_CUDA(cudaFreeAsync(scratch, s));
    return 0;
}

int mopa_xm_checkAsyncError(void *stream) {
    MOPA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return xm_take_async_error();
}

size_t mopa_xm_maskConsStatsBytes(int batch, int max_ids, int classes) {
    return ((size_t)batch * max_ids * mc_stride(classes) + (size_t)batch) * 8 + 64;
}

int mopa_xm_MaskConsLoss_updateOutput(const float *probs, const int32_t *masks, int batch, int64_t pixels, int classes,
                                      int max_ids, int min_entropy, float entropy_norm, void *stats, float *loss_out,
                                      void *stream) {
    MOPA_CHECK(probs && masks && stats && loss_out, "MaskConsLoss: null argument");
    MOPA_CHECK(batch >= 1 && pixels >= 1 && classes >= 1 && classes <= 64 && max_ids >= 1 && max_ids <= 4096,
               "MaskConsLoss: batch, pixels >= 1; classes in [1, 64]; max_ids in [1, 4096]");
    const size_t smem = (size_t)max_ids * mc_stride(classes) * 8;
    MOPA_CHECK(smem <= 200 * 1024, "MaskConsLoss: max_ids x (classes + 2) accumulators do not fit in shared memory");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = mopa_xm_maskConsStatsBytes(batch, max_ids, classes);
    MOPA_CUDA(cudaMemsetAsync(stats, 0, bytes, s));
    MOPA_TRY(xm_take_async_error());
    int *err = xm_err_flag();
    static std::atomic<uint64_t> configured{0};
    MOPA_TRY(once_per_device(configured, [] {
        MOPA_CUDA(cudaFuncSetAttribute(k_mask_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        return 0;
    }));
    // blocks per image: enough to fill the GPU, few enough that the per-block flush (max_ids x (C + 2) atomics) stays small
    int64_t bx = ceil_div(pixels, 256 * 16);
    const int64_t cap = ceil_div(2 * (int64_t)num_sms(), batch);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    k_mask_stats<<<dim3((unsigned)bx, (unsigned)batch), 256, smem, s>>>(probs, masks, pixels, classes, max_ids,
                                                                       reinterpret_cast<double *>(stats), err);
    MOPA_LAUNCHED();
    k_mask_finalise<<<1, 256, 0, s>>>(reinterpret_cast<double *>(stats), batch, classes, max_ids, min_entropy, entropy_norm,
                                      loss_out);
    MOPA_LAUNCHED();
    return 0;
}

int mopa_xm_MaskConsLoss_backward(const float *probs, const int32_t *masks, int batch, int64_t pixels, int classes,
                                  int max_ids, int min_entropy, float entropy_norm, const void *stats, float grad_scale,
                                  float *d_probs, void *stream) {
    MOPA_CHECK(probs && masks && stats && d_probs, "MaskConsLoss_backward: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t blocks = ceil_div((int64_t)batch * pixels, 256);
    if (blocks > 8 * (int64_t)num_sms()) blocks = 8 * (int64_t)num_sms();
    k_mask_bwd<<<(unsigned)blocks, 256, 0, s>>>(probs, masks, batch, pixels, classes, max_ids, min_entropy, entropy_norm,
                                               reinterpret_cast<const double *>(stats), grad_scale, d_probs);
    MOPA_LAUNCHED();
    return 0;
}

}  // extern "C"
