// BatchNorm(+leaky ReLU) and the Input/Output layer feature kernels.
// Replaces [UPSTREAM] SparseConvNet SCN/CUDA/{BatchNormalization,IOLayers}.cu, reached from
// mopa/models/scn_unet.py:26,29,30 and from scn.UNet's BatchNormLeakyReLU layers.
// All of these are pure HBM streaming kernels: 128-bit accesses along the plane axis, grid sized from the SM count,
// per-block partial statistics combined in fixed order by the last block to finish (deterministic, no float atomics).
#include <stdlib.h>

#include "geometry.cuh"
#include "mopa_scn.h"

namespace mopa {

// ------------------------------------------------------------------------------------------------ BatchNorm
constexpr int kBnMaxPlanes = 256;
// workspace layout (floats): [0] block counter (int), [8 .. 8+2C) gradMean / k of the backward pass,
// [8 + 512 ...) 2C fp64 accumulators (8-byte aligned). Zero between calls.
// The fused (single cooperative kernel) path keeps its own block further up: [kBnFusedOff] parity (int), [+2], [+3] arrival
// counters of the two parities, [+8 ...) two sets of 2 * 256 fp64 accumulators. A call uses the set its parity selects
// and clears the other one for the next call, so nothing has to be zeroed between kernels.
constexpr int kBnFusedOff = 1560;
__host__ __device__ inline size_t bn_ws_floats(int) { return kBnFusedOff + 8 + 2 * 2 * 2 * (size_t)kBnMaxPlanes; }
__device__ __forceinline__ double *bn_ws_acc(float *ws) { return reinterpret_cast<double *>(ws + 8 + 2 * kBnMaxPlanes); }

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ void get(const float *p, float (&v)[4]) {
        float4 x = *reinterpret_cast<const float4 *>(p);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    }
    static __device__ __forceinline__ void put(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Vec<1> {
    static __device__ __forceinline__ void get(const float *p, float (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void put(float *p, const float (&v)[1]) { *p = v[0]; }
};

// block = (planes / VEC, rows_per_block). BWD = false: S1 = sum(x - x0), S2 = sum((x - x0)^2)   (x0 = row 0, a shift
// that keeps the single-pass variance well conditioned). BWD = true: S1 = sum(d), S2 = sum((x - mean) d) with
// d = d_out masked by the sign of the recomputed output.
template <int VEC, bool BWD>
__global__ void __launch_bounds__(256) k_bn_stats(const float *__restrict__ x, int64_t ld_x, const float *__restrict__ dout,
                                                  int64_t ld_dout, int64_t n, int planes, float *__restrict__ ws,
                                                  const float *__restrict__ mean_in, const float *__restrict__ invstd_in,
                                                  const float *__restrict__ weight, const float *__restrict__ bias,
                                                  float leakiness, int train, float eps, float momentum,
                                                  float *__restrict__ save_mean, float *__restrict__ save_invstd,
                                                  float *__restrict__ running_mean, float *__restrict__ running_var,
                                                  float *__restrict__ d_weight, float *__restrict__ d_bias) {
    extern __shared__ float sred[];  // [blockDim.y][2 * planes]
    pdl_wait();  // (launched through launch_pdl: x / dout come from the previous kernel in the stream)
    pdl_trigger();
    const int c0 = threadIdx.x * VEC;
    float s1[VEC], s2[VEC], ref[VEC], sc[VEC], sh[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        s1[e] = s2[e] = 0.f;
        if (BWD) {
            ref[e] = mean_in[c0 + e];
            sc[e] = invstd_in[c0 + e] * weight[c0 + e];
            sh[e] = bias[c0 + e] - ref[e] * sc[e];
        } else {
            ref[e] = x[c0 + e];
        }
    }
    auto accumulate = [&](const float (&v)[VEC], const float (&d)[VEC]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (BWD) {
                const float y = fmaf(v[e], sc[e], sh[e]);
                const float dm = y > 0.f ? d[e] : d[e] * leakiness;
                s1[e] += dm;
                s2[e] = fmaf(v[e] - ref[e], dm, s2[e]);
            } else {
                const float dv = v[e] - ref[e];
                s1[e] += dv;
                s2[e] = fmaf(dv, dv, s2[e]);
            }
        }
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.y;
    int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    for (; r + 3 * stride < n; r += 4 * stride) {  // four independent rows in flight per thread
        float v[4][VEC], d[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Vec<VEC>::get(x + (r + q * stride) * ld_x + c0, v[q]);
            if (BWD) Vec<VEC>::get(dout + (r + q * stride) * ld_dout + c0, d[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) accumulate(v[q], d[q]);
    }
    for (; r < n; r += stride) {
        float v[VEC], d[VEC];
        Vec<VEC>::get(x + r * ld_x + c0, v);
        if (BWD) Vec<VEC>::get(dout + r * ld_dout + c0, d);
        accumulate(v, d);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        sred[threadIdx.y * 2 * planes + c0 + e] = s1[e];
        sred[threadIdx.y * 2 * planes + planes + c0 + e] = s2[e];
    }
    __syncthreads();
    // ---- block partial -> fp64 atomics on 2C global accumulators (a double sum is order-insensitive far below fp32
    //      resolution); the last block to finish finalises and leaves the workspace zeroed for the next call
    double *acc = bn_ws_acc(ws);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int i = tid; i < 2 * planes; i += nthr) {
        float s = 0.f;
        for (int y = 0; y < (int)blockDim.y; ++y) s += sred[y * 2 * planes + i];
        atomicAdd(acc + i, (double)s);
    }
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int *counter = reinterpret_cast<int *>(ws);
        const int done = atomicAdd(counter, 1);
        is_last = done == (int)gridDim.x - 1;
        if (is_last) *counter = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = tid; c < planes; c += nthr) {
        const double a = __ldcg(acc + c), b = __ldcg(acc + planes + c);
        acc[c] = 0.0;
        acc[planes + c] = 0.0;
        const double dn = (double)n;
        if (!BWD) {
            const double shift = (double)x[c];
            const double mean = shift + a / dn;
            double m2 = b - a * a / dn;  // sum (x - mean)^2
            if (m2 < 0.0) m2 = 0.0;
            const float invstd = (float)(1.0 / sqrt(m2 / dn + (double)eps));
            save_mean[c] = (float)mean;
            save_invstd[c] = invstd;
            running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * (float)mean;
            running_var[c] = momentum * running_var[c] + (1.f - momentum) * (float)(m2 / (n > 1 ? dn - 1.0 : 1.0));
        } else {
            const float invstd = invstd_in[c];
            if (d_weight) d_weight[c] = (float)(b * (double)invstd);
            if (d_bias) d_bias[c] = (float)a;
            ws[8 + c] = train ? (float)(a / dn) : 0.f;                                          // gradMean
            ws[8 + planes + c] = train ? (float)(b * (double)invstd * (double)invstd / dn) : 0.f;  // k
        }
    }
}

__global__ void k_bn_eval_stats(const float *__restrict__ running_mean, const float *__restrict__ running_var, float eps,
                                int planes, float *__restrict__ save_mean, float *__restrict__ save_invstd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= planes) return;
    save_mean[c] = running_mean[c];
    save_invstd[c] = 1.f / sqrtf(running_var[c] + eps);
}

template <int VEC>
__global__ void __launch_bounds__(256) k_bn_apply(const float *__restrict__ x, int64_t ld_x, float *__restrict__ out,
                                                  int64_t ld_out, int64_t n, int planes, const float *__restrict__ mean,
                                                  const float *__restrict__ invstd, const float *__restrict__ weight,
                                                  const float *__restrict__ bias, float leakiness) {
    const int c0 = threadIdx.x * VEC;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        sc[e] = invstd[c0 + e] * weight[c0 + e];
        sh[e] = bias[c0 + e] - mean[c0 + e] * sc[e];
    }
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < n; r += (int64_t)gridDim.x * blockDim.y) {
        float v[VEC];
        Vec<VEC>::get(x + r * ld_x + c0, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float y = fmaf(v[e], sc[e], sh[e]);
            v[e] = y > 0.f ? y : y * leakiness;
        }
        Vec<VEC>::put(out + r * ld_out + c0, v);
    }
}

// sums != nullptr: the two column sums S1 = sum d, S2 = sum (x - mean) d were reduced by the producing d_input convolution's
// epilogue (conv_tc.cu); every block derives gradMean / k from them, block 0 also writes the affine gradients: the whole
// BatchNorm backward is this one streaming kernel.
template <int VEC>
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float *__restrict__ x, int64_t ld_x,
                                                      const float *__restrict__ dout, int64_t ld_dout,
                                                      float *__restrict__ din, int64_t ld_din, int64_t n, int planes,
                                                      const float *__restrict__ mean, const float *__restrict__ invstd,
                                                      const float *__restrict__ weight, const float *__restrict__ bias,
                                                      const float *__restrict__ ws, float leakiness, int accumulate,
                                                      const double *__restrict__ sums, int train,
                                                      float *__restrict__ d_weight, float *__restrict__ d_bias, float inv_n) {
    pdl_wait();
    pdl_trigger();
    const int c0 = threadIdx.x * VEC;
    float sc[VEC], sh[VEC], mu[VEC], gm[VEC], kk[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        mu[e] = mean[c0 + e];
        sc[e] = invstd[c0 + e] * weight[c0 + e];
        sh[e] = bias[c0 + e] - mu[e] * sc[e];
        if (sums) {
            // (the sums are fp64 because they are accumulated with atomics in any order; their values fit fp32 math)
            const float a = (float)sums[c0 + e], b = (float)sums[kStatsLd + c0 + e], is = invstd[c0 + e];
            gm[e] = train ? a * inv_n : 0.f;
            kk[e] = train ? b * is * is * inv_n : 0.f;
            if (blockIdx.x == 0 && threadIdx.y == 0) {
                if (d_weight) d_weight[c0 + e] = b * is;
                if (d_bias) d_bias[c0 + e] = a;
            }
        } else {
            gm[e] = ws[8 + c0 + e];
            kk[e] = ws[8 + planes + c0 + e];
        }
    }
    auto one = [&](float (&v)[VEC], const float (&d)[VEC]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float y = fmaf(v[e], sc[e], sh[e]);
            const float dm = y > 0.f ? d[e] : d[e] * leakiness;
            v[e] = (dm - gm[e] - (v[e] - mu[e]) * kk[e]) * sc[e];
        }
    };
    // (one row per iteration: a thread sees ~6 rows at the largest level; four rows in flight measured 2 us slower per call)
    const int64_t stride = (int64_t)gridDim.x * blockDim.y;
    int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    for (; r < n; r += stride) {
        float v[VEC], d[VEC];
        Vec<VEC>::get(x + r * ld_x + c0, v);
        Vec<VEC>::get(dout + r * ld_dout + c0, d);
        one(v, d);
        if (accumulate) {
            float old[VEC];
            Vec<VEC>::get(din + r * ld_din + c0, old);
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = __fadd_rn(v[e], old[e]);
        }
        Vec<VEC>::put(din + r * ld_din + c0, v);
    }
}

struct BnShapeArgs {
    const float *x; int64_t ld_x; const float *dout; int64_t ld_dout; float *out; int64_t ld_out; int64_t n; int planes;
    float *ws; const float *mean_in; const float *invstd_in; const float *weight; const float *bias; float leakiness;
    int train; float eps; float momentum; float *save_mean; float *save_invstd; float *running_mean; float *running_var;
    float *d_weight; float *d_bias; int accumulate;
};
// ---- single-kernel BatchNorm (train forward, and backward): statistics, a grid-wide barrier, then the apply pass over
// the same slab of rows (second read comes from L2). Launched cooperatively so that all blocks are co-resident.
// Replaces [UPSTREAM] BatchNormalization_ForwardPass / _BackwardPass (two reduction kernels + elementwise).
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int VEC, bool BWD>
__global__ void __launch_bounds__(256) k_bn_fused(const float *__restrict__ x, int64_t ld_x, const float *__restrict__ dout,
                                                  int64_t ld_dout, float *__restrict__ out, int64_t ld_out, int64_t n,
                                                  int planes, float *__restrict__ ws, const float *__restrict__ mean_in,
                                                  const float *__restrict__ invstd_in, const float *__restrict__ weight,
                                                  const float *__restrict__ bias, float leakiness, int train, float eps,
                                                  float momentum, float *__restrict__ save_mean,
                                                  float *__restrict__ save_invstd, float *__restrict__ running_mean,
                                                  float *__restrict__ running_var, float *__restrict__ d_weight,
                                                  float *__restrict__ d_bias, int accumulate) {
    extern __shared__ float sred[];  // [blockDim.y][2 * planes]
    int *hdr = reinterpret_cast<int *>(ws + kBnFusedOff);
    const int par = *reinterpret_cast<volatile int *>(hdr) & 1;
    double *acc = reinterpret_cast<double *>(ws + kBnFusedOff + 8) + par * 2 * kBnMaxPlanes;
    unsigned *ctr = reinterpret_cast<unsigned *>(hdr) + 2 + par;
    const int c0 = threadIdx.x * VEC;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const int64_t rpb = ceil_div(n, (int64_t)gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = min(n, r0 + rpb);
    const int64_t stride = blockDim.y;
    float s1[VEC], s2[VEC], ref[VEC], sc[VEC], sh[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        s1[e] = s2[e] = 0.f;
        if (BWD) {
            ref[e] = mean_in[c0 + e];
            sc[e] = invstd_in[c0 + e] * weight[c0 + e];
            sh[e] = bias[c0 + e] - ref[e] * sc[e];
        } else {
            ref[e] = x[c0 + e];
        }
    }
    auto accumulate_row = [&](const float (&v)[VEC], const float (&d)[VEC]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (BWD) {
                const float y = fmaf(v[e], sc[e], sh[e]);
                const float dm = y > 0.f ? d[e] : d[e] * leakiness;
                s1[e] += dm;
                s2[e] = fmaf(v[e] - ref[e], dm, s2[e]);
            } else {
                const float dv = v[e] - ref[e];
                s1[e] += dv;
                s2[e] = fmaf(dv, dv, s2[e]);
            }
        }
    };
    int64_t r = r0 + threadIdx.y;
    for (; r + 3 * stride < r1; r += 4 * stride) {
        float v[4][VEC], d[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Vec<VEC>::get(x + (r + q * stride) * ld_x + c0, v[q]);
            if (BWD) Vec<VEC>::get(dout + (r + q * stride) * ld_dout + c0, d[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) accumulate_row(v[q], d[q]);
    }
    for (; r < r1; r += stride) {
        float v[VEC], d[VEC];
        Vec<VEC>::get(x + r * ld_x + c0, v);
        if (BWD) Vec<VEC>::get(dout + r * ld_dout + c0, d);
        accumulate_row(v, d);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        sred[threadIdx.y * 2 * planes + c0 + e] = s1[e];
        sred[threadIdx.y * 2 * planes + planes + c0 + e] = s2[e];
    }
    __syncthreads();
    for (int i = tid; i < 2 * planes; i += nthr) {
        float s = 0.f;
        for (int y = 0; y < (int)blockDim.y; ++y) s += sred[y * 2 * planes + i];
        atomicAdd(acc + i, (double)s);
    }
    // ---- grid-wide barrier (all blocks are resident: cooperative launch)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < gridDim.x) {}
    }
    __syncthreads();
    // ---- per-plane coefficients from the fp64 sums
    const double dn = (double)n;
    float gm[VEC], kk[VEC], mu[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const int c = c0 + e;
        const double a = __ldcg(acc + c), b = __ldcg(acc + planes + c);
        const bool writer = blockIdx.x == 0 && threadIdx.y == 0;
        if (!BWD) {
            const double mean = (double)ref[e] + a / dn;
            double m2 = b - a * a / dn;  // sum (x - mean)^2
            if (m2 < 0.0) m2 = 0.0;
            const float invstd = (float)(1.0 / sqrt(m2 / dn + (double)eps));
            sc[e] = invstd * weight[c];
            sh[e] = bias[c] - (float)mean * sc[e];
            if (writer) {
                save_mean[c] = (float)mean;
                save_invstd[c] = invstd;
                running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * (float)mean;
                running_var[c] = momentum * running_var[c] + (1.f - momentum) * (float)(m2 / (n > 1 ? dn - 1.0 : 1.0));
            }
        } else {
            const float invstd = invstd_in[c];
            mu[e] = ref[e];
            gm[e] = train ? (float)(a / dn) : 0.f;
            kk[e] = train ? (float)(b * (double)invstd * (double)invstd / dn) : 0.f;
            if (writer) {
                if (d_weight) d_weight[c] = (float)(b * (double)invstd);
                if (d_bias) d_bias[c] = (float)a;
            }
        }
    }
    if (blockIdx.x == 0) {  // leave the other parity's block clean for the next call, then flip
        double *other = reinterpret_cast<double *>(ws + kBnFusedOff + 8) + (par ^ 1) * 2 * kBnMaxPlanes;
        for (int i = tid; i < 2 * kBnMaxPlanes; i += nthr) other[i] = 0.0;
        if (tid == 0) {
            reinterpret_cast<unsigned *>(hdr)[2 + (par ^ 1)] = 0u;
            hdr[0] = par ^ 1;
        }
    }
    if (out == nullptr) return;
    // ---- apply over the same slab
    auto apply_row = [&](float (&v)[VEC], const float (&d)[VEC]) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float y = fmaf(v[e], sc[e], sh[e]);
            if (BWD) {
                const float dm = y > 0.f ? d[e] : d[e] * leakiness;
                v[e] = (dm - gm[e] - (v[e] - mu[e]) * kk[e]) * sc[e];
            } else {
                v[e] = y > 0.f ? y : y * leakiness;
            }
        }
    };
    r = r0 + threadIdx.y;
    for (; r + 3 * stride < r1; r += 4 * stride) {
        float v[4][VEC], d[4][VEC], o[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Vec<VEC>::get(x + (r + q * stride) * ld_x + c0, v[q]);
            if (BWD) Vec<VEC>::get(dout + (r + q * stride) * ld_dout + c0, d[q]);
            if (BWD && accumulate) Vec<VEC>::get(out + (r + q * stride) * ld_out + c0, o[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            apply_row(v[q], d[q]);
            if (BWD && accumulate) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[q][e] = __fadd_rn(v[q][e], o[q][e]);
            }
            Vec<VEC>::put(out + (r + q * stride) * ld_out + c0, v[q]);
        }
    }
    for (; r < r1; r += stride) {
        float v[VEC], d[VEC];
        Vec<VEC>::get(x + r * ld_x + c0, v);
        if (BWD) Vec<VEC>::get(dout + r * ld_dout + c0, d);
        apply_row(v, d);
        if (BWD && accumulate) {
            float o[VEC];
            Vec<VEC>::get(out + r * ld_out + c0, o);
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = __fadd_rn(v[e], o[e]);
        }
        Vec<VEC>::put(out + r * ld_out + c0, v);
    }
}

// blocks of the cooperative launch: enough rows per thread to amortise the barrier, never more than can be co-resident
template <int VEC, bool BWD>
static int launch_bn_fused(const BnShapeArgs &A, cudaStream_t s);

// ---- train-mode forward when the producing convolution already accumulated the column sums (conv_tc epilogue):
// every block derives mean / invstd from the fp64 sums (sum x, sum x^2) and applies them; no statistics pass, no barrier.
template <int VEC>
__global__ void __launch_bounds__(256) k_bn_apply_sums(const float *__restrict__ x, int64_t ld_x, float *__restrict__ out,
                                                       int64_t ld_out, int64_t n, int planes,
                                                       const double *__restrict__ sums, const float *__restrict__ weight,
                                                       const float *__restrict__ bias, float leakiness, float eps,
                                                       float momentum, float *__restrict__ save_mean,
                                                       float *__restrict__ save_invstd, float *__restrict__ running_mean,
                                                       float *__restrict__ running_var, double inv_n) {
    pdl_wait();
    pdl_trigger();
    const int c0 = threadIdx.x * VEC;
    const double dn = (double)n;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const int c = c0 + e;
        // every thread repeats this per-column prologue for a handful of rows: fp64 only where the cancellation needs it
        // (two multiplies and one FMA), no fp64 division or square root
        const double mean = sums[c] * inv_n;
        double m2 = fma(-mean * dn, mean, sums[kStatsLd + c]);  // sum (x - mean)^2
        if (m2 < 0.0) m2 = 0.0;
        const float var = (float)(m2 * inv_n) + eps;
        float invstd = rsqrtf(var);
        invstd = invstd * fmaf(-0.5f * var * invstd, invstd, 1.5f);  // one Newton step: full fp32 accuracy
        sc[e] = invstd * weight[c];
        sh[e] = bias[c] - (float)mean * sc[e];
        if (blockIdx.x == 0 && threadIdx.y == 0) {
            save_mean[c] = (float)mean;
            save_invstd[c] = invstd;
            running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * (float)mean;
            running_var[c] = momentum * running_var[c] + (1.f - momentum) * (float)(m2 / (n > 1 ? dn - 1.0 : 1.0));
        }
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.y;
    int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    for (; r + 3 * stride < n; r += 4 * stride) {
        float v[4][VEC];
#pragma unroll
        for (int q = 0; q < 4; ++q) Vec<VEC>::get(x + (r + q * stride) * ld_x + c0, v[q]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float y = fmaf(v[q][e], sc[e], sh[e]);
                v[q][e] = y > 0.f ? y : y * leakiness;
            }
            Vec<VEC>::put(out + (r + q * stride) * ld_out + c0, v[q]);
        }
    }
    for (; r < n; r += stride) {
        float v[VEC];
        Vec<VEC>::get(x + r * ld_x + c0, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float y = fmaf(v[e], sc[e], sh[e]);
            v[e] = y > 0.f ? y : y * leakiness;
        }
        Vec<VEC>::put(out + r * ld_out + c0, v);
    }
}

struct BnShape {
    int vec;
    dim3 block;
    unsigned grid;
    size_t smem;
};
static BnShape bn_shape(int64_t n, int planes, bool vec_ok) {
    BnShape s;
    s.vec = (vec_ok && planes % 4 == 0) ? 4 : 1;
    int tx = planes / s.vec;
    int ty = 256 / tx;
    if (ty < 1) ty = 1;
    if (ty > 64) ty = 64;
    s.block = dim3(tx, ty);
    int64_t want = ceil_div(n > 0 ? n : 1, (int64_t)ty * 4);  // >= 4 rows per thread
    const int64_t cap = 4 * (int64_t)num_sms();
    s.grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    s.smem = (size_t)ty * 2 * planes * 4;
    return s;
}
static bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }

static bool bn_fused_enabled() {  // MOPA_SCN_NO_BNFUSED=1: the two-kernel path (A/B measurements; read per call)
    const char *e = getenv("MOPA_SCN_NO_BNFUSED");
    return !(e && e[0] == '1');
}
// The cooperative single-kernel BatchNorm (statistics -> grid barrier -> apply) is only used from this many elements
// (rows x planes) on; the default is "never". Measured in round 2 (profiles/r02_bn_paths.txt): a cooperative launch needs
// its whole grid co-resident, so it cannot start while the d_weight kernels of the side stream occupy SMs, and nothing else
// can start while it spins on its grid barrier: in the backward pass that serialises the two streams. Two plain kernels
// per BatchNorm cost ~10 % more device time per op in isolation and make the whole step 0.15-0.25 ms faster.
// MOPA_SCN_BN_FUSED_MIN=<elements> re-enables the cooperative kernel from that size on (A/B measurements, tests).
static int64_t bn_fused_min_elems() {
    static const int64_t v = [] {
        const char *e = getenv("MOPA_SCN_BN_FUSED_MIN");
        return e ? (int64_t)atoll(e) : (int64_t)1 << 62;
    }();
    return v;
}
// ... and up to this many elements (default 0: never). Small tensors are the case where the cooperative kernel cannot hurt
// the other stream (its grid is a handful of blocks) and where the second launch + the fence / counter chain of the
// two-kernel path cost most relative to the work: MOPA_SCN_BN_FUSED_MAX=<elements>.
static int64_t bn_fused_max_elems() {
    static const int64_t v = [] {
        const char *e = getenv("MOPA_SCN_BN_FUSED_MAX");
        return e ? (int64_t)atoll(e) : (int64_t)0;
    }();
    return v;
}
static bool bn_use_fused(int64_t elems) { return elems >= bn_fused_min_elems() || elems <= bn_fused_max_elems(); }
template <int VEC, bool BWD>
static int launch_bn_fused(const BnShapeArgs &A, cudaStream_t s) {
    const int tx = A.planes / VEC;
    int ty = 256 / tx;
    if (ty < 1) ty = 1;
    if (ty > 64) ty = 64;
    const size_t smem = (size_t)ty * 2 * A.planes * 4;
    auto kern = k_bn_fused<VEC, BWD>;
    static std::atomic<int> blocks_of[64];  // per instantiation and device: co-resident blocks of a cooperative launch
    int dev = 0;
    MOPA_CUDA(cudaGetDevice(&dev));
    MOPA_CHECK(dev >= 0 && dev < 64, "device index out of range");
    int max_blocks = blocks_of[dev].load(std::memory_order_relaxed);
    if (max_blocks == 0) {
        int occ = 0;
        MOPA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 16 * 1024));
        if (occ > 2) occ = 2;  // (4 blocks per SM measured slower: the grid barrier and the 2C atomics per block grow)
        MOPA_CHECK(occ >= 1, "BatchNormalization: the fused kernel does not fit on this device");
        max_blocks = occ * num_sms();
        blocks_of[dev].store(max_blocks, std::memory_order_relaxed);
    }
    int64_t want = ceil_div(A.n, (int64_t)ty * 8);  // >= 8 rows per thread
    if (want < 1) want = 1;
    if (want > max_blocks) want = max_blocks;
    BnShapeArgs a = A;
    void *args[] = {&a.x, &a.ld_x, &a.dout, &a.ld_dout, &a.out, &a.ld_out, &a.n, &a.planes, &a.ws, &a.mean_in, &a.invstd_in,
                    &a.weight, &a.bias, &a.leakiness, &a.train, &a.eps, &a.momentum, &a.save_mean, &a.save_invstd,
                    &a.running_mean, &a.running_var, &a.d_weight, &a.d_bias, &a.accumulate};
    MOPA_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)want), dim3(tx, ty), args, smem, s));
    MOPA_LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------------ IO layers
// InputLayer mode 4: out[v][c] = sum over the voxel's rows, ascending, of (1/n_v) * in[row][c]  (multiply, then add)
__global__ void __launch_bounds__(256) k_pool_fwd(const float *__restrict__ in, int64_t ld_in, int planes,
                                                  const int32_t *__restrict__ off, const int32_t *__restrict__ rows,
                                                  int64_t V, float *__restrict__ out, int64_t ld_out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= V * planes) return;
    const int64_t v = idx / planes;
    const int c = (int)(idx - v * planes);
    const int beg = off[v], end = off[v + 1];
    const float inv = 1.f / (float)(end - beg);
    float acc = 0.f;
    for (int j = beg; j < end; ++j) acc = __fadd_rn(acc, __fmul_rn(in[(int64_t)rows[j] * ld_in + c], inv));
    out[v * ld_out + c] = acc;
}
__global__ void __launch_bounds__(256) k_pool_bwd(float *__restrict__ din, int64_t ld_din, int planes,
                                                  const int32_t *__restrict__ off, const int32_t *__restrict__ p2v,
                                                  int64_t n, const float *__restrict__ dout, int64_t ld_dout) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * planes) return;
    const int64_t i = idx / planes;
    const int c = (int)(idx - i * planes);
    const int v = p2v[i];
    const float inv = 1.f / (float)(off[v + 1] - off[v]);
    din[i * ld_din + c] = __fmul_rn(dout[(int64_t)v * ld_dout + c], inv);
}
// OutputLayer: out[i] = in[voxel(i)]; one thread moves VEC planes
template <int VEC>
__global__ void __launch_bounds__(256) k_unpool_fwd(const float *__restrict__ in, int64_t ld_in, int planes,
                                                    const int32_t *__restrict__ p2v, int64_t n, float *__restrict__ out,
                                                    int64_t ld_out) {
    const int per_row = planes / VEC;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * per_row) return;
    const int64_t i = idx / per_row;
    const int c = (int)(idx - i * per_row) * VEC;
    float v[VEC];
    Vec<VEC>::get(in + (int64_t)p2v[i] * ld_in + c, v);
    Vec<VEC>::put(out + i * ld_out + c, v);
}
template <int VEC>
__global__ void __launch_bounds__(256) k_unpool_bwd(float *__restrict__ din, int64_t ld_din, int planes,
                                                    const int32_t *__restrict__ off, const int32_t *__restrict__ rows,
                                                    int64_t V, const float *__restrict__ dout, int64_t ld_dout) {
    const int per_row = planes / VEC;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= V * per_row) return;
    const int64_t v = idx / per_row;
    const int c = (int)(idx - v * per_row) * VEC;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int j = off[v]; j < off[v + 1]; ++j) {
        float d[VEC];
        Vec<VEC>::get(dout + (int64_t)rows[j] * ld_dout + c, d);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = __fadd_rn(acc[e], d[e]);
    }
    Vec<VEC>::put(din + v * ld_din + c, acc);
}

size_t bn_workspace_bytes(int planes) { return bn_ws_floats(planes) * 4; }

int bn_forward(const float *in, int64_t ld_in, float *out, int64_t ld_out, float *save_mean, float *save_invstd,
               float *running_mean, float *running_var, const float *weight, const float *bias, float eps, float momentum,
               int train, float leakiness, int64_t n_active, int planes, void *workspace, cudaStream_t s,
               const double *stats) {
    MOPA_CHECK(planes > 0 && planes <= 256, "BatchNormalization: planes must be in [1, 256]");
    MOPA_CHECK(weight && bias, "BatchNormalization: affine parameters are required");
    if (n_active == 0) return 0;
    const bool vec_ok = al16(in) && al16(out) && ld_in % 4 == 0 && ld_out % 4 == 0;
    BnShape sh = bn_shape(n_active, planes, vec_ok);
    MOPA_CHECK(sh.block.x * sh.block.y <= 256, "BatchNormalization: unaligned features with more than 256 planes");
    float *ws = reinterpret_cast<float *>(workspace);
    const int prof = prof_begin(40, nullptr, planes, planes, n_active, s);
    if (train && stats && sh.vec == 4) {  // the producing convolution accumulated the column sums in its epilogue
        MOPA_CUDA(launch_pdl(k_bn_apply_sums<4>, dim3(sh.grid), sh.block, 0, s, in, ld_in, out, ld_out, n_active, planes, stats,
                             weight, bias, leakiness, eps, momentum, save_mean, save_invstd, running_mean, running_var,
                             1.0 / (double)n_active));
        MOPA_LAUNCHED();
        prof_end(prof, s);
        return 0;
    }
    if (train && sh.vec == 4 && bn_fused_enabled() && bn_use_fused(n_active * planes)) {
        const BnShapeArgs A{in, ld_in, nullptr, 0, out, ld_out, n_active, planes, ws, nullptr, nullptr, weight, bias, leakiness,
                            1, eps, momentum, save_mean, save_invstd, running_mean, running_var, nullptr, nullptr, 0};
        const int rc = launch_bn_fused<4, false>(A, s);
        prof_end(prof, s);
        return rc;
    }
    if (train) {
        if (sh.vec == 4)
            k_bn_stats<4, false><<<sh.grid, sh.block, sh.smem, s>>>(in, ld_in, nullptr, 0, n_active, planes, ws, nullptr,
                                                                    nullptr, nullptr, nullptr, leakiness, 1, eps, momentum,
                                                                    save_mean, save_invstd, running_mean, running_var,
                                                                    nullptr, nullptr);
        else
            k_bn_stats<1, false><<<sh.grid, sh.block, sh.smem, s>>>(in, ld_in, nullptr, 0, n_active, planes, ws, nullptr,
                                                                    nullptr, nullptr, nullptr, leakiness, 1, eps, momentum,
                                                                    save_mean, save_invstd, running_mean, running_var,
                                                                    nullptr, nullptr);
    } else {
        k_bn_eval_stats<<<(unsigned)ceil_div(planes, 128), 128, 0, s>>>(running_mean, running_var, eps, planes, save_mean,
                                                                        save_invstd);
    }
    MOPA_LAUNCHED();
    if (sh.vec == 4)
        k_bn_apply<4><<<sh.grid, sh.block, 0, s>>>(in, ld_in, out, ld_out, n_active, planes, save_mean, save_invstd, weight,
                                                   bias, leakiness);
    else
        k_bn_apply<1><<<sh.grid, sh.block, 0, s>>>(in, ld_in, out, ld_out, n_active, planes, save_mean, save_invstd, weight,
                                                   bias, leakiness);
    prof_end(prof, s);
    MOPA_LAUNCHED();
    return 0;
}

int bn_backward(const float *in, int64_t ld_in, float *d_in, int64_t ld_din, const float *d_out, int64_t ld_dout,
                const float *save_mean, const float *save_invstd, const float *weight, const float *bias, float *d_weight,
                float *d_bias, float leakiness, int train, int64_t n_active, int planes, void *workspace, int accumulate,
                cudaStream_t s, const double *sums) {
    MOPA_CHECK(planes > 0 && planes <= 256, "BatchNormalization: planes must be in [1, 256]");
    if (n_active == 0) {
        if (d_weight) MOPA_CUDA(cudaMemsetAsync(d_weight, 0, (size_t)planes * 4, s));
        if (d_bias) MOPA_CUDA(cudaMemsetAsync(d_bias, 0, (size_t)planes * 4, s));
        return 0;
    }
    const bool vec_ok = al16(in) && al16(d_in) && al16(d_out) && ld_in % 4 == 0 && ld_din % 4 == 0 && ld_dout % 4 == 0;
    BnShape sh = bn_shape(n_active, planes, vec_ok);
    MOPA_CHECK(sh.block.x * sh.block.y <= 256, "BatchNormalization: unaligned features with more than 256 planes");
    float *ws = reinterpret_cast<float *>(workspace);
    const int prof = prof_begin(50, nullptr, planes, planes, n_active, s);
    if (sums && d_in && sh.vec == 4) {  // the producing d_input convolution reduced the column sums in its epilogue
        MOPA_CUDA(launch_pdl(k_bn_bwd_apply<4>, dim3(sh.grid), sh.block, 0, s, in, ld_in, d_out, ld_dout, d_in, ld_din, n_active,
                             planes, save_mean, save_invstd, weight, bias, (const float *)ws, leakiness, accumulate, sums, train,
                             d_weight, d_bias, (float)(1.0 / (double)n_active)));
        MOPA_LAUNCHED();
        prof_end(prof, s);
        return 0;
    }
    if (sh.vec == 4 && bn_fused_enabled() && bn_use_fused(n_active * planes)) {
        const BnShapeArgs A{in, ld_in, d_out, ld_dout, d_in, ld_din, n_active, planes, ws, save_mean, save_invstd, weight, bias,
                            leakiness, train, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, d_weight, d_bias, accumulate};
        const int rc = launch_bn_fused<4, true>(A, s);
        prof_end(prof, s);
        return rc;
    }
    if (sh.vec == 4)
        MOPA_CUDA(launch_pdl(k_bn_stats<4, true>, dim3(sh.grid), sh.block, sh.smem, s, in, ld_in, d_out, ld_dout, n_active, planes,
                             ws, save_mean, save_invstd, weight, bias, leakiness, train, 0.f, 0.f, (float *)nullptr,
                             (float *)nullptr, (float *)nullptr, (float *)nullptr, d_weight, d_bias));
    else
        k_bn_stats<1, true><<<sh.grid, sh.block, sh.smem, s>>>(in, ld_in, d_out, ld_dout, n_active, planes, ws, save_mean,
                                                               save_invstd, weight, bias, leakiness, train, 0.f, 0.f,
                                                               nullptr, nullptr, nullptr, nullptr, d_weight, d_bias);
    MOPA_LAUNCHED();
    if (d_in) {
        if (sh.vec == 4)
            MOPA_CUDA(launch_pdl(k_bn_bwd_apply<4>, dim3(sh.grid), sh.block, 0, s, in, ld_in, d_out, ld_dout, d_in, ld_din,
                                 n_active, planes, save_mean, save_invstd, weight, bias, (const float *)ws, leakiness, accumulate,
                                 (const double *)nullptr, train, (float *)nullptr, (float *)nullptr, 0.f));
        else
            k_bn_bwd_apply<1><<<sh.grid, sh.block, 0, s>>>(in, ld_in, d_out, ld_dout, d_in, ld_din, n_active, planes,
                                                           save_mean, save_invstd, weight, bias, ws, leakiness, accumulate,
                                                           nullptr, train, nullptr, nullptr, 0.f);
        MOPA_LAUNCHED();
    }
    prof_end(prof, s);
    return 0;
}

}  // namespace mopa

using namespace mopa;

extern "C" {

size_t mopa_scn_bnWorkspaceBytes(int planes) { return bn_workspace_bytes(planes); }

int mopa_scn_BatchNormalization_updateOutput(const float *in, int64_t ld_in, float *out, int64_t ld_out,
                                             float *save_mean, float *save_invstd, float *running_mean,
                                             float *running_var, const float *weight, const float *bias, float eps,
                                             float momentum, int train, float leakiness, int64_t n_active, int planes,
                                             void *workspace, size_t workspace_bytes, void *stream) {
    MOPA_CHECK(workspace && workspace_bytes >= bn_workspace_bytes(planes), "BatchNormalization: workspace too small");
    return bn_forward(in, ld_in, out, ld_out, save_mean, save_invstd, running_mean, running_var, weight, bias, eps, momentum,
                      train, leakiness, n_active, planes, workspace, (cudaStream_t)stream);
}

int mopa_scn_BatchNormalization_backward(const float *in, int64_t ld_in, float *d_in, int64_t ld_din,
                                         const float *d_out, int64_t ld_dout, const float *save_mean,
                                         const float *save_invstd, const float *weight, const float *bias,
                                         float *d_weight, float *d_bias, float leakiness, int train, int64_t n_active,
                                         int planes, void *workspace, size_t workspace_bytes, void *stream) {
    MOPA_CHECK(workspace && workspace_bytes >= bn_workspace_bytes(planes), "BatchNormalization: workspace too small");
    return bn_backward(in, ld_in, d_in, ld_din, d_out, ld_dout, save_mean, save_invstd, weight, bias, d_weight, d_bias,
                       leakiness, train, n_active, planes, workspace, 0, (cudaStream_t)stream);
}

static int io_check(mopa_scn_metadata *m, int planes) {
    MOPA_CHECK(m != nullptr && !m->levels.empty(), "metadata has no input layer");
    MOPA_CHECK(planes > 0, "planes must be positive");
    MOPA_CUDA(cudaSetDevice(m->device));
    return 0;
}

int mopa_scn_InputLayer_updateOutput(mopa_scn_metadata *m, const float *in, int64_t ld_in, int planes, float *out,
                                     int64_t ld_out, void *stream) {
    MOPA_TRY(io_check(m, planes));
    const int64_t V = m->levels[0].V;
    if (V == 0) return 0;
    k_pool_fwd<<<(unsigned)ceil_div(V * planes, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, planes, m->csr_off,
                                                                                      m->csr_rows, V, out, ld_out);
    MOPA_LAUNCHED();
    return 0;
}

int mopa_scn_InputLayer_updateGradInput(mopa_scn_metadata *m, float *d_in, int64_t ld_din, const float *d_out,
                                        int64_t ld_dout, int planes, void *stream) {
    MOPA_TRY(io_check(m, planes));
    const int64_t n = m->n_points;
    if (n == 0) return 0;
    k_pool_bwd<<<(unsigned)ceil_div(n * planes, 256), 256, 0, (cudaStream_t)stream>>>(d_in, ld_din, planes, m->csr_off,
                                                                                      m->p2v, n, d_out, ld_dout);
    MOPA_LAUNCHED();
    return 0;
}

int mopa_scn_OutputLayer_updateOutput(mopa_scn_metadata *m, const float *in, int64_t ld_in, int planes, float *out,
                                      int64_t ld_out, void *stream) {
    MOPA_TRY(io_check(m, planes));
    const int64_t n = m->n_points;
    if (n == 0) return 0;
    const bool v4 = planes % 4 == 0 && al16(in) && al16(out) && ld_in % 4 == 0 && ld_out % 4 == 0;
    if (v4)
        k_unpool_fwd<4><<<(unsigned)ceil_div(n * (planes / 4), 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, planes, m->p2v,
                                                                                                     n, out, ld_out);
    else
        k_unpool_fwd<1><<<(unsigned)ceil_div(n * planes, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, planes, m->p2v, n,
                                                                                               out, ld_out);
    MOPA_LAUNCHED();
    return 0;
}

int mopa_scn_OutputLayer_updateGradInput(mopa_scn_metadata *m, float *d_in, int64_t ld_din, const float *d_out,
                                         int64_t ld_dout, int planes, void *stream) {
    MOPA_TRY(io_check(m, planes));
    const int64_t V = m->levels[0].V;
    if (V == 0) return 0;
    const bool v4 = planes % 4 == 0 && al16(d_in) && al16(d_out) && ld_din % 4 == 0 && ld_dout % 4 == 0;
    if (v4)
        k_unpool_bwd<4><<<(unsigned)ceil_div(V * (planes / 4), 256), 256, 0, (cudaStream_t)stream>>>(
            d_in, ld_din, planes, m->csr_off, m->csr_rows, V, d_out, ld_dout);
    else
        k_unpool_bwd<1><<<(unsigned)ceil_div(V * planes, 256), 256, 0, (cudaStream_t)stream>>>(
            d_in, ld_din, planes, m->csr_off, m->csr_rows, V, d_out, ld_dout);
    MOPA_LAUNCHED();
    return 0;
}

}  // extern "C"
