// Sparse convolution kernels: output-stationary gather -> TF32 MMA -> single coalesced store, and the weight-gradient
// reduction. Replaces [UPSTREAM] SparseConvNet SCN/CUDA/{Convolution,Deconvolution}.cu (dConvolution_KMxKN_forward*,
// _backward_dW*), reached from mopa/models/scn_unet.py:27-28.
//
// Upstream: one launch per filter offset, each preceded by a blocking H2D copy of that offset's rule list, fp32 SIMT FMA,
// read-modify-write of the output rows. Here: ONE launch per layer. A CTA owns a tile of OUTPUT rows; for every filter
// offset k it gathers the needed input rows straight into MMA A-fragments (128-bit loads, the K axis permuted so that a
// thread's float4 feeds two k-steps), multiplies with W[k] staged in shared memory by cp.async (double buffered,
// pre-packed in fragment order so a warp reads it with conflict-free LDS.128), accumulates in registers in ascending k,
// and writes each output row exactly once with 128-bit stores. No atomics; results are deterministic.
//
//   forward  submanifold : T = nbr table (27, V)          out[o] = sum_k in[T[k][o]] W[k]
//   forward  convolution : T = child table (8, Vcoarse)   out[p] = sum_k in[T[k][p]] W[k]
//   forward  deconv      : select(parent, kidx)           out[c] = in[parent[c]] W[kidx[c]]
//   d_input  submanifold : same table, weights W[26-k]^T  (rule (i,o,k) <-> rule (o,i,26-k))
//   d_input  convolution : select form with W[k]^T ;  d_input deconv : child table with W[k]^T
//   d_weight             : per (k, row chunk): block-local ordered compaction of the chunk's rules, cp.async staging of
//                          the gathered in/dOut rows, MMA with M = nIn, N = nOut, K = rules; per-chunk partials are
//                          summed in fixed order by a second kernel (deterministic).
#include <stdlib.h>

#include "geometry.cuh"
#include "mopa_scn.h"
#include "ptx.cuh"

namespace mopa {

// ------------------------------------------------------------------------------------------------ weight packing
// The contraction axis is cut into chunks of 32 channels (the last one 16 when c_in % 32 == 16); one (offset, chunk) pair
// is one pipeline step whose weights (chunk x c_out floats, fragment order) are one contiguous block: a single TMA
// bulk copy. Channel of (k-step s, fragment slot q) inside a chunk, and output channel of (n-tile p, column c): the
// permutations that let one float4 feed two k-steps and one thread own 4 contiguous output channels.
constexpr int kChunk = 32;
__host__ __device__ inline int frag_k_channel(int s, int q) { return 16 * (s >> 1) + 4 * (q & 3) + 2 * (s & 1) + (q >> 2); }
__host__ __device__ inline int frag_n_channel(int p, int c) { return 16 * (p >> 1) + 4 * (c >> 1) + 2 * (p & 1) + (c & 1); }

// packed[k][chunk][half][kstep][npair][lane][4]; half = hi / lo tf32 parts when split
__global__ void __launch_bounds__(256) k_pack_weights(const float *__restrict__ w, int volume, int n_in0, int n_out0,
                                                      int transpose, int flip, int split, float *__restrict__ packed,
                                                      int64_t total) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c_in = transpose ? n_out0 : n_in0, c_out = transpose ? n_in0 : n_out0;
    const int npair = c_out / 16, halves = split ? 2 : 1;
    const int64_t per_k = (int64_t)c_in * c_out * halves;
    const int k = (int)(idx / per_k);
    int64_t r = idx - (int64_t)k * per_k;
    const int chunk = (int)(r / ((int64_t)kChunk * c_out * halves));
    r -= (int64_t)chunk * kChunk * c_out * halves;
    const int kc = min(kChunk, c_in - chunk * kChunk);
    const int half = (int)(r / ((int64_t)kc * c_out));
    r -= (int64_t)half * kc * c_out;
    const int e4 = r & 3; r >>= 2;
    const int lane = r & 31; r >>= 5;
    const int u = (int)(r % npair);
    const int s = (int)(r / npair);
    const int g = lane >> 2, t = lane & 3;
    const int ci = chunk * kChunk + frag_k_channel(s, t + 4 * (e4 & 1));
    const int co = frag_n_channel(2 * u + (e4 >> 1), g);
    const int ks = flip ? volume - 1 - k : k;
    float v = transpose ? w[((int64_t)ks * n_in0 + co) * n_out0 + ci] : w[((int64_t)ks * n_in0 + ci) * n_out0 + co];
    float hi = __uint_as_float(to_tf32(v));
    packed[idx] = half == 0 ? hi : __uint_as_float(to_tf32(v - hi));
}

// ------------------------------------------------------------------------------------------------ gather-MMA kernel
// 8 consumer warps (each MT x 16 output rows) + 1 producer warp that streams the per-step weight blocks through a
// ring of shared-memory stages with TMA bulk copies. Consumers never meet at a block barrier: they wait on the stage's
// "full" mbarrier, multiply, and release it on the "empty" mbarrier; the gather of step s+1 is in flight while step s
// is multiplied.
constexpr int kConvWarps = 8;
constexpr int kConvThreads = (kConvWarps + 1) * 32;
constexpr int kMaxStages = 4;

// resident CTAs per SM the register allocation aims for (9 warps per CTA)
constexpr int conv_min_blocks(int mt, int np, bool split) { return split ? 1 : (mt * np <= 2 ? 3 : (mt * np <= 6 ? 2 : 1)); }

template <int MT, int NP, bool SPLIT>
__global__ void __launch_bounds__(kConvThreads, conv_min_blocks(MT, NP, SPLIT)) k_gather_mma(Gather gt, const float *__restrict__ in, int64_t ld_in,
                                                             float *__restrict__ out, int64_t ld_out,
                                                             const float *__restrict__ packed, int c_in, int n_stages) {
    constexpr int TM = kConvWarps * MT * 16;  // output rows per CTA
    constexpr int C_OUT = NP * 16;
    constexpr int H = SPLIT ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = gt.volume;
    float *sW = reinterpret_cast<float *>(smem_raw);  // n_stages x (kChunk * C_OUT * H) floats, 128-byte aligned
    const int stage_floats = kChunk * C_OUT * H;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(sW + (int64_t)n_stages * stage_floats);
    uint64_t *bar_empty = bar_full + kMaxStages;
    int32_t *sActive = reinterpret_cast<int32_t *>(bar_empty + kMaxStages);  // [32] active offsets, [32] = count
    int32_t *sT = sActive + 40;                                             // [K][TM]
    const int nchunk = (c_in + kChunk - 1) / kChunk;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t row0 = (int64_t)blockIdx.x * TM;

    // ---- stage the tile's slice of the gather table; find the offsets that touch this tile
    if (tid < 40) sActive[tid] = 0;
    if (tid == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, kConvWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    for (int idx = tid; idx < K * TM; idx += kConvThreads) {
        const int k = idx / TM, r = idx - k * TM;
        const int64_t row = row0 + r;
        int v = row < gt.n_out ? gather_lookup(gt, k, row) : -1;
        sT[idx] = v;
        if (v >= 0) sActive[k] = 1;  // benign race
    }
    __syncthreads();
    if (warp == 0) {
        const bool on = lane < K && sActive[lane] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        __syncwarp();
        if (on) sActive[__popc(m & ((1u << lane) - 1))] = lane;
        if (lane == 0) sActive[32] = __popc(m);
    }
    __syncthreads();  // last block-wide barrier: the producer warp leaves early
    const int n_steps = sActive[32] * nchunk;

    if (warp == kConvWarps) {  // ===== producer: one lane feeds the ring =====
        if (lane == 0) {
            for (int step = 0; step < n_steps; ++step) {
                const int slot = step % n_stages, use = step / n_stages;
                const int a = step / nchunk, ch = step - a * nchunk;
                const int kc = min(kChunk, c_in - ch * kChunk);
                const uint32_t bytes = (uint32_t)kc * C_OUT * H * 4;
                const float *src = packed + ((int64_t)sActive[a] * c_in + ch * kChunk) * C_OUT * H;
                mbar_wait(bar_empty + slot, (use & 1) ^ 1);
                mbar_expect_tx(bar_full + slot, bytes);
                tma_load_1d(sW + (int64_t)slot * stage_floats, src, bytes, bar_full + slot);
            }
        }
        return;
    }

    float acc[MT][2 * NP][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int p = 0; p < 2 * NP; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[m][p][e] = 0.f;

    const int warp_row = warp * MT * 16;
    const float *abase = in + 4 * t;

    // gather the A rows of one step (up to 32 channels = two float4 per row half); rows without a rule stay zero
    auto load_a = [&](int step, float4 (&x)[MT][2][2], bool &any) {
        any = false;
        if (step >= n_steps) return;
        const int a = step / nchunk, ch = step - a * nchunk;
        const int k = sActive[a];
        const int nj = min(kChunk, c_in - ch * kChunk) / 16;
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int idx = sT[k * TM + warp_row + m * 16 + g + 8 * h];
                any |= idx >= 0;
                const float *p = abase + (int64_t)idx * ld_in + ch * kChunk;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    x[m][h][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx >= 0 && j < nj) x[m][h][j] = __ldg(reinterpret_cast<const float4 *>(p + 16 * j));
                }
            }
        any = __any_sync(0xffffffffu, any);
    };

    auto compute = [&](int step, const float4 (&x)[MT][2][2], bool any) {
        const int slot = step % n_stages, use = step / n_stages;
        mbar_wait(bar_full + slot, use & 1);
        if (any) {
            const int ch = step % nchunk;
            const int kc = min(kChunk, c_in - ch * kChunk);
            const float4 *w = reinterpret_cast<const float4 *>(sW + (int64_t)slot * stage_floats) + lane;
            const int lo_off = (kc / 8) * NP * 32;  // float4 offset of the lo half (split mode)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j * 16 >= kc) break;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    uint32_t a_hi[MT][4], a_lo[MT][4];
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const float v0 = e ? x[m][0][j].z : x[m][0][j].x, v2 = e ? x[m][0][j].w : x[m][0][j].y;
                        const float v1 = e ? x[m][1][j].z : x[m][1][j].x, v3 = e ? x[m][1][j].w : x[m][1][j].y;
                        a_hi[m][0] = to_tf32(v0); a_hi[m][1] = to_tf32(v1); a_hi[m][2] = to_tf32(v2); a_hi[m][3] = to_tf32(v3);
                        if (SPLIT) {
                            a_lo[m][0] = to_tf32(v0 - __uint_as_float(a_hi[m][0]));
                            a_lo[m][1] = to_tf32(v1 - __uint_as_float(a_hi[m][1]));
                            a_lo[m][2] = to_tf32(v2 - __uint_as_float(a_hi[m][2]));
                            a_lo[m][3] = to_tf32(v3 - __uint_as_float(a_hi[m][3]));
                        }
                    }
                    const float4 *wf = w + (2 * j + e) * NP * 32;
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        const float4 b = wf[u * 32];
                        if (SPLIT) {
                            const float4 bl = wf[lo_off + u * 32];
#pragma unroll
                            for (int m = 0; m < MT; ++m) {
                                mma_tf32(acc[m][2 * u], a_lo[m], __float_as_uint(b.x), __float_as_uint(b.y));
                                mma_tf32(acc[m][2 * u + 1], a_lo[m], __float_as_uint(b.z), __float_as_uint(b.w));
                                mma_tf32(acc[m][2 * u], a_hi[m], __float_as_uint(bl.x), __float_as_uint(bl.y));
                                mma_tf32(acc[m][2 * u + 1], a_hi[m], __float_as_uint(bl.z), __float_as_uint(bl.w));
                            }
                        }
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            mma_tf32(acc[m][2 * u], a_hi[m], __float_as_uint(b.x), __float_as_uint(b.y));
                            mma_tf32(acc[m][2 * u + 1], a_hi[m], __float_as_uint(b.z), __float_as_uint(b.w));
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + slot);
    };

    // two register sets ping-pong: while set A is multiplied, set B's gathers are in flight
    float4 xa[MT][2][2], xb[MT][2][2];
    bool any_a = false, any_b = false;
    load_a(0, xa, any_a);
    for (int step = 0; step < n_steps; step += 2) {
        load_a(step + 1, xb, any_b);
        compute(step, xa, any_a);
        if (step + 1 < n_steps) {
            load_a(step + 2, xa, any_a);
            compute(step + 1, xb, any_b);
        }
    }

    // ---- each output row is written once: thread (g, t) owns channels 16u + 4t .. +3 of rows g and g + 8
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int64_t r_lo = row0 + warp_row + m * 16 + g, r_hi = r_lo + 8;
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            if (r_lo < gt.n_out) {
                float4 *dst = reinterpret_cast<float4 *>(out + r_lo * ld_out + 16 * u + 4 * t);
                float4 v = make_float4(acc[m][2 * u][0], acc[m][2 * u][1], acc[m][2 * u + 1][0], acc[m][2 * u + 1][1]);
                if (gt.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                *dst = v;
            }
            if (r_hi < gt.n_out) {
                float4 *dst = reinterpret_cast<float4 *>(out + r_hi * ld_out + 16 * u + 4 * t);
                float4 v = make_float4(acc[m][2 * u][2], acc[m][2 * u][3], acc[m][2 * u + 1][2], acc[m][2 * u + 1][3]);
                if (gt.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                *dst = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ small-c_in conv
// The 1 -> 16 input layer (scn_unet.py:27): four threads per output row, four output channels each (a warp writes 8
// rows = 512 contiguous bytes), weights in shared memory, all table reads of a 9-offset group in flight before the
// first feature read. out[o][co] = sum_k sum_ci in[T[k][o]][ci] * W[k][ci][co]. HBM-bound: table 4 K V + out 4 V Cout.
__global__ void __launch_bounds__(256) k_conv_smallcin(Gather gt, const float *__restrict__ in, int64_t ld_in,
                                                       float *__restrict__ out, int64_t ld_out,
                                                       const float *__restrict__ w, int c_in, int c_out) {
    extern __shared__ __align__(16) float sw[];  // [K][c_in][c_out]
    for (int i = threadIdx.x; i < gt.volume * c_in * c_out; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int64_t o = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2);
    const int cb = blockIdx.y * 16 + 4 * (threadIdx.x & 3);
    if (o >= gt.n_out) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < gt.volume; k0 += 9) {
        int idx[9];
#pragma unroll
        for (int u = 0; u < 9; ++u) idx[u] = k0 + u < gt.volume ? gather_lookup(gt, k0 + u, o) : -1;
        for (int ci = 0; ci < c_in; ++ci) {
            float x[9];
            // (guarded loads: absent neighbours -- ~85 % of them -- are not fetched; a branch-free variant with clamped rows
            // was measured slower, 49 vs 43 us)
#pragma unroll
            for (int u = 0; u < 9; ++u) x[u] = idx[u] >= 0 ? __ldg(in + (int64_t)idx[u] * ld_in + ci) : 0.f;
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                if (k0 + u < gt.volume) {
                    const float4 wv = *reinterpret_cast<const float4 *>(sw + ((k0 + u) * c_in + ci) * c_out + cb);
                    acc.x = fmaf(x[u], wv.x, acc.x); acc.y = fmaf(x[u], wv.y, acc.y);
                    acc.z = fmaf(x[u], wv.z, acc.z); acc.w = fmaf(x[u], wv.w, acc.w);
                }
            }
        }
    }
    float4 *dst = reinterpret_cast<float4 *>(out + o * ld_out + cb);
    if (gt.accumulate) { const float4 e = *dst; acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w; }
    *dst = acc;
}

// ------------------------------------------------------------------------------------------------ generic SIMT conv
// Shapes the MMA path does not cover (channel counts that are not multiples of 16: the 1 -> 16 input layer).
// out[o][co] = sum_k sum_ci in[T[k][o]][ci] * Wm[k][ci][co], Wm read from the unpacked (volume, n_in0, n_out0) weight.
__global__ void __launch_bounds__(256) k_conv_generic(Gather gt, const float *__restrict__ in, int64_t ld_in,
                                                      float *__restrict__ out, int64_t ld_out,
                                                      const float *__restrict__ w, int n_in0, int n_out0,
                                                      int transpose, int flip) {
    const int c_in = transpose ? n_out0 : n_in0, c_out = transpose ? n_in0 : n_out0;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= gt.n_out * c_out) return;
    const int64_t o = idx / c_out;
    const int co = (int)(idx - o * c_out);
    float acc = 0.f;
    for (int k = 0; k < gt.volume; ++k) {
        const int i = gather_lookup(gt, k, o);
        if (i < 0) continue;
        const int ks = flip ? gt.volume - 1 - k : k;
        const float *row = in + (int64_t)i * ld_in;
        for (int ci = 0; ci < c_in; ++ci) {
            const float wv = transpose ? __ldg(w + ((int64_t)ks * n_in0 + co) * n_out0 + ci)
                                       : __ldg(w + ((int64_t)ks * n_in0 + ci) * n_out0 + co);
            acc = fmaf(__ldg(row + ci), wv, acc);
        }
    }
    out[o * ld_out + co] = gt.accumulate ? out[o * ld_out + co] + acc : acc;
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int kDwThreads = 256;
constexpr int kDwSub = 1024;  // rows compacted per pass

__device__ __forceinline__ int block_exclusive_scan_256(int v, int *sh /*>= 9 ints*/) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
    }
    __syncthreads();
    if (lane == 31) sh[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < 8 ? sh[lane] : 0, si = s;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            int x = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += x;
        }
        if (lane < 8) sh[lane] = si - s;
        if (lane == 7) sh[8] = si;
    }
    __syncthreads();
    return incl - v + sh[w];
}

// grid (nchunks, volume). Warp w: block (bm, bn) of 32x32 output tiles b = w % nblk (+ 8, 16, ... up to BPW), rule
// k-steps ks = w / nblk mod WK. partial[(k * nchunks + chunk) * WK + wk][n_in][n_out].
template <int BPW, bool SPLIT>
__global__ void __launch_bounds__(kDwThreads) k_dw_mma(Gather gt, const float *__restrict__ in, int64_t ld_in,
                                                       const float *__restrict__ dout, int64_t ld_dout, int n_in,
                                                       int n_out, int rows_per_chunk, int RT, int WK,
                                                       float *__restrict__ partial, int kmap_center) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // offset order: the centre offset of a submanifold filter carries V rules (6x the others) -> schedule it first
    int k = blockIdx.y;
    if (kmap_center >= 0) k = (k == 0) ? kmap_center : (k <= kmap_center ? k - 1 : k);
    const int chunk = blockIdx.x, nchunks = gridDim.x;
    const int ldA = n_in + 8, ldB = n_out + 8;
    int32_t *sIn = reinterpret_cast<int32_t *>(smem_raw);  // [kDwSub]
    int32_t *sOut = sIn + kDwSub;                          // [kDwSub]
    int *sScan = sOut + kDwSub;                            // [16]
    float *sA = reinterpret_cast<float *>(sScan + 16);     // 2 stages x RT x ldA
    float *sB = sA + 2 * RT * ldA;                         // 2 stages x RT x ldB

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nbm = (n_in + 31) / 32, nbn = (n_out + 31) / 32, nblk = nbm * nbn;
    const int wk = (warp / nblk) % WK;           // which k-steps of a rule tile this warp multiplies
    const bool warp_active = warp < nblk * WK || nblk >= 8;
    const int first_blk = nblk >= 8 ? warp : warp % nblk;
    const int blk_stride = nblk >= 8 ? 8 : nblk * 1000;  // blocks of this warp: first_blk, +8, +16, ...

    float acc[BPW][2][4][4];
#pragma unroll
    for (int b = 0; b < BPW; ++b)
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[b][mi][ni][e] = 0.f;

    const int64_t r_begin = (int64_t)chunk * rows_per_chunk;
    const int64_t r_end = min(gt.n_out, r_begin + rows_per_chunk);
    const int cpr_a = n_in / 4, cpr_b = n_out / 4;  // 16-byte pieces per row

    for (int64_t sub = r_begin; sub < r_end; sub += kDwSub) {
        // ---- ordered compaction of this sub-chunk's rules (thread owns 4 consecutive rows)
        int v[4], cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t row = sub + tid * 4 + q;
            v[q] = row < r_end ? gather_lookup(gt, k, row) : -1;
            cnt += v[q] >= 0;
        }
        __syncthreads();  // previous pass finished reading sIn/sOut and the staging buffers
        int pos = block_exclusive_scan_256(cnt, sScan);
        const int n_rules = sScan[8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (v[q] >= 0) {
                sIn[pos] = v[q];
                sOut[pos] = (int)(sub - r_begin) + tid * 4 + q;  // relative to r_begin (fits int)
                ++pos;
            }
        __syncthreads();
        const int n_tiles = (n_rules + RT - 1) / RT;

        auto stage = [&](int tile, int buf) {
            float *a = sA + buf * RT * ldA, *b = sB + buf * RT * ldB;
            const int base = tile * RT;
            for (int i = tid; i < RT * cpr_a; i += kDwThreads) {
                const int r = i / cpr_a, c = i - r * cpr_a;
                const bool ok = base + r < n_rules;
                const float *src = ok ? in + (int64_t)sIn[base + r] * ld_in + 4 * c : in;
                cp_async16(a + r * ldA + 4 * c, src, ok);
            }
            for (int i = tid; i < RT * cpr_b; i += kDwThreads) {
                const int r = i / cpr_b, c = i - r * cpr_b;
                const bool ok = base + r < n_rules;
                const float *src = ok ? dout + (r_begin + sOut[base + r]) * ld_dout + 4 * c : dout;
                cp_async16(b + r * ldB + 4 * c, src, ok);
            }
            cp_async_commit();
        };

        if (n_tiles > 0) stage(0, 0);
        for (int tile = 0; tile < n_tiles; ++tile) {
            cp_async_wait<0>();
            __syncthreads();
            if (tile + 1 < n_tiles) stage(tile + 1, (tile + 1) & 1);
            if (!warp_active) continue;
            const float *a = sA + (tile & 1) * RT * ldA, *b = sB + (tile & 1) * RT * ldB;
            for (int ks = wk; ks < RT / 8; ks += WK) {
                const float *a0 = a + (ks * 8 + t) * ldA, *a1 = a0 + 4 * ldA;
                const float *b0 = b + (ks * 8 + t) * ldB, *b1 = b0 + 4 * ldB;
#pragma unroll
                for (int bb = 0; bb < BPW; ++bb) {
                    const int blk = first_blk + bb * blk_stride;
                    if (blk >= nblk) break;
                    const int cm = (blk / nbn) * 32, cn = (blk % nbn) * 32;
                    uint32_t ah[2][4], al[2][4];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        const int c = cm + mi * 16 + g;
                        const bool okm = cm + mi * 16 < n_in;
                        const float f0 = okm ? a0[c] : 0.f, f1 = okm ? a0[c + 8] : 0.f;
                        const float f2 = okm ? a1[c] : 0.f, f3 = okm ? a1[c + 8] : 0.f;
                        ah[mi][0] = to_tf32(f0); ah[mi][1] = to_tf32(f1); ah[mi][2] = to_tf32(f2); ah[mi][3] = to_tf32(f3);
                        if (SPLIT) {
                            al[mi][0] = to_tf32(f0 - __uint_as_float(ah[mi][0]));
                            al[mi][1] = to_tf32(f1 - __uint_as_float(ah[mi][1]));
                            al[mi][2] = to_tf32(f2 - __uint_as_float(ah[mi][2]));
                            al[mi][3] = to_tf32(f3 - __uint_as_float(ah[mi][3]));
                        }
                    }
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {
                        const int c = cn + ni * 8 + g;
                        const bool okn = cn + ni * 8 < n_out;
                        const float g0 = okn ? b0[c] : 0.f, g1 = okn ? b1[c] : 0.f;
                        const uint32_t bh0 = to_tf32(g0), bh1 = to_tf32(g1);
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi) {
                            if (SPLIT) {
                                const uint32_t bl0 = to_tf32(g0 - __uint_as_float(bh0));
                                const uint32_t bl1 = to_tf32(g1 - __uint_as_float(bh1));
                                mma_tf32(acc[bb][mi][ni], al[mi], bh0, bh1);
                                mma_tf32(acc[bb][mi][ni], ah[mi], bl0, bl1);
                            }
                            mma_tf32(acc[bb][mi][ni], ah[mi], bh0, bh1);
                        }
                    }
                }
            }
        }
    }

    // ---- per-(chunk, wk) partial; warps that own no block (nblk * WK < 8) write nothing
    if (!warp_active) return;
    float *dst = partial + (((int64_t)k * nchunks + chunk) * WK + wk) * n_in * n_out;
#pragma unroll
    for (int bb = 0; bb < BPW; ++bb) {
        const int blk = first_blk + bb * blk_stride;
        if (blk >= nblk) break;
        const int cm = (blk / nbn) * 32, cn = (blk % nbn) * 32;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int r = cm + mi * 16 + g, c = cn + ni * 8 + 2 * t;
                if (cm + mi * 16 < n_in && cn + ni * 8 < n_out) {
                    *reinterpret_cast<float2 *>(dst + (int64_t)r * n_out + c) = make_float2(acc[bb][mi][ni][0], acc[bb][mi][ni][1]);
                    *reinterpret_cast<float2 *>(dst + (int64_t)(r + 8) * n_out + c) = make_float2(acc[bb][mi][ni][2], acc[bb][mi][ni][3]);
                }
            }
    }
}

// generic dW for shapes off the MMA path: grid (nchunks, volume, n_in); block: threads over rows, 16 output channels each
__global__ void __launch_bounds__(256) k_dw_generic(Gather gt, const float *__restrict__ in, int64_t ld_in,
                                                    const float *__restrict__ dout, int64_t ld_dout, int n_in,
                                                    int n_out, int rows_per_chunk, float *__restrict__ partial) {
    const int chunk = blockIdx.x, nchunks = gridDim.x, k = blockIdx.y, ci = blockIdx.z;
    const int64_t r_begin = (int64_t)chunk * rows_per_chunk, r_end = min(gt.n_out, r_begin + rows_per_chunk);
    __shared__ float red[8][16];
    for (int c0 = 0; c0 < n_out; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        for (int64_t o = r_begin + threadIdx.x; o < r_end; o += blockDim.x) {
            const int i = gather_lookup(gt, k, o);
            if (i < 0) continue;
            const float x = __ldg(in + (int64_t)i * ld_in + ci);
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c0 + c < n_out) acc[c] = fmaf(x, __ldg(dout + o * ld_dout + c0 + c), acc[c]);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float v = acc[c];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
        }
        __syncthreads();
        if (threadIdx.x < 16 && c0 + threadIdx.x < n_out) {
            float v = 0.f;
            for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
            partial[(((int64_t)k * nchunks + chunk) * n_in + ci) * n_out + c0 + threadIdx.x] = v;
        }
        __syncthreads();
    }
}

// d_weight of a convolution with very few input planes (the network's first layer, 1 -> 16; TF32 mode), where the rows are
// the only long dimension: dW[k][ci][co] = sum_o x[in(k, o)][ci] dy[o][co] = (Xg^T dY)[k][co] with Xg[o][k] = x[in(k, o)][ci],
// a (32 x rows) x (rows x n_out) product. mma.sync m16n8k8 takes 8 rows per step: a lane gathers its eight A values
// (offset m = g, g + 8, g + 16, g + 24; rows t, t + 4) scalar by scalar straight into the fragment layout, dy is read once
// per row instead of once per offset and row (k_dw_generic: 27 passes over dy, 83 us at 231k rows; this kernel: ~20 us).
// grid (blocks, 1, n_in); every warp strides over 8-row chunks; warps, then blocks, are summed in index order.
template <int NT8>
__global__ void __launch_bounds__(256) k_dw_rows_mma(Gather gt, const float *__restrict__ in, int64_t ld_in,
                                                     const float *__restrict__ dout, int64_t ld_dout, int n_in,
                                                     float *__restrict__ partial) {
    constexpr int n_out = 8 * NT8;
    __shared__ float red[8][32][n_out + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3, ci = blockIdx.z;
    const int K = gt.volume;
    float acc[2][NT8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < NT8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;
    const int64_t n_chunks = (gt.n_out + 7) >> 3, stride = (int64_t)gridDim.x * 8;
#pragma unroll 2
    for (int64_t c = (int64_t)blockIdx.x * 8 + warp; c < n_chunks; c += stride) {
        const int64_t o0 = (c << 3) + t, o1 = o0 + 4;
        // Branch-free: all eight rule lookups first, then all eight gathers (clamped addresses, results selected afterwards).
        // With `if (valid) { i = lookup; if (i >= 0) x = load; }` per element the compiler emitted eight reconvergence
        // regions, i.e. eight index -> value round trips one after the other per chunk (55 us instead of ~15).
        int src[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = 16 * mt + g + 8 * (e & 1);
                const int64_t o = (e & 2) ? o1 : o0;
                const bool ok = m < K && o < gt.n_out;
                const int mc = m < K ? m : K - 1;
                const int64_t oc = o < gt.n_out ? o : gt.n_out - 1;
                int i;
                if (gt.table) {  // (uniform)
                    i = __ldg(gt.table + (int64_t)mc * gt.ld + oc);
                } else {
                    const int kx = __ldg(gt.kidx + oc), pr = __ldg(gt.parent + oc);
                    i = kx == m ? pr : -1;
                }
                src[mt][e] = ok ? i : -1;
            }
        }
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = src[mt][e];
                const float x = __ldg(in + (int64_t)(i < 0 ? 0 : i) * ld_in + ci);
                a[mt][e] = to_tf32(i < 0 ? 0.f : x);
            }
        }
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            const uint32_t b0 = to_tf32(o0 < gt.n_out ? __ldg(dout + o0 * ld_dout + 8 * j + g) : 0.f);
            const uint32_t b1 = to_tf32(o1 < gt.n_out ? __ldg(dout + o1 * ld_dout + 8 * j + g) : 0.f);
            mma_tf32(acc[0][j], a[0], b0, b1);
            mma_tf32(acc[1][j], a[1], b0, b1);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            red[warp][16 * mt + g][8 * j + 2 * t] = acc[mt][j][0];
            red[warp][16 * mt + g][8 * j + 2 * t + 1] = acc[mt][j][1];
            red[warp][16 * mt + g + 8][8 * j + 2 * t] = acc[mt][j][2];
            red[warp][16 * mt + g + 8][8 * j + 2 * t + 1] = acc[mt][j][3];
        }
    __syncthreads();
    for (int idx = threadIdx.x; idx < K * n_out; idx += blockDim.x) {
        const int k = idx / n_out, co = idx - k * n_out;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][k][co];
        partial[(((int64_t)k * gridDim.x + blockIdx.x) * n_in + ci) * n_out + co] = v;
    }
}

// d_weight[k][e] = sum over the k-th group of `per_k` partial slices, in index order (deterministic)
// grid (element blocks of 32, offsets), block (32 elements, 8 slice groups): thread (x, y) adds slices y, y + 8, ... in order,
// the 8 sums are combined in y order: a fixed tree, with 8x shorter load chains than one thread per element.
__global__ void __launch_bounds__(256) k_dw_reduce(const float *__restrict__ partial, int per_k, int64_t mat,
                                                   float *__restrict__ dw) {
    __shared__ float red[8][33];
    const int k = blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * 32 + threadIdx.x;
    float s0 = 0.f, s1 = 0.f;
    if (e < mat) {
        const float *p = partial + (int64_t)k * per_k * mat + e;
        int c = threadIdx.y;
        for (; c + 8 < per_k; c += 16) {
            s0 += p[(int64_t)c * mat];
            s1 += p[(int64_t)(c + 8) * mat];
        }
        if (c < per_k) s0 += p[(int64_t)c * mat];
    }
    red[threadIdx.y][threadIdx.x] = s0 + s1;
    __syncthreads();
    if (threadIdx.y == 0 && e < mat) {
        float s = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x];
        dw[(int64_t)k * mat + e] = s;
    }
}

// ------------------------------------------------------------------------------------------------ host dispatch
bool conv_tc_enabled() {  // MOPA_SCN_NO_TC=1 keeps every layer on the mma.sync kernel (A/B measurements)
    static const bool on = [] {
        const char *e = getenv("MOPA_SCN_NO_TC");
        return !(e && e[0] == '1');
    }();
    return on;
}
bool conv_uses_packed(int c_in, int c_out) {
    if (c_in % 16 || c_out % 16 || c_in < 16 || c_out < 16) return false;
    const int np = c_out / 16;
    return np <= 8 || np == 10 || np == 12;
}
static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

template <int MT, int NP, bool SPLIT>
static int launch_gather_mma(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out,
                             const float *packed, int c_in, cudaStream_t s) {
    constexpr int TM = kConvWarps * MT * 16;
    const size_t stage = (size_t)kChunk * NP * 16 * (SPLIT ? 2 : 1) * 4;
    const size_t fixed = (size_t)2 * kMaxStages * 8 + 40 * 4 + (size_t)gt.volume * TM * 4;
    int stages = kMaxStages;
    while (stages > 2 && stages * stage + fixed > 72 * 1024) --stages;  // keep three CTAs per SM where possible
    const int nsteps_max = gt.volume * (int)ceil_div(c_in, kChunk);
    if (stages > nsteps_max) stages = nsteps_max < 1 ? 1 : nsteps_max;
    const size_t smem = stages * stage + fixed;
    auto kern = k_gather_mma<MT, NP, SPLIT>;
    static std::atomic<uint64_t> configured{0};  // per template instantiation, one bit per device
    MOPA_TRY(once_per_device(configured, [&] {
        MOPA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MOPA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        return 0;
    }));
    const unsigned grid = (unsigned)ceil_div(gt.n_out, TM);
    kern<<<grid, kConvThreads, smem, s>>>(gt, in, ld_in, out, ld_out, packed, c_in, stages);
    MOPA_LAUNCHED();
    return 0;
}

template <int NP>
static int dispatch_np(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out,
                       const float *packed, int c_in, int split, cudaStream_t s) {
    // two m-tiles per warp (B fragments re-used twice) while the accumulators fit and the grid still fills the chip
    const bool wide_tile = false;  // measured: the second m-tile costs more occupancy than the B re-use saves
    if (split) return launch_gather_mma<1, NP, true>(gt, in, ld_in, out, ld_out, packed, c_in, s);
    if (wide_tile) return launch_gather_mma<(NP <= 4 ? 2 : 1), NP, false>(gt, in, ld_in, out, ld_out, packed, c_in, s);
    return launch_gather_mma<1, NP, false>(gt, in, ld_in, out, ld_out, packed, c_in, s);
}

// out (n_out rows, c_out) = gather-conv of in with the (possibly transposed / flipped) weights
int conv_apply(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out, const float *weight,
               const float *packed, int n_in0, int n_out0, int transpose, int flip, int precision, cudaStream_t s,
               double *stats, bool *stats_done, const TcBnBwd *bn) {
    // stats: per-column sum / sum of squares of the output rows are added there IF the tcgen05 kernel runs this op
    // (*stats_done tells the caller); the BatchNorm that follows then skips its own statistics pass
    if (stats_done) *stats_done = false;
    if (gt.n_out == 0) return 0;
    const int c_in = transpose ? n_out0 : n_in0, c_out = transpose ? n_in0 : n_out0;
    const int prof = prof_begin((transpose ? 20 : 10) + gt.op, &gt, c_in, c_out, gt.n_out, s);
    struct ProfEnd {
        int idx; cudaStream_t s;
        ~ProfEnd() { prof_end(idx, s); }
    } prof_guard{prof, s};
    const bool fast = conv_uses_packed(c_in, c_out) && packed && aligned16(in) && aligned16(out) && ld_in % 4 == 0 &&
                      ld_out % 4 == 0 && aligned16(packed);
    if (!fast && !transpose && !flip && c_in <= 8 && c_out % 16 == 0 && (size_t)gt.volume * c_in * c_out * 4 <= 40 * 1024 &&
        weight && aligned16(out) && ld_out % 4 == 0) {
        dim3 grid((unsigned)ceil_div(gt.n_out, 64), c_out / 16);
        k_conv_smallcin<<<grid, 256, (size_t)gt.volume * c_in * c_out * 4, s>>>(gt, in, ld_in, out, ld_out, weight, c_in, c_out);
        MOPA_LAUNCHED();
        return 0;
    }
    if (!fast) {
        MOPA_CHECK(weight != nullptr, "generic conv path needs the unpacked weight");
        const int64_t total = gt.n_out * c_out;
        k_conv_generic<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(gt, in, ld_in, out, ld_out, weight, n_in0, n_out0,
                                                                      transpose, flip);
        MOPA_LAUNCHED();
        return 0;
    }
    const int split = precision == MOPA_SCN_PREC_FP32;
    if (!split && conv_tc_enabled() && conv_tc_supported(c_in, c_out))
    {
        if (stats_done) *stats_done = stats != nullptr;
        return conv_apply_tc(gt, in, ld_in, out, ld_out, packed, c_in, c_out, stats, s, stats ? bn : nullptr);
    }
    switch (c_out / 16) {
        case 1: return dispatch_np<1>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 2: return dispatch_np<2>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 3: return dispatch_np<3>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 4: return dispatch_np<4>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 5: return dispatch_np<5>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 6: return dispatch_np<6>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 7: return dispatch_np<7>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 8: return dispatch_np<8>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 10: return dispatch_np<10>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
        case 12: return dispatch_np<12>(gt, in, ld_in, out, ld_out, packed, c_in, split, s);
    }
    MOPA_FAIL("unreachable conv shape");
}

struct DwPlan {
    bool mma;
    int nchunks, rows_per_chunk, RT, WK, BPW, per_k;
    size_t partial_bytes;
};
static DwPlan dw_plan(int volume, int n_in, int n_out, int64_t n_rows) {
    DwPlan p{};
    p.mma = n_in % 16 == 0 && n_out % 16 == 0 && n_in >= 16 && n_out >= 16 &&
            ((n_in + 31) / 32) * ((n_out + 31) / 32) <= 32;
    int nch = (int)ceil_div(n_rows > 0 ? n_rows : 1, 512);
    p.nchunks = nch < 1 ? 1 : (nch > 32 ? 32 : nch);
    p.rows_per_chunk = (int)round_up(ceil_div(n_rows > 0 ? n_rows : 1, p.nchunks), 4);
    if (p.mma) {
        const int nblk = ((n_in + 31) / 32) * ((n_out + 31) / 32);
        p.RT = (n_in + n_out <= 128) ? 64 : 32;
        p.WK = nblk >= 8 ? 1 : 8 / nblk;
        if (p.WK > p.RT / 8) p.WK = p.RT / 8;
        p.BPW = nblk >= 8 ? (nblk + 7) / 8 : 1;
    } else {
        p.RT = 0; p.WK = 1; p.BPW = 1;
    }
    p.per_k = p.nchunks * p.WK;
    p.partial_bytes = (size_t)volume * p.per_k * n_in * n_out * 4;
    return p;
}

// blocks of k_dw_rows_mma: >= 6 chunks of 8 rows per warp, at most two blocks per SM of the reference device (a function of
// the row count only: the summation tree must not depend on the caller's workspace or on the device at hand)
static int dw_rows_blocks(int64_t n_rows) {
    int64_t b = ceil_div(ceil_div(n_rows > 0 ? n_rows : 1, 8), 8 * 6);
    return (int)(b < 1 ? 1 : (b > 2 * kNumSMs ? 2 * kNumSMs : b));
}
static bool dw_rows_supported(int volume, int n_in, int n_out) { return n_in <= 4 && (n_out == 16 || n_out == 32) && volume <= 32; }

size_t dw_workspace_bytes(int volume, int n_in, int n_out, int64_t n_rows) {
    size_t b = dw_plan(volume, n_in, n_out, n_rows).partial_bytes;
    if (dw_rows_supported(volume, n_in, n_out)) {
        const size_t t = (size_t)volume * dw_rows_blocks(n_rows) * n_in * n_out * 4;
        if (t > b) b = t;
    }
    if (dw_tc_supported(n_in, n_out)) {
        const size_t t = dw_tc_workspace_bytes(volume, n_in, n_out, n_rows);
        if (t > b) b = t;
    }
    return b + 256;
}

template <int BPW, bool SPLIT>
static int launch_dw(const Gather &gt, const float *in, int64_t ld_in, const float *dout, int64_t ld_dout, int n_in,
                     int n_out, const DwPlan &p, float *partial, int center, cudaStream_t s) {
    const size_t smem = (size_t)(2 * kDwSub + 16) * 4 + (size_t)2 * p.RT * (n_in + 8 + n_out + 8) * 4;
    auto kern = k_dw_mma<BPW, SPLIT>;
    static std::atomic<uint64_t> configured{0};
    MOPA_TRY(once_per_device(configured, [&] {
        MOPA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MOPA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        return 0;
    }));
    dim3 grid(p.nchunks, gt.volume);
    kern<<<grid, kDwThreads, smem, s>>>(gt, in, ld_in, dout, ld_dout, n_in, n_out, p.rows_per_chunk, p.RT, p.WK, partial,
                                        center);
    MOPA_LAUNCHED();
    return 0;
}

// d_weight (volume, n_in, n_out) = sum over rules in[in_row]^T dout[out_row]; gt indexes by OUTPUT row
int conv_dweight(const Gather &gt, const float *in, int64_t ld_in, const float *dout, int64_t ld_dout, float *dw,
                 int n_in, int n_out, int precision, void *workspace, size_t workspace_bytes, cudaStream_t s) {
    const int64_t mat = (int64_t)n_in * n_out;
    if (gt.n_out == 0) {
        MOPA_CUDA(cudaMemsetAsync(dw, 0, (size_t)gt.volume * mat * 4, s));
        return 0;
    }
    DwPlan p = dw_plan(gt.volume, n_in, n_out, gt.n_out);
    MOPA_CHECK(workspace && workspace_bytes >= p.partial_bytes, "backward workspace too small");
    const int prof = prof_begin(30 + gt.op, &gt, n_in, n_out, gt.n_out, s);
    struct ProfEnd {
        int idx; cudaStream_t s;
        ~ProfEnd() { prof_end(idx, s); }
    } prof_guard{prof, s};
    float *partial = reinterpret_cast<float *>(workspace);
    const bool fast = p.mma && aligned16(in) && aligned16(dout) && ld_in % 4 == 0 && ld_dout % 4 == 0;
    if (precision == MOPA_SCN_PREC_TF32 && dw_tc_enabled() && dw_tc_supported(n_in, n_out) && aligned16(in) &&
        aligned16(dout) && ld_in % 4 == 0 && ld_dout % 4 == 0) {
        MOPA_CHECK(workspace_bytes >= dw_tc_workspace_bytes(gt.volume, n_in, n_out, gt.n_out), "backward workspace too small");
        return conv_dweight_tc(gt, in, ld_in, dout, ld_dout, dw, n_in, n_out, partial, s);
    }
    if (fast) {
        const int center = gt.volume == 27 && gt.table ? 13 : -1;
        const bool split = precision == MOPA_SCN_PREC_FP32;
#define MOPA_DW(B)                                                                                            \
    (split ? launch_dw<B, true>(gt, in, ld_in, dout, ld_dout, n_in, n_out, p, partial, center, s)              \
           : launch_dw<B, false>(gt, in, ld_in, dout, ld_dout, n_in, n_out, p, partial, center, s))
        switch (p.BPW) {
            case 1: MOPA_TRY(MOPA_DW(1)); break;
            case 2: MOPA_TRY(MOPA_DW(2)); break;
            case 3: MOPA_TRY(MOPA_DW(3)); break;
            case 4: MOPA_TRY(MOPA_DW(4)); break;
            default: MOPA_FAIL("d_weight: channel counts too large");
        }
#undef MOPA_DW
    } else if (precision == MOPA_SCN_PREC_TF32 && dw_rows_supported(gt.volume, n_in, n_out)) {
        // few input planes (the first layer): rows as the GEMM's K dimension
        const int blocks = dw_rows_blocks(gt.n_out);
        MOPA_CHECK(workspace_bytes >= (size_t)gt.volume * blocks * mat * 4, "backward workspace too small");
        p.per_k = blocks;
        dim3 grid(blocks, 1, n_in);
        if (n_out == 16) k_dw_rows_mma<2><<<grid, 256, 0, s>>>(gt, in, ld_in, dout, ld_dout, n_in, partial);
        else k_dw_rows_mma<4><<<grid, 256, 0, s>>>(gt, in, ld_in, dout, ld_dout, n_in, partial);
        MOPA_LAUNCHED();
    } else {
        p.WK = 1;
        p.per_k = p.nchunks;
        dim3 grid(p.nchunks, gt.volume, n_in);
        k_dw_generic<<<grid, 256, 0, s>>>(gt, in, ld_in, dout, ld_dout, n_in, n_out, p.rows_per_chunk, partial);
        MOPA_LAUNCHED();
    }
    k_dw_reduce<<<dim3((unsigned)ceil_div(mat, 32), gt.volume), dim3(32, 8), 0, s>>>(partial, p.per_k, mat, dw);
    MOPA_LAUNCHED();
    return 0;
}

bool conv_packs_tc(int c_in, int c_out, int precision) {
    return precision != MOPA_SCN_PREC_FP32 && c_in % 16 == 0 && c_out % 16 == 0 && conv_uses_packed(c_in, c_out) &&
           conv_tc_enabled() && conv_tc_supported(c_in, c_out);
}

int pack_weights(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, int precision,
                 float *packed, cudaStream_t s) {
    const int c_in = transpose ? n_out : n_in, c_out = transpose ? n_in : n_out;
    MOPA_CHECK(c_in % 16 == 0 && c_out % 16 == 0, "packWeights: channel counts must be multiples of 16");
    const int split = precision == MOPA_SCN_PREC_FP32;
    if (!split && conv_tc_enabled() && conv_tc_supported(c_in, c_out))
        return pack_weights_tc(weight, volume, n_in, n_out, transpose, flip, packed, s);
    const int64_t total = (int64_t)volume * n_in * n_out * (split ? 2 : 1);
    k_pack_weights<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(weight, volume, n_in, n_out, transpose, flip, split, packed,
                                                                  total);
    MOPA_LAUNCHED();
    return 0;
}

Gather subm_gather(const Level &L) {
    Gather g;
    g.table = L.nbr; g.ld = L.nbr_ld; g.volume = 27; g.n_out = L.V; g.n_in = L.V; g.op = 1;
    g.tl = L.tl_subm; g.tm = L.tm_subm;
    return g;
}
Gather child_gather(const Level &fine, const Level &coarse, int op) {  // rows = coarse sites, inputs = fine sites
    Gather g;
    g.op = op;
    g.table = fine.child; g.ld = fine.child_ld; g.volume = 8; g.n_out = coarse.V; g.n_in = fine.V;
    g.tl = fine.tl_child; g.tm = fine.tm_child;
    return g;
}
Gather select_gather(const Level &fine, const Level &coarse, int op) {  // rows = fine sites, inputs = coarse sites
    Gather g;
    g.op = op;
    g.parent = fine.parent; g.kidx = fine.kidx; g.volume = 8; g.n_out = fine.V; g.n_in = coarse.V;
    g.tl = fine.tl_sel; g.tm = fine.tm_sel;
    return g;
}

}  // namespace mopa

using namespace mopa;

extern "C" {

int64_t mopa_scn_packedWeightFloats(int volume, int n_in, int n_out, int precision) {
    // room for either operand orientation in either kernel's layout (the tcgen05 layout pads channels to 32)
    const int64_t a = (int64_t)volume * n_in * n_out * (precision == MOPA_SCN_PREC_FP32 ? 2 : 1);
    const int64_t b = (int64_t)volume * round_up(n_in, 32) * round_up(n_out, 32);
    return a > b ? a : b;
}

int mopa_scn_packWeights(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, int precision,
                         float *packed, void *stream) {
    return pack_weights(weight, volume, n_in, n_out, transpose, flip, precision, packed, (cudaStream_t)stream);
}

size_t mopa_scn_backwardWorkspaceBytes(int volume, int n_in, int n_out, int64_t n_rows) {
    return dw_workspace_bytes(volume, n_in, n_out, n_rows);
}

static int get_level(mopa_scn_metadata *m, int64_t spatial, int &l) {
    MOPA_CHECK(m != nullptr, "null metadata");
    l = m->level_of(spatial);
    MOPA_CHECK(l >= 0, "no grid at this spatial size");
    return 0;
}

int mopa_scn_SubmanifoldConvolution_updateOutput(mopa_scn_metadata *m, int64_t spatial_size, int filter_size,
                                                 const float *in, int64_t ld_in, float *out, int64_t ld_out,
                                                 const float *weight, const float *packed, int n_in, int n_out,
                                                 int precision, void *stream) {
    int l;
    MOPA_TRY(get_level(m, spatial_size, l));
    MOPA_CHECK(filter_size == 3, "only 3x3x3 submanifold filters are implemented");
    cudaStream_t s = (cudaStream_t)stream;
    m->last_stream = s;
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_subm(m, l, s));
    return conv_apply(subm_gather(m->levels[l]), in, ld_in, out, ld_out, weight, packed, n_in, n_out, 0, 0, precision, s);
}

int mopa_scn_SubmanifoldConvolution_backward(mopa_scn_metadata *m, int64_t spatial_size, int filter_size,
                                             const float *in, int64_t ld_in, float *d_in, int64_t ld_din,
                                             const float *d_out, int64_t ld_dout, const float *weight,
                                             const float *packed_t, float *d_weight, int n_in, int n_out,
                                             int precision, void *workspace, size_t workspace_bytes, void *stream) {
    int l;
    MOPA_TRY(get_level(m, spatial_size, l));
    MOPA_CHECK(filter_size == 3, "only 3x3x3 submanifold filters are implemented");
    cudaStream_t s = (cudaStream_t)stream;
    m->last_stream = s;
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_subm(m, l, s));
    const Gather g = subm_gather(m->levels[l]);
    if (d_in)  // d_in[i] = sum_k d_out[nbr_k(i)] W[26-k]^T
        MOPA_TRY(conv_apply(g, d_out, ld_dout, d_in, ld_din, weight, packed_t, n_in, n_out, 1, 1, precision, s));
    if (d_weight)
        MOPA_TRY(conv_dweight(g, in, ld_in, d_out, ld_dout, d_weight, n_in, n_out, precision, workspace, workspace_bytes, s));
    return 0;
}

static int strided_levels(mopa_scn_metadata *m, int64_t fine_size, int64_t coarse_size, int filter_size,
                          int filter_stride, cudaStream_t s, int &l) {
    MOPA_TRY(get_level(m, fine_size, l));
    MOPA_CHECK(filter_size == 2 && filter_stride == 2, "only size-2 stride-2 (de)convolutions are implemented");
    MOPA_CHECK(coarse_size * 2 == fine_size, "spatial sizes do not match a size-2 stride-2 (de)convolution");
    m->last_stream = s;
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_down(m, l, s));
    return 0;
}

int mopa_scn_Convolution_updateOutput(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                      int filter_size, int filter_stride, const float *in, int64_t ld_in, float *out,
                                      int64_t ld_out, const float *weight, const float *packed, int n_in, int n_out,
                                      int precision, void *stream) {
    int l;
    cudaStream_t s = (cudaStream_t)stream;
    MOPA_TRY(strided_levels(m, in_spatial_size, out_spatial_size, filter_size, filter_stride, s, l));
    return conv_apply(child_gather(m->levels[l], m->levels[l + 1], 2), in, ld_in, out, ld_out, weight, packed, n_in, n_out, 0,
                      0, precision, s);
}

int mopa_scn_Convolution_backward(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                  int filter_size, int filter_stride, const float *in, int64_t ld_in, float *d_in,
                                  int64_t ld_din, const float *d_out, int64_t ld_dout, const float *weight,
                                  const float *packed_t, float *d_weight, int n_in, int n_out, int precision,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    int l;
    cudaStream_t s = (cudaStream_t)stream;
    MOPA_TRY(strided_levels(m, in_spatial_size, out_spatial_size, filter_size, filter_stride, s, l));
    if (d_in)  // d_in[c] = d_out[parent(c)] W[k(c)]^T
        MOPA_TRY(conv_apply(select_gather(m->levels[l], m->levels[l + 1], 2), d_out, ld_dout, d_in, ld_din, weight, packed_t,
                            n_in, n_out, 1, 0, precision, s));
    if (d_weight)  // rules (in = fine c, out = coarse p)
        MOPA_TRY(conv_dweight(child_gather(m->levels[l], m->levels[l + 1], 2), in, ld_in, d_out, ld_dout, d_weight, n_in,
                              n_out, precision, workspace, workspace_bytes, s));
    return 0;
}

int mopa_scn_Deconvolution_updateOutput(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                        int filter_size, int filter_stride, const float *in, int64_t ld_in, float *out,
                                        int64_t ld_out, const float *weight, const float *packed, int n_in, int n_out,
                                        int precision, void *stream) {
    int l;
    cudaStream_t s = (cudaStream_t)stream;
    MOPA_TRY(strided_levels(m, out_spatial_size, in_spatial_size, filter_size, filter_stride, s, l));
    return conv_apply(select_gather(m->levels[l], m->levels[l + 1], 3), in, ld_in, out, ld_out, weight, packed, n_in, n_out, 0,
                      0, precision, s);
}

int mopa_scn_Deconvolution_backward(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                    int filter_size, int filter_stride, const float *in, int64_t ld_in, float *d_in,
                                    int64_t ld_din, const float *d_out, int64_t ld_dout, const float *weight,
                                    const float *packed_t, float *d_weight, int n_in, int n_out, int precision,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    int l;
    cudaStream_t s = (cudaStream_t)stream;
    MOPA_TRY(strided_levels(m, out_spatial_size, in_spatial_size, filter_size, filter_stride, s, l));
    if (d_in)  // d_in[p] = sum_{children c} d_out[c] Wd[k(c)]^T
        MOPA_TRY(conv_apply(child_gather(m->levels[l], m->levels[l + 1], 3), d_out, ld_dout, d_in, ld_din, weight, packed_t,
                            n_in, n_out, 1, 0, precision, s));
    if (d_weight)  // rules (in = coarse parent(c), out = fine c)
        MOPA_TRY(conv_dweight(select_gather(m->levels[l], m->levels[l + 1], 3), in, ld_in, d_out, ld_dout, d_weight, n_in,
                              n_out, precision, workspace, workspace_bytes, s));
    return 0;
}

}  // extern "C"
