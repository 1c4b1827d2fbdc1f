// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
// Same contract as k_gather_mma in conv.cu (replaces [UPSTREAM] SparseConvNet SCN/CUDA/Convolution.cu's per-offset
// gather-FMA-scatter kernels reached from mopa/models/scn_unet.py:27-28), TF32 precision mode only.
//
// Why tcgen05 here: not for FLOPs (the layers are HBM/L2-bound) but for INSTRUCTION economy. With mma.sync every gathered
// row travels global -> registers -> cvt -> MMA fragments and the products come back as register fragments; at ~4 rules
// per site that costs hundreds of issue slots per useful MMA (ncu: 2.7% of the issued instructions were HMMA). Here a
// gathered row goes global -> shared memory with cp.async (LDGSTS, no registers), one elected thread issues the MMAs
// straight from shared memory, and the accumulator never leaves TMEM until the tile is finished.
//
// Output-stationary, no scatter at all: a CTA owns 256 consecutive OUTPUT rows = two M = 128 accumulator tiles in TMEM
// (D row = output row, N = C_out columns). For every filter offset k and 32-channel chunk, the A operand tile is
// "input row of the neighbour at offset k, or zero": rows without a neighbour stay zero in shared memory, so the dense MMA
// adds nothing for them. The tensor core runs ~6x more MACs than there are rules (a lidar site has ~4 of 27 neighbours),
// which still costs less time than the HBM floor of every layer; what is saved is all per-rule bookkeeping.
//
//   warps 0-7  gather  : four warps per M tile, warp w owns rows [32 (w & 3), +32) of its tile (= TMEM lane quarter w & 3).
//                        Neighbour ids are looked up kTcLook offsets ahead (register ring). Per step (k, chunk) a warp
//                        issues one ballot, 8 shuffles and 8 guarded cp.async (ignore-src form: the same instruction
//                        copies a live 16-byte piece or zeroes a stale one) straight into the A stage (UMMA K-major
//                        SWIZZLE_128B layout); which rows a stage holds from its previous use lives in registers. The
//                        copies' completion arrives on the stage's mbarrier asynchronously (cp.async.mbarrier.arrive
//                        .noinc): a warp never waits for data, all stages of the ring can be in flight.
//   warp  8    weights : one lane streams W[k] chunks (pre-packed N x K K-major, same swizzle) by TMA bulk copy.
//   warps 9,10 MMA     : one issuer warp per M tile: an elected lane issues tcgen05.mma (M = 128, N = C_out, K = 8 per
//                        instruction, fp32 accumulate in TMEM) and commits the stage releases to mbarriers.
//   epilogue           : when the last MMA has retired, the gather warps read their TMEM lanes (one output row per
//                        thread) and write every output row to HBM once (optionally added to what is there); in the
//                        forward pass of a training step they also reduce the per-column sum / sum of squares of the
//                        rows for the BatchNorm that follows (shuffle transpose-reduce, one fp64 atomic per column and CTA).
// Variants: one M tile per CTA on levels too small for 256-row CTAs; two 32-channel atoms per step (k_conv_tc<2>) where a
// CTA is alone on its SM and bound by the per-step latency.
// Summation order is fixed (k ascending, chunks ascending, hardware order inside an MMA): outputs are deterministic.
#include <stdlib.h>

#include "geometry.cuh"
#include "mopa_scn.h"
#include "ptx.cuh"

namespace mopa {

constexpr int kTcTM = 256;            // output rows per CTA (two M = 128 tiles)
constexpr int kTcThreads = 11 * 32;  // 8 gather warps, weight producer, two MMA issuers
constexpr int kTcChunk = 32;          // input channels per pipeline step (one 128-byte swizzle row)
constexpr int kTcAStage = 128 * 128;  // bytes: 128 rows x 128 bytes
constexpr int kTcMaxSA = 12, kTcMaxSB = 8;
constexpr int kTcStatsLd = 256;      // BatchNorm statistics block of a buffer: [sum x | sum x^2], 256 doubles each
constexpr int kTcLook = 6;          // neighbour ids are looked up this many offsets ahead

__host__ __device__ inline int tc_tmem_cols(int nt, int tpc = 2) {  // power of two >= 32 holding tpc accumulators of nt columns
    int c = 32;
    while (c < tpc * nt) c <<= 1;
    return c;
}

// packed[k][chunk][NT rows x 128 bytes, SWIZZLE_128B]: B operand (N x K, K-major) of one pipeline step.
// element (n, c) of a chunk = W'[ci = 32 chunk + c][co = n], rounded to TF32 (zero for ci >= c_in).
__device__ __forceinline__ void tc_pack_element(const float *__restrict__ w, int volume, int n_in0, int n_out0, int transpose,
                                                int flip, float *__restrict__ packed, int64_t idx) {
    const int c_in = transpose ? n_out0 : n_in0, nt = transpose ? n_in0 : n_out0;
    const int nchunk = (c_in + kTcChunk - 1) / kTcChunk;
    const int64_t per_chunk = (int64_t)nt * kTcChunk, per_k = per_chunk * nchunk;
    const int k = (int)(idx / per_k);
    int64_t r = idx - (int64_t)k * per_k;
    const int chunk = (int)(r / per_chunk);
    r -= (int64_t)chunk * per_chunk;
    // r = float index inside the swizzled tile: row n = r / 32, physical 16-byte piece pp = (r % 32) / 4
    const int n = (int)(r >> 5), pp = (int)(r & 31) >> 2, e = (int)(r & 3);
    const int c = ((pp ^ (n & 7)) << 2) + e;  // logical channel inside the chunk
    const int ci = chunk * kTcChunk + c, co = n;
    const int ks = flip ? volume - 1 - k : k;
    float v = 0.f;
    if (ci < c_in) v = transpose ? w[((int64_t)ks * n_in0 + co) * n_out0 + ci] : w[((int64_t)ks * n_in0 + ci) * n_out0 + co];
    packed[idx] = __uint_as_float(to_tf32(v));
}
__global__ void __launch_bounds__(256) k_pack_weights_tc(const float *__restrict__ w, int volume, int n_in0, int n_out0,
                                                         int transpose, int flip, float *__restrict__ packed,
                                                         int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < total) tc_pack_element(w, volume, n_in0, n_out0, transpose, flip, packed, idx);
}
// every convolution of a pass in one launch (grid.y = job): the whole-network executor packs all weights up front
__global__ void __launch_bounds__(256) k_pack_weights_tc_batch(const __grid_constant__ TcPackJobs jobs) {
    const TcPackJob &j = jobs.job[blockIdx.y];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < j.total; idx += (int64_t)gridDim.x * blockDim.x)
        tc_pack_element(j.w, j.volume, j.n_in, j.n_out, j.transpose, j.flip, j.packed, idx);
}

#ifdef MOPA_TC_TRACE
// debug timeline (build with -DMOPA_TC_TRACE, run scratch/tc_trace.py): clock64 stamps of one mid-grid CTA, rows: 0 gather
// step top, 1 a_empty acquired, 2 copies issued + arrive, 4 issuer before a_full wait, 5 a_full seen, 6 MMAs + commit issued
__device__ long long g_tc_trace[8][512];
#define TC_STAMP(cond, row, idx) do { if ((cond) && (idx) < 512) g_tc_trace[row][idx] = clock64(); } while (0)
#else
#define TC_STAMP(cond, row, idx) do { } while (0)
#endif

struct TcSmem {  // byte offsets inside the dynamic shared memory block (base aligned to 1024)
    int a, b, mask, bars, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int nt, int sa, int sb, int na = 1) {  // na: 32-channel atoms per stage
    TcSmem L;
    L.a = 0;
    L.b = L.a + sa * na * kTcAStage;
    L.mask = L.b + sb * na * nt * 128;
    L.bars = L.mask + 64;
    L.total = L.bars + 8 * (2 * kTcMaxSA + 2 * kTcMaxSB + 1) + 16;
    return L;
}

template <int NA>
__global__ void __launch_bounds__(kTcThreads, NA == 2 ? 1 : 2)  // the two-atom variant only runs one CTA per SM
    k_conv_tc(Gather gt, const float *__restrict__ in, int64_t ld_in, float *__restrict__ out, int64_t ld_out,
              const float *__restrict__ packed, int c_in, int NT, int SA, int SB, int TPC, double *__restrict__ stats) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const TcSmem L = tc_smem_layout(NT, SA, SB, NA);
    unsigned char *sA = smem + L.a, *sB = smem + L.b;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L.bars), *a_empty = a_full + kTcMaxSA;
    uint64_t *b_full = a_empty + kTcMaxSA, *b_empty = b_full + kTcMaxSB;
    uint64_t *d_full = b_empty + kTcMaxSB;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(d_full + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = gt.volume;
    // NA = 32-channel atoms per pipeline step (1, or 2 on the small levels: their CTAs are alone on an SM and bound by
    // the per-step latency, so half as many, twice as wide steps)
    const int chw = kTcChunk * NA;                       // channels per step
    const int nchunk = (c_in + chw - 1) / chw;           // steps per offset
    const int nchunk32 = (c_in + kTcChunk - 1) / kTcChunk;  // packed weight chunks per offset
    const uint32_t a_stage = (uint32_t)NA * kTcAStage, b_stage = (uint32_t)NA * NT * 128;
    // TPC = M tiles per CTA: 2 for the large levels; 1 for levels too small to fill the GPU with 256-row CTAs (twice
    // the CTAs, and the whole stage ring serves the one tile)
    const int64_t row0 = (int64_t)blockIdx.x * (128 * TPC);
    const int n_mt = (TPC == 2 && gt.n_out - row0 > 128) ? 2 : 1;  // M tiles of this CTA that hold rows
    const uint32_t tmem_cols = (uint32_t)tc_tmem_cols(NT, TPC);

    if (tid == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, n_mt); }
        mbar_init(d_full, n_mt);
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr, tmem_cols);
    if (warp < 8) {  // all A stages start all-zero; nothing has been written by any warp yet
        for (int i = tid; i < SA * NA * kTcAStage / 16; i += 256) reinterpret_cast<float4 *>(sA)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 8) {
        // ================================================================= gather warps
        // warps 0-3 feed M tile 0, warps 4-7 M tile 1; warp w owns tile rows [32 (w & 3), +32) = TMEM lane quarter w & 3
        const int mt = warp >> 2, wq = warp & 3;
        const bool live = mt < n_mt;
        const int64_t row = row0 + 128 * mt + 32 * wq + lane;
        const bool row_ok = live && row < gt.n_out;
        // everything below addresses shared memory through 32-bit shared-window addresses
        const uint32_t tile0_a = smem_u32(sA) + wq * 4096;     // + stage * 16 KB: this warp's 32 rows (4 swizzle groups)
        const uint32_t full0_a = smem_u32(a_full), empty0_a = smem_u32(a_empty);
        // ring bookkeeping without divisions: this warp's steps are s = TPC j + mt -> stage s % SA (SA even when TPC = 2)
        int st = mt;
        uint32_t ph = 1;  // parity of the a_empty wait of the next step to issue
        const int n_k = live ? K : 0;
        // neighbour ids are fetched kTcLook offsets ahead (a ring of registers with static indices: the k loop is unrolled
        // kTcLook times). One offset ahead was not enough: a step is shorter than an L2/HBM round trip, and ncu showed the
        // gather warps spending most of their time on the scoreboard of this load.
        int sel_k = -2, sel_p = -1;  // select mode: the row's only offset and its source row
        if (!gt.table && row_ok) { sel_k = __ldg(gt.kidx + row); sel_p = __ldg(gt.parent + row); }
        auto look = [&](int k) -> int {
            if (!row_ok || k >= K) return -1;
            if (gt.table) return __ldg(gt.table + (int64_t)k * gt.ld + row);
            return sel_k == k ? sel_p : -1;
        };
        int nq[kTcLook];
#pragma unroll
        for (int u = 0; u < kTcLook; ++u) nq[u] = look(u);
        // Per step (offset k, 32-channel chunk) a warp (lane = row for the lookup) does one ballot and then NP passes of
        // 32 / LPR rows x LPR lanes (one 16-byte piece per lane): rows with a rule get their bytes by cp.async; rows the
        // warp wrote in the stage's previous use and does not rewrite are zero-filled through the same instruction
        // (ignore-src operand), so a stage is zero outside its live rules. Which rows were written last time lives in
        // registers (lane s keeps the record of stage s): no shared-memory lists, no __syncwarp, no branches inside a
        // pass. Everything that does not change per step (row of each pass, its bit, its swizzled offset) is hoisted.
        // History: the first version kept compacted row lists in shared memory (~220 instructions per step), the second
        // used the src-size form of cp.async (~370 SASS instructions per step after ptxas expanded it); a device-side
        // timeline (scratch/tc_trace.py) showed the gather warps spend ~1800 cycles per step ISSUING and ~80 waiting.
        uint32_t oldm = 0, oldp = 0;
        int tstep = 0;  // trace builds only
        const int LPR = c_in == 16 ? 4 : 8;      // lanes per row (64-byte rows need 4)
        const int NP = LPR;                      // passes per step: 32 rows / (32 / LPR rows per pass)
        const int cl = lane & (LPR - 1);         // this lane's 16-byte piece
        const int rl = lane / LPR;               // row inside a pass
        uint32_t doff[8];
        int rj[8];  // row of pass j (32 for the passes a narrow layout does not have: its bit shifts out)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = (j * (32 / LPR) + rl) & 31;
            rj[j] = j < NP ? r : 32;
            doff[j] = (uint32_t)r * 128 + (uint32_t)((cl ^ (r & 7)) << 4);
        }
        const char *in_c = reinterpret_cast<const char *>(in) + 16 * cl;
        const uint32_t ldb = (uint32_t)ld_in * 4;  // row pitch in bytes (feature matrices are far below 4 GB)
        auto step = [&](const int nv) {
            const uint32_t m_new = __ballot_sync(0xffffffffu, nv >= 0);
            const char *rp[8];  // source row of each pass (row 0 where there is no rule: never read, must be mapped)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int srow = __shfl_sync(0xffffffffu, nv, rj[j] & 31);
                rp[j] = in_c + (uint64_t)(uint32_t)max(srow, 0) * (uint64_t)ldb;
            }
            for (int ch = 0; ch < nchunk; ++ch) {
                // 16-byte pieces per row in each 32-channel atom of this step (the last step of an offset may be short)
                const int left = c_in - ch * chw;
                const uint32_t pa0 = (uint32_t)min(kTcChunk, left) / 4, pa1 = NA == 2 ? (uint32_t)max(0, min(kTcChunk, left - kTcChunk)) / 4 : 0u;
                const uint32_t pieces = pa0 | (pa1 << 8);
                TC_STAMP(warp == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 0, tstep);
                mbar_wait_s(empty0_a + 8 * st, ph);
                TC_STAMP(warp == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 1, tstep);
                const uint32_t m_old = __shfl_sync(0xffffffffu, oldm, st), p_old = __shfl_sync(0xffffffffu, oldp, st);
                const uint32_t tile_a = tile0_a + st * a_stage;
                const int cho = ch * (chw * 4);
                {
                    const uint32_t a_new = (uint32_t)cl < pa0 ? m_new : 0u;  // rows this lane copies / may have to clear
                    const uint32_t a_old = (uint32_t)cl < (p_old & 0xffu) ? m_old : 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // bit of this pass's row: clamped funnel shift, so that a shift count of 32 yields 0 (absent pass)
                        const uint32_t is_new = __funnelshift_rc(a_new, 0u, rj[j]) & 1u;
                        cp_async16_zfill_pred_s(tile_a + doff[j], rp[j] + cho, is_new | (__funnelshift_rc(a_old, 0u, rj[j]) & 1u),
                                                is_new ^ 1u);
                    }
                }
                if (NA == 2) {  // second atom of the step: channels [32, 64) of the chunk, 16 KB further in the stage
                    const uint32_t a_new = (uint32_t)cl < pa1 ? m_new : 0u;
                    const uint32_t a_old = (uint32_t)cl < (p_old >> 8) ? m_old : 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t is_new = __funnelshift_rc(a_new, 0u, rj[j]) & 1u;
                        cp_async16_zfill_pred_s(tile_a + kTcAStage + doff[j], rp[j] + cho + 128,
                                                is_new | (__funnelshift_rc(a_old, 0u, rj[j]) & 1u), is_new ^ 1u);
                    }
                }
                // every lane: "my copies of this step have landed" arrives on the stage's barrier asynchronously (128
                // arrivals complete it); the warp never waits for data, so all stages of the ring can be in flight
                cp_async_mbar_arrive_noinc_s(full0_a + 8 * st);
                TC_STAMP(warp == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 2, tstep);
                ++tstep;
                if (lane == st) { oldm = m_new; oldp = pieces; }
                st += TPC;
                if (st >= SA) { st -= SA; ph ^= 1; }
            }
        };
        for (int k0 = 0; k0 < n_k; k0 += kTcLook) {
#pragma unroll
            for (int u = 0; u < kTcLook; ++u) {
                if (k0 + u < n_k) {  // warp-uniform
                    const int nv = nq[u];
                    nq[u] = look(k0 + u + kTcLook);
                    step(nv);
                }
            }
        }

        // ================================================================= epilogue: TMEM -> HBM, one output row per thread
        if (live) {
            mbar_wait(d_full, 0);
            tc_fence_after_sync();
            float *dst = out + row * ld_out;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * wq) << 16) + (uint32_t)(mt * NT);
            // per-column sums of the output rows for the BatchNorm that follows (stats != nullptr): warp transpose-reduce,
            // one partial per warp in shared memory (the stage ring is idle once d_full has completed), combined after
            // the final barrier and added to the fp64 accumulators with one atomic per column and CTA
            float *s_part = reinterpret_cast<float *>(sA) + (size_t)warp * 2 * NT;
            for (int q = 0; q < NT / 16; ++q) {
                float v[16];
                tmem_ld16(taddr + 16 * q, v);  // warp-collective: every lane takes part, also beyond n_out
                if (row < gt.n_out) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float4 o = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        float4 *p = reinterpret_cast<float4 *>(dst + 16 * q + 4 * e);
                        if (gt.accumulate) { const float4 x = *p; o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w; }
                        *p = o;
                    }
                }
                if (stats) {  // warp-uniform
                    float a[16], b[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        a[e] = row < gt.n_out ? v[e] : 0.f;
                        b[e] = a[e] * a[e];
                    }
                    // butterfly: after the steps with strides 16, 8, 4, 2 a lane holds ONE column's sum over 16 of the 32
                    // rows (column = bits 4..1 of the lane, msb first); stride 1 adds the two halves
#pragma unroll
                    for (int w = 8, stride = 16; w >= 1; w >>= 1, stride >>= 1) {
                        const bool up = lane & stride;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (e < w) {
                                const float sa = up ? a[e] : a[e + w], ka = up ? a[e + w] : a[e];
                                const float sb2 = up ? b[e] : b[e + w], kb = up ? b[e + w] : b[e];
                                a[e] = ka + __shfl_xor_sync(0xffffffffu, sa, stride);
                                b[e] = kb + __shfl_xor_sync(0xffffffffu, sb2, stride);
                            }
                        }
                    }
                    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                    b[0] += __shfl_xor_sync(0xffffffffu, b[0], 1);
                    if (!(lane & 1)) {
                        const int col = 16 * q + (((lane >> 4) & 1) << 3) + (((lane >> 3) & 1) << 2) + (((lane >> 2) & 1) << 1) + ((lane >> 1) & 1);
                        s_part[col] = a[0];
                        s_part[NT + col] = b[0];
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ================================================================= weight producer (TMA bulk copies)
        if (lane == 0) {
            int st = 0, ph = 1;
            for (int k = 0; k < K; ++k)
                for (int ch = 0; ch < nchunk; ++ch) {
                    const int n32 = min(NA, nchunk32 - ch * NA);  // packed 32-channel chunks in this step (contiguous)
                    const uint32_t bytes = (uint32_t)n32 * NT * 128;
                    mbar_wait(b_empty + st, ph);
                    mbar_expect_tx(b_full + st, bytes);
                    tma_load_1d(sB + (size_t)st * b_stage, packed + ((int64_t)k * nchunk32 + ch * NA) * NT * kTcChunk, bytes,
                                b_full + st);
                    if (++st == SB) { st = 0; ph ^= 1; }
                }
        }
    } else if (warp - 9 < n_mt) {
        // ================================================================= MMA issuers: warp 9 -> M tile 0, warp 10 -> M tile 1
        // One issuer per tile: a single issuer served the two tiles in turn, so a late stage of one tile held back the
        // other (the timeline showed ~450 cycles per tile step in this warp, two tiles back to back per step).
        // The whole warp runs the (warp-uniform) loops and waits; one elected lane executes the tcgen05 instructions. Keeping
        // the control flow uniform keeps the descriptors in uniform registers (issuing from an `if (lane == 0)` branch cost
        // ~130 cycles of R2UR traffic per MMA).
        const int mt = warp - 9;
        const uint32_t idesc = umma_idesc_tf32(NT);
        const uint64_t desc_hi = umma_desc_sw128(0);  // everything but the start address
        const uint32_t d = tmem_base + (uint32_t)(mt * NT);
        int st = mt, ph = 0, stb = 0, phb = 0, tstep = 0;
        for (int k = 0; k < K; ++k) {
            for (int ch = 0; ch < nchunk; ++ch) {
                mbar_wait(b_full + stb, phb);
                const int left = c_in - ch * chw;
                const int nk0 = min(kTcChunk, left) / 8;                                   // MMAs (K = 8 each) from atom 0
                const int nk1 = NA == 2 ? max(0, min(kTcChunk, left - kTcChunk)) / 8 : 0;  // ... and from atom 1
                const uint64_t b_desc = desc_hi | (uint64_t)((smem_u32(sB + (size_t)stb * b_stage) & 0x3FFFFu) >> 4);
                TC_STAMP(mt == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 4, tstep);
                mbar_wait(a_full + st, ph);
                TC_STAMP(mt == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 5, tstep);
                fence_proxy_async_smem();  // the gather warps' cp.async writes (generic proxy) -> UMMA reads
                tc_fence_after_sync();
                const uint64_t a_desc = desc_hi | (uint64_t)((smem_u32(sA + (size_t)st * a_stage) & 0x3FFFFu) >> 4);
                const uint32_t first = (k > 0 || ch > 0) ? 1u : 0u;
                if (elect_one()) {
                    for (int j = 0; j < nk0; ++j)  // + 32 bytes of K per MMA = + 2 in the address field
                        umma_tf32(d, a_desc + 2 * j, b_desc + 2 * j, idesc, j > 0 ? 1u : first);
                    for (int j = 0; j < nk1; ++j)  // second atom: 16 KB further in A, NT x 128 bytes further in B
                        umma_tf32(d, a_desc + (kTcAStage >> 4) + 2 * j, b_desc + (uint64_t)((NT * 128) >> 4) + 2 * j, idesc, 1u);
                    umma_commit(a_empty + st);
                    umma_commit(b_empty + stb);  // n_mt arrivals (one per issuer) release the weight stage
                }
                __syncwarp();
                TC_STAMP(mt == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 6, tstep);
                ++tstep;
                st += TPC;
                if (st >= SA) { st -= SA; ph ^= 1; }
                if (++stb == SB) { stb = 0; phb ^= 1; }
            }
        }
        if (elect_one()) umma_commit(d_full);  // n_mt arrivals complete it
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, tmem_cols);
    if (stats && tid < 2 * NT) {  // [0, NT): sum x, [NT, 2 NT): sum x^2; warps of the tiles that hold rows, in order
        const float *s_all = reinterpret_cast<const float *>(sA);
        float sum = 0.f;
        for (int w = 0; w < 4 * n_mt; ++w) sum += s_all[(size_t)w * 2 * NT + tid];
        atomicAdd(stats + (tid < NT ? tid : kTcStatsLd + (tid - NT)), (double)sum);
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool conv_tc_supported(int c_in, int c_out) {
    return c_in % 16 == 0 && c_in >= 16 && c_out % 16 == 0 && c_out >= 16 && c_out <= 256;
}

int64_t tc_packed_floats(int volume, int c_in, int c_out) {
    return (int64_t)volume * ceil_div(c_in, kTcChunk) * kTcChunk * c_out;
}

int pack_weights_tc(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, float *packed,
                    cudaStream_t s) {
    const int c_in = transpose ? n_out : n_in, c_out = transpose ? n_in : n_out;
    MOPA_CHECK(conv_tc_supported(c_in, c_out), "packWeights: shape is not on the tcgen05 path");
    const int64_t total = tc_packed_floats(volume, c_in, c_out);
    k_pack_weights_tc<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(weight, volume, n_in, n_out, transpose, flip, packed,
                                                                    total);
    MOPA_LAUNCHED();
    return 0;
}

int pack_weights_tc_batch(TcPackJobs &jobs, int n_jobs, cudaStream_t s) {
    if (n_jobs == 0) return 0;
    MOPA_CHECK(n_jobs <= kTcMaxPackJobs, "packWeights: too many jobs in one batch");
    int64_t most = 0;
    for (int i = 0; i < n_jobs; ++i) {
        TcPackJob &j = jobs.job[i];
        const int c_in = j.transpose ? j.n_out : j.n_in, c_out = j.transpose ? j.n_in : j.n_out;
        MOPA_CHECK(conv_tc_supported(c_in, c_out), "packWeights: shape is not on the tcgen05 path");
        j.total = tc_packed_floats(j.volume, c_in, c_out);
        if (j.total > most) most = j.total;
    }
    int64_t bx = ceil_div(most, 256 * 4);  // ~4 elements per thread for the largest job
    if (bx < 1) bx = 1;
    k_pack_weights_tc_batch<<<dim3((unsigned)bx, (unsigned)n_jobs), 256, 0, s>>>(jobs);
    MOPA_LAUNCHED();
    return 0;
}

int conv_apply_tc(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out, const float *packed,
                  int c_in, int c_out, double *stats, cudaStream_t s) {
    const int nt = c_out;
    // ring depths. An A stage is 16 KB but carries only the few rows that have a rule at that offset, so the gather bytes
    // in flight are set by the NUMBER of stages (kept even for two-tile CTAs: tile 0 uses the even stages, tile 1 the odd).
    static const int want_ctas = [] { const char *e = getenv("MOPA_TC_CTAS"); return e ? atoi(e) : 2; }();
    static const int want_sb = [] { const char *e = getenv("MOPA_TC_SB"); return e ? atoi(e) : 4; }();
    static const int want_tpc = [] { const char *e = getenv("MOPA_TC_TPC"); return e ? atoi(e) : 0; }();
    // one M tile per CTA only for levels so small that 256-row CTAs would leave SMs empty (the weight tiles are
    // streamed per CTA: halving the rows per CTA doubles that traffic, which costs more than it gains on mid-size levels)
    const int tpc = want_tpc ? want_tpc : (ceil_div(gt.n_out, kTcTM) > kNumSMs ? 2 : 1);
    int sb = want_sb < 2 ? 2 : (want_sb > kTcMaxSB ? kTcMaxSB : want_sb);
    // two 32-channel atoms per step on the small levels (one CTA per SM, bound by the per-step latency)
    static const int want_na = [] { const char *e = getenv("MOPA_TC_NA"); return e ? atoi(e) : 0; }();
    const int na = want_na ? want_na : ((tpc == 1 && c_in >= 64 && ceil_div(gt.n_out, 128) <= kNumSMs) ? 2 : 1);
    while (sb > 2 && (size_t)sb * na * nt * 128 > (size_t)(tpc == 2 ? 64 : 32 * na) * 1024) --sb;
    // two CTAs per SM when that helps: not when the whole grid fits one CTA per SM anyway (then the one CTA gets all stages)
    bool two = want_ctas >= 2 && tc_tmem_cols(nt, tpc) <= 256 && (nt <= 64 || tpc == 1) &&
               ceil_div(gt.n_out, 128 * tpc) > kNumSMs;
    // (Three CTAs per SM with a 4-stage ring were measured on the narrow, large levels: within noise of two CTAs with six
    // stages, and the 11-warp CTA does not fit three times in the register file without spills; not kept.)
    int sa = 0;
    size_t cap = 0;
    for (;;) {
        cap = two ? (size_t)113 * 1024 : (size_t)226 * 1024;
        sa = kTcMaxSA;
        while (sa > 2 && (size_t)tc_smem_layout(nt, sa, sb, na).total + 1024 > cap) sa -= tpc;
        if (!two || sa >= 4) break;
        two = false;  // too few stages at two CTAs per SM: take the whole SM
    }
    MOPA_CHECK((size_t)tc_smem_layout(nt, sa, sb, na).total + 1024 <= cap && sa >= 2, "conv_tc: shared memory layout does not fit");
    const size_t smem = (size_t)tc_smem_layout(nt, sa, sb, na).total + 1024;
    static bool configured = false;
    if (!configured) {
        MOPA_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MOPA_CUDA(cudaFuncSetAttribute(k_conv_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(gt.n_out, 128 * tpc));
    if (na == 2)
        k_conv_tc<2><<<grid, kTcThreads, smem, s>>>(gt, in, ld_in, out, ld_out, packed, c_in, nt, sa, sb, tpc, stats);
    else
        k_conv_tc<1><<<grid, kTcThreads, smem, s>>>(gt, in, ld_in, out, ld_out, packed, c_in, nt, sa, sb, tpc, stats);
    MOPA_LAUNCHED();
    return 0;
}

}  // namespace mopa

#ifdef MOPA_TC_TRACE
extern "C" int mopa_scn_debug_tc_trace(long long *host) {  // copies the timeline of the last launch out, then clears it
    if (cudaMemcpyFromSymbol(host, mopa::g_tc_trace, sizeof(long long) * 8 * 512) != cudaSuccess) return 1;
    static long long zeros[8 * 512];
    return cudaMemcpyToSymbol(mopa::g_tc_trace, zeros, sizeof(zeros)) != cudaSuccess;
}
#endif
