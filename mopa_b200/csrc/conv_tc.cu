// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
// Replaces [UPSTREAM] SparseConvNet SCN/CUDA/Convolution.cu's per-offset gather-FMA-scatter kernels reached from
// mopa/models/scn_unet.py:27-28 (SubmanifoldConvolution, Convolution, Deconvolution forward and input gradient), TF32 mode.
//
// Output-stationary, no scatter: a CTA owns one or two tiles of 128 consecutive OUTPUT rows; a tile's accumulator is
// (128 TMEM lanes = rows) x (C_out fp32 columns). For every filter offset k that has a rule in the tile, and every
// 32-channel chunk, the A operand is "input row of the rule whose output is tile row r" placed AT row r of a 128 x 32
// K-major SWIZZLE_128B stage; tcgen05.mma (kind::tf32, M = 128, N = C_out, K = 8) accumulates with the per-row write mask
// of that (tile, offset) (PTX disable-output-lane), so rows without a rule are untouched whatever the stage holds there:
// nothing is ever zero-filled.
//
// What round 1 / the first half of round 2 measured (profiles/r01_*, r02_conv_tc_history.txt): with the rulebook as a
// dense (offset, row) -> row table, the gather warps spend ~130 instructions per (warp, offset) on ballots, shuffles and
// bookkeeping for ~4 live rows. Moving that out of the kernel (geometry.cu::k_tile_lists_batch builds, once per level and
// shared by the six convolution passes that use the level, per (tile, offset) a compact list of (input row, tile row) pairs
// and the 128-bit row mask) cut the instruction count ~4x and did NOT change the time: a stage-step is a ~2200-cycle chain
// (copies issued -> data landed -> MMAs issued -> commit -> stage free) and what bounds the kernel is how many such chains
// an SM keeps in flight (profiles/r02_tc_timeline.txt). Hence the roles:
//   warps 0-7  gather : 4 per tile, in GW groups; a group fills every GW-th stage-step of its tile, so a stage is always
//                       filled by the same warps (mbarrier parity waits cannot tell phases two apart). A warp loads 32
//                       list entries with one coalesced LDG (the next step's entries are prefetched) and per pass copies
//                       32 / LPR rows with one LDGSTS (LPR lanes x 16 bytes per row, cp.async straight into the swizzled
//                       stage). Completion arrives on the stage's mbarrier asynchronously (cp.async.mbarrier.arrive.noinc).
//   warp  8    weights: one lane streams W[k] chunks of the live offsets (pre-packed N x K K-major) by TMA bulk copy.
//   warps 9..  MMA    : ACC issuer warps per tile, each with its OWN TMEM accumulator set and every ACC-th stage-step:
//                       elected lane issues the masked tcgen05.mma's and commits the stage releases.
//   epilogue          : gather warp (tile mt, quarter wq) reads TMEM lanes [32 wq, +32), adds the ACC sets, writes each
//                       output row once (optionally accumulating), and reduces two per-column sums for the BatchNorm next
//                       to the convolution (forward: sum x, sum x^2; optional d_input variant: the backward sums).
// Offsets with no rule in any tile of the CTA are skipped by all three roles (same 27-bit live set, from the masks). On the
// small levels the live offsets of a tile are split over several CTAs (TcExtra::split). Summation order is fixed: outputs
// are deterministic run to run.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "geometry.cuh"
#include "mopa_scn.h"
#include "ptx.cuh"

namespace mopa {

constexpr int kTcTM = 256;            // output rows per CTA (two M = 128 tiles)
constexpr int kTcMaxThreads = 17 * 32;  // 8 gather warps, weight producer, up to 8 MMA issuers (tiles x accumulator sets)
constexpr int kTcChunk = 32;          // input channels per pipeline step (one 128-byte swizzle row)
constexpr int kTcAStage = 128 * 128;  // bytes: 128 rows x 128 bytes
constexpr int kTcMaxSA = 12, kTcMaxSB = 8;
constexpr int kTcAhead = 1;           // a gather warp fetches its rule-list entries this many of its stage-steps ahead
constexpr int kTcStatsLd = 256;      // BatchNorm statistics block of a buffer: [sum x | sum x^2], 256 doubles each

__host__ __device__ inline int tc_tmem_cols(int nt, int tpc = 2) {  // power of two >= 32 holding tpc accumulators of nt columns
    int c = 32;
    while (c < tpc * nt) c <<= 1;
    return c;
}

// packed[k][chunk][NT rows x 128 bytes, SWIZZLE_128B]: B operand (N x K, K-major) of one pipeline step.
// element (n, c) of a chunk = W'[ci = 32 chunk + c][co = n], rounded to TF32 (zero for ci >= c_in).
// c_in == 16 (half a swizzle row per offset): tile p holds a PAIR of offsets, channels [0, 16) = offset 2p, [16, 32) = offset
// 2p + 1 (zero past the last offset); tiles ceil(K / 2) .. K - 1 are unused. The kernel then runs one stage-step per pair.
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
    int ci = chunk * kTcChunk + c, kk = k;
    const int co = n;
    if (c_in == 16) {  // paired offsets
        kk = 2 * k + (c >> 4);
        ci = c & 15;
        if (kk >= volume) ci = c_in;  // -> zero
    }
    const int ks = flip ? volume - 1 - kk : kk;
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
    int a, b, bars, cst, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int nt, int sa, int sb, int na = 1, bool bn = false) {  // na: 32-channel atoms per stage
    TcSmem L;
    L.a = 0;
    L.b = L.a + sa * na * kTcAStage;
    L.bars = L.b + sb * na * nt * 128;
    L.cst = (L.bars + 8 * (2 * kTcMaxSA + 2 * kTcMaxSB + 1) + 16 + 15) & ~15;  // [scale | shift | mean] of the BatchNorm in front (d_input pass)
    L.total = L.cst + (bn ? 3 * nt * 4 : 0);
    return L;
}

// The less common arguments of a launch.
//  * split > 1: `split` CTAs (blockIdx.y) share one output tile, each takes every split-th live offset. A CTA parks its
//    partial accumulator in `partial` ([tile][part][column][128 rows], L2-resident), takes a ticket, and the LAST CTA of
//    the tile to arrive adds the parts in part order (deterministic whatever the arrival order) and runs the epilogue.
//    For the small levels: 29-72 tiles would leave most SMs idle while each CTA walks 27 offsets one after the other.
//  * bn_x != nullptr (d_input pass whose output is the gradient of a BatchNormReLU's output): the epilogue also reduces,
//    per column, S1 = sum d and S2 = sum (x - mean) d with d = the gradient masked by the sign of the recomputed BatchNorm
//    output, into `stats` -- the BatchNorm backward that follows then needs no reduction pass of its own.
struct TcExtra {
    float *partial;
    unsigned *tickets;
    int split;
    const float *bn_x;
    int64_t ld_bn_x;
    const float *bn_mean, *bn_invstd, *bn_weight, *bn_bias;
    float leak;
    int bn_ring;  // the tile's rows of x are staged through the A ring (32-column slices) instead of read row by row
};

__device__ __forceinline__ float4 tc_lds128(uint32_t addr) {  // explicit shared-space load (pointers derived from the aligned
    float4 t;                                                   // dynamic block are generic to the compiler: LD, not LDS)
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
    return t;
}
__device__ __forceinline__ void tc_bar_sync_128(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// WIDE = 1: up to 17 warps, one CTA per SM; WIDE = 0: up to 13 warps (two tiles x two issuers), two CTAs per SM;
// WIDE = 2: 11 warps (two tiles, one issuer each), three CTAs per SM (narrow layers: small weight stages)
// BN = true: the d_input pass in front of a BatchNorm (TcExtra::bn_x); a separate instantiation, so that the forward
// kernels do not carry its code: the prologue and the epilogue run once per CTA, from a cold instruction cache.
template <int NA, int LPR, int WIDE, bool BN>
__global__ void __launch_bounds__(WIDE == 1 ? kTcMaxThreads : (WIDE == 2 ? 11 * 32 : 13 * 32), WIDE == 1 ? 1 : (WIDE == 2 ? 3 : 2))
    k_conv_tc(Gather gt, const float *__restrict__ in, int64_t ld_in, float *__restrict__ out, int64_t ld_out,
              const float *__restrict__ packed, int c_in, int NT, int SA, int SB, int TPC, int GW, int ACC,
              double *__restrict__ stats, const __grid_constant__ TcExtra X) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const TcSmem L = tc_smem_layout(NT, SA, SB, NA, BN);
    unsigned char *sA = smem + L.a, *sB = smem + L.b;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L.bars), *a_empty = a_full + kTcMaxSA;
    uint64_t *b_full = a_empty + kTcMaxSA, *b_empty = b_full + kTcMaxSB;
    uint64_t *d_full = b_empty + kTcMaxSB;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(d_full + 1);
    volatile int *final_flag = reinterpret_cast<volatile int *>(tmem_ptr + 1);  // split launches: this CTA runs the epilogue
    float *cst = reinterpret_cast<float *>(smem + L.cst);
    const int S = X.split;

    // LPR == 4 is the 16-input-channel instantiation: a row is half a swizzle row, so a stage-step carries TWO offsets
    // (2u in bytes [0, 64) of the rows, 2u + 1 in [64, 128)), each multiplied under its own row mask: half as many trips
    // through the ~2200-cycle hand-off chain for the layers of the largest level.
    constexpr bool PAIR = LPR == 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2, 7, 0);
    const int K = gt.volume;                                // 8 or 27: one lane per offset below
    const int chw = kTcChunk * NA;                          // channels per step
    const int nchunk = (c_in + chw - 1) / chw;              // steps per offset
    const int nchunk32 = (c_in + kTcChunk - 1) / kTcChunk;  // packed weight chunks per offset
    const uint32_t a_stage = (uint32_t)NA * kTcAStage, b_stage = (uint32_t)NA * NT * 128;
    const int64_t tile0 = (int64_t)blockIdx.x * TPC, row0 = tile0 * 128;
    const int n_mt = (TPC == 2 && gt.n_out - row0 > 128) ? 2 : 1;  // tiles of this CTA that hold rows
    const uint32_t tmem_cols = (uint32_t)tc_tmem_cols(NT, TPC * ACC);  // accumulator (set ai, tile mt) at column (ai TPC + mt) NT

    if (tid == 0) {
        for (int i = 0; i < SA; ++i) { mbar_init(a_full + i, 32 * (4 / GW)); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, n_mt); }
        mbar_init(d_full, n_mt * ACC);
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr, tmem_cols);
    // every warp: lane k holds the row masks of offset k for the CTA's tiles; the offsets any tile uses = the live set
    uint4 mk0 = make_uint4(0, 0, 0, 0), mk1 = mk0;
    if (lane < K) {
        mk0 = __ldg(gt.tm + tile0 * K + lane);
        if (n_mt == 2) mk1 = __ldg(gt.tm + (tile0 + 1) * K + lane);
    }
    if (BN && tid < NT) {
        const float mu = X.bn_mean[tid], sc = X.bn_invstd[tid] * X.bn_weight[tid];
        cst[tid] = sc;
        cst[NT + tid] = X.bn_bias[tid] - mu * sc;
        cst[2 * NT + tid] = mu;
    }
    if (tid == 0) *final_flag = 1;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp < 4 * n_mt) {  // zero the accumulators (every MMA accumulates under a row mask): warp = (tile, TMEM lane quarter)
        for (int ai = 0; ai < ACC; ++ai) {
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((ai * TPC + (warp >> 2)) * NT);
            for (int q = 0; q < NT / 16; ++q) tmem_st16_zero(taddr + 16 * q);
        }
        tmem_st_wait();
    }
    // (the masks are consumed only here: their load latency runs under the first barrier and the TMEM zeroing)
    uint32_t liveset = __ballot_sync(0xffffffffu, (mk0.x | mk0.y | mk0.z | mk0.w | mk1.x | mk1.y | mk1.z | mk1.w) != 0u);
    if (PAIR) {  // from here on a bit of the live set is a UNIT of the pipeline: the pair of offsets (2u, 2u + 1)
        const uint32_t t = (liveset | (liveset >> 1)) & 0x15555555u;
        uint32_t units = 0;
#pragma unroll
        for (int u = 0; u < 14; ++u) units |= ((t >> (2 * u)) & 1u) << u;
        liveset = units;
    }
    if (S > 1) {  // this CTA's share: live offsets number blockIdx.y, blockIdx.y + S, ...
        uint32_t mine = 0;
        int r = 0;
        for (uint32_t rest = liveset; rest; rest &= rest - 1, ++r)
            if (r % S == (int)blockIdx.y) mine |= rest & (0u - rest);
        liveset = mine;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    // everything above read only the tile rulebook and per-channel constants (older than the previous kernel in the stream);
    // from here on the activations / gradients it wrote are read
    pdl_wait();
    pdl_trigger();
    TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2, 7, 1);

    if (warp < 8) {
        // ================================================================= gather warps
        const int mt = warp >> 2, wq = warp & 3;
        const bool live = mt < n_mt;
        const uint4 mk = mt ? mk1 : mk0;
        const int cnt = __popc(mk.x) + __popc(mk.y) + __popc(mk.z) + __popc(mk.w);  // lane k: rules of my tile at offset k
        const int32_t *tl_tile = gt.tl + (((tile0 + mt) * K) << 7);
        constexpr int RPP = 32 / LPR;           // rows per copy instruction
        const int cl = lane & (LPR - 1);        // this lane's 16-byte piece of a row
        const int rl = lane / LPR;              // this lane's row slot inside a pass
        const char *in_c = reinterpret_cast<const char *>(in) + 16 * cl;
        const uint32_t ldb = (uint32_t)ld_in * 4;  // row pitch in bytes (feature matrices are far below 4 GB)
        const uint32_t sA_a = smem_u32(sA), full0_a = smem_u32(a_full), empty0_a = smem_u32(a_empty);
        // The tile's stage-steps g = 0, 1, ... (live offsets in order x chunks) use stages mt + TPC g round the ring. The
        // tile's four gather warps form GW groups of CW = 4 / GW warps: group gi fills the stage-steps g = gi (mod GW), and
        // inside a group warp mi copies the passes j = mi (mod CW) of each of its stage-steps (a pass = 32 / LPR rows). GW
        // divides the stages per tile, so a stage is always filled by the SAME warps: their a_empty waits are never more
        // than one phase ahead of the barrier (mbarrier waits only tell odd from even phases; warps taking turns on a
        // stage could run two uses ahead and alias). GW = 4: one warp per stage-step (sparse tiles, ~15 rules per step);
        // GW = 1: all four warps share every stage-step (dense tiles: a step's copy time is what bounds the ring).
        const int CW = 4 / GW, gi = wq % GW, mi = wq / GW;
        uint32_t rest = live ? liveset : 0u;  // live offsets from the current stage-step's on
        int ch = gi;                           // chunk of the current stage-step
        while (rest && ch >= nchunk) { ch -= nchunk; rest &= rest - 1; }
        int st = mt + TPC * gi;
        uint32_t ph = 1;  // parity of the a_empty wait of the next stage to fill
        while (st >= SA) { st -= SA; ph ^= 1; }
        // List entries are fetched kTcAhead of this warp's stage-steps before they are used (register ring with static
        // indices; [0, 32) always, [32, 64) when the step has them): a stage-step is shorter than an L2 / HBM round trip,
        // and the device timeline showed the warp stalling on this load when it ran only one step ahead.
        auto advance = [&](uint32_t &r, int &c) {
            c += GW;
            while (r && c >= nchunk) { c -= nchunk; r &= r - 1; }
        };
        int en[kTcAhead], en1[kTcAhead], eb[kTcAhead], eb1[kTcAhead];  // (eb*: the second offset of a pair)
        uint32_t rest_p = rest;  // prefetch cursor
        int ch_p = ch;
        auto prefetch = [&](int &e, int &e1, int &f, int &f1) {
            if (!rest_p) return;
            const int un = __ffs(rest_p) - 1, kn = PAIR ? 2 * un : un;
            const int32_t *src = tl_tile + (kn << 7) + lane;
            e = __ldg(src);
            if (__shfl_sync(0xffffffffu, cnt, kn) > 32) e1 = __ldg(src + 32);
            if (PAIR) {
                const int nb2 = __shfl_sync(0xffffffffu, cnt, kn + 1);  // (lanes >= K hold 0)
                if (nb2 > 0) f = __ldg(src + 128);
                if (nb2 > 32) f1 = __ldg(src + 160);
            }
            advance(rest_p, ch_p);
        };
#pragma unroll
        for (int u = 0; u < kTcAhead; ++u) { en[u] = en1[u] = eb[u] = eb1[u] = 0; prefetch(en[u], en1[u], eb[u], eb1[u]); }
        int tstep = 0;  // trace builds only
        while (rest) {
#pragma unroll
          for (int u = 0; u < kTcAhead; ++u) {
            if (!rest) break;  // warp-uniform
            const int k = PAIR ? 2 * (__ffs(rest) - 1) : __ffs(rest) - 1;
            const int n = __shfl_sync(0xffffffffu, cnt, k);  // rules of my tile at this offset (0: only the other tile has some)
            const int n2 = PAIR ? __shfl_sync(0xffffffffu, cnt, k + 1) : 0;  // ... at the pair's second offset
            const int e0 = en[u], e1 = en1[u], f0 = eb[u], f1 = eb1[u];
            prefetch(en[u], en1[u], eb[u], eb1[u]);
            const int cho = ch * (chw * 4);                  // byte offset of this step's channels in a feature row
            const int left = c_in - ch * chw;
            const bool has0 = cl < min(kTcChunk, left) / 4;                                 // this lane's piece exists in atom 0
            const bool has1 = NA == 2 && cl < max(0, min(kTcChunk, left - kTcChunk)) / 4;  // ... in atom 1
            advance(rest, ch);
            TC_STAMP(warp == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 0, tstep);
            mbar_wait_s(empty0_a + 8 * st, ph);
            TC_STAMP(warp == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 1, tstep);
            const uint32_t tile_a = sA_a + st * a_stage;
            // one block of up to 32 list entries (held one per lane in `e`): pass j copies rows j RPP .. j RPP + RPP - 1
            // `half`: 0, or 4 for the second offset of a pair (16-byte pieces 4..7 of the row)
            auto copy_pass = [&](int e, int nb, int j, uint32_t half) {
                const int idx = j * RPP + rl;
                const int ent = __shfl_sync(0xffffffffu, e, idx);
                const uint32_t r = (uint32_t)ent >> kTileRowShift;
                const char *src = in_c + (uint64_t)((uint32_t)ent & ((1u << kTileRowShift) - 1u)) * (uint64_t)ldb + cho;
                const uint32_t dst = tile_a + r * 128u + ((((uint32_t)cl + half) ^ (r & 7u)) << 4);
                const bool ok = idx < nb;
                cp_async16_guard_s(dst, src, (ok && has0) ? 1u : 0u);
                if (NA == 2) cp_async16_guard_s(dst + kTcAStage, src + 128, (ok && has1) ? 1u : 0u);
            };
            auto copy_block = [&](int e, int nb, uint32_t half) {  // nb >= 1, warp-uniform; this warp's passes: mi, mi + CW, ...
                for (int j = mi; j * RPP < nb; j += CW) copy_pass(e, nb, j, half);
            };
            if (n > 0) copy_block(e0, min(n, 32), 0u);
            if (n > 32) copy_block(e1, min(n - 32, 32), 0u);
            for (int base = 64; base < n; base += 32)  // dense tiles and the centre offset
                copy_block(__ldg(tl_tile + (k << 7) + base + lane), min(n - base, 32), 0u);
            if (PAIR) {
                if (n2 > 0) copy_block(f0, min(n2, 32), 4u);
                if (n2 > 32) copy_block(f1, min(n2 - 32, 32), 4u);
                for (int base = 64; base < n2; base += 32)
                    copy_block(__ldg(tl_tile + ((k + 1) << 7) + base + lane), min(n2 - base, 32), 4u);
            }
            // every lane: "my copies into this stage have landed" arrives asynchronously (32 CW arrivals complete it)
            cp_async_mbar_arrive_noinc_s(full0_a + 8 * st);
            TC_STAMP(warp == 0 && lane == 0 && blockIdx.x == gridDim.x / 2, 2, tstep);
            ++tstep;
            st += TPC * GW;
            if (st >= SA) { st -= SA; ph ^= 1; }
          }
        }

        // d_input pass in front of a BatchNorm: the tile's 128 rows of that BatchNorm's input x follow the gathered stages
        // through the ring as extra stage-steps, one per 32-column slice and in the same swizzled row layout (fully
        // coalesced copies that land while the last MMAs drain; the issuers never see them). Slice j is copied by group
        // j mod GW into the stage its next stage-step would use, so the stage / parity bookkeeping above carries on.
        const int nslice = (NT + kTcChunk - 1) / kTcChunk;
        if (BN && X.bn_ring > 0 && live) {
            const char *xsrc = reinterpret_cast<const char *>(X.bn_x + (row0 + 128 * mt) * X.ld_bn_x);
            const uint32_t xldb = (uint32_t)X.ld_bn_x * 4;
            const int rows_here = (int)min((int64_t)128, gt.n_out - (row0 + 128 * mt));
            for (int j = gi; j < nslice; j += GW) {
                mbar_wait_s(empty0_a + 8 * st, ph);
                const uint32_t tile_a = sA_a + st * a_stage;
                for (int it = mi; it < 32; it += CW) {
                    const int pidx = it * 32 + lane, r = pidx >> 3, pc = pidx & 7;
                    const bool ok = r < rows_here && kTcChunk * j + 4 * pc < NT;
                    cp_async16_guard_s(tile_a + (uint32_t)r * 128u + (uint32_t)((pc ^ (r & 7)) << 4),
                                       xsrc + (uint64_t)r * xldb + (uint32_t)(128 * j + 16 * pc), ok ? 1u : 0u);
                }
                st += TPC * GW;
                if (st >= SA) { st -= SA; ph ^= 1; }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            named_barrier_sync(2 + mt, 128);  // the tile's four warps: every slice is in shared memory
        }

        // ================================================================= epilogue: TMEM -> HBM, one output row per thread
        TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2, 7, 2);
        if (live) {
            const int64_t row = row0 + 128 * mt + 32 * wq + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * wq) << 16) + (uint32_t)(mt * NT);
            mbar_wait(d_full, 0);
            tc_fence_after_sync();
            TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2, 7, 3);
            auto load_acc = [&](int q, float (&v)[16]) {  // columns [16 q, +16) of this thread's row, the ACC sets in order
                tmem_ld16(taddr + 16 * q, v);  // warp-collective: every lane takes part, also beyond n_out
#pragma unroll 1
                for (int ai = 1; ai < ACC; ++ai) {
                    float u[16];
                    tmem_ld16(taddr + (uint32_t)(ai * TPC * NT) + 16 * q, u);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += u[e];
                }
            };
            bool run = true;
            const float *parts = nullptr;
            if (S > 1) {  // (split launches have one tile per CTA: mt = 0, warps 0-3)
                float *mine = X.partial + ((size_t)(tile0 * S + blockIdx.y) * NT) * 128 + 32 * wq + lane;
#pragma unroll 1
                for (int q = 0; q < NT / 16; ++q) {
                    float v[16];
                    load_acc(q, v);
#pragma unroll
                    for (int e = 0; e < 16; ++e) __stcg(mine + (size_t)(16 * q + e) * 128, v[e]);
                }
                __threadfence();
                tc_bar_sync_128(1);
                if (warp == 0 && lane == 0) {
                    const unsigned t = atomicAdd(X.tickets + tile0, 1u);
                    const int last = t == (unsigned)(S - 1);
                    if (last) X.tickets[tile0] = 0;  // nobody else touches it any more: ready for the next launch
                    *final_flag = last;
                }
                tc_bar_sync_128(1);
                run = *final_flag != 0;
                if (run) {
                    __threadfence();
                    parts = X.partial + ((size_t)tile0 * S * NT) * 128 + 32 * wq + lane;
                }
            }
            float *dst = out + row * ld_out;
            float *s_part = reinterpret_cast<float *>(sB) + (size_t)warp * 2 * NT;  // the weight ring is idle by now (>= 256 NT bytes)
#pragma unroll 1
            for (int q = 0; run && q < NT / 16; ++q) {
                float v[16];
                if (S > 1) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __ldcg(parts + (size_t)(16 * q + e) * 128);
#pragma unroll 1
                    for (int part = 1; part < S; ++part) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] += __ldcg(parts + ((size_t)part * NT + 16 * q + e) * 128);
                    }
                } else {
                    load_acc(q, v);
                }
                TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2 && q == 0, 7, 6);
                if (row < gt.n_out) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float4 o = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        float4 *p = reinterpret_cast<float4 *>(dst + 16 * q + 4 * e);
                        if (gt.accumulate) {
                            const float4 x = *p;
                            o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                            v[4 * e] = o.x; v[4 * e + 1] = o.y; v[4 * e + 2] = o.z; v[4 * e + 3] = o.w;
                        }
                        *p = o;
                    }
                }
                TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2 && q == 0, 7, 7);
                if (stats) {  // warp-uniform: two per-column sums for the BatchNorm next to this convolution
                    // Eight columns at a time (register budget: 56 per thread at three CTAs per SM; sixteen at a time spilled
                    // ~70 bytes per thread into the epilogue's dependency chain).
                    const bool in_range = row < gt.n_out;
                    uint32_t x_a = 0;  // ring mode: shared address of this thread's row in the slice that holds columns 16 q ..
                    if (BN && X.bn_ring > 0) {
                        // slice j = q / 2 went into the stage of step n_j: the first step index >= G congruent to
                        // j mod GW, plus GW (j / GW), G = the stage-steps of the main loop
                        const int G = __popc(liveset) * nchunk, j = q >> 1, g = j % GW;
                        const int nj = G + ((g - G) % GW + GW) % GW + GW * (j / GW);
                        x_a = sA_a + (uint32_t)((mt + TPC * nj) % SA) * a_stage + (uint32_t)(32 * wq + lane) * 128u;
                    }
                    const uint32_t c_a = smem_u32(cst) + 64u * (uint32_t)q;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float a[8], b[8];
                        if (BN) {  // d_input pass: S1 = sum d, S2 = sum (x - mean) d, d = gradient through the ReLU
#pragma unroll
                            for (int e4 = 0; e4 < 2; ++e4) {
                                const int p4 = 2 * h + e4;  // 16-byte piece of this thread's 64-byte row segment
                                float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (X.bn_ring > 0) x4 = tc_lds128(x_a + (uint32_t)((((q & 1) * 4 + p4) ^ (lane & 7)) << 4));
                                else if (in_range) x4 = __ldg(reinterpret_cast<const float4 *>(X.bn_x + row * X.ld_bn_x + 16 * q) + p4);
                                const float4 sc4 = tc_lds128(c_a + 16 * p4), sh4 = tc_lds128(c_a + 4 * NT + 16 * p4),
                                             mu4 = tc_lds128(c_a + 8 * NT + 16 * p4);
                                const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w},
                                            sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w}, mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float x = in_range ? xs[i] : 0.f;  // (rows past the end: whatever the stage holds)
                                    const float y = fmaf(x, sc[i], sh[i]);
                                    const float g = v[8 * h + 4 * e4 + i];
                                    const float d = in_range ? (y > 0.f ? g : g * X.leak) : 0.f;
                                    a[4 * e4 + i] = d;
                                    b[4 * e4 + i] = (x - mu[i]) * d;
                                }
                            }
                        } else {  // forward: sum x, sum x^2 of the rows just written
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                a[e] = in_range ? v[8 * h + e] : 0.f;
                                b[e] = a[e] * a[e];
                            }
                        }
                        TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2 && q == 0, 7, 8 + 2 * h);
                        // butterfly: after strides 16, 8, 4 a lane holds ONE column's sum over 8 of the 32 rows (column =
                        // bits 4..2 of the lane, msb first); strides 2 and 1 add the four quarters
#pragma unroll
                        for (int w = 4, stride = 16; w >= 1; w >>= 1, stride >>= 1) {
                            const bool up = lane & stride;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (e < w) {
                                    const float sa = up ? a[e] : a[e + w], ka = up ? a[e + w] : a[e];
                                    const float sb2 = up ? b[e] : b[e + w], kb = up ? b[e + w] : b[e];
                                    a[e] = ka + __shfl_xor_sync(0xffffffffu, sa, stride);
                                    b[e] = kb + __shfl_xor_sync(0xffffffffu, sb2, stride);
                                }
                            }
                        }
                        a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
                        b[0] += __shfl_xor_sync(0xffffffffu, b[0], 2);
                        a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                        b[0] += __shfl_xor_sync(0xffffffffu, b[0], 1);
                        TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2 && q == 0, 7, 9 + 2 * h);
                        if (!(lane & 3)) {
                            const int col = 16 * q + 8 * h + (((lane >> 4) & 1) << 2) + (((lane >> 3) & 1) << 1) + ((lane >> 2) & 1);
                            s_part[col] = a[0];
                            s_part[NT + col] = b[0];
                        }
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ================================================================= weight producer (TMA bulk copies)
        if (lane == 0) {
            int st = 0, ph = 1;
            for (uint32_t rest = liveset; rest; rest &= rest - 1) {
                const int k = __ffs(rest) - 1;
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
        }
    } else if (warp - 9 < TPC * ACC && (warp - 9) / ACC < n_mt) {
        // ================================================================= MMA issuers: warp 9 + mt ACC + ai
        // Issuing one stage-step (proxy fence, descriptors through the uniform datapath, 4 tcgen05.mma, 2 commits) is a
        // ~800-cycle serial chain per warp (device timeline, profiles/r02_tc_timeline.txt) while the tensor pipe itself is
        // nearly idle, so a tile has ACC issuer warps: issuer ai takes the stage-steps g = ai (mod ACC) and accumulates them
        // in ITS OWN TMEM accumulator (no ordering between MMAs of different threads is needed); the epilogue adds the ACC
        // partial accumulators. ACC divides the A stages per tile and the B stages, so a given stage is always consumed by
        // the same issuer and its barrier waits never run two phases ahead.
        // The whole warp runs the (warp-uniform) loops and waits; one elected lane executes the tcgen05 instructions.
        const int mt = (warp - 9) / ACC, ai = (warp - 9) % ACC;
        const uint4 mk = mt ? mk1 : mk0;
        const uint32_t idesc = umma_idesc_tf32(NT);
        const uint64_t desc_hi = umma_desc_sw128(0);  // everything but the start address
        const uint32_t d = tmem_base + (uint32_t)((ai * TPC + mt) * NT);
        uint32_t rest = liveset;
        int ch = ai;
        while (rest && ch >= nchunk) { ch -= nchunk; rest &= rest - 1; }
        int st = mt + TPC * ai, ph = 0, stb = ai, phb = 0, tstep = 0;
        while (st >= SA) { st -= SA; ph ^= 1; }
        while (stb >= SB) { stb -= SB; phb ^= 1; }
        while (rest) {
            const int k = PAIR ? 2 * (__ffs(rest) - 1) : __ffs(rest) - 1;
            const uint32_t m0 = __shfl_sync(0xffffffffu, mk.x, k), m1 = __shfl_sync(0xffffffffu, mk.y, k),
                           m2 = __shfl_sync(0xffffffffu, mk.z, k), m3 = __shfl_sync(0xffffffffu, mk.w, k);
            const bool any = (m0 | m1 | m2 | m3) != 0;  // warp-uniform: does THIS tile have a rule at offset k
            // the pair's second offset (lanes >= K hold empty masks)
            const uint32_t p0 = PAIR ? __shfl_sync(0xffffffffu, mk.x, k + 1) : 0u, p1 = PAIR ? __shfl_sync(0xffffffffu, mk.y, k + 1) : 0u,
                           p2 = PAIR ? __shfl_sync(0xffffffffu, mk.z, k + 1) : 0u, p3 = PAIR ? __shfl_sync(0xffffffffu, mk.w, k + 1) : 0u;
            const bool any2 = (p0 | p1 | p2 | p3) != 0;
            const int left = c_in - ch * chw;
            const int nk0 = min(kTcChunk, left) / 8;                                   // MMAs (K = 8 each) from atom 0
            const int nk1 = NA == 2 ? max(0, min(kTcChunk, left - kTcChunk)) / 8 : 0;  // ... and from atom 1
            ch += ACC;
            while (rest && ch >= nchunk) { ch -= nchunk; rest &= rest - 1; }
            mbar_wait(b_full + stb, phb);
            const uint64_t b_desc = desc_hi | (uint64_t)((smem_u32(sB + (size_t)stb * b_stage) & 0x3FFFFu) >> 4);
            TC_STAMP(warp == 9 && lane == 0 && blockIdx.x == gridDim.x / 2, 4, tstep);
            mbar_wait(a_full + st, ph);
            TC_STAMP(warp == 9 && lane == 0 && blockIdx.x == gridDim.x / 2, 5, tstep);
            fence_proxy_async_smem();  // the gather warps' cp.async writes (generic proxy) -> UMMA reads
            tc_fence_after_sync();
            TC_STAMP(warp == 9 && lane == 0 && blockIdx.x == gridDim.x / 2, 3, tstep);
            const uint64_t a_desc = desc_hi | (uint64_t)((smem_u32(sA + (size_t)st * a_stage) & 0x3FFFFu) >> 4);
            if (elect_one()) {
                if (PAIR) {  // channels [0, 16) of the rows = offset k under its mask, [16, 32) = offset k + 1 under its own
                    if (any)
                        for (int j = 0; j < 2; ++j) umma_tf32_masked(d, a_desc + 2 * j, b_desc + 2 * j, idesc, ~m0, ~m1, ~m2, ~m3);
                    if (any2)
                        for (int j = 2; j < 4; ++j) umma_tf32_masked(d, a_desc + 2 * j, b_desc + 2 * j, idesc, ~p0, ~p1, ~p2, ~p3);
                } else if (any) {
                    for (int j = 0; j < nk0; ++j)  // + 32 bytes of K per MMA = + 2 in the address field
                        umma_tf32_masked(d, a_desc + 2 * j, b_desc + 2 * j, idesc, ~m0, ~m1, ~m2, ~m3);
                    for (int j = 0; j < nk1; ++j)  // second atom: 16 KB further in A, NT x 128 bytes further in B
                        umma_tf32_masked(d, a_desc + (kTcAStage >> 4) + 2 * j, b_desc + (uint64_t)((NT * 128) >> 4) + 2 * j,
                                         idesc, ~m0, ~m1, ~m2, ~m3);
                }
                umma_commit(a_empty + st);
                umma_commit(b_empty + stb);  // n_mt arrivals (one per tile) release the weight stage
            }
            __syncwarp();
            TC_STAMP(warp == 9 && lane == 0 && blockIdx.x == gridDim.x / 2, 6, tstep);
            ++tstep;
            st += TPC * ACC;
            if (st >= SA) { st -= SA; ph ^= 1; }
            stb += ACC;
            if (stb >= SB) { stb -= SB; phb ^= 1; }
        }
        if (elect_one()) umma_commit(d_full);  // n_mt * ACC arrivals complete it
        __syncwarp();
    }
    TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2, 7, 4);
    tc_fence_before_sync();
    __syncthreads();
    TC_STAMP(tid == 0 && blockIdx.x == gridDim.x / 2, 7, 5);
    if (warp == 9) tmem_dealloc(tmem_base, tmem_cols);
    if (stats && *final_flag) {  // [0, NT): first sum, [NT, 2 NT): second; warps of the tiles that hold rows, in order
        const float *s_all = reinterpret_cast<const float *>(sB);
        for (int i = tid; i < 2 * NT; i += (int)blockDim.x) {  // (2 NT = 384 on the widest d_input pass, 352 threads)
            float sum = 0.f;
            for (int w = 0; w < 4 * n_mt; ++w) sum += s_all[(size_t)w * 2 * NT + i];
            atomicAdd(stats + (i < NT ? i : kTcStatsLd + (i - NT)), (double)sum);
        }
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

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Parking space of the split launches: [1024 tickets | partial accumulators], one block per (device, stream), grown on
// demand with the stream-ordered allocator (the old block is freed behind the kernels that may still use it).
constexpr int kTcSplitMaxTiles = 1024;
static int split_workspace(size_t partial_bytes, cudaStream_t s, float **partial, unsigned **tickets) {
    struct Block { int device; cudaStream_t stream; char *p; size_t bytes; };
    static std::mutex mu;
    static std::vector<Block> blocks;
    int dev = 0;
    MOPA_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    Block *b = nullptr;
    for (Block &x : blocks)
        if (x.device == dev && x.stream == s) b = &x;
    if (!b) { blocks.push_back(Block{dev, s, nullptr, 0}); b = &blocks.back(); }
    const size_t need = (size_t)kTcSplitMaxTiles * 4 + partial_bytes;
    if (b->bytes < need) {
        size_t want = (size_t)32 << 20;
        while (want < need) want <<= 1;
        char *fresh = nullptr;
        MOPA_CUDA(cudaMallocAsync((void **)&fresh, want, s));
        MOPA_CUDA(cudaMemsetAsync(fresh, 0, (size_t)kTcSplitMaxTiles * 4, s));
        if (b->p) MOPA_CUDA(cudaFreeAsync(b->p, s));
        b->p = fresh;
        b->bytes = want;
    }
    *tickets = reinterpret_cast<unsigned *>(b->p);
    *partial = reinterpret_cast<float *>(b->p + (size_t)kTcSplitMaxTiles * 4);
    return 0;
}

template <bool BN>
static int tc_configure() {
#define MOPA_TC_ATTR(NA_, LPR_, W_) \
    MOPA_CUDA(cudaFuncSetAttribute(k_conv_tc<NA_, LPR_, W_, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    MOPA_TC_ATTR(1, 4, 0); MOPA_TC_ATTR(1, 8, 0); MOPA_TC_ATTR(2, 8, 0);
    MOPA_TC_ATTR(1, 4, 1); MOPA_TC_ATTR(1, 8, 1); MOPA_TC_ATTR(2, 8, 1);
    MOPA_TC_ATTR(1, 4, 2); MOPA_TC_ATTR(1, 8, 2);
#undef MOPA_TC_ATTR
    return 0;
}

int conv_apply_tc(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out, const float *packed,
                  int c_in, int c_out, double *stats, cudaStream_t s, const TcBnBwd *bn) {
    const int nt = c_out;
    MOPA_CHECK(gt.tl && gt.tm, "conv_tc: the gather has no tile rulebook");
    MOPA_CHECK(!bn || stats, "conv_tc: BatchNorm-backward sums requested without a statistics block");
    // Shape of the CTA (all overridable for sweeps: MOPA_TC_{CTAS,TPC,SA,SB,NA,ACC}):
    //   tpc  tiles (128 output rows) per CTA: 2 on the large levels (the weight tiles are streamed per CTA), 1 where
    //        256-row CTAs would leave SMs empty
    //   na   32-channel atoms per stage-step: 2 on the small levels (half the steps where the per-step latency dominates)
    //   acc  issuer warps / TMEM accumulator sets per tile (the issue chain of a stage-step is the serial bottleneck)
    //   sa   A stages per CTA (sx = sa / tpc per tile), sb weight stages; acc and the gather warps per tile (gw) must
    //        divide sx, acc must divide sb: every stage then always meets the same producer and consumer warp
    static const int want_ctas = env_int("MOPA_TC_CTAS", 3), want_sb = env_int("MOPA_TC_SB", 4), want_tpc = env_int("MOPA_TC_TPC", 0);
    static const int want_sa = env_int("MOPA_TC_SA", 0), want_na = env_int("MOPA_TC_NA", 0), want_acc = env_int("MOPA_TC_ACC", 2);
    const int sms = num_sms();
    const int tpc = want_tpc ? want_tpc : (ceil_div(gt.n_out, kTcTM) > sms ? 2 : 1);
    const int na = want_na ? want_na : ((tpc == 1 && c_in >= 64 && ceil_div(gt.n_out, 128) <= sms) ? 2 : 1);
    bool two = want_ctas >= 2 && (nt <= 64 || tpc == 1) && ceil_div(gt.n_out, 128 * tpc) > sms;  // two CTAs per SM
    // three CTAs per SM (MOPA_TC_CTAS=3): narrow layers of the large levels, two stages per tile, one issuer per tile
    static const int bn_no3 = env_int("MOPA_TC_BN_NO3", 0);  // A/B: no three-CTA shape for launches that carry BatchNorm sums
    const bool three = want_ctas >= 3 && two && tpc == 2 && na == 1 && nt <= 32 && ceil_div(gt.n_out, 256) > 2 * sms &&
                       !(bn && bn_no3) &&
                       (size_t)tc_smem_layout(nt, 4, 2, 1, bn != nullptr).total + 1024 <= (size_t)75 * 1024;
    int acc = 1, sb = 2, sa = 0;
    size_t cap = 0;
    for (;;) {
        if (three) { acc = 1; sb = 2; sa = 4; cap = (size_t)75 * 1024; break; }
        cap = two ? (size_t)113 * 1024 : (size_t)226 * 1024;
        acc = want_acc >= 4 ? 4 : (want_acc >= 2 ? 2 : 1);
        while (acc > 1 && (tc_tmem_cols(nt, tpc * acc) > (two ? 256 : 512) || 32 * (9 + tpc * acc) > (two ? 13 * 32 : kTcMaxThreads))) acc >>= 1;
        sb = want_sb >= 4 ? 4 : 2;
        // weight ring budget: a CTA that has the SM to itself (the small levels) may spend up to 120 KB on it, so that wide
        // layers (28 KB per weight stage) still get two stages per issuer and the TMA loads run one step ahead
        // (level-5 convolutions 39 -> 35, 47 -> 41, 36 -> 30 us); shared SMs keep the ring small
        static const int bcap_kb = env_int("MOPA_TC_BCAP_KB", 120);
        const size_t bcap = !two ? (size_t)bcap_kb * 1024 : (size_t)(tpc == 2 ? 64 : 32 * na) * 1024;
        while (sb > 2 && (size_t)sb * na * nt * 128 > bcap) sb -= 2;
        if (acc > sb) acc = sb;
        const int unit = tpc * acc;  // sa must be a multiple of it
        sa = want_sa >= unit && want_sa <= kTcMaxSA ? want_sa : kTcMaxSA;
        sa -= sa % unit;
        while (sa > unit && (size_t)tc_smem_layout(nt, sa, sb, na, bn != nullptr).total + 1024 > cap) sa -= unit;
        const bool fits = (size_t)tc_smem_layout(nt, sa, sb, na, bn != nullptr).total + 1024 <= cap && tc_tmem_cols(nt, tpc * acc) <= (two ? 256 : 512);
        if (fits && (!two || sa >= 2 * tpc)) break;
        if (two) { two = false; continue; }  // too little room at two CTAs per SM: take the whole SM
        MOPA_CHECK(fits, "conv_tc: shared memory / TMEM layout does not fit");
        break;
    }
    const size_t smem = (size_t)tc_smem_layout(nt, sa, sb, na, bn != nullptr).total + 1024;
    static std::atomic<uint64_t> configured{0};
    MOPA_TRY(once_per_device(configured, [] {
        MOPA_TRY((tc_configure<false>()));
        return tc_configure<true>();
    }));
    const int sx = sa / tpc;
    static const int want_gw = env_int("MOPA_TC_GW", 2);  // gather groups per tile (1: all four warps share every stage-step)
    int gw = want_gw >= 4 ? 4 : (want_gw >= 2 ? 2 : 1);
    while (gw > 1 && sx % gw) gw >>= 1;
    const unsigned threads = 32u * (unsigned)(9 + tpc * acc);
    dim3 grid((unsigned)ceil_div(gt.n_out, 128 * tpc));
    const bool wide = threads > 13 * 32;
    TcExtra X{};
    X.split = 1;
    if (bn) {
        X.bn_x = bn->x; X.ld_bn_x = bn->ld_x; X.bn_mean = bn->mean; X.bn_invstd = bn->invstd; X.bn_weight = bn->weight;
        X.bn_bias = bn->bias; X.leak = bn->leak;
        static const int want_ring = env_int("MOPA_TC_BN_RING", 1);
        X.bn_ring = want_ring > 0 && ceil_div(nt, kTcChunk) + gw - 1 <= sx;
        if (want_ring < 0) X.bn_ring = -1;  // A/B: row-by-row loads, no L2 prefetch either
    }
    // offsets split over several CTAs per tile where the tiles alone cannot fill the SMs (MOPA_TC_SPLIT=0 disables,
    // n > 1 forces n parts): as many parts as fit in one wave, at least ~3 live offsets each
    static const int want_split = env_int("MOPA_TC_SPLIT", -1);
    if (want_split != 0 && tpc == 1 && !two && (int)grid.x <= kTcSplitMaxTiles) {
        int parts = want_split > 0 ? want_split : sms / (int)grid.x;
        if (parts > gt.volume / 3) parts = gt.volume / 3;
        if (parts > 1) {
            MOPA_TRY(split_workspace((size_t)grid.x * parts * nt * 128 * 4, s, &X.partial, &X.tickets));
            X.split = parts;
            grid.y = (unsigned)parts;
        }
    }
    const bool bnk = X.bn_x != nullptr;
#define MOPA_TC_GO(NA_, LPR_, W_)                                                                                              \
    do {                                                                                                                       \
        if (bnk)                                                                                                               \
            MOPA_CUDA(launch_pdl(k_conv_tc<NA_, LPR_, W_, true>, grid, dim3(threads), smem, s, gt, in, ld_in, out, ld_out,    \
                                 packed, c_in, nt, sa, sb, tpc, gw, acc, stats, X));                                           \
        else                                                                                                                   \
            MOPA_CUDA(launch_pdl(k_conv_tc<NA_, LPR_, W_, false>, grid, dim3(threads), smem, s, gt, in, ld_in, out, ld_out,   \
                                 packed, c_in, nt, sa, sb, tpc, gw, acc, stats, X));                                           \
    } while (0)
#define MOPA_TC_LAUNCH(NA_, LPR_)               \
    do {                                        \
        if (three) MOPA_TC_GO(NA_, LPR_, 2);    \
        else if (wide) MOPA_TC_GO(NA_, LPR_, 1); \
        else MOPA_TC_GO(NA_, LPR_, 0);          \
    } while (0)
    if (na == 2) {  // (never together with `three`)
        if (wide) MOPA_TC_GO(2, 8, 1);
        else MOPA_TC_GO(2, 8, 0);
    } else if (c_in == 16) MOPA_TC_LAUNCH(1, 4);
    else MOPA_TC_LAUNCH(1, 8);
#undef MOPA_TC_LAUNCH
#undef MOPA_TC_GO
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
