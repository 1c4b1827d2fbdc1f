// d_weight of the sparse convolutions on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only, TF32 mode.
// Replaces the weight-gradient half of [UPSTREAM] SparseConvNet SCN/CUDA/Convolution.cu (dConvolution_KMxKN_backward_dW*,
// reached from mopa/train/train_xmuda_mopa.py:417-418 through the autograd of mopa/models/scn_unet.py:27-28).
//
//   d_weight[k] (C_in x C_out) = sum over the rules (i -> o) of offset k of  in[i]^T  d_out[o]
//
// This is the one contraction of the path whose reduction axis IS the rule list, so the rules can be compacted: a
// pipeline stage holds 32 live rules and nothing else (the forward kernel has to carry zero rows for absent neighbours).
// Both operands are "MN-major" for the tensor core: a gathered feature row (channels contiguous) is one K-row of the
// UMMA operand tile, so cp.async drops rows straight from HBM/L2 into the operand layout; no transpose, no registers.
// For 32-bit operands the only MN-major layout the hardware accepts is SWIZZLE_128B_BASE32B (every other layout type
// silently multiplies by zero); its element map was decoded on a B200 with scratch/umma_mn_probe2.cu:
//   (mn, k) -> (mn/32) LBO + (k/4) SBO + (k%4) 128 + (((mn%32)/8) ^ (k%4)) 32 + (mn%8) 4      [bytes]
// i.e. atoms of 4 K-rows (rules) x 128 bytes (32 channels), 32-byte chunks XOR-swizzled by the row.
//
//   work item  : (offset k, range of output rows). The centre offset of a submanifold filter has one rule per row
//                (6-12x the others) and is split 8x finer, so all items carry a similar number of rules.
//   warps 0-3  : producers. Look up 512 rows per pass (coalesced table reads, prefetched one pass ahead), compact the
//                live rules in row order with ballots + a 16-entry scan, append (m_row, n_row) to a pending ring, and for
//                every 32 pending rules fill the next ring stage with cp.async (4 threads per rule, 16-byte pieces);
//                completion is signalled asynchronously (cp.async.mbarrier.arrive.noinc): nobody waits for data.
//   warp  4    : one elected lane issues 4 x tcgen05.mma (M = 128, N = C_n, K = 8) per stage and commits the stage
//                release; the accumulator (C_m lanes x C_n columns, fp32) stays in TMEM for the whole item.
//   epilogue   : warps 0-3 read TMEM (lane = m) and write the item's partial d_weight slice; k_dw_tc_reduce sums the
//                slices of each offset in index order (deterministic; no float atomics).
// M is always 128 for the hardware; rows >= C_m of the A tile alias whatever follows in shared memory and only feed
// accumulator lanes that are never read. The larger channel count sits on M when it is <= 128 (MMA time ~ N).
#include <stdlib.h>

#include "geometry.cuh"
#include "mopa_scn.h"
#include "ptx.cuh"

namespace mopa {

constexpr int kDwTcThreads = 5 * 32;
constexpr int kDwTcTile = 32;    // rules per stage
constexpr int kDwTcSub = 512;    // granularity of an item's row range (a multiple of the 128-row rulebook tiles)
constexpr int kDwTcMaxTiles = 2048;  // rulebook tiles per item (prefix counts in shared memory); caps rows per item
constexpr int kDwTcList = kDwTcMaxTiles / 2 + 4;  // 8-byte slots of the prefix block (kDwTcMaxTiles + 1 ints)
constexpr int kDwTcMaxStages = 12;
constexpr int kDwTcAhead = 8;  // stages between fetching a rule's list entry and copying its rows

struct DwTcPlan {
    int centre;  // 13 for a submanifold table, -1 otherwise
    int rpi;     // rows per item, ordinary offsets
    int n_o;     // items per ordinary offset
    int rpi_c;   // rows per item, centre offset
    int n_c;     // items of the centre offset
    int items;
};
__host__ __device__ inline DwTcPlan dw_tc_plan(int volume, bool subm_table, int64_t n_rows) {
    DwTcPlan p;
    const int64_t rows = n_rows > 0 ? n_rows : 1;
    const int target = 4 * kNumSMs;
    p.centre = (subm_table && volume == 27) ? 13 : -1;
    const int groups = p.centre >= 0 ? 26 + 8 : volume;
    int64_t rpi = round_up(ceil_div(rows * groups, target), kDwTcSub);
    if (rpi < 1024) rpi = 1024;
    if (rpi > (int64_t)(kDwTcMaxTiles - 1) * 128) rpi = (int64_t)(kDwTcMaxTiles - 1) * 128;  // prefix block of the kernel
    p.rpi = (int)rpi;
    p.n_o = (int)ceil_div(rows, rpi);
    if (p.centre >= 0) {
        int64_t rc = round_up(ceil_div(rpi, 8), 128);
        if (rc < 512) rc = 512;
        p.rpi_c = (int)rc;
        p.n_c = (int)ceil_div(rows, rc);
        p.items = p.n_c + 26 * p.n_o;
    } else {
        p.rpi_c = 0;
        p.n_c = 0;
        p.items = volume * p.n_o;
    }
    return p;
}

struct DwTcSmem {
    int a, b, list, cnt, info, bars, total;
};
__host__ __device__ inline DwTcSmem dw_tc_smem(int n_ma, int n_na, int stages) {
    DwTcSmem L;
    L.a = 0;
    L.b = L.a + stages * n_ma * 4096;
    L.list = L.b + stages * n_na * 4096 + 4096;  // + slack (the MMA always reads 4 M atoms = 2 KB from every group base)
    L.cnt = L.list + kDwTcList * 8;
    L.info = L.cnt + 16 * 4;
    L.bars = L.info + kDwTcMaxStages * 4 + 16;
    L.total = L.bars + 8 * (2 * kDwTcMaxStages + 1) + 16;
    return L;
}

// MN-major 32-bit operand, SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 K-rows x 128 bytes (32 channels); the atoms of
// one 4-rule group are contiguous (LBO = 512), consecutive 4-rule groups `sbo_bytes` apart
__device__ __forceinline__ uint64_t umma_desc_mn_b32(uint32_t sbo_bytes) {
    return ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int n) {  // both operands MN-major, M = 128, N = n
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

// srcM / srcN: the feature matrices that feed the M and N side. swap = 0: M = in (row i), N = d_out (row o), D = d_weight;
// swap = 1: M = d_out (row o), N = in (row i), D = d_weight^T.
__global__ void __launch_bounds__(kDwTcThreads)
    k_dw_tc(Gather gt, const float *__restrict__ in, int64_t ld_in, const float *__restrict__ dout, int64_t ld_dout,
            int n_in, int n_out, int swap, DwTcPlan plan, int stages, float *__restrict__ partial) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int c_m = swap ? n_out : n_in, c_n = swap ? n_in : n_out;
    const int n_ma = (c_m + 31) >> 5, n_na = (c_n + 31) >> 5;
    const DwTcSmem L = dw_tc_smem(n_ma, n_na, stages);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bars), *empty = full + kDwTcMaxStages;
    uint64_t *d_full = empty + kDwTcMaxStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(d_full + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- work item -> (k, rows)
    int k;
    int64_t r_begin, r_end;
    {
        const int item = blockIdx.x;
        if (plan.centre >= 0) {
            if (item < plan.n_c) {
                k = plan.centre;
                r_begin = (int64_t)item * plan.rpi_c;
                r_end = r_begin + plan.rpi_c;
            } else {
                const int kk = (item - plan.n_c) / plan.n_o, j = (item - plan.n_c) - kk * plan.n_o;
                k = kk < plan.centre ? kk : kk + 1;
                r_begin = (int64_t)j * plan.rpi;
                r_end = r_begin + plan.rpi;
            }
        } else {
            k = item / plan.n_o;
            r_begin = (int64_t)(item - k * plan.n_o) * plan.rpi;
            r_end = r_begin + plan.rpi;
        }
        if (r_end > gt.n_out) r_end = gt.n_out;
    }

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < c_n) tmem_cols <<= 1;
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(full + i, 129); mbar_init(empty + i, 1); }
        mbar_init(d_full, 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
        // ================================================================= producers
        const uint32_t a0 = smem_u32(smem + L.a), b0 = smem_u32(smem + L.b);
        const uint32_t list_a = smem_u32(smem + L.list), cnt_a = smem_u32(smem + L.cnt), info_a = smem_u32(smem + L.info);
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const uint32_t stage_a = (uint32_t)n_ma * 4096, stage_b = (uint32_t)n_na * 4096;
        const float *src_m = swap ? dout : in, *src_n = swap ? in : dout;
        const int64_t ld_m = swap ? ld_dout : ld_in, ld_n = swap ? ld_in : ld_dout;
        const int pm = c_m >> 2, pn = c_n >> 2;  // 16-byte pieces per row
        const int rt = tid >> 2, sub4 = tid & 3;  // rule of the tile this thread copies, piece phase
        // byte offset of this thread's rule row inside a stage (4-rule group, row in group); 32-byte chunks XOR (rt & 3)
        const uint32_t row_a = (uint32_t)(rt >> 2) * (uint32_t)n_ma * 512 + (uint32_t)(rt & 3) * 128;
        const uint32_t row_b = (uint32_t)(rt >> 2) * (uint32_t)n_na * 512 + (uint32_t)(rt & 3) * 128;
        const uint32_t sw = (uint32_t)(rt & 3);
        const uint32_t off0 = ((((uint32_t)sub4 >> 1) ^ sw) << 5) + (((uint32_t)sub4 & 1) << 4);        // piece sub4
        const uint32_t off1 = (((((uint32_t)sub4 >> 1) + 2) ^ sw) << 5) + (((uint32_t)sub4 & 1) << 4);  // piece sub4 + 4
        const char *src_m_c = reinterpret_cast<const char *>(src_m) + 16 * sub4;
        const char *src_n_c = reinterpret_cast<const char *>(src_n) + 16 * sub4;
        const uint32_t ldm_b = (uint32_t)ld_m * 4, ldn_b = (uint32_t)ld_n * 4;
        int st = 0;
        uint32_t ph = 1;

        auto emit = [&](uint32_t m_row, uint32_t n_row, uint32_t pad, uint32_t last) {
            // pad: this thread's rule slot lies beyond the end of the item: zero rows (ignore-src), rows 0 are never read
            mbar_wait_s(empty_a + 8 * st, ph);
            // 32-bit row offsets (a feature matrix is far below 4 GB); this thread copies pieces sub4 and sub4 + 4 of every
            // 32-channel atom: their swizzled offsets (off0, off1) do not depend on the atom
            const char *gm = src_m_c + (uint64_t)m_row * (uint64_t)ldm_b;
            const char *gn = src_n_c + (uint64_t)n_row * (uint64_t)ldn_b;
            uint32_t ta = a0 + (uint32_t)st * stage_a + row_a, tb = b0 + (uint32_t)st * stage_b + row_b;
            for (int p8 = 0; p8 < pm; p8 += 8) {
                cp_async16_zfill_s(ta + off0, gm, pad);
                if (p8 + 4 < pm) cp_async16_zfill_s(ta + off1, gm + 64, pad);
                ta += 512;
                gm += 128;
            }
            for (int p8 = 0; p8 < pn; p8 += 8) {
                cp_async16_zfill_s(tb + off0, gn, pad);
                if (p8 + 4 < pn) cp_async16_zfill_s(tb + off1, gn + 64, pad);
                tb += 512;
                gn += 128;
            }
            if (tid == 0) {
                sts_u32(info_a + 4 * st, last);
                mbar_arrive_s(full_a + 8 * st);  // release: publishes the flag
            }
            cp_async_mbar_arrive_noinc_s(full_a + 8 * st);
            if (++st == stages) { st = 0; ph ^= 1; }
        };

        // The item's rules come straight from the tile rulebook (geometry.cu::k_tile_lists_batch): per tile of 128 output rows a
        // compact, row-ordered list of this offset's rules. The producers first put the running rule counts of the item's
        // tiles into shared memory; rule g of the item is then entry g - prefix[j] of tile j, and because a thread's g grows
        // by 32 per stage it finds j with a running pointer. No table lookups, ballots or block barriers per pass.
        const int K = gt.volume;
        const int64_t t0 = r_begin >> 7;
        const int ntile = r_end > r_begin ? (int)(((r_end + 127) >> 7) - t0) : 0;  // <= kDwTcMaxTiles (dw_tc_plan)
        const uint32_t pre_a = list_a;  // prefix[0 .. ntile] (int32), in the block that used to hold the pending-rule ring
        {
            constexpr int PER = kDwTcMaxTiles / 128;  // tiles per thread
            int c[PER], sum = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int t = tid * PER + i;
                c[i] = 0;
                if (t < ntile) {
                    const uint4 m = __ldg(gt.tm + (t0 + t) * K + k);
                    c[i] = __popc(m.x) + __popc(m.y) + __popc(m.z) + __popc(m.w);
                }
                sum += c[i];
            }
            int incl = sum;  // inclusive scan over the 128 producer threads: shuffles inside a warp, warp totals through smem
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) sts_u32(cnt_a + 4 * warp, (uint32_t)incl);
            named_barrier_sync(1, 128);
            int run = incl - sum;
            for (int w = 0; w < warp; ++w) run += (int)lds_u32(cnt_a + 4 * w);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int t = tid * PER + i;
                if (t <= ntile) sts_u32(pre_a + 4 * t, (uint32_t)run);  // t == ntile: the item's rule count
                run += c[i];
            }
            named_barrier_sync(1, 128);
        }
        const int n_rules = ntile > 0 ? (int)lds_u32(pre_a + 4 * ntile) : 0;
        const int n_stage = n_rules > 0 ? (n_rules + kDwTcTile - 1) / kDwTcTile : 1;  // an empty item still zeroes its slice
        const int32_t *tl_k = gt.tl + ((t0 * K + k) << 7);
        // running tile pointer of this thread with the tile's rule range [lo, hi) in registers: the common case (the next
        // rule is in the same tile) touches no shared memory (a dependent LDS chain per stage cost ~25 % of the kernel)
        int j = 0, lo = 0, hi = ntile > 0 ? (int)lds_u32(pre_a + 4) : 0;
        auto locate = [&](int g, int &tile) -> int {  // list entry of rule g (g < n_rules)
            while (g >= hi) {
                ++j;
                lo = hi;
                hi = (int)lds_u32(pre_a + 4 * (j + 1));
            }
            tile = j;
            return __ldg(tl_k + (((int64_t)j * K) << 7) + (g - lo));
        };
        // entries are fetched kDwTcAhead stages before their stage is filled (register ring with static indices): a stage
        // is shorter than an L2 / HBM round trip
        int e_q[kDwTcAhead], t_q[kDwTcAhead];
#pragma unroll
        for (int u = 0; u < kDwTcAhead; ++u) {
            e_q[u] = 0;
            t_q[u] = 0;
            if (rt + kDwTcTile * u < n_rules) e_q[u] = locate(rt + kDwTcTile * u, t_q[u]);
        }
        for (int s0 = 0; s0 < n_stage; s0 += kDwTcAhead) {
#pragma unroll
            for (int u = 0; u < kDwTcAhead; ++u) {
                const int sidx = s0 + u;
                if (sidx < n_stage) {  // uniform over the producers
                    const int g = rt + kDwTcTile * sidx, e_cur = e_q[u], t_cur = t_q[u];
                    const int g_ahead = g + kDwTcTile * kDwTcAhead;
                    if (g_ahead < n_rules) e_q[u] = locate(g_ahead, t_q[u]);
                    const uint32_t in_row = (uint32_t)e_cur & ((1u << kTileRowShift) - 1u);
                    const uint32_t out_row = (uint32_t)((t0 + t_cur) << 7) + ((uint32_t)e_cur >> kTileRowShift);
                    emit(swap ? out_row : in_row, swap ? in_row : out_row, g < n_rules ? 0u : 1u, sidx == n_stage - 1 ? 1u : 0u);
                }
            }
        }

        // ================================================================= epilogue: TMEM -> partial slice
        mbar_wait(d_full, 0);
        tc_fence_after_sync();
        float *dst = partial + (int64_t)blockIdx.x * n_in * n_out;
        const int mrow = 32 * warp + lane;
        if (32 * warp < c_m) {  // warp-uniform
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
            for (int q = 0; q < c_n / 16; ++q) {
                float x[16];
                tmem_ld16(taddr + 16 * q, x);
                if (mrow < c_m) {
                    if (!swap) {  // D[ci][co]
                        float4 *p = reinterpret_cast<float4 *>(dst + (int64_t)mrow * n_out + 16 * q);
#pragma unroll
                        for (int e = 0; e < 4; ++e) p[e] = make_float4(x[4 * e], x[4 * e + 1], x[4 * e + 2], x[4 * e + 3]);
                    } else {  // D[co][ci]
#pragma unroll
                        for (int e = 0; e < 16; ++e) dst[(int64_t)(16 * q + e) * n_out + mrow] = x[e];
                    }
                }
            }
        }
    } else {
        // ================================================================= MMA issuer (warp-uniform control flow)
        const uint32_t idesc = umma_idesc_tf32_mn(c_n);
        const uint32_t kg_a = (uint32_t)n_ma * 1024, kg_b = (uint32_t)n_na * 1024;  // bytes per MMA (8 rules = two 4-rule groups)
        const uint64_t da_hi = umma_desc_mn_b32(kg_a >> 1), db_hi = umma_desc_mn_b32(kg_b >> 1);
        const uint32_t a0 = smem_u32(smem + L.a), b0 = smem_u32(smem + L.b), info_a = smem_u32(smem + L.info);
        int st = 0;
        uint32_t ph = 0, acc = 0;
        for (;;) {
            mbar_wait(full + st, ph);
            fence_proxy_async_smem();  // producers' cp.async writes -> tensor-core (async proxy) reads
            tc_fence_after_sync();
            const uint32_t last = lds_u32(info_a + 4 * st);
            const uint64_t a_desc = da_hi | (uint64_t)(((a0 + (uint32_t)st * kg_a * 4) & 0x3FFFFu) >> 4);
            const uint64_t b_desc = db_hi | (uint64_t)(((b0 + (uint32_t)st * kg_b * 4) & 0x3FFFFu) >> 4);
            __syncwarp();
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < kDwTcTile / 8; ++j)
                    umma_tf32(tmem_base, a_desc + (uint64_t)((kg_a >> 4) * j), b_desc + (uint64_t)((kg_b >> 4) * j), idesc,
                              j > 0 ? 1u : acc);
                umma_commit(empty + st);
            }
            __syncwarp();
            acc = 1;
            if (++st == stages) { st = 0; ph ^= 1; }
            if (last) break;
        }
        if (elect_one()) umma_commit(d_full);
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
}

// d_weight[k][e] = sum of the item slices of offset k. grid (element blocks of 32, offsets), block (32 elements, 8 slice
// groups): thread (x, y) adds slices y, y + 8, ... in order, the 8 partial sums are combined in y order: a fixed summation
// tree (deterministic, no float atomics) with 8x shorter load chains than one thread per element.
__global__ void __launch_bounds__(256) k_dw_tc_reduce(const float *__restrict__ partial, int64_t mat, DwTcPlan plan,
                                                      float *__restrict__ dw) {
    __shared__ float red[8][33];
    const int k = blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * 32 + threadIdx.x;
    int beg, cnt;
    if (plan.centre >= 0) {
        if (k == plan.centre) { beg = 0; cnt = plan.n_c; }
        else { beg = plan.n_c + (k < plan.centre ? k : k - 1) * plan.n_o; cnt = plan.n_o; }
    } else {
        beg = k * plan.n_o;
        cnt = plan.n_o;
    }
    float s0 = 0.f, s1 = 0.f;
    if (e < mat) {
        const float *p = partial + (int64_t)beg * mat + e;
        int c = threadIdx.y;
        for (; c + 8 < cnt; c += 16) {
            s0 += p[(int64_t)c * mat];
            s1 += p[(int64_t)(c + 8) * mat];
        }
        if (c < cnt) s0 += p[(int64_t)c * mat];
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

// ------------------------------------------------------------------------------------------------ host side
bool dw_tc_enabled() {  // MOPA_SCN_NO_DWTC=1 keeps d_weight on the mma.sync kernel (A/B measurements; read per call)
    const char *e = getenv("MOPA_SCN_NO_DWTC");
    return !(e && e[0] == '1');
}
bool dw_tc_supported(int n_in, int n_out) {
    if (n_in % 16 || n_out % 16 || n_in < 16 || n_out < 16) return false;
    const int lo = n_in < n_out ? n_in : n_out, hi = n_in < n_out ? n_out : n_in;
    return lo <= 128 && hi <= 256;
}
size_t dw_tc_workspace_bytes(int volume, int n_in, int n_out, int64_t n_rows) {
    // sized for either table kind (the submanifold plan has the most items)
    const DwTcPlan a = dw_tc_plan(volume, true, n_rows), b = dw_tc_plan(volume, false, n_rows);
    const int items = a.items > b.items ? a.items : b.items;
    return (size_t)items * n_in * n_out * 4;
}

int conv_dweight_tc(const Gather &gt, const float *in, int64_t ld_in, const float *dout, int64_t ld_dout, float *dw,
                    int n_in, int n_out, float *partial, cudaStream_t s) {
    const DwTcPlan plan = dw_tc_plan(gt.volume, gt.table != nullptr, gt.n_out);
    const int hi = n_in < n_out ? n_out : n_in;
    // M side: the larger channel count when it fits the 128 accumulator lanes (MMA time ~ N); ties keep d_weight untransposed
    const int swap = hi <= 128 ? (n_out > n_in) : (n_out < n_in);
    const int c_m = swap ? n_out : n_in, c_n = swap ? n_in : n_out;
    const int n_ma = (c_m + 31) / 32, n_na = (c_n + 31) / 32;
    // The producers are issue-latency bound (ncu: 'wait' + 'selected' > 50 % of their samples), so resident warps matter
    // more than ring depth: as many CTAs per SM (up to 4) as still leave every CTA a ring of >= 3 stages.
    const size_t stage_bytes = (size_t)(n_ma + n_na) * 4096;
    const size_t fixed = (size_t)dw_tc_smem(n_ma, n_na, 0).total + 2048;  // list, flags, barriers, alignment, per-CTA reserve
    int stages = 2;
    for (int ctas = 4; ctas >= 1; --ctas) {
        const size_t per_cta = (size_t)227 * 1024 / ctas;
        if (per_cta <= fixed) continue;
        const int st = (int)((per_cta - fixed) / stage_bytes);
        if (st >= (ctas == 1 ? 2 : 3)) {
            stages = st > kDwTcMaxStages ? kDwTcMaxStages : st;
            break;
        }
    }
    const size_t smem = (size_t)dw_tc_smem(n_ma, n_na, stages).total + 1024;
    static std::atomic<uint64_t> configured{0};
    MOPA_TRY(once_per_device(configured, [] {
        MOPA_CUDA(cudaFuncSetAttribute(k_dw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        return 0;
    }));
    k_dw_tc<<<(unsigned)plan.items, kDwTcThreads, smem, s>>>(gt, in, ld_in, dout, ld_dout, n_in, n_out, swap, plan, stages,
                                                           partial);
    MOPA_LAUNCHED();
    const int64_t mat = (int64_t)n_in * n_out;
    k_dw_tc_reduce<<<dim3((unsigned)ceil_div(mat, 32), gt.volume), dim3(32, 8), 0, s>>>(partial, mat, plan, dw);
    MOPA_LAUNCHED();
    return 0;
}

}  // namespace mopa

extern "C" int mopa_scn_debug_dweightPlan(int volume, int subm_table, int64_t n_rows, int *plan_out) {
    if (!plan_out || (volume != 27 && volume != 8) || n_rows < 0) return 1;
    const mopa::DwTcPlan p = mopa::dw_tc_plan(volume, subm_table != 0, n_rows);
    plan_out[0] = p.centre; plan_out[1] = p.rpi; plan_out[2] = p.n_o; plan_out[3] = p.rpi_c; plan_out[4] = p.n_c;
    plan_out[5] = p.items;
    return 0;
}
