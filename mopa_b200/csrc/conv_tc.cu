// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
// Same contract as k_gather_mma in conv.cu (replaces [UPSTREAM] SparseConvNet SCN/CUDA/Convolution.cu's per-offset
// gather-FMA-scatter kernels reached from mopa/models/scn_unet.py:27-28), TF32 precision mode only.
//
// Why tcgen05 here: not for FLOPs (the layers are HBM-bound) but for INSTRUCTION economy. With mma.sync every gathered
// row travels global -> registers -> (cvt) -> MMA fragments and every product comes back as a register fragment that has
// to be scattered; at ~4 rules per site and offset that costs hundreds of issue slots per 16 rules (ncu: 2.7% of the
// issued instructions were HMMA). Here a gathered row goes global -> shared memory with cp.async (LDGSTS, no registers),
// one elected thread issues the MMA straight from shared memory, and a product row comes back as ONE TMEM lane per
// thread, which that thread adds to its output row.
//
// CTA = 256 output rows, 10 warps:
//   warps 0-3  gather   : warp w owns rows [64w, 64w+64). Per filter offset k it compacts the rows that have a rule
//                         (ballot -> slot = rank), publishes the slot -> (input row, output row) lists, and copies the
//                         slots' input rows (32 channels = 128 bytes per step) into its quarter (32 slots) of the
//                         128-row A tile of a ring stage, in the UMMA K-major SWIZZLE_128B layout. More than 32 rules
//                         per warp and offset (dense clouds; always the centre offset) take a second pass.
//   warp  8    weights  : one lane streams W[k] chunks (pre-packed in the same swizzled K-major layout) by TMA bulk copy.
//   warp  9    MMA      : one lane issues tcgen05.mma (M = 128 slots, N = C_out, K = 8 per instruction) into a TMEM
//                         accumulator (one per offset and pass, NBUF offsets in flight), commits to mbarriers.
//   warps 4-7  epilogue : warp 4+w reads TMEM lanes 32w.. (its slots), and adds each valid slot's product row to the
//                         slot's output row of the CTA's fp32 accumulator tile in shared memory. An output row occurs at
//                         most once per offset and belongs to one warp: plain read-modify-write, fixed order (k ascending).
// Every output row is written to HBM once, coalesced.
#include "geometry.cuh"
#include "mopa_scn.h"
#include "ptx.cuh"

namespace mopa {

constexpr int kTcRW = 64;            // output rows per gather / epilogue warp
constexpr int kTcTM = 4 * kTcRW;     // output rows per CTA
constexpr int kTcThreads = 10 * 32;
constexpr int kTcListBytes = kTcRW * 4 + kTcRW;  // one list: input rows (int32) + output rows inside the CTA tile (uint8)
constexpr int kTcChunk = 32;         // input channels per pipeline step (one 128-byte swizzle row)
constexpr int kTcAStage = 128 * 128; // bytes: 128 slots x 128 bytes

__host__ __device__ constexpr int tc_col_stride(int nt) { return nt <= 16 ? 16 : (nt <= 32 ? 32 : (nt <= 64 ? 64 : 128)); }
__host__ __device__ constexpr int tc_nbuf(int nt) { return nt <= 64 ? 4 : 2; }  // TMEM accumulators: NBUF offsets x 2 passes
__host__ __device__ constexpr int tc_tmem_cols(int nt) {
    return tc_nbuf(nt) * 2 * tc_col_stride(nt) < 32 ? 32 : tc_nbuf(nt) * 2 * tc_col_stride(nt);
}
__host__ __device__ constexpr int tc_ldo(int nt) { return nt + 4; }  // accumulator tile row stride (floats)

// packed[slice][k][chunk][NT rows x 128 bytes, SWIZZLE_128B]: B operand (N x K, K-major) of one pipeline step.
// element (n, c) of a chunk = W'[ci = 32 chunk + c][co = slice NT + n], rounded to TF32 (zero for ci >= c_in).
__global__ void __launch_bounds__(256) k_pack_weights_tc(const float *__restrict__ w, int volume, int n_in0, int n_out0,
                                                         int transpose, int flip, int nt, float *__restrict__ packed,
                                                         int64_t total) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c_in = transpose ? n_out0 : n_in0;
    const int nchunk = (c_in + kTcChunk - 1) / kTcChunk;
    const int64_t per_chunk = (int64_t)nt * kTcChunk, per_k = per_chunk * nchunk, per_slice = per_k * volume;
    const int slice = (int)(idx / per_slice);
    int64_t r = idx - (int64_t)slice * per_slice;
    const int k = (int)(r / per_k);
    r -= (int64_t)k * per_k;
    const int chunk = (int)(r / per_chunk);
    r -= (int64_t)chunk * per_chunk;
    // r = float index inside the swizzled tile: row n = r / 32, physical 16-byte piece pp = (r % 32) / 4
    const int n = (int)(r >> 5), pp = (int)(r & 31) >> 2, e = (int)(r & 3);
    const int c = ((pp ^ (n & 7)) << 2) + e;  // logical channel inside the chunk
    const int ci = chunk * kTcChunk + c, co = slice * nt + n;
    const int ks = flip ? volume - 1 - k : k;
    float v = 0.f;
    if (ci < c_in) v = transpose ? w[((int64_t)ks * n_in0 + co) * n_out0 + ci] : w[((int64_t)ks * n_in0 + ci) * n_out0 + co];
    packed[idx] = __uint_as_float(to_tf32(v));
}

struct TcSmem {  // byte offsets inside the dynamic shared memory block (base aligned to 1024)
    int a, b, out, lists, cnt, bars, total;
};
// lr = depth of the list / count ring: the gather warps run at most SA + NBUF offsets ahead of the epilogue warps
__host__ __device__ inline int tc_list_ring(int nt, int sa) { return sa + tc_nbuf(nt) + 1 <= 8 ? 8 : 16; }
__host__ __device__ inline TcSmem tc_smem_layout(int nt, int sa, int sb) {
    const int lr = tc_list_ring(nt, sa);
    TcSmem L;
    L.a = 0;
    L.b = L.a + sa * kTcAStage;
    L.out = L.b + sb * nt * 128;
    L.lists = L.out + kTcTM * tc_ldo(nt) * 4;
    L.cnt = L.lists + 4 * lr * kTcListBytes;
    L.bars = L.cnt + lr * 8 * 4;  // per ring slot: 4 counts + pass count (+ pad)
    L.total = L.bars + 8 * (2 * 8 + 2 * 4 + 2 * 4) + 16;     // mbarriers: A full/empty [8], B [4], D [4]; tmem pointer
    return L;
}

template <int NT>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_conv_tc(Gather gt, const float *__restrict__ in, int64_t ld_in, float *__restrict__ out, int64_t ld_out,
              const float *__restrict__ packed, int c_in, int SA, int SB) {
    constexpr int LDO = tc_ldo(NT);
    constexpr int NBUF = tc_nbuf(NT);
    constexpr int CS = tc_col_stride(NT);
    constexpr uint32_t IDESC = umma_idesc_tf32(NT);
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const TcSmem L = tc_smem_layout(NT, SA, SB);
    unsigned char *sA = smem + L.a, *sB = smem + L.b;
    float *sOut = reinterpret_cast<float *>(smem + L.out);
    const int LR = tc_list_ring(NT, SA);
    unsigned char *sLists = smem + L.lists;                      // [warp][LR][kTcListBytes]
    int32_t *sCnt = reinterpret_cast<int32_t *>(smem + L.cnt);  // [LR][8]: n_w (4), passes, pad
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + L.bars), *a_empty = a_full + 8;
    uint64_t *b_full = a_empty + 8, *b_empty = b_full + 4;
    uint64_t *d_full = b_empty + 4, *d_empty = d_full + 4;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(d_empty + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = gt.volume;
    const int nchunk = (c_in + kTcChunk - 1) / kTcChunk;
    packed += (int64_t)blockIdx.y * K * nchunk * NT * kTcChunk;
    out += (int64_t)blockIdx.y * NT;

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); mbar_init(d_full + i, 1); mbar_init(d_empty + i, 4); }
        mbar_fence_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr, tc_tmem_cols(NT));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
        // ================================================================= gather warps
        const int w = warp;
        const int64_t wrow0 = (int64_t)blockIdx.x * kTcTM + (int64_t)w * kTcRW;
        unsigned char *myLists = sLists + (size_t)w * LR * kTcListBytes;
        int nv[2], nn[2];
        auto lookup = [&](int k, int (&v)[2]) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int64_t row = wrow0 + 32 * j + lane;
                v[j] = (k < K && row < gt.n_out) ? gather_lookup(gt, k, row) : -1;
            }
        };
        lookup(0, nv);
        int s = 0;          // A-ring step counter (k, chunk, pass)
        int arrived = 0;    // steps whose copies have landed and been signalled
        const int DA = SA - 1;  // copies of up to DA later steps may still be in flight when a step is signalled
        auto signal_upto = [&](int upto) {  // all lanes: their copies of steps < upto are complete
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0)
                for (int q = arrived; q < upto; ++q) mbar_arrive(a_full + (q % SA));
            arrived = upto;
        };
        for (int k = 0; k < K; ++k) {
            lookup(k + 1, nn);
            // ---- ordered compaction: slot = rank of the row among this warp's rows that have a rule at offset k
            int32_t *lin = reinterpret_cast<int32_t *>(myLists + (k % LR) * kTcListBytes);
            unsigned char *lrow = reinterpret_cast<unsigned char *>(lin + kTcRW);
            int n = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const unsigned m = __ballot_sync(0xffffffffu, nv[j] >= 0);
                if (nv[j] >= 0) {
                    const int pos = n + __popc(m & ((1u << lane) - 1));
                    lin[pos] = nv[j];
                    lrow[pos] = (unsigned char)(w * kTcRW + 32 * j + lane);
                }
                n += __popc(m);
            }
            nv[0] = nn[0]; nv[1] = nn[1];
            int32_t *cnt = sCnt + (k % LR) * 8;
            if (lane == 0) cnt[w] = n;
            named_barrier_sync(1, 128);  // the four gather warps agree on the number of passes of this offset
            const int nmax = max(max(cnt[0], cnt[1]), max(cnt[2], cnt[3]));
            const int P = nmax > 32 ? 2 : 1;
            if (w == 0 && lane == 0) cnt[4] = P;
            __syncwarp();
            for (int ch = 0; ch < nchunk; ++ch) {
                const int kc16 = min(kTcChunk, c_in - ch * kTcChunk) / 4;  // valid 16-byte pieces per row
                for (int p = 0; p < P; ++p, ++s) {
                    const int st = s % SA;
                    mbar_wait(a_empty + st, ((s / SA) & 1) ^ 1);
                    const int np = min(32, n - 32 * p);
                    unsigned char *tile = sA + (size_t)st * kTcAStage;
                    const int c = lane & 7;
                    for (int j = lane >> 3; j < np; j += 4) {
                        const int idx = lin[32 * p + j];
                        const int r = 32 * w + j;  // row of the A tile
                        if (c < kc16)
                            cp_async16(tile + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4),
                                       in + (int64_t)idx * ld_in + ch * kTcChunk + 4 * c, true);
                    }
                    cp_async_commit();
                    if (s + 1 - arrived > DA) {  // keep at most DA steps in flight: wait for the oldest, signal it
                        switch (DA) {
                            case 1: cp_async_wait_upto<1>(); break;
                            case 2: cp_async_wait_upto<2>(); break;
                            case 3: cp_async_wait_upto<3>(); break;
                            case 4: cp_async_wait_upto<4>(); break;
                            case 5: cp_async_wait_upto<5>(); break;
                            case 6: cp_async_wait_upto<6>(); break;
                            default: cp_async_wait_upto<0>(); break;
                        }
                        signal_upto(DA <= 6 ? s + 1 - DA : s + 1);
                    }
                }
            }
        }
        cp_async_wait_upto<0>();
        signal_upto(s);
    } else if (warp < 8) {
        // ================================================================= epilogue warps
        const int w = warp - 4;
        const int64_t wrow0 = (int64_t)blockIdx.x * kTcTM + (int64_t)w * kTcRW;
        float *myOut = sOut + w * kTcRW * LDO;
        for (int i = lane; i < kTcRW * LDO / 4; i += 32) reinterpret_cast<float4 *>(myOut)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const unsigned char *myLists = sLists + (size_t)w * LR * kTcListBytes;
        for (int k = 0; k < K; ++k) {
            const int b = k % NBUF;
            mbar_wait(d_full + b, (k / NBUF) & 1);
            tc_fence_after_sync();
            const int32_t *cnt = sCnt + (k % LR) * 8;
            const int n = cnt[w], P = cnt[4];
            const unsigned char *lrow = myLists + (k % LR) * kTcListBytes + kTcRW * 4;
            for (int p = 0; p < P; ++p) {
                const int np = min(32, n - 32 * p);
                if (np <= 0) break;
                float *orow = lane < np ? sOut + lrow[32 * p + lane] * LDO : nullptr;
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * w) << 16) + (uint32_t)((b * 2 + p) * CS);
#pragma unroll
                for (int q = 0; q < NT / 16; ++q) {
                    float v[16];
                    tmem_ld16(taddr + 16 * q, v);
                    if (orow) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float4 o = *reinterpret_cast<float4 *>(orow + 16 * q + 4 * e);
                            o.x += v[4 * e]; o.y += v[4 * e + 1]; o.z += v[4 * e + 2]; o.w += v[4 * e + 3];
                            *reinterpret_cast<float4 *>(orow + 16 * q + 4 * e) = o;
                        }
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty + b);
        }
        __syncwarp();
        // ---- every output row is written once, coalesced
        constexpr int F4 = NT / 4;
        for (int i = lane; i < kTcRW * F4; i += 32) {
            const int r = i / F4, c = i - r * F4;
            const int64_t row = wrow0 + r;
            if (row >= gt.n_out) break;
            float4 v = *reinterpret_cast<const float4 *>(myOut + r * LDO + 4 * c);
            float4 *dst = reinterpret_cast<float4 *>(out + row * ld_out + 4 * c);
            if (gt.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *dst = v;
        }
    } else if (warp == 8) {
        // ================================================================= weight producer (TMA bulk copies)
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)NT * 128;
            const int n_steps = K * nchunk;
            for (int sb = 0; sb < n_steps; ++sb) {
                const int st = sb % SB;
                mbar_wait(b_empty + st, ((sb / SB) & 1) ^ 1);
                mbar_expect_tx(b_full + st, bytes);
                tma_load_1d(sB + (size_t)st * bytes, packed + (int64_t)sb * NT * kTcChunk, bytes, b_full + st);
            }
        }
    } else {
        // ================================================================= MMA issuer
        if (lane == 0) {
            const int LRi = LR;
            (void)LRi;
            int s = 0, sb = 0;
            for (int k = 0; k < K; ++k) {
                const int b = k % NBUF;
                mbar_wait(d_empty + b, ((k / NBUF) & 1) ^ 1);
                int P = 1;
                for (int ch = 0; ch < nchunk; ++ch, ++sb) {
                    const int stb = sb % SB;
                    mbar_wait(b_full + stb, (sb / SB) & 1);
                    const int nk = min(kTcChunk, c_in - ch * kTcChunk) / 8;  // MMAs (K = 8 each) in this chunk
                    const uint32_t b_addr = smem_u32(sB + (size_t)stb * NT * 128);
                    for (int p = 0; p < P; ++p, ++s) {
                        const int st = s % SA;
                        mbar_wait(a_full + st, (s / SA) & 1);
                        if (ch == 0 && p == 0) P = *(volatile int32_t *)(sCnt + (k % LR) * 8 + 4);
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(sA + (size_t)st * kTcAStage);
                        const uint32_t d = tmem_base + (uint32_t)((b * 2 + p) * CS);
                        for (int j = 0; j < nk; ++j)
                            umma_tf32(d, umma_desc_sw128(a_addr + 32 * j), umma_desc_sw128(b_addr + 32 * j), IDESC,
                                      (ch > 0 || j > 0) ? 1u : 0u);
                        umma_commit(a_empty + st);
                    }
                    umma_commit(b_empty + stb);
                }
                umma_commit(d_full + b);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, tc_tmem_cols(NT));
}

// ------------------------------------------------------------------------------------------------ host side
int tc_col_splits(int c_out) {  // slices of at most 112 output channels; 0 = not on the tcgen05 path
    if (c_out % 16 || c_out < 16) return 0;
    const int np = c_out / 16;
    if (np <= 7) return 1;
    if (np % 2 == 0 && np / 2 <= 7) return 2;
    return 0;
}
bool conv_tc_supported(int c_in, int c_out) { return c_in % 16 == 0 && c_in >= 16 && tc_col_splits(c_out) > 0; }

int64_t tc_packed_floats(int volume, int c_in, int c_out) {
    return (int64_t)volume * ceil_div(c_in, kTcChunk) * kTcChunk * c_out;
}

int pack_weights_tc(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, float *packed,
                    cudaStream_t s) {
    const int c_in = transpose ? n_out : n_in, c_out = transpose ? n_in : n_out;
    const int cs = tc_col_splits(c_out);
    MOPA_CHECK(cs > 0 && c_in % 16 == 0, "packWeights: shape is not on the tcgen05 path");
    const int64_t total = tc_packed_floats(volume, c_in, c_out);
    k_pack_weights_tc<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(weight, volume, n_in, n_out, transpose, flip, c_out / cs,
                                                                    packed, total);
    MOPA_LAUNCHED();
    return 0;
}

template <int NT>
static int launch_conv_tc(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out,
                          const float *packed, int c_in, int col_splits, cudaStream_t s) {
    // ring depths: as many A stages as fit beside the accumulator tile (two CTAs per SM while the tile is small)
    const int sb = NT <= 64 ? 3 : 2;
    const size_t cap = NT <= 32 ? (size_t)113 * 1024 : (size_t)226 * 1024;
    int sa = 8;
    while (sa > 2 && (size_t)tc_smem_layout(NT, sa, sb).total + 1024 > cap) --sa;
    if (sa > 7) sa = 7;
    const size_t smem = (size_t)tc_smem_layout(NT, sa, sb).total + 1024;
    auto kern = k_conv_tc<NT>;
    static bool configured = false;
    if (!configured) {
        MOPA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(gt.n_out, kTcTM), (unsigned)col_splits);
    kern<<<grid, kTcThreads, smem, s>>>(gt, in, ld_in, out, ld_out, packed, c_in, sa, sb);
    MOPA_LAUNCHED();
    return 0;
}

int conv_apply_tc(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out, const float *packed,
                  int c_in, int c_out, cudaStream_t s) {
    const int cs = tc_col_splits(c_out);
    switch (c_out / cs) {
        case 16: return launch_conv_tc<16>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
        case 32: return launch_conv_tc<32>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
        case 48: return launch_conv_tc<48>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
        case 64: return launch_conv_tc<64>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
        case 80: return launch_conv_tc<80>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
        case 96: return launch_conv_tc<96>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
        case 112: return launch_conv_tc<112>(gt, in, ld_in, out, ld_out, packed, c_in, cs, s);
    }
    MOPA_FAIL("unreachable tcgen05 conv shape");
}

}  // namespace mopa
