// Coordinate hashing, voxelisation and rulebook construction on the GPU.
//
// Replaces, for the path mopa/models/scn_unet.py:25-30 reaches, the host-side code of [UPSTREAM] SparseConvNet
// SCN/Metadata/{Metadata.cpp, IOLayersRules.h, SubmanifoldConvolutionRules.h, ConvolutionRules.h}. Upstream builds
// google::dense_hash_map grids and std::vector rulebooks on the CPU every forward; here every grid is an
// open-addressing table in HBM and every "rulebook" is a dense (offset, out-row) -> in-row table that the conv
// kernels read coalesced. All of it is integer work; results are bit-exact against oracle/scn_oracle.py.
//
// Voxel numbering: level-0 ids = order of first occurrence among the input rows; coarse ids =
// order of first occurrence when scanning fine ids ascending. Both come from one routine, unique_first():
//   insert all keys (atomicCAS) and atomicMin the element index into the slot  -> slot holds the first index
//   flag elements that ARE their slot's first index, exclusive-scan the flags  -> rank = id
//   write id back into the slot, then every element reads its id from its slot.
#include <mutex>

#include "geometry.cuh"
#include "mopa_scn.h"

namespace mopa {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

// ------------------------------------------------------------------------------------------------ profiling
struct ProfRec {
    int tag, volume, c_in, c_out;
    int64_t rows_out, rows_in, rules;
    int count_slot;  // index into g_prof_counts when the rule count is computed on the device, else -1
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static unsigned long long *g_prof_counts = nullptr;  // device, kProfSlots entries
constexpr int kProfSlots = 1 << 14;

__global__ void __launch_bounds__(256) k_count_valid(const int32_t *__restrict__ table, int64_t ld, int64_t V, int K,
                                                     unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V * K; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / V, o = i - k * V;
        c += table[k * ld + o] >= 0;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int prof_begin(int tag, const Gather *gt, int c_in, int c_out, int64_t rows, cudaStream_t s) {
    if (!g_prof_on) return -1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on || (int)g_prof.size() >= kProfSlots) return -1;
    ProfRec r{};
    r.tag = tag; r.c_in = c_in; r.c_out = c_out; r.rows_out = rows; r.rows_in = rows; r.rules = rows; r.count_slot = -1;
    if (gt) {
        r.volume = gt->volume; r.rows_out = gt->n_out; r.rows_in = gt->n_in;
        r.rules = gt->table && gt->volume == 8 ? gt->n_in : gt->n_out;  // strided: one rule per fine site
        if (gt->table && gt->volume == 27 && gt->n_out > 0 && g_prof_counts) {  // submanifold: count the table's rules
            r.count_slot = (int)g_prof.size();
            k_count_valid<<<num_sms() * 4, 256, 0, s>>>(gt->table, gt->ld, gt->n_out, 27, g_prof_counts + r.count_slot);
        }
    }
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
    cudaEventRecord(r.e0, s);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void prof_end(int idx, cudaStream_t s) {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].e1, s);
}

// ------------------------------------------------------------------------------------------------ allocation
int meta_alloc(mopa_scn_metadata *m, void **p, size_t bytes, cudaStream_t s) {
    if (bytes == 0) bytes = 16;
    MOPA_CUDA(cudaMallocAsync(p, bytes, s));
    m->allocs.push_back(*p);
    m->alloc_stream = s;
    return 0;
}

static int tmp_alloc(void **p, size_t bytes, cudaStream_t s) {
    MOPA_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, s));
    return 0;
}

// pinned 256-byte staging blocks, recycled process-wide (cudaHostAlloc costs tens of microseconds)
static std::mutex g_pin_mu;
static std::vector<int32_t *> g_pin_free;
static int32_t *pin_get() {
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) {
            int32_t *p = g_pin_free.back();
            g_pin_free.pop_back();
            return p;
        }
    }
    int32_t *p = nullptr;
    if (cudaHostAlloc((void **)&p, 256, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
static void pin_put(int32_t *p) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back(p);
}

// ------------------------------------------------------------------------------------------------ scan
__device__ __forceinline__ int block_exclusive_scan_1024(int v, int *total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();  // protect warp_sums from the previous call
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += t;
        }
        warp_sums[lane] = si - s;
        if (lane == 31 && total) *total = si;
    }
    __syncthreads();
    return incl - v + warp_sums[w];
}

__global__ void __launch_bounds__(1024) k_scan_blocks(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                      int64_t n, int32_t *__restrict__ bsum) {
    __shared__ int tot;
    int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int v = i < n ? in[i] : 0;
    int e = block_exclusive_scan_1024(v, &tot);
    if (i < n) out[i] = e;
    __syncthreads();
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single block: exclusive scan of bsum[0..nb) in place, grand total to *total
__global__ void __launch_bounds__(1024) k_scan_sums(int32_t *__restrict__ bsum, int64_t nb, int32_t *__restrict__ total) {
    __shared__ int tot;
    int carry = 0;
    for (int64_t base = 0; base < nb; base += 1024) {
        int64_t i = base + threadIdx.x;
        int v = i < nb ? bsum[i] : 0;
        int e = block_exclusive_scan_1024(v, &tot);
        if (i < nb) bsum[i] = e + carry;
        __syncthreads();
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(1024) k_scan_add(int32_t *__restrict__ out, int64_t n, const int32_t *__restrict__ bsum) {
    int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += bsum[blockIdx.x];
}

// exclusive scan of int32; `bsum` scratch of ceil(n/1024) ints; total (device) may be null
int exclusive_scan(const int32_t *in, int32_t *out, int64_t n, int32_t *bsum, int32_t *total, cudaStream_t s) {
    if (n <= 0) {
        if (total) MOPA_CUDA(cudaMemsetAsync(total, 0, 4, s));
        return 0;
    }
    int64_t nb = ceil_div(n, 1024);
    k_scan_blocks<<<(unsigned)nb, 1024, 0, s>>>(in, out, n, bsum);
    MOPA_LAUNCHED();
    k_scan_sums<<<1, 1024, 0, s>>>(bsum, nb, total);
    MOPA_LAUNCHED();
    if (nb > 1) {
        k_scan_add<<<(unsigned)nb, 1024, 0, s>>>(out, n, bsum);
        MOPA_LAUNCHED();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ unique_first
// MODE 0: keys from an int64 (n, ncols) coordinate matrix (validated); MODE 1: stride-2 parents of fine keys.
template <int MODE>
__global__ void __launch_bounds__(256) k_insert(const int64_t *__restrict__ coords, int ncols, int64_t spatial,
                                                const uint64_t *__restrict__ fine_keys, int64_t n,
                                                uint64_t *__restrict__ tab_keys, int32_t *__restrict__ tab_vals,
                                                uint32_t mask, int32_t *__restrict__ slot_of,
                                                uint64_t *__restrict__ key_of, int32_t *__restrict__ kidx,
                                                int32_t *__restrict__ err, const int32_t *__restrict__ n_dev) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;  // n: launch bound; *n_dev: the element count only the device knows
    uint64_t key;
    if (MODE == 0) {
        int64_t x, y, z, b = 0;
        if (ncols == 4) {
            const longlong2 *row = reinterpret_cast<const longlong2 *>(coords + 4 * i);
            longlong2 a = __ldg(row), c = __ldg(row + 1);
            x = a.x; y = a.y; z = c.x; b = c.y;
        } else {
            x = coords[3 * i]; y = coords[3 * i + 1]; z = coords[3 * i + 2];
        }
        if (x < 0 || y < 0 || z < 0 || x >= spatial || y >= spatial || z >= spatial || b < 0 || b >= 65535) {
            atomicExch(err, 1);
            slot_of[i] = -1;
            key_of[i] = kEmptyKey;
            return;
        }
        key = pack_key((uint32_t)x, (uint32_t)y, (uint32_t)z, (uint32_t)b);
    } else {
        uint64_t fk = fine_keys[i];
        key = parent_key(fk);
        // filter position of the fine site inside its 2x2x2 parent: (x&1)*4 + (y&1)*2 + (z&1)
        kidx[i] = (int)(((fk >> 32) & 1) * 4 + ((fk >> 16) & 1) * 2 + (fk & 1));
    }
    uint32_t s = hash_key(key) & mask;
    while (true) {
        uint64_t cur = tab_keys[s];
        if (cur == key) break;
        if (cur == kEmptyKey) {
            unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(tab_keys + s),
                                                (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (prev == kEmptyKey || prev == key) break;
        }
        s = (s + 1) & mask;
    }
    atomicMin(tab_vals + s, (int32_t)i);
    slot_of[i] = (int32_t)s;
    key_of[i] = key;
}

// k_flag_first + exclusive scan + k_assign_ids in ONE pass (decoupled look-back over 1024-element blocks): element i is a
// first occurrence iff its slot's first-index entry equals i; its id is the number of first occurrences before it; the id
// goes into tab_ids[slot] (a separate array: tab_first is still being read by other blocks) and its key into uniq_keys[id].
// status[b]: bits 63..62 = 1 (block aggregate published) / 2 (inclusive prefix published), low bits = the value; blocks take
// their index from an atomic ticket, so a block only ever waits for blocks that have already started.
__global__ void __launch_bounds__(1024) k_unique_scan(const int32_t *__restrict__ slot_of, const int32_t *__restrict__ tab_first,
                                                      const uint64_t *__restrict__ key_of, int64_t n,
                                                      const int32_t *__restrict__ n_dev, int32_t *__restrict__ tab_ids,
                                                      uint64_t *__restrict__ uniq_keys, int32_t *__restrict__ count_out,
                                                      unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket) {
    __shared__ int s_bid, s_prefix, s_total;
    if (threadIdx.x == 0) s_bid = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int bid = s_bid;
    const int64_t n_eff = n_dev ? min(n, (int64_t)*n_dev) : n;
    // launches are sized by an upper bound of n: blocks past the true count leave at once (nobody waits for them: a block
    // only looks back at lower indices, and every block below a working block is a working block)
    if (bid > 0 && (int64_t)bid * 1024 >= n_eff) return;
    const int64_t i = (int64_t)bid * 1024 + threadIdx.x;
    int slot = -1, flag = 0;
    if (i < n_eff) {
        slot = slot_of[i];
        flag = (slot >= 0 && tab_first[slot] == (int32_t)i) ? 1 : 0;
    }
    const int e = block_exclusive_scan_1024(flag, &s_total);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long total = (unsigned long long)s_total;
        unsigned long long prefix = 0;
        if (bid == 0) {
            __threadfence();
            atomicExch(status, (2ull << 62) | total);
        } else {
            __threadfence();
            atomicExch(status + bid, (1ull << 62) | total);
            for (int j = bid - 1; j >= 0; --j) {
                unsigned long long v;
                do { v = *(volatile unsigned long long *)(status + j); } while ((v >> 62) == 0);
                prefix += v & ((1ull << 62) - 1);
                if ((v >> 62) == 2) break;
            }
            __threadfence();
            atomicExch(status + bid, (2ull << 62) | (prefix + total));
        }
        s_prefix = (int)prefix;
        if ((int64_t)(bid + 1) * 1024 >= n_eff) *count_out = (int)(prefix + total);  // the block that holds the last element
    }
    __syncthreads();
    if (flag) {
        const int id = s_prefix + e;
        tab_ids[slot] = id;
        uniq_keys[id] = key_of[i];
    }
}

__global__ void __launch_bounds__(256) k_read_ids(const int32_t *__restrict__ slot_of,
                                                  const int32_t *__restrict__ tab_vals, int64_t n,
                                                  int32_t *__restrict__ ids, int32_t *__restrict__ counts,
                                                  const int32_t *__restrict__ n_dev) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    int s = slot_of[i];
    int id = s >= 0 ? tab_vals[s] : -1;
    ids[i] = id;
    if (counts && id >= 0) atomicAdd(counts + id, 1);
}

static uint32_t table_capacity(int64_t n) {
    uint64_t cap = 1024;
    while (cap < (uint64_t)(2 * n + 2)) cap <<= 1;
    return (uint32_t)cap;
}

// Builds the table + ids; *count_dev receives the number of unique keys. uniq_keys must hold n entries.
// n: number of elements, or (n_dev != nullptr) an upper bound of it with the true count in *n_dev on the device.
// Three kernels: hash insert (atomicCAS key, atomicMin first index), fused flag / scan / id assignment, id read-back.
static int unique_first(int mode, const int64_t *coords, int ncols, int64_t spatial, const uint64_t *fine_keys,
                        int64_t n, uint64_t *tab_keys, int32_t *tab_vals, uint32_t cap, uint64_t *uniq_keys,
                        int32_t *ids, int32_t *kidx, int32_t *counts, int32_t *count_dev, int32_t *err_dev,
                        cudaStream_t s, const int32_t *n_dev = nullptr) {
    MOPA_CUDA(cudaMemsetAsync(tab_keys, 0xFF, (size_t)cap * 8, s));
    if (n == 0) {
        MOPA_CUDA(cudaMemsetAsync(count_dev, 0, 4, s));
        return 0;
    }
    int32_t *slot_of, *tab_first;
    uint64_t *key_of;
    unsigned long long *status;
    const int64_t nb = ceil_div(n, 1024);
    MOPA_TRY(tmp_alloc((void **)&slot_of, n * 4, s));
    MOPA_TRY(tmp_alloc((void **)&tab_first, (size_t)cap * 4, s));
    MOPA_TRY(tmp_alloc((void **)&key_of, n * 8, s));
    MOPA_TRY(tmp_alloc((void **)&status, (size_t)(nb + 1) * 8, s));
    MOPA_CUDA(cudaMemsetAsync(tab_first, 0x7F, (size_t)cap * 4, s));
    MOPA_CUDA(cudaMemsetAsync(status, 0, (size_t)(nb + 1) * 8, s));  // last entry: the block ticket
    unsigned g = (unsigned)ceil_div(n, 256);
    if (mode == 0)
        k_insert<0><<<g, 256, 0, s>>>(coords, ncols, spatial, nullptr, n, tab_keys, tab_first, cap - 1, slot_of, key_of,
                                      nullptr, err_dev, n_dev);
    else
        k_insert<1><<<g, 256, 0, s>>>(nullptr, 0, 0, fine_keys, n, tab_keys, tab_first, cap - 1, slot_of, key_of, kidx,
                                      err_dev, n_dev);
    MOPA_LAUNCHED();
    k_unique_scan<<<(unsigned)nb, 1024, 0, s>>>(slot_of, tab_first, key_of, n, n_dev, tab_vals, uniq_keys, count_dev, status,
                                               reinterpret_cast<unsigned int *>(status + nb));
    MOPA_LAUNCHED();
    k_read_ids<<<g, 256, 0, s>>>(slot_of, tab_vals, n, ids, counts, n_dev);
    MOPA_LAUNCHED();
    MOPA_CUDA(cudaFreeAsync(slot_of, s));
    MOPA_CUDA(cudaFreeAsync(tab_first, s));
    MOPA_CUDA(cudaFreeAsync(key_of, s));
    MOPA_CUDA(cudaFreeAsync(status, s));
    return 0;
}

// ------------------------------------------------------------------------------------------------ input rules (CSR)
__global__ void __launch_bounds__(256) k_csr_fill(const int32_t *__restrict__ p2v, const int32_t *__restrict__ off,
                                                  int64_t n, int32_t *__restrict__ cursor, int32_t *__restrict__ tmp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = p2v[i];
    if (v < 0) return;
    int pos = atomicAdd(cursor + v, 1);
    tmp[off[v] + pos] = (int32_t)i;
}

// order each voxel's row list ascending: rank of row i = number of rows of the same voxel that are smaller
__global__ void __launch_bounds__(256) k_csr_order(const int32_t *__restrict__ p2v, const int32_t *__restrict__ off,
                                                   const int32_t *__restrict__ tmp, int64_t n,
                                                   int32_t *__restrict__ rows) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = p2v[i];
    if (v < 0) return;
    int beg = off[v], end = off[v + 1];
    int r = 0;
    for (int j = beg; j < end; ++j) r += (tmp[j] < (int32_t)i);
    rows[beg + r] = (int32_t)i;
}

// ------------------------------------------------------------------------------------------------ strided links, tile rulebooks
// Each kernel serves SEVERAL levels / links in one launch. After the single host round trip of the voxel pyramid every
// level's structures are independent of each other: one launch each for all child tables, all strided tile rulebooks and
// all submanifold levels instead of ~30 small ones (the small levels alone cannot fill the SMs, and a launch costs more
// than their work). Blocks are numbered through the jobs; first[j] = first block of job j.
//   k_tile_lists_batch: block = (tile of 128 output rows, offset k): thread r looks up the input row of (k, row 128 t + r)
//     in the dense table (or the select pair), the block compacts the hits in row order (ballot + popc ranks) and writes
//     the row mask: the warp-cooperative rulebook build the conv kernels consume (they never ballot / compact).
//   k_subm_tiles_batch: the same for a submanifold level, fused with the hash probes: thread r probes the grid for the
//     neighbour of site 128 t + r at offset k and also writes the dense table entry (fp32-mode / small-channel kernels).
constexpr int kGeoMaxJobs = 16;
struct GeoJobs {
    int n;
    int first[kGeoMaxJobs + 1];
    int K[kGeoMaxJobs];
    int64_t V[kGeoMaxJobs], ld[kGeoMaxJobs];
    const int32_t *table[kGeoMaxJobs], *parent[kGeoMaxJobs], *kidx[kGeoMaxJobs];
    int32_t *out[kGeoMaxJobs];  // child table (scatter) or tile lists
    uint4 *tm[kGeoMaxJobs];
    // submanifold jobs
    const uint64_t *keys[kGeoMaxJobs], *tab_keys[kGeoMaxJobs];
    const int32_t *tab_vals[kGeoMaxJobs];
    int32_t *nbr[kGeoMaxJobs];
    uint32_t mask[kGeoMaxJobs];
    int spatial[kGeoMaxJobs];
};
__device__ __forceinline__ int geo_job_of(const GeoJobs &J, int block) {
    int j = 0;
    while (j + 1 < J.n && block >= J.first[j + 1]) ++j;
    return j;
}
__global__ void __launch_bounds__(256) k_child_scatter_batch(const __grid_constant__ GeoJobs J) {
    const int j = geo_job_of(J, blockIdx.x);
    const int64_t c = (int64_t)(blockIdx.x - J.first[j]) * blockDim.x + threadIdx.x;
    if (c >= J.V[j]) return;
    J.out[j][(int64_t)J.kidx[j][c] * J.ld[j] + J.parent[j][c]] = (int32_t)c;
}
__global__ void __launch_bounds__(128) k_tile_lists_batch(const __grid_constant__ GeoJobs J) {
    __shared__ uint32_t wm[4];
    const int j = geo_job_of(J, blockIdx.x);
    const int K = J.K[j], local = blockIdx.x - J.first[j];
    const int r = threadIdx.x, lane = r & 31, warp = r >> 5, k = local % K;
    const int64_t t = local / K, o = t * kTileRows + r;
    const int32_t *table = J.table[j];
    int id = -1;
    if (o < J.V[j]) id = table ? __ldg(table + (int64_t)k * J.ld[j] + o) : (__ldg(J.kidx[j] + o) == k ? __ldg(J.parent[j] + o) : -1);
    const uint32_t m = __ballot_sync(0xffffffffu, id >= 0);
    if (lane == 0) wm[warp] = m;
    __syncthreads();
    int rank = __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) rank += __popc(wm[w]);
    const int64_t slot = t * K + k;
    if (id >= 0) J.out[j][(slot << 7) + rank] = id | (r << kTileRowShift);
    if (r == 0) J.tm[j][slot] = make_uint4(wm[0], wm[1], wm[2], wm[3]);
}
__global__ void __launch_bounds__(128) k_subm_tiles_batch(const __grid_constant__ GeoJobs J) {
    __shared__ uint32_t wm[4];
    const int j = geo_job_of(J, blockIdx.x);
    const int local = blockIdx.x - J.first[j];
    const int r = threadIdx.x, lane = r & 31, warp = r >> 5, k = local % 27;
    const int64_t t = local / 27, o = t * kTileRows + r;
    int id = -1;
    if (o < J.V[j]) {
        if (k == 13) {
            id = (int)o;
        } else {
            const int dx = k / 9 - 1, dy = (k / 3) % 3 - 1, dz = k % 3 - 1, spatial = J.spatial[j];
            int x, y, z, b;
            unpack_key(J.keys[j][o], x, y, z, b);
            x += dx; y += dy; z += dz;
            if (x >= 0 && y >= 0 && z >= 0 && x < spatial && y < spatial && z < spatial) {
                const uint64_t q = pack_key((uint32_t)x, (uint32_t)y, (uint32_t)z, (uint32_t)b);
                const uint64_t *tab_keys = J.tab_keys[j];
                const uint32_t mask = J.mask[j];
                uint32_t s = hash_key(q) & mask;
                while (true) {
                    const uint64_t cur = __ldg(tab_keys + s);
                    if (cur == q) { id = __ldg(J.tab_vals[j] + s); break; }
                    if (cur == kEmptyKey) break;
                    s = (s + 1) & mask;
                }
            }
        }
        J.nbr[j][(int64_t)k * J.ld[j] + o] = id;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, id >= 0);
    if (lane == 0) wm[warp] = m;
    __syncthreads();
    int rank = __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) rank += __popc(wm[w]);
    const int64_t slot = t * 27 + k;
    if (id >= 0) J.out[j][(slot << 7) + rank] = id | (r << kTileRowShift);
    if (r == 0) J.tm[j][slot] = make_uint4(wm[0], wm[1], wm[2], wm[3]);
}

// allocates the tile rulebook of one (rows, K) and appends the job; the launch follows in flush_tile_jobs
static int add_tile_job(mopa_scn_metadata *m, GeoJobs &J, const int32_t *table, int64_t ld, const int32_t *parent,
                        const int32_t *kidx, int64_t V, int64_t n_in, int K, int32_t **tl, uint4 **tm, cudaStream_t s) {
    MOPA_CHECK(n_in < ((int64_t)1 << kTileRowShift), "more than 2^25 active sites at one level: tile rulebook entries overflow");
    MOPA_CHECK(J.n < kGeoMaxJobs, "too many tile rulebook jobs in one batch");
    const int64_t tiles = ceil_div(V > 0 ? V : 1, kTileRows);
    MOPA_TRY(meta_alloc(m, (void **)tl, (size_t)tiles * K * kTileRows * 4, s));
    MOPA_TRY(meta_alloc(m, (void **)tm, (size_t)tiles * K * sizeof(uint4), s));
    const int j = J.n++;
    J.table[j] = table; J.ld[j] = ld; J.parent[j] = parent; J.kidx[j] = kidx; J.V[j] = V; J.K[j] = K;
    J.out[j] = *tl; J.tm[j] = *tm;
    J.first[j + 1] = J.first[j] + (int)(tiles * K);
    return 0;
}
static int flush_tile_jobs(GeoJobs &J, cudaStream_t s) {
    if (J.n == 0) return 0;
    k_tile_lists_batch<<<(unsigned)J.first[J.n], 128, 0, s>>>(J);
    MOPA_LAUNCHED();
    J.n = 0;
    return 0;
}

// ------------------------------------------------------------------------------------------------ rulebook compaction
// Warp-cooperative ordered compaction of a dense (K, ld) table into offset-major (in, out) pairs, ascending out row.
__global__ void __launch_bounds__(1024) k_rule_flags(const int32_t *__restrict__ table, int64_t ld, int64_t V, int K,
                                                     int32_t *__restrict__ flags) {
    int64_t o = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int k = blockIdx.y;
    if (o < V && k < K) flags[(int64_t)k * V + o] = table[(int64_t)k * ld + o] >= 0;
}
__global__ void __launch_bounds__(1024) k_rule_write(const int32_t *__restrict__ table, int64_t ld, int64_t V, int K,
                                                     const int32_t *__restrict__ rank, int32_t *__restrict__ pairs) {
    int64_t o = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int k = blockIdx.y;
    if (o >= V || k >= K) return;
    int in = table[(int64_t)k * ld + o];
    if (in < 0) return;
    int64_t r = rank[(int64_t)k * V + o];
    pairs[2 * r] = in;
    pairs[2 * r + 1] = (int32_t)o;
}
__global__ void k_rule_offsets(const int32_t *__restrict__ rank, const int32_t *__restrict__ total, int64_t V, int K,
                               int32_t *__restrict__ offs) {
    int k = threadIdx.x;
    if (k < K) offs[k] = rank[(int64_t)k * V];
    if (k == K) offs[K] = *total;
}

// ------------------------------------------------------------------------------------------------ Metadata ops
int ensure_subm_many(mopa_scn_metadata *m, const int *levels, int n, cudaStream_t s) {
    MOPA_TRY(finish_levels(m, s));
    GeoJobs J{};
    for (int i = 0; i < n; ++i) {
        Level &L = m->levels[levels[i]];
        if (L.nbr) continue;
        if (J.n == kGeoMaxJobs) {
            k_subm_tiles_batch<<<(unsigned)J.first[J.n], 128, 0, s>>>(J);
            MOPA_LAUNCHED();
            J = GeoJobs{};
        }
        L.nbr_ld = round_up(L.V > 0 ? L.V : 1, 32);
        MOPA_TRY(meta_alloc(m, (void **)&L.nbr, (size_t)27 * L.nbr_ld * 4, s));
        MOPA_CHECK(L.V < ((int64_t)1 << kTileRowShift), "more than 2^25 active sites at one level: tile rulebook entries overflow");
        const int64_t tiles = ceil_div(L.V > 0 ? L.V : 1, kTileRows);
        MOPA_TRY(meta_alloc(m, (void **)&L.tl_subm, (size_t)tiles * 27 * kTileRows * 4, s));
        MOPA_TRY(meta_alloc(m, (void **)&L.tm_subm, (size_t)tiles * 27 * sizeof(uint4), s));
        const int j = J.n++;
        J.keys[j] = L.keys; J.V[j] = L.V; J.spatial[j] = (int)L.spatial; J.tab_keys[j] = L.tab_keys; J.tab_vals[j] = L.tab_vals;
        J.mask[j] = (uint32_t)(L.cap - 1); J.nbr[j] = L.nbr; J.ld[j] = L.nbr_ld; J.out[j] = L.tl_subm; J.tm[j] = L.tm_subm;
        J.first[j + 1] = J.first[j] + (int)(tiles * 27);
    }
    if (J.n) {
        k_subm_tiles_batch<<<(unsigned)J.first[J.n], 128, 0, s>>>(J);
        MOPA_LAUNCHED();
    }
    return 0;
}
int ensure_subm(mopa_scn_metadata *m, int level, cudaStream_t s) { return ensure_subm_many(m, &level, 1, s); }

static int read_back(mopa_scn_metadata *m, const int32_t *dev, int n_ints, cudaStream_t s) {
    MOPA_CUDA(cudaMemcpyAsync(m->pinned, dev, (size_t)n_ints * 4, cudaMemcpyDeviceToHost, s));
    MOPA_CUDA(cudaStreamSynchronize(s));
    return 0;
}

static int cnt_block(mopa_scn_metadata *m, cudaStream_t s) {
    if (m->cnt_dev) return 0;
    MOPA_TRY(meta_alloc(m, (void **)&m->cnt_dev, 32 * 4, s));
    MOPA_CUDA(cudaMemsetAsync(m->cnt_dev, 0, 32 * 4, s));
    return 0;
}

// Hashes level + 1 (stride-2 parents of level's sites): table, coarse keys in id order, parent / kidx of every fine
// site. No host round trip: when the fine level's site count is still device-only, its upper bound sizes the
// allocations and the launches, and the kernels read the true count from cnt_dev. child / tile rulebooks (which need exact
// row counts) follow in finish_levels.
int ensure_down_async(mopa_scn_metadata *m, int level, cudaStream_t s) {
    if (m->levels[level].has_down) return 0;
    MOPA_CHECK((int)m->levels.size() == level + 1, "strided levels must be created in order");
    MOPA_CHECK(level + 1 < 31, "too many levels");
    MOPA_CHECK(m->levels[level].spatial % 2 == 0, "input spatial size must be even for a size-2 stride-2 convolution");
    MOPA_TRY(cnt_block(m, s));
    m->levels.emplace_back();
    Level &L = m->levels[level];
    Level &N = m->levels[level + 1];
    N.spatial = L.spatial / 2;
    const bool pending = L.V < 0;
    const int64_t Vf = pending ? L.V_bound : L.V;  // elements to launch / allocate for
    N.V = -1;
    N.V_bound = Vf;
    N.cap = table_capacity(Vf);
    MOPA_TRY(meta_alloc(m, (void **)&N.tab_keys, (size_t)N.cap * 8, s));
    MOPA_TRY(meta_alloc(m, (void **)&N.tab_vals, (size_t)N.cap * 4, s));
    MOPA_TRY(meta_alloc(m, (void **)&N.keys, (size_t)Vf * 8, s));
    MOPA_TRY(meta_alloc(m, (void **)&L.parent, (size_t)Vf * 4, s));
    MOPA_TRY(meta_alloc(m, (void **)&L.kidx, (size_t)Vf * 4, s));
    MOPA_TRY(unique_first(1, nullptr, 0, 0, L.keys, Vf, N.tab_keys, N.tab_vals, N.cap, N.keys, L.parent, L.kidx, nullptr,
                          m->cnt_dev + level + 1, m->cnt_dev + 31, s, pending ? m->cnt_dev + level : nullptr));
    if (m->pending_from < 0) m->pending_from = level + 1;
    L.has_down = true;
    return 0;
}

// ONE synchronisation for every level hashed since the last call: reads the site counts back, then builds the structures
// that need exact row counts (child tables, tile rulebooks of the strided links).
int finish_levels(mopa_scn_metadata *m, cudaStream_t s) {
    if (m->pending_from < 0) return 0;
    const int first = m->pending_from, n_lv = (int)m->levels.size();
    MOPA_CUDA(cudaMemcpyAsync(m->pinned, m->cnt_dev, 32 * 4, cudaMemcpyDeviceToHost, s));
    MOPA_CUDA(cudaStreamSynchronize(s));
    m->pending_from = -1;
    for (int l = first; l < n_lv; ++l)
        if (m->levels[l].V < 0) m->levels[l].V = m->pinned[l];
    MOPA_CHECK(m->pinned[31] == 0, "InputLayer: coordinates outside [0, spatial_size) or batch index outside [0, 65535)");
    // every link's child table in one block (one fill, one scatter launch), then every strided tile rulebook in one launch
    std::vector<int> todo;
    size_t child_ints = 0;
    for (int l = (first > 0 ? first - 1 : 0); l + 1 < n_lv; ++l) {
        Level &L = m->levels[l];
        if (!L.has_down || L.child) continue;
        L.child_ld = round_up(m->levels[l + 1].V > 0 ? m->levels[l + 1].V : 1, 32);
        child_ints += (size_t)8 * L.child_ld;
        todo.push_back(l);
    }
    if (todo.empty()) return 0;
    int32_t *block = nullptr;
    MOPA_TRY(meta_alloc(m, (void **)&block, child_ints * 4, s));
    MOPA_CUDA(cudaMemsetAsync(block, 0xFF, child_ints * 4, s));
    for (size_t b = 0; b < todo.size(); b += kGeoMaxJobs / 2) {  // two tile rulebooks per link
        GeoJobs S{}, T{};
        for (size_t i = b; i < todo.size() && i < b + kGeoMaxJobs / 2; ++i) {
            Level &L = m->levels[todo[i]];
            Level &N = m->levels[todo[i] + 1];
            L.child = block;
            block += (size_t)8 * L.child_ld;
            if (L.V > 0) {
                const int j = S.n++;
                S.parent[j] = L.parent; S.kidx[j] = L.kidx; S.V[j] = L.V; S.out[j] = L.child; S.ld[j] = L.child_ld;
                S.first[j + 1] = S.first[j] + (int)ceil_div(L.V, 256);
            }
            MOPA_TRY(add_tile_job(m, T, L.child, L.child_ld, nullptr, nullptr, N.V, L.V, 8, &L.tl_child, &L.tm_child, s));
            MOPA_TRY(add_tile_job(m, T, nullptr, 0, L.parent, L.kidx, L.V, N.V, 8, &L.tl_sel, &L.tm_sel, s));
        }
        if (S.n) {
            k_child_scatter_batch<<<(unsigned)S.first[S.n], 256, 0, s>>>(S);
            MOPA_LAUNCHED();
        }
        MOPA_TRY(flush_tile_jobs(T, s));
    }
    return 0;
}

int ensure_down(mopa_scn_metadata *m, int level, cudaStream_t s) {
    if (m->levels[level].has_down && m->pending_from < 0) return 0;
    MOPA_TRY(ensure_down_async(m, level, s));
    return finish_levels(m, s);
}

// compact a (K, ld) table with V columns into pairs on the HOST side buffers (inspection only; synchronises)
static int rulebook_to_host(mopa_scn_metadata *m, const int32_t *table, int64_t ld, int64_t V, int K,
                            int64_t *counts_host, int32_t *pairs_host, cudaStream_t s) {
    for (int k = 0; k < K; ++k) counts_host[k] = 0;
    if (V == 0) return 0;
    int64_t n = (int64_t)K * V;
    int32_t *flags, *rank, *bsum, *offs, *pairs = nullptr;
    MOPA_TRY(tmp_alloc((void **)&flags, n * 4, s));
    MOPA_TRY(tmp_alloc((void **)&rank, n * 4, s));
    MOPA_TRY(tmp_alloc((void **)&bsum, ceil_div(n, 1024) * 4, s));
    MOPA_TRY(tmp_alloc((void **)&offs, 64 * 4, s));
    dim3 grid((unsigned)ceil_div(V, 1024), K);
    k_rule_flags<<<grid, 1024, 0, s>>>(table, ld, V, K, flags);
    MOPA_LAUNCHED();
    MOPA_TRY(exclusive_scan(flags, rank, n, bsum, offs + 63, s));
    k_rule_offsets<<<1, 64, 0, s>>>(rank, offs + 63, V, K, offs);
    MOPA_LAUNCHED();
    MOPA_TRY(read_back(m, offs, K + 1, s));
    int64_t total = m->pinned[K];
    for (int k = 0; k < K; ++k) counts_host[k] = m->pinned[k + 1] - m->pinned[k];
    if (pairs_host && total > 0) {
        MOPA_TRY(tmp_alloc((void **)&pairs, total * 8, s));
        k_rule_write<<<grid, 1024, 0, s>>>(table, ld, V, K, rank, pairs);
        MOPA_LAUNCHED();
        MOPA_CUDA(cudaMemcpyAsync(pairs_host, pairs, total * 8, cudaMemcpyDeviceToHost, s));
        MOPA_CUDA(cudaStreamSynchronize(s));
        MOPA_CUDA(cudaFreeAsync(pairs, s));
    }
    MOPA_CUDA(cudaFreeAsync(flags, s));
    MOPA_CUDA(cudaFreeAsync(rank, s));
    MOPA_CUDA(cudaFreeAsync(bsum, s));
    MOPA_CUDA(cudaFreeAsync(offs, s));
    return 0;
}

int set_locations(mopa_scn_metadata *m, int64_t spatial_size, const int64_t *coords, int64_t n, int ncols,
                  int coords_on_device, cudaStream_t s, bool defer_sync) {
    MOPA_CHECK(m != nullptr, "null metadata");
    MOPA_CHECK(ncols == 3 || ncols == 4, "coords must have 3 or 4 columns");
    MOPA_CHECK(spatial_size > 0 && spatial_size <= 65536, "spatial_size must be in (0, 65536]");
    MOPA_CHECK(m->levels.empty(), "setLocations called twice on one Metadata");
    MOPA_CHECK(n >= 0 && n < (int64_t)1 << 30, "point count out of range");
    MOPA_CUDA(cudaSetDevice(m->device));
    m->levels.emplace_back();
    Level &L = m->levels[0];
    L.spatial = spatial_size;
    m->n_points = n;

    const int64_t *dcoords = coords;
    int64_t *staged = nullptr;
    if ((!coords_on_device || ((uintptr_t)coords & 15) != 0) && n > 0) {
        MOPA_TRY(tmp_alloc((void **)&staged, (size_t)n * ncols * 8, s));
        MOPA_CUDA(cudaMemcpyAsync(staged, coords, (size_t)n * ncols * 8,
                                  coords_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
        dcoords = staged;
    }
    L.cap = table_capacity(n);
    MOPA_TRY(meta_alloc(m, (void **)&L.tab_keys, (size_t)L.cap * 8, s));
    MOPA_TRY(meta_alloc(m, (void **)&L.tab_vals, (size_t)L.cap * 4, s));
    MOPA_TRY(meta_alloc(m, (void **)&L.keys, (size_t)n * 8, s));
    MOPA_TRY(meta_alloc(m, (void **)&m->p2v, (size_t)n * 4, s));
    MOPA_TRY(meta_alloc(m, (void **)&m->csr_off, (size_t)(n + 1) * 4, s));
    MOPA_TRY(meta_alloc(m, (void **)&m->csr_rows, (size_t)n * 4, s));
    int32_t *counts, *tmp_rows, *bsum;
    MOPA_TRY(cnt_block(m, s));
    MOPA_TRY(tmp_alloc((void **)&counts, (size_t)(n + 1) * 4, s));
    MOPA_TRY(tmp_alloc((void **)&tmp_rows, (size_t)n * 4, s));
    MOPA_TRY(tmp_alloc((void **)&bsum, (size_t)ceil_div(n + 1, 1024) * 4, s));
    MOPA_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n + 1) * 4, s));
    MOPA_TRY(unique_first(0, dcoords, ncols, spatial_size, nullptr, n, L.tab_keys, L.tab_vals, L.cap, L.keys, m->p2v,
                          nullptr, counts, m->cnt_dev, m->cnt_dev + 31, s));
    if (n > 0) {
        // CSR offsets over the (n + 1)-long zero-padded count array: off[v] valid for v <= V0, off[V0] = n
        MOPA_TRY(exclusive_scan(counts, m->csr_off, n + 1, bsum, nullptr, s));
        MOPA_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * 4, s));  // reuse as per-voxel cursor
        unsigned g = (unsigned)ceil_div(n, 256);
        k_csr_fill<<<g, 256, 0, s>>>(m->p2v, m->csr_off, n, counts, tmp_rows);
        MOPA_LAUNCHED();
        k_csr_order<<<g, 256, 0, s>>>(m->p2v, m->csr_off, tmp_rows, n, m->csr_rows);
        MOPA_LAUNCHED();
    } else {
        MOPA_CUDA(cudaMemsetAsync(m->csr_off, 0, 4, s));
    }
    MOPA_CUDA(cudaFreeAsync(counts, s));
    MOPA_CUDA(cudaFreeAsync(tmp_rows, s));
    MOPA_CUDA(cudaFreeAsync(bsum, s));
    if (staged) MOPA_CUDA(cudaFreeAsync(staged, s));
    L.V = -1;  // device-only until finish_levels
    L.V_bound = n;
    m->pending_from = 0;
    if (!defer_sync) MOPA_TRY(finish_levels(m, s));
    return 0;
}


}  // namespace mopa

using namespace mopa;

// ================================================================================================ C ABI
extern "C" {

int mopa_scn_abi_version(void) { return MOPA_SCN_ABI_VERSION; }
const char *mopa_scn_last_error(void) { return g_last_error.c_str(); }
int64_t mopa_scn_kernelLaunchCount(void) { return (int64_t)g_launches.load(); }

int mopa_scn_Profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfRec &r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    if (on) {
        if (!g_prof_counts) MOPA_CUDA(cudaMalloc((void **)&g_prof_counts, (size_t)kProfSlots * 8));
        MOPA_CUDA(cudaMemset(g_prof_counts, 0, (size_t)kProfSlots * 8));
    }
    g_prof_on = on != 0;
    return 0;
}
int64_t mopa_scn_Profile_count(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    return (int64_t)g_prof.size();
}
int mopa_scn_Profile_read(mopa_scn_profile_record *out, int64_t max_records) {
    MOPA_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::vector<unsigned long long> counts(g_prof.size());
    if (!g_prof.empty() && g_prof_counts)
        MOPA_CUDA(cudaMemcpy(counts.data(), g_prof_counts, g_prof.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < g_prof.size() && (int64_t)i < max_records; ++i) {
        const ProfRec &r = g_prof[i];
        float ms = 0.f;
        MOPA_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        out[i].tag = r.tag; out[i].volume = r.volume; out[i].c_in = r.c_in; out[i].c_out = r.c_out;
        out[i].rows_out = r.rows_out; out[i].rows_in = r.rows_in;
        out[i].rules = r.count_slot >= 0 ? (int64_t)counts[r.count_slot] : r.rules;
        out[i].ms = ms; out[i].reserved = 0;
    }
    return 0;
}

struct Parked {
    int device;
    cudaEvent_t done;
    cudaStream_t stream;
    std::vector<void *> ptrs;
};
static std::mutex g_park_mu;
static std::vector<Parked> g_parked;
// Also the throttle of a loop that never synchronises: with the geometry on its own stream nothing else stops the host
// from queueing step after step (measured: the pool kept growing and single steps stalled for 100-350 ms). More than
// kMaxParked forwards whose memory is still in use -> wait for the oldest: the host stays about one step ahead.
constexpr int kMaxParked = 1;
static void reap_parked(int device) {  // current device == device
    std::lock_guard<std::mutex> lk(g_park_mu);
    int mine = 0;
    for (const Parked &e : g_parked) mine += e.device == device;
    for (size_t i = 0; i < g_parked.size();) {
        Parked &e = g_parked[i];
        if (e.device != device) { ++i; continue; }
        if (mine > kMaxParked) cudaEventSynchronize(e.done);  // (entries are in order of deletion: this is the oldest)
        if (cudaEventQuery(e.done) != cudaSuccess) { ++i; continue; }
        for (void *p : e.ptrs) cudaFreeAsync(p, e.stream);
        cudaEventDestroy(e.done);
        g_parked.erase(g_parked.begin() + (long)i);
        --mine;
    }
    cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
}

mopa_scn_metadata *mopa_scn_Metadata_new(int dimension, int device) {
    if (dimension != 3) {
        fail(__FILE__, __LINE__, "only dimension 3 is supported (scn_unet.py: DIMENSION = 3)");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(__FILE__, __LINE__, "cudaSetDevice failed: no usable CUDA device (this library has no CPU path)");
        return nullptr;
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;  // keep freed blocks cached: per-forward scratch is re-used, not re-mapped
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    reap_parked(device);
    auto *m = new mopa_scn_metadata();
    m->device = device;
    m->pinned = pin_get();
    if (!m->pinned) {
        fail(__FILE__, __LINE__, "cudaHostAlloc failed");
        delete m;
        return nullptr;
    }
    return m;
}

void mopa_scn_Metadata_delete(mopa_scn_metadata *m) {
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(m->device);
    // Blocks allocated on another stream than the one that used them last (compiled networks: the geometry stream) are
    // parked with an event of the last user and freed ON THEIR ALLOCATION STREAM once that event has completed (reap_parked,
    // called when the next Metadata is created): the pool then recycles them in plain stream order and nothing ever waits.
    // Freed on the compute stream instead, the geometry stream's next allocations (issued while the compute stream is a
    // step behind) found only blocks whose release was still pending and the pool grew by fresh device memory (rare
    // 30-90 ms stalls); made to wait for the compute stream, the geometry of step i+1 no longer overlapped step i.
    bool parked = false;
    if (!m->allocs.empty() && m->alloc_stream != m->last_stream) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
            if (cudaEventRecord(ev, m->last_stream) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_park_mu);
                g_parked.push_back(Parked{m->device, ev, m->alloc_stream, std::move(m->allocs)});
                parked = true;
            } else {
                cudaEventDestroy(ev);
            }
        }
    }
    if (!parked)
        for (void *p : m->allocs) cudaFreeAsync(p, m->last_stream);
    if (m->geom_done) cudaEventDestroy(m->geom_done);
    cudaSetDevice(prev);
    pin_put(m->pinned);
    delete m;
}

int mopa_scn_InputLayer_setLocations(mopa_scn_metadata *m, int64_t spatial_size, const int64_t *coords, int64_t n,
                                     int ncols, int coords_on_device, int mode, void *stream, int64_t *n_active_out) {
    MOPA_CHECK(m != nullptr, "null metadata");
    MOPA_CHECK(mode == 4, "only InputLayer mode 4 (mean) is implemented (scn_unet.py:26)");
    m->last_stream = (cudaStream_t)stream;
    MOPA_TRY(set_locations(m, spatial_size, coords, n, ncols, coords_on_device, (cudaStream_t)stream));
    if (n_active_out) *n_active_out = m->levels[0].V;
    return 0;
}

int mopa_scn_Metadata_prepareSubmanifold(mopa_scn_metadata *m, int64_t spatial_size, int filter_size, void *stream,
                                         int64_t *n_active_out) {
    MOPA_CHECK(m != nullptr, "null metadata");
    MOPA_CHECK(filter_size == 3, "only 3x3x3 submanifold filters are implemented");
    int l = m->level_of(spatial_size);
    MOPA_CHECK(l >= 0, "no grid at this spatial size (call InputLayer / Convolution first)");
    m->last_stream = (cudaStream_t)stream;
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_subm(m, l, (cudaStream_t)stream));
    if (n_active_out) *n_active_out = m->levels[l].V;
    return 0;
}

int mopa_scn_Metadata_prepareConvolution(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                         int filter_size, int filter_stride, void *stream, int64_t *n_active_out) {
    MOPA_CHECK(m != nullptr, "null metadata");
    MOPA_CHECK(filter_size == 2 && filter_stride == 2, "only size-2 stride-2 (de)convolutions are implemented");
    MOPA_CHECK((out_spatial_size - 1) * filter_stride + filter_size == in_spatial_size,
               "Convolution: (out - 1) * stride + size must equal the input spatial size");
    int l = m->level_of(in_spatial_size);
    MOPA_CHECK(l >= 0, "no grid at the input spatial size");
    m->last_stream = (cudaStream_t)stream;
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_down(m, l, (cudaStream_t)stream));
    if (n_active_out) *n_active_out = m->levels[l + 1].V;
    return 0;
}

int64_t mopa_scn_Metadata_getNActive(mopa_scn_metadata *m, int64_t spatial_size) {
    if (!m) return -1;
    int l = m->level_of(spatial_size);
    return l < 0 ? -1 : m->levels[l].V;
}
int64_t mopa_scn_Metadata_getNPoints(mopa_scn_metadata *m) { return m ? m->n_points : -1; }

int64_t mopa_scn_Metadata_getSubmanifoldRuleCount(mopa_scn_metadata *m, int64_t spatial_size) {
    // upper bound usable for workspace sizing without a device round trip: every site has at most 27 rules
    if (!m) return -1;
    int l = m->level_of(spatial_size);
    return l < 0 ? -1 : 27 * m->levels[l].V;
}

int mopa_scn_Metadata_getSpatialLocations(mopa_scn_metadata *m, int64_t spatial_size, int64_t *coords_host) {
    MOPA_CHECK(m != nullptr, "null metadata");
    int l = m->level_of(spatial_size);
    MOPA_CHECK(l >= 0, "no grid at this spatial size");
    const Level &L = m->levels[l];
    std::vector<uint64_t> keys((size_t)L.V);
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_CUDA(cudaStreamSynchronize(m->last_stream));
    if (L.V) MOPA_CUDA(cudaMemcpy(keys.data(), L.keys, (size_t)L.V * 8, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < L.V; ++i) {
        int x, y, z, b;
        unpack_key(keys[(size_t)i], x, y, z, b);
        coords_host[4 * i] = x; coords_host[4 * i + 1] = y; coords_host[4 * i + 2] = z; coords_host[4 * i + 3] = b;
    }
    return 0;
}

int mopa_scn_Metadata_getPointToVoxel(mopa_scn_metadata *m, int32_t *p2v_host) {
    MOPA_CHECK(m != nullptr && !m->levels.empty(), "metadata has no input layer");
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_CUDA(cudaStreamSynchronize(m->last_stream));
    if (m->n_points) MOPA_CUDA(cudaMemcpy(p2v_host, m->p2v, (size_t)m->n_points * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int mopa_scn_Metadata_getInputRules(mopa_scn_metadata *m, int32_t *off_host, int32_t *rows_host) {
    MOPA_CHECK(m != nullptr && !m->levels.empty(), "metadata has no input layer");
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_CUDA(cudaStreamSynchronize(m->last_stream));
    MOPA_CUDA(cudaMemcpy(off_host, m->csr_off, (size_t)(m->levels[0].V + 1) * 4, cudaMemcpyDeviceToHost));
    if (m->n_points) MOPA_CUDA(cudaMemcpy(rows_host, m->csr_rows, (size_t)m->n_points * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int mopa_scn_Metadata_getSubmanifoldRuleBook(mopa_scn_metadata *m, int64_t spatial_size, int64_t *counts_host,
                                             int32_t *pairs_host) {
    MOPA_CHECK(m != nullptr, "null metadata");
    int l = m->level_of(spatial_size);
    MOPA_CHECK(l >= 0, "no grid at this spatial size");
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_subm(m, l, m->last_stream));
    const Level &L = m->levels[l];
    return rulebook_to_host(m, L.nbr, L.nbr_ld, L.V, 27, counts_host, pairs_host, m->last_stream);
}

int mopa_scn_Metadata_getTileRuleBook(mopa_scn_metadata *m, int64_t spatial_size, int kind, int64_t *tiles_out,
                                      int32_t *lists_host, uint32_t *masks_host) {
    MOPA_CHECK(m != nullptr && kind >= 0 && kind <= 2, "getTileRuleBook: bad arguments");
    int l = m->level_of(spatial_size);
    MOPA_CHECK(l >= 0, "no grid at this spatial size");
    MOPA_CUDA(cudaSetDevice(m->device));
    if (kind == 0) MOPA_TRY(ensure_subm(m, l, m->last_stream));
    else MOPA_TRY(ensure_down(m, l, m->last_stream));
    const Level &L = m->levels[l];
    const int K = kind == 0 ? 27 : 8;
    const int64_t rows = kind == 1 ? m->levels[l + 1].V : L.V;
    const int64_t tiles = ceil_div(rows > 0 ? rows : 1, kTileRows);
    if (tiles_out) *tiles_out = tiles;
    const int32_t *tl = kind == 0 ? L.tl_subm : (kind == 1 ? L.tl_child : L.tl_sel);
    const uint4 *tm = kind == 0 ? L.tm_subm : (kind == 1 ? L.tm_child : L.tm_sel);
    MOPA_CUDA(cudaStreamSynchronize(m->last_stream));
    if (lists_host) MOPA_CUDA(cudaMemcpy(lists_host, tl, (size_t)tiles * K * kTileRows * 4, cudaMemcpyDeviceToHost));
    if (masks_host) MOPA_CUDA(cudaMemcpy(masks_host, tm, (size_t)tiles * K * 16, cudaMemcpyDeviceToHost));
    return 0;
}

int mopa_scn_Metadata_getConvolutionRuleBook(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t *counts_host,
                                             int32_t *pairs_host) {
    MOPA_CHECK(m != nullptr, "null metadata");
    int l = m->level_of(in_spatial_size);
    MOPA_CHECK(l >= 0, "no grid at this spatial size");
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(ensure_down(m, l, m->last_stream));
    const Level &L = m->levels[l];
    return rulebook_to_host(m, L.child, L.child_ld, m->levels[l + 1].V, 8, counts_host, pairs_host, m->last_stream);
}

}  // extern "C"
