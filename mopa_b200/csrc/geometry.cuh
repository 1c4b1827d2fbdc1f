// Metadata: per-forward GPU hash grids, neighbour tables and rulebooks (replaces [UPSTREAM] Metadata<3>).
#pragma once
#include <vector>

#include "common.cuh"

namespace mopa {

struct Level {
    int64_t spatial = 0;
    int64_t V = 0;                 // active sites (host value; -1 while only the device knows it, see cnt_dev)
    int64_t V_bound = 0;           // upper bound used for allocations / grids of this level's hashing structures
    uint64_t *keys = nullptr;      // [V] packed site keys, id order
    uint64_t *tab_keys = nullptr;  // open-addressing table: key -> id
    int32_t *tab_vals = nullptr;
    uint32_t cap = 0;  // power of two
    // 3x3x3 submanifold neighbour table: nbr[k * nbr_ld + o] = id at coord(o) + delta_k, or -1
    int32_t *nbr = nullptr;
    int64_t nbr_ld = 0;
    // stride-2 link to the next (coarser) level
    bool has_down = false;
    int32_t *parent = nullptr;  // [V] coarse id
    int32_t *kidx = nullptr;    // [V] (x&1)*4 + (y&1)*2 + (z&1)
    int32_t *child = nullptr;   // [8 * child_ld] fine id or -1, indexed by coarse id
    int64_t child_ld = 0;
    // tile rulebooks (what the tcgen05 conv kernel reads): for every 128-row tile t of the OUTPUT rows and offset k,
    //   tl[(t * K + k) * 128 + i], i < n(t, k): the tile's rules at that offset, ascending row: in_row | row_in_tile << 25
    //   tm[t * K + k]: 128-bit mask of the tile rows that have a rule at offset k (n = its popcount)
    int32_t *tl_subm = nullptr;   // rows = this level's sites, K = 27 (from nbr)
    uint4 *tm_subm = nullptr;
    int32_t *tl_child = nullptr;  // rows = the NEXT (coarser) level's sites, K = 8, inputs = this level's sites (from child)
    uint4 *tm_child = nullptr;
    int32_t *tl_sel = nullptr;    // rows = this level's sites, K = 8, inputs = the next level's sites (from parent / kidx)
    uint4 *tm_sel = nullptr;
};
constexpr int kTileRows = 128;
constexpr int kTileRowShift = 25;  // in_row < 2^25 (checked when the lists are built)

}  // namespace mopa

struct mopa_scn_metadata {
    int device = 0;
    std::vector<mopa::Level> levels;  // levels[0] = InputLayer's spatial size, levels[l].spatial = spatial >> l
    int64_t n_points = 0;
    int32_t *p2v = nullptr;       // [n_points]
    int32_t *csr_off = nullptr;   // [V0 + 1]
    int32_t *csr_rows = nullptr;  // [n_points] ascending inside a voxel
    std::vector<void *> allocs;
    cudaStream_t last_stream = nullptr;
    cudaStream_t alloc_stream = nullptr;  // stream of the (last) meta_alloc; see Metadata_delete
    int32_t *pinned = nullptr;  // small host staging block for counts
    cudaEvent_t geom_done = nullptr;  // recorded on the geometry stream when grids/tables are complete
    // device-side site counts: cnt_dev[l] = V of level l, cnt_dev[31] = coordinate error flag. Levels are hashed back to
    // back with upper-bound sizes and the counts come back in ONE read (finish_levels), not one per level.
    int32_t *cnt_dev = nullptr;
    int pending_from = -1;  // first level whose V is still device-only, or -1

    int level_of(int64_t spatial) const {
        for (size_t l = 0; l < levels.size(); ++l)
            if (levels[l].spatial == spatial) return (int)l;
        return -1;
    }
};

namespace mopa {
// how a conv kernel finds the input row of (offset k, output row o)
struct Gather {
    // table mode: table[k * ld + o]; select mode (table == nullptr): kidx[o] == k ? parent[o] : -1
    const int32_t *table = nullptr;
    int64_t ld = 0;
    const int32_t *parent = nullptr;
    const int32_t *kidx = nullptr;
    int volume = 0;     // 27 or 8
    int64_t n_out = 0;  // output rows
    int64_t n_in = 0;   // input rows (for bounds/debug)
    const int32_t *tl = nullptr;  // tile rulebook of the same rules (see Level), always present for library-built gathers
    const uint4 *tm = nullptr;
    int accumulate = 0; // conv_apply: add to the existing output rows instead of overwriting them
    int op = 0;         // 1 submanifold, 2 convolution, 3 deconvolution (profiling tag only)
};

// d_input pass whose output is the gradient of a BatchNormReLU's output: what the conv epilogue needs to reduce the two
// per-column sums of that BatchNorm's backward pass (x = the BatchNorm's INPUT; planes = the conv's output channels)
struct TcBnBwd {
    const float *x;
    int64_t ld_x;
    const float *mean, *invstd, *weight, *bias;
    float leak;
};
int meta_alloc(mopa_scn_metadata *m, void **p, size_t bytes, cudaStream_t s);
int set_locations(mopa_scn_metadata *m, int64_t spatial_size, const int64_t *coords, int64_t n, int ncols,
                  int coords_on_device, cudaStream_t s, bool defer_sync = false);
int conv_apply(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out, const float *weight,
               const float *packed, int n_in0, int n_out0, int transpose, int flip, int precision, cudaStream_t s,
               double *stats = nullptr, bool *stats_done = nullptr, const TcBnBwd *bn = nullptr);
int conv_dweight(const Gather &gt, const float *in, int64_t ld_in, const float *dout, int64_t ld_dout, float *dw,
                 int n_in, int n_out, int precision, void *workspace, size_t workspace_bytes, cudaStream_t s);
size_t dw_workspace_bytes(int volume, int n_in, int n_out, int64_t n_rows);
int pack_weights(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, int precision,
                 float *packed, cudaStream_t s);
bool conv_uses_packed(int c_in, int c_out);
// tcgen05 path (conv_tc.cu): TF32 mode, channel counts that are multiples of 16
bool conv_tc_enabled();
bool conv_tc_supported(int c_in, int c_out);
int64_t tc_packed_floats(int volume, int c_in, int c_out);
int pack_weights_tc(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, float *packed,
                    cudaStream_t s);
// all weight packs of one pass in one launch (program.cu)
constexpr int kTcMaxPackJobs = 40;
struct TcPackJob {
    const float *w;
    float *packed;
    int64_t total;  // filled by pack_weights_tc_batch
    int volume, n_in, n_out, transpose, flip;
};
struct TcPackJobs {
    TcPackJob job[kTcMaxPackJobs];
};
int pack_weights_tc_batch(TcPackJobs &jobs, int n_jobs, cudaStream_t s);
bool conv_packs_tc(int c_in, int c_out, int precision);  // this shape's weights go through pack_weights_tc
int conv_apply_tc(const Gather &gt, const float *in, int64_t ld_in, float *out, int64_t ld_out, const float *packed,
                  int c_in, int c_out, double *stats, cudaStream_t s, const TcBnBwd *bn = nullptr);
// tcgen05 d_weight (conv_dw_tc.cu): TF32 mode, channel counts that are multiples of 16
bool dw_tc_enabled();
bool dw_tc_supported(int n_in, int n_out);
size_t dw_tc_workspace_bytes(int volume, int n_in, int n_out, int64_t n_rows);
int conv_dweight_tc(const Gather &gt, const float *in, int64_t ld_in, const float *dout, int64_t ld_dout, float *dw,
                    int n_in, int n_out, float *partial, cudaStream_t s);
Gather subm_gather(const Level &L);
Gather child_gather(const Level &fine, const Level &coarse, int op = 0);
Gather select_gather(const Level &fine, const Level &coarse, int op = 0);
int bn_forward(const float *in, int64_t ld_in, float *out, int64_t ld_out, float *save_mean, float *save_invstd,
               float *running_mean, float *running_var, const float *weight, const float *bias, float eps, float momentum,
               int train, float leakiness, int64_t n_active, int planes, void *workspace, cudaStream_t s,
               const double *stats = nullptr);
constexpr int kStatsLd = 256;  // per-buffer statistics block: [sum x | sum x^2], kStatsLd doubles each (conv epilogue -> BatchNorm)
int bn_backward(const float *in, int64_t ld_in, float *d_in, int64_t ld_din, const float *d_out, int64_t ld_dout,
                const float *save_mean, const float *save_invstd, const float *weight, const float *bias, float *d_weight,
                float *d_bias, float leakiness, int train, int64_t n_active, int planes, void *workspace, int accumulate,
                cudaStream_t s, const double *sums = nullptr);
size_t bn_workspace_bytes(int planes);
int ensure_subm(mopa_scn_metadata *m, int level, cudaStream_t s);
int ensure_subm_many(mopa_scn_metadata *m, const int *levels, int n, cudaStream_t s);  // one launch for all of them
int ensure_down(mopa_scn_metadata *m, int level, cudaStream_t s);
int ensure_down_async(mopa_scn_metadata *m, int level, cudaStream_t s);  // hashes level + 1 without a host round trip
int finish_levels(mopa_scn_metadata *m, cudaStream_t s);                  // ONE synchronisation: all pending counts
}  // namespace mopa
