// Whole-network executor: runs a compiled scn.Sequential (InputLayer ... OutputLayer) forward or backward as ONE
// C call that enqueues every kernel back to back on the stream.
//
// The reference drives SparseConvNet module by module from Python (mopa/models/scn_unet.py:32-34 ->
// scn.Sequential.forward); at ~90 modules per UNetSCN that costs more host time than the GPU needs for the arithmetic.
// mopa_b200.scn keeps that module tree (state_dict, .train()/.eval(), deepcopy ...) but compiles it once into the op list
// executed here. Same kernels as the per-module ABI; additionally
//   * all grids / neighbour tables of the pyramid are built up front on a dedicated high-priority stream, so the host
//     synchronisations of the build (one count read-back per level) never wait for the previous step's backward pass;
//   * JoinTable is free: the joined buffer is allocated once and its producers write straight into column slices;
//   * gradients that meet at a fan-out are accumulated inside the producing kernels' epilogues.
#include <map>
#include <mutex>

#include <stdlib.h>

#include "geometry.cuh"
#include "mopa_scn.h"

namespace mopa {

enum { OP_SUBM = 1, OP_CONV = 2, OP_DECONV = 3, OP_BN = 4 };
constexpr int kOpWidth = 12, kBufWidth = 4;

struct POp {
    int type, in, out, level_in, level_out, n_in, n_out, param;
    float leak, eps, momentum;
};
struct PBuf {
    int level, channels, parent, col_off;
};
struct Layout {  // resolved per forward from the level sizes
    std::vector<size_t> buf_off;   // byte offset of root buffers in the activation / gradient arena
    std::vector<int64_t> buf_ld;   // row stride (floats) of every buffer
    std::vector<size_t> bn_off;    // per op: byte offset of (save_mean, save_invstd) in the activation arena
    std::vector<size_t> pk_off;    // per conv op: byte offset of its packed weights in the scratch block
    size_t act_bytes = 0, grad_bytes = 0, packed_bytes = 0, dw_bytes = 0, stats_bytes = 0;
};

static std::mutex g_stream_mu;
static std::map<int, cudaStream_t> g_geom_streams;
static int geom_stream(int device, cudaStream_t *out) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_geom_streams.find(device);
    if (it == g_geom_streams.end()) {
        int lo = 0, hi = 0;
        MOPA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        cudaStream_t s;
        MOPA_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
        it = g_geom_streams.emplace(device, s).first;
    }
    *out = it->second;
    return 0;
}

// d_weight kernels run on a second (lowest-priority) stream, forked from the main stream once the op's output gradient
// is complete: they depend on nothing the rest of the backward pass produces, and fill the SMs the last (partial) wave
// of the d_input kernels leaves idle. MOPA_SCN_NO_DW_OVERLAP=1 keeps everything on one stream (A/B measurements).
static std::map<int, cudaStream_t> g_dw_streams;
static int dw_stream(int device, cudaStream_t *out) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_dw_streams.find(device);
    if (it == g_dw_streams.end()) {
        int lo = 0, hi = 0;
        MOPA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        cudaStream_t s;
        MOPA_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, lo));
        it = g_dw_streams.emplace(device, s).first;
    }
    *out = it->second;
    return 0;
}
static bool dw_overlap_enabled() {
    const char *e = getenv("MOPA_SCN_NO_DW_OVERLAP");
    return !(e && e[0] == '1');
}

}  // namespace mopa

struct mopa_scn_program {
    std::vector<mopa::POp> ops;
    std::vector<mopa::PBuf> bufs;
    int in_planes = 0, in_buf = 0, out_buf = 0, n_levels = 1, device = 0;
    int64_t spatial = 0;
    void *bn_ws = nullptr;  // persistent, zero-initialised (the BN kernels leave it zeroed)
    std::vector<cudaEvent_t> fork_events;  // one per op (main stream -> d_weight stream), created on first use
    cudaEvent_t join_event = nullptr;
};

namespace mopa {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int make_layout(const mopa_scn_program *p, const mopa_scn_metadata *m, int precision, Layout &L) {
    MOPA_CHECK((int)m->levels.size() >= p->n_levels, "metadata has fewer levels than the program needs");
    const size_t nb = p->bufs.size();
    L.buf_off.assign(nb, 0);
    L.buf_ld.assign(nb, 0);
    size_t off = 0;
    for (size_t b = 0; b < nb; ++b) {
        const PBuf &B = p->bufs[b];
        if (B.parent >= 0) continue;
        L.buf_ld[b] = round_up(B.channels, 4);
        L.buf_off[b] = off;
        off += align256((size_t)m->levels[B.level].V * L.buf_ld[b] * 4);
    }
    for (size_t b = 0; b < nb; ++b) {
        const PBuf &B = p->bufs[b];
        if (B.parent < 0) continue;
        MOPA_CHECK(p->bufs[B.parent].parent < 0, "nested joins are not supported");
        L.buf_ld[b] = L.buf_ld[B.parent];
        L.buf_off[b] = L.buf_off[B.parent] + (size_t)B.col_off * 4;
    }
    L.grad_bytes = off + 256;
    for (const POp &o : p->ops)  // fail up front, not in the middle of a pass (MODEL_3D.SCN.m >= 22 reaches 12 m > 256 planes)
        MOPA_CHECK(o.type != OP_BN || (o.n_in >= 1 && o.n_in <= 256), "BatchNormalization over more than 256 planes is not supported");
    L.bn_off.assign(p->ops.size(), 0);
    L.pk_off.assign(p->ops.size(), 0);
    L.packed_bytes = 0;
    L.dw_bytes = 256;
    for (size_t i = 0; i < p->ops.size(); ++i) {
        const POp &o = p->ops[i];
        if (o.type == OP_BN) {
            L.bn_off[i] = off;
            off += align256((size_t)2 * o.n_in * 4);
        } else {
            const int volume = o.type == OP_SUBM ? 27 : 8;
            L.pk_off[i] = L.packed_bytes;  // every op keeps its own slot: all packs of a pass are one launch
            L.packed_bytes += align256((size_t)mopa_scn_packedWeightFloats(volume, o.n_in, o.n_out, precision) * 4);
            const int64_t rows = m->levels[o.level_out].V;  // d_weight chunks run over the op's OUTPUT rows
            const size_t dw = dw_workspace_bytes(volume, o.n_in, o.n_out, rows);
            if (dw > L.dw_bytes) L.dw_bytes = align256(dw);
        }
    }
    L.packed_bytes += 256;
    // forward reuses the d_weight workspace for the BatchNorm statistics blocks (one per buffer, conv epilogue -> BN)
    const size_t stats_bytes = align256(nb * (size_t)2 * kStatsLd * sizeof(double));
    if (stats_bytes > L.dw_bytes) L.dw_bytes = stats_bytes;
    L.stats_bytes = stats_bytes;  // backward: its own block behind the d_weight workspace (which is busy on the side stream)
    L.act_bytes = off + 256;
    return 0;
}

// packs the weights of every convolution of a pass: the tcgen05 layouts in one launch, the others one by one
static int pack_all(const mopa_scn_program *p, const Layout &L, const void *const *params, int precision, bool backward,
                    const std::vector<char> *wanted, void *scratch, cudaStream_t s) {
    TcPackJobs jobs;
    int n = 0;
    for (size_t i = 0; i < p->ops.size(); ++i) {
        const POp &o = p->ops[i];
        if (o.type == OP_BN || (wanted && !(*wanted)[i])) continue;
        const int volume = o.type == OP_SUBM ? 27 : 8;
        const int c_in = backward ? o.n_out : o.n_in, c_out = backward ? o.n_in : o.n_out;
        if (!conv_uses_packed(c_in, c_out)) continue;
        float *dst = reinterpret_cast<float *>(reinterpret_cast<char *>(scratch) + L.pk_off[i]);
        const float *w = (const float *)params[o.param];
        const int flip = backward && o.type == OP_SUBM ? 1 : 0;
        if (conv_packs_tc(c_in, c_out, precision)) {
            if (n == kTcMaxPackJobs) {
                MOPA_TRY(pack_weights_tc_batch(jobs, n, s));
                n = 0;
            }
            jobs.job[n++] = TcPackJob{w, dst, 0, volume, o.n_in, o.n_out, backward ? 1 : 0, flip};
        } else {
            MOPA_TRY(pack_weights(w, volume, o.n_in, o.n_out, backward ? 1 : 0, flip, precision, dst, s));
        }
    }
    return pack_weights_tc_batch(jobs, n, s);
}

struct BufView {
    float *ptr;
    int64_t ld;
};
static BufView view(const Layout &L, void *arena, int b) {
    return BufView{reinterpret_cast<float *>(reinterpret_cast<char *>(arena) + L.buf_off[b]), L.buf_ld[b]};
}

static Gather op_gather(const POp &o, const mopa_scn_metadata *m, bool for_dinput) {
    // forward / d_weight index by the op's OUTPUT rows; d_input by its INPUT rows
    if (o.type == OP_SUBM) return subm_gather(m->levels[o.level_in]);
    const bool down = o.type == OP_CONV;  // fine -> coarse
    const Level &fine = m->levels[down ? o.level_in : o.level_out];
    const Level &coarse = m->levels[down ? o.level_out : o.level_in];
    const bool rows_are_coarse = down != for_dinput;
    Gather g = rows_are_coarse ? child_gather(fine, coarse) : select_gather(fine, coarse);
    g.op = down ? 2 : 3;
    return g;
}

}  // namespace mopa

using namespace mopa;

extern "C" {

mopa_scn_program *mopa_scn_Program_new(const int32_t *ops, int n_ops, const int32_t *bufs, int n_bufs, int in_planes,
                                       int in_buf, int out_buf, int64_t spatial_size, int n_levels, int device) {
    auto *p = new mopa_scn_program();
    p->in_planes = in_planes; p->in_buf = in_buf; p->out_buf = out_buf; p->spatial = spatial_size; p->n_levels = n_levels;
    p->device = device;
    for (int i = 0; i < n_ops; ++i) {
        const int32_t *r = ops + (size_t)i * kOpWidth;
        POp o;
        o.type = r[0]; o.in = r[1]; o.out = r[2]; o.level_in = r[4]; o.level_out = r[5]; o.n_in = r[6]; o.n_out = r[7];
        o.param = r[8];
        memcpy(&o.leak, r + 9, 4); memcpy(&o.eps, r + 10, 4); memcpy(&o.momentum, r + 11, 4);
        p->ops.push_back(o);
    }
    for (int i = 0; i < n_bufs; ++i) {
        const int32_t *r = bufs + (size_t)i * kBufWidth;
        p->bufs.push_back(PBuf{r[0], r[1], r[2], r[3]});
    }
    const size_t ws = bn_workspace_bytes(256);
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p->bn_ws, ws) != cudaSuccess ||
        cudaMemset(p->bn_ws, 0, ws) != cudaSuccess) {
        fail(__FILE__, __LINE__, "Program_new: cannot allocate the BatchNorm workspace on the device");
        delete p;
        return nullptr;
    }
    return p;
}

void mopa_scn_Program_delete(mopa_scn_program *p) {
    if (!p) return;
    if (p->bn_ws) cudaFree(p->bn_ws);
    for (cudaEvent_t e : p->fork_events)
        if (e) cudaEventDestroy(e);
    if (p->join_event) cudaEventDestroy(p->join_event);
    delete p;
}

// Voxelise + build every level / table the program uses (geometry stream), then size the arenas.
// sizes_out: [act_bytes, grad_bytes, scratch_bytes]; n_active_out: n_levels entries.
int mopa_scn_Program_prepare(mopa_scn_program *p, mopa_scn_metadata *m, const int64_t *coords, int64_t n, int ncols,
                             int coords_on_device, int precision, void *stream, int64_t *n_active_out,
                             uint64_t *sizes_out) {
    MOPA_CHECK(p && m, "null program / metadata");
    cudaStream_t main = (cudaStream_t)stream, gs = nullptr;
    MOPA_CUDA(cudaSetDevice(m->device));
    MOPA_TRY(geom_stream(m->device, &gs));
    m->last_stream = main;
    if (coords_on_device == 1) {  // the coordinates may have been produced on the caller's stream (2: known complete)
        if (!m->geom_done) MOPA_CUDA(cudaEventCreateWithFlags(&m->geom_done, cudaEventDisableTiming));
        MOPA_CUDA(cudaEventRecord(m->geom_done, main));
        MOPA_CUDA(cudaStreamWaitEvent(gs, m->geom_done, 0));
    }
    // the whole pyramid is hashed back to back; ONE host round trip brings every level's site count (and the coordinate
    // error flag) back, instead of one per level
    MOPA_TRY(set_locations(m, p->spatial, coords, n, ncols, coords_on_device != 0, gs, /*defer_sync=*/true));
    for (int l = 0; l + 1 < p->n_levels; ++l) MOPA_TRY(ensure_down_async(m, l, gs));
    MOPA_TRY(finish_levels(m, gs));
    bool need_subm[32] = {false};
    for (const POp &o : p->ops)
        if (o.type == OP_SUBM) need_subm[o.level_in] = true;
    int subm_levels[32], n_subm = 0;
    for (int l = 0; l < p->n_levels; ++l)
        if (need_subm[l]) subm_levels[n_subm++] = l;
    MOPA_TRY(ensure_subm_many(m, subm_levels, n_subm, gs));
    if (!m->geom_done) MOPA_CUDA(cudaEventCreateWithFlags(&m->geom_done, cudaEventDisableTiming));
    MOPA_CUDA(cudaEventRecord(m->geom_done, gs));
    MOPA_CUDA(cudaStreamWaitEvent(main, m->geom_done, 0));
    Layout L;
    MOPA_TRY(make_layout(p, m, precision, L));
    for (int l = 0; l < p->n_levels; ++l) n_active_out[l] = m->levels[l].V;
    sizes_out[0] = L.act_bytes;
    sizes_out[1] = L.grad_bytes;
    sizes_out[2] = L.packed_bytes + L.dw_bytes + L.stats_bytes;
    return 0;
}

// params[op.param ...]: conv -> {weight}; BN -> {weight, bias, running_mean, running_var}
int mopa_scn_Program_forward(mopa_scn_program *p, mopa_scn_metadata *m, const float *feats, int64_t ld_feats,
                             const void *const *params, int train, int precision, void *act_arena, void *scratch,
                             float *out, int64_t ld_out, void *stream) {
    MOPA_CHECK(p && m && act_arena && scratch, "Program_forward: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    MOPA_CUDA(cudaSetDevice(m->device));
    m->last_stream = s;
    Layout L;
    MOPA_TRY(make_layout(p, m, precision, L));
    MOPA_TRY(pack_all(p, L, params, precision, false, nullptr, scratch, s));
    // BatchNorm statistics ride on the producing convolution's epilogue (tcgen05 kernel, train mode): one block of
    // [sum x | sum x^2] per buffer, children of a joined buffer own column ranges of their parent's block
    const char *nofuse = getenv("MOPA_SCN_NO_BNSTATS_FUSION");  // read per call (A/B measurements, bitwise eager/compiled test)
    const bool fuse_stats = !(nofuse && nofuse[0] == '1');
    const size_t nb = p->bufs.size();
    double *stats_base = reinterpret_cast<double *>(reinterpret_cast<char *>(scratch) + L.packed_bytes);
    std::vector<char> stats_ok(nb, 0);
    const bool want_stats = train && fuse_stats;
    if (want_stats) MOPA_CUDA(cudaMemsetAsync(stats_base, 0, nb * (size_t)2 * kStatsLd * sizeof(double), s));
    auto stats_of = [&](int b) -> double * {  // nullptr: the (joined) buffer is wider than a statistics block
        const PBuf &B = p->bufs[b];
        const int root = B.parent >= 0 ? B.parent : b;
        if (p->bufs[root].channels > kStatsLd) return nullptr;
        return stats_base + (size_t)root * 2 * kStatsLd + (B.parent >= 0 ? B.col_off : 0);
    };
    auto stats_ready = [&](int b) {  // every column of buffer b has been accumulated
        if (stats_ok[b]) return true;
        int covered = 0, kids = 0;
        for (size_t c = 0; c < nb; ++c)
            if (p->bufs[c].parent == b) { ++kids; if (stats_ok[c]) covered += p->bufs[c].channels; }
        return kids > 0 && covered == p->bufs[b].channels;
    };
    BufView b0 = view(L, act_arena, p->in_buf);
    MOPA_TRY(mopa_scn_InputLayer_updateOutput(m, feats, ld_feats, p->in_planes, b0.ptr, b0.ld, s));
    for (size_t i = 0; i < p->ops.size(); ++i) {
        const POp &o = p->ops[i];
        BufView in = view(L, act_arena, o.in), ob = view(L, act_arena, o.out);
        if (o.type == OP_BN) {
            float *save = reinterpret_cast<float *>(reinterpret_cast<char *>(act_arena) + L.bn_off[i]);
            MOPA_TRY(bn_forward(in.ptr, in.ld, ob.ptr, ob.ld, save, save + o.n_in, (float *)params[o.param + 2],
                                (float *)params[o.param + 3], (const float *)params[o.param],
                                (const float *)params[o.param + 1], o.eps, o.momentum, train, o.leak,
                                m->levels[o.level_in].V, o.n_in, p->bn_ws, s,
                                want_stats && stats_ready(o.in) ? stats_of(o.in) : nullptr));
            continue;
        }
        const float *w = (const float *)params[o.param];
        const float *pk = conv_uses_packed(o.n_in, o.n_out)
                              ? reinterpret_cast<const float *>(reinterpret_cast<const char *>(scratch) + L.pk_off[i])
                              : nullptr;
        bool done = false;
        double *st_out = want_stats ? stats_of(o.out) : nullptr;
        MOPA_TRY(conv_apply(op_gather(o, m, false), in.ptr, in.ld, ob.ptr, ob.ld, w, pk, o.n_in, o.n_out, 0, 0, precision, s,
                            st_out, &done));
        stats_ok[o.out] = done && st_out != nullptr;
    }
    BufView last = view(L, act_arena, p->out_buf);
    return mopa_scn_OutputLayer_updateOutput(m, last.ptr, last.ld, p->bufs[p->out_buf].channels, out, ld_out, s);
}

// param_grads mirrors params (entries may be NULL: no gradient wanted). d_feats may be NULL.
int mopa_scn_Program_backward(mopa_scn_program *p, mopa_scn_metadata *m, const void *const *params,
                              void *const *param_grads, int train, int precision, const void *act_arena,
                              void *grad_arena, void *scratch, const float *d_out, int64_t ld_dout, float *d_feats,
                              int64_t ld_dfeats, void *stream) {
    MOPA_CHECK(p && m && act_arena && grad_arena && scratch, "Program_backward: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    MOPA_CUDA(cudaSetDevice(m->device));
    m->last_stream = s;
    Layout L;
    MOPA_TRY(make_layout(p, m, precision, L));
    void *dw_ws = reinterpret_cast<char *>(scratch) + L.packed_bytes;
    {  // transposed (and for 3x3x3, offset-flipped) weights of every op that produces an input gradient, one launch
        std::vector<char> wanted(p->ops.size(), 0);
        for (size_t i = 0; i < p->ops.size(); ++i)
            wanted[i] = p->ops[i].type != OP_BN && (p->ops[i].in != p->in_buf || d_feats != nullptr);
        MOPA_TRY(pack_all(p, L, params, precision, true, &wanted, scratch, s));
    }
    void *act = const_cast<void *>(act_arena);
    const size_t nb = p->bufs.size();
    std::vector<char> written(nb, 0);
    auto mark = [&](int b) {
        written[b] = 1;
        for (size_t c = 0; c < nb; ++c)
            if (p->bufs[c].parent == b) written[c] = 1;
    };
    cudaStream_t s2 = s;  // stream of the d_weight kernels
    bool forked = false;
    if (dw_overlap_enabled()) {
        MOPA_TRY(dw_stream(m->device, &s2));
        if (p->fork_events.size() != p->ops.size()) p->fork_events.assign(p->ops.size(), nullptr);
        if (!p->join_event) MOPA_CUDA(cudaEventCreateWithFlags(&p->join_event, cudaEventDisableTiming));
    }
    // BatchNorm backward sums ride on the epilogue of the d_input convolution that produces the BatchNorm's output
    // gradient (tcgen05 kernel): one [S1 | S2] block per buffer; the BatchNorm backward is then a single streaming kernel
    const char *nofuse = getenv("MOPA_SCN_NO_BNSTATS_FUSION");  // read per call, as in the forward pass
    // Off by default: measured (profiles/r02_bn_bwd_fusion.txt) the epilogue work costs the d_input kernels what the
    // BatchNorm statistics kernels save, and d_input time is harder to overlap with the d_weight stream than BatchNorm time.
    // MOPA_SCN_BNBWD_FUSION=1 turns it on.
    const char *fuse_b = getenv("MOPA_SCN_BNBWD_FUSION");
    const bool fuse_stats = !(nofuse && nofuse[0] == '1') && (fuse_b && fuse_b[0] == '1');
    double *bwd_stats = reinterpret_cast<double *>(reinterpret_cast<char *>(scratch) + L.packed_bytes + L.dw_bytes);
    std::vector<char> sums_ok(nb, 0);
    std::vector<int> bn_of(nb, -1);  // buffer -> the BatchNorm op that wrote it in the forward pass
    if (fuse_stats) {
        MOPA_CUDA(cudaMemsetAsync(bwd_stats, 0, nb * (size_t)2 * kStatsLd * sizeof(double), s));
        for (size_t i = 0; i < p->ops.size(); ++i)
            if (p->ops[i].type == OP_BN && p->bufs[p->ops[i].out].parent < 0) bn_of[p->ops[i].out] = (int)i;
        // only where ONE op reads the BatchNorm's output (a second reader would add to the gradient after the sums)
        std::vector<int> readers(nb, 0);
        for (const POp &o : p->ops) {
            ++readers[o.in];
            if (p->bufs[o.in].parent >= 0) ++readers[p->bufs[o.in].parent];
        }
        for (size_t b = 0; b < nb; ++b)
            if (readers[b] != 1 || (int)b == p->out_buf) bn_of[b] = -1;
    }
    BufView g_last = view(L, grad_arena, p->out_buf);
    MOPA_TRY(mopa_scn_OutputLayer_updateGradInput(m, g_last.ptr, g_last.ld, d_out, ld_dout, p->bufs[p->out_buf].channels, s));
    mark(p->out_buf);
    for (int i = (int)p->ops.size() - 1; i >= 0; --i) {
        const POp &o = p->ops[i];
        const int volume = o.type == OP_SUBM ? 27 : 8;
        if (!written[o.out]) {  // output feeds nothing that reaches the loss: parameters get a zero gradient
            const size_t n = o.type == OP_BN ? (size_t)o.n_in : (size_t)volume * o.n_in * o.n_out;
            if (param_grads[o.param]) MOPA_CUDA(cudaMemsetAsync(param_grads[o.param], 0, n * 4, s));
            if (o.type == OP_BN && param_grads[o.param + 1]) MOPA_CUDA(cudaMemsetAsync(param_grads[o.param + 1], 0, n * 4, s));
            continue;
        }
        if (p->bufs[o.in].parent < 0) {  // overwriting a joined buffer would clobber slices that were already written
            for (size_t c = 0; c < nb; ++c)
                MOPA_CHECK(!(p->bufs[c].parent == o.in && written[c] && !written[o.in]),
                           "unsupported program: a joined buffer receives its gradient after one of its slices");
        }
        const int accumulate = written[o.in];
        const bool need_din = o.in != p->in_buf || d_feats != nullptr;
        BufView x = view(L, act, o.in), dx = view(L, grad_arena, o.in), dy = view(L, grad_arena, o.out);
        if (o.type == OP_BN) {
            const float *save = reinterpret_cast<const float *>(reinterpret_cast<const char *>(act) + L.bn_off[i]);
            MOPA_TRY(bn_backward(x.ptr, x.ld, need_din ? dx.ptr : nullptr, dx.ld, dy.ptr, dy.ld, save, save + o.n_in,
                                 (const float *)params[o.param], (const float *)params[o.param + 1],
                                 (float *)param_grads[o.param], (float *)param_grads[o.param + 1], o.leak, train,
                                 m->levels[o.level_in].V, o.n_in, p->bn_ws, accumulate, s,
                                 sums_ok[o.out] ? bwd_stats + (size_t)o.out * 2 * kStatsLd : nullptr));
        } else {
            const float *w = (const float *)params[o.param];
            // d_weight first: it is forked to its own stream and overlaps the d_input kernel of the same op
            if (param_grads[o.param]) {
                if (s2 != s) {  // dy is complete on the main stream here; nothing later in this pass writes it again
                    if (!p->fork_events[i]) MOPA_CUDA(cudaEventCreateWithFlags(&p->fork_events[i], cudaEventDisableTiming));
                    MOPA_CUDA(cudaEventRecord(p->fork_events[i], s));
                    MOPA_CUDA(cudaStreamWaitEvent(s2, p->fork_events[i], 0));
                    forked = true;
                }
                MOPA_TRY(conv_dweight(op_gather(o, m, false), x.ptr, x.ld, dy.ptr, dy.ld, (float *)param_grads[o.param],
                                      o.n_in, o.n_out, precision, dw_ws, L.dw_bytes, s2));
            }
            if (need_din) {
                const float *pk = conv_uses_packed(o.n_out, o.n_in)
                                      ? reinterpret_cast<const float *>(reinterpret_cast<const char *>(scratch) + L.pk_off[i])
                                      : nullptr;
                Gather g = op_gather(o, m, true);
                g.accumulate = accumulate;
                // this op is the only producer of d(o.in), and o.in is the output of a BatchNormReLU: reduce that
                // BatchNorm's backward sums while the gradient rows are still in registers
                const int bi = accumulate ? -1 : bn_of[o.in];
                TcBnBwd bn{};
                double *sums = nullptr;
                bool done = false;
                if (bi >= 0 && p->ops[bi].n_in == o.n_in && o.n_in <= kStatsLd) {
                    const POp &b = p->ops[bi];
                    const float *save = reinterpret_cast<const float *>(reinterpret_cast<const char *>(act) + L.bn_off[bi]);
                    BufView bx = view(L, act, b.in);
                    bn = TcBnBwd{bx.ptr, bx.ld, save, save + b.n_in, (const float *)params[b.param],
                                 (const float *)params[b.param + 1], b.leak};
                    sums = bwd_stats + (size_t)o.in * 2 * kStatsLd;
                }
                MOPA_TRY(conv_apply(g, dy.ptr, dy.ld, dx.ptr, dx.ld, w, pk, o.n_in, o.n_out, 1, o.type == OP_SUBM ? 1 : 0,
                                    precision, s, sums, &done, sums ? &bn : nullptr));
                if (sums && done) sums_ok[o.in] = 1;
            }
        }
        if (need_din) mark(o.in);
    }
    if (forked) {  // the caller's stream owns every buffer again only after the d_weight stream has drained
        MOPA_CUDA(cudaEventRecord(p->join_event, s2));
        MOPA_CUDA(cudaStreamWaitEvent(s, p->join_event, 0));
    }
    if (d_feats) {
        BufView g0 = view(L, grad_arena, p->in_buf);
        MOPA_CHECK(written[p->in_buf], "input gradient requested but nothing reaches the input");
        MOPA_TRY(mopa_scn_InputLayer_updateGradInput(m, d_feats, ld_dfeats, g0.ptr, g0.ld, p->in_planes, s));
    }
    return 0;
}

}  // extern "C"
