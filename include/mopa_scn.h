/* mopa_scn.h -- C ABI of libmopa_scn.so, the B200 (sm_100a) replacement for the native layer behind
 * `import sparseconvnet as scn` on MoPA's 3D-branch hot path.
 *
 * The reference reaches native code only through the SparseConvNet Python modules assembled at
 *   /root/reference/mopa/models/scn_unet.py:25-30   (InputLayer, SubmanifoldConvolution, UNet, BatchNormReLU, OutputLayer)
 * and, through scn.UNet, Convolution / Deconvolution / BatchNormLeakyReLU / ConcatTable / JoinTable.
 * Upstream binds those modules to a pybind11 module `sparseconvnet.SCN` ([UPSTREAM] sparseconvnet/SCN/pybind.cpp,
 * sparseconvnet.h; not vendored under /root/reference). Every entry point below names
 * the upstream binding it replaces; argument meaning follows that binding, with these deliberate differences:
 *   - plain C: raw pointers + sizes + an explicit CUDA stream, no torch types, status code + last_error();
 *   - the callee never allocates feature tensors: a *_prepare / setLocations call returns the row count, the
 *     caller (PyTorch) allocates, then *_updateOutput fills;
 *   - coordinates are hashed ON THE GPU (upstream builds grids and rulebooks on the host even for CUDA tensors);
 *   - feature matrices carry an explicit row stride (ld, in floats) so JoinTable can be a view of one buffer.
 *
 * All feature / weight / gradient pointers are DEVICE pointers to float32 unless a parameter says HOST.
 * All functions return 0 on success, non-zero on error (message via mopa_scn_last_error(), thread-local).
 * The library is re-entrant per metadata handle and keeps no global mutable state besides the per-thread error string.
 */
#ifndef MOPA_SCN_H_
#define MOPA_SCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOPA_SCN_ABI_VERSION 2

/* arithmetic mode of the conv contractions (fp32 storage and fp32 accumulation in both) */
#define MOPA_SCN_PREC_FP32 0 /* 3xTF32 split-operand MMA: fp32-equivalent products (parity mode) */
#define MOPA_SCN_PREC_TF32 1 /* single TF32 MMA, operands rounded to nearest tf32 (default, fast) */

typedef struct mopa_scn_metadata mopa_scn_metadata; /* replaces [UPSTREAM] Metadata<3> (Metadata/Metadata.h) */

int mopa_scn_abi_version(void);
const char *mopa_scn_last_error(void);

/* ---- Metadata<3>: owns the per-level GPU hash grids, neighbour tables and rulebooks of ONE forward ------------- */
mopa_scn_metadata *mopa_scn_Metadata_new(int dimension /* must be 3 */, int device);
void mopa_scn_Metadata_delete(mopa_scn_metadata *m);

/* replaces InputLayer_updateOutput's rule-building half ([UPSTREAM] IOLayersRules.h::inputLayerRules, mode 4).
 * coords: int64 (n, ncols), ncols 3 or 4 (x, y, z[, batch]) -- collate.py:182-186 layout. coords_on_device = 0 means a
 * HOST pointer (what the reference passes; copied H2D on `stream`). Voxel ids = order of first occurrence.
 * Fails (non-zero) on coordinates outside [0, spatial_size) -- the reference filters them (nuscenes_dataloader.py:422).
 * Synchronises `stream` once to return the active-site count. */
int mopa_scn_InputLayer_setLocations(mopa_scn_metadata *m, int64_t spatial_size, const int64_t *coords, int64_t n,
                                     int ncols, int coords_on_device, int mode /* 4 */, void *stream,
                                     int64_t *n_active_out);
/* replaces InputLayer_updateOutput (feature half): out[v] = sum_{i in v, ascending} (1/n_v) * in[i]; in (n, planes) */
int mopa_scn_InputLayer_updateOutput(mopa_scn_metadata *m, const float *in, int64_t ld_in, int planes, float *out,
                                     int64_t ld_out, void *stream);
/* replaces InputLayer_updateGradInput: d_in[i] = (1/n_v) * d_out[voxel(i)] */
int mopa_scn_InputLayer_updateGradInput(mopa_scn_metadata *m, float *d_in, int64_t ld_din, const float *d_out,
                                        int64_t ld_dout, int planes, void *stream);
/* replaces OutputLayer_updateOutput: out[i] = in[voxel(i)]  (n rows out) */
int mopa_scn_OutputLayer_updateOutput(mopa_scn_metadata *m, const float *in, int64_t ld_in, int planes, float *out,
                                      int64_t ld_out, void *stream);
/* replaces OutputLayer_updateGradInput: d_in[v] = sum_{i in v, ascending} d_out[i] */
int mopa_scn_OutputLayer_updateGradInput(mopa_scn_metadata *m, float *d_in, int64_t ld_din, const float *d_out,
                                         int64_t ld_dout, int planes, void *stream);

/* replaces Metadata::getSubmanifoldRuleBook (built lazily, cached per spatial size). filter_size must be 3. */
int mopa_scn_Metadata_prepareSubmanifold(mopa_scn_metadata *m, int64_t spatial_size, int filter_size, void *stream,
                                         int64_t *n_active_out);
/* replaces Metadata::getRuleBook(in, out, size, stride) for size == stride == 2: creates the out_size grid if absent. */
int mopa_scn_Metadata_prepareConvolution(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                         int filter_size, int filter_stride, void *stream, int64_t *n_active_out);
/* number of active sites at a spatial size already built, or -1 */
int64_t mopa_scn_Metadata_getNActive(mopa_scn_metadata *m, int64_t spatial_size);
int64_t mopa_scn_Metadata_getNPoints(mopa_scn_metadata *m);

/* ---- inspection (parity tests): copy integer structures to HOST buffers; each synchronises ------------------- */
/* voxel coordinates (n_active, 4) int64 [x, y, z, batch] in id order */
int mopa_scn_Metadata_getSpatialLocations(mopa_scn_metadata *m, int64_t spatial_size, int64_t *coords_host);
/* point -> voxel id (n_points) */
int mopa_scn_Metadata_getPointToVoxel(mopa_scn_metadata *m, int32_t *p2v_host);
/* per-voxel contributing rows, CSR: off (n_active + 1), rows (n_points) ascending within a voxel */
int mopa_scn_Metadata_getInputRules(mopa_scn_metadata *m, int32_t *off_host, int32_t *rows_host);
/* submanifold rulebook: counts_host[27]; pairs_host (sum counts, 2) int32 [in, out], offset-major, ascending out.
 * pairs_host may be NULL to query counts only. */
int mopa_scn_Metadata_getSubmanifoldRuleBook(mopa_scn_metadata *m, int64_t spatial_size, int64_t *counts_host,
                                             int32_t *pairs_host);
/* strided rulebook in_size -> in_size/2: counts_host[8]; pairs (n_active(in), 2) [fine, coarse], offset-major */
int mopa_scn_Metadata_getConvolutionRuleBook(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t *counts_host,
                                             int32_t *pairs_host);

/* tile rulebooks (the form the tcgen05 conv kernels consume; no upstream counterpart -- upstream's kernels walk the pair
 * lists above): for every tile of 128 output rows and filter offset k, lists_host[(t * K + k) * 128 + i], i < n(t, k), holds
 * the tile's rules at that offset in ascending row order as in_row | (row_in_tile << 25), and masks_host[(t * K + k) * 4 ..]
 * the 128-bit mask of the tile rows that have a rule (entries beyond n(t, k) are unspecified). kind 0: 3x3x3 submanifold
 * at spatial_size (K = 27); kind 1: Convolution rules spatial_size -> spatial_size/2 indexed by COARSE rows (K = 8);
 * kind 2: the same rules indexed by FINE rows (Deconvolution forward / Convolution input gradient, K = 8).
 * tiles_out receives the tile count; pass NULL buffers to query it. */
int mopa_scn_Metadata_getTileRuleBook(mopa_scn_metadata *m, int64_t spatial_size, int kind, int64_t *tiles_out,
                                      int32_t *lists_host, uint32_t *masks_host);

/* ---- weights: repack (volume, nIn, nOut) fp32 into the MMA fragment order the conv kernels stream -------------
 * transpose = 0: forward operand. transpose = 1: operand of the input-gradient pass (W[k]^T; for submanifold
 * filters additionally offset-flipped, k -> volume-1-k). Output size from _packedWeightFloats. */
int64_t mopa_scn_packedWeightFloats(int volume, int n_in, int n_out, int precision);
int mopa_scn_packWeights(const float *weight, int volume, int n_in, int n_out, int transpose, int flip, int precision,
                         float *packed, void *stream);

/* ---- convolutions ---------------------------------------------------------------------------------------------
 * *_updateOutput replace SubmanifoldConvolution_updateOutput / Convolution_updateOutput / Deconvolution_updateOutput
 * ([UPSTREAM] CUDA/Convolution.cu, Deconvolution.cu); bias is not supported (scn_unet.py passes bias=False everywhere).
 * `packed` = mopa_scn_packWeights(weight, transpose=0). Output rows are written once (no atomics, k ascending). */
int mopa_scn_SubmanifoldConvolution_updateOutput(mopa_scn_metadata *m, int64_t spatial_size, int filter_size,
                                                 const float *in, int64_t ld_in, float *out, int64_t ld_out,
                                                 const float *weight, const float *packed, int n_in, int n_out,
                                                 int precision, void *stream);
int mopa_scn_Convolution_updateOutput(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                      int filter_size, int filter_stride, const float *in, int64_t ld_in, float *out,
                                      int64_t ld_out, const float *weight, const float *packed, int n_in, int n_out,
                                      int precision, void *stream);
/* in_spatial_size is the COARSE size, out_spatial_size the fine one (existing grid) */
int mopa_scn_Deconvolution_updateOutput(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                        int filter_size, int filter_stride, const float *in, int64_t ld_in, float *out,
                                        int64_t ld_out, const float *weight, const float *packed, int n_in, int n_out,
                                        int precision, void *stream);

/* *_backward replace the matching upstream *_backward. d_in may be NULL (input needs no gradient); d_weight may be
 * NULL. `packed_t` = mopa_scn_packWeights(weight, transpose=1 [, flip=1 for submanifold]). d_weight (volume, nIn, nOut)
 * is OVERWRITTEN (deterministic two-stage reduction; autograd accumulates). workspace: device scratch of at least
 * mopa_scn_backwardWorkspaceBytes(...) bytes. */
size_t mopa_scn_backwardWorkspaceBytes(int volume, int n_in, int n_out, int64_t n_rules);
/* Host-only inspection of how the tcgen05 d_weight kernel splits `n_rows` output rows of a `volume`-offset rulebook into
 * work items (subm_table != 0: 3x3x3 submanifold table whose centre offset carries one rule per row and is split finer).
 * plan_out[6] = {centre offset or -1, rows per item, items per ordinary offset, rows per centre item, centre items,
 * total items}. No device call; used by the CPU tests (every row of every offset is covered exactly once). */
int mopa_scn_debug_dweightPlan(int volume, int subm_table, int64_t n_rows, int *plan_out);
int mopa_scn_SubmanifoldConvolution_backward(mopa_scn_metadata *m, int64_t spatial_size, int filter_size,
                                             const float *in, int64_t ld_in, float *d_in, int64_t ld_din,
                                             const float *d_out, int64_t ld_dout, const float *weight,
                                             const float *packed_t, float *d_weight, int n_in, int n_out,
                                             int precision, void *workspace, size_t workspace_bytes, void *stream);
int mopa_scn_Convolution_backward(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                  int filter_size, int filter_stride, const float *in, int64_t ld_in, float *d_in,
                                  int64_t ld_din, const float *d_out, int64_t ld_dout, const float *weight,
                                  const float *packed_t, float *d_weight, int n_in, int n_out, int precision,
                                  void *workspace, size_t workspace_bytes, void *stream);
int mopa_scn_Deconvolution_backward(mopa_scn_metadata *m, int64_t in_spatial_size, int64_t out_spatial_size,
                                    int filter_size, int filter_stride, const float *in, int64_t ld_in, float *d_in,
                                    int64_t ld_din, const float *d_out, int64_t ld_dout, const float *weight,
                                    const float *packed_t, float *d_weight, int n_in, int n_out, int precision,
                                    void *workspace, size_t workspace_bytes, void *stream);
/* rule count of a prepared rulebook (host value), for sizing the backward workspace; -1 if not built */
int64_t mopa_scn_Metadata_getSubmanifoldRuleCount(mopa_scn_metadata *m, int64_t spatial_size);

/* ---- BatchNormalization (+ leaky ReLU) -------------------------------------------------------------------------
 * replaces BatchNormalization_updateOutput / _backward ([UPSTREAM] CUDA/BatchNormalization.cu): per-plane statistics
 * over the n_active rows; eps inside the sqrt; `momentum` is the KEEP fraction of the running stats (0.9);
 * running_var uses the unbiased estimate; out = y > 0 ? y : leakiness * y. workspace: >= mopa_scn_bnWorkspaceBytes. */
size_t mopa_scn_bnWorkspaceBytes(int planes);
int mopa_scn_BatchNormalization_updateOutput(const float *in, int64_t ld_in, float *out, int64_t ld_out,
                                             float *save_mean, float *save_invstd, float *running_mean,
                                             float *running_var, const float *weight, const float *bias, float eps,
                                             float momentum, int train, float leakiness, int64_t n_active, int planes,
                                             void *workspace, size_t workspace_bytes, void *stream);
int mopa_scn_BatchNormalization_backward(const float *in, int64_t ld_in, float *d_in, int64_t ld_din,
                                         const float *d_out, int64_t ld_dout, const float *save_mean,
                                         const float *save_invstd, const float *weight, const float *bias,
                                         float *d_weight, float *d_bias, float leakiness, int train, int64_t n_active,
                                         int planes, void *workspace, size_t workspace_bytes, void *stream);

/* ---- whole-network executor ------------------------------------------------------------------------------------
 * Replaces the Python-level module loop of scn.Sequential.forward (mopa/models/scn_unet.py:32-34 -> [UPSTREAM]
 * sequential.py) for a compiled InputLayer ... OutputLayer chain: one call enqueues every kernel of the forward (or
 * backward) pass. ops: n_ops rows of 12 int32 {type (1 subm, 2 conv, 3 deconv, 4 batchnorm), in_buf, out_buf, 0,
 * level_in, level_out, n_in, n_out, param_index, leakiness, eps, momentum (float bits)}; bufs: n_bufs rows of 4 int32
 * {level, channels, parent_buf (-1 = owns storage; else a column slice of the joined parent), column offset}.
 * params / param_grads: arrays of DEVICE pointers indexed by param_index: conv -> {weight}; batchnorm -> {weight, bias,
 * running_mean, running_var}. Arenas are caller-allocated device scratch of the sizes _prepare reports. */
typedef struct mopa_scn_program mopa_scn_program;
mopa_scn_program *mopa_scn_Program_new(const int32_t *ops, int n_ops, const int32_t *bufs, int n_bufs, int in_planes,
                                       int in_buf, int out_buf, int64_t spatial_size, int n_levels, int device);
void mopa_scn_Program_delete(mopa_scn_program *p);
/* voxelise `coords` (as InputLayer_setLocations) and build every grid / table the program needs on a dedicated
 * high-priority stream; `stream` is made to wait for it. n_active_out[n_levels]; sizes_out = {activation arena bytes,
 * gradient arena bytes, scratch bytes}. coords_on_device: 0 host pointer; 1 device pointer whose contents may still be
 * in flight on `stream` (the geometry stream waits for everything queued there); 2 device pointer whose contents are
 * complete (no wait: the geometry of this forward overlaps whatever `stream` is still running, e.g. the previous
 * step's backward pass). */
int mopa_scn_Program_prepare(mopa_scn_program *p, mopa_scn_metadata *m, const int64_t *coords, int64_t n, int ncols,
                             int coords_on_device, int precision, void *stream, int64_t *n_active_out,
                             uint64_t *sizes_out);
int mopa_scn_Program_forward(mopa_scn_program *p, mopa_scn_metadata *m, const float *feats, int64_t ld_feats,
                             const void *const *params, int train, int precision, void *act_arena, void *scratch,
                             float *out, int64_t ld_out, void *stream);
int mopa_scn_Program_backward(mopa_scn_program *p, mopa_scn_metadata *m, const void *const *params,
                              void *const *param_grads, int train, int precision, const void *act_arena,
                              void *grad_arena, void *scratch, const float *d_out, int64_t ld_dout, float *d_feats,
                              int64_t ld_dfeats, void *stream);

/* ---- instrumentation: number of kernels this library has launched in this process (bench.py `gpu_launches`) --- */
int64_t mopa_scn_kernelLaunchCount(void);

/* Per-kernel timing for bench.py's roofline leg (no reference counterpart). While enabled, every conv / BatchNorm /
 * IO-layer op brackets its kernels with CUDA events ON THE LAUNCHING STREAM and records its shape. tag = 10 * class + op:
 * class 1 conv forward, 2 conv d_input, 3 conv d_weight, 4 BatchNorm forward, 5 BatchNorm backward, 6 IO layers;
 * op 1 submanifold, 2 convolution, 3 deconvolution, 0 n/a. `rules` is the exact rule count of the op's rulebook.
 * _enable(on) clears the log; _read synchronises the device. Off by default; costs nothing when off. */
typedef struct mopa_scn_profile_record {
    int32_t tag, volume, c_in, c_out;
    int64_t rows_out, rows_in, rules;
    float ms;
    int32_t reserved;
} mopa_scn_profile_record;
int mopa_scn_Profile_enable(int on);
int64_t mopa_scn_Profile_count(void);
int mopa_scn_Profile_read(mopa_scn_profile_record *records, int64_t max_records);

#ifdef __cplusplus
}
#endif
#endif /* MOPA_SCN_H_ */
