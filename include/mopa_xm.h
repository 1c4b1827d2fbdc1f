/* mopa_xm.h -- C ABI of the cross-modal / data-side operators next to the UNetSCN hot path (SURVEY.md section 8(f), rows
 * N2-N4), exported by the same libmopa_scn.so. Each entry point names the reference code it replaces. Conventions as
 * in mopa_scn.h: plain C, DEVICE pointers unless a parameter says HOST, explicit CUDA stream, 0 on success and
 * mopa_scn_last_error() otherwise; the callee never allocates result tensors.
 */
#ifndef MOPA_XM_H_
#define MOPA_XM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Invalid indices (image index outside the feature map, mask id >= max_ids) are reported asynchronously, like torch's CUDA
 * indexing: the kernel records them without a stream synchronisation, the NEXT mopa_xm_* call (or this one, which
 * synchronises `stream` first) returns the error. */
int mopa_xm_checkAsyncError(void *stream);

/* ---- N2: 2D -> 3D feature lifting + segmentation heads ------------------------------------------------------------
 * replaces Net2DSeg.forward's per-sample Python loop and its two nn.Linear calls
 *   /root/reference/mopa/models/xmuda_arch.py:62-65   img_feats[i] = x.permute(0,2,3,1)[i][idx[:,0], idx[:,1]]; cat
 *   /root/reference/mopa/models/xmuda_arch.py:73-77   seg_logit = linear(img_feats), seg_logit2 = linear2(img_feats)
 * x: (B, C, H, W) fp32, contiguous NCHW (what the 2D network emits). img_indices: (N, 2) int64 [row, col] of all samples
 * concatenated; sample_offsets: (B + 1) int64 HOST prefix (point n belongs to sample b iff offsets[b] <= n < offsets[b+1]).
 * feats: (N, C) out. w1/b1 (classes, C)/(classes): logit = feats w1^T + b1 -> (N, classes); w2/b2/logit2 may be NULL
 * (single head). Negative indices wrap like torch advanced indexing; an index outside the image is an asynchronous error. */
int mopa_xm_PixelGatherHeads_updateOutput(const float *x, int batch, int channels, int height, int width,
                                          const int64_t *img_indices, const int64_t *sample_offsets_host, int64_t n,
                                          const float *w1, const float *b1, const float *w2, const float *b2, int classes,
                                          float *feats, float *logit, float *logit2, void *stream);
/* backward of the above. d_feats / d_logit / d_logit2 may be NULL (no gradient flows through that output).
 * d_x (B, C, H, W) must be ZEROED by the caller and receives the scattered gradient (points sharing a pixel add up;
 * fp32 atomics like torch's index backward). d_w* (classes, C), d_b* (classes) are OVERWRITTEN (may be NULL).
 * workspace: >= mopa_xm_pixelGatherWorkspaceBytes(channels, classes) bytes. */
size_t mopa_xm_pixelGatherWorkspaceBytes(int channels, int classes);
int mopa_xm_PixelGatherHeads_backward(const float *feats, const int64_t *img_indices, const int64_t *sample_offsets_host,
                                      int batch, int channels, int height, int width, int64_t n, const float *w1,
                                      const float *w2, int classes, const float *d_feats, const float *d_logit,
                                      const float *d_logit2, float *d_x, float *d_w1, float *d_b1, float *d_w2, float *d_b2,
                                      void *workspace, size_t workspace_bytes, void *stream);

/* ---- N2: cross-modal KL loss ----------------------------------------------------------------------------------------
 * replaces /root/reference/mopa/train/train_xmuda_mopa.py:389-398, 440-445 (and train_xmuda.py:248-257, 296-303):
 *   F.kl_div(F.log_softmax(student, 1), F.softmax(teacher.detach(), 1), reduction='none').sum(1).mean()
 * student, teacher: (N, classes) fp32 logits. loss_out: one float (device). d_student (N, classes), optional: the gradient
 * of the loss w.r.t. `student` times *grad_scale_host (computed in the same pass; pass NULL to skip). */
int mopa_xm_KLDivLoss_updateOutput(const float *student, const float *teacher, int64_t n, int classes, float *loss_out,
                                   float *d_student, float grad_scale, void *stream);

/* ---- N4: SAM mask-consistency loss ------------------------------------------------------------------------------------
 * replaces mask_cons_loss, /root/reference/mopa/common/utils/loss.py:241-283 (called at train_xmuda_mopa.py:472-480 with
 * softmaxed logits laid out (B, H, W, C)): a Python loop over masks.unique() with boolean indexing becomes one segmented
 * reduction. probs: (B, H, W, C) fp32; masks: (B, H, W) int32, ids < 0 are ignored, ids must be < max_ids.
 * entropy_norm: the divisor of the entropy term, log2(num_classes) as the reference computes it (it reads num_classes
 * from all_logits.shape[1]); min_entropy = 0 drops the term. loss_out: one float. stats: device scratch of at least
 * mopa_xm_maskConsStatsBytes(batch, max_ids, classes) bytes, kept for the backward call. */
size_t mopa_xm_maskConsStatsBytes(int batch, int max_ids, int classes);
int mopa_xm_MaskConsLoss_updateOutput(const float *probs, const int32_t *masks, int batch, int64_t pixels, int classes,
                                      int max_ids, int min_entropy, float entropy_norm, void *stats, float *loss_out,
                                      void *stream);
/* d_probs (B, H, W, C) = grad_scale * d loss / d probs, OVERWRITTEN (zero at ignored pixels) */
int mopa_xm_MaskConsLoss_backward(const float *probs, const int32_t *masks, int batch, int64_t pixels, int classes,
                                  int max_ids, int min_entropy, float entropy_norm, const void *stats, float grad_scale,
                                  float *d_probs, void *stream);

/* ---- N3: VGI post-processing -------------------------------------------------------------------------------------------
 * replaces the numpy / torch round trips of post_process, /root/reference/mopa/data/mixmatch_ss.py:458-559, for one scan:
 *   range_projection(..., obj_mask) occlusion test   /root/reference/mopa/data/utils/augmentation_3d.py:161-280, 81-111
 *   augment_and_scale_3d                             /root/reference/mopa/data/utils/augmentation_3d.py:6-60
 *   receptive-field filter + int64 coords            mixmatch_ss.py:531-538
 * points: (n, 3) fp64 xyz (scene points followed / interleaved with inserted object points), obj_mask: (n) uint8.
 * Step 1 (occlusion): in every range-image pixel (proj_H x proj_W, fov_up / fov_down in rad) that holds an object point only
 * the nearest point survives (ties: lowest index); keep_out (n) uint8. use_proj = 0 keeps everything.
 * Step 2: kept points p -> round((p . rot) * scale) - min + offset, offset = clip(full_scale - max - 0.001, 0) * rand3
 * (rot: 9 fp64 row-major or NULL, rand3: 3 fp64 in [0,1) or NULL for transl = False; both drawn by the caller from numpy's
 * RNG exactly as the reference does), truncated to int64; rows outside [0, full_scale) are dropped.
 * coords_out: (n, 4) int64 capacity, [x, y, z, batch_index]; sel_out: (n) int64 capacity, the ORIGINAL row of every output
 * row (for gathering labels / masks); aug_points_out: (n, 3) fp64 capacity or NULL (rotated points of the output rows).
 * n_out_host: HOST, receives the number of output rows (the call synchronises the stream once).
 * workspace: >= mopa_xm_vgiWorkspaceBytes(n, proj_H, proj_W) bytes. */
size_t mopa_xm_vgiWorkspaceBytes(int64_t n, int proj_h, int proj_w);
int mopa_xm_VgiPostProcess(const double *points, const uint8_t *obj_mask, int64_t n, int use_proj, double fov_up,
                           double fov_down, int proj_w, int proj_h, const double *rot_host, const double *rand3_host,
                           double scale, int64_t full_scale, int batch_index, uint8_t *keep_out, int64_t *coords_out,
                           int64_t *sel_out, double *aug_points_out, int64_t *n_out_host, void *workspace,
                           size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MOPA_XM_H_ */
