// VGI post-processing on the GPU, sm_100a. C ABI: include/mopa_xm.h::mopa_xm_VgiPostProcess.
// Replaces, per mix-matched scan, the numpy / torch round trips of post_process (/root/reference/mopa/data/mixmatch_ss.py:
// 458-559): range_projection's occlusion test (/root/reference/mopa/data/utils/augmentation_3d.py:161-280 with
// occulusion_detector :81-111), augment_and_scale_3d (:6-60) and the receptive-field filter (mixmatch_ss.py:531-538).
// Integer / index work is bit-exact; the float64 arithmetic follows numpy's operation order without FMA contraction
// (explicit __dmul_rn / __dadd_rn), so the rounded voxel coordinates agree with the reference's.
//
// Occlusion rule restated: a range-image pixel that holds at least one inserted-object point keeps only its nearest point
// (object or scene; ties -> lowest row index, the stable lexsort order of occulusion_detector); all other pixels keep
// every point. Two atomicMin passes per flagged pixel (depth bits, then row index) instead of a sort.
#include <math.h>

#include "geometry.cuh"
#include "mopa_xm.h"

namespace mopa {

int exclusive_scan(const int32_t *in, int32_t *out, int64_t n, int32_t *bsum, int32_t *total, cudaStream_t s);  // geometry.cu

constexpr double kPi = 3.141592653589793;  // numpy.pi

// projection of one point: (pixel index, depth); augmentation_3d.py:198-232, float64 throughout
__device__ __forceinline__ void vgi_project(const double *__restrict__ p, double fov_down_abs, double fov, int W, int H, int &pix,
                                            double &depth) {
    const double x = p[0], y = p[1], z = p[2];
    depth = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));  // np.linalg.norm(points, 2, axis=1)
    const double yaw = -atan2(y, x);
    const double pitch = asin(z / depth);
    double px = __dmul_rn(__dmul_rn(0.5, __dadd_rn(yaw / kPi, 1.0)), (double)W);                  // 0.5 * (yaw / pi + 1.0) * W
    double py = __dmul_rn(__dadd_rn(1.0, -(__dadd_rn(pitch, fov_down_abs) / fov)), (double)H);    // (1.0 - (pitch + |fov_down|) / fov) * H
    px = floor(px);
    px = fmin((double)(W - 1), px);
    px = fmax(0.0, px);
    py = floor(py);
    py = fmin((double)(H - 1), py);
    py = fmax(0.0, py);
    pix = (int)py * W + (int)px;
}

__global__ void __launch_bounds__(256) k_vgi_flag(const double *__restrict__ pts, const uint8_t *__restrict__ obj, int64_t n,
                                                 double fov_down_abs, double fov, int W, int H, int32_t *__restrict__ pix_of,
                                                 unsigned long long *__restrict__ depth_of, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int pix;
    double d;
    vgi_project(pts + 3 * i, fov_down_abs, fov, W, H, pix, d);
    pix_of[i] = pix;
    depth_of[i] = (unsigned long long)__double_as_longlong(d);  // depth >= 0: the bit pattern orders like the value
    if (obj[i]) flag[pix] = 1;
}
__global__ void __launch_bounds__(256) k_vgi_min_depth(const int32_t *__restrict__ pix_of,
                                                      const unsigned long long *__restrict__ depth_of, int64_t n,
                                                      const int32_t *__restrict__ flag, unsigned long long *__restrict__ best) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pix = pix_of[i];
    if (flag[pix]) atomicMin(best + pix, depth_of[i]);
}
__global__ void __launch_bounds__(256) k_vgi_min_index(const int32_t *__restrict__ pix_of,
                                                      const unsigned long long *__restrict__ depth_of, int64_t n,
                                                      const int32_t *__restrict__ flag, const unsigned long long *__restrict__ best,
                                                      int32_t *__restrict__ first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pix = pix_of[i];
    if (flag[pix] && depth_of[i] == best[pix]) atomicMin(first + pix, (int32_t)i);
}
__global__ void __launch_bounds__(256) k_vgi_keep(const int32_t *__restrict__ pix_of, int64_t n, const int32_t *__restrict__ flag,
                                                 const int32_t *__restrict__ first, int use_proj, uint8_t *__restrict__ keep,
                                                 int32_t *__restrict__ keep32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = 1;
    if (use_proj) {
        const int pix = pix_of[i];
        k = !flag[pix] || first[pix] == (int32_t)i;
    }
    keep[i] = (uint8_t)k;
    keep32[i] = k;
}

// kept rows: rotated point q = p . rot (numpy's (N,3) x (3,3) dot: q_j = (p0 R0j + p1 R1j) + p2 R2j), r = rint(q * scale);
// per-axis min / max of r over the kept rows (exact: r is integral and small)
__global__ void __launch_bounds__(256) k_vgi_transform(const double *__restrict__ pts, const int32_t *__restrict__ keep32,
                                                      const int32_t *__restrict__ rank, int64_t n, const double *__restrict__ rot,
                                                      double scale, double *__restrict__ q_out, double *__restrict__ r_out,
                                                      int32_t *__restrict__ row_out, long long *__restrict__ mm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep32[i]) return;
    const int64_t o = rank[i];
    const double p0 = pts[3 * i], p1 = pts[3 * i + 1], p2 = pts[3 * i + 2];
    row_out[o] = (int32_t)i;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double q = j == 0 ? p0 : (j == 1 ? p1 : p2);
        if (rot) q = __dadd_rn(__dadd_rn(__dmul_rn(p0, rot[j]), __dmul_rn(p1, rot[3 + j])), __dmul_rn(p2, rot[6 + j]));
        const double r = rint(__dmul_rn(q, scale));  // np.round: half to even
        q_out[3 * o + j] = q;
        r_out[3 * o + j] = r;
        atomicMin(mm + j, (long long)r);
        atomicMax(mm + 3 + j, (long long)r);
    }
}
// coords = (r - min) + offset; valid = all axes in [0, full_scale) on the float value; int64 cast truncates
__global__ void __launch_bounds__(256) k_vgi_valid(const double *__restrict__ r_in, const int32_t *__restrict__ m_dev, int64_t n,
                                                  const long long *__restrict__ mm, const double *__restrict__ rand3,
                                                  int has_transl, double full_scale, double *__restrict__ c_out,
                                                  int32_t *__restrict__ valid) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    if (o >= *m_dev) { valid[o] = 0; return; }  // slots beyond the kept rows (their count is only known on the device)
    int ok = 1;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double mn = (double)mm[j], mx = (double)mm[3 + j];
        double c = __dadd_rn(r_in[3 * o + j], -mn);  // coords -= coords.min(0)
        if (has_transl) {  // offset = clip(full_scale - coords.max(0) - 0.001, 0, None) * rand(3)
            double off = __dadd_rn(__dadd_rn(full_scale, -__dadd_rn(mx, -mn)), -0.001);
            off = __dmul_rn(fmax(off, 0.0), rand3[j]);
            c = __dadd_rn(c, off);
        }
        c_out[3 * o + j] = c;
        ok &= (c >= 0.0) && (c < full_scale);
    }
    valid[o] = ok;
}
__global__ void __launch_bounds__(256) k_vgi_emit(const double *__restrict__ c_in, const double *__restrict__ q_in,
                                                 const int32_t *__restrict__ row_in, const int32_t *__restrict__ valid,
                                                 const int32_t *__restrict__ rank, int64_t n, int batch_index,
                                                 int64_t *__restrict__ coords, int64_t *__restrict__ sel, double *__restrict__ aug) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n || !valid[o]) return;
    const int64_t w = rank[o];
    coords[4 * w] = (int64_t)c_in[3 * o];
    coords[4 * w + 1] = (int64_t)c_in[3 * o + 1];
    coords[4 * w + 2] = (int64_t)c_in[3 * o + 2];
    coords[4 * w + 3] = batch_index;
    sel[w] = row_in[o];
    if (aug) { aug[3 * w] = q_in[3 * o]; aug[3 * w + 1] = q_in[3 * o + 1]; aug[3 * w + 2] = q_in[3 * o + 2]; }
}

struct VgiWs {  // byte offsets inside the workspace
    size_t pix, depth, flag, best, first, keep32, rank, bsum, q, r, c, row, valid, rank2, mm, rot, cnt, total;
};
static VgiWs vgi_layout(int64_t n, int64_t pixels) {
    VgiWs L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t nn = (size_t)(n > 0 ? n : 1);
    L.depth = take(nn * 8); L.best = take((size_t)pixels * 8);
    L.q = take(nn * 24); L.r = take(nn * 24); L.c = take(nn * 24);
    L.mm = take(64); L.rot = take(128);
    L.pix = take(nn * 4); L.flag = take((size_t)pixels * 4); L.first = take((size_t)pixels * 4);
    L.keep32 = take(nn * 4); L.rank = take(nn * 4); L.bsum = take((nn / 1024 + 2) * 4);
    L.row = take(nn * 4); L.valid = take(nn * 4); L.rank2 = take(nn * 4); L.cnt = take(64);
    L.total = off;
    return L;
}

}  // namespace mopa

using namespace mopa;

extern "C" {

size_t mopa_xm_vgiWorkspaceBytes(int64_t n, int proj_h, int proj_w) { return vgi_layout(n, (int64_t)proj_h * proj_w).total + 256; }

int mopa_xm_VgiPostProcess(const double *points, const uint8_t *obj_mask, int64_t n, int use_proj, double fov_up,
                           double fov_down, int proj_w, int proj_h, const double *rot_host, const double *rand3_host,
                           double scale, int64_t full_scale, int batch_index, uint8_t *keep_out, int64_t *coords_out,
                           int64_t *sel_out, double *aug_points_out, int64_t *n_out_host, void *workspace,
                           size_t workspace_bytes, void *stream) {
    MOPA_CHECK(points && obj_mask && keep_out && coords_out && sel_out && n_out_host && workspace, "VgiPostProcess: null argument");
    MOPA_CHECK(n >= 0 && n < ((int64_t)1 << 31) && proj_w >= 1 && proj_h >= 1 && (int64_t)proj_w * proj_h < ((int64_t)1 << 30),
               "VgiPostProcess: sizes out of range");
    *n_out_host = 0;
    if (n == 0) return 0;
    const int64_t pixels = (int64_t)proj_w * proj_h;
    const VgiWs L = vgi_layout(n, pixels);
    MOPA_CHECK(workspace_bytes >= L.total, "VgiPostProcess: workspace too small");
    MOPA_CHECK(((uintptr_t)workspace & 255) == 0, "VgiPostProcess: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    auto at = [&](size_t off) { return ws + off; };
    int32_t *pix = (int32_t *)at(L.pix), *flag = (int32_t *)at(L.flag), *first = (int32_t *)at(L.first);
    int32_t *keep32 = (int32_t *)at(L.keep32), *rank = (int32_t *)at(L.rank), *bsum = (int32_t *)at(L.bsum);
    int32_t *row = (int32_t *)at(L.row), *valid = (int32_t *)at(L.valid), *rank2 = (int32_t *)at(L.rank2);
    int32_t *cnt = (int32_t *)at(L.cnt);
    unsigned long long *depth = (unsigned long long *)at(L.depth), *best = (unsigned long long *)at(L.best);
    double *q = (double *)at(L.q), *r = (double *)at(L.r), *c = (double *)at(L.c), *rot = (double *)at(L.rot);
    long long *mm = (long long *)at(L.mm);
    const unsigned g = (unsigned)ceil_div(n, 256);
    if (use_proj) {
        MOPA_CUDA(cudaMemsetAsync(flag, 0, (size_t)pixels * 4, s));
        MOPA_CUDA(cudaMemsetAsync(best, 0xFF, (size_t)pixels * 8, s));
        MOPA_CUDA(cudaMemsetAsync(first, 0x7F, (size_t)pixels * 4, s));
        const double fov = fabs(fov_down) + fabs(fov_up);
        k_vgi_flag<<<g, 256, 0, s>>>(points, obj_mask, n, fabs(fov_down), fov, proj_w, proj_h, pix, depth, flag);
        MOPA_LAUNCHED();
        k_vgi_min_depth<<<g, 256, 0, s>>>(pix, depth, n, flag, best);
        MOPA_LAUNCHED();
        k_vgi_min_index<<<g, 256, 0, s>>>(pix, depth, n, flag, best, first);
        MOPA_LAUNCHED();
    }
    k_vgi_keep<<<g, 256, 0, s>>>(pix, n, flag, first, use_proj, keep_out, keep32);
    MOPA_LAUNCHED();
    MOPA_TRY(exclusive_scan(keep32, rank, n, bsum, cnt, s));
    // rotation matrix / random translation factors: 12 doubles, staged through the workspace
    double host12[12];
    for (int i = 0; i < 9; ++i) host12[i] = rot_host ? rot_host[i] : (i % 4 == 0 ? 1.0 : 0.0);
    for (int i = 0; i < 3; ++i) host12[9 + i] = rand3_host ? rand3_host[i] : 0.0;
    MOPA_CUDA(cudaMemcpyAsync(rot, host12, sizeof(host12), cudaMemcpyHostToDevice, s));
    const long long mm_init[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
    MOPA_CUDA(cudaMemcpyAsync(mm, mm_init, sizeof(mm_init), cudaMemcpyHostToDevice, s));
    k_vgi_transform<<<g, 256, 0, s>>>(points, keep32, rank, n, rot_host ? rot : nullptr, scale, q, r, row, mm);
    MOPA_LAUNCHED();
    // the kept count is only known on the device: the remaining kernels run over n slots and mask by the device count
    k_vgi_valid<<<g, 256, 0, s>>>(r, cnt, n, mm, rot + 9, rand3_host != nullptr, (double)full_scale, c, valid);
    MOPA_LAUNCHED();
    MOPA_TRY(exclusive_scan(valid, rank2, n, bsum, cnt + 1, s));
    k_vgi_emit<<<g, 256, 0, s>>>(c, q, row, valid, rank2, n, batch_index, coords_out, sel_out, aug_points_out);
    MOPA_LAUNCHED();
    int32_t out_host = 0;
    MOPA_CUDA(cudaMemcpyAsync(&out_host, cnt + 1, 4, cudaMemcpyDeviceToHost, s));
    MOPA_CUDA(cudaStreamSynchronize(s));  // the one synchronisation: row count back (and host12 / mm_init consumed)
    *n_out_host = out_host;
    return 0;
}

}  // extern "C"
