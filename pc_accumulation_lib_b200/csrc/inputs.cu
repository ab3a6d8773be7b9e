// inputs.cu — the dataloader's input side: what produces the
// (N,7) record rows and pc_cam_idx that pcacc_integrate_records consumes.
//
//   k_assign_boxes     box loop of inst_centric_get_sweeps, datasets/nuscenes_utils.py:412-470,
//                      with find_points_in_box :317-329 and apply_tf :233-243
//   k_project_cameras  obs_dataloaders/nuscenes_obs_dataloader.py:176-198 with
//                      homo_transform (datasets/nuscenes_utils.py:46-60) and
//                      NuScenesCamera.project_pts3d (:112-136; nuscenes-devkit view_points)
//
// One thread per point; the per-box / per-camera matrices sit in shared memory.  All
// products are the float64 FMA chains numpy's matmul performs for these shapes (measured in
// the build container), divisions and compares are IEEE, so masks and indices are bit-exact.
#include <math.h>
#include <string.h>

#include "common.cuh"

#define INB 256
#define MAX_BOXES_PER_LAUNCH 256

struct BoxDev {
    double M[12];    // rows 0..2 of inv(target_from_box)
    double size[3];  // dxdydz
};

template <typename T>
__global__ void __launch_bounds__(INB)
k_assign_boxes(const T *__restrict__ pts, int64_t n, int64_t stride, const BoxDev *__restrict__ boxes,
               int n_boxes, int box0, double limit, int32_t *__restrict__ out_box,
               int32_t *__restrict__ out_count) {
    __shared__ BoxDev s_box[MAX_BOXES_PER_LAUNCH];
    for (int i = threadIdx.x; i < n_boxes * (int)(sizeof(BoxDev) / 8); i += INB)
        ((double *)s_box)[i] = ((const double *)boxes)[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * INB + threadIdx.x;
    const bool have = i < n;
    double x = 0.0, y = 0.0, z = 0.0;
    if (have) {
        x = (double)pts[i * stride];
        y = (double)pts[i * stride + 1];
        z = (double)pts[i * stride + 2];
    }
    int last = -1;
    for (int b = 0; b < n_boxes; b++) {
        double bx, by, bz;
        affine_chain(s_box[b].M, 4, x, y, z, bx, by, bz);
        // np.all(np.abs(box_points / dxdydz) < 0.5 + tolerance)
        const bool in = have && fabs(__ddiv_rn(bx, s_box[b].size[0])) < limit &&
                        fabs(__ddiv_rn(by, s_box[b].size[1])) < limit &&
                        fabs(__ddiv_rn(bz, s_box[b].size[2])) < limit;
        if (in) last = box0 + b;   // a later box overwrites an earlier one
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(&out_count[box0 + b], __popc(m));
    }
    if (have && last >= 0) out_box[i] = last;
}

__global__ void k_fill_i32(int32_t *__restrict__ p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int pcacc_assign_boxes(pcacc_t h, const void *pts_dev, int pts_f32, int64_t n, int64_t stride,
                                  const double *box_from_target, const double *dxdydz, int n_boxes,
                                  double tolerance, int32_t *out_box_dev, int32_t *out_count_dev,
                                  void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || n_boxes < 0 || stride < 3 || (n && (!pts_dev || !out_box_dev)) ||
        (n_boxes && (!box_from_target || !dxdydz || !out_count_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad assign_boxes arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t blocks = (n + INB - 1) / INB;
    if (n) {
        k_fill_i32<<<(unsigned)blocks, INB, 0, st>>>(out_box_dev, n, -1);
        PCACC_CUDA(h, cudaGetLastError());
    }
    if (n_boxes) PCACC_CUDA(h, cudaMemsetAsync(out_count_dev, 0, (size_t)n_boxes * 4, st));
    const double limit = 0.5 + tolerance;
    for (int b0 = 0; b0 < n_boxes && n; b0 += MAX_BOXES_PER_LAUNCH) {
        const int nb = n_boxes - b0 < MAX_BOXES_PER_LAUNCH ? n_boxes - b0 : MAX_BOXES_PER_LAUNCH;
        BoxDev tmp[MAX_BOXES_PER_LAUNCH];
        for (int b = 0; b < nb; b++) {
            for (int k = 0; k < 12; k++) tmp[b].M[k] = box_from_target[(size_t)(b0 + b) * 16 + k];
            for (int k = 0; k < 3; k++) tmp[b].size[k] = dxdydz[(size_t)(b0 + b) * 3 + k];
        }
        void *d_boxes = nullptr;
        int rc = pcacc_arena_put(h, tmp, (size_t)nb * sizeof(BoxDev), &d_boxes, st);
        if (rc) return rc;
        h->launches[PCACC_K_INTEGRATE]++;
        if (pts_f32)
            k_assign_boxes<float><<<(unsigned)blocks, INB, 0, st>>>(
                (const float *)pts_dev, n, stride, (const BoxDev *)d_boxes, nb, b0, limit, out_box_dev,
                out_count_dev);
        else
            k_assign_boxes<double><<<(unsigned)blocks, INB, 0, st>>>(
                (const double *)pts_dev, n, stride, (const BoxDev *)d_boxes, nb, b0, limit, out_box_dev,
                out_count_dev);
        PCACC_CUDA(h, cudaGetLastError());
    }
    return PCACC_OK;
}

// ---------------------------------------------------------------------------
struct CamDev {
    double cam_from_glob[12];
    double viewpad[12];   // rows 0..2 of the 4x4 identity with cam_K in its top-left corner
    double w1, h1;        // img_wh - 1
};
struct CamSet {
    double glob_from_ego[12];
    CamDev cam[PCACC_MAX_CAMS];
    int n;
};

__global__ void __launch_bounds__(INB)
k_project_cameras(const double *__restrict__ pc_ego, int64_t n, int64_t stride, const CamSet *__restrict__ cs,
                  double depth_thres, double *__restrict__ out_uv, long long *__restrict__ out_cam) {
    __shared__ CamSet s;
    for (int i = threadIdx.x; i < (int)(sizeof(CamSet) / 8); i += INB) ((double *)&s)[i] = ((const double *)cs)[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * INB + threadIdx.x;
    if (i >= n) return;
    double gx, gy, gz;
    affine_chain(s.glob_from_ego, 4, pc_ego[i * stride], pc_ego[i * stride + 1], pc_ego[i * stride + 2], gx, gy,
                 gz);
    double u_out = 0.0, v_out = 0.0;   // pc_uv starts as zeros, pc_cam_idx as -1
    long long cam = -1;
    for (int j = 0; j < s.n; j++) {
        double cx, cy, cz;
        affine_chain(s.cam[j].cam_from_glob, 4, gx, gy, gz, cx, cy, cz);
        if (!(cz > depth_thres)) continue;   // out stays -10: never inside the image
        double kx, ky, kz;
        affine_chain(s.cam[j].viewpad, 4, cx, cy, cz, kx, ky, kz);
        const double u = __ddiv_rn(kx, kz), v = __ddiv_rn(ky, kz);
        if (u > 1.0 && u < s.cam[j].w1 && v > 1.0 && v < s.cam[j].h1) {
            u_out = u;
            v_out = v;
            cam = j;   // a later camera overwrites an earlier one
        }
    }
    out_uv[2 * i] = u_out;
    out_uv[2 * i + 1] = v_out;
    out_cam[i] = cam;
}

extern "C" int pcacc_project_cameras(pcacc_t h, const double *pc_ego_dev, int64_t n, int64_t stride,
                                     const double *glob_from_ego, const double *cam_from_glob,
                                     const double *cam_K, const double *img_wh, int n_cams,
                                     double depth_thres, double *out_uv_dev, int64_t *out_cam_idx_dev,
                                     void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || stride < 3 || n_cams < 0 || n_cams > PCACC_MAX_CAMS || !glob_from_ego ||
        (n_cams && (!cam_from_glob || !cam_K || !img_wh)) || (n && (!pc_ego_dev || !out_uv_dev || !out_cam_idx_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad project_cameras arguments (at most %d cameras)", PCACC_MAX_CAMS);
    if (n == 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    CamSet cs;
    memset(&cs, 0, sizeof(cs));
    for (int k = 0; k < 12; k++) cs.glob_from_ego[k] = glob_from_ego[k];
    cs.n = n_cams;
    for (int j = 0; j < n_cams; j++) {
        for (int k = 0; k < 12; k++) cs.cam[j].cam_from_glob[k] = cam_from_glob[(size_t)j * 16 + k];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) cs.cam[j].viewpad[4 * r + c] = cam_K[(size_t)j * 9 + 3 * r + c];
            cs.cam[j].viewpad[4 * r + 3] = 0.0;
        }
        cs.cam[j].w1 = img_wh[2 * j] - 1.0;
        cs.cam[j].h1 = img_wh[2 * j + 1] - 1.0;
    }
    void *d_cs = nullptr;
    int rc = pcacc_arena_put(h, &cs, sizeof(cs), &d_cs, st);
    if (rc) return rc;
    h->launches[PCACC_K_INTEGRATE]++;
    k_project_cameras<<<(unsigned)((n + INB - 1) / INB), INB, 0, st>>>(
        pc_ego_dev, n, stride, (const CamSet *)d_cs, depth_thres, out_uv_dev, (long long *)out_cam_idx_dev);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

// ---------------------------------------------------------------------------
// pts_feat_from_img, datasets/nuscenes_utils.py:181-214 — both branches.
// One thread per (point, channel).  'nearest': img[rint(v), rint(u)] (half to even).
// 'bilinear': the reference's weights in its own operation order, every product and sum
// rounded separately (numpy evaluates the expression element-wise without contraction):
//   total = (uc - uf) * (vc - vf);  w_ff = (uc - u) * (vc - v) / total;  w_cc = (u - uf) * (v - vf) / total;
//   w_fc = (u - uf) * (vc - v) / total;  w_cf = 1 - (w_ff + w_cc + w_fc);
//   feat = ((w_ff * I[vf,uf] + w_cc * I[vc,uc]) + w_cf * I[vc,uf]) + w_fc * I[vf,uc]
// (an integral u or v makes total 0 and the result NaN / inf, as in the reference).  The
// reference multiplies (N,) weights with (N,C) features, which numpy only broadcasts for a 2-D
// image; here every channel gets the per-point weights, which is that result for C = 1.
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double img_at(const T *img, int64_t pix, int C, int c) {
    return (double)img[pix * C + c];
}

template <typename T>
__global__ void __launch_bounds__(INB)
k_pts_feat(const double *__restrict__ uv, int64_t n, const T *__restrict__ img, int H, int W, int C,
           int bilinear, double *__restrict__ out, uint32_t *__restrict__ flags) {
    const int64_t t = (int64_t)blockIdx.x * INB + threadIdx.x;
    if (t >= n * C) return;
    const int64_t i = t / C;
    const int c = (int)(t - i * C);
    const double u = uv[2 * i], v = uv[2 * i + 1];
    // assert np.all((pts_uv > 1) & (pts_uv < img_wh - 1))
    const bool inside = (u > 1.0) && (u < (double)W - 1.0) && (v > 1.0) && (v < (double)H - 1.0);
    if (!inside) {
        if (c == 0) atomicOr(flags, PCACC_FLAG_UV_OUT_OF_IMAGE);
        out[t] = 0.0;
        return;
    }
    if (!bilinear) {
        const int64_t pix = (int64_t)rint_even(v) * W + (int64_t)rint_even(u);
        out[t] = img_at(img, pix, C, c);
        return;
    }
    const double uf = floor(u), uc = ceil(u), vf = floor(v), vc = ceil(v);
    const double total = __dmul_rn(__dsub_rn(uc, uf), __dsub_rn(vc, vf));
    const double w_ff = __ddiv_rn(__dmul_rn(__dsub_rn(uc, u), __dsub_rn(vc, v)), total);
    const double w_cc = __ddiv_rn(__dmul_rn(__dsub_rn(u, uf), __dsub_rn(v, vf)), total);
    const double w_fc = __ddiv_rn(__dmul_rn(__dsub_rn(u, uf), __dsub_rn(vc, v)), total);
    const double w_cf = __dsub_rn(1.0, __dadd_rn(__dadd_rn(w_ff, w_cc), w_fc));
    const int64_t iuf = (int64_t)uf, iuc = (int64_t)uc, ivf = (int64_t)vf, ivc = (int64_t)vc;
    const double a = __dmul_rn(w_ff, img_at(img, ivf * W + iuf, C, c));
    const double b = __dmul_rn(w_cc, img_at(img, ivc * W + iuc, C, c));
    const double d = __dmul_rn(w_cf, img_at(img, ivc * W + iuf, C, c));
    const double e = __dmul_rn(w_fc, img_at(img, ivf * W + iuc, C, c));
    out[t] = __dadd_rn(__dadd_rn(__dadd_rn(a, b), d), e);
}

extern "C" int pcacc_pts_feat_from_img(pcacc_t h, const double *uv_dev, int64_t n, const void *img_dev,
                                       int img_dtype, int img_h, int img_w, int channels, int bilinear,
                                       double *out_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || img_h <= 0 || img_w <= 0 || channels <= 0 || (n > 0 && (!uv_dev || !img_dev || !out_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad pts_feat_from_img arguments");
    if (n == 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t blocks = (n * channels + INB - 1) / INB;
    h->launches[PCACC_K_INTEGRATE]++;
#define LAUNCH_PF(T)                                                                                    \
    k_pts_feat<T><<<(unsigned)blocks, INB, 0, st>>>(uv_dev, n, (const T *)img_dev, img_h, img_w, channels, \
                                                    bilinear, out_dev, h->d_flags)
    switch (img_dtype) {
        case PCACC_SEM_U8: LAUNCH_PF(uint8_t); break;
        case PCACC_SEM_I16: LAUNCH_PF(int16_t); break;
        case PCACC_SEM_I32: LAUNCH_PF(int32_t); break;
        case PCACC_SEM_I64: LAUNCH_PF(long long); break;
        case PCACC_IMG_F32: LAUNCH_PF(float); break;
        case PCACC_IMG_F64: LAUNCH_PF(double); break;
        default: return pcacc_fail(h, PCACC_ERR_ARG, "unsupported image dtype %d", img_dtype);
    }
#undef LAUNCH_PF
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

// ---------------------------------------------------------------------------
// static_obj_partitioning_by_elev, bev_generator/sem_bev.py:556-591: per-cell minimum z of
// the (grid-coordinate) cloud, then every point more than elev_thresh above its cell's
// minimum gets column 8 set to 1.  pc rows are the (M,10) float64 records whose columns 0, 1
// hold grid coordinates (output of pos2grid); `.astype(int)` truncates toward zero and numpy
// wraps negative indices once, which is mirrored; anything outside [-P, P) is the reference's
// IndexError and raises PCACC_FLAG_ATTR_RANGE here.  The minimum is order-independent except
// for NaN z (the reference keeps a NaN that arrives first): NaN z raises the same flag.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool grid_index(double g, int P, int &idx) {
    if (!(g > -2147483648.0 && g < 2147483647.0)) return false;   // NaN / inf / beyond int
    int i = (int)g;   // truncation toward zero, as ndarray.astype(int)
    if (i < -P || i >= P) return false;
    idx = i < 0 ? i + P : i;
    return true;
}

__global__ void __launch_bounds__(INB)
k_elev_min(const double *__restrict__ pc, int64_t n, int cols, int P, unsigned long long *__restrict__ zmin,
           uint32_t *__restrict__ flags) {
    const int64_t t = (int64_t)blockIdx.x * INB + threadIdx.x;
    if (t >= n) return;
    const double *r = pc + t * cols;
    int i, j;
    const double z = r[2];
    // j_rev = P - 1 - j is what is indexed (and wrapped) in the reference
    double jr = (double)(P - 1) - trunc(r[1]);
    if (!grid_index(r[0], P, i) || !(r[1] > -2147483648.0 && r[1] < 2147483647.0) || !grid_index(jr, P, j) ||
        z != z) {
        atomicOr(flags, PCACC_FLAG_ATTR_RANGE);
        return;
    }
    atomicMin(&zmin[(int64_t)j * P + i], ord_encode(z));
}

__global__ void __launch_bounds__(INB)
k_elev_flag(double *__restrict__ pc, int64_t n, int P, double elev_thresh,
            const unsigned long long *__restrict__ zmin) {
    const int64_t t = (int64_t)blockIdx.x * INB + threadIdx.x;
    if (t >= n) return;
    double *r = pc + t * 10;
    int i, j;
    double jr = (double)(P - 1) - trunc(r[1]);
    if (!grid_index(r[0], P, i) || !(r[1] > -2147483648.0 && r[1] < 2147483647.0) || !grid_index(jr, P, j) ||
        r[2] != r[2])
        return;
    const double m = ord_decode(zmin[(int64_t)j * P + i]);
    if (r[2] > __dadd_rn(m, elev_thresh)) r[8] = 1.0;
}

__global__ void __launch_bounds__(INB)
k_elev_export(const unsigned long long *__restrict__ zmin, int64_t cells, double *__restrict__ elevmap,
              uint8_t *__restrict__ obs) {
    const int64_t t = (int64_t)blockIdx.x * INB + threadIdx.x;
    if (t >= cells) return;
    const unsigned long long u = zmin[t];
    const bool seen = u != AABB_EMPTY_MIN;
    elevmap[t] = seen ? ord_decode(u) : 0.0;
    obs[t] = seen ? 1 : 0;
}

extern "C" int pcacc_static_obj_partitioning(pcacc_t h, double *pc_dev, int64_t n, int P, double elev_thresh,
                                             double *elevmap_dev, uint8_t *obs_mask_dev,
                                             unsigned long long *scratch_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || P <= 0 || !elevmap_dev || !obs_mask_dev || !scratch_dev || (n > 0 && !pc_dev))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad static_obj_partitioning arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t cells = (int64_t)P * P;
    PCACC_CUDA(h, cudaMemsetAsync(scratch_dev, 0xff, (size_t)cells * 8, st));   // AABB_EMPTY_MIN everywhere
    h->launches[PCACC_K_EXPORT] += 3;
    if (n > 0) {
        const int64_t nb = (n + INB - 1) / INB;
        k_elev_min<<<(unsigned)nb, INB, 0, st>>>(pc_dev, n, 10, P, scratch_dev, h->d_flags);
        k_elev_flag<<<(unsigned)nb, INB, 0, st>>>(pc_dev, n, P, elev_thresh, scratch_dev);
    }
    k_elev_export<<<(unsigned)((cells + INB - 1) / INB), INB, 0, st>>>(scratch_dev, cells, elevmap_dev,
                                                                     obs_mask_dev);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

// get_elevation_map, bev_generator/sem_bev.py:535-553: the per-cell minimum z alone, for rows of
// any width (x, y, z in columns 0..2).
extern "C" int pcacc_elevation_map(pcacc_t h, const double *pc_dev, int64_t n, int cols, int P,
                                   double *elevmap_dev, uint8_t *obs_mask_dev,
                                   unsigned long long *scratch_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || cols < 3 || P <= 0 || !elevmap_dev || !obs_mask_dev || !scratch_dev || (n > 0 && !pc_dev))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad elevation_map arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t cells = (int64_t)P * P;
    PCACC_CUDA(h, cudaMemsetAsync(scratch_dev, 0xff, (size_t)cells * 8, st));
    h->launches[PCACC_K_EXPORT] += 2;
    if (n > 0)
        k_elev_min<<<(unsigned)((n + INB - 1) / INB), INB, 0, st>>>(pc_dev, n, cols, P, scratch_dev,
                                                                     h->d_flags);
    k_elev_export<<<(unsigned)((cells + INB - 1) / INB), INB, 0, st>>>(scratch_dev, cells, elevmap_dev,
                                                                     obs_mask_dev);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}
