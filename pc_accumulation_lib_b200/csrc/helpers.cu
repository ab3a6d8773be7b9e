// Stand-alone forms of the per-function steps that the fused rasteriser (raster.cu) folds
// together: the reference exposes each of them as a public method of BEVGenerator /
// SemBEVGenerator / SemanticPointCloudAccumulator, so each has an entry point of its own
// with the reference's argument meaning.  None of these is on the headline path; they are
// one- or two-launch kernels over (n, cols) float64 row clouds that stay on the device.
#include "common.cuh"

#define HB 256

namespace {

struct Mat34 { double m[12]; };
struct Mat33 { double m[9]; };
#define HELPER_MAX_SEMS 32   /* classes one selection may list (PCACC_MAX_SEMS in pcacc.h) */
struct SemSel { int n; double v[HELPER_MAX_SEMS]; };

// statically indexed and predicated (a loop to sel.n would index the by-value kernel parameter at run
// time and send it to local memory)
__device__ __forceinline__ bool sel_hit(const SemSel &sel, double s) {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < HELPER_MAX_SEMS; k++) hit |= (k < sel.n) && (s == sel.v[k]);
    return hit;
}

struct LookBackH {
    unsigned long long *state;
    uint32_t *ticket;
    uint32_t epoch;
    uint32_t n_tiles;
};

// sem_pc_accum.py:347-366: rows of (3,4) P times [x y z 1]
template <typename T>
__global__ void __launch_bounds__(HB)
k_velo2frame(const T *__restrict__ pts, int64_t n, int stride, Mat34 P, double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * HB + threadIdx.x;
    if (t >= n) return;
    const T *r = pts + t * stride;
    double ox, oy, oz;
    affine_chain(P.m, 4, (double)r[0], (double)r[1], (double)r[2], ox, oy, oz);
    out[t * 3 + 0] = ox;
    out[t * 3 + 1] = oy;
    out[t * 3 + 2] = oz;
}

// bev_generator.py:127-160 (cloud part), :207-256, :737-747.  One row per thread; the kept rows
// keep their input order (decoupled look-back over 256-row tiles).
__global__ void __launch_bounds__(HB)
k_preprocess_pc(const double *__restrict__ pc, int64_t n, int cols, int do_rot, Mat33 R, double dx,
                double dy, int do_crop, double view, int do_hf, double height_filter, int to_grid,
                double grid_view, int P, double *__restrict__ out, int64_t *__restrict__ n_kept,
                LookBackH lb) {
    __shared__ uint32_t s_warp[HB / 32 + 1];
    __shared__ uint32_t s_tile;
    const uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const int64_t t = (int64_t)tile * HB + threadIdx.x;
    bool keep = false;
    double x = 0.0, y = 0.0, z = 0.0;
    if (t < n) {
        const double *r = pc + t * cols;
        x = r[0];
        y = r[1];
        z = r[2];
        if (do_rot) {     // strided 3x3 dgemm: one multiply, two fused multiply-adds over k
            const double rx = __fma_rn(R.m[2], z, __fma_rn(R.m[1], y, __dmul_rn(R.m[0], x)));
            const double ry = __fma_rn(R.m[5], z, __fma_rn(R.m[4], y, __dmul_rn(R.m[3], x)));
            const double rz = __fma_rn(R.m[8], z, __fma_rn(R.m[7], y, __dmul_rn(R.m[6], x)));
            x = __dadd_rn(rx, dx);
            y = __dadd_rn(ry, dy);
            z = rz;
        }
        keep = true;
        if (do_crop) {    // strict on both sides; NaN never passes
            const double hv = __dmul_rn(0.5, view);
            keep = (x > -hv) && (x < hv) && (y > -hv) && (y < hv);
        }
        if (do_hf) keep = keep && (z < height_filter);
        if (to_grid) {    // div, mul, add separately rounded, then floor
            const double dP = (double)P, hP = __dmul_rn(0.5, dP);
            x = floor(__dadd_rn(__dmul_rn(__ddiv_rn(x, grid_view), dP), hP));
            y = floor(__dadd_rn(__dmul_rn(__ddiv_rn(y, grid_view), dP), hP));
        }
    }
    uint32_t tile_end;
    const uint32_t rank = compact_rank<HB>(keep, lb.state, lb.epoch, tile, s_warp, &tile_end);
    if (keep) {
        const double *r = pc + t * cols;
        double *o = out + (int64_t)rank * cols;
        o[0] = x;
        o[1] = y;
        o[2] = z;
        for (int c = 3; c < cols; c++) o[c] = r[c];
    }
    if (tile == lb.n_tiles - 1 && threadIdx.x == 0) *n_kept = (int64_t)tile_end;
}

// np.histogram2d over [0, P] x [0, P] with P bins a side (bev_generator.py:436-453): a value v
// lands in bin floor(v) for 0 <= v < P, the right edge P itself in the last bin, anything else
// (NaN included) nowhere.
__device__ __forceinline__ bool hist_bin(double v, int P, int &b) {
    if (!(v >= 0.0 && v <= (double)P)) return false;
    b = (v == (double)P) ? P - 1 : (int)v;
    return true;
}

// partition_semantic_pc + gen_gridmap_count_map (bev_generator.py:411-453): per cell, the number
// of selected and of other points and the weight sum of the selected ones.  Counts are exact in
// float64 whatever the order; the weight sum is a float64 atomic (order-dependent rounding).
__global__ void __launch_bounds__(HB)
k_cell_stats(const double *__restrict__ pc, int64_t n, int cols, int P, int sem_col, SemSel sel,
             int weight_col, const double *__restrict__ weights, double *__restrict__ cnt_sel,
             double *__restrict__ cnt_rest, double *__restrict__ wsum) {
    const int64_t t = (int64_t)blockIdx.x * HB + threadIdx.x;
    if (t >= n) return;
    const double *r = pc + t * cols;
    bool hit = sel.n < 0;
    if (!hit) {
        hit = sel_hit(sel, r[sem_col]);
    }
    int bi, bj;
    // first histogram axis = column 1 (j), second = column 0 (i); then np.flip(axis=0)
    if (!hist_bin(r[1], P, bj) || !hist_bin(r[0], P, bi)) return;
    const int64_t cell = (int64_t)(P - 1 - bj) * P + bi;
    if (hit) {
        if (cnt_sel) atomicAdd(&cnt_sel[cell], 1.0);
        if (wsum) atomicAdd(&wsum[cell], weights ? weights[t] : r[weight_col]);
    } else if (cnt_rest) {
        atomicAdd(&cnt_rest[cell], 1.0);
    }
}

// finish = 1: dirichlet_dist_expectation of [sel, rest] with a uniform prior (bev_generator.py:
// 455-481); finish = 2: weight sum / (count + 1) (bev_generator.py:393-409)
__global__ void __launch_bounds__(HB)
k_cell_finish(int64_t cells, int finish, double *__restrict__ cnt_sel, double *__restrict__ cnt_rest,
              double *__restrict__ wsum) {
    const int64_t t = (int64_t)blockIdx.x * HB + threadIdx.x;
    if (t >= cells) return;
    if (finish == 1) {
        const double a = __dadd_rn(cnt_sel[t], 1.0), b = __dadd_rn(cnt_rest[t], 1.0);
        const double a0 = __dadd_rn(a, b);
        cnt_sel[t] = __ddiv_rn(a, a0);
        cnt_rest[t] = __ddiv_rn(b, a0);
    } else if (finish == 2) {
        wsum[t] = __ddiv_rn(wsum[t], __dadd_rn(cnt_sel[t], 1.0));
    }
}

// dirichlet_dist_expectation on G stacked maps, in place (bev_generator.py:455-481): scale by
// obs_weight, add the uniform prior, divide by the sum over maps (np.sum over axis 0 adds the
// maps in order)
__global__ void __launch_bounds__(HB)
k_dirichlet(double *__restrict__ maps, int G, int64_t cells, double obs_weight) {
    const int64_t t = (int64_t)blockIdx.x * HB + threadIdx.x;
    if (t >= cells) return;
    double a0 = 0.0;
    for (int g = 0; g < G; g++) {
        const double v = __dadd_rn(__dmul_rn(maps[(int64_t)g * cells + t], obs_weight), 1.0);
        maps[(int64_t)g * cells + t] = v;
        a0 = g == 0 ? v : __dadd_rn(a0, v);
    }
    for (int g = 0; g < G; g++) maps[(int64_t)g * cells + t] = __ddiv_rn(maps[(int64_t)g * cells + t], a0);
}

// sem_bev.py:593-617: int_scaler * sigmoid(int_sep_scaler * (I - int_mid_threshold)), capped at 1;
// sigmoid_only: 1 / (1 + exp(-z))
__global__ void __launch_bounds__(HB)
k_road_marking(const double *__restrict__ in, int64_t n, double int_scaler, double int_sep_scaler,
               double int_mid_threshold, int sigmoid_only, double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * HB + threadIdx.x;
    if (t >= n) return;
    const double v = in[t];
    if (sigmoid_only) {
        out[t] = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-v)));
        return;
    }
    const double z = __dmul_rn(int_sep_scaler, __dsub_rn(v, int_mid_threshold));
    const double s = __dmul_rn(int_scaler, __ddiv_rn(1.0, __dadd_rn(1.0, exp(-z))));
    out[t] = s > 1.0 ? 1.0 : s;
}

// partition_semantic_pc, bev_generator.py:411-432: rows whose column sem_col equals one of the
// listed classes go to out_sel, the others to out_rest, both in input order.  One scan serves both:
// a row's rank among the others is its index minus its rank among the selected.
__global__ void __launch_bounds__(HB)
k_partition_sem(const double *__restrict__ pc, int64_t n, int cols, int sem_col, SemSel sel,
                double *__restrict__ out_sel, double *__restrict__ out_rest, int64_t *__restrict__ n_sel,
                LookBackH lb) {
    __shared__ uint32_t s_warp[HB / 32 + 1];
    __shared__ uint32_t s_tile;
    const uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const int64_t t = (int64_t)tile * HB + threadIdx.x;
    bool hit = false;
    if (t < n) {
        hit = sel_hit(sel, pc[t * cols + sem_col]);
    }
    uint32_t tile_end;
    const uint32_t rank = compact_rank<HB>(hit, lb.state, lb.epoch, tile, s_warp, &tile_end);
    if (t < n) {
        const double *r = pc + t * cols;
        double *o = hit ? out_sel + (int64_t)rank * cols : out_rest + (t - (int64_t)rank) * cols;
        for (int c = 0; c < cols; c++) o[c] = r[c];
    }
    if (tile == lb.n_tiles - 1 && threadIdx.x == 0) *n_sel = (int64_t)tile_end;
}

}  // namespace

static inline unsigned blocks_for(int64_t n) { return (unsigned)((n + HB - 1) / HB); }

extern "C" int pcacc_velo2frame(pcacc_t h, const void *pts_dev, int pts_f64, int64_t n, int stride,
                                const double *P, double *out_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || stride < 3 || !P || (n > 0 && (!pts_dev || !out_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad velo2frame arguments");
    if (n == 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    Mat34 pm;
    memcpy(pm.m, P, sizeof(pm.m));
    h->launches[PCACC_K_EXPORT]++;
    if (pts_f64)
        k_velo2frame<double><<<blocks_for(n), HB, 0, st>>>((const double *)pts_dev, n, stride, pm, out_dev);
    else
        k_velo2frame<float><<<blocks_for(n), HB, 0, st>>>((const float *)pts_dev, n, stride, pm, out_dev);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_preprocess_pc(pcacc_t h, const double *pc_dev, int64_t n, int cols, const double *rot,
                                   double trans_dx, double trans_dy, double crop_view, double height_filter,
                                   double grid_view, int P, double *out_dev, int64_t *n_kept_dev,
                                   void *stream) {
    if (!h) return PCACC_ERR_ARG;
    const bool to_grid = grid_view == grid_view;
    if (n < 0 || cols < 3 || !n_kept_dev || (to_grid && (P <= 0 || grid_view == 0.0)) ||
        (n > 0 && (!pc_dev || !out_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad preprocess_pc arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    if (n == 0) {
        PCACC_CUDA(h, cudaMemsetAsync(n_kept_dev, 0, sizeof(int64_t), st));
        return PCACC_OK;
    }
    const int64_t tiles = (n + HB - 1) / HB;
    int rc = pcacc_ensure_tiles(h, tiles);
    if (rc) return rc;
    LookBackH lb{h->d_tile_state, h->d_ticket, pcacc_next_epoch(h), (uint32_t)tiles};
    Mat33 R;
    memset(&R, 0, sizeof(R));
    if (rot) memcpy(R.m, rot, sizeof(R.m));
    h->launches[PCACC_K_EXPORT]++;
    k_preprocess_pc<<<(unsigned)tiles, HB, 0, st>>>(
        pc_dev, n, cols, rot ? 1 : 0, R, trans_dx, trans_dy, crop_view == crop_view ? 1 : 0, crop_view,
        height_filter == height_filter ? 1 : 0, height_filter, to_grid ? 1 : 0, grid_view, P, out_dev,
        n_kept_dev, lb);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_cell_stats(pcacc_t h, const double *pc_dev, int64_t n, int cols, int P, int sem_col,
                                const int32_t *sems, int n_sems, int weight_col, const double *weights_dev,
                                int finish, double *count_sel_dev, double *count_rest_dev,
                                double *wsum_sel_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || cols < 2 || P <= 0 || n_sems > HELPER_MAX_SEMS || (n_sems > 0 && !sems) ||
        (n_sems >= 0 && (sem_col < 0 || sem_col >= cols)) || weight_col >= cols || (n > 0 && !pc_dev) ||
        (wsum_sel_dev && weight_col < 0 && !weights_dev) || finish < 0 || finish > 2 ||
        (finish == 1 && (!count_sel_dev || !count_rest_dev)) ||
        (finish == 2 && (!count_sel_dev || !wsum_sel_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad cell_stats arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t cells = (int64_t)P * P;
    if (count_sel_dev) PCACC_CUDA(h, cudaMemsetAsync(count_sel_dev, 0, (size_t)cells * 8, st));
    if (count_rest_dev) PCACC_CUDA(h, cudaMemsetAsync(count_rest_dev, 0, (size_t)cells * 8, st));
    if (wsum_sel_dev) PCACC_CUDA(h, cudaMemsetAsync(wsum_sel_dev, 0, (size_t)cells * 8, st));
    SemSel sel;
    sel.n = n_sems;
    for (int k = 0; k < n_sems; k++) sel.v[k] = (double)sems[k];
    if (n > 0) {
        h->launches[PCACC_K_EXPORT]++;
        k_cell_stats<<<blocks_for(n), HB, 0, st>>>(pc_dev, n, cols, P, sem_col, sel, weight_col, weights_dev,
                                                   count_sel_dev, count_rest_dev, wsum_sel_dev);
    }
    if (finish) {
        h->launches[PCACC_K_EXPORT]++;
        k_cell_finish<<<blocks_for(cells), HB, 0, st>>>(cells, finish, count_sel_dev, count_rest_dev,
                                                        wsum_sel_dev);
    }
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_dirichlet_expectation(pcacc_t h, double *maps_dev, int n_maps, int64_t cells,
                                           double obs_weight, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n_maps < 1 || cells < 0 || (cells > 0 && !maps_dev))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad dirichlet_expectation arguments");
    if (cells == 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    h->launches[PCACC_K_EXPORT]++;
    k_dirichlet<<<blocks_for(cells), HB, 0, st>>>(maps_dev, n_maps, cells, obs_weight);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_road_marking(pcacc_t h, const double *in_dev, int64_t n, double int_scaler,
                                  double int_sep_scaler, double int_mid_threshold, int sigmoid_only,
                                  double *out_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || (n > 0 && (!in_dev || !out_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad road_marking arguments");
    if (n == 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    h->launches[PCACC_K_EXPORT]++;
    k_road_marking<<<blocks_for(n), HB, 0, st>>>(in_dev, n, int_scaler, int_sep_scaler, int_mid_threshold,
                                                 sigmoid_only, out_dev);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_partition_semantic_pc(pcacc_t h, const double *pc_dev, int64_t n, int cols, int sem_col,
                                           const int32_t *sems, int n_sems, double *out_sel_dev,
                                           double *out_rest_dev, int64_t *n_sel_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || cols < 1 || sem_col < 0 || sem_col >= cols || n_sems < 0 || n_sems > HELPER_MAX_SEMS ||
        (n_sems > 0 && !sems) || !n_sel_dev || (n > 0 && (!pc_dev || !out_sel_dev || !out_rest_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad partition_semantic_pc arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    if (n == 0) {
        PCACC_CUDA(h, cudaMemsetAsync(n_sel_dev, 0, sizeof(int64_t), st));
        return PCACC_OK;
    }
    const int64_t tiles = (n + HB - 1) / HB;
    int rc = pcacc_ensure_tiles(h, tiles);
    if (rc) return rc;
    LookBackH lb{h->d_tile_state, h->d_ticket, pcacc_next_epoch(h), (uint32_t)tiles};
    SemSel sel;
    sel.n = n_sems;
    for (int k = 0; k < n_sems; k++) sel.v[k] = (double)sems[k];
    h->launches[PCACC_K_EXPORT]++;
    k_partition_sem<<<(unsigned)tiles, HB, 0, st>>>(pc_dev, n, cols, sem_col, sel, out_sel_dev, out_rest_dev,
                                                    n_sel_dev, lb);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}
