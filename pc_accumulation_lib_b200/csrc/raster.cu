// raster.cu — BEV rasterisation of the resident ring (K7-K12 of SURVEY.md §2.1).
//
//   k_bev_bin      shift / lazy re-base / rotate / translate / crop / height /
//                  static filter / pos2grid -> cell key; per-(cell,window)
//                  counting (the returned count is the point's rank in its
//                  segment) and block-aggregated append of a 16 B record
//   k_scan         single-pass chained exclusive scan of the counters
//   k_bev_scatter  counting-sort placement: sorted[start[key] + rank] = record
//   k_bev_reduce   per-cell reductions (counts, fixed-point intensity sum,
//                  min/max z, 256-bin histogram medians) for present / future /
//                  full, Dirichlet + sigmoid finalisation, float16 planes
//
// Every reduction is order-independent (integers, min/max, exact fixed-point
// sums, histograms), so the unordered placement inside a cell segment never
// shows in the output: results are deterministic run to run.
//
// Reference behaviour restated: bev_generator/bev_generator.py:127-160,207-255,
// 373-480,737-747; bev_generator/sem_bev.py:54-118,196-257,535-554,593-669;
// window split + origin shift kitti360_sem_pc_accum.py:179-213.
#include <math.h>
#include <string.h>

#include "common.cuh"

#define BIN_BLOCK 256
#define BIN_ITEMS 4
#define BIN_TILE (BIN_BLOCK * BIN_ITEMS)
#define MAX_VGROUP 32

#define SCAN_BLOCK 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)

#define RED_WARPS 4

// guard band (metres) inside which a lazily re-based point is re-evaluated with
// the exact sequential chain; composed-vs-sequential error is < 1e-11 m for
// chains of hundreds of frames, so 1e-7 is four orders of margin
#define GUARD_M 1e-7

static_assert(sizeof(pcacc_bev_params) == 208, "pcacc_bev_params layout is part of the ABI");

struct BinArgs {
    RingDev ring;
    const int64_t *frame_off, *frame_cnt, *frame_epoch;
    const double *comp, *chain;
    int max_frames;
    int64_t frame_lo;      // absolute id of blockIdx.y == 0
    int64_t epoch_now;
    const pcacc_bev_params *params;  // device, n_var entries
    int n_var;
    int v_per_block;       // variants handled by one block (blockIdx.z selects the group)
    int P;
    uint32_t *counts;      // n_var * 2*P*P (+1)
    uint32_t *tmp_key, *tmp_rank;
    uint4 *tmp_rec;
    unsigned long long *n_append;  // device counter
    unsigned long long *n_replay;
    int64_t cap;           // capacity of the tmp arrays
    int32_t *dbg_cell;     // optional
    uint32_t *flags;
};

struct Eval {
    bool keep, near;
    int cell;
    double z;
};

// One point, one variant: bev_generator.py:224-255,737-747 on a point that is already in
// the accumulator's current frame.  (px, py) are always known; pz is fetched through
// `zload` only when it is needed: always if the rotation mixes z into x / y (never for
// rotation_matrix_3d) and otherwise only for points that pass the x / y crop.
// want_near: also report whether the point lies within GUARD_M of any decision boundary.
template <typename ZLoad>
__device__ __forceinline__ Eval eval_point(const pcacc_bev_params &bp, int P, double px, double py,
                                           ZLoad zload, bool want_near) {
    Eval e;
    e.keep = false;
    e.near = false;
    e.cell = -1;
    e.z = 0.0;
    const bool zfree = (bp.R[2] == 0.0) && (bp.R[5] == 0.0);
    // origin shift: one subtract per component (kitti360_sem_pc_accum.py:193)
    const double sx = __dsub_rn(px, bp.origin[0]);
    const double sy = __dsub_rn(py, bp.origin[1]);
    double sz = 0.0;
    if (!zfree) sz = __dsub_rn(zload(), bp.origin[2]);
    // rotation: FMA chain over k = 0..2 (strided 3x3 dgemm, bev_generator.py:227);
    // a zero coefficient contributes exactly nothing for finite z, so that link is skipped
    double q0 = __fma_rn(bp.R[1], sy, __dmul_rn(bp.R[0], sx));
    double q1 = __fma_rn(bp.R[4], sy, __dmul_rn(bp.R[3], sx));
    if (!zfree) {
        q0 = __fma_rn(bp.R[2], sz, q0);
        q1 = __fma_rn(bp.R[5], sz, q1);
    }
    q0 = __dadd_rn(q0, bp.trans_dx);
    q1 = __dadd_rn(q1, bp.trans_dy);
    const double hv = __dmul_rn(0.5, bp.view);
    bool in = (q0 > -hv) && (q0 < hv) && (q1 > -hv) && (q1 < hv);
    const bool cand = want_near ? ((fabs(q0) < hv + GUARD_M) && (fabs(q1) < hv + GUARD_M)) : in;
    if (!cand) return e;
    if (zfree) {
        const double pz = zload();
        // a non-finite z poisons x and y in the reference (0 * inf = NaN): the point is dropped
        if (!(fabs(pz) <= 1.7976931348623157e308)) return e;
        sz = __dsub_rn(pz, bp.origin[2]);
    }
    const double q2 = __fma_rn(bp.R[8], sz, __fma_rn(bp.R[7], sy, __dmul_rn(bp.R[6], sx)));
    e.z = q2;
    const bool hf_on = (bp.height_filter == bp.height_filter);
    if (hf_on) in = in && (q2 < bp.height_filter);
    const double dP = (double)P, hP = __dmul_rn(0.5, dP);
    const double g0 = __dadd_rn(__dmul_rn(__ddiv_rn(q0, bp.view), dP), hP);
    const double g1 = __dadd_rn(__dmul_rn(__ddiv_rn(q1, bp.view), dP), hP);
    if (want_near) {
        const double eg = GUARD_M * dP / bp.view;
        bool n = (fabs(fabs(q0) - hv) < GUARD_M) || (fabs(fabs(q1) - hv) < GUARD_M);
        if (hf_on) n = n || (fabs(q2 - bp.height_filter) < GUARD_M);
        n = n || (fabs(g0 - rint(g0)) < eg) || (fabs(g1 - rint(g1)) < eg);
        e.near = n;
    }
    if (in) {
        const double fi = floor(g0), fj = floor(g1);
        if (fi >= 0.0 && fi < dP && fj >= 0.0 && fj < dP) {
            const int i = (int)fi, j = (int)fj;
            e.cell = (P - 1 - j) * P + i;  // row = P-1-j, col = i (bev_generator.py:453)
            e.keep = true;
        }
    }
    return e;
}

__global__ void __launch_bounds__(BIN_BLOCK, 2)
k_bev_bin(BinArgs a) {
    __shared__ pcacc_bev_params s_par[MAX_VGROUP];
    __shared__ double s_comp[12];
    __shared__ uint32_t s_warp[BIN_BLOCK / 32];
    __shared__ unsigned long long s_base;

    const int64_t fid = a.frame_lo + blockIdx.y;
    const int slot = (int)(fid % a.max_frames);
    const int64_t cnt = a.frame_cnt[slot];
    const int64_t tile0 = (int64_t)blockIdx.x * BIN_TILE;
    if (tile0 >= cnt) return;
    const int64_t off = a.frame_off[slot];
    const int64_t e0 = a.frame_epoch[slot];
    const bool lazy = e0 < a.epoch_now;

    // stage this block's variant parameters and the frame's composed matrix
    const int v_begin = (int)blockIdx.z * a.v_per_block;
    const int v_end = min(a.n_var, v_begin + a.v_per_block);
    {
        const uint32_t *src = (const uint32_t *)(a.params + v_begin);
        uint32_t *dst = (uint32_t *)s_par;
        const int words = (v_end - v_begin) * (int)(sizeof(pcacc_bev_params) / 4);
        for (int k = threadIdx.x; k < words; k += BIN_BLOCK) dst[k] = src[k];
        if (threadIdx.x < 12) s_comp[threadIdx.x] = a.comp[(int64_t)slot * 12 + threadIdx.x];
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int PP = a.P * a.P;

    // This thread's points: two pairs of neighbours (16 B loads; frame offsets are
    // multiples of 4 records, so the pairs are aligned).  x and y are always needed; z
    // only up front when the frame is lazily re-based (the composed matrix mixes it in).
    int64_t idx[BIN_ITEMS];
    double px[BIN_ITEMS], py[BIN_ITEMS], pz[BIN_ITEMS];
    bool valid[BIN_ITEMS];
#pragma unroll
    for (int h = 0; h < BIN_ITEMS / 2; h++) {
        const int64_t i0 = tile0 + (int64_t)h * (2 * BIN_BLOCK) + 2 * threadIdx.x;
        idx[2 * h] = i0;
        idx[2 * h + 1] = i0 + 1;
        valid[2 * h] = i0 < cnt;
        valid[2 * h + 1] = i0 + 1 < cnt;
        double2 X = make_double2(0, 0), Y = make_double2(0, 0), Z = make_double2(0, 0);
        if (valid[2 * h]) {
            X = *(const double2 *)(a.ring.x + off + i0);
            Y = *(const double2 *)(a.ring.y + off + i0);
            if (lazy) Z = *(const double2 *)(a.ring.z + off + i0);
        }
        px[2 * h] = X.x; px[2 * h + 1] = X.y;
        py[2 * h] = Y.x; py[2 * h + 1] = Y.y;
        pz[2 * h] = Z.x; pz[2 * h + 1] = Z.y;
    }
    if (lazy) {
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            double nx, ny, nz;
            affine_chain(s_comp, 4, px[k], py[k], pz[k], nx, ny, nz);
            px[k] = nx; py[k] = ny; pz[k] = nz;
        }
    }

    for (int v = v_begin; v < v_end; v++) {
        const pcacc_bev_params &bp = s_par[v - v_begin];
        const bool in_range = fid >= bp.frame_begin && fid < bp.frame_end;  // block-uniform
        if (!in_range) continue;
        const uint32_t win = fid >= bp.frame_split ? 1u : 0u;
        bool keep[BIN_ITEMS];
        uint32_t key[BIN_ITEMS], rank[BIN_ITEMS];
        uint4 rec[BIN_ITEMS];
        uint32_t my_cnt = 0;
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            keep[k] = false;
            key[k] = 0;
            rank[k] = 0;
            rec[k] = make_uint4(0, 0, 0, 0);
            if (valid[k]) {
                const int64_t gi = off + idx[k];
                Eval e;
                if (lazy) {
                    const double zk = pz[k];
                    e = eval_point(bp, a.P, px[k], py[k], [zk]() { return zk; }, true);
                    if (e.near) {
                        // exact sequential re-base chain (update_sem_pcs, sem_pc_accum.py:167-183)
                        double ex = a.ring.x[gi], ey = a.ring.y[gi], ez = a.ring.z[gi];
                        for (int64_t ep = e0; ep < a.epoch_now; ep++) {
                            const double *T = a.chain + (ep % a.max_frames) * 12;
                            double nx, ny, nz;
                            affine_chain(T, 4, ex, ey, ez, nx, ny, nz);
                            ex = nx; ey = ny; ez = nz;
                        }
                        e = eval_point(bp, a.P, ex, ey, [ez]() { return ez; }, false);
                        atomicAdd(a.n_replay, 1ull);
                    }
                } else {
                    const double *zp = a.ring.z + gi;
                    e = eval_point(bp, a.P, px[k], py[k], [zp]() { return *zp; }, false);
                }
                if (e.keep && a.ring.dyn[gi] == 1) e.keep = false;  // static points only
                if (a.dbg_cell && v == 0) a.dbg_cell[gi] = e.keep ? e.cell : -1;
                if (e.keep) {
                    keep[k] = true;
                    key[k] = ((uint32_t)v * (uint32_t)PP + (uint32_t)e.cell) * 2u + win;
                    rank[k] = atomicAdd(&a.counts[key[k]], 1u);
                    unsigned long long zb = (unsigned long long)__double_as_longlong(e.z);
                    rec[k].x = (uint32_t)zb;
                    rec[k].y = (uint32_t)(zb >> 32);
                    rec[k].z = a.ring.rgbs[gi];
                    rec[k].w = __float_as_uint(a.ring.inten[gi]);
                    my_cnt++;
                }
            }
        }
        // block-aggregated append: one global atomic per block and variant
        uint32_t incl = my_cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, btotal = 0;
#pragma unroll
        for (int w = 0; w < BIN_BLOCK / 32; w++) {
            uint32_t c = s_warp[w];
            if (w < (int)warp) wbase += c;
            btotal += c;
        }
        if (threadIdx.x == 0) s_base = btotal ? atomicAdd(a.n_append, (unsigned long long)btotal) : 0ull;
        __syncthreads();
        unsigned long long pos = s_base + wbase + (incl - my_cnt);
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            if (keep[k]) {
                if ((int64_t)pos < a.cap) {
                    a.tmp_key[pos] = key[k];
                    a.tmp_rank[pos] = rank[k];
                    a.tmp_rec[pos] = rec[k];
                }
                pos++;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// exclusive scan of n u32 counters, in place (chained single pass)
// ---------------------------------------------------------------------------
struct ScanLB {
    unsigned long long *state;
    uint32_t *ticket;
    uint32_t epoch, n_tiles;
};

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan(uint32_t *__restrict__ data, int64_t n, ScanLB lb) {
    __shared__ uint32_t s_warp[SCAN_BLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n) {
        const uint4 *p = (const uint4 *)(data + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            uint4 q = p[k];
            v[4 * k] = q.x;
            v[4 * k + 1] = q.y;
            v[4 * k + 2] = q.z;
            v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? data[base + k] : 0u;
    }
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) tsum += v[k];
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = SCAN_BLOCK / 32;
        uint32_t c = (lane < NW) ? s_warp[lane] : 0u;
        uint32_t wi = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += t;
        }
        uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
        uint32_t ex = lb_exclusive_prefix(lb.state, lb.epoch, tile, total);
        if (lane < NW) s_warp[lane] = ex + wi - c;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - tsum;
    uint32_t o[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        o[k] = run;
        run += v[k];
    }
    if (base + SCAN_ITEMS <= n) {
        uint4 *p = (uint4 *)(data + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++)
            p[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++)
            if (base + k < n) data[base + k] = o[k];
    }
}

// ---------------------------------------------------------------------------
// scatter
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bev_scatter(const uint32_t *__restrict__ start, const uint32_t *__restrict__ tmp_key,
              const uint32_t *__restrict__ tmp_rank, const uint4 *__restrict__ tmp_rec,
              const unsigned long long *__restrict__ n_append, int64_t cap, uint4 *__restrict__ sorted) {
    unsigned long long n = *n_append;
    if (n > (unsigned long long)cap) n = (unsigned long long)cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * 256) {
        uint32_t pos = start[tmp_key[i]] + tmp_rank[i];
        sorted[pos] = tmp_rec[i];
    }
}

// ---------------------------------------------------------------------------
// per-cell reduction + finalisation
// ---------------------------------------------------------------------------
#define FX_SCALE 1099511627776.0        /* 2^40 */
#define FX_INV 9.094947017729282e-13    /* 2^-40 */

__device__ __forceinline__ double fx_to_double(long long hi, long long lo) {
    // hi*2^32 + lo with both parts far below 2^53: one rounding at the add
    return __dmul_rn(__dadd_rn(__dmul_rn((double)hi, 4294967296.0), (double)lo), FX_INV);
}

// value of rank k (0-based) in a 256-bin histogram spread 8 bins per lane
__device__ __forceinline__ int hist_kth(const uint32_t c[8], uint32_t excl, uint32_t incl, uint32_t k,
                                        unsigned lane) {
    bool mine = (excl <= k) && (k < incl);
    int bin = 0;
    if (mine) {
        uint32_t t = k - excl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (t < c[j]) {
                bin = (int)(lane * 8 + j);
                break;
            }
            t -= c[j];
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, mine);
    return __shfl_sync(0xffffffffu, bin, __ffs(m) - 1);
}

// Accumulates one record into the per-window statistics (everything except the medians).
struct WinAcc {
    uint32_t n_road[2], n_veh[2];
    long long fx_hi[2], fx_lo[2];
    double ext_z[2];
};

__device__ __forceinline__ void acc_record(WinAcc &a, const uint4 &r, int w, const pcacc_bev_params &bp,
                                           double intensity_div, bool want_max) {
    const double z = __longlong_as_double((long long)(((unsigned long long)r.y << 32) | r.x));
    const int sem = (int)(r.z >> 24);
    const bool road = sem == bp.road_cls;
    const bool veh = (sem == bp.veh_cls[0]) || (sem == bp.veh_cls[1]) || (sem == bp.veh_cls[2]) ||
                     (sem == bp.veh_cls[3]);
    if (road) {
        double iv = __ddiv_rn((double)__uint_as_float(r.w), intensity_div);
        long long fx = __double2ll_rn(__dmul_rn(iv, FX_SCALE));
        a.fx_hi[w] += fx >> 32;
        a.fx_lo[w] += fx & 0xffffffffll;
        a.n_road[w]++;
    }
    if (veh) a.n_veh[w]++;
    a.ext_z[w] = want_max ? fmax(a.ext_z[w], z) : fmin(a.ext_z[w], z);
}

// 7 planes of one window from its statistics (bev_generator.py:396-415,457-480;
// sem_bev.py:593-617,665-667)
__device__ __forceinline__ void finalise_window(const pcacc_bev_params &bp, double n, double n_road,
                                                double n_veh, double isum, double ez, const int med2[3],
                                                double plane[7]) {
    double a = __dadd_rn(n_road, 1.0), b = __dadd_rn(__dsub_rn(n, n_road), 1.0);
    plane[0] = __ddiv_rn(a, __dadd_rn(a, b));
    double c = __dadd_rn(n_veh, 1.0), d = __dadd_rn(__dsub_rn(n, n_veh), 1.0);
    plane[5] = __ddiv_rn(c, __dadd_rn(c, d));
    double I = __ddiv_rn(isum, __dadd_rn(n_road, 1.0));
    double t = __dmul_rn(bp.int_sep_scaler, __dsub_rn(I, bp.int_mid_threshold));
    double sg = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-t)));
    double val = __dmul_rn(bp.int_scaler, sg);
    plane[1] = val > 1.0 ? 1.0 : val;
#pragma unroll
    for (int k = 0; k < 3; k++) plane[2 + k] = __ddiv_rn(__dmul_rn((double)med2[k], 0.5), 255.0);
    plane[6] = ez;
}

#define SMALL_T 16      /* cells with <= SMALL_T points are reduced by their own lane */
#define SMALL_STRIDE 17 /* odd stride: lanes reading the same j hit distinct banks */

template <bool F64OUT>
__global__ void __launch_bounds__(RED_WARPS * 32)
k_bev_reduce(const uint32_t *__restrict__ start, const uint4 *__restrict__ sorted,
             const pcacc_bev_params *__restrict__ params, int n_var, int P, double intensity_div,
             __half *__restrict__ out16, double *__restrict__ out64) {
    __shared__ __align__(16) uint32_t s_hist[RED_WARPS][2][3][256];
    __shared__ uint32_t s_small[RED_WARPS][32 * SMALL_STRIDE];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int PP = P * P;
    const int64_t n_cells = (int64_t)n_var * PP;
    const int64_t gw = (int64_t)blockIdx.x * RED_WARPS + warp;
    const int64_t cell0 = gw * 32;
    if (cell0 >= n_cells) return;
    const int var = (int)(cell0 / PP);
    const int cell_in = (int)(cell0 - (int64_t)var * PP) + (int)lane;
    const pcacc_bev_params &bp = params[var];
    uint32_t(*hist)[3][256] = s_hist[warp];
    const bool want_max = bp.elevation_max != 0;

    // segment bounds of my cell: [s0, s1) present, [s1, s2) future
    const int64_t gc = cell0 + lane;
    const uint2 s01 = ((const uint2 *)start)[gc];
    uint32_t s2 = __shfl_down_sync(0xffffffffu, s01.x, 1);
    if (lane == 31) s2 = start[2 * gc + 2];
    const uint32_t my_np = s01.y - s01.x, my_nf = s2 - s01.y, my_nt = my_np + my_nf;

    WinAcc st;
#pragma unroll
    for (int w = 0; w < 2; w++) {
        st.n_road[w] = st.n_veh[w] = 0;
        st.fx_hi[w] = st.fx_lo[w] = 0;
        st.ext_z[w] = want_max ? -INFINITY : INFINITY;
    }
    int med2[3][3];
#pragma unroll
    for (int w = 0; w < 3; w++)
#pragma unroll
        for (int c = 0; c < 3; c++) med2[w][c] = 0;

    // ---- small cells: each lane reduces its own cell ---------------------------------
    if (my_nt > 0 && my_nt <= SMALL_T) {
        uint32_t *mine = s_small[warp] + lane * SMALL_STRIDE;
        for (uint32_t i = 0; i < my_nt; i++) {
            const uint4 r = sorted[s01.x + i];
            mine[i] = r.z & 0x00ffffffu;
            acc_record(st, r, i >= my_np ? 1 : 0, bp, intensity_div, want_max);
        }
        // rank of every colour byte inside its window and inside the full cell; ties
        // broken by position, so ranks are a permutation and rank == k picks the
        // k-th order statistic.  r, g, b are processed together, one byte lane each.
        const uint32_t kp_lo = my_np ? (my_np - 1) / 2 * 0x010101u : 0xffffffffu;
        const uint32_t kp_hi = my_np ? my_np / 2 * 0x010101u : 0xffffffffu;
        const uint32_t kf_lo = my_nf ? (my_nf - 1) / 2 * 0x010101u : 0xffffffffu;
        const uint32_t kf_hi = my_nf ? my_nf / 2 * 0x010101u : 0xffffffffu;
        const uint32_t ka_lo = (my_nt - 1) / 2 * 0x010101u, ka_hi = my_nt / 2 * 0x010101u;
        uint32_t lo_p = 0, hi_p = 0, lo_f = 0, hi_f = 0, lo_a = 0, hi_a = 0;
        for (uint32_t i = 0; i < my_nt; i++) {
            const uint32_t vi = mine[i];
            const bool wi = i >= my_np;
            uint32_t rank_w = 0, rank_a = 0;
            for (uint32_t j = 0; j < my_nt; j++) {
                const uint32_t vj = mine[j];
                uint32_t m = __vcmpltu4(vj, vi);
                if (j < i) m |= __vcmpeq4(vj, vi);
                m &= 0x00010101u;
                rank_a += m;
                if ((j >= my_np) == wi) rank_w += m;
            }
            uint32_t m;
            m = __vcmpeq4(rank_a, ka_lo) & 0x00ffffffu;
            lo_a = (vi & m) | (lo_a & ~m);
            m = __vcmpeq4(rank_a, ka_hi) & 0x00ffffffu;
            hi_a = (vi & m) | (hi_a & ~m);
            if (wi) {
                m = __vcmpeq4(rank_w, kf_lo) & 0x00ffffffu;
                lo_f = (vi & m) | (lo_f & ~m);
                m = __vcmpeq4(rank_w, kf_hi) & 0x00ffffffu;
                hi_f = (vi & m) | (hi_f & ~m);
            } else {
                m = __vcmpeq4(rank_w, kp_lo) & 0x00ffffffu;
                lo_p = (vi & m) | (lo_p & ~m);
                m = __vcmpeq4(rank_w, kp_hi) & 0x00ffffffu;
                hi_p = (vi & m) | (hi_p & ~m);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            med2[0][c] = (int)((lo_p >> (8 * c)) & 255u) + (int)((hi_p >> (8 * c)) & 255u);
            med2[1][c] = (int)((lo_f >> (8 * c)) & 255u) + (int)((hi_f >> (8 * c)) & 255u);
            med2[2][c] = (int)((lo_a >> (8 * c)) & 255u) + (int)((hi_a >> (8 * c)) & 255u);
        }
    }
    __syncwarp();

    // ---- large cells: the whole warp reduces one cell at a time with histograms ---------
    unsigned todo = __ballot_sync(0xffffffffu, my_nt > SMALL_T);
    while (todo) {
        const int owner = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t b0 = __shfl_sync(0xffffffffu, s01.x, owner);
        const uint32_t np = __shfl_sync(0xffffffffu, my_np, owner);
        const uint32_t nf = __shfl_sync(0xffffffffu, my_nf, owner);
        const uint32_t nt = np + nf;
        {
            uint4 *h4 = (uint4 *)&hist[0][0][0];
#pragma unroll
            for (int k = 0; k < 12; k++) h4[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        WinAcc a;
#pragma unroll
        for (int w = 0; w < 2; w++) {
            a.n_road[w] = a.n_veh[w] = 0;
            a.fx_hi[w] = a.fx_lo[w] = 0;
            a.ext_z[w] = want_max ? -INFINITY : INFINITY;
        }
        for (uint32_t i = lane; i < nt; i += 32) {
            const uint4 r = sorted[b0 + i];
            const int w = i >= np ? 1 : 0;
            const uint32_t c = r.z;
            atomicAdd(&hist[w][0][c & 255u], 1u);
            atomicAdd(&hist[w][1][(c >> 8) & 255u], 1u);
            atomicAdd(&hist[w][2][(c >> 16) & 255u], 1u);
            acc_record(a, r, w, bp, intensity_div, want_max);
        }
#pragma unroll
        for (int w = 0; w < 2; w++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a.n_road[w] += __shfl_xor_sync(0xffffffffu, a.n_road[w], o);
                a.n_veh[w] += __shfl_xor_sync(0xffffffffu, a.n_veh[w], o);
                a.fx_hi[w] += __shfl_xor_sync(0xffffffffu, a.fx_hi[w], o);
                a.fx_lo[w] += __shfl_xor_sync(0xffffffffu, a.fx_lo[w], o);
                double oz = __shfl_xor_sync(0xffffffffu, a.ext_z[w], o);
                a.ext_z[w] = want_max ? fmax(a.ext_z[w], oz) : fmin(a.ext_z[w], oz);
            }
        }
        __syncwarp();
        int m2[3][3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            uint32_t cp[8], cf[8];
            {
                const uint4 *hp = (const uint4 *)&hist[0][ch][lane * 8];
                const uint4 *hf = (const uint4 *)&hist[1][ch][lane * 8];
                uint4 x0 = hp[0], x1 = hp[1], y0 = hf[0], y1 = hf[1];
                cp[0] = x0.x; cp[1] = x0.y; cp[2] = x0.z; cp[3] = x0.w;
                cp[4] = x1.x; cp[5] = x1.y; cp[6] = x1.z; cp[7] = x1.w;
                cf[0] = y0.x; cf[1] = y0.y; cf[2] = y0.z; cf[3] = y0.w;
                cf[4] = y1.x; cf[5] = y1.y; cf[6] = y1.z; cf[7] = y1.w;
            }
#pragma unroll
            for (int w = 0; w < 3; w++) {
                uint32_t c[8];
#pragma unroll
                for (int j = 0; j < 8; j++) c[j] = (w == 0) ? cp[j] : (w == 1) ? cf[j] : cp[j] + cf[j];
                const uint32_t nw = (w == 0) ? np : (w == 1) ? nf : nt;
                int v = 0;
                if (nw > 0) {  // warp-uniform
                    uint32_t s = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) s += c[j];
                    uint32_t incl = s;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= (unsigned)o) incl += t;
                    }
                    int lo = hist_kth(c, incl - s, incl, (nw - 1) / 2, lane);
                    int hi = (nw & 1u) ? lo : hist_kth(c, incl - s, incl, nw / 2, lane);
                    v = lo + hi;
                }
                m2[w][ch] = v;
            }
        }
        if ((int)lane == owner) {
            st = a;
#pragma unroll
            for (int w = 0; w < 3; w++)
#pragma unroll
                for (int c = 0; c < 3; c++) med2[w][c] = m2[w][c];
        }
        __syncwarp();
    }

    // ---- finalise my cell: 3 windows x 7 planes ------------------------------------------
    // values of a window without any point: computed once per warp, not per cell
    double empty[7];
    {
        double e1 = 0.0;
        if (lane == 0) {
            const int zero3[3] = {0, 0, 0};
            double pl[7];
            finalise_window(bp, 0.0, 0.0, 0.0, 0.0, 0.0, zero3, pl);
            e1 = pl[1];
        }
        e1 = __shfl_sync(0xffffffffu, e1, 0);
        empty[0] = 0.5;
        empty[1] = e1;
        empty[2] = empty[3] = empty[4] = __ddiv_rn(bp.rgb_fill, 255.0);
        empty[5] = 0.5;
        empty[6] = 0.0;
    }
#pragma unroll
    for (int w = 0; w < 3; w++) {
        double plane[7];
        const uint32_t nw = (w == 0) ? my_np : (w == 1) ? my_nf : my_nt;
        if (nw == 0) {
#pragma unroll
            for (int p = 0; p < 7; p++) plane[p] = empty[p];
        } else {
            double n_road, n_veh, isum, ez;
            if (w < 2) {
                n_road = (double)st.n_road[w];
                n_veh = (double)st.n_veh[w];
                isum = fx_to_double(st.fx_hi[w], st.fx_lo[w]);
                ez = st.ext_z[w];
            } else {
                n_road = (double)(st.n_road[0] + st.n_road[1]);
                n_veh = (double)(st.n_veh[0] + st.n_veh[1]);
                isum = fx_to_double(st.fx_hi[0] + st.fx_hi[1], st.fx_lo[0] + st.fx_lo[1]);
                // the unused side still holds +-inf, the identity of min / max
                ez = want_max ? fmax(st.ext_z[0], st.ext_z[1]) : fmin(st.ext_z[0], st.ext_z[1]);
            }
            finalise_window(bp, (double)nw, n_road, n_veh, isum, ez, med2[w], plane);
        }
        const int64_t o = (((int64_t)var * 3 + w) * 7) * PP + cell_in;
#pragma unroll
        for (int p = 0; p < 7; p++) {
            out16[o + (int64_t)p * PP] = __double2half(plane[p]);
            if (F64OUT) out64[o + (int64_t)p * PP] = plane[p];
        }
    }
}

// ===========================================================================
// host
// ===========================================================================
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int ensure_ws(pcacc_t h, size_t bytes) {
    if (bytes <= h->ws_size) return PCACC_OK;
    if (h->d_ws) {
        PCACC_CUDA(h, cudaDeviceSynchronize());
        cudaFree(h->d_ws);
        h->d_ws = nullptr;
        h->ws_size = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&h->d_ws, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&h->d_ws, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return pcacc_fail(h, PCACC_ERR_NOMEM, "raster workspace of %zu bytes: %s", bytes,
                          cudaGetErrorString(e));
    }
    h->ws_size = want;
    return PCACC_OK;
}

extern "C" int pcacc_rasterise(pcacc_t h, const pcacc_bev_params *params, int n_variants, int P,
                               void *out_f16_dev, double *out_f64_dev, int32_t *dbg_cell_dev,
                               void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (!params || n_variants <= 0 || P <= 0 || (P * P) % 32 != 0 || !out_f16_dev)
        return pcacc_fail(h, PCACC_ERR_ARG, "bad rasterise arguments (P*P must be a multiple of 32)");
    if (h->inten_div == 0.0) h->inten_div = 1.0;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t PP = (int64_t)P * P;

    for (int v0 = 0; v0 < n_variants; v0 += MAX_VGROUP) {
        const int nv = n_variants - v0 < MAX_VGROUP ? n_variants - v0 : MAX_VGROUP;
        if ((int64_t)nv * PP * 2 + 1 > 0xffffffffll)
            return pcacc_fail(h, PCACC_ERR_ARG, "grid too large for 32-bit keys");
        // frame span and visit upper bound of this group
        int64_t flo = INT64_MAX, fhi = INT64_MIN;
        for (int v = 0; v < nv; v++) {
            const pcacc_bev_params &bp = params[v0 + v];
            if (bp.frame_begin < h->first_id || bp.frame_end > h->next_id ||
                bp.frame_begin > bp.frame_split || bp.frame_split > bp.frame_end)
                return pcacc_fail(h, PCACC_ERR_ARG,
                                  "variant %d: frames [%lld,%lld,%lld) outside live range [%lld,%lld)",
                                  v0 + v, (long long)bp.frame_begin, (long long)bp.frame_split,
                                  (long long)bp.frame_end, (long long)h->first_id,
                                  (long long)h->next_id);
            if (!(bp.view > 0.0)) return pcacc_fail(h, PCACC_ERR_ARG, "view must be positive");
            if (bp.frame_begin < flo) flo = bp.frame_begin;
            if (bp.frame_end > fhi) fhi = bp.frame_end;
        }
        int64_t max_cnt = 0;
        std::vector<int64_t> ub((size_t)(fhi > flo ? fhi - flo : 0));
        for (int64_t f = flo; f < fhi; f++) {
            FrameHost &fh = h->frames[(int)(f % h->max_frames)];
            int64_t c = fh.exact ? fh.cnt : fh.n_in;
            ub[(size_t)(f - flo)] = c;
            if (c > max_cnt) max_cnt = c;
        }
        int64_t cap = 0;
        for (int v = 0; v < nv; v++) {
            const pcacc_bev_params &bp = params[v0 + v];
            for (int64_t f = bp.frame_begin; f < bp.frame_end; f++) cap += ub[(size_t)(f - flo)];
        }
        if (cap > 0xfffffff0ll)
            return pcacc_fail(h, PCACC_ERR_CAPACITY, "more than 2^32 point visits in one batch");
        const int64_t n_keys = (int64_t)nv * PP * 2 + 1;
        // workspace layout
        size_t o_counts = 0;
        size_t o_cnt2 = align_up(o_counts + (size_t)n_keys * 4, 256);  // two u64 counters
        size_t o_key = align_up(o_cnt2 + 16, 256);
        size_t o_rank = align_up(o_key + (size_t)cap * 4, 256);
        size_t o_rec = align_up(o_rank + (size_t)cap * 4, 256);
        size_t o_sorted = align_up(o_rec + (size_t)cap * 16, 256);
        size_t total = align_up(o_sorted + (size_t)cap * 16, 256);
        int rc = ensure_ws(h, total);
        if (rc) return rc;
        char *ws = (char *)h->d_ws;
        uint32_t *counts = (uint32_t *)(ws + o_counts);
        unsigned long long *ctr = (unsigned long long *)(ws + o_cnt2);
        // counters + append/replay counters are contiguous up to o_key: one memset
        PCACC_CUDA(h, cudaMemsetAsync(ws, 0, o_key, st));

        void *d_params = nullptr;
        rc = pcacc_arena_put(h, params + v0, (size_t)nv * sizeof(pcacc_bev_params), &d_params, st);
        if (rc) return rc;

        const bool want_f64 = out_f64_dev != nullptr;
        __half *o16 = (__half *)out_f16_dev + (int64_t)v0 * 21 * PP;
        double *o64 = want_f64 ? out_f64_dev + (int64_t)v0 * 21 * PP : nullptr;

        if (cap > 0 && fhi > flo && max_cnt > 0) {
            BinArgs a;
            a.ring = h->ring;
            a.frame_off = h->d_frame_off;
            a.frame_cnt = h->d_frame_cnt;
            a.frame_epoch = h->d_frame_epoch;
            a.comp = h->d_comp;
            a.chain = h->d_chain;
            a.max_frames = h->max_frames;
            a.frame_lo = flo;
            a.epoch_now = h->rebase_epoch;
            a.params = (const pcacc_bev_params *)d_params;
            a.n_var = nv;
            a.P = P;
            a.counts = counts;
            a.tmp_key = (uint32_t *)(ws + o_key);
            a.tmp_rank = (uint32_t *)(ws + o_rank);
            a.tmp_rec = (uint4 *)(ws + o_rec);
            a.n_append = ctr;
            a.n_replay = ctr + 1;
            a.cap = cap;
            a.dbg_cell = (v0 == 0) ? dbg_cell_dev : nullptr;
            a.flags = h->d_flags;
            int64_t nf = fhi - flo;
            // spread the variants over blockIdx.z until the grid fills the chip
            int64_t useful = 0;
            for (int64_t f = flo; f < fhi; f++) useful += (ub[(size_t)(f - flo)] + BIN_TILE - 1) / BIN_TILE;
            int groups = (int)(148 * 6 / (useful > 0 ? useful : 1));
            if (groups < 1) groups = 1;
            if (groups > nv) groups = nv;
            a.v_per_block = (nv + groups - 1) / groups;
            groups = (nv + a.v_per_block - 1) / a.v_per_block;
            for (int64_t f0 = 0; f0 < nf; f0 += 65535) {
                int64_t ny = nf - f0 < 65535 ? nf - f0 : 65535;
                BinArgs b = a;
                b.frame_lo = flo + f0;
                dim3 grid((unsigned)((max_cnt + BIN_TILE - 1) / BIN_TILE), (unsigned)ny,
                          (unsigned)groups);
                size_t pe = pcacc_prof_begin(h, PCACC_K_BIN, st);
                k_bev_bin<<<grid, BIN_BLOCK, 0, st>>>(b);
                PCACC_CUDA(h, cudaGetLastError());
                pcacc_prof_end(h, PCACC_K_BIN, pe, st);
            }
            // scan
            int64_t tiles = (n_keys + SCAN_TILE - 1) / SCAN_TILE;
            rc = pcacc_ensure_tiles(h, tiles);
            if (rc) return rc;
            ScanLB lb;
            lb.state = h->d_tile_state;
            lb.ticket = h->d_ticket;
            lb.epoch = pcacc_next_epoch(h);
            lb.n_tiles = (uint32_t)tiles;
            size_t pe = pcacc_prof_begin(h, PCACC_K_SCAN, st);
            k_scan<<<(unsigned)tiles, SCAN_BLOCK, 0, st>>>(counts, n_keys, lb);
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_SCAN, pe, st);
            // scatter
            int64_t sb = (cap + 255) / 256;
            if (sb > 148 * 16) sb = 148 * 16;
            pe = pcacc_prof_begin(h, PCACC_K_SCATTER, st);
            k_bev_scatter<<<(unsigned)sb, 256, 0, st>>>(counts, a.tmp_key, a.tmp_rank, a.tmp_rec, ctr,
                                                        cap, (uint4 *)(ws + o_sorted));
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_SCATTER, pe, st);
        }
        // reduce + finalise (also correct on all-zero counters: every cell empty)
        int64_t warps = (int64_t)nv * PP / 32;
        int64_t blocks = (warps + RED_WARPS - 1) / RED_WARPS;
        size_t pr = pcacc_prof_begin(h, PCACC_K_REDUCE, st);
        if (want_f64)
            k_bev_reduce<true><<<(unsigned)blocks, RED_WARPS * 32, 0, st>>>(
                counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params, nv, P,
                h->inten_div, o16, o64);
        else
            k_bev_reduce<false><<<(unsigned)blocks, RED_WARPS * 32, 0, st>>>(
                counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params, nv, P,
                h->inten_div, o16, nullptr);
        PCACC_CUDA(h, cudaGetLastError());
        pcacc_prof_end(h, PCACC_K_REDUCE, pr, st);
        // keep the counters of the last group for pcacc_raster_stats
        PCACC_CUDA(h, cudaMemcpyAsync(h->d_rstats + 1, ctr, 16, cudaMemcpyDeviceToDevice, st));
        h->last_visit_ub = cap;
    }
    return PCACC_OK;
}

extern "C" int pcacc_raster_stats(pcacc_t h, int64_t stats[3], void *stream) {
    if (!h || !stats) return PCACC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    PCACC_CUDA(h, cudaMemcpyAsync(h->h_mail, h->d_rstats, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PCACC_CUDA(h, cudaStreamSynchronize(st));
    stats[0] = h->last_visit_ub;
    stats[1] = h->h_mail[1];
    stats[2] = h->h_mail[2];
    return PCACC_OK;
}
