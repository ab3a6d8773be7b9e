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

// one point, one variant: bev_generator.py:224-255,737-747 on (px,py,pz) that is
// already in the accumulator's current frame
__device__ __forceinline__ Eval eval_point(const pcacc_bev_params &bp, int P, double px, double py,
                                           double pz, bool want_near) {
    Eval e;
    e.keep = false;
    e.near = false;
    e.cell = -1;
    // origin shift: one subtract per component (kitti360_sem_pc_accum.py:193)
    double sx = __dsub_rn(px, bp.origin[0]);
    double sy = __dsub_rn(py, bp.origin[1]);
    double sz = __dsub_rn(pz, bp.origin[2]);
    // rotation: FMA chain over k = 0..2 (strided 3x3 dgemm, bev_generator.py:227)
    double q0 = __fma_rn(bp.R[2], sz, __fma_rn(bp.R[1], sy, __dmul_rn(bp.R[0], sx)));
    double q1 = __fma_rn(bp.R[5], sz, __fma_rn(bp.R[4], sy, __dmul_rn(bp.R[3], sx)));
    double q2 = __fma_rn(bp.R[8], sz, __fma_rn(bp.R[7], sy, __dmul_rn(bp.R[6], sx)));
    q0 = __dadd_rn(q0, bp.trans_dx);
    q1 = __dadd_rn(q1, bp.trans_dy);
    e.z = q2;
    const double hv = __dmul_rn(0.5, bp.view);
    bool in = (q0 > -hv) && (q0 < hv) && (q1 > -hv) && (q1 < hv);
    bool hf_on = (bp.height_filter == bp.height_filter);
    if (hf_on) in = in && (q2 < bp.height_filter);
    double g0 = 0, g1 = 0;
    const double dP = (double)P, hP = __dmul_rn(0.5, dP);
    if (in || want_near) {
        g0 = __dadd_rn(__dmul_rn(__ddiv_rn(q0, bp.view), dP), hP);
        g1 = __dadd_rn(__dmul_rn(__ddiv_rn(q1, bp.view), dP), hP);
    }
    if (want_near) {
        const double eg = GUARD_M * dP / bp.view;
        bool n = (fabs(fabs(q0) - hv) < GUARD_M) || (fabs(fabs(q1) - hv) < GUARD_M);
        if (hf_on) n = n || (fabs(q2 - bp.height_filter) < GUARD_M);
        n = n || (fabs(g0 - rint(g0)) < eg) || (fabs(g1 - rint(g1)) < eg);
        // only points that are, or could become, part of the view matter
        e.near = n && (fabs(q0) < hv + GUARD_M) && (fabs(q1) < hv + GUARD_M);
    }
    if (in) {
        double fi = floor(g0), fj = floor(g1);
        if (fi >= 0.0 && fi < dP && fj >= 0.0 && fj < dP) {
            int i = (int)fi, j = (int)fj;
            e.cell = (P - 1 - j) * P + i;  // row = P-1-j, col = i (bev_generator.py:453)
            e.keep = true;
        }
    }
    return e;
}

__global__ void __launch_bounds__(BIN_BLOCK)
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

    // stage variant parameters and the frame's composed matrix
    {
        const uint32_t *src = (const uint32_t *)a.params;
        uint32_t *dst = (uint32_t *)s_par;
        const int words = a.n_var * (int)(sizeof(pcacc_bev_params) / 4);
        for (int k = threadIdx.x; k < words; k += BIN_BLOCK) dst[k] = src[k];
        if (threadIdx.x < 12) s_comp[threadIdx.x] = a.comp[(int64_t)slot * 12 + threadIdx.x];
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int PP = a.P * a.P;

    // load this thread's points once; they are reused by every variant
    double sx[BIN_ITEMS], sy[BIN_ITEMS], sz[BIN_ITEMS];   // source-frame coordinates
    double px[BIN_ITEMS], py[BIN_ITEMS], pz[BIN_ITEMS];   // current-frame coordinates
    bool valid[BIN_ITEMS];
#pragma unroll
    for (int k = 0; k < BIN_ITEMS; k++) {
        int64_t i = tile0 + k * BIN_BLOCK + threadIdx.x;
        valid[k] = i < cnt;
        if (valid[k]) {
            sx[k] = a.ring.x[off + i];
            sy[k] = a.ring.y[off + i];
            sz[k] = a.ring.z[off + i];
        } else {
            sx[k] = sy[k] = sz[k] = 0.0;
        }
        if (lazy) affine_chain(s_comp, 4, sx[k], sy[k], sz[k], px[k], py[k], pz[k]);
        else {
            px[k] = sx[k];
            py[k] = sy[k];
            pz[k] = sz[k];
        }
    }

    for (int v = 0; v < a.n_var; v++) {
        const pcacc_bev_params &bp = s_par[v];
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
            int64_t i = tile0 + k * BIN_BLOCK + threadIdx.x;
            if (valid[k]) {
                Eval e = eval_point(bp, a.P, px[k], py[k], pz[k], lazy);
                if (lazy && e.near) {
                    // exact sequential re-base chain (update_sem_pcs, sem_pc_accum.py:167-183)
                    double ex = sx[k], ey = sy[k], ez = sz[k];
                    for (int64_t ep = e0; ep < a.epoch_now; ep++) {
                        const double *T = a.chain + (ep % a.max_frames) * 12;
                        double nx, ny, nz;
                        affine_chain(T, 4, ex, ey, ez, nx, ny, nz);
                        ex = nx;
                        ey = ny;
                        ez = nz;
                    }
                    e = eval_point(bp, a.P, ex, ey, ez, false);
                    atomicAdd(a.n_replay, 1ull);
                }
                if (e.keep && a.ring.dyn[off + i] == 1) e.keep = false;  // static points only
                if (a.dbg_cell && v == 0) a.dbg_cell[off + i] = e.keep ? e.cell : -1;
                if (e.keep) {
                    keep[k] = true;
                    key[k] = ((uint32_t)v * (uint32_t)PP + (uint32_t)e.cell) * 2u + win;
                    rank[k] = atomicAdd(&a.counts[key[k]], 1u);
                    unsigned long long zb = (unsigned long long)__double_as_longlong(e.z);
                    rec[k].x = (uint32_t)zb;
                    rec[k].y = (uint32_t)(zb >> 32);
                    rec[k].z = a.ring.rgbs[off + i];
                    rec[k].w = __float_as_uint(a.ring.inten[off + i]);
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
struct CellStats {
    uint32_t n[2], n_road[2], n_veh[2];
    long long fx_hi[2], fx_lo[2];  // intensity sum of road points, 2^-40 fixed point, split at 2^32
    double ext_z[2];               // min (or max) z
    int med2[3][3];                // [window p/f/full][channel]: twice the median
};

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

template <bool F64OUT>
__global__ void __launch_bounds__(RED_WARPS * 32)
k_bev_reduce(const uint32_t *__restrict__ start, const uint4 *__restrict__ sorted,
             const pcacc_bev_params *__restrict__ params, int n_var, int P, double intensity_div,
             __half *__restrict__ out16, double *__restrict__ out64) {
    __shared__ __align__(16) uint32_t s_hist[RED_WARPS][2][3][256];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int PP = P * P;
    const int64_t n_cells = (int64_t)n_var * PP;
    const int64_t gw = (int64_t)blockIdx.x * RED_WARPS + warp;
    const int64_t cell0 = gw * 32;
    if (cell0 >= n_cells) return;
    const int var = (int)(cell0 / PP);
    const int cell_in = (int)(cell0 - (int64_t)var * PP) + (int)lane;
    const pcacc_bev_params bp = params[var];
    uint32_t(*hist)[3][256] = s_hist[warp];

    // segment bounds of my cell: [s0, s1) present, [s1, s2) future
    const int64_t gc = cell0 + lane;
    uint2 s01 = ((const uint2 *)start)[gc];
    uint32_t s2 = __shfl_down_sync(0xffffffffu, s01.x, 1);
    if (lane == 31) s2 = start[2 * gc + 2];
    const uint32_t my_np = s01.y - s01.x, my_nf = s2 - s01.y;

    CellStats st;
    st.n[0] = my_np;
    st.n[1] = my_nf;
#pragma unroll
    for (int w = 0; w < 2; w++) {
        st.n_road[w] = st.n_veh[w] = 0;
        st.fx_hi[w] = st.fx_lo[w] = 0;
        st.ext_z[w] = 0.0;
    }
#pragma unroll
    for (int w = 0; w < 3; w++)
        for (int c = 0; c < 3; c++) st.med2[w][c] = 0;

    const bool want_max = bp.elevation_max != 0;
    unsigned todo = __ballot_sync(0xffffffffu, (my_np + my_nf) > 0);
    while (todo) {
        const int owner = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t b0 = __shfl_sync(0xffffffffu, s01.x, owner);
        const uint32_t np = __shfl_sync(0xffffffffu, my_np, owner);
        const uint32_t nf = __shfl_sync(0xffffffffu, my_nf, owner);
        const uint32_t nt = np + nf;

        // clear the six histograms (1536 words)
        {
            uint4 *h4 = (uint4 *)&hist[0][0][0];
#pragma unroll
            for (int k = 0; k < 12; k++) h4[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();

        uint32_t a_road[2] = {0, 0}, a_veh[2] = {0, 0};
        long long a_hi[2] = {0, 0}, a_lo[2] = {0, 0};
        double a_z[2];
        a_z[0] = a_z[1] = want_max ? -INFINITY : INFINITY;
        for (uint32_t i = lane; i < nt; i += 32) {
            const uint4 r = sorted[b0 + i];
            const int w = i >= np ? 1 : 0;
            const double z =
                __longlong_as_double((long long)(((unsigned long long)r.y << 32) | r.x));
            const uint32_t c = r.z;
            const int sem = (int)(c >> 24);
            atomicAdd(&hist[w][0][c & 255u], 1u);
            atomicAdd(&hist[w][1][(c >> 8) & 255u], 1u);
            atomicAdd(&hist[w][2][(c >> 16) & 255u], 1u);
            const bool road = sem == bp.road_cls;
            const bool veh = (sem == bp.veh_cls[0]) || (sem == bp.veh_cls[1]) ||
                             (sem == bp.veh_cls[2]) || (sem == bp.veh_cls[3]);
            if (road) {
                double iv = __ddiv_rn((double)__uint_as_float(r.w), intensity_div);
                long long fx = __double2ll_rn(__dmul_rn(iv, FX_SCALE));
                a_hi[w] += fx >> 32;
                a_lo[w] += fx & 0xffffffffll;
                a_road[w]++;
            }
            if (veh) a_veh[w]++;
            a_z[w] = want_max ? fmax(a_z[w], z) : fmin(a_z[w], z);
        }
#pragma unroll
        for (int w = 0; w < 2; w++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a_road[w] += __shfl_xor_sync(0xffffffffu, a_road[w], o);
                a_veh[w] += __shfl_xor_sync(0xffffffffu, a_veh[w], o);
                a_hi[w] += __shfl_xor_sync(0xffffffffu, a_hi[w], o);
                a_lo[w] += __shfl_xor_sync(0xffffffffu, a_lo[w], o);
                double oz = __shfl_xor_sync(0xffffffffu, a_z[w], o);
                a_z[w] = want_max ? fmax(a_z[w], oz) : fmin(a_z[w], oz);
            }
        }
        __syncwarp();

        int med2[3][3];
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
                uint32_t s = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) s += c[j];
                uint32_t incl = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (unsigned)o) incl += t;
                }
                int m2 = 0;
                if (nw > 0) {  // warp-uniform
                    int lo = hist_kth(c, incl - s, incl, (nw - 1) / 2, lane);
                    int hi = (nw & 1u) ? lo : hist_kth(c, incl - s, incl, nw / 2, lane);
                    m2 = lo + hi;
                }
                med2[w][ch] = m2;
            }
        }
        if ((int)lane == owner) {
#pragma unroll
            for (int w = 0; w < 2; w++) {
                st.n_road[w] = a_road[w];
                st.n_veh[w] = a_veh[w];
                st.fx_hi[w] = a_hi[w];
                st.fx_lo[w] = a_lo[w];
                st.ext_z[w] = a_z[w];
            }
#pragma unroll
            for (int w = 0; w < 3; w++)
                for (int c = 0; c < 3; c++) st.med2[w][c] = med2[w][c];
        }
        __syncwarp();
    }

    // ---- finalise my cell: 3 windows x 7 planes ----------------------------
#pragma unroll
    for (int w = 0; w < 3; w++) {
        double n, n_road, n_veh, isum, ez;
        bool empty;
        if (w < 2) {
            n = (double)st.n[w];
            n_road = (double)st.n_road[w];
            n_veh = (double)st.n_veh[w];
            isum = fx_to_double(st.fx_hi[w], st.fx_lo[w]);
            empty = st.n[w] == 0;
            ez = empty ? 0.0 : st.ext_z[w];
        } else {
            n = (double)(st.n[0] + st.n[1]);
            n_road = (double)(st.n_road[0] + st.n_road[1]);
            n_veh = (double)(st.n_veh[0] + st.n_veh[1]);
            isum = fx_to_double(st.fx_hi[0] + st.fx_hi[1], st.fx_lo[0] + st.fx_lo[1]);
            empty = (st.n[0] + st.n[1]) == 0;
            if (st.n[0] == 0) ez = st.n[1] == 0 ? 0.0 : st.ext_z[1];
            else if (st.n[1] == 0) ez = st.ext_z[0];
            else ez = want_max ? fmax(st.ext_z[0], st.ext_z[1]) : fmin(st.ext_z[0], st.ext_z[1]);
        }
        double plane[7];
        // Dirichlet expectation, uniform prior (bev_generator.py:457-480)
        {
            double a = __dadd_rn(n_road, 1.0), b = __dadd_rn(__dsub_rn(n, n_road), 1.0);
            plane[0] = __ddiv_rn(a, __dadd_rn(a, b));
            double c = __dadd_rn(n_veh, 1.0), d = __dadd_rn(__dsub_rn(n, n_veh), 1.0);
            plane[5] = __ddiv_rn(c, __dadd_rn(c, d));
        }
        // intensity mean over (count+1) then road_marking_transform (sem_bev.py:593-617)
        {
            double I = __ddiv_rn(isum, __dadd_rn(n_road, 1.0));
            double t = __dmul_rn(bp.int_sep_scaler, __dsub_rn(I, bp.int_mid_threshold));
            double sg = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-t)));
            double val = __dmul_rn(bp.int_scaler, sg);
            plane[1] = val > 1.0 ? 1.0 : val;
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double med = empty ? bp.rgb_fill : __dmul_rn((double)st.med2[w][c], 0.5);
            plane[2 + c] = __ddiv_rn(med, 255.0);
        }
        plane[6] = ez;
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
            for (int64_t f0 = 0; f0 < nf; f0 += 65535) {
                int64_t ny = nf - f0 < 65535 ? nf - f0 : 65535;
                BinArgs b = a;
                b.frame_lo = flo + f0;
                dim3 grid((unsigned)((max_cnt + BIN_TILE - 1) / BIN_TILE), (unsigned)ny);
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
