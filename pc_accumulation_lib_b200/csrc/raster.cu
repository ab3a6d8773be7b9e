// raster.cu — BEV rasterisation of the resident ring (K7-K12 of SURVEY.md §2.1).
//
//   k_bev_cull     per (frame, variant): can the frame's bounding box touch the view?
//   k_bev_classify streaming pass over the ring: lazy re-base matrix, shift / rotate /
//                  translate, x / y crop -> block-aggregated candidate list
//   k_bev_bin      per candidate: exact re-base (guard band / chain replay), height and
//                  static filters, pos2grid -> cell key, cell counter (RED), 16 B record
//   k_scan         single-pass chained inclusive scan of the counters
//   k_bev_scatter  counting-sort placement: sorted[--cursor[key]] = record
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
#include <stddef.h>
#include <string.h>

#include "common.cuh"

#define BIN_BLOCK 256
#define BIN_ITEMS 4
#define BIN_TILE (BIN_BLOCK * BIN_ITEMS)
#define MAX_VGROUP 32

#define SCAN_BLOCK 512
#define SCAN_ITEMS 32
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)

// warps per block of the two reduce passes: small blocks, because the work per warp ranges
// from a handful of stores (32 empty cells) to thousands of instructions and a block's slot
// is held until its slowest warp is done (measured: 2 -> 78 us, 4 -> 84 us, 8 -> 100 us)
#define RED_WARPS 2
// pass B: a warp per queued cell with 6 KB of histograms; 4 warps per block share the
// dynamic queue's ticket traffic (2 -> 142 us, 4 -> 104 us on the long-horizon window)
#define REDB_WARPS 4

// guard band (metres) inside which a lazily re-based point is re-evaluated with
// the exact sequential chain; composed-vs-sequential error is < 1e-11 m for
// chains of hundreds of frames, so 1e-7 is four orders of margin
#define GUARD_M 1e-7

static_assert(sizeof(pcacc_bev_params) == 208, "pcacc_bev_params layout is part of the ABI");

// per-variant constants, computed once per rasterise call by k_bev_consts
struct BevConsts {
    double empty[7];    // planes of a window without any point
    __half empty_h[4];  // the same as float16: road, intensity, rgb (r = g = b), elevation
    double pv;          // P / view: grid_floor's estimate
};
__device__ __forceinline__ double bev_pv(const BevConsts *c, int v) { return c[v].pv; }

// per (frame of the launch, variant): the x / y part of  R (M p - origin) + trans  folded
// into one 2x4 affine map of the STORED coordinates (M = the frame's pending lazy matrix,
// identity when there is none).  Not the reference's rounding sequence: it only selects
// candidates (k_bev_classify, with a guard band) and feeds the float32 screening of
// k_bev_bin; every decision that matters is re-taken in exact arithmetic.
struct FrameVar {
    double A[8];
    float Af[8];
};

struct BinArgs {
    RingDev ring;
    const int64_t *frame_off, *frame_cnt, *frame_epoch;
    const double *comp, *chain;
    const double *cull;              // total transform of each frame since insertion
    const unsigned long long *aabb;  // order-encoded source-frame bounding boxes
    int max_frames;
    int64_t frame_lo;      // absolute id of the first frame of this launch
    int64_t epoch_now;
    const pcacc_bev_params *params;  // device, n_var entries
    const BevConsts *consts;         // device, n_var entries
    int n_var;
    int n_frames;          // frames covered by this launch (<= BIN_MAXF), first = frame_lo
    int P;
    uint32_t *counts;      // n_var * 2*P*P (+1)
    uint32_t *frame_tiles; // per frame of the launch: tiles to visit (k_bev_cull)
    FrameVar *fvar;        // n_frames x n_var (k_bev_cull)
    uint32_t *cand_gi;     // candidate list: ring position ...
    uint32_t *cand_meta;   // ... and variant | frame-in-launch << 8
    uint32_t *tmp_key;
    uint4 *tmp_rec;
    unsigned long long *n_append;  // device counter (candidates)
    unsigned long long *n_replay;
    int64_t cap;           // capacity of the candidate / tmp arrays
    int32_t *dbg_cell;     // optional
    uint32_t *flags;
};

struct Eval {
    bool keep, near;
    int cell;
    double z;
};

#define KEY_INVALID 0xffffffffu


// x / y part of bev_generator.py:224-231 for one point of the accumulator's current
// frame: origin shift (one subtract per component, kitti360_sem_pc_accum.py:193), rotation
// as an FMA chain over k = 0..2 (strided 3x3 dgemm), translation.  A zero coefficient
// contributes exactly nothing for finite z, so that link of the chain is skipped (zfree).
__device__ __forceinline__ void bev_xy(const pcacc_bev_params &bp, bool zfree, double px, double py,
                                       double pz, double &q0, double &q1) {
    const double sx = __dsub_rn(px, bp.origin[0]);
    const double sy = __dsub_rn(py, bp.origin[1]);
    q0 = __fma_rn(bp.R[1], sy, __dmul_rn(bp.R[0], sx));
    q1 = __fma_rn(bp.R[4], sy, __dmul_rn(bp.R[3], sx));
    if (!zfree) {
        const double sz = __dsub_rn(pz, bp.origin[2]);
        q0 = __fma_rn(bp.R[2], sz, q0);
        q1 = __fma_rn(bp.R[5], sz, q1);
    }
    q0 = __dadd_rn(q0, bp.trans_dx);
    q1 = __dadd_rn(q1, bp.trans_dy);
}

// The whole of bev_generator.py:224-255,737-747 for one point: crop (strict), height
// filter, pos2grid (div, mul, add separately rounded), row = P-1-j, col = i.
// want_near: also report whether any decision lies within GUARD_M of its boundary.
// floor of pos2grid's  q / V * P + 0.5 P  (div, mul, add separately rounded,
// bev_generator.py:737-747).  The fused estimate q * (P/V) + 0.5 P differs from the
// reference value by < 1e-12 for |g| <= 1e4, so its floor is the reference's floor unless
// it lies within 1e-9 of an integer — only then is the exact sequence evaluated.
__device__ __forceinline__ double grid_floor(double q, double view, double dP, double hP, double pv,
                                             double &g_est) {
    g_est = fma(q, pv, hP);
    const double f = floor(g_est), t = g_est - f;
    if (t > 1e-9 && t < 1.0 - 1e-9 && fabs(g_est) < 1e4) return f;
    return floor(__dadd_rn(__dmul_rn(__ddiv_rn(q, view), dP), hP));
}

__device__ __forceinline__ Eval eval_point(const pcacc_bev_params &bp, int P, double pv, double px,
                                           double py, double pz, bool want_near) {
    Eval e;
    e.keep = false;
    e.near = false;
    e.cell = -1;
    e.z = 0.0;
    const bool zfree = (bp.R[2] == 0.0) && (bp.R[5] == 0.0);
    // a non-finite z poisons x and y in the reference (0 * inf = NaN): the point is dropped
    if (zfree && !(fabs(pz) <= 1.7976931348623157e308)) return e;
    double q0, q1;
    bev_xy(bp, zfree, px, py, pz, q0, q1);
    const double sx = __dsub_rn(px, bp.origin[0]);
    const double sy = __dsub_rn(py, bp.origin[1]);
    const double sz = __dsub_rn(pz, bp.origin[2]);
    const double q2 = __fma_rn(bp.R[8], sz, __fma_rn(bp.R[7], sy, __dmul_rn(bp.R[6], sx)));
    e.z = q2;
    const double hv = __dmul_rn(0.5, bp.view);
    bool in = (q0 > -hv) && (q0 < hv) && (q1 > -hv) && (q1 < hv);
    const bool hf_on = (bp.height_filter == bp.height_filter);
    if (hf_on) in = in && (q2 < bp.height_filter);
    const double dP = (double)P, hP = __dmul_rn(0.5, dP);
    double g0, g1;  // estimates; the floors below are exact
    const double fi = grid_floor(q0, bp.view, dP, hP, pv, g0);
    const double fj = grid_floor(q1, bp.view, dP, hP, pv, g1);
    if (want_near) {
        const double eg = GUARD_M * pv;
        bool n = (fabs(fabs(q0) - hv) < GUARD_M) || (fabs(fabs(q1) - hv) < GUARD_M);
        if (hf_on) n = n || (fabs(q2 - bp.height_filter) < GUARD_M);
        n = n || (fabs(g0 - rint(g0)) < eg) || (fabs(g1 - rint(g1)) < eg);
        e.near = n;
    }
    if (in) {
        if (fi >= 0.0 && fi < dP && fj >= 0.0 && fj < dP) {
            const int i = (int)fi, j = (int)fj;
            e.cell = (P - 1 - j) * P + i;  // row = P-1-j, col = i (bev_generator.py:453)
            e.keep = true;
        }
    }
    return e;
}

#define BIN_MAXF 2048 /* frames per launch */

// Frame culling: can any point of the frame fall into the view of a variant?  The 8
// corners of the frame's bounding box (source frame, recorded at integrate time) go
// through the frame's total transform and the variant's shift / rotation; the image of
// the box is inside the bounding box of the transformed corners.  Conservative (margin
// far above fp64 rounding and above the lazy-vs-sequential chain difference); NaN / inf
// boxes are never culled.  One thread per (frame, variant); frame_tiles[] starts at 0.
__device__ void bev_write_consts(const pcacc_bev_params &bp, int P, BevConsts *__restrict__ out);

__global__ void k_bev_cull(BinArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_frames * a.n_var) return;
    const int f = t / a.n_var, v = t - f * a.n_var;
    const int64_t fid = a.frame_lo + f;
    const int slot = (int)(fid % a.max_frames);
    // Everything this thread may read is addressed by (f, v) alone: all of it is requested here, before
    // the first decision, so the loads travel together.  Taken where they are used — behind the window
    // test, the empty-frame test and the box test — they were a chain of four dependent DRAM round trips,
    // most of this small kernel's 8 us on every rasterise call's critical path.
    const pcacc_bev_params bp = a.params[v];
    const int64_t cnt = a.frame_cnt[slot];
    double Mp[12], M[12];
    unsigned long long bbw[6];
#pragma unroll
    for (int k = 0; k < 12; k++) {
        Mp[k] = a.comp[(int64_t)slot * 12 + k];   // identity unless lazily re-based
        M[k] = a.cull[(int64_t)slot * 12 + k];
    }
#pragma unroll
    for (int k = 0; k < 6; k++) bbw[k] = a.aabb[(int64_t)slot * 6 + k];
    if (f == 0) bev_write_consts(bp, a.P, (BevConsts *)a.consts + v);
    if (!(fid >= bp.frame_begin && fid < bp.frame_end)) return;
    const uint32_t tiles = (uint32_t)((cnt + BIN_TILE - 1) / BIN_TILE);
    if (tiles == 0) return;
    {
        FrameVar fv;
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double acc = bp.R[3 * r] * Mp[c] + bp.R[3 * r + 1] * Mp[4 + c] + bp.R[3 * r + 2] * Mp[8 + c];
                if (c == 3)
                    acc += (r ? bp.trans_dy : bp.trans_dx) -
                           (bp.R[3 * r] * bp.origin[0] + bp.R[3 * r + 1] * bp.origin[1] +
                            bp.R[3 * r + 2] * bp.origin[2]);
                fv.A[4 * r + c] = acc;
                fv.Af[4 * r + c] = (float)acc;
            }
        }
        a.fvar[(int64_t)f * a.n_var + v] = fv;
    }
    double lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = ord_decode(bbw[k]);
        hi[k] = ord_decode(bbw[3 + k]);
    }
    bool touch = true;
    if (lo[0] <= hi[0]) {  // a tracked, non-empty box (otherwise let the points decide)
        double q0lo = INFINITY, q0hi = -INFINITY, q1lo = INFINITY, q1hi = -INFINITY, amax = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const double x = (c & 1) ? hi[0] : lo[0], y = (c & 2) ? hi[1] : lo[1],
                         z = (c & 4) ? hi[2] : lo[2];
            const double cx = M[0] * x + M[1] * y + M[2] * z + M[3];
            const double cy = M[4] * x + M[5] * y + M[6] * z + M[7];
            const double cz = M[8] * x + M[9] * y + M[10] * z + M[11];
            amax = fmax(amax, fmax(fabs(cx), fmax(fabs(cy), fabs(cz))));
            const double sx = cx - bp.origin[0], sy = cy - bp.origin[1], sz = cz - bp.origin[2];
            const double q0 = bp.R[0] * sx + bp.R[1] * sy + bp.R[2] * sz + bp.trans_dx;
            const double q1 = bp.R[3] * sx + bp.R[4] * sy + bp.R[5] * sz + bp.trans_dy;
            q0lo = fmin(q0lo, q0);
            q0hi = fmax(q0hi, q0);
            q1lo = fmin(q1lo, q1);
            q1hi = fmax(q1hi, q1);
        }
        const double lim = 0.5 * bp.view + 1e-5 + 1e-9 * amax;
        const bool outside = (q0lo > lim) || (q0hi < -lim) || (q1lo > lim) || (q1hi < -lim);
        touch = !outside;  // also true when anything is NaN
    }
    if (touch) atomicMax(&a.frame_tiles[f], tiles);
}

// ---------------------------------------------------------------------------
// pass 1a — k_bev_classify: pure streaming.  Persistent blocks walk a (frame, tile,
// variant-group) work list that is derived on the device from the frame table (the
// launch never depends on counts the host has not fetched); each thread loads x, y
// (z only when it is needed) of 4 points, applies the lazy matrix, and tests the x / y
// crop of every variant.  Candidates are appended (block-aggregated: one global atomic
// per tile and variant) as (ring position, variant | frame).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BIN_BLOCK, 4)
k_bev_classify(BinArgs a) {
    __shared__ pcacc_bev_params s_par[MAX_VGROUP];
    __shared__ uint32_t s_tiles[BIN_MAXF + 1];
    __shared__ uint32_t s_warp[BIN_BLOCK / 32];
    __shared__ double s_A[MAX_VGROUP][8];
    __shared__ int s_vpb;
    __shared__ unsigned long long s_base;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

    // stage all variant parameters; build the exclusive prefix of tiles per frame
    {
        const uint32_t *src = (const uint32_t *)a.params;
        uint32_t *dst = (uint32_t *)s_par;
        const int words = a.n_var * (int)(sizeof(pcacc_bev_params) / 4);
        for (int k = threadIdx.x; k < words; k += BIN_BLOCK) dst[k] = src[k];
        constexpr int PER = BIN_MAXF / BIN_BLOCK;
        uint32_t loc[PER];
        uint32_t sum = 0;
        const bool mine = (int)threadIdx.x * PER <= a.n_frames;   // few threads own frames
#pragma unroll
        for (int k = 0; k < PER; k++) loc[k] = 0;
        if (mine) {
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int f = (int)threadIdx.x * PER + k;
                const uint32_t t = f < a.n_frames ? a.frame_tiles[f] : 0u;  // 0 = empty or culled
                loc[k] = sum;
                sum += t;
            }
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < BIN_BLOCK / 32; w++)
            if (w < (int)warp) wbase += s_warp[w];
        const uint32_t excl = wbase + incl - sum;
        if (mine) {
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int f = (int)threadIdx.x * PER + k;
                if (f <= a.n_frames) s_tiles[f] = excl + loc[k];
            }
        }
        __syncthreads();
    }
    const uint32_t total_tiles = s_tiles[a.n_frames];
    if (total_tiles == 0) return;
    // A work item = one tile x a group of vpb variants.  Few variants per item spread the work
    // over more blocks but re-read the tile for every group; the block with the most items sets
    // the kernel's duration, so pick the vpb that minimises  max items per block x (vpb + cost
    // of loading a tile, about three variant tests).
    if (threadIdx.x == 0) {
        uint32_t best = 0xffffffffu;
        int pick = a.n_var;
        for (int c = 1; c <= a.n_var; c = (c < a.n_var && 2 * c > a.n_var) ? a.n_var : 2 * c) {
            const uint32_t g = (uint32_t)((a.n_var + c - 1) / c);
            const uint32_t per_block = (total_tiles * g + gridDim.x - 1) / gridDim.x;
            const uint32_t cost = per_block * (uint32_t)(c + 3);
            if (cost < best) {
                best = cost;
                pick = c;
            }
            if (c == a.n_var) break;
        }
        s_vpb = pick;
    }
    __syncthreads();
    const int vpb = s_vpb;
    const int groups = (a.n_var + vpb - 1) / vpb;
    const uint32_t n_items = total_tiles * (uint32_t)groups;

    // The per-item bookkeeping is block-uniform: warp 0 does it once — which frame the tile
    // belongs to (every lane probes a slice of the tile prefix), the frame's ring offset and
    // count, whether z has to be streamed — and hands it over in shared memory.  Done by every
    // thread (a binary search over s_tiles, two 64-bit modulos, dependent loads of the frame
    // table) it was a third of this kernel's instructions.
    __shared__ struct {
        long long off, cnt, tile0, fid;
        int fl, v_begin, v_end, need_z;
    } s_item;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();   // the previous item's readers of s_item / s_A are done
        if (warp == 0) {
            const uint32_t grp = groups == 1 ? 0u : item / total_tiles;
            const uint32_t tile_lin = item - grp * total_tiles;
            int fl = 0;   // the frame f with s_tiles[f] <= tile_lin < s_tiles[f + 1] (tiles > 0 there)
            for (int f0 = 0; f0 < a.n_frames; f0 += 32) {
                const int f = f0 + (int)lane;
                const bool hit = f < a.n_frames && s_tiles[f] <= tile_lin && tile_lin < s_tiles[f + 1];
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    fl = f0 + __ffs(m) - 1;
                    break;
                }
            }
            const int vb = (int)grp * vpb, ve = min(a.n_var, vb + vpb);
            const long long fid = a.frame_lo + fl;
            int slot = (int)(a.frame_lo % a.max_frames) + fl;
            if (slot >= a.max_frames) slot -= a.max_frames;
            // z is only streamed when some variant's map mixes it into x / y (a lazily re-based
            // frame, or a rotation that is not about the z axis)
            bool nz = false;
            const FrameVar *fvw = a.fvar + (int64_t)fl * a.n_var;
            for (int v = vb + (int)lane; v < ve; v += 32) {
                const pcacc_bev_params &bp = s_par[v];
                if (fid >= bp.frame_begin && fid < bp.frame_end)
                    nz = nz || (fvw[v].A[2] != 0.0) || (fvw[v].A[6] != 0.0);
            }
            nz = __any_sync(0xffffffffu, nz);
            if (lane == 0) {
                s_item.off = a.frame_off[slot];
                s_item.cnt = a.frame_cnt[slot];
                s_item.tile0 = (long long)(tile_lin - s_tiles[fl]) * BIN_TILE;
                s_item.fid = fid;
                s_item.fl = fl;
                s_item.v_begin = vb;
                s_item.v_end = ve;
                s_item.need_z = nz ? 1 : 0;
            }
        }
        __syncthreads();
        const int fl = s_item.fl, v_begin = s_item.v_begin, v_end = s_item.v_end;
        const int64_t fid = s_item.fid, cnt = s_item.cnt, tile0 = s_item.tile0, off = s_item.off;
        const bool need_z = s_item.need_z != 0;
        const FrameVar *fv = a.fvar + (int64_t)fl * a.n_var;
        // the item's candidate maps go to shared memory in one cooperative load: read from
        // global inside the variant loop, each was a dependent L2 round trip ahead of its tests
        const bool staged = a.n_var > 1;   // a single variant reads its map directly
        if (staged) {
            for (int k = threadIdx.x; k < (v_end - v_begin) * 8; k += BIN_BLOCK)
                s_A[k >> 3][k & 7] = fv[v_begin + (k >> 3)].A[k & 7];
            __syncthreads();
        }

        // 4 points per thread: two pairs of neighbours (16 B loads; frame offsets are
        // multiples of 4 records, so the pairs are aligned)
        double px[BIN_ITEMS], py[BIN_ITEMS], pz[BIN_ITEMS];
        unsigned vmask = 0;
#pragma unroll
        for (int h = 0; h < BIN_ITEMS / 2; h++) {
            const int64_t i0 = tile0 + (int64_t)h * (2 * BIN_BLOCK) + 2 * threadIdx.x;
            double2 X = make_double2(0, 0), Y = make_double2(0, 0), Z = make_double2(0, 0);
            if (i0 < cnt) {
                vmask |= 1u << (2 * h);
                if (i0 + 1 < cnt) vmask |= 2u << (2 * h);
                X = *(const double2 *)(a.ring.x + off + i0);
                Y = *(const double2 *)(a.ring.y + off + i0);
                if (need_z) Z = *(const double2 *)(a.ring.z + off + i0);
            }
            px[2 * h] = X.x; px[2 * h + 1] = X.y;
            py[2 * h] = Y.x; py[2 * h + 1] = Y.y;
            pz[2 * h] = Z.x; pz[2 * h + 1] = Z.y;
        }

        for (int v = v_begin; v < v_end; v++) {
            const pcacc_bev_params &bp = s_par[v];
            if (!(fid >= bp.frame_begin && fid < bp.frame_end)) continue;  // block-uniform
            // candidate = within the view plus a guard band; k_bev_bin takes the real decision
            const double lim = __dmul_rn(0.5, bp.view) + GUARD_M;
            double A[8];
#pragma unroll
            for (int k = 0; k < 8; k++) A[k] = staged ? s_A[v - v_begin][k] : fv[v].A[k];
            unsigned cmask = 0;
#pragma unroll
            for (int k = 0; k < BIN_ITEMS; k++) {
                const double q0 = fma(A[0], px[k], fma(A[1], py[k], fma(A[2], pz[k], A[3])));
                const double q1 = fma(A[4], px[k], fma(A[5], py[k], fma(A[6], pz[k], A[7])));
                if ((fabs(q0) < lim) && (fabs(q1) < lim)) cmask |= 1u << k;  // NaN fails
            }
            cmask &= vmask;
            const uint32_t my_cnt = __popc(cmask);
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
                const uint32_t c = s_warp[w];
                if (w < (int)warp) wbase += c;
                btotal += c;
            }
            if (threadIdx.x == 0)
                s_base = btotal ? atomicAdd(a.n_append, (unsigned long long)btotal) : 0ull;
            __syncthreads();
            if (cmask) {
                unsigned long long pos = s_base + wbase + (incl - my_cnt);
                const uint32_t meta = (uint32_t)v | ((uint32_t)fl << 8);
#pragma unroll
                for (int k = 0; k < BIN_ITEMS; k++) {
                    if ((cmask >> k) & 1u) {
                        if ((int64_t)pos < a.cap) {
                            a.cand_gi[pos] = (uint32_t)(off + tile0 + (int64_t)(k >> 1) * (2 * BIN_BLOCK) +
                                                        2 * threadIdx.x + (k & 1));
                            a.cand_meta[pos] = meta;
                        }
                        pos++;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// pass 1a for batches of variants (n_var >= CLS_MV_MIN) — k_bev_classify_mv.  k_bev_classify lets
// the whole block cooperate on one variant at a time: a block-wide scan, a returning atomic and two
// barriers per (tile, variant), i.e. a chain of ~1 us round trips per variant.  Here a work item is
// one tile x CLS_MV_W variants: the block stages the tile's x / y (/ z) in shared memory once, then
// every warp takes ONE variant and works alone — it counts its candidates over the 1024 points,
// reserves their slots with one atomic, and walks the tile a second time to write them (the test is
// a handful of FMAs; repeating it is cheaper than keeping 32 masks per lane).  No barrier inside an
// item, eight independent chains per block.
// ---------------------------------------------------------------------------
#define CLS_MV_W (BIN_BLOCK / 32)   /* variants per item = warps per block */
#define CLS_MV_MIN 8

__global__ void __launch_bounds__(BIN_BLOCK, 4)
k_bev_classify_mv(BinArgs a) {
    __shared__ pcacc_bev_params s_par[MAX_VGROUP];
    __shared__ uint32_t s_tiles[BIN_MAXF + 1];
    __shared__ uint32_t s_warp[BIN_BLOCK / 32];
    __shared__ __align__(16) double s_x[BIN_TILE], s_y[BIN_TILE], s_z[BIN_TILE];
    __shared__ struct {
        long long off, cnt, tile0, fid;
        int fl, v_begin, v_end, need_z;
    } s_item;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    {   // stage all variant parameters; exclusive prefix of tiles per frame (as k_bev_classify)
        const uint32_t *src = (const uint32_t *)a.params;
        uint32_t *dst = (uint32_t *)s_par;
        const int words = a.n_var * (int)(sizeof(pcacc_bev_params) / 4);
        for (int k = threadIdx.x; k < words; k += BIN_BLOCK) dst[k] = src[k];
        constexpr int PER = BIN_MAXF / BIN_BLOCK;
        uint32_t loc[PER];
        uint32_t sum = 0;
        const bool mine = (int)threadIdx.x * PER <= a.n_frames;
#pragma unroll
        for (int k = 0; k < PER; k++) loc[k] = 0;
        if (mine) {
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int f = (int)threadIdx.x * PER + k;
                const uint32_t t = f < a.n_frames ? a.frame_tiles[f] : 0u;
                loc[k] = sum;
                sum += t;
            }
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < BIN_BLOCK / 32; w++)
            if (w < (int)warp) wbase += s_warp[w];
        const uint32_t excl = wbase + incl - sum;
        if (mine) {
#pragma unroll
            for (int k = 0; k < PER; k++) {
                const int f = (int)threadIdx.x * PER + k;
                if (f <= a.n_frames) s_tiles[f] = excl + loc[k];
            }
        }
        __syncthreads();
    }
    const uint32_t total_tiles = s_tiles[a.n_frames];
    if (total_tiles == 0) return;
    const uint32_t groups = (uint32_t)((a.n_var + CLS_MV_W - 1) / CLS_MV_W);
    const uint32_t n_items = total_tiles * groups;

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();   // the previous item's readers of s_item / the tile are done
        if (warp == 0) {
            // consecutive items share the tile (its lines stay in L2 for the other groups)
            const uint32_t tile_lin = item / groups, grp = item - tile_lin * groups;
            int fl = 0;
            for (int f0 = 0; f0 < a.n_frames; f0 += 32) {
                const int f = f0 + (int)lane;
                const bool hit = f < a.n_frames && s_tiles[f] <= tile_lin && tile_lin < s_tiles[f + 1];
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    fl = f0 + __ffs(m) - 1;
                    break;
                }
            }
            const int vb = (int)grp * CLS_MV_W, ve = min(a.n_var, vb + CLS_MV_W);
            const long long fid = a.frame_lo + fl;
            int slot = (int)(a.frame_lo % a.max_frames) + fl;
            if (slot >= a.max_frames) slot -= a.max_frames;
            bool nz = false;
            const FrameVar *fvw = a.fvar + (int64_t)fl * a.n_var;
            for (int v = vb + (int)lane; v < ve; v += 32) {
                const pcacc_bev_params &bp = s_par[v];
                if (fid >= bp.frame_begin && fid < bp.frame_end)
                    nz = nz || (fvw[v].A[2] != 0.0) || (fvw[v].A[6] != 0.0);
            }
            nz = __any_sync(0xffffffffu, nz);
            if (lane == 0) {
                s_item.off = a.frame_off[slot];
                s_item.cnt = a.frame_cnt[slot];
                s_item.tile0 = (long long)(tile_lin - s_tiles[fl]) * BIN_TILE;
                s_item.fid = fid;
                s_item.fl = fl;
                s_item.v_begin = vb;
                s_item.v_end = ve;
                s_item.need_z = nz ? 1 : 0;
            }
        }
        __syncthreads();
        const int fl = s_item.fl;
        const int64_t fid = s_item.fid, cnt = s_item.cnt, tile0 = s_item.tile0, off = s_item.off;
        const bool need_z = s_item.need_z != 0;
        const int n_here = (int)min((int64_t)BIN_TILE, cnt - tile0);   // points of this tile (>= 1)
        {   // the tile: pairs of neighbours, 16 B loads (frame offsets are multiples of 4 records)
#pragma unroll
            for (int h = 0; h < BIN_ITEMS / 2; h++) {
                const int i0 = h * (2 * BIN_BLOCK) + 2 * (int)threadIdx.x;
                double2 X = make_double2(0, 0), Y = make_double2(0, 0), Z = make_double2(0, 0);
                if (i0 < n_here) {
                    X = *(const double2 *)(a.ring.x + off + tile0 + i0);
                    Y = *(const double2 *)(a.ring.y + off + tile0 + i0);
                    if (need_z) Z = *(const double2 *)(a.ring.z + off + tile0 + i0);
                }
                *(double2 *)&s_x[i0] = X;
                *(double2 *)&s_y[i0] = Y;
                *(double2 *)&s_z[i0] = Z;
            }
        }
        __syncthreads();
        const int v = s_item.v_begin + (int)warp;
        if (v >= s_item.v_end) continue;
        const pcacc_bev_params &bp = s_par[v];
        if (!(fid >= bp.frame_begin && fid < bp.frame_end)) continue;   // warp-uniform
        // candidate = within the view plus a guard band; k_bev_bin takes the real decision
        const double lim = __dmul_rn(0.5, bp.view) + GUARD_M;
        double A[8];
        {
            const double *Ag = a.fvar[(int64_t)fl * a.n_var + v].A;
#pragma unroll
            for (int k = 0; k < 8; k++) A[k] = Ag[k];
        }
        const int n_round = (n_here + 31) >> 5;
        uint32_t total = 0;
        for (int r = 0; r < n_round; r++) {
            const int i = (r << 5) + (int)lane;
            const double q0 = fma(A[0], s_x[i], fma(A[1], s_y[i], fma(A[2], s_z[i], A[3])));
            const double q1 = fma(A[4], s_x[i], fma(A[5], s_y[i], fma(A[6], s_z[i], A[7])));
            const bool c = i < n_here && (fabs(q0) < lim) && (fabs(q1) < lim);   // NaN fails
            total += (uint32_t)__popc(__ballot_sync(0xffffffffu, c));
        }
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.n_append, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t meta = (uint32_t)v | ((uint32_t)fl << 8);
        const uint32_t gi0 = (uint32_t)(off + tile0);
        for (int r = 0; r < n_round; r++) {
            const int i = (r << 5) + (int)lane;
            const double q0 = fma(A[0], s_x[i], fma(A[1], s_y[i], fma(A[2], s_z[i], A[3])));
            const double q1 = fma(A[4], s_x[i], fma(A[5], s_y[i], fma(A[6], s_z[i], A[7])));
            const bool c = i < n_here && (fabs(q0) < lim) && (fabs(q1) < lim);
            const unsigned m = __ballot_sync(0xffffffffu, c);
            if (c) {
                const unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
                if ((int64_t)pos < a.cap) {
                    a.cand_gi[pos] = gi0 + (uint32_t)i;
                    a.cand_meta[pos] = meta;
                }
            }
            base += (unsigned long long)__popc(m);
        }
    }
}

// ---------------------------------------------------------------------------
// pass 1b — k_bev_bin: one thread per candidate, no block synchronisation: exact
// per-point work (lazy matrix + guard band / chain replay, height filter, pos2grid,
// static filter), the cell counter (its old value is the rank in the cell segment) and
// the 16 B record.  A candidate rejected here leaves a hole (KEY_INVALID).
// ---------------------------------------------------------------------------
// Exact evaluation of one candidate: lazy matrix + guard band, exact chain replay when a
// decision is within GUARD_M of its boundary (update_sem_pcs, sem_pc_accum.py:167-183).
__device__ __forceinline__ Eval exact_candidate(const BinArgs &a, const pcacc_bev_params &bp, double pv,
                                             int slot, int64_t e0, double x, double y, double z) {
    Eval e;
    if (e0 < a.epoch_now) {
        double cx, cy, cz;
        affine_chain(a.comp + (int64_t)slot * 12, 4, x, y, z, cx, cy, cz);
        e = eval_point(bp, a.P, pv, cx, cy, cz, true);
        if (e.near) {
            for (int64_t ep = e0; ep < a.epoch_now; ep++) {
                const double *T = a.chain + (ep % a.max_frames) * 12;
                double nx, ny, nz;
                affine_chain(T, 4, x, y, z, nx, ny, nz);
                x = nx; y = ny; z = nz;
            }
            e = eval_point(bp, a.P, pv, x, y, z, false);
            atomicAdd(a.n_replay, 1ull);
        }
    } else {
        e = eval_point(bp, a.P, pv, x, y, z, false);
    }
    return e;
}

#ifndef BIN_OCC
#define BIN_OCC 3
#endif
__global__ void __launch_bounds__(256, BIN_OCC)
k_bev_bin(BinArgs a) {
    __shared__ pcacc_bev_params s_par[MAX_VGROUP];
    __shared__ double s_pv[MAX_VGROUP];
    {
        const uint32_t *src = (const uint32_t *)a.params;
        uint32_t *dst = (uint32_t *)s_par;
        const int words = a.n_var * (int)(sizeof(pcacc_bev_params) / 4);
        for (int k = threadIdx.x; k < words; k += 256) dst[k] = src[k];
        if ((int)threadIdx.x < a.n_var) s_pv[threadIdx.x] = bev_pv(a.consts, threadIdx.x);
        __syncthreads();
    }
    unsigned long long n = *a.n_append;
    if (n > (unsigned long long)a.cap) n = (unsigned long long)a.cap;
    const int PP = a.P * a.P;
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    const int slot_lo = (int)(a.frame_lo % a.max_frames);

    // exact evaluation of one candidate whose point data is already in registers
    auto process = [&](unsigned long long c, uint32_t gi32, uint32_t meta, double x, double y, double z,
                       uint32_t rgbs, float inten, uint32_t dyn) {
        const int64_t gi = (int64_t)gi32;
        const int v = (int)(meta & 255u);
        const int64_t fid = a.frame_lo + (int64_t)(meta >> 8);
        // (frame_lo + f) % max_frames without a 64-bit division per candidate (f < max_frames)
        int slot = slot_lo + (int)(meta >> 8);
        if (slot >= a.max_frames) slot -= a.max_frames;
        const pcacc_bev_params &bp = s_par[v];
        const double pv = s_pv[v];
        const int64_t e0 = a.frame_epoch[slot];
        Eval e;
        // ---- float32 screening -------------------------------------------------------------
        // The composed map of the stored coordinates, evaluated in float32 with a running
        // bound of its own rounding error.  When the point is farther than that bound (plus
        // 1e-5 m) from the crop boundary and from every cell boundary, the cell index is
        // decided here and only z is computed in float64; everything else falls through to
        // the exact path below.  err: 16 float ulps of the sum of the term magnitudes covers
        // the roundings of A -> float, p -> float and the three FMAs.
        if (bp.R[6] == 0.0 && bp.R[7] == 0.0 && bp.R[8] == 1.0) {
            // the float32 map of this (frame, variant): two 16-byte loads (32-bit index math:
            // frames x variants of a launch is far below 2^31)
            const float4 *af4 = (const float4 *)a.fvar[(meta >> 8) * (uint32_t)a.n_var + (uint32_t)v].Af;
            const float4 r0 = af4[0], r1 = af4[1];
            const float xf = (float)x, yf = (float)y, zf = (float)z;
            const float q0 = fmaf(r0.x, xf, fmaf(r0.y, yf, fmaf(r0.z, zf, r0.w)));
            const float q1 = fmaf(r1.x, xf, fmaf(r1.y, yf, fmaf(r1.z, zf, r1.w)));
            const float viewf = (float)bp.view;
            const float m0 = 1e-5f + 9.6e-7f * (fabsf(r0.x * xf) + fabsf(r0.y * yf) + fabsf(r0.z * zf) +
                                               fabsf(r0.w) + viewf);
            const float m1 = 1e-5f + 9.6e-7f * (fabsf(r1.x * xf) + fabsf(r1.y * yf) + fabsf(r1.z * zf) +
                                               fabsf(r1.w) + viewf);
            const float hvf = 0.5f * viewf;
            const float d0 = hvf - fabsf(q0), d1 = hvf - fabsf(q1);   // > 0 inside the view
            if (d0 < -m0 || d1 < -m1) {
                // certainly outside the view (the guard band let it through): not binned
                if (a.dbg_cell && v == 0) a.dbg_cell[gi] = -1;
                a.tmp_key[c] = KEY_INVALID;
                return;
            }
            if (d0 > m0 && d1 > m1) {
                const float pvf = (float)pv, hPf = 0.5f * (float)a.P;
                const float g0 = fmaf(q0, pvf, hPf), g1 = fmaf(q1, pvf, hPf);
                const float f0 = floorf(g0), f1 = floorf(g1);
                const float e0g = m0 * pvf + 4.8e-7f * (fabsf(g0) + hPf + 1.f), e1g = m1 * pvf + 4.8e-7f * (fabsf(g1) + hPf + 1.f);
                const float t0 = g0 - f0, t1 = g1 - f1;
                if (t0 > e0g && t0 < 1.f - e0g && t1 > e1g && t1 < 1.f - e1g && f0 >= 0.f &&
                    f0 < (float)a.P && f1 >= 0.f && f1 < (float)a.P) {
                    // z in the reference's arithmetic: third row of the lazy matrix, then the shift
                    double cz = z;
                    if (e0 < a.epoch_now) {
                        const double *m2 = a.comp + (int64_t)slot * 12 + 8;
                        cz = __fma_rn(m2[3], 1.0, __fma_rn(m2[2], z, __fma_rn(m2[1], y, __dmul_rn(m2[0], x))));
                    }
                    // + 0.0: the exact path's  fma(1, sz, +0)  turns -0.0 into +0.0
                    const double q2 = __dadd_rn(__dsub_rn(cz, bp.origin[2]), 0.0);
                    const bool hf_on = (bp.height_filter == bp.height_filter);
                    // q2 == q2 rejects a non-finite z like the exact path does
                    if (q2 == q2 && fabs(q2) <= 1.7976931348623157e308 &&
                        (!hf_on || fabs(q2 - bp.height_filter) > 1e-5)) {
                        const bool keep = (dyn != 1u) && (!hf_on || q2 < bp.height_filter);
                        const int cell = (a.P - 1 - (int)f1) * a.P + (int)f0;
                        if (a.dbg_cell && v == 0) a.dbg_cell[gi] = keep ? cell : -1;
                        if (keep) {
                            const uint32_t win = fid >= bp.frame_split ? 1u : 0u;
                            const uint32_t key = ((uint32_t)v * (uint32_t)PP + (uint32_t)cell) * 2u + win;
                            const unsigned long long zb = (unsigned long long)__double_as_longlong(q2);
                            a.tmp_key[c] = key;
                            a.tmp_rec[c] = make_uint4((uint32_t)zb, (uint32_t)(zb >> 32), rgbs, __float_as_uint(inten));
                            atomicAdd(&a.counts[key], 1u);
                        } else {
                            a.tmp_key[c] = KEY_INVALID;
                        }
                        return;
                    }
                }
            }
        }
        // ---- exact path (rare) -------------------------------------------------------------
        e = exact_candidate(a, bp, pv, slot, e0, x, y, z);
        if (dyn == 1u) e.keep = false;  // static points only (sem_bev.py:54-58)
        if (a.dbg_cell && v == 0) a.dbg_cell[gi] = e.keep ? e.cell : -1;
        if (e.keep) {
            const uint32_t win = fid >= bp.frame_split ? 1u : 0u;
            const uint32_t key = ((uint32_t)v * (uint32_t)PP + (uint32_t)e.cell) * 2u + win;
            const unsigned long long zb = (unsigned long long)__double_as_longlong(e.z);
            a.tmp_key[c] = key;
            a.tmp_rec[c] = make_uint4((uint32_t)zb, (uint32_t)(zb >> 32), rgbs, __float_as_uint(inten));
            atomicAdd(&a.counts[key], 1u);  // result unused: a fire-and-forget reduction
        } else {
            a.tmp_key[c] = KEY_INVALID;  // hole: rejected by the exact test
        }
    };

    // Software pipeline over the thread's candidates c, c + stride, ...: while candidate k is
    // evaluated, the gathers of k + 1 and the list entry of k + 2 are in flight (the list entry is the
    // address of the gathers, the gathers feed the evaluation: taken one after the other, each
    // candidate paid both round trips).
    struct Pt { double x, y, z; uint32_t rgbs, dyn; float inten; };
    auto gather = [&](uint32_t g) {
        Pt p;
        p.x = a.ring.x[g]; p.y = a.ring.y[g]; p.z = a.ring.z[g];
        p.rgbs = a.ring.rgbs[g]; p.inten = a.ring.inten[g]; p.dyn = a.ring.dyn[g];
        return p;
    };
    unsigned long long c = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= n) return;
    uint32_t g0 = a.cand_gi[c], m0 = a.cand_meta[c];
    uint32_t g1 = g0, m1 = m0;
    if (c + stride < n) { g1 = a.cand_gi[c + stride]; m1 = a.cand_meta[c + stride]; }
    Pt p0 = gather(g0);
    for (; c < n; c += stride) {
        const Pt p1 = gather(g1);                       // valid address even past the end (a repeat)
        uint32_t g2 = g1, m2 = m1;
        if (c + 2 * stride < n) { g2 = a.cand_gi[c + 2 * stride]; m2 = a.cand_meta[c + 2 * stride]; }
        process(c, g0, m0, p0.x, p0.y, p0.z, p0.rgbs, p0.inten, p0.dyn);
        p0 = p1; g0 = g1; m0 = m1; g1 = g2; m1 = m2;
    }
}

// ---------------------------------------------------------------------------
// inclusive scan of n u32 counters, in place (chained single pass)
// ---------------------------------------------------------------------------
struct ScanLB {
    unsigned long long *state;
    uint32_t *ticket;
    uint32_t epoch, n_tiles;
};

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan(uint32_t *__restrict__ data, int64_t n, ScanLB lb) {
    // Warp w of the block owns 1024 consecutive counters; lane l holds the 16-byte pieces
    // k*32 + l (k = 0..7) of them, so every load and store of a warp is one contiguous
    // 512-byte run.  Order of the elements: (k, lane, component).
    __shared__ uint32_t s_warp[SCAN_BLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    constexpr int NK = SCAN_ITEMS / 4;
    const int64_t wbase = (int64_t)tile * SCAN_TILE + (int64_t)warp * (32 * SCAN_ITEMS);
    uint4 v[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int64_t e = wbase + ((int64_t)k * 32 + lane) * 4;
        if (e + 4 <= n) {
            v[k] = *(const uint4 *)(data + e);
        } else {
            v[k].x = e < n ? data[e] : 0u;
            v[k].y = e + 1 < n ? data[e + 1] : 0u;
            v[k].z = e + 2 < n ? data[e + 2] : 0u;
            v[k].w = 0u;
        }
    }
    // exclusive offset of every piece inside the warp's run
    uint32_t off[NK], carry = 0;
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const uint32_t sum = v[k].x + v[k].y + v[k].z + v[k].w;
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        off[k] = carry + incl - sum;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_warp[warp] = carry;   // the warp's total
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
    const uint32_t wex = s_warp[warp];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        // inclusive: the scatter counts each segment down to its start
        uint4 o;
        o.x = wex + off[k] + v[k].x;
        o.y = o.x + v[k].y;
        o.z = o.y + v[k].z;
        o.w = o.z + v[k].w;
        const int64_t e = wbase + ((int64_t)k * 32 + lane) * 4;
        if (e + 4 <= n) {
            *(uint4 *)(data + e) = o;
        } else {
            if (e < n) data[e] = o.x;
            if (e + 1 < n) data[e + 1] = o.y;
            if (e + 2 < n) data[e + 2] = o.z;
        }
    }
}

// ---------------------------------------------------------------------------
// scatter
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bev_scatter(uint32_t *__restrict__ cursor, const uint32_t *__restrict__ tmp_key,
              const uint4 *__restrict__ tmp_rec, const unsigned long long *__restrict__ n_append,
              int64_t cap, uint4 *__restrict__ sorted) {
    // cursor[key] enters as the inclusive prefix (= end of the key's segment); every record
    // takes the slot below it, so the array leaves as the segments' starts — which is what
    // the reduction reads.  The order inside a segment is arbitrary (see the file header).
    unsigned long long n = *n_append;
    if (n > (unsigned long long)cap) n = (unsigned long long)cap;
    // Records arrive in point order, and neighbouring lidar returns often fall into the same
    // cell: the lanes of a warp that hold the same key take their slots with ONE atomic (the
    // cursors of crowded cells are the serial bottleneck of this kernel).
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    const unsigned long long i_first = (unsigned long long)blockIdx.x * 256 + (threadIdx.x & ~31u);
    for (unsigned long long iw = i_first; iw < n; iw += stride) {   // warp-uniform trip count
        const unsigned long long i = iw + lane;
        const uint32_t key = i < n ? tmp_key[i] : KEY_INVALID;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (key != KEY_INVALID) rec = tmp_rec[i];
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        const uint32_t cnt = (uint32_t)__popc(peers), rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if ((int)lane == leader && key != KEY_INVALID) base = atomicSub(&cursor[key], cnt) - cnt;
        base = __shfl_sync(0xffffffffu, base, leader);
        if (key != KEY_INVALID) sorted[base + rank] = rec;
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

// Per-window statistics of one cell (everything except the medians).
struct WinAcc {
    uint32_t n_road[2], n_veh[2];
    long long fx_hi[2], fx_lo[2];   // sum of raw road intensities, 2^-40 fixed point split at 2^32
    double ext_z[2];
};

__device__ __forceinline__ void acc_init(WinAcc &a, bool want_max) {
#pragma unroll
    for (int w = 0; w < 2; w++) {
        a.n_road[w] = a.n_veh[w] = 0;
        a.fx_hi[w] = a.fx_lo[w] = 0;
        a.ext_z[w] = want_max ? -INFINITY : INFINITY;
    }
}

// w selects the window with conditional moves, not with an index: indexed, the accumulators lived in
// local memory and every record paid a load-modify-store round trip through L1 per field.
__device__ __forceinline__ void acc_record(WinAcc &a, const uint4 &r, int w, int road_cls, int v0,
                                           int v1, int v2, int v3, bool want_max) {
    const double z = __longlong_as_double((long long)(((unsigned long long)r.y << 32) | r.x));
    const int sem = (int)(r.z >> 24);
    const bool w1 = w != 0;
    if (sem == road_cls) {
        // raw float32 intensity: exact in 2^-40 fixed point for |v| >= 2^-17
        const long long fx = __double2ll_rn(__dmul_rn((double)__uint_as_float(r.w), FX_SCALE));
        const long long hi = fx >> 32, lo = fx & 0xffffffffll;
        a.fx_hi[0] += w1 ? 0ll : hi;
        a.fx_hi[1] += w1 ? hi : 0ll;
        a.fx_lo[0] += w1 ? 0ll : lo;
        a.fx_lo[1] += w1 ? lo : 0ll;
        a.n_road[0] += w1 ? 0u : 1u;
        a.n_road[1] += w1 ? 1u : 0u;
    }
    if ((sem == v0) || (sem == v1) || (sem == v2) || (sem == v3)) {
        a.n_veh[0] += w1 ? 0u : 1u;
        a.n_veh[1] += w1 ? 1u : 0u;
    }
    const double e0 = want_max ? fmax(a.ext_z[0], z) : fmin(a.ext_z[0], z);
    const double e1 = want_max ? fmax(a.ext_z[1], z) : fmin(a.ext_z[1], z);
    a.ext_z[0] = w1 ? a.ext_z[0] : e0;
    a.ext_z[1] = w1 ? e1 : a.ext_z[1];
}

// (m2 * 0.5) / 255. for m2 = twice the median, 0..510: filled once per handle
#define RGB_LUT_N 511
// followed by the Dirichlet expectation (c+1)/(n+2) for the small cells, n, c = 0..15
#define DIR_LUT_OFF 512
#define LUT_TOTAL (DIR_LUT_OFF + 256)

// road and dynamic planes of one window: Dirichlet expectation with a uniform prior,
// (c+1) / ((c+1) + ((n-c)+1)) (bev_generator.py:457-480); the denominator is n+2 exactly
__device__ __forceinline__ double dirichlet(uint32_t c, uint32_t n) {
    return __ddiv_rn((double)(c + 1u), (double)(n + 2u));
}

// intensity plane: mean over (count+1) then road_marking_transform (sem_bev.py:593-617).
// raw_sum = sum of the raw road intensities; the reference divides every intensity by
// intensity_div (255 on nuScenes) before it sums: sum(raw) / (div * (n+1)) rounds once
// instead of n+1 times (div * (n+1) is exact), DESIGN.md §6.
__device__ __forceinline__ double intensity_plane(const pcacc_bev_params &bp, double raw_sum, uint32_t n_road,
                                                  double intensity_div) {
    double I = __ddiv_rn(raw_sum, __dmul_rn(intensity_div, (double)(n_road + 1u)));
    double t = __dmul_rn(bp.int_sep_scaler, __dsub_rn(I, bp.int_mid_threshold));
    double sg = __ddiv_rn(1.0, __dadd_rn(1.0, exp(-t)));
    double val = __dmul_rn(bp.int_scaler, sg);
    return val > 1.0 ? 1.0 : val;
}

// per-variant constants (empty-window planes, P/view)
__device__ void bev_write_consts(const pcacc_bev_params &bp, int P, BevConsts *__restrict__ out) {
    BevConsts c;
    c.empty[0] = dirichlet(0, 0);
    c.empty[1] = intensity_plane(bp, 0.0, 0, 1.0);
    c.empty[2] = c.empty[3] = c.empty[4] = __ddiv_rn(bp.rgb_fill, 255.0);
    c.empty[5] = dirichlet(0, 0);
    c.empty[6] = 0.0;
    c.empty_h[0] = __double2half(c.empty[0]);
    c.empty_h[1] = __double2half(c.empty[1]);
    c.empty_h[2] = __double2half(c.empty[2]);
    c.empty_h[3] = __double2half(c.empty[6]);
    c.pv = (double)P / bp.view;
    *out = c;
}

// only launched when no frame is visited (k_bev_cull writes the constants otherwise)
__global__ void k_bev_consts(const pcacc_bev_params *__restrict__ params, int n_var, int P,
                             BevConsts *__restrict__ consts) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_var) return;
    bev_write_consts(params[v], P, consts + v);
}

__global__ void k_rgb_lut(double *__restrict__ lut) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < RGB_LUT_N) lut[m] = __ddiv_rn(__dmul_rn((double)m, 0.5), 255.0);
    else if (m == RGB_LUT_N) lut[m] = 0.0;
    else if (m < LUT_TOTAL) {
        const uint32_t n = (uint32_t)(m - DIR_LUT_OFF) >> 4, c = (uint32_t)(m - DIR_LUT_OFF) & 15u;
        lut[m] = dirichlet(c, n);
    }
}
// the same table rounded to fp16 behind the doubles (the fp16-only output reads these)
__global__ void k_rgb_lut16(double *__restrict__ lut) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < LUT_TOTAL) ((__half *)(lut + LUT_TOTAL))[m] = __double2half(lut[m]);
}

// window statistics -> 7 planes.  SMALL: n <= 15, the Dirichlet quotient comes from the table
template <bool SMALL>
__device__ __forceinline__ void window_planes(const pcacc_bev_params &bp, const BevConsts &cst,
                                              const double *__restrict__ lut, uint32_t n,
                                              uint32_t n_road, uint32_t n_veh, long long fx_hi,
                                              long long fx_lo, double ez, const int med2[3],
                                              double intensity_div, double plane[7]) {
    if (SMALL) {
        plane[0] = lut[DIR_LUT_OFF + n * 16 + n_road];
        plane[5] = lut[DIR_LUT_OFF + n * 16 + n_veh];
    } else {
        plane[0] = dirichlet(n_road, n);
        plane[5] = dirichlet(n_veh, n);
    }
    if (n_road == 0) {
        plane[1] = cst.empty[1];   // sum 0 over count 0: the same arithmetic as an empty window
    } else {
        plane[1] = intensity_plane(bp, fx_to_double(fx_hi, fx_lo), n_road, intensity_div);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) plane[2 + k] = lut[med2[k]];
    plane[6] = ez;
}

template <bool F64OUT>
__device__ __forceinline__ void store_empty(__half *__restrict__ out16, double *__restrict__ out64,
                                            int64_t o, int PP, const BevConsts &cst) {
    out16[o] = cst.empty_h[0];
    out16[o + (int64_t)PP] = cst.empty_h[1];
    out16[o + 2 * (int64_t)PP] = cst.empty_h[2];
    out16[o + 3 * (int64_t)PP] = cst.empty_h[2];
    out16[o + 4 * (int64_t)PP] = cst.empty_h[2];
    out16[o + 5 * (int64_t)PP] = cst.empty_h[0];
    out16[o + 6 * (int64_t)PP] = cst.empty_h[3];
    if (F64OUT) {
#pragma unroll
        for (int p = 0; p < 7; p++) out64[o + (int64_t)p * PP] = cst.empty[p];
    }
}

template <bool F64OUT>
__device__ __forceinline__ void store_planes(__half *__restrict__ out16, double *__restrict__ out64,
                                             int64_t o, int PP, const double plane[7]) {
#pragma unroll
    for (int p = 0; p < 7; p++) {
        out16[o + (int64_t)p * PP] = __double2half(plane[p]);
        if (F64OUT) out64[o + (int64_t)p * PP] = plane[p];
    }
}

#define SMALL_T 15      /* cells with <= SMALL_T points are reduced by their own lane */
#define SMALL_STRIDE 17 /* odd stride: lanes reading the same j hit distinct banks */

// r, g, b in three 10-bit fields (bit 9 of each field is a guard bit): one subtraction
// compares all three channels at once.
#define F_ONE 0x00100401u    /* 1 in every field */
#define F_GUARD 0x20080200u  /* bit 9 of every field */
__device__ __forceinline__ uint32_t f_pack(uint32_t rgbs) {
    return (rgbs & 255u) | ((rgbs & 0xff00u) << 2) | ((rgbs & 0xff0000u) << 4);
}
// bit 0 of field c = (a_c >= b_c); ag = a | F_GUARD
__device__ __forceinline__ uint32_t f_ge(uint32_t ag, uint32_t b) { return ((ag - b) & F_GUARD) >> 9; }

// per-warp shared-memory state of pass A
// Per-cell arrays are indexed [k][cell]: the lanes of a warp (one cell each, or neighbouring
// cells in pass 1) then hit consecutive banks.  The shared-memory data pipe is this kernel's
// busiest unit (ncu: 69 % of its peak), a third of its wavefronts were bank-conflict replays
// of [cell][k] layouts.
struct __align__(16) SmallWarp {
    uint32_t val[32 * SMALL_T];      // packed r,g,b of the small cells' points, back to back
    // ---- zeroed together with 16-byte stores --------------------------------------
    uint32_t cnt[32];                // four 8-bit counts: road present / future, vehicle present / future
    unsigned long long fx_hi[2][32], fx_lo[2][32];
    uint32_t med[6][32];             // lo/hi order statistics: present, future, full
    // --------------------------------------------------------------------------------
    unsigned long long zc[2][32];
    uint32_t cs[33];                 // first compacted index of each cell (cs[32] = total)
    uint32_t s0[32];                 // first record of each cell in `sorted`
    uint8_t meta[32 * SMALL_T];      // cell | window << 5 of each point
    uint8_t np[32], nt[32];
};
#define SW_ZERO_BYTES (32 * 4 + 2 * 2 * 32 * 8 + 6 * 32 * 4)
static_assert(offsetof(SmallWarp, cnt) % 16 == 0 && SW_ZERO_BYTES % 16 == 0 &&
              offsetof(SmallWarp, zc) == offsetof(SmallWarp, cnt) + SW_ZERO_BYTES, "zeroed block");

// ---------------------------------------------------------------------------
// pass A: one warp per 32 consecutive cells.  Empty cells take the per-variant
// constants, cells with <= SMALL_T points are reduced by their own lane, larger
// cells are queued for pass B.
// ---------------------------------------------------------------------------
template <bool F64OUT>
__global__ void __launch_bounds__(RED_WARPS * 32)
k_bev_reduce(const uint32_t *__restrict__ start, const uint4 *__restrict__ sorted,
             const pcacc_bev_params *__restrict__ params, const BevConsts *__restrict__ consts,
             const double *__restrict__ lut, int n_var, int P, double intensity_div,
             uint32_t *__restrict__ big_list, uint32_t *__restrict__ big_count,
             __half *__restrict__ out16, double *__restrict__ out64) {
    __shared__ SmallWarp s_sw[RED_WARPS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int PP = P * P;
    // blockIdx.y = variant, blockIdx.x = 32*RED_WARPS consecutive cells of it; when PP is not a
    // multiple of 32 the lanes past the last cell of the last warp are idle (valid = false)
    const int var = (int)blockIdx.y;
    const int cw = ((int)blockIdx.x * RED_WARPS + (int)warp) * 32;
    if (cw >= PP) return;
    const bool valid = cw + (int)lane < PP;
    const int cell_in = valid ? cw + (int)lane : PP - 1;
    const pcacc_bev_params &bp = params[var];
    const BevConsts &cst = consts[var];
    const bool want_max = bp.elevation_max != 0;

    // segment bounds of my cell: [s0, s1) present, [s1, s2) future
    const uint32_t gc = (uint32_t)var * (uint32_t)PP + (uint32_t)cell_in;
    const uint2 s01 = ((const uint2 *)start)[gc];
    uint32_t s2 = __shfl_down_sync(0xffffffffu, s01.x, 1);
    if (lane == 31 || cw + (int)lane + 1 >= PP) s2 = start[2 * (size_t)gc + 2];
    const uint32_t my_np = valid ? s01.y - s01.x : 0u, my_nf = valid ? s2 - s01.y : 0u, my_nt = my_np + my_nf;
    const int64_t o0 = ((int64_t)var * 3 * 7) * PP + cell_in;

    // all 32 cells empty: nothing to read
    if (__ballot_sync(0xffffffffu, my_nt != 0) == 0) {
        if (valid) {
#pragma unroll
            for (int w = 0; w < 3; w++) store_empty<F64OUT>(out16, out64, o0 + (int64_t)w * 7 * PP, PP, cst);
        }
        return;
    }

    // queue the large cells for pass B
    {
        const bool big = my_nt > SMALL_T;
        const unsigned bm = __ballot_sync(0xffffffffu, big);
        if (bm) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(big_count, (uint32_t)__popc(bm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (big) big_list[base + __popc(bm & ((1u << lane) - 1u))] = gc;
        }
    }

    // ---- small cells, point-parallel ---------------------------------------------------
    // The points of the warp's small cells are laid out back to back (compacted index q)
    // and spread over the lanes, so a lane's work is one point, not one cell: loads are
    // coalesced and the rank loops of neighbouring lanes have (nearly) the same length.
    const bool small = my_nt > 0 && my_nt <= SMALL_T;
    uint32_t cs_incl = small ? my_nt : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, cs_incl, o);
        if (lane >= (unsigned)o) cs_incl += t;
    }
    const uint32_t T = __shfl_sync(0xffffffffu, cs_incl, 31);
    SmallWarp &sw = s_sw[warp];
    sw.cs[lane] = cs_incl - (small ? my_nt : 0u);
    if (lane == 31) sw.cs[32] = T;
    sw.s0[lane] = s01.x;
    sw.np[lane] = (uint8_t)(small ? my_np : 0u);
    sw.nt[lane] = (uint8_t)(small ? my_nt : 0u);
    {
        uint4 *z4 = (uint4 *)sw.cnt;
#pragma unroll
        for (int k = 0; k < (SW_ZERO_BYTES / 16 + 31) / 32; k++)
            if (lane + 32 * k < SW_ZERO_BYTES / 16) z4[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        sw.zc[0][lane] = sw.zc[1][lane] = want_max ? 0ull : ~0ull;   // order-encoded -inf / +inf
    }
    if (small) {   // tag my points with cell | window << 5
        const uint32_t b = cs_incl - my_nt;
        for (uint32_t j = 0; j < my_nt; j++) sw.meta[b + j] = (uint8_t)(lane | (j >= my_np ? 32u : 0u));
    }
    __syncwarp();

    if (T) {
        const int road_cls = bp.road_cls, v0 = bp.veh_cls[0], v1 = bp.veh_cls[1], v2 = bp.veh_cls[2],
                  v3 = bp.veh_cls[3];
        // pass 1: one record per lane and round; statistics through shared-memory atomics
        for (uint32_t q = lane; q < T; q += 32) {
            const int c = sw.meta[q] & 31, w = sw.meta[q] >> 5;
            const uint32_t li = q - sw.cs[c];
            const uint4 r = sorted[sw.s0[c] + li];
            sw.val[q] = f_pack(r.z);
            const int sem = (int)(r.z >> 24);
            uint32_t add = 0;
            if (sem == road_cls) {
                const long long fx = __double2ll_rn(__dmul_rn((double)__uint_as_float(r.w), FX_SCALE));
                add = 1u << (8 * w);
                atomicAdd(&sw.fx_hi[w][c], (unsigned long long)(fx >> 32));
                atomicAdd(&sw.fx_lo[w][c], (unsigned long long)(fx & 0xffffffffll));
            }
            if ((sem == v0) || (sem == v1) || (sem == v2) || (sem == v3)) add += 1u << (16 + 8 * w);
            if (add) atomicAdd(&sw.cnt[c], add);   // at most SMALL_T per 8-bit field
            const unsigned long long zc =
                ord_encode(__longlong_as_double((long long)(((unsigned long long)r.y << 32) | r.x)));
            if (want_max) atomicMax(&sw.zc[w][c], zc); else atomicMin(&sw.zc[w][c], zc);
        }
        __syncwarp();
        // pass 2: order statistics without sorting.  Point i occupies the ranks [L_i, E_i) of
        // its set (L = #{v_j < v_i}, E = #{v_j <= v_i}); it is the k-th smallest iff
        // L <= k < E.  Counted separately over the present and the future part of the cell,
        // so the window's own ranks and the full cell's ranks come out of the same loop;
        // r, g, b are compared together in three 10-bit fields.
        for (uint32_t q = lane; q < T; q += 32) {
            const int c = sw.meta[q] & 31, wi = sw.meta[q] >> 5;
            const uint32_t base = sw.cs[c], np = sw.np[c], nt = sw.nt[c], nf = nt - np;
            const uint32_t vi = sw.val[q], vig = vi | F_GUARD;
            uint32_t e1 = 0, g1 = 0, e2 = 0, g2 = 0;
#pragma unroll 2
            for (uint32_t j = 0; j < np; j++) {
                const uint32_t vj = sw.val[base + j];
                e1 += f_ge(vig, vj);
                g1 += f_ge(vj | F_GUARD, vi);
            }
#pragma unroll 2
            for (uint32_t j = np; j < nt; j++) {
                const uint32_t vj = sw.val[base + j];
                e2 += f_ge(vig, vj);
                g2 += f_ge(vj | F_GUARD, vi);
            }
            {   // own window
                const uint32_t nw = wi ? nf : np;
                const uint32_t E = wi ? e2 : e1, L = nw * F_ONE - (wi ? g2 : g1);
                const uint32_t klo = (nw - 1) / 2 * F_ONE, khi = nw / 2 * F_ONE;
                const uint32_t m = (f_ge(klo | F_GUARD, L) & f_ge(E | F_GUARD, klo + F_ONE)) * 0x3ffu;
                const uint32_t h = (f_ge(khi | F_GUARD, L) & f_ge(E | F_GUARD, khi + F_ONE)) * 0x3ffu;
                if (m) atomicOr(&sw.med[2 * wi][c], vi & m);
                if (h) atomicOr(&sw.med[2 * wi + 1][c], vi & h);
            }
            {   // full cell
                const uint32_t E = e1 + e2, L = nt * F_ONE - (g1 + g2);
                const uint32_t klo = (nt - 1) / 2 * F_ONE, khi = nt / 2 * F_ONE;
                const uint32_t m = (f_ge(klo | F_GUARD, L) & f_ge(E | F_GUARD, klo + F_ONE)) * 0x3ffu;
                const uint32_t h = (f_ge(khi | F_GUARD, L) & f_ge(E | F_GUARD, khi + F_ONE)) * 0x3ffu;
                if (m) atomicOr(&sw.med[4][c], vi & m);
                if (h) atomicOr(&sw.med[5][c], vi & h);
            }
        }
        __syncwarp();
    }

    // back to one cell per lane (large cells are written by pass B)
    const bool fin = valid && my_nt <= SMALL_T;
    uint32_t nr[2], nv[2];
    long long hi[2], lo[2];
    double ez[2];
    {
        const uint32_t cw4 = sw.cnt[lane];
        nr[0] = cw4 & 255u;
        nr[1] = (cw4 >> 8) & 255u;
        nv[0] = (cw4 >> 16) & 255u;
        nv[1] = cw4 >> 24;
    }
#pragma unroll
    for (int w = 0; w < 2; w++) {
        hi[w] = (long long)sw.fx_hi[w][lane];
        lo[w] = (long long)sw.fx_lo[w][lane];
        ez[w] = ord_decode(sw.zc[w][lane]);   // an empty window keeps the identity of min / max
    }
    int med2[3][3];
#pragma unroll
    for (int w = 0; w < 3; w++) {
        // lo + hi per 10-bit field in one add (at most 510 per field)
        const uint32_t sm = sw.med[2 * w][lane] + sw.med[2 * w + 1][lane];
#pragma unroll
        for (int c = 0; c < 3; c++) med2[w][c] = (int)((sm >> (10 * c)) & 1023u);
    }

    // intensity planes: the (cell, window) pairs that hold road points are compacted over the
    // warp, so the division / exp chain runs once per 32 pairs instead of once per window.
    // 'full' needs its own evaluation only if both halves hold road points; otherwise its sum
    // and count are those of the non-empty half (same operands, same result).
    const bool need0 = fin && nr[0] != 0, need1 = fin && nr[1] != 0, need2 = need0 && need1;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned m0 = __ballot_sync(0xffffffffu, need0), m1 = __ballot_sync(0xffffffffu, need1),
                   m2 = __ballot_sync(0xffffffffu, need2);
    const uint32_t b1 = (uint32_t)__popc(m0), b2 = b1 + (uint32_t)__popc(m1), tot = b2 + (uint32_t)__popc(m2);
    const uint32_t k0 = (uint32_t)__popc(m0 & lt), k1 = b1 + (uint32_t)__popc(m1 & lt),
                   k2 = b2 + (uint32_t)__popc(m2 & lt);
    // the packed values are dead after pass 2: 96 x (hi, lo, n) fit exactly
    long long *e_hi = (long long *)sw.val, *e_lo = e_hi + 96;
    uint32_t *e_n = (uint32_t *)(e_lo + 96);
    static_assert(sizeof(sw.val) >= 96 * (2 * sizeof(long long) + sizeof(uint32_t)), "evaluation slots");
    if (need0) { e_hi[k0] = hi[0]; e_lo[k0] = lo[0]; e_n[k0] = nr[0]; }
    if (need1) { e_hi[k1] = hi[1]; e_lo[k1] = lo[1]; e_n[k1] = nr[1]; }
    if (need2) { e_hi[k2] = hi[0] + hi[1]; e_lo[k2] = lo[0] + lo[1]; e_n[k2] = nr[0] + nr[1]; }
    __syncwarp();
    for (uint32_t i = lane; i < tot; i += 32) {
        const double val = intensity_plane(bp, fx_to_double(e_hi[i], e_lo[i]), e_n[i], intensity_div);
        e_hi[i] = __double_as_longlong(val);
    }
    __syncwarp();
    double I[3];
    I[0] = need0 ? __longlong_as_double(e_hi[k0]) : cst.empty[1];   // no road points: sum 0 over count 0,
    I[1] = need1 ? __longlong_as_double(e_hi[k1]) : cst.empty[1];   // the arithmetic of an empty window
    I[2] = need2 ? __longlong_as_double(e_hi[k2]) : (need0 ? I[0] : I[1]);

    // finalise: 3 windows x 7 planes, one store sequence for empty and non-empty windows
    // (the Dirichlet table's entry n = c = 0 is the empty window's 1/2)
    if (fin) {
        const __half *lut16 = (const __half *)(lut + LUT_TOTAL);
        __half *ob = out16 + o0;
        const __half eh_rgb = cst.empty_h[2], eh_z = cst.empty_h[3];
        __half *op = ob;
#pragma unroll
        for (int w = 0; w < 3; w++) {
            const uint32_t nw = (w == 0) ? my_np : (w == 1) ? my_nf : my_nt;
            const uint32_t nrw = (w < 2) ? nr[w & 1] : nr[0] + nr[1], nvw = (w < 2) ? nv[w & 1] : nv[0] + nv[1];
            const double zz = (w < 2) ? ez[w & 1] : (want_max ? fmax(ez[0], ez[1]) : fmin(ez[0], ez[1]));
            const bool ne = nw != 0;
            const uint32_t di = DIR_LUT_OFF + nw * 16;
            const size_t ow = (size_t)w * 7 * PP;
            if (F64OUT) {
                double plane[7];
                plane[0] = lut[di + nrw];
                plane[1] = I[w];
#pragma unroll
                for (int k = 0; k < 3; k++) plane[2 + k] = ne ? lut[med2[w][k]] : cst.empty[2];
                plane[5] = lut[di + nvw];
                plane[6] = ne ? zz : cst.empty[6];
                store_planes<true>(out16, out64, o0 + (int64_t)ow, PP, plane);
            } else {
                // branch-free: the table is read for empty windows too (index 0) and the
                // per-variant fill value selected afterwards
                __half h[7];
                h[0] = lut16[di + nrw];
                h[1] = __double2half(I[w]);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const __half t = lut16[med2[w][k]];
                    h[2 + k] = ne ? t : eh_rgb;
                }
                h[5] = lut16[di + nvw];
                const __half hz = __double2half(zz);
                h[6] = ne ? hz : eh_z;
                // the 21 planes of a cell are PP elements apart: walk one pointer
#pragma unroll
                for (int p = 0; p < 7; p++) {
                    *op = h[p];
                    op += PP;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// pass A, chunked (float16 output only, P*P a multiple of CH_CELLS): one warp per CH_CELLS
// consecutive cells.  In the dataset-generation workload a 32-cell strip holds ~16 points in
// ~7 non-empty cells, and k_bev_reduce pays its fixed per-strip work (segment scans, zeroing,
// the 21-plane finalisation of 32 lanes) for every strip that is not completely empty.  Here
// the non-empty small cells of a 128-cell chunk are compacted first and reduced 32 at a time
// (lane = one non-empty cell: the finalisation never runs for an empty cell), results go to
// a shared-memory tile of the chunk's 21 x 128 planes that starts out as the empty-window
// constants and leaves as 21 fully coalesced 256-byte row pieces.
// ---------------------------------------------------------------------------
#define CH_WARPS 2
#ifndef CH_CPL
#define CH_CPL 4          /* cells per lane: a chunk is 32 * CH_CPL consecutive cells (8: 62 us, 16: 96 us against 52 us: fewer groups, but longer serial chains per warp) */
#endif
#define CH_CELLS (32 * CH_CPL)
#define CH_PER_LANE CH_CPL

// per-warp state of one group of <= 32 small cells.  Arrays are [k][cell] (see SmallWarp).
// Every accumulator is a 32-bit word updated with native shared-memory atomics: the 64-bit
// ones of k_bev_reduce compile to compare-and-swap loops (ATOMS.CAST.SPIN).
struct __align__(16) GroupWarp {
    uint32_t val[32 * SMALL_T];      // packed r,g,b of the points; later the intensity evaluation slots
    // ---- zeroed together with 16-byte stores --------------------------------------
    uint32_t cnt[32];                // four 8-bit counts: road present / future, vehicle present / future
    uint32_t fx0[2][32], fx1[2][32]; // raw road intensity sums, 2^-40 fixed point, in three 21-bit limbs:
    int32_t fx2[2][32];              //   fx = fx0 + fx1 * 2^21 + fx2 * 2^42 (fx2 signed)
    uint32_t med[6][32];             // lo/hi order statistics: present, future, full
    // --------------------------------------------------------------------------------
    uint32_t zc[2][32];              // order-encoded float16 of the extreme z per window
    uint32_t cs[33];                 // first compacted index of each cell (cs[32] = total)
    uint32_t s0[32];                 // first record of each cell in `sorted`
    uint8_t meta[32 * SMALL_T];      // cell | window << 5 of each point
    uint8_t np[32], nt[32];
};
#define GW_ZERO_BYTES (32 * 4 + 3 * 2 * 32 * 4 + 6 * 32 * 4)
static_assert(offsetof(GroupWarp, cnt) % 16 == 0 && GW_ZERO_BYTES % 16 == 0 &&
              offsetof(GroupWarp, zc) == offsetof(GroupWarp, cnt) + GW_ZERO_BYTES, "zeroed block");

struct __align__(16) ChunkWarp {
    GroupWarp gw;
    uint32_t g_s0[CH_CELLS];     // compacted non-empty small cells: first record in `sorted`
    uint16_t g_cell[CH_CELLS];   // ... cell index inside the chunk
    uint8_t g_np[CH_CELLS], g_nt[CH_CELLS];
};

// float16 bits <-> unsigned keys with the same order (rounding to float16 is monotone, so
// the float16 of the smallest z is the smallest float16: the extreme can be taken after the cast)
__device__ __forceinline__ uint32_t h_ord(__half h) {
    const uint32_t b = __half_as_ushort(h);
    return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u);
}
__device__ __forceinline__ __half h_unord(uint32_t u) {
    const uint32_t b = (u & 0x8000u) ? (u & 0x7fffu) : (~u & 0xffffu);
    return __ushort_as_half((unsigned short)b);
}

__global__ void __launch_bounds__(CH_WARPS * 32, 16)
k_bev_reduce_chunk(const uint32_t *__restrict__ start, const uint4 *__restrict__ sorted,
                   const pcacc_bev_params *__restrict__ params, const BevConsts *__restrict__ consts,
                   const double *__restrict__ lut, int P, double intensity_div, uint32_t small_t,
                   uint32_t *__restrict__ big_list, uint32_t *__restrict__ big_count,
                   __half *__restrict__ out16) {
    __shared__ ChunkWarp s_cw[CH_WARPS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int PP = P * P;
    // Longest chunks first: the points are densest around the ego vehicle, i.e. in the middle
    // rows of the grid, and a chunk there is four groups of serial work while one at the rim is
    // 21 stores.  Blocks are dispatched x-fastest, so x = variant and y walks the chunks from the
    // centre row outwards — the dense chunks of every variant start in the first wave instead of
    // forming the kernel's tail.
    const int var = (int)blockIdx.x;
    const int n_ch = PP / CH_CELLS;
    const int ord = (int)blockIdx.y * CH_WARPS + (int)warp;
    if (ord >= n_ch) return;
    const int mid = n_ch >> 1;
    int chunk = (ord & 1) ? mid - ((ord + 1) >> 1) : mid + (ord >> 1);
    if (chunk < 0) chunk += n_ch;            // odd chunk counts: the last odd step wraps below 0
    if (chunk >= n_ch) chunk -= n_ch;
    const int c0 = chunk * CH_CELLS;
    ChunkWarp &cw = s_cw[warp];
    GroupWarp &sw = cw.gw;
    const pcacc_bev_params &bp = params[var];
    const BevConsts &cst = consts[var];
    const bool want_max = bp.elevation_max != 0;

    // segment bounds of my CH_PER_LANE cells: words 2*gc .. 2*gc + 2*CH_PER_LANE (coalesced 16 B loads)
    const uint32_t gc0 = (uint32_t)var * (uint32_t)PP + (uint32_t)c0 + CH_PER_LANE * lane;
    uint32_t sb[2 * CH_PER_LANE + 1];
    {
        const uint4 *s4 = (const uint4 *)(start + 2 * (size_t)gc0);
#pragma unroll
        for (int k = 0; k < CH_PER_LANE / 2; k++) {
            const uint4 v = s4[k];
            sb[4 * k] = v.x; sb[4 * k + 1] = v.y; sb[4 * k + 2] = v.z; sb[4 * k + 3] = v.w;
        }
        uint32_t nxt = __shfl_down_sync(0xffffffffu, sb[0], 1);
        if (lane == 31) nxt = start[2 * (size_t)gc0 + 2 * CH_PER_LANE];
        sb[2 * CH_PER_LANE] = nxt;
    }
    uint32_t c_np[CH_PER_LANE], c_nt[CH_PER_LANE];
    uint32_t n_big = 0, n_small = 0;
#pragma unroll
    for (int k = 0; k < CH_PER_LANE; k++) {
        c_np[k] = sb[2 * k + 1] - sb[2 * k];
        c_nt[k] = sb[2 * k + 2] - sb[2 * k];
        n_big += c_nt[k] > small_t;
        n_small += c_nt[k] > 0 && c_nt[k] <= small_t;
    }
    // The empty-window constants go out first, for the whole chunk: plane p of window w of the
    // chunk is CH_CELLS halves = 32 lanes x 8 bytes.  The non-empty cells overwrite theirs
    // afterwards (ordered by the __syncwarp below).
    __half *const ob = out16 + ((int64_t)var * 21) * PP + c0;
    static_assert(CH_PER_LANE == 4 || CH_PER_LANE % 8 == 0, "8- or 16-byte pieces per lane and plane");
    {
        const uint32_t r = __half_as_ushort(cst.empty_h[0]), i = __half_as_ushort(cst.empty_h[1]),
                       g = __half_as_ushort(cst.empty_h[2]), z = __half_as_ushort(cst.empty_h[3]);
        const uint32_t e_road = r | (r << 16), e_int = i | (i << 16), e_rgb = g | (g << 16), e_z = z | (z << 16);
        __half *o = ob + CH_PER_LANE * lane;
#pragma unroll
        for (int w = 0; w < 3; w++) {
#pragma unroll
            for (int pl = 0; pl < 7; pl++) {
                const uint32_t e = pl == 0 || pl == 5 ? e_road : pl == 1 ? e_int : pl == 6 ? e_z : e_rgb;
                if (CH_PER_LANE == 4) {
                    *(uint2 *)o = make_uint2(e, e);
                } else {
#pragma unroll
                    for (int q = 0; q < CH_PER_LANE / 8; q++) *(uint4 *)(o + 8 * q) = make_uint4(e, e, e, e);
                }
                o += PP;
            }
        }
    }
    // queue the large cells for pass B; compact the small non-empty ones
    {
        uint32_t incl = n_big | (n_small << 16);     // both scans in one (<= 128 each)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot == 0) return;                        // nothing but empty cells
        const uint32_t excl = incl - (n_big | (n_small << 16));
        uint32_t bbase = 0;
        if ((tot & 0xffffu) != 0) {
            if (lane == 0) bbase = atomicAdd(big_count, tot & 0xffffu);
            bbase = __shfl_sync(0xffffffffu, bbase, 0);
        }
        uint32_t bi = bbase + (excl & 0xffffu), si = excl >> 16;
#pragma unroll
        for (int k = 0; k < CH_PER_LANE; k++) {
            if (c_nt[k] > small_t) {
                big_list[bi++] = gc0 + k;
            } else if (c_nt[k] > 0) {
                cw.g_s0[si] = sb[2 * k];
                cw.g_cell[si] = (uint16_t)(CH_PER_LANE * lane + k);
                cw.g_np[si] = (uint8_t)c_np[k];
                cw.g_nt[si] = (uint8_t)c_nt[k];
                si++;
            }
        }
        n_small = tot >> 16;      // now: the chunk's total
    }
    __syncwarp();                 // lists visible; constant stores ordered before the overwrites

    const int road_cls = bp.road_cls, v0 = bp.veh_cls[0], v1 = bp.veh_cls[1], v2 = bp.veh_cls[2],
              v3 = bp.veh_cls[3];
    const __half *lut16 = (const __half *)(lut + LUT_TOTAL);
    const __half eh_rgb = cst.empty_h[2], eh_z = cst.empty_h[3], eh_int = cst.empty_h[1];
    const uint32_t z_init = want_max ? 0u : 0xffffffffu;
    for (uint32_t g0 = 0; g0 < n_small; g0 += 32) {
        // ---- one group of <= 32 non-empty small cells: lane = cell --------------------------
        const bool have = g0 + lane < n_small;
        const uint32_t my_np = have ? cw.g_np[g0 + lane] : 0u, my_nt = have ? cw.g_nt[g0 + lane] : 0u;
        const uint32_t my_nf = my_nt - my_np;
        uint32_t cs_incl = my_nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, cs_incl, o);
            if (lane >= (unsigned)o) cs_incl += t;
        }
        const uint32_t T = __shfl_sync(0xffffffffu, cs_incl, 31);
        sw.cs[lane] = cs_incl - my_nt;
        if (lane == 31) sw.cs[32] = T;
        sw.s0[lane] = have ? cw.g_s0[g0 + lane] : 0u;
        sw.np[lane] = (uint8_t)my_np;
        sw.nt[lane] = (uint8_t)my_nt;
        {
            uint4 *z4 = (uint4 *)sw.cnt;
#pragma unroll
            for (int k = 0; k < (GW_ZERO_BYTES / 16 + 31) / 32; k++)
                if (lane + 32 * k < GW_ZERO_BYTES / 16) z4[lane + 32 * k] = make_uint4(0, 0, 0, 0);
            sw.zc[0][lane] = sw.zc[1][lane] = z_init;
        }
        {
            const uint32_t b = cs_incl - my_nt;
            for (uint32_t j = 0; j < my_nt; j++) sw.meta[b + j] = (uint8_t)(lane | (j >= my_np ? 32u : 0u));
        }
        __syncwarp();
        // pass 1: statistics, one record per lane and round
        for (uint32_t q = lane; q < T; q += 32) {
            const int c = sw.meta[q] & 31, w = sw.meta[q] >> 5;
            const uint32_t li = q - sw.cs[c];
            const uint4 r = sorted[sw.s0[c] + li];
            sw.val[q] = f_pack(r.z);
            const int sem = (int)(r.z >> 24);
            uint32_t add = 0;
            if (sem == road_cls) {
                const long long fx = __double2ll_rn(__dmul_rn((double)__uint_as_float(r.w), FX_SCALE));
                add = 1u << (8 * w);
                atomicAdd(&sw.fx0[w][c], (uint32_t)fx & 0x1fffffu);
                atomicAdd(&sw.fx1[w][c], (uint32_t)(fx >> 21) & 0x1fffffu);
                atomicAdd(&sw.fx2[w][c], (int32_t)(fx >> 42));
            }
            if ((sem == v0) || (sem == v1) || (sem == v2) || (sem == v3)) add += 1u << (16 + 8 * w);
            if (add) atomicAdd(&sw.cnt[c], add);
            const uint32_t zk =
                h_ord(__double2half(__longlong_as_double((long long)(((unsigned long long)r.y << 32) | r.x))));
            if (want_max) atomicMax(&sw.zc[w][c], zk); else atomicMin(&sw.zc[w][c], zk);
        }
        __syncwarp();
        // pass 2: rank intervals (see k_bev_reduce)
        for (uint32_t q = lane; q < T; q += 32) {
            const int c = sw.meta[q] & 31, wi = sw.meta[q] >> 5;
            const uint32_t base = sw.cs[c], np = sw.np[c], nt = sw.nt[c], nf = nt - np;
            const uint32_t vi = sw.val[q], vig = vi | F_GUARD;
            if (nt == 1) {                      // the commonest cell: its one point is every median
                sw.med[2 * wi][c] = vi;
                sw.med[2 * wi + 1][c] = vi;
                sw.med[4][c] = vi;
                sw.med[5][c] = vi;
                continue;
            }
            uint32_t e1 = 0, g1 = 0, e2 = 0, g2 = 0;
#pragma unroll 2
            for (uint32_t j = 0; j < np; j++) {
                const uint32_t vj = sw.val[base + j];
                e1 += f_ge(vig, vj);
                g1 += f_ge(vj | F_GUARD, vi);
            }
#pragma unroll 2
            for (uint32_t j = np; j < nt; j++) {
                const uint32_t vj = sw.val[base + j];
                e2 += f_ge(vig, vj);
                g2 += f_ge(vj | F_GUARD, vi);
            }
            {
                const uint32_t nw = wi ? nf : np;
                const uint32_t E = wi ? e2 : e1, L = nw * F_ONE - (wi ? g2 : g1);
                const uint32_t klo = (nw - 1) / 2 * F_ONE, khi = nw / 2 * F_ONE;
                const uint32_t m = (f_ge(klo | F_GUARD, L) & f_ge(E | F_GUARD, klo + F_ONE)) * 0x3ffu;
                const uint32_t h = (f_ge(khi | F_GUARD, L) & f_ge(E | F_GUARD, khi + F_ONE)) * 0x3ffu;
                if (m) atomicOr(&sw.med[2 * wi][c], vi & m);
                if (h) atomicOr(&sw.med[2 * wi + 1][c], vi & h);
            }
            {
                const uint32_t E = e1 + e2, L = nt * F_ONE - (g1 + g2);
                const uint32_t klo = (nt - 1) / 2 * F_ONE, khi = nt / 2 * F_ONE;
                const uint32_t m = (f_ge(klo | F_GUARD, L) & f_ge(E | F_GUARD, klo + F_ONE)) * 0x3ffu;
                const uint32_t h = (f_ge(khi | F_GUARD, L) & f_ge(E | F_GUARD, khi + F_ONE)) * 0x3ffu;
                if (m) atomicOr(&sw.med[4][c], vi & m);
                if (h) atomicOr(&sw.med[5][c], vi & h);
            }
        }
        __syncwarp();

        uint32_t nr[2], nv[2];
        uint32_t zk[2];
        {
            const uint32_t cw4 = sw.cnt[lane];
            nr[0] = cw4 & 255u;
            nr[1] = (cw4 >> 8) & 255u;
            nv[0] = (cw4 >> 16) & 255u;
            nv[1] = cw4 >> 24;
        }
        zk[0] = sw.zc[0][lane];
        zk[1] = sw.zc[1][lane];
        uint32_t med2[3];
#pragma unroll
        for (int w = 0; w < 3; w++) med2[w] = sw.med[2 * w][lane] + sw.med[2 * w + 1][lane];

        // intensity: (cell, window) pairs with road points compacted over the warp
        const bool need0 = have && nr[0] != 0, need1 = have && nr[1] != 0, need2 = need0 && need1;
        const unsigned lt = (1u << lane) - 1u;
        const unsigned m0 = __ballot_sync(0xffffffffu, need0), m1 = __ballot_sync(0xffffffffu, need1),
                       m2 = __ballot_sync(0xffffffffu, need2);
        const uint32_t b1 = (uint32_t)__popc(m0), b2 = b1 + (uint32_t)__popc(m1), tot = b2 + (uint32_t)__popc(m2);
        const uint32_t k0 = (uint32_t)__popc(m0 & lt), k1 = b1 + (uint32_t)__popc(m1 & lt),
                       k2 = b2 + (uint32_t)__popc(m2 & lt);
        // slot = the three limb sums (the exact integer sum; converted by the evaluating lane) + count
        uint4 *e_sl = (uint4 *)sw.val;
        static_assert(sizeof(sw.val) >= 96 * sizeof(uint4), "evaluation slots");
        {
            const uint32_t a0 = sw.fx0[0][lane], a1 = sw.fx1[0][lane], c0_ = sw.fx0[1][lane], c1_ = sw.fx1[1][lane];
            const int32_t a2 = sw.fx2[0][lane], c2_ = sw.fx2[1][lane];
            if (need0) e_sl[k0] = make_uint4(a0, a1, (uint32_t)a2, nr[0]);
            if (need1) e_sl[k1] = make_uint4(c0_, c1_, (uint32_t)c2_, nr[1]);
            if (need2) e_sl[k2] = make_uint4(a0 + c0_, a1 + c1_, (uint32_t)(a2 + c2_), nr[0] + nr[1]);
        }
        __syncwarp();
        for (uint32_t i = lane; i < tot; i += 32) {
            const uint4 sl = e_sl[i];
            // S = l0 + l1 2^21 + l2 2^42 rounded once, as fx_to_double rounds hi 2^32 + lo
            const double raw = __dmul_rn(__dadd_rn(__dmul_rn((double)(int32_t)sl.z, 4398046511104.0),
                                                   (double)(((long long)sl.y << 21) + (long long)sl.x)), FX_INV);
            const double val = intensity_plane(bp, raw, sl.w, intensity_div);
            ((__half *)e_sl)[8 * i] = __double2half(val);     // in place: slot i is read by this lane only
        }
        __syncwarp();
        __half I[3];
        I[0] = need0 ? ((const __half *)e_sl)[8 * k0] : eh_int;
        I[1] = need1 ? ((const __half *)e_sl)[8 * k1] : eh_int;
        I[2] = need2 ? ((const __half *)e_sl)[8 * k2] : (need0 ? I[0] : I[1]);

        if (have) {
            __half *o = ob + cw.g_cell[g0 + lane];
#pragma unroll
            for (int w = 0; w < 3; w++) {
                const uint32_t nw = (w == 0) ? my_np : (w == 1) ? my_nf : my_nt;
                const uint32_t nrw = (w < 2) ? nr[w & 1] : nr[0] + nr[1], nvw = (w < 2) ? nv[w & 1] : nv[0] + nv[1];
                const uint32_t zz = (w < 2) ? zk[w & 1] : (want_max ? max(zk[0], zk[1]) : min(zk[0], zk[1]));
                const bool ne = nw != 0;
                const uint32_t di = DIR_LUT_OFF + nw * 16;
                __half h[7];
                h[0] = lut16[di + nrw];
                h[1] = I[w];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const __half t = lut16[(med2[w] >> (10 * k)) & 1023u];
                    h[2 + k] = ne ? t : eh_rgb;
                }
                h[5] = lut16[di + nvw];
                h[6] = ne ? h_unord(zz) : eh_z;
#pragma unroll
                for (int pl = 0; pl < 7; pl++) {
                    *o = h[pl];
                    o += PP;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// pass B: one warp per queued large cell (dynamic queue), 256-bin shared-memory
// histograms per window and channel.
// ---------------------------------------------------------------------------
// warp-wide reductions on the REDUX unit (one instruction per 32-bit word)
__device__ __forceinline__ long long warp_sum_i64(long long x) {
    // three 21-bit limbs: the 32 partial limbs sum without overflow
    const unsigned l0 = (unsigned)(x & 0x1fffff), l1 = (unsigned)((x >> 21) & 0x1fffff);
    const int l2 = (int)(x >> 42);
    const unsigned s0 = __reduce_add_sync(0xffffffffu, l0), s1 = __reduce_add_sync(0xffffffffu, l1);
    const int s2 = __reduce_add_sync(0xffffffffu, l2);
    return (long long)s0 + ((long long)s1 << 21) + ((long long)s2 << 42);
}
__device__ __forceinline__ double warp_min_f64(double v) {
    const unsigned long long u = ord_encode(v);
    const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ord_decode(((unsigned long long)mh << 32) | ml);
}
__device__ __forceinline__ double warp_max_f64(double v) {
    const unsigned long long u = ord_encode(v);
    const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ord_decode(((unsigned long long)mh << 32) | ml);
}

// parked (cell, window) statistics of k_bev_reduce_big, one warp's worth
struct BigPend {
    long long fx_hi[32], fx_lo[32];
    double ez[32];
    uint32_t nw[32], nr[32], nv[32], med[32], gc[32];   // med: three 10-bit doubled medians | window << 30
};

template <bool F64OUT>
__device__ __forceinline__ void big_flush(const BigPend &pd, uint32_t n, unsigned lane,
                                          const pcacc_bev_params *__restrict__ params,
                                          const BevConsts *__restrict__ consts, const double *__restrict__ lut,
                                          int PP, double intensity_div, __half *__restrict__ out16,
                                          double *__restrict__ out64) {
    if (lane < n) {
        const uint32_t gc = pd.gc[lane], m = pd.med[lane];
        const int w = (int)(m >> 30);
        const int var = (int)(gc / (uint32_t)PP);
        const int cell_in = (int)(gc - (uint32_t)var * (uint32_t)PP);
        const int md[3] = {(int)(m & 1023u), (int)((m >> 10) & 1023u), (int)((m >> 20) & 1023u)};
        double plane[7];
        window_planes<false>(params[var], consts[var], lut, pd.nw[lane], pd.nr[lane], pd.nv[lane], pd.fx_hi[lane],
                             pd.fx_lo[lane], pd.ez[lane], md, intensity_div, plane);
        store_planes<F64OUT>(out16, out64, (((int64_t)var * 3 + w) * 7) * PP + cell_in, PP, plane);
    }
}

#ifndef REDB_BATCH
#define REDB_BATCH 1   /* queue entries per ticket (4: 109 vs 94 us on the long-horizon window, 8: 123) */
#endif
template <bool F64OUT>
__global__ void __launch_bounds__(REDB_WARPS * 32, 7)
k_bev_reduce_big(const uint32_t *__restrict__ start, const uint4 *__restrict__ sorted,
                 const pcacc_bev_params *__restrict__ params, const BevConsts *__restrict__ consts,
                 const double *__restrict__ lut, int P, double intensity_div,
                 const uint32_t *__restrict__ big_list, const uint32_t *__restrict__ big_count,
                 uint32_t *__restrict__ next, __half *__restrict__ out16, double *__restrict__ out64) {
    __shared__ __align__(16) uint32_t s_hist[REDB_WARPS][2][3][256];
    __shared__ BigPend s_pend[REDB_WARPS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t(*hist)[3][256] = s_hist[warp];
    BigPend &pd = s_pend[warp];
    uint32_t pend_n = 0;
    const int PP = P * P;
    const uint32_t n_big = *big_count;
    // Cells are drawn from the queue REDB_BATCH at a time: one ticket, then the cells' ids, segment
    // bounds and a prefetch of their first records fetched by one lane each, side by side.  Drawn one
    // by one, every cell paid four dependent round trips (ticket -> id -> bounds -> records) before
    // its first record arrived, and with 28 warps per SM that chain, not the arithmetic, set the
    // kernel's duration.
    while (true) {
        uint32_t q0 = 0;
        if (lane == 0) q0 = atomicAdd(next, (uint32_t)REDB_BATCH);
        q0 = __shfl_sync(0xffffffffu, q0, 0);
        if (q0 >= n_big) break;
        const uint32_t n_batch = min((uint32_t)REDB_BATCH, n_big - q0);
        uint32_t l_gc = 0, l_b0 = 0, l_b1 = 0, l_b2 = 0;
        if (lane < n_batch) {
            l_gc = big_list[q0 + lane];
            l_b0 = start[2 * (int64_t)l_gc];
            l_b1 = start[2 * (int64_t)l_gc + 1];
            l_b2 = start[2 * (int64_t)l_gc + 2];
            const char *rec = (const char *)(sorted + l_b0);
            const uint32_t bytes = min((l_b2 - l_b0) * 16u, 2048u);
            for (uint32_t o = 0; o < bytes; o += 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + o));
        }
#pragma unroll 1
    for (uint32_t bi = 0; bi < n_batch; bi++) {
        const uint32_t gc = __shfl_sync(0xffffffffu, l_gc, (int)bi);
        const int var = (int)(gc / (uint32_t)PP);
        const int cell_in = (int)(gc - (uint32_t)var * (uint32_t)PP);
        const pcacc_bev_params &bp = params[var];
        const bool want_max = bp.elevation_max != 0;
        const int road_cls = bp.road_cls, v0 = bp.veh_cls[0], v1 = bp.veh_cls[1], v2 = bp.veh_cls[2],
                  v3 = bp.veh_cls[3];
        const uint32_t b0 = __shfl_sync(0xffffffffu, l_b0, (int)bi), b1 = __shfl_sync(0xffffffffu, l_b1, (int)bi),
                       b2 = __shfl_sync(0xffffffffu, l_b2, (int)bi);
        const uint32_t np = b1 - b0, nf = b2 - b1, nt = np + nf;
        // Cells of at most 32 points (one per lane): the rank intervals of pass A with the
        // values exchanged by shuffles — about n x 10 instructions instead of the ~800 it takes
        // to clear and scan nine 256-bin histograms.  Larger cells: histograms.
        const bool medium = nt <= 32;
        WinAcc a;
        acc_init(a, want_max);
        int m2[3][3];
        if (medium) {
            const bool have = lane < nt;
            const uint4 r = have ? sorted[b0 + lane] : make_uint4(0, 0, 0, 0);
            const int w = lane >= np ? 1 : 0;
            if (have) acc_record(a, r, w, road_cls, v0, v1, v2, v3, want_max);
            // per field: guard bit of (vi|G) - vj = (vi >= vj), of (vi|G) - 1 - vj = (vi > vj)
            const uint32_t vi = f_pack(r.z), vig = vi | F_GUARD, vig1 = vig - F_ONE;
            uint32_t e1 = 0, l1 = 0, e2 = 0, l2 = 0;
            for (uint32_t j = 0; j < np; j++) {
                const uint32_t vj = __shfl_sync(0xffffffffu, vi, (int)j);
                e1 += ((vig - vj) & F_GUARD) >> 9;
                l1 += ((vig1 - vj) & F_GUARD) >> 9;
            }
            for (uint32_t j = np; j < nt; j++) {
                const uint32_t vj = __shfl_sync(0xffffffffu, vi, (int)j);
                e2 += ((vig - vj) & F_GUARD) >> 9;
                l2 += ((vig1 - vj) & F_GUARD) >> 9;
            }
            uint32_t own_lo = 0, own_hi = 0, all_lo = 0, all_hi = 0;
            if (have) {
                {   // own window (its size is at least 1: this lane is in it)
                    const uint32_t nw = w ? nf : np;
                    const uint32_t E = w ? e2 : e1, L = w ? l2 : l1;
                    const uint32_t klo = (nw - 1) / 2 * F_ONE, khi = nw / 2 * F_ONE;
                    own_lo = vi & ((f_ge(klo | F_GUARD, L) & f_ge(E | F_GUARD, klo + F_ONE)) * 0x3ffu);
                    own_hi = vi & ((f_ge(khi | F_GUARD, L) & f_ge(E | F_GUARD, khi + F_ONE)) * 0x3ffu);
                }
                {   // full cell
                    const uint32_t E = e1 + e2, L = l1 + l2;
                    const uint32_t klo = (nt - 1) / 2 * F_ONE, khi = nt / 2 * F_ONE;
                    all_lo = vi & ((f_ge(klo | F_GUARD, L) & f_ge(E | F_GUARD, klo + F_ONE)) * 0x3ffu);
                    all_hi = vi & ((f_ge(khi | F_GUARD, L) & f_ge(E | F_GUARD, khi + F_ONE)) * 0x3ffu);
                }
            }
            // every point that holds the wanted rank carries the same value: OR them together
            uint32_t sm[3];
            sm[0] = __reduce_or_sync(0xffffffffu, w == 0 ? own_lo : 0u) + __reduce_or_sync(0xffffffffu, w == 0 ? own_hi : 0u);
            sm[1] = __reduce_or_sync(0xffffffffu, w == 1 ? own_lo : 0u) + __reduce_or_sync(0xffffffffu, w == 1 ? own_hi : 0u);
            sm[2] = __reduce_or_sync(0xffffffffu, all_lo) + __reduce_or_sync(0xffffffffu, all_hi);
#pragma unroll
            for (int ww = 0; ww < 3; ww++)
#pragma unroll
                for (int c = 0; c < 3; c++) m2[ww][c] = (int)((sm[ww] >> (10 * c)) & 1023u);
        } else {
        {
            uint4 *h4 = (uint4 *)&hist[0][0][0];
#pragma unroll
            for (int k = 0; k < 12; k++) h4[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        for (uint32_t i0 = 0; i0 < nt; i0 += 128) {
            // four independent 16 B loads in flight per lane
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + 32 * u + lane;
                r[u] = i < nt ? sorted[b0 + i] : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + 32 * u + lane;
                if (i < nt) {
                    const int w = i >= np ? 1 : 0;
                    const uint32_t c = r[u].z;
                    atomicAdd(&hist[w][0][c & 255u], 1u);
                    atomicAdd(&hist[w][1][(c >> 8) & 255u], 1u);
                    atomicAdd(&hist[w][2][(c >> 16) & 255u], 1u);
                    acc_record(a, r[u], w, road_cls, v0, v1, v2, v3, want_max);
                }
            }
        }
        }   // histogram path: records
#pragma unroll
        for (int w = 0; w < 2; w++) {
            a.n_road[w] = __reduce_add_sync(0xffffffffu, a.n_road[w]);
            a.n_veh[w] = __reduce_add_sync(0xffffffffu, a.n_veh[w]);
            a.fx_hi[w] = warp_sum_i64(a.fx_hi[w]);
            a.fx_lo[w] = warp_sum_i64(a.fx_lo[w]);
            a.ext_z[w] = want_max ? warp_max_f64(a.ext_z[w]) : warp_min_f64(a.ext_z[w]);
        }
        __syncwarp();
        if (!medium) {
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
        }   // histogram path: medians
        __syncwarp();
        // The windows' statistics are parked in shared memory and finalised 32 at a time with one
        // lane per (cell, window): done on the spot, three lanes ran the float64 divisions and the
        // sigmoid of a cell while 29 waited — 28 % of this kernel's instructions on the 24 M-point
        // window.  Empty windows only store constants and are not parked.
        {
            const uint32_t k0 = np != 0 ? 1u : 0u, k1 = nf != 0 ? 1u : 0u;
            if (lane < 3) {
                const int w = (int)lane;
                const uint32_t nw = (w == 0) ? np : (w == 1) ? nf : nt;
                if (nw == 0) {
                    store_empty<F64OUT>(out16, out64, (((int64_t)var * 3 + w) * 7) * PP + cell_in, PP, consts[var]);
                } else {
                    const uint32_t e = pend_n + (w == 0 ? 0u : w == 1 ? k0 : k0 + k1);
                    const int ww = w & 1;
                    pd.nw[e] = nw;
                    pd.gc[e] = gc;
                    const uint32_t mp0 = (uint32_t)m2[0][0] | ((uint32_t)m2[0][1] << 10) | ((uint32_t)m2[0][2] << 20),
                                   mp1 = (uint32_t)m2[1][0] | ((uint32_t)m2[1][1] << 10) | ((uint32_t)m2[1][2] << 20),
                                   mp2 = (uint32_t)m2[2][0] | ((uint32_t)m2[2][1] << 10) | ((uint32_t)m2[2][2] << 20);
                    pd.med[e] = ((w == 0) ? mp0 : (w == 1) ? mp1 : mp2) | ((uint32_t)w << 30);
                    if (w < 2) {
                        pd.nr[e] = ww ? a.n_road[1] : a.n_road[0];
                        pd.nv[e] = ww ? a.n_veh[1] : a.n_veh[0];
                        pd.fx_hi[e] = ww ? a.fx_hi[1] : a.fx_hi[0];
                        pd.fx_lo[e] = ww ? a.fx_lo[1] : a.fx_lo[0];
                        pd.ez[e] = ww ? a.ext_z[1] : a.ext_z[0];
                    } else {
                        pd.nr[e] = a.n_road[0] + a.n_road[1];
                        pd.nv[e] = a.n_veh[0] + a.n_veh[1];
                        pd.fx_hi[e] = a.fx_hi[0] + a.fx_hi[1];
                        pd.fx_lo[e] = a.fx_lo[0] + a.fx_lo[1];
                        pd.ez[e] = want_max ? fmax(a.ext_z[0], a.ext_z[1]) : fmin(a.ext_z[0], a.ext_z[1]);
                    }
                }
            }
            pend_n += k0 + k1 + 1u;      // a queued cell is never empty: its full window always counts
        }
        __syncwarp();
        if (pend_n > 29u) {
            big_flush<F64OUT>(pd, pend_n, lane, params, consts, lut, PP, intensity_div, out16, out64);
            pend_n = 0;
            __syncwarp();
        }
    }   // cells of the batch
    }
    if (pend_n) big_flush<F64OUT>(pd, pend_n, lane, params, consts, lut, PP, intensity_div, out16, out64);
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

int pcacc_init_tables(pcacc_t h) {
    PCACC_CUDA(h, cudaMalloc(&h->d_rgb_lut, LUT_TOTAL * (sizeof(double) + sizeof(__half))));
    k_rgb_lut<<<(LUT_TOTAL + 127) / 128, 128>>>(h->d_rgb_lut);
    k_rgb_lut16<<<(LUT_TOTAL + 127) / 128, 128>>>(h->d_rgb_lut);
    PCACC_CUDA(h, cudaGetLastError());
    PCACC_CUDA(h, cudaDeviceSynchronize());
    return PCACC_OK;
}

extern "C" int pcacc_rasterise(pcacc_t h, const pcacc_bev_params *params, int n_variants, int P,
                               void *out_f16_dev, double *out_f64_dev, int32_t *dbg_cell_dev,
                               void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (!params || n_variants <= 0 || P <= 0 || P > 16384 || !out_f16_dev)
        return pcacc_fail(h, PCACC_ERR_ARG, "bad rasterise arguments");
    if (h->inten_div == 0.0) h->inten_div = 1.0;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int64_t PP = (int64_t)P * P;

    const bool want_f64_early = out_f64_dev != nullptr;
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
        // workspace layout: [counters | append/replay/big-queue counters] are zeroed per call
        size_t o_counts = 0;
        size_t o_cnt2 = align_up(o_counts + (size_t)n_keys * 4, 256);  // 4 x u64 counters
        size_t o_ftiles = align_up(o_cnt2 + 32, 256);                  // per-frame tile counts
        size_t o_key = align_up(o_ftiles + (size_t)BIN_MAXF * 4, 256);
        // the candidate list (8 B) is dead once k_bev_bin has run: it shares the space of `sorted`
        size_t o_rec = align_up(o_key + (size_t)cap * 4, 256);
        size_t o_sorted = align_up(o_rec + (size_t)cap * 16, 256);
        size_t o_consts = align_up(o_sorted + (size_t)cap * 16, 256);
        size_t o_fvar = align_up(o_consts + (size_t)nv * sizeof(BevConsts), 256);
        size_t o_big = align_up(o_fvar + (size_t)(fhi > flo ? fhi - flo : 0) * nv * sizeof(FrameVar), 256);
        // Cells above small_t points are queued for pass B (one warp per cell, dynamically
        // balanced); the rest is reduced 32 cells per warp by pass A.  (Measured on the
        // long-horizon window, one variant = 512 chunk warps: small_t = 4 takes pass A from 35 to
        // 12 us but pass B from 94 to 128 us — a warp per 5..15-point cell costs more than the
        // group reduction it replaces — so the threshold stays at SMALL_T for every launch.)
        const bool chunked = !want_f64_early && PP % CH_CELLS == 0 && !h->reduce_strips &&
                             (PP / CH_CELLS + CH_WARPS - 1) / CH_WARPS <= 65535;   // grid.y
        const uint32_t small_t = (uint32_t)SMALL_T;
        // so the queue never exceeds cap / (small_t + 1)
        int64_t big_cap = cap / (small_t + 1) + 1;
        if (big_cap > (int64_t)nv * PP) big_cap = (int64_t)nv * PP;
        size_t total = align_up(o_big + (size_t)big_cap * 4, 256);
        int rc = ensure_ws(h, total);
        if (rc) return rc;
        char *ws = (char *)h->d_ws;
        uint32_t *counts = (uint32_t *)(ws + o_counts);
        unsigned long long *ctr = (unsigned long long *)(ws + o_cnt2);
        uint32_t *big_count = (uint32_t *)(ctr + 2), *big_next = (uint32_t *)(ctr + 3);
        BevConsts *d_consts = (BevConsts *)(ws + o_consts);
        uint32_t *big_list = (uint32_t *)(ws + o_big);
        PCACC_CUDA(h, cudaMemsetAsync(ws, 0, o_key, st));

        void *d_params = nullptr;
        rc = pcacc_arena_put(h, params + v0, (size_t)nv * sizeof(pcacc_bev_params), &d_params, st);
        if (rc) return rc;

        // per-variant constants (empty-window planes, P/view): written by k_bev_cull when
        // there are frames to visit
        const bool visit = cap > 0 && fhi > flo && max_cnt > 0;
        if (!visit) {
            h->launches[PCACC_K_REDUCE]++;
            k_bev_consts<<<(nv + 31) / 32, 32, 0, st>>>((const pcacc_bev_params *)d_params, nv, P, d_consts);
            PCACC_CUDA(h, cudaGetLastError());
        }

        const bool want_f64 = out_f64_dev != nullptr;
        __half *o16 = (__half *)out_f16_dev + (int64_t)v0 * 21 * PP;
        double *o64 = want_f64 ? out_f64_dev + (int64_t)v0 * 21 * PP : nullptr;

        if (visit) {
            BinArgs a;
            a.ring = h->ring;
            a.frame_off = h->d_frame_off;
            a.frame_cnt = h->d_frame_cnt;
            a.frame_epoch = h->d_frame_epoch;
            a.comp = h->d_comp;
            a.chain = h->d_chain;
            a.cull = h->d_cull;
            a.aabb = h->d_aabb;
            a.max_frames = h->max_frames;
            a.frame_lo = flo;
            a.epoch_now = h->rebase_epoch;
            a.params = (const pcacc_bev_params *)d_params;
            a.consts = d_consts;
            a.n_var = nv;
            a.P = P;
            a.counts = counts;
            a.frame_tiles = (uint32_t *)(ws + o_ftiles);
            a.fvar = (FrameVar *)(ws + o_fvar);
            a.cand_gi = (uint32_t *)(ws + o_sorted);
            a.cand_meta = (uint32_t *)(ws + o_sorted + (size_t)cap * 4);
            a.tmp_key = (uint32_t *)(ws + o_key);
            a.tmp_rec = (uint4 *)(ws + o_rec);
            a.n_append = ctr;
            a.n_replay = ctr + 1;
            a.cap = cap;
            a.dbg_cell = (v0 == 0) ? dbg_cell_dev : nullptr;
            a.flags = h->d_flags;
            int64_t nf = fhi - flo;
            if (nf >= BIN_MAXF)   // k_bev_classify's tile prefix has BIN_MAXF entries incl. the total
                return pcacc_fail(h, PCACC_ERR_ARG, "more than %d frames in one rasterise call", BIN_MAXF - 1);
            a.frame_lo = flo;
            a.n_frames = (int)nf;
            size_t pe = pcacc_prof_begin(h, PCACC_K_CLASSIFY, st);
            k_bev_cull<<<(a.n_frames * nv + 127) / 128, 128, 0, st>>>(a);
            PCACC_CUDA(h, cudaGetLastError());
            h->launches[PCACC_K_CLASSIFY]++;
            if (nv >= CLS_MV_MIN && !h->classify_single)
                k_bev_classify_mv<<<h->n_sm * h->cls_mult, BIN_BLOCK, 0, st>>>(a);
            else
                k_bev_classify<<<h->n_sm * h->cls_mult, BIN_BLOCK, 0, st>>>(a);
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_CLASSIFY, pe, st);
            pe = pcacc_prof_begin(h, PCACC_K_BIN, st);
            // grid in multiples of the SM count (3 blocks are resident per SM): two full waves for a batch of
            // variants (23.0 vs 25.2 us per scene with 8), three for few variants over many points (88 vs 92 us
            // on the long-horizon window, 97 with 6); PCACC_BIN_MULT overrides
            const int bmult = h->bin_mult ? h->bin_mult : (nv >= CLS_MV_MIN ? 6 : 9);
            k_bev_bin<<<h->n_sm * bmult, 256, 0, st>>>(a);
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_BIN, pe, st);
            // scan
            int64_t tiles = (n_keys + SCAN_TILE - 1) / SCAN_TILE;
            rc = pcacc_ensure_tiles(h, tiles);
            if (rc) return rc;
            ScanLB lb;
            lb.state = h->d_tile_state;
            lb.ticket = h->d_ticket;
            lb.epoch = pcacc_next_epoch(h);
            lb.n_tiles = (uint32_t)tiles;
            pe = pcacc_prof_begin(h, PCACC_K_SCAN, st);
            k_scan<<<(unsigned)tiles, SCAN_BLOCK, 0, st>>>(counts, n_keys, lb);
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_SCAN, pe, st);
            // scatter
            int64_t sb = (cap + 255) / 256;
            if (sb > (int64_t)h->n_sm * 16) sb = (int64_t)h->n_sm * 16;
            pe = pcacc_prof_begin(h, PCACC_K_SCATTER, st);
            k_bev_scatter<<<(unsigned)sb, 256, 0, st>>>(counts, a.tmp_key, a.tmp_rec, ctr, cap,
                                                        (uint4 *)(ws + o_sorted));
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_SCATTER, pe, st);
        }
        // reduce + finalise (also correct on all-zero counters: every cell empty)
        const dim3 blocks((unsigned)(((PP + 31) / 32 + RED_WARPS - 1) / RED_WARPS), (unsigned)nv);
        size_t pr = pcacc_prof_begin(h, PCACC_K_REDUCE, st);
        if (want_f64)
            k_bev_reduce<true><<<blocks, RED_WARPS * 32, 0, st>>>(
                counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params, d_consts,
                h->d_rgb_lut, nv, P, h->inten_div, big_list, big_count, o16, o64);
        else if (chunked)
            k_bev_reduce_chunk<<<dim3((unsigned)nv, (unsigned)((PP / CH_CELLS + CH_WARPS - 1) / CH_WARPS)),
                                 CH_WARPS * 32, 0, st>>>(
                counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params, d_consts,
                h->d_rgb_lut, P, h->inten_div, small_t, big_list, big_count, o16);
        else
            k_bev_reduce<false><<<blocks, RED_WARPS * 32, 0, st>>>(
                counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params, d_consts,
                h->d_rgb_lut, nv, P, h->inten_div, big_list, big_count, o16, nullptr);
        PCACC_CUDA(h, cudaGetLastError());
        pcacc_prof_end(h, PCACC_K_REDUCE, pr, st);
        if (cap > (int64_t)small_t) {
            int64_t bb = (big_cap + REDB_WARPS - 1) / REDB_WARPS;
            if (bb > (int64_t)h->n_sm * 8) bb = (int64_t)h->n_sm * 8;
            pr = pcacc_prof_begin(h, PCACC_K_REDUCE_BIG, st);
            if (want_f64)
                k_bev_reduce_big<true><<<(unsigned)bb, REDB_WARPS * 32, 0, st>>>(
                    counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params,
                    d_consts, h->d_rgb_lut, P, h->inten_div, big_list, big_count, big_next, o16, o64);
            else
                k_bev_reduce_big<false><<<(unsigned)bb, REDB_WARPS * 32, 0, st>>>(
                    counts, (const uint4 *)(ws + o_sorted), (const pcacc_bev_params *)d_params,
                    d_consts, h->d_rgb_lut, P, h->inten_div, big_list, big_count, big_next, o16,
                    nullptr);
            PCACC_CUDA(h, cudaGetLastError());
            pcacc_prof_end(h, PCACC_K_REDUCE_BIG, pr, st);
        }
        // keep the counters of the last group for pcacc_raster_stats
        PCACC_CUDA(h, cudaMemcpyAsync(h->d_rstats + 1, ctr, 16, cudaMemcpyDeviceToDevice, st));
        h->last_visit_ub = cap;
    }
    return PCACC_OK;
}

// ---------------------------------------------------------------------------
// polynomial warp of finished planes (a separable gather)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_warp_planes(const __half *__restrict__ in, __half *__restrict__ out, const int32_t *__restrict__ imap,
              const int32_t *__restrict__ jmap, int n_planes, int P, int64_t total) {
    const int64_t PP = (int64_t)P * P;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int iw = (int)(t % P);
        const int jw = (int)((t / P) % P);
        const int64_t bp = t / PP;             // bev * n_planes + plane
        const int b = (int)(bp / n_planes);
        const int i = imap[(int64_t)b * P + iw], j = jmap[(int64_t)b * P + jw];
        out[t] = in[bp * PP + (int64_t)j * P + i];
    }
}

extern "C" int pcacc_warp_planes(pcacc_t h, const void *in_f16_dev, void *out_f16_dev, int n_bevs,
                                 int n_planes, int P, const int32_t *imap, const int32_t *jmap,
                                 void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (!in_f16_dev || !out_f16_dev || in_f16_dev == out_f16_dev || n_bevs <= 0 || n_planes <= 0 ||
        P <= 0 || !imap || !jmap)
        return pcacc_fail(h, PCACC_ERR_ARG, "bad warp_planes arguments");
    for (int64_t k = 0; k < (int64_t)n_bevs * P; k++)
        if (imap[k] < 0 || imap[k] >= P || jmap[k] < 0 || jmap[k] >= P)
            return pcacc_fail(h, PCACC_ERR_ARG, "warp index outside [0, P)");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    std::vector<int32_t> buf((size_t)2 * n_bevs * P);
    memcpy(buf.data(), imap, (size_t)n_bevs * P * 4);
    memcpy(buf.data() + (size_t)n_bevs * P, jmap, (size_t)n_bevs * P * 4);
    void *dev = nullptr;
    int rc = pcacc_arena_put(h, buf.data(), buf.size() * 4, &dev, st);
    if (rc) return rc;
    const int64_t total = (int64_t)n_bevs * n_planes * P * P;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)h->n_sm * 16) blocks = (int64_t)h->n_sm * 16;
    h->launches[PCACC_K_EXPORT]++;
    k_warp_planes<<<(unsigned)blocks, 256, 0, st>>>((const __half *)in_f16_dev, (__half *)out_f16_dev,
                                                    (const int32_t *)dev,
                                                    (const int32_t *)dev + (size_t)n_bevs * P, n_planes,
                                                    P, total);
    PCACC_CUDA(h, cudaGetLastError());
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
