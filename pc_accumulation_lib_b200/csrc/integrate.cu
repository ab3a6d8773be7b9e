// integrate.cu — per-frame fusion kernels: projection, in-image mask, map
// gathers, class filter, pose transform and order-preserving append into the
// device-resident SoA ring; plus re-basing, dynamic flags and frame export.
//
// Reference behaviour restated (paths relative to the reference root):
//   sem_pc_accum.py:317-402   velo2frame / velo2img / gen_semantic_pc / filter_semseg_pc
//   kitti360_sem_pc_accum.py:129-156   record build (inst = 0, dyn = 0)
//   nuscenes_oracle_sem_pc_accum.py:454-501, datasets/nuscenes_utils.py:46-60,181-214
//   sem_pc_accum.py:167-183   update_sem_pcs
//   nuscenes_oracle_sem_pc_accum.py:223-230,243-250   dynamic flags
#include <math.h>
#include <algorithm>
#include <utility>
#include <vector>
#include <string.h>

#include "common.cuh"

#define IBLOCK 256

struct FrameSlots {
    int64_t *frame_off, *frame_cnt, *frame_epoch;
    double *comp;  // this frame's composed matrix (12 doubles), reset to identity
    double *cull;  // this frame's total transform since insertion (12 doubles), reset to identity
    unsigned long long *aabb;       // this frame's bounding box (already initialised to empty)
    unsigned long long *aabb_next;  // the next slot's box: emptied by the tile that ends this frame
    int slot, next_slot;
    int64_t base_override;  // >= 0: use this ring offset instead of frame_off[slot]
    int64_t epoch;
    int64_t capacity;
    int64_t slot_id;  // absolute frame id
    int write_next;  // 0 inside a batch (the next frame has its own fixed base)
};

struct Filters {
    int32_t v[PCACC_MAX_FILTERS];
    int n;
};

struct LookBack {
    unsigned long long *state;
    uint32_t *ticket;
    uint32_t epoch;
    uint32_t n_tiles;
};

// All PCACC_MAX_FILTERS slots are compared (statically indexed, predicated on k < n): a loop bounded
// by f.n indexes the by-value kernel parameter dynamically, which makes the compiler copy the struct
// to local memory and read it back through L1 for every point.
__device__ __forceinline__ bool class_filtered(const Filters &f, int cls) {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < PCACC_MAX_FILTERS; k++) hit |= (k < f.n) && (cls == f.v[k]);
    return hit;
}

// the tile that ends the frame records its size and the next frame's offset
__device__ __forceinline__ void finish_frame(const FrameSlots &fs, int64_t base, uint32_t total) {
    fs.frame_off[fs.slot] = base;
    fs.frame_cnt[fs.slot] = (int64_t)total;
    fs.frame_epoch[fs.slot] = fs.epoch;
    int64_t nxt = base + (((int64_t)total + PCACC_ALIGN_PTS - 1) / PCACC_ALIGN_PTS) * PCACC_ALIGN_PTS;
    if (fs.write_next) {
        fs.frame_off[fs.next_slot] = nxt;
        for (int k = 0; k < 6; k++) fs.aabb_next[k] = k < 3 ? AABB_EMPTY_MIN : AABB_EMPTY_MAX;
    }
    for (int k = 0; k < 12; k++) {
        const double id = (k == 0 || k == 5 || k == 10) ? 1.0 : 0.0;
        fs.comp[k] = id;
        fs.cull[k] = id;
    }
}

__device__ __forceinline__ int64_t frame_base(const FrameSlots &fs) {
    return fs.base_override >= 0 ? fs.base_override : fs.frame_off[fs.slot];
}

template <int DT>
__device__ __forceinline__ int load_class(const void *map, int64_t pix, int K) {
    if (DT == PCACC_SEM_U8) return (int)((const uint8_t *)map)[pix];
    if (DT == PCACC_SEM_I32) return ((const int32_t *)map)[pix];
    if (DT == PCACC_SEM_I16) return (int)((const int16_t *)map)[pix];
    if (DT == PCACC_SEM_I64) {
        long long v = ((const long long *)map)[pix];
        return (v < -2147483647ll || v > 2147483647ll) ? -2147483647 : (int)v;
    }
    if (DT == PCACC_SEM_SAMPLED) return ((const int32_t *)map)[2 * pix + 1];
    // float32 probabilities: first index of the maximum (np.argmax).  All loads of the row are
    // issued before the first compare.  A row of K floats starts at byte 4*K*pix: when K is a
    // multiple of 4 (a producer that pads 19 classes to 20 with a zero channel) and the map is
    // 16-byte aligned the row is read with K/4 vector loads, otherwise with K scalar loads
    // (a 76-byte row straddles three or four 32-byte sectors either way; L1 keeps them).
    const float *p = (const float *)map + pix * K;
    float best;
    int arg = 0;
    if ((K & 3) == 0 && K <= 32 && ((uintptr_t)map & 15u) == 0) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (4 * k < K) v[k] = __ldg((const float4 *)p + k);
        best = v[0].x;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (4 * k < K) {
                const float c[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (c[j] > best) {
                        best = c[j];
                        arg = 4 * k + j;
                    }
                }
            }
        }
        return arg;
    }
    if (K == 19) {      // the semseg head of the reference (19 Cityscapes classes): fully unrolled
        float c[19];
#pragma unroll
        for (int k = 0; k < 19; k++) c[k] = __ldg(p + k);
        best = c[0];
#pragma unroll
        for (int k = 1; k < 19; k++) {
            if (c[k] > best) {
                best = c[k];
                arg = k;
            }
        }
        return arg;
    }
    best = p[0];
    for (int k = 1; k < K; k++) {
        float v = p[k];
        if (v > best) {
            best = v;
            arg = k;
        }
    }
    return arg;
}

// projection of one point: sem_pc_accum.py:358-396
struct Proj {
    double u, v, d;
    bool in_img;
};
__device__ __forceinline__ Proj project_point(const double *P, float x, float y, float z, int img_h,
                                              int img_w, double max_depth) {
    double X0, X1, X2;
    affine_chain(P, 4, (double)x, (double)y, (double)z, X0, X1, X2);
    double d = X2;
    if (d == 0.0) d = -1e-6;
    double ad = fabs(d);
    Proj r;
    r.u = rint_even(__ddiv_rn(X0, ad));
    r.v = rint_even(__ddiv_rn(X1, ad));
    r.d = d;
    // integer compares of the reference done on the (integral) doubles: NaN /
    // inf / out-of-int64 values fail them exactly as astype(int) garbage does
    r.in_img = (r.u >= 0.0) && (r.u < (double)img_w) && (r.v >= 0.0) && (r.v < (double)img_h) &&
               (d > 0.0) && (d < max_depth);
    return r;
}

struct PMat {
    double m[12];
};

// ---------------------------------------------------------------------------
// stand-alone projection (all points)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(IBLOCK)
k_project(const float *__restrict__ pts, int64_t n, int stride, PMat P, int img_h, int img_w,
          double max_depth, int32_t *__restrict__ u, int32_t *__restrict__ v,
          uint8_t *__restrict__ mask) {
    int64_t i = (int64_t)blockIdx.x * IBLOCK + threadIdx.x;
    if (i >= n) return;
    const float *p = pts + i * stride;
    Proj r = project_point(P.m, p[0], p[1], p[2], img_h, img_w, max_depth);
    u[i] = sat_i32(r.u);
    v[i] = sat_i32(r.v);
    mask[i] = r.in_img ? 1 : 0;
}

// ---------------------------------------------------------------------------
// gen_semantic_pc: generic K-channel gather into (M,4+K) float64
// ---------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(IBLOCK)
k_gen_semantic_pc(const float4 *__restrict__ pts, int64_t n, PMat P, const void *__restrict__ map,
                  int img_h, int img_w, int K, double *__restrict__ out,
                  int64_t *__restrict__ n_kept, LookBack lb) {
    __shared__ uint32_t s_warp[IBLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    int64_t i = (int64_t)tile * IBLOCK + threadIdx.x;
    bool keep = false;
    float4 p = make_float4(0, 0, 0, 0);
    Proj r;
    r.u = r.v = 0;
    if (i < n) {
        p = pts[i];
        r = project_point(P.m, p.x, p.y, p.z, img_h, img_w, INFINITY);
        keep = r.in_img;
    }
    uint32_t tile_end;
    uint32_t rank = compact_rank<IBLOCK>(keep, lb.state, lb.epoch, tile, s_warp, &tile_end);
    if (keep) {
        double *o = out + (int64_t)rank * (4 + K);
        o[0] = (double)p.x;
        o[1] = (double)p.y;
        o[2] = (double)p.z;
        o[3] = (double)p.w;
        int64_t pix = ((int64_t)r.v * img_w + (int64_t)r.u) * K;
        for (int k = 0; k < K; k++) {
            double f;
            if (DT == PCACC_SEM_U8) f = (double)((const uint8_t *)map)[pix + k];
            else if (DT == PCACC_SEM_I32) f = (double)((const int32_t *)map)[pix + k];
            else if (DT == PCACC_SEM_I64) f = (double)((const long long *)map)[pix + k];
            else if (DT == PCACC_SEM_I16) f = (double)((const int16_t *)map)[pix + k];
            else f = (double)((const float *)map)[pix + k];
            o[4 + k] = f;
        }
    }
    if (tile == lb.n_tiles - 1 && threadIdx.x == 0) *n_kept = (int64_t)tile_end;
}

// ---------------------------------------------------------------------------
// velo2img as the reference returns it (sem_pc_accum.py:367-402): the rows [pc_velo, u, v] of the
// points that fall into the image, float64, in input order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(IBLOCK)
k_velo2img(const float *__restrict__ pts, int64_t n, int stride, PMat P, int img_h, int img_w,
           double max_depth, double *__restrict__ out, int64_t *__restrict__ n_kept, LookBack lb) {
    __shared__ uint32_t s_warp[IBLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    const uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const int64_t i = (int64_t)tile * IBLOCK + threadIdx.x;
    bool keep = false;
    Proj r;
    r.u = r.v = 0;
    if (i < n) {
        const float *p = pts + i * stride;
        r = project_point(P.m, p[0], p[1], p[2], img_h, img_w, max_depth);
        keep = r.in_img;
    }
    uint32_t tile_end;
    const uint32_t rank = compact_rank<IBLOCK>(keep, lb.state, lb.epoch, tile, s_warp, &tile_end);
    if (keep) {
        const float *p = pts + i * stride;
        double *o = out + (int64_t)rank * (stride + 2);
        for (int c = 0; c < stride; c++) o[c] = (double)p[c];
        o[stride] = r.u;       // integral doubles inside the image: what astype(int) -> float64 gives
        o[stride + 1] = r.v;
    }
    if (tile == lb.n_tiles - 1 && threadIdx.x == 0) *n_kept = (int64_t)tile_end;
}

// ---------------------------------------------------------------------------
// KITTI-360 frustum integrate (K1-K4 fused)
// ---------------------------------------------------------------------------
// KI_ITEMS points per thread (element k of thread t = point k * IBLOCK + t of the tile): a 120 000-point
// frame is 235 tiles instead of 938 — a shorter look-back chain — and every thread has its KI_ITEMS
// independent point loads and map gathers in flight at once.
#define KI_ITEMS 4
template <int DT>
__global__ void __launch_bounds__(IBLOCK)
k_integrate_frustum(const float4 *__restrict__ pts, int64_t n, PMat P,
                    const uint8_t *__restrict__ rgb, const void *__restrict__ sem, int K, int img_h,
                    int img_w, double max_depth, Filters filt, RingDev ring, FrameSlots fs,
                    LookBack lb, uint32_t *__restrict__ flags) {
    __shared__ uint32_t s_cnt[KI_ITEMS * IBLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    const uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const int64_t base = frame_base(fs);
    const int64_t i0 = (int64_t)tile * (IBLOCK * KI_ITEMS) + threadIdx.x;
    bool keep[KI_ITEMS];
    float4 p[KI_ITEMS];
    uint32_t packed[KI_ITEMS];
#pragma unroll
    for (int k = 0; k < KI_ITEMS; k++) {
        const int64_t i = i0 + k * IBLOCK;
        p[k] = i < n ? pts[i] : make_float4(0, 0, 0, 0);
        keep[k] = false;
        packed[k] = 0;
    }
#pragma unroll
    for (int k = 0; k < KI_ITEMS; k++) {
        if (i0 + k * IBLOCK < n) {
            const Proj r = project_point(P.m, p[k].x, p[k].y, p[k].z, img_h, img_w, max_depth);
            if (r.in_img) {
                const int64_t pix = (int64_t)r.v * img_w + (int64_t)r.u;
                int cls = load_class<DT>(sem, pix, K);
                keep[k] = !class_filtered(filt, cls);
                if (keep[k]) {
                    if (cls < 0 || cls > 255) {
                        atomicOr(flags, PCACC_FLAG_ATTR_RANGE);
                        cls &= 255;
                    }
                    const uint8_t *c = rgb + pix * 3;
                    packed[k] = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) |
                                ((uint32_t)cls << 24);
                }
            }
        }
    }
    uint32_t rank[KI_ITEMS], tile_end;
    compact_rank_multi<IBLOCK, KI_ITEMS>(keep, lb.state, lb.epoch, tile, s_cnt, rank, &tile_end);
    double bb[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < KI_ITEMS; k++) {
        if (keep[k]) {
            const double x = (double)p[k].x, y = (double)p[k].y, z = (double)p[k].z;
            bb[0] = fmin(bb[0], x); bb[1] = fmin(bb[1], y); bb[2] = fmin(bb[2], z);
            bb[3] = fmax(bb[3], x); bb[4] = fmax(bb[4], y); bb[5] = fmax(bb[5], z);
            const int64_t o = base + rank[k];
            if (o < fs.capacity) {
                ring.x[o] = x;
                ring.y[o] = y;
                ring.z[o] = z;
                ring.inten[o] = p[k].w;
                ring.rgbs[o] = packed[k];
                ring.inst[o] = 0;
                ring.dyn[o] = 0;
            }
        }
    }
    aabb_update_v<IBLOCK>(fs.aabb, bb);
    if (tile == lb.n_tiles - 1 && threadIdx.x == 0) finish_frame(fs, base, tile_end);
}

// KITTI-360 use_gt_sem path: all points, class from sem_gt, rgb = 0
__global__ void __launch_bounds__(IBLOCK)
k_integrate_gt(const float4 *__restrict__ pts, int64_t n, const int16_t *__restrict__ sem_gt,
               Filters filt, RingDev ring, FrameSlots fs, LookBack lb, uint32_t *__restrict__ flags) {
    __shared__ uint32_t s_cnt[KI_ITEMS * IBLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    const uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    const int64_t base = frame_base(fs);
    const int64_t i0 = (int64_t)tile * (IBLOCK * KI_ITEMS) + threadIdx.x;
    bool keep[KI_ITEMS];
    float4 p[KI_ITEMS];
    int cls[KI_ITEMS];
#pragma unroll
    for (int k = 0; k < KI_ITEMS; k++) {
        const int64_t i = i0 + k * IBLOCK;
        const bool in = i < n;
        p[k] = in ? pts[i] : make_float4(0, 0, 0, 0);
        cls[k] = in ? (int)sem_gt[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < KI_ITEMS; k++) {
        keep[k] = (i0 + k * IBLOCK < n) && !class_filtered(filt, cls[k]);
        if (keep[k] && (cls[k] < 0 || cls[k] > 255)) {
            atomicOr(flags, PCACC_FLAG_ATTR_RANGE);
            cls[k] &= 255;
        }
    }
    uint32_t rank[KI_ITEMS], tile_end;
    compact_rank_multi<IBLOCK, KI_ITEMS>(keep, lb.state, lb.epoch, tile, s_cnt, rank, &tile_end);
    double bb[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < KI_ITEMS; k++) {
        if (keep[k]) {
            const double x = (double)p[k].x, y = (double)p[k].y, z = (double)p[k].z;
            bb[0] = fmin(bb[0], x); bb[1] = fmin(bb[1], y); bb[2] = fmin(bb[2], z);
            bb[3] = fmax(bb[3], x); bb[4] = fmax(bb[4], y); bb[5] = fmax(bb[5], z);
            const int64_t o = base + rank[k];
            if (o < fs.capacity) {
                ring.x[o] = x;
                ring.y[o] = y;
                ring.z[o] = z;
                ring.inten[o] = p[k].w;
                ring.rgbs[o] = (uint32_t)cls[k] << 24;
                ring.inst[o] = 0;
                ring.dyn[o] = 0;
            }
        }
    }
    aabb_update_v<IBLOCK>(fs.aabb, bb);
    if (tile == lb.n_tiles - 1 && threadIdx.x == 0) finish_frame(fs, base, tile_end);
}

// ---------------------------------------------------------------------------
// nuScenes oracle-pose integrate (a9-a11)
// ---------------------------------------------------------------------------
struct CamMaps {
    const uint8_t *rgb[PCACC_MAX_CAMS];
    const void *sem[PCACC_MAX_CAMS];
    int n;
};
struct TMat {
    double m[16];
};

// Records kernels: 128 threads x 8 points.  A scene's 40 sweeps x 34 tiles are 1360 blocks;
// with 256-thread blocks only 8 fit an SM (1184 slots on 148 SMs) and the last 176 blocks
// ran as a second, nearly empty wave.
#define RBLOCK 128
#define REC_ITEMS 8                       /* points per thread in the records kernels */
#define REC_TILE (RBLOCK * REC_ITEMS)

// One tile (1024 consecutive points) of the nuScenes record path.  Item k of thread t is
// point tile*1024 + k*128 + t.  Only the two pixel coordinates are read for every point;
// xyz / intensity / instance of the (few) kept points are fetched after the gather.
template <int DT, int ITEMS>
__device__ __forceinline__ void records_tile(const double *__restrict__ pc,
                                             const long long *__restrict__ cam_idx, int64_t n,
                                             const CamMaps &maps, int img_h, int img_w,
                                             const double *__restrict__ T, const Filters &filt,
                                             const RingDev &ring, const FrameSlots &fs,
                                             unsigned long long *state, uint32_t epoch, uint32_t tile,
                                             uint32_t n_tiles, int64_t base,
                                             uint32_t *__restrict__ flags, uint32_t *s_cnt) {
    bool keep[ITEMS];
    uint32_t packed[ITEMS];
    long long cam[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int64_t i = (int64_t)tile * (RBLOCK * ITEMS) + k * RBLOCK + threadIdx.x;
        // sampled source: the rows are the visible points already (camera 0 of 1)
        cam[k] = i < n ? (DT == PCACC_SEM_SAMPLED ? 0ll : cam_idx[i]) : -1;
        keep[k] = false;
        packed[k] = 0;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        if (cam[k] >= 0 && cam[k] < maps.n) {
            const int64_t i = (int64_t)tile * (RBLOCK * ITEMS) + k * RBLOCK + threadIdx.x;
            const double uf = pc[i * 7 + 4], vf = pc[i * 7 + 5];
            // pts_feat_from_img bounds assertion, datasets/nuscenes_utils.py:190-195
            const bool inside = (uf > 1.0) && (uf < (double)img_w - 1.0) && (vf > 1.0) &&
                                (vf < (double)img_h - 1.0);
            if (!inside) {
                atomicOr(flags, PCACC_FLAG_UV_OUT_OF_IMAGE);
            } else {
                int cls;
                uint32_t rgbp = 0;
                int64_t pix = 0;
                if (DT == PCACC_SEM_SAMPLED) {
                    const uint2 smp = ((const uint2 *)maps.sem[0])[i];
                    rgbp = smp.x;
                    cls = (int)smp.y;
                } else {
                    pix = (int64_t)rint_even(vf) * img_w + (int64_t)rint_even(uf);
                    cls = load_class<DT>(maps.sem[cam[k]], pix, 1);
                }
                keep[k] = (cls >= 0) && !class_filtered(filt, cls);
                if (keep[k]) {
                    if (cls > 255) {
                        atomicOr(flags, PCACC_FLAG_ATTR_RANGE);
                        cls &= 255;
                    }
                    if (DT != PCACC_SEM_SAMPLED) {
                        const uint8_t *c = maps.rgb[cam[k]] + pix * 3;
                        rgbp = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
                    }
                    packed[k] = (rgbp & 0xffffffu) | ((uint32_t)cls << 24);
                }
            }
        }
    }
    uint32_t rank[ITEMS], tile_end;
    compact_rank_multi<RBLOCK, ITEMS>(keep, state, epoch, tile, s_cnt, rank, &tile_end);
    double bb[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        if (keep[k]) {
            const int64_t i = (int64_t)tile * (RBLOCK * ITEMS) + k * RBLOCK + threadIdx.x;
            const double *row = pc + i * 7;
            double wx, wy, wz;
            affine_chain(T, 4, row[0], row[1], row[2], wx, wy, wz);
            bb[0] = fmin(bb[0], wx); bb[1] = fmin(bb[1], wy); bb[2] = fmin(bb[2], wz);
            bb[3] = fmax(bb[3], wx); bb[4] = fmax(bb[4], wy); bb[5] = fmax(bb[5], wz);
            const int64_t o = base + rank[k];
            if (o < fs.capacity) {
                ring.x[o] = wx;
                ring.y[o] = wy;
                ring.z[o] = wz;
                const double inten = row[3], inst_f = row[6];
                const float fi = (float)inten;
                if ((double)fi != inten) atomicOr(flags, PCACC_FLAG_INTENSITY_F32);
                ring.inten[o] = fi;
                ring.rgbs[o] = packed[k];
                const int32_t ii = sat_i32(inst_f);
                if ((double)ii != inst_f) atomicOr(flags, PCACC_FLAG_ATTR_RANGE);
                ring.inst[o] = ii;
                ring.dyn[o] = 0;
            }
        }
    }
    aabb_update_v<RBLOCK>(fs.aabb, bb);
    if (tile == n_tiles - 1 && threadIdx.x == 0) finish_frame(fs, base, tile_end);
}

// Per-sweep launch: a sweep of 34,688 points is 34 tiles of 1024 — 34 blocks, each a chain of
// dependent reads (cam_idx -> u, v -> class -> rgb / row) that go over the bus when the arrays are
// page-locked host memory.  Two points per thread (256-point tiles, 136 blocks) put four times as
// many chains in flight; the batched kernel below keeps 8 points per thread (its 40 sweeps fill
// the machine already).
#define REC_ITEMS_SWEEP 2
template <int DT>
__global__ void __launch_bounds__(RBLOCK)
k_integrate_records(const double *__restrict__ pc, const long long *__restrict__ cam_idx, int64_t n,
                    CamMaps maps, int img_h, int img_w, TMat T, Filters filt, RingDev ring,
                    FrameSlots fs, LookBack lb, uint32_t *__restrict__ flags) {
    __shared__ uint32_t s_cnt[REC_ITEMS_SWEEP * RBLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    const uint32_t tile = lb_take_ticket(lb.ticket, lb.n_tiles, &s_tile);
    records_tile<DT, REC_ITEMS_SWEEP>(pc, cam_idx, n, maps, img_h, img_w, T.m, filt, ring, fs, lb.state, lb.epoch,
                                      tile, lb.n_tiles, frame_base(fs), flags, s_cnt);
}

// ---------------------------------------------------------------------------
// batched nuScenes integrate: every sweep of a scene in ONE launch
// (blockIdx.y = sweep).  Each sweep is still its own frame with its own
// order-preserving compaction; frames sit at fixed upper-bound offsets so no
// sweep waits for another sweep's count.
// ---------------------------------------------------------------------------
struct SweepDesc {
    const double *pc;
    const long long *cam;
    int64_t n;
    CamMaps maps;
    TMat T;
    FrameSlots fs;
    uint32_t n_tiles;
    uint32_t state_off;   // first tile-state word of this sweep
};

// empties the bounding boxes of frames first_id+1 .. first_id+n-1 (the first one was emptied
// by the frame before it, or at creation)
__global__ void k_aabb_empty(unsigned long long *__restrict__ aabb, int64_t first_id, int n,
                             int max_frames) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n - 1) * 6) return;
    int slot = (int)((first_id + 1 + t / 6) % max_frames);
    aabb[(int64_t)slot * 6 + t % 6] = (t % 6) < 3 ? AABB_EMPTY_MIN : AABB_EMPTY_MAX;
}

template <int DT>
__global__ void __launch_bounds__(RBLOCK)
k_integrate_records_batch(const SweepDesc *__restrict__ sweeps, int img_h, int img_w, Filters filt,
                          RingDev ring, unsigned long long *__restrict__ state,
                          uint32_t *__restrict__ tickets, uint32_t epoch,
                          uint32_t *__restrict__ flags) {
    __shared__ uint32_t s_cnt[REC_ITEMS * RBLOCK / 32 + 1];
    __shared__ uint32_t s_tile;
    const SweepDesc &sw = sweeps[blockIdx.y];
    if (blockIdx.x >= sw.n_tiles) return;
    const uint32_t tile = lb_take_ticket(tickets + blockIdx.y, sw.n_tiles, &s_tile);
    records_tile<DT, REC_ITEMS>(sw.pc, sw.cam, sw.n, sw.maps, img_h, img_w, sw.T.m, filt, ring, sw.fs,
                     state + sw.state_off, epoch, tile, sw.n_tiles, sw.fs.base_override, flags, s_cnt);
}

// ---------------------------------------------------------------------------
// import of a ready-made (n,10) float64 cloud
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(IBLOCK)
k_integrate_cloud(const double *__restrict__ rec, int64_t n, RingDev ring, FrameSlots fs,
                  uint32_t *__restrict__ flags) {
    const int64_t base = frame_base(fs);
    int64_t i = (int64_t)blockIdx.x * IBLOCK + threadIdx.x;
    {
        const bool have = i < n;
        aabb_update<IBLOCK>(fs.aabb, have, have ? rec[i * 10] : 0.0, have ? rec[i * 10 + 1] : 0.0,
                            have ? rec[i * 10 + 2] : 0.0);
    }
    if (i < n) {
        const double *r = rec + i * 10;
        int64_t o = base + i;
        if (o < fs.capacity) {
            ring.x[o] = r[0];
            ring.y[o] = r[1];
            ring.z[o] = r[2];
            float fi = (float)r[3];
            if ((double)fi != r[3]) atomicOr(flags, PCACC_FLAG_INTENSITY_F32);
            ring.inten[o] = fi;
            uint32_t packed = 0;
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double c = r[4 + k];
                int ci = (int)c;
                if (!((double)ci == c) || ci < 0 || ci > 255) {
                    bad = true;
                    ci &= 255;
                }
                packed |= (uint32_t)ci << (8 * k);
            }
            ring.rgbs[o] = packed;
            int32_t ii = sat_i32(r[8]);
            if ((double)ii != r[8]) bad = true;
            ring.inst[o] = ii;
            double dy = r[9];
            if (dy != 0.0 && dy != 1.0) bad = true;
            ring.dyn[o] = (dy == 1.0) ? 1 : 0;
            if (bad) atomicOr(flags, PCACC_FLAG_ATTR_RANGE);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) finish_frame(fs, base, (uint32_t)n);
}

// ---------------------------------------------------------------------------
// re-basing
// ---------------------------------------------------------------------------
// lazy: chain[epoch] = T; comp[f] = T @ comp[f] for every live frame
__global__ void k_rebase_lazy(double *__restrict__ chain_slot, double *__restrict__ comp,
                              double *__restrict__ cull, TMat T, int first_slot, int n_live,
                              int max_frames) {
    int t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t < 12) chain_slot[t] = T.m[(t / 4) * 4 + (t % 4)];
    // both per-frame matrices take the new link: `comp` (pending, reset when materialised)
    // and `cull` (everything since insertion)
    if (t < 2 * n_live) {
        double *c = (t < n_live ? comp : cull) + (int64_t)((first_slot + (t % n_live)) % max_frames) * 12;
        double o[12];
        // 3x4 affine composition: rows of T times [c; 0 0 0 1]
        for (int r = 0; r < 3; r++) {
            for (int k = 0; k < 4; k++) {
                double a = __dmul_rn(T.m[r * 4 + 0], c[0 * 4 + k]);
                a = __fma_rn(T.m[r * 4 + 1], c[1 * 4 + k], a);
                a = __fma_rn(T.m[r * 4 + 2], c[2 * 4 + k], a);
                if (k == 3) a = __dadd_rn(a, T.m[r * 4 + 3]);
                o[r * 4 + k] = a;
            }
        }
        for (int k = 0; k < 12; k++) c[k] = o[k];
    }
}

// exact sequential replay of the re-base chain for one point
__device__ __forceinline__ void replay_chain(const double *__restrict__ chain, int max_frames,
                                             int64_t e0, int64_t e1, double &x, double &y, double &z) {
    for (int64_t e = e0; e < e1; e++) {
        const double *T = chain + (e % max_frames) * 12;
        double nx, ny, nz;
        affine_chain(T, 4, x, y, z, nx, ny, nz);
        x = nx;
        y = ny;
        z = nz;
    }
}

// materialise: apply every pending transform of each live frame in place
__global__ void __launch_bounds__(IBLOCK)
k_materialise(RingDev ring, const int64_t *__restrict__ frame_off,
              const int64_t *__restrict__ frame_cnt, const int64_t *__restrict__ frame_epoch,
              const double *__restrict__ chain, int first_slot, int max_frames, int64_t epoch_now) {
    int slot = (first_slot + blockIdx.y) % max_frames;
    int64_t cnt = frame_cnt[slot], off = frame_off[slot], e0 = frame_epoch[slot];
    if (e0 >= epoch_now) return;
    for (int64_t i = (int64_t)blockIdx.x * IBLOCK + threadIdx.x; i < cnt;
         i += (int64_t)gridDim.x * IBLOCK) {
        double x = ring.x[off + i], y = ring.y[off + i], z = ring.z[off + i];
        replay_chain(chain, max_frames, e0, epoch_now, x, y, z);
        ring.x[off + i] = x;
        ring.y[off + i] = y;
        ring.z[off + i] = z;
    }
}
__global__ void k_materialise_done(int64_t *__restrict__ frame_epoch, double *__restrict__ comp,
                                   int first_slot, int n_live, int max_frames, int64_t epoch_now) {
    int t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t >= n_live) return;
    int slot = (first_slot + t) % max_frames;
    frame_epoch[slot] = epoch_now;
    double *c = comp + (int64_t)slot * 12;
    for (int k = 0; k < 12; k++) c[k] = (k == 0 || k == 5 || k == 10) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------
// dynamic flags
// ---------------------------------------------------------------------------
// One block row per FRAME that has marks (the host groups the (frame, instance) pairs):
// every point's instance index is read once and compared with the frame's few instances.
__global__ void __launch_bounds__(IBLOCK)
k_mark_dynamic(RingDev ring, const int64_t *__restrict__ frame_off,
               const int64_t *__restrict__ frame_cnt, const int32_t *__restrict__ grp /* slot, begin, end */,
               const int32_t *__restrict__ inst_sorted) {
    const int slot = grp[3 * blockIdx.y], b = grp[3 * blockIdx.y + 1], e = grp[3 * blockIdx.y + 2];
    const int64_t cnt = frame_cnt[slot], off = frame_off[slot];
    for (int64_t i = (int64_t)blockIdx.x * IBLOCK + threadIdx.x; i < cnt;
         i += (int64_t)gridDim.x * IBLOCK) {
        const int32_t inst = ring.inst[off + i];
        bool hit = false;
        for (int k = b; k < e; k++) hit |= inst == inst_sorted[k];
        if (hit) ring.dyn[off + i] = 1;
    }
}

// ---------------------------------------------------------------------------
// export one frame as (M,10) float64
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(IBLOCK)
k_export_frame(RingDev ring, const int64_t *__restrict__ frame_off,
               const int64_t *__restrict__ frame_cnt, const int64_t *__restrict__ frame_epoch,
               const double *__restrict__ chain, int slot, int max_frames, int64_t epoch_now,
               double intensity_div, double *__restrict__ out) {
    int64_t cnt = frame_cnt[slot], off = frame_off[slot], e0 = frame_epoch[slot];
    int64_t i = (int64_t)blockIdx.x * IBLOCK + threadIdx.x;
    if (i >= cnt) return;
    double x = ring.x[off + i], y = ring.y[off + i], z = ring.z[off + i];
    replay_chain(chain, max_frames, e0, epoch_now, x, y, z);
    double *o = out + i * 10;
    uint32_t c = ring.rgbs[off + i];
    o[0] = x;
    o[1] = y;
    o[2] = z;
    o[3] = __ddiv_rn((double)ring.inten[off + i], intensity_div);
    o[4] = (double)(c & 255u);
    o[5] = (double)((c >> 8) & 255u);
    o[6] = (double)((c >> 16) & 255u);
    o[7] = (double)(c >> 24);
    o[8] = (double)ring.inst[off + i];
    o[9] = (double)ring.dyn[off + i];
}

// ===========================================================================
// host side
// ===========================================================================
static int fill_filters(pcacc_t h, const int32_t *filters, int n, Filters *f) {
    if (n < 0 || n > PCACC_MAX_FILTERS || (n > 0 && !filters))
        return pcacc_fail(h, PCACC_ERR_ARG, "n_filters must be 0..%d", PCACC_MAX_FILTERS);
    f->n = n;
    for (int k = 0; k < PCACC_MAX_FILTERS; k++) f->v[k] = k < n ? filters[k] : 0;
    return PCACC_OK;
}

static int set_inten_div(pcacc_t h, double div) {
    if (!(div > 0.0)) return pcacc_fail(h, PCACC_ERR_ARG, "intensity_div must be positive");
    if (h->next_id == h->first_id || h->inten_div == 0.0) h->inten_div = div;
    if (h->inten_div != div)
        return pcacc_fail(h, PCACC_ERR_STATE, "ring holds intensity/%g frames; cannot add intensity/%g",
                          h->inten_div, div);
    return PCACC_OK;
}

static int make_lookback(pcacc_t h, int64_t n_tiles, LookBack *lb) {
    int rc = pcacc_ensure_tiles(h, n_tiles);
    if (rc) return rc;
    lb->state = h->d_tile_state;
    lb->ticket = h->d_ticket;
    lb->epoch = pcacc_next_epoch(h);
    lb->n_tiles = (uint32_t)n_tiles;
    return PCACC_OK;
}

// Conservative extent of a live frame on the host: exact after pcacc_sync; a frame placed at a
// fixed base (batch integrate, first frame, post-wrap frame) has an exact offset and an upper
// bound of its count before that.
static inline int64_t fr_off(const FrameHost &g) { return g.exact ? g.off : (g.off >= 0 ? g.off : g.off_ub); }
static inline int64_t fr_cnt(const FrameHost &g) { return g.exact ? g.cnt : g.n_in; }
static inline int64_t align_pts(int64_t n) { return (n + PCACC_ALIGN_PTS - 1) / PCACC_ALIGN_PTS * PCACC_ALIGN_PTS; }

static bool valid_sem_dtype(int dt, bool allow_prob) {
    return dt == PCACC_SEM_U8 || dt == PCACC_SEM_I32 || dt == PCACC_SEM_I64 || dt == PCACC_SEM_I16 ||
           (allow_prob && dt == PCACC_SEM_F32_PROB);
}

// Reserve a frame slot for up to n_in records. Decides the ring offset policy
// on the host from upper bounds; syncs the table only when it has to.
// allow_sync = false: the caller has synchronised already and is registering frames whose
// kernels are not enqueued yet (batch integrate) — pcacc_sync would overwrite their host
// entries with stale device-table values, so the placement works from the host-side bounds
// (exact for every frame older than the batch, fixed base + upper-bound count inside it).
static int begin_frame(pcacc_t h, int64_t n_in, cudaStream_t st, FrameSlots *fs, int64_t *frame_id,
                       bool fixed_slot = false, bool allow_sync = true) {
    if (n_in < 0) return pcacc_fail(h, PCACC_ERR_ARG, "negative point count");
    if (n_in > h->capacity)
        return pcacc_fail(h, PCACC_ERR_CAPACITY, "frame of %lld points exceeds ring capacity %lld",
                          (long long)n_in, (long long)h->capacity);
    int n_live = (int)(h->next_id - h->first_id);
    if (n_live >= h->max_frames - 1)
        return pcacc_fail(h, PCACC_ERR_CAPACITY, "frame table full (%d live frames)", n_live);
    int64_t id = h->next_id;
    int slot = (int)(id % h->max_frames);
    FrameHost &f = h->frames[slot];
    int64_t override_base = -1;
    int64_t ub = 0;
    if (n_live == 0) {
        override_base = 0;
        ub = 0;
        h->wrapped = false;
    } else {
        FrameHost &prev = h->frames[(int)((id - 1) % h->max_frames)];
        ub = prev.off_ub + align_pts(prev.n_in);
        if (h->wrapped || ub + n_in > h->capacity) {
            // need exact numbers
            if (allow_sync) {
                int rc = pcacc_sync(h, nullptr, (void *)st);
                if (rc) return rc;
            }
            FrameHost &pv = h->frames[(int)((id - 1) % h->max_frames)];
            int64_t nxt = fr_off(pv) + align_pts(fr_cnt(pv));
            if (nxt + n_in > h->capacity) {
                nxt = 0;
                h->wrapped = true;
            }
            // the new region must not touch any live frame
            bool any_after = false;
            for (int64_t k = h->first_id; k < h->next_id; k++) {
                FrameHost &g = h->frames[(int)(k % h->max_frames)];
                const int64_t go = fr_off(g), gc = fr_cnt(g);
                if (gc > 0 && go < nxt + n_in && nxt < go + gc)
                    return pcacc_fail(h, PCACC_ERR_CAPACITY,
                                      "ring full: %lld resident points, capacity %lld",
                                      (long long)pcacc_resident_points(h), (long long)h->capacity);
                if (go >= nxt) any_after = true;
            }
            if (!any_after) h->wrapped = false;
            override_base = nxt;
            ub = nxt;
        } else if (fixed_slot) {
            override_base = ub;  // place at the upper bound: no dependence on earlier counts
        }
    }
    f.n_in = n_in;
    f.off_ub = ub;
    f.off = override_base >= 0 ? override_base : -1;
    f.cnt = -1;
    f.epoch = h->rebase_epoch;
    f.exact = false;
    fs->frame_off = h->d_frame_off;
    fs->frame_cnt = h->d_frame_cnt;
    fs->frame_epoch = h->d_frame_epoch;
    fs->comp = h->d_comp + (int64_t)slot * 12;
    fs->cull = h->d_cull + (int64_t)slot * 12;
    fs->aabb = h->d_aabb + (int64_t)slot * 6;
    fs->aabb_next = h->d_aabb + (int64_t)((id + 1) % h->max_frames) * 6;
    fs->slot = slot;
    fs->next_slot = (int)((id + 1) % h->max_frames);
    fs->base_override = override_base;
    fs->epoch = h->rebase_epoch;
    fs->capacity = h->capacity;
    fs->write_next = 1;
    fs->slot_id = id;
    h->next_id = id + 1;
    if (frame_id) *frame_id = id;
    return PCACC_OK;
}

extern "C" int pcacc_project(const float *pts_dev, int64_t n, int pts_stride, const double *P,
                             int img_h, int img_w, double max_depth, int32_t *u_dev, int32_t *v_dev,
                             uint8_t *mask_dev, void *stream) {
    if (n < 0 || !P || pts_stride < 3) return PCACC_ERR_ARG;
    if (n == 0) return PCACC_OK;
    PMat pm;
    memcpy(pm.m, P, sizeof(pm.m));
    int64_t blocks = (n + IBLOCK - 1) / IBLOCK;
    k_project<<<(unsigned)blocks, IBLOCK, 0, (cudaStream_t)stream>>>(
        pts_dev, n, pts_stride, pm, img_h, img_w, max_depth, u_dev, v_dev, mask_dev);
    return cudaGetLastError() == cudaSuccess ? PCACC_OK : PCACC_ERR_CUDA;
}

extern "C" int pcacc_gen_semantic_pc(pcacc_t h, const float *pts_dev, int64_t n, const double *P,
                                     const void *map_dev, int map_dtype, int img_h, int img_w, int K,
                                     double *out_dev, int64_t *n_kept_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || !P || K < 1) return pcacc_fail(h, PCACC_ERR_ARG, "bad gen_semantic_pc arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    if (n == 0) {
        PCACC_CUDA(h, cudaMemsetAsync(n_kept_dev, 0, sizeof(int64_t), st));
        return PCACC_OK;
    }
    PMat pm;
    memcpy(pm.m, P, sizeof(pm.m));
    int64_t tiles = (n + IBLOCK - 1) / IBLOCK;
    LookBack lb;
    int rc = make_lookback(h, tiles, &lb);
    if (rc) return rc;
    const float4 *p4 = (const float4 *)pts_dev;
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
#define LAUNCH_GSP(DT)                                                                      \
    k_gen_semantic_pc<DT><<<(unsigned)tiles, IBLOCK, 0, st>>>(p4, n, pm, map_dev, img_h, img_w, K, \
                                                               out_dev, n_kept_dev, lb)
    switch (map_dtype) {
        case PCACC_SEM_U8: LAUNCH_GSP(PCACC_SEM_U8); break;
        case PCACC_SEM_I32: LAUNCH_GSP(PCACC_SEM_I32); break;
        case PCACC_SEM_I64: LAUNCH_GSP(PCACC_SEM_I64); break;
        case PCACC_SEM_I16: LAUNCH_GSP(PCACC_SEM_I16); break;
        case PCACC_SEM_F32_PROB: LAUNCH_GSP(PCACC_SEM_F32_PROB); break;
        default: return pcacc_fail(h, PCACC_ERR_ARG, "unknown map dtype %d", map_dtype);
    }
#undef LAUNCH_GSP
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_velo2img(pcacc_t h, const float *pts_dev, int64_t n, int pts_stride, const double *P,
                              int img_h, int img_w, double max_depth, double *out_dev, int64_t *n_kept_dev,
                              void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || !P || pts_stride < 3 || img_h <= 0 || img_w <= 0 || !n_kept_dev || (n > 0 && (!pts_dev || !out_dev)))
        return pcacc_fail(h, PCACC_ERR_ARG, "bad velo2img arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    if (n == 0) {
        PCACC_CUDA(h, cudaMemsetAsync(n_kept_dev, 0, sizeof(int64_t), st));
        return PCACC_OK;
    }
    PMat pm;
    memcpy(pm.m, P, sizeof(pm.m));
    const int64_t tiles = (n + IBLOCK - 1) / IBLOCK;
    LookBack lb;
    int rc = make_lookback(h, tiles, &lb);
    if (rc) return rc;
    h->launches[PCACC_K_INTEGRATE]++;
    k_velo2img<<<(unsigned)tiles, IBLOCK, 0, st>>>(pts_dev, n, pts_stride, pm, img_h, img_w, max_depth, out_dev,
                                                   n_kept_dev, lb);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_integrate_frustum(pcacc_t h, const float *pts_dev, int64_t n, const double *P,
                                       const uint8_t *rgb_dev, const void *sem_dev, int sem_dtype,
                                       int K, int img_h, int img_w, double max_depth,
                                       const int32_t *filters, int n_filters, int64_t *frame_id,
                                       void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (!P || img_h <= 0 || img_w <= 0 || K < 1)
        return pcacc_fail(h, PCACC_ERR_ARG, "bad integrate_frustum arguments");
    // validated before a frame is registered: an error return must not leave a phantom frame
    if (!valid_sem_dtype(sem_dtype, true)) return pcacc_fail(h, PCACC_ERR_ARG, "unknown sem dtype %d", sem_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    Filters filt;
    int rc = fill_filters(h, filters, n_filters, &filt);
    if (rc) return rc;
    rc = set_inten_div(h, 1.0);
    if (rc) return rc;
    int64_t tiles = (n + IBLOCK * KI_ITEMS - 1) / (IBLOCK * KI_ITEMS);
    if (tiles == 0) tiles = 1;
    LookBack lb;
    rc = make_lookback(h, tiles, &lb);
    if (rc) return rc;
    FrameSlots fs;
    rc = begin_frame(h, n, st, &fs, frame_id);
    if (rc) return rc;
    PMat pm;
    memcpy(pm.m, P, sizeof(pm.m));
    const float4 *p4 = (const float4 *)pts_dev;
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
#define LAUNCH_IF(DT)                                                                          \
    k_integrate_frustum<DT><<<(unsigned)tiles, IBLOCK, 0, st>>>(p4, n, pm, rgb_dev, sem_dev, K, img_h, \
                                                                 img_w, max_depth, filt, h->ring, fs, \
                                                                 lb, h->d_flags)
    switch (sem_dtype) {
        case PCACC_SEM_U8: LAUNCH_IF(PCACC_SEM_U8); break;
        case PCACC_SEM_I32: LAUNCH_IF(PCACC_SEM_I32); break;
        case PCACC_SEM_I64: LAUNCH_IF(PCACC_SEM_I64); break;
        case PCACC_SEM_I16: LAUNCH_IF(PCACC_SEM_I16); break;
        case PCACC_SEM_F32_PROB: LAUNCH_IF(PCACC_SEM_F32_PROB); break;
        default: return pcacc_fail(h, PCACC_ERR_ARG, "unknown sem dtype %d", sem_dtype);
    }
#undef LAUNCH_IF
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_integrate_gt(pcacc_t h, const float *pts_dev, int64_t n,
                                  const int16_t *sem_gt_dev, const int32_t *filters, int n_filters,
                                  int64_t *frame_id, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    Filters filt;
    int rc = fill_filters(h, filters, n_filters, &filt);
    if (rc) return rc;
    rc = set_inten_div(h, 1.0);
    if (rc) return rc;
    int64_t tiles = (n + IBLOCK * KI_ITEMS - 1) / (IBLOCK * KI_ITEMS);
    if (tiles == 0) tiles = 1;
    LookBack lb;
    rc = make_lookback(h, tiles, &lb);
    if (rc) return rc;
    FrameSlots fs;
    rc = begin_frame(h, n, st, &fs, frame_id);
    if (rc) return rc;
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
    k_integrate_gt<<<(unsigned)tiles, IBLOCK, 0, st>>>((const float4 *)pts_dev, n, sem_gt_dev, filt,
                                                       h->ring, fs, lb, h->d_flags);
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

extern "C" int pcacc_integrate_records(pcacc_t h, const double *pc_dev, const int64_t *cam_idx_dev,
                                       int64_t n, const uint8_t *const *rgb_maps,
                                       const void *const *sem_maps, int n_cams, int sem_dtype,
                                       int img_h, int img_w, const double *T_ego_world,
                                       double intensity_div, const int32_t *filters, int n_filters,
                                       int64_t *frame_id, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n_cams < 0 || n_cams > PCACC_MAX_CAMS || !T_ego_world)
        return pcacc_fail(h, PCACC_ERR_ARG, "n_cams must be 0..%d", PCACC_MAX_CAMS);
    if (!valid_sem_dtype(sem_dtype, false)) return pcacc_fail(h, PCACC_ERR_ARG, "unsupported sem dtype %d", sem_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    Filters filt;
    int rc = fill_filters(h, filters, n_filters, &filt);
    if (rc) return rc;
    rc = set_inten_div(h, intensity_div);  // applied where intensity is read (export, rasterise)
    if (rc) return rc;
    CamMaps maps;
    maps.n = n_cams;
    for (int k = 0; k < PCACC_MAX_CAMS; k++) {
        maps.rgb[k] = k < n_cams ? rgb_maps[k] : nullptr;
        maps.sem[k] = k < n_cams ? sem_maps[k] : nullptr;
    }
    TMat T;
    memcpy(T.m, T_ego_world, sizeof(T.m));
    int64_t tiles = (n + RBLOCK * REC_ITEMS_SWEEP - 1) / (RBLOCK * REC_ITEMS_SWEEP);
    if (tiles == 0) tiles = 1;
    LookBack lb;
    rc = make_lookback(h, tiles, &lb);
    if (rc) return rc;
    FrameSlots fs;
    rc = begin_frame(h, n, st, &fs, frame_id);
    if (rc) return rc;
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
#define LAUNCH_IR(DT)                                                                           \
    k_integrate_records<DT><<<(unsigned)tiles, RBLOCK, 0, st>>>(                                \
        pc_dev, (const long long *)cam_idx_dev, n, maps, img_h, img_w, T, filt, h->ring, fs, lb, \
        h->d_flags)
    switch (sem_dtype) {
        case PCACC_SEM_U8: LAUNCH_IR(PCACC_SEM_U8); break;
        case PCACC_SEM_I32: LAUNCH_IR(PCACC_SEM_I32); break;
        case PCACC_SEM_I64: LAUNCH_IR(PCACC_SEM_I64); break;
        case PCACC_SEM_I16: LAUNCH_IR(PCACC_SEM_I16); break;
        default: return pcacc_fail(h, PCACC_ERR_ARG, "unsupported sem dtype %d", sem_dtype);
    }
#undef LAUNCH_IR
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}


extern "C" int pcacc_integrate_records_batch(pcacc_t h, int n_sweeps, const double *const *pc_dev,
                                             const int64_t *const *cam_idx_dev, const int64_t *n,
                                             const uint8_t *const *rgb_maps,
                                             const void *const *sem_maps, int n_cams, int sem_dtype,
                                             int img_h, int img_w, const double *T_ego_world,
                                             double intensity_div, const int32_t *filters,
                                             int n_filters, int64_t *first_frame_id, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n_sweeps <= 0 || n_sweeps > 65535 || n_cams < 0 || n_cams > PCACC_MAX_CAMS || !pc_dev ||
        !cam_idx_dev || !n || !T_ego_world)
        return pcacc_fail(h, PCACC_ERR_ARG, "bad integrate_records_batch arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    Filters filt;
    int rc = fill_filters(h, filters, n_filters, &filt);
    if (rc) return rc;
    rc = set_inten_div(h, intensity_div);
    if (rc) return rc;
    if (!valid_sem_dtype(sem_dtype, false)) return pcacc_fail(h, PCACC_ERR_ARG, "unsupported sem dtype %d", sem_dtype);
    if ((int)(h->next_id - h->first_id) + n_sweeps >= h->max_frames - 1)
        return pcacc_fail(h, PCACC_ERR_CAPACITY, "frame table too small for %d more frames", n_sweeps);
    // The capacity / wrap decision needs exact numbers at most ONCE for the whole batch, and
    // before any of its frames is registered: pcacc_sync refreshes every live frame from the
    // device table, which knows nothing yet about frames whose kernel is not enqueued.
    if (h->next_id > h->first_id) {
        const FrameHost &prev = h->frames[(int)((h->next_id - 1) % h->max_frames)];
        int64_t need = prev.off_ub + align_pts(prev.n_in);
        for (int k = 0; k < n_sweeps; k++) need += align_pts(n[k] < 0 ? 0 : n[k]);
        if (h->wrapped || need > h->capacity) {
            rc = pcacc_sync(h, nullptr, (void *)st);
            if (rc) return rc;
        }
    }
    // any failure below un-registers the batch's frames (their table entries were never written)
    const int64_t id_entry = h->next_id;
    const bool wrapped_entry = h->wrapped;
#define BATCH_FAIL(rc_)                \
    do {                               \
        h->next_id = id_entry;         \
        h->wrapped = wrapped_entry;    \
        return (rc_);                  \
    } while (0)
    std::vector<SweepDesc> desc((size_t)n_sweeps);
    int64_t total_tiles = 0, max_tiles = 1;
    for (int k = 0; k < n_sweeps; k++) {
        int64_t tiles = (n[k] + REC_TILE - 1) / REC_TILE;
        if (tiles <= 0) tiles = 1;
        SweepDesc &d = desc[(size_t)k];
        d.pc = pc_dev[k];
        d.cam = (const long long *)cam_idx_dev[k];
        d.n = n[k];
        d.maps.n = n_cams;
        for (int c = 0; c < PCACC_MAX_CAMS; c++) {
            d.maps.rgb[c] = c < n_cams ? rgb_maps[(size_t)k * n_cams + c] : nullptr;
            d.maps.sem[c] = c < n_cams ? sem_maps[(size_t)k * n_cams + c] : nullptr;
        }
        memcpy(d.T.m, T_ego_world + (size_t)k * 16, sizeof(d.T.m));
        d.n_tiles = (uint32_t)tiles;
        d.state_off = (uint32_t)total_tiles;
        total_tiles += tiles;
        if (tiles > max_tiles) max_tiles = tiles;
        int64_t fid = -1;
        rc = begin_frame(h, n[k], st, &d.fs, &fid, /*fixed_slot=*/true, /*allow_sync=*/false);
        if (rc) BATCH_FAIL(rc);
        if (d.fs.base_override < 0)
            BATCH_FAIL(pcacc_fail(h, PCACC_ERR_STATE, "internal: batch frame without a fixed base"));
        d.fs.write_next = (k == n_sweeps - 1) ? 1 : 0;
        if (k == 0 && first_frame_id) *first_frame_id = fid;
    }
    rc = pcacc_ensure_tiles(h, total_tiles);
    if (rc) BATCH_FAIL(rc);
    // per-sweep tickets live behind the descriptors in the arena; zeroed by the upload
    size_t desc_bytes = desc.size() * sizeof(SweepDesc);
    std::vector<char> blob(desc_bytes + (size_t)n_sweeps * 4, 0);
    memcpy(blob.data(), desc.data(), desc_bytes);
    void *dev = nullptr;
    rc = pcacc_arena_put(h, blob.data(), blob.size(), &dev, st);
    if (rc) BATCH_FAIL(rc);
#undef BATCH_FAIL
    const SweepDesc *d_desc = (const SweepDesc *)dev;
    uint32_t *d_tickets = (uint32_t *)((char *)dev + desc_bytes);
    uint32_t epoch = pcacc_next_epoch(h);
    if (n_sweeps > 1) {
        k_aabb_empty<<<((n_sweeps - 1) * 6 + 127) / 128, 128, 0, st>>>(h->d_aabb, desc[0].fs.slot_id,
                                                                     n_sweeps, h->max_frames);
        PCACC_CUDA(h, cudaGetLastError());
    }
    dim3 grid((unsigned)max_tiles, (unsigned)n_sweeps);
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
#define LAUNCH_IRB(DT)                                                                           \
    k_integrate_records_batch<DT><<<grid, RBLOCK, 0, st>>>(d_desc, img_h, img_w, filt, h->ring,  \
                                                           h->d_tile_state, d_tickets, epoch,    \
                                                           h->d_flags)
    switch (sem_dtype) {
        case PCACC_SEM_U8: LAUNCH_IRB(PCACC_SEM_U8); break;
        case PCACC_SEM_I32: LAUNCH_IRB(PCACC_SEM_I32); break;
        case PCACC_SEM_I64: LAUNCH_IRB(PCACC_SEM_I64); break;
        case PCACC_SEM_I16: LAUNCH_IRB(PCACC_SEM_I16); break;
        default: return pcacc_fail(h, PCACC_ERR_ARG, "unsupported sem dtype %d", sem_dtype);
    }
#undef LAUNCH_IRB
    PCACC_CUDA(h, cudaGetLastError());
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    return PCACC_OK;
}

// ---------------------------------------------------------------------------
// host-buffer entry point of the nuScenes record path (the e2e boundary)
// ---------------------------------------------------------------------------
// Device-visible address of a host pointer, or nullptr when the memory is pageable.
static const void *pinned_dev_ptr(const void *p) {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged) return at.devicePointer;
    if (at.type == cudaMemoryTypeDevice) return at.devicePointer;   // already resident: fine too
    return nullptr;
}

extern "C" int pcacc_host_is_pinned(const void *p) { return pinned_dev_ptr(p) != nullptr; }

static inline int32_t host_class(const void *map, int dt, int64_t pix) {
    switch (dt) {
        case PCACC_SEM_U8: return (int32_t)((const uint8_t *)map)[pix];
        case PCACC_SEM_I16: return (int32_t)((const int16_t *)map)[pix];
        case PCACC_SEM_I32: return ((const int32_t *)map)[pix];
        default: {
            const long long v = ((const long long *)map)[pix];
            return (v < -2147483647ll || v > 2147483647ll) ? -2147483647 : (int32_t)v;
        }
    }
}

static int stage_slot(pcacc_t h, size_t bytes, pcacc_s::StageSlot **out) {
    pcacc_s::StageSlot &sl = h->stage[h->stage_turn];
    h->stage_turn = (h->stage_turn + 1) % PCACC_STAGE_SLOTS;
    if (sl.busy) {   // the kernel that read this slot PCACC_STAGE_SLOTS integrates ago
        PCACC_CUDA(h, cudaEventSynchronize(sl.ev));
        sl.busy = false;
    }
    if (!sl.ev) PCACC_CUDA(h, cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    if (sl.cap < bytes) {
        if (sl.host) cudaFreeHost(sl.host);
        sl.host = nullptr;
        sl.cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc((void **)&sl.host, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return pcacc_fail(h, PCACC_ERR_NOMEM, "pinned staging slot of %zu bytes: %s", want,
                              cudaGetErrorString(e));
        }
        sl.cap = want;
    }
    *out = &sl;
    return PCACC_OK;
}

extern "C" int pcacc_integrate_records_host(pcacc_t h, const double *pc_host,
                                            const int64_t *cam_idx_host, int64_t n,
                                            const uint8_t *const *rgb_maps_host,
                                            const void *const *sem_maps_host, int n_cams,
                                            int sem_dtype, int img_h, int img_w,
                                            const double *T_ego_world, double intensity_div,
                                            const int32_t *filters, int n_filters, int staging,
                                            int *staging_used, int64_t *frame_id, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n < 0 || n_cams < 0 || n_cams > PCACC_MAX_CAMS || !T_ego_world || (n > 0 && (!pc_host || !cam_idx_host)) ||
        (n_cams > 0 && (!rgb_maps_host || !sem_maps_host)) || img_h <= 0 || img_w <= 0)
        return pcacc_fail(h, PCACC_ERR_ARG, "bad integrate_records_host arguments");
    if (!valid_sem_dtype(sem_dtype, false)) return pcacc_fail(h, PCACC_ERR_ARG, "unsupported sem dtype %d", sem_dtype);
    if (staging != PCACC_STAGE_AUTO && staging != PCACC_STAGE_DIRECT && staging != PCACC_STAGE_SPARSE)
        return pcacc_fail(h, PCACC_ERR_ARG, "unknown staging mode %d", staging);
    PCACC_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;

    // DIRECT: every array is page-locked (or resident): the kernel reads it in place
    const void *d_pc = nullptr, *d_cam = nullptr;
    const uint8_t *d_rgb[PCACC_MAX_CAMS];
    const void *d_sem[PCACC_MAX_CAMS];
    bool direct = staging != PCACC_STAGE_SPARSE;
    if (direct) {
        d_pc = n ? pinned_dev_ptr(pc_host) : (const void *)pc_host;
        d_cam = n ? pinned_dev_ptr(cam_idx_host) : (const void *)cam_idx_host;
        direct = (n == 0) || (d_pc && d_cam);
        for (int c = 0; c < n_cams && direct; c++) {
            d_rgb[c] = (const uint8_t *)pinned_dev_ptr(rgb_maps_host[c]);
            d_sem[c] = pinned_dev_ptr(sem_maps_host[c]);
            direct = d_rgb[c] && d_sem[c];
        }
    }
    if (!direct && staging == PCACC_STAGE_DIRECT)
        return pcacc_fail(h, PCACC_ERR_ARG, "PCACC_STAGE_DIRECT needs page-locked (pinned) host arrays");
    if (staging_used) *staging_used = direct ? PCACC_STAGE_DIRECT : PCACC_STAGE_SPARSE;
    if (direct)
        return pcacc_integrate_records(h, (const double *)d_pc, (const int64_t *)d_cam, n, d_rgb, d_sem, n_cams,
                                       sem_dtype, img_h, img_w, T_ego_world, intensity_div, filters,
                                       n_filters, frame_id, stream);

    const int sem_bytes = sem_dtype == PCACC_SEM_U8 ? 1 : sem_dtype == PCACC_SEM_I16 ? 2 : sem_dtype == PCACC_SEM_I32 ? 4 : 8;
    // SPARSE: the rows a camera sees and their (rgb, class) samples, in input order, go to a
    // pinned slot the kernel reads in place: 64 B per visible point instead of the whole
    // (n,7) array and the camera frames.
    Filters filt;
    int rc = fill_filters(h, filters, n_filters, &filt);
    if (rc) return rc;
    rc = set_inten_div(h, intensity_div);
    if (rc) return rc;
    // pass 0: indices of the points a camera sees — a branch-free compaction (one point in ten is
    // visible, in no predictable pattern: a branch per point costs more than the store it avoids)
    std::vector<uint32_t> &vidx = h->stage_idx;
    std::vector<int64_t> &vpix = h->stage_pix;
    vidx.resize((size_t)n + 1);
    size_t nv = 0;
    {
        uint32_t *vi = vidx.data();
        const uint64_t nc = (uint64_t)n_cams;
        for (int64_t i = 0; i < n; i++) {
            vi[nv] = (uint32_t)i;
            nv += (uint64_t)cam_idx_host[i] < nc;     // 0 <= c < n_cams in one unsigned compare
        }
    }
    vidx.resize(nv);
    vpix.resize(nv);
    // pass 1: pixel addresses of the visible points and a prefetch of the two cache lines each of
    // them will read (the gathers are the cache-missing part of the staging; the iterations are
    // independent, so the row reads of this pass overlap as well)
    const double wlim = (double)img_w - 1.0, hlim = (double)img_h - 1.0;
    for (size_t k = 0; k < nv; k++) {
        if (k + 24 < nv) {   // the rows are ~560 bytes apart: no hardware prefetcher follows them
            const double *nxt = pc_host + 7 * (int64_t)vidx[k + 24];
            __builtin_prefetch(nxt);
            __builtin_prefetch(nxt + 6);
        }
        const int64_t i = vidx[k], c = cam_idx_host[i];
        const double uf = pc_host[7 * i + 4], vf = pc_host[7 * i + 5];
        int64_t pix = -1;
        // outside the image the kernel raises PCACC_FLAG_UV_OUT_OF_IMAGE and never looks at the sample
        if ((uf > 1.0) && (uf < wlim) && (vf > 1.0) && (vf < hlim)) {
            pix = (int64_t)nearbyint(vf) * img_w + (int64_t)nearbyint(uf);   // half to even
            __builtin_prefetch(rgb_maps_host[c] + pix * 3);
            __builtin_prefetch((const char *)sem_maps_host[c] + pix * sem_bytes);
        }
        vpix[k] = pix;
    }
    const int64_t n_vis = (int64_t)vidx.size();
    pcacc_s::StageSlot *sl = nullptr;
    const size_t row_bytes = (size_t)n_vis * 7 * sizeof(double);
    rc = stage_slot(h, row_bytes + (size_t)n_vis * 8 + 64, &sl);
    if (rc) return rc;
    double *rows = (double *)sl->host;
    uint32_t *samp = (uint32_t *)(sl->host + row_bytes);
    // pass 2: rows and samples, in input order
    for (int64_t j = 0; j < n_vis; j++) {
        const int64_t i = vidx[(size_t)j];
        memcpy(rows + 7 * j, pc_host + 7 * i, 7 * sizeof(double));
        const int64_t c = cam_idx_host[i], pix = vpix[(size_t)j];
        uint32_t rgbp = 0;
        int32_t cls = 0;
        if (pix >= 0) {
            const uint8_t *px = rgb_maps_host[c] + pix * 3;
            rgbp = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
            cls = host_class(sem_maps_host[c], sem_dtype, pix);
        }
        samp[2 * j] = rgbp;
        samp[2 * j + 1] = (uint32_t)cls;
    }
    CamMaps maps;
    maps.n = 1;
    for (int k = 0; k < PCACC_MAX_CAMS; k++) {
        maps.rgb[k] = nullptr;
        maps.sem[k] = k == 0 ? (const void *)samp : nullptr;
    }
    TMat T;
    memcpy(T.m, T_ego_world, sizeof(T.m));
    int64_t tiles = (n_vis + RBLOCK * REC_ITEMS_SWEEP - 1) / (RBLOCK * REC_ITEMS_SWEEP);
    if (tiles == 0) tiles = 1;
    LookBack lb;
    rc = make_lookback(h, tiles, &lb);
    if (rc) return rc;
    FrameSlots fs;
    rc = begin_frame(h, n_vis, st, &fs, frame_id);
    if (rc) return rc;
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
    k_integrate_records<PCACC_SEM_SAMPLED><<<(unsigned)tiles, RBLOCK, 0, st>>>(
        rows, nullptr, n_vis, maps, img_h, img_w, T, filt, h->ring, fs, lb, h->d_flags);
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    PCACC_CUDA(h, cudaGetLastError());
    PCACC_CUDA(h, cudaEventRecord(sl->ev, st));
    sl->busy = true;
    return PCACC_OK;
}

extern "C" int pcacc_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    if (!bytes) return PCACC_OK;
    if (!dst_host || !src_dev) return PCACC_ERR_ARG;
    return cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream) == cudaSuccess
               ? PCACC_OK
               : PCACC_ERR_CUDA;
}

extern "C" int pcacc_integrate_cloud(pcacc_t h, const double *rec_dev, int64_t n, int64_t *frame_id,
                                     void *stream) {
    if (!h) return PCACC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    int rc = set_inten_div(h, 1.0);
    if (rc) return rc;
    FrameSlots fs;
    rc = begin_frame(h, n, st, &fs, frame_id);
    if (rc) return rc;
    int64_t blocks = (n + IBLOCK - 1) / IBLOCK;
    if (blocks == 0) blocks = 1;
    size_t pe = pcacc_prof_begin(h, PCACC_K_INTEGRATE, st);
    k_integrate_cloud<<<(unsigned)blocks, IBLOCK, 0, st>>>(rec_dev, n, h->ring, fs, h->d_flags);
    pcacc_prof_end(h, PCACC_K_INTEGRATE, pe, st);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}

// apply every pending transform of every live frame in place (exact chain)
static int materialise_all(pcacc_t h, cudaStream_t st) {
    int n_live = (int)(h->next_id - h->first_id);
    if (n_live == 0 || !h->any_lazy) return PCACC_OK;
    int first_slot = (int)(h->first_id % h->max_frames);
    int64_t max_n = 0;
    for (int64_t k = h->first_id; k < h->next_id; k++) {
        FrameHost &f = h->frames[(int)(k % h->max_frames)];
        int64_t c = f.exact ? f.cnt : f.n_in;
        if (c > max_n) max_n = c;
    }
    int64_t bx = (max_n + IBLOCK - 1) / IBLOCK;
    if (bx < 1) bx = 1;
    if (bx > 4096) bx = 4096;
    dim3 grid((unsigned)bx, (unsigned)n_live);
    size_t pe = pcacc_prof_begin(h, PCACC_K_REBASE, st);
    k_materialise<<<grid, IBLOCK, 0, st>>>(h->ring, h->d_frame_off, h->d_frame_cnt, h->d_frame_epoch,
                                           h->d_chain, first_slot, h->max_frames, h->rebase_epoch);
    PCACC_CUDA(h, cudaGetLastError());
    h->launches[PCACC_K_REBASE]++;
    k_materialise_done<<<(n_live + 127) / 128, 128, 0, st>>>(h->d_frame_epoch, h->d_comp, first_slot,
                                                             n_live, h->max_frames, h->rebase_epoch);
    PCACC_CUDA(h, cudaGetLastError());
    pcacc_prof_end(h, PCACC_K_REBASE, pe, st);
    for (int64_t k = h->first_id; k < h->next_id; k++)
        h->frames[(int)(k % h->max_frames)].epoch = h->rebase_epoch;
    h->any_lazy = false;
    return PCACC_OK;
}

extern "C" int pcacc_rebase(pcacc_t h, const double *T_new_prev, int eager, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (!T_new_prev) return pcacc_fail(h, PCACC_ERR_ARG, "T_new_prev is null");
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    int n_live = (int)(h->next_id - h->first_id);
    if (n_live == 0) return PCACC_OK;  // nothing stored: the reference loop is empty too
    // the chain ring holds max_frames transforms: epochs [oldest live, now]
    int64_t oldest = h->rebase_epoch;
    for (int64_t k = h->first_id; k < h->next_id; k++) {
        int64_t e = h->frames[(int)(k % h->max_frames)].epoch;
        if (e < oldest) oldest = e;
    }
    if (h->rebase_epoch + 1 - oldest > h->max_frames) {
        int rc = materialise_all(h, st);  // fold the pending chain before it is overwritten
        if (rc) return rc;
    }
    TMat T;
    memcpy(T.m, T_new_prev, sizeof(T.m));
    int first_slot = (int)(h->first_id % h->max_frames);
    double *slot = h->d_chain + (h->rebase_epoch % h->max_frames) * 12;
    int threads = 2 * n_live < 12 ? 12 : 2 * n_live;
    size_t pe = pcacc_prof_begin(h, PCACC_K_REBASE, st);
    k_rebase_lazy<<<(threads + 127) / 128, 128, 0, st>>>(slot, h->d_comp, h->d_cull, T, first_slot,
                                                         n_live, h->max_frames);
    PCACC_CUDA(h, cudaGetLastError());
    pcacc_prof_end(h, PCACC_K_REBASE, pe, st);
    h->rebase_epoch += 1;
    h->any_lazy = true;
    if (eager) return materialise_all(h, st);
    return PCACC_OK;
}

extern "C" int pcacc_mark_dynamic(pcacc_t h, const int64_t *frame_ids, const int32_t *inst_idx,
                                  int n_pairs, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    if (n_pairs <= 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    // group the pairs by frame slot (stable: order inside a frame is irrelevant)
    std::vector<std::pair<int32_t, int32_t>> pr((size_t)n_pairs);
    int64_t max_n = 0;
    for (int k = 0; k < n_pairs; k++) {
        FrameHost *f = pcacc_frame(h, frame_ids[k]);
        if (!f) return pcacc_fail(h, PCACC_ERR_ARG, "frame %lld is not live", (long long)frame_ids[k]);
        pr[k] = {(int32_t)(frame_ids[k] % h->max_frames), inst_idx[k]};
        int64_t c = f->exact ? f->cnt : f->n_in;
        if (c > max_n) max_n = c;
    }
    std::sort(pr.begin(), pr.end());
    std::vector<int32_t> buf;
    buf.reserve(4 * (size_t)n_pairs);
    std::vector<int32_t> inst((size_t)n_pairs);
    int n_grp = 0;
    for (int k = 0; k < n_pairs; k++) {
        inst[k] = pr[k].second;
        if (k == 0 || pr[k].first != pr[k - 1].first) {
            if (n_grp) buf[3 * (n_grp - 1) + 2] = k;
            buf.push_back(pr[k].first);
            buf.push_back(k);
            buf.push_back(n_pairs);
            n_grp++;
        }
    }
    const size_t grp_words = buf.size();
    buf.insert(buf.end(), inst.begin(), inst.end());
    void *dev = nullptr;
    int rc = pcacc_arena_put(h, buf.data(), buf.size() * sizeof(int32_t), &dev, st);
    if (rc) return rc;
    int64_t bx = (max_n + IBLOCK - 1) / IBLOCK;
    if (bx < 1) bx = 1;
    // grid-stride beyond a few blocks per frame: unsynchronised frames only have upper bounds
    // of their counts (often 10x the kept points) and a block without work still costs a launch slot
    if (bx > 16) bx = 16;
    for (int g0 = 0; g0 < n_grp; g0 += 32768) {
        int ny = n_grp - g0 < 32768 ? n_grp - g0 : 32768;
        dim3 grid((unsigned)bx, (unsigned)ny);
        size_t pe = pcacc_prof_begin(h, PCACC_K_MARK, st);
        k_mark_dynamic<<<grid, IBLOCK, 0, st>>>(h->ring, h->d_frame_off, h->d_frame_cnt,
                                                (const int32_t *)dev + 3 * g0,
                                                (const int32_t *)dev + grp_words);
        PCACC_CUDA(h, cudaGetLastError());
        pcacc_prof_end(h, PCACC_K_MARK, pe, st);
    }
    return PCACC_OK;
}

extern "C" int pcacc_export_frame(pcacc_t h, int64_t frame_id, double *out_dev, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    FrameHost *f = pcacc_frame(h, frame_id);
    if (!f) return pcacc_fail(h, PCACC_ERR_ARG, "frame %lld is not live", (long long)frame_id);
    if (!f->exact) return pcacc_fail(h, PCACC_ERR_STATE, "pcacc_sync() needed before export");
    if (f->cnt == 0) return PCACC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    int64_t blocks = (f->cnt + IBLOCK - 1) / IBLOCK;
    h->launches[PCACC_K_EXPORT]++;
    k_export_frame<<<(unsigned)blocks, IBLOCK, 0, st>>>(
        h->ring, h->d_frame_off, h->d_frame_cnt, h->d_frame_epoch, h->d_chain,
        (int)(frame_id % h->max_frames), h->max_frames, h->rebase_epoch,
        h->inten_div > 0.0 ? h->inten_div : 1.0, out_dev);
    PCACC_CUDA(h, cudaGetLastError());
    return PCACC_OK;
}
