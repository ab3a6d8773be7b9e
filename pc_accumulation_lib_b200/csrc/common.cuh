// common.cuh — shared device helpers and the handle layout of libpcacc.
//
// Exactness rules (DESIGN.md "Arithmetic"): every coordinate expression whose
// result feeds an index is written with explicit round-to-nearest intrinsics
// (__dmul_rn/__dadd_rn/__ddiv_rn never contract; __fma_rn where the reference's
// dgemm fuses), in the op order of SURVEY.md Appendix B.  The library is also
// compiled with -fmad=false so no stray a*b+c is contracted.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <vector>

#include "../../include/pcacc.h"

#define PCACC_MAX_FILTERS 16
#define PCACC_MAX_CAMS 8
#define PCACC_ALIGN_PTS 4  // frame offsets are multiples of this many records
// internal map dtype of the sparse host staging mode: maps.sem[0] points to (n,2) uint32
// {r | g<<8 | b<<16, class as int32} already gathered on the host; cam_idx is absent
#define PCACC_SEM_SAMPLED 5
#define PCACC_STAGE_SLOTS 4
#define PCACC_ARENA_SEGS 4
#define PCACC_ARENA_SEG_BYTES (1u << 20)

// ---------------------------------------------------------------------------
// exact small transforms
// ---------------------------------------------------------------------------
// rows 0..2 of M(3x4, row-major, leading dim ld) @ [x y z 1]: mul then three
// fused multiply-adds over k, the arithmetic numpy's float64 matmul performs
// for these shapes (sem_pc_accum.py:362, :179; datasets/nuscenes_utils.py:58).
__device__ __forceinline__ void affine_chain(const double *__restrict__ M, int ld, double x, double y,
                                             double z, double &ox, double &oy, double &oz) {
    const double *m0 = M, *m1 = M + ld, *m2 = M + 2 * ld;
    ox = __fma_rn(m0[3], 1.0, __fma_rn(m0[2], z, __fma_rn(m0[1], y, __dmul_rn(m0[0], x))));
    oy = __fma_rn(m1[3], 1.0, __fma_rn(m1[2], z, __fma_rn(m1[1], y, __dmul_rn(m1[0], x))));
    oz = __fma_rn(m2[3], 1.0, __fma_rn(m2[2], z, __fma_rn(m2[1], y, __dmul_rn(m2[0], x))));
}

// half-to-even rounding of np.round on float64
__device__ __forceinline__ double rint_even(double v) { return rint(v); }

__device__ __forceinline__ int32_t sat_i32(double v) {
    if (!(v == v)) return INT32_MIN;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

// ---------------------------------------------------------------------------
// single-pass chained scan ("decoupled look-back") over tile aggregates.
// state word: [63:34] launch epoch | [33:32] flag | [31:0] value
// A word whose epoch differs from the launch's is "not written yet", so the
// array never needs clearing between launches.
// ---------------------------------------------------------------------------
#define LB_FLAG_AGG 1ull
#define LB_FLAG_PREFIX 2ull

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, unsigned long long flag,
                                                      uint32_t value) {
    return ((unsigned long long)epoch << 34) | (flag << 32) | (unsigned long long)value;
}

__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_store(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Takes a ticket (tile index in launch order). Call with all threads; uses one
// shared word.  The thread that draws the last ticket re-arms the counter.
__device__ __forceinline__ uint32_t lb_take_ticket(uint32_t *ticket, uint32_t n_tiles,
                                                   uint32_t *s_tile) {
    if (threadIdx.x == 0) {
        uint32_t t = atomicAdd(ticket, 1u);
        if (t == n_tiles - 1) atomicExch(ticket, 0u);
        *s_tile = t;
    }
    __syncthreads();
    return *s_tile;
}

// Called by warp 0 only (all 32 lanes). Publishes this tile's aggregate and
// returns the exclusive prefix of all earlier tiles.
__device__ __forceinline__ uint32_t lb_exclusive_prefix(unsigned long long *state, uint32_t epoch,
                                                        uint32_t tile, uint32_t aggregate) {
    const unsigned lane = threadIdx.x & 31u;
    if (tile == 0) {
        if (lane == 0) lb_store(&state[0], lb_pack(epoch, LB_FLAG_PREFIX, aggregate));
        return 0;
    }
    if (lane == 0) lb_store(&state[tile], lb_pack(epoch, LB_FLAG_AGG, aggregate));
    uint32_t exclusive = 0;
    int look = (int)tile - 1;
    while (true) {
        int idx = look - (int)lane;
        unsigned long long s;
        unsigned long long flag;
        do {
            if (idx >= 0) {
                s = lb_load(&state[idx]);
                flag = ((uint32_t)(s >> 34) == epoch) ? ((s >> 32) & 3ull) : 0ull;
            } else {
                s = 0;
                flag = LB_FLAG_PREFIX;  // virtual tile before the first: prefix 0
            }
        } while (__any_sync(0xffffffffu, flag == 0ull));
        unsigned pm = __ballot_sync(0xffffffffu, flag == LB_FLAG_PREFIX);
        int first = pm ? (__ffs(pm) - 1) : 31;  // nearest tile that already has an inclusive prefix
        uint32_t v = ((int)lane <= first) ? (uint32_t)s : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        exclusive += v;
        if (pm) break;
        look -= 32;
    }
    if (lane == 0) lb_store(&state[tile], lb_pack(epoch, LB_FLAG_PREFIX, exclusive + aggregate));
    return exclusive;
}

// Block-wide order-preserving compaction rank for ONE flag per thread.
// Returns the global rank (exclusive count of kept items before this thread in
// launch order) and the tile's inclusive end in *tile_end (valid in all threads).
template <int BLOCK>
__device__ __forceinline__ uint32_t compact_rank(bool keep, unsigned long long *state, uint32_t epoch,
                                                 uint32_t tile, uint32_t *s_warp /*BLOCK/32+2*/,
                                                 uint32_t *tile_end) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned m = __ballot_sync(0xffffffffu, keep);
    uint32_t in_warp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = BLOCK / 32;
        uint32_t c = (lane < NW) ? s_warp[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = lb_exclusive_prefix(state, epoch, tile, total);
        if (lane < NW) s_warp[lane] = base + incl - c;
        if (lane == 0) s_warp[NW] = base + total;
    }
    __syncthreads();
    *tile_end = s_warp[BLOCK / 32];
    return s_warp[warp] + in_warp;
}

// The same for ITEMS flags per thread, where item k of thread t is element k*BLOCK + t of
// the tile (ITEMS * BLOCK/32 <= 32).
template <int BLOCK, int ITEMS>
__device__ __forceinline__ void compact_rank_multi(const bool (&keep)[ITEMS], unsigned long long *state,
                                                   uint32_t epoch, uint32_t tile,
                                                   uint32_t *s_cnt /*ITEMS*BLOCK/32+1*/,
                                                   uint32_t (&rank)[ITEMS], uint32_t *tile_end) {
    constexpr int NW = BLOCK / 32;
    static_assert(ITEMS * NW <= 32, "one warp scans all sub-row counts");
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t in_warp[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, keep[k]);
        in_warp[k] = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[k * NW + warp] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t c = (lane < ITEMS * NW) ? s_cnt[lane] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t base = lb_exclusive_prefix(state, epoch, tile, total);
        if (lane < ITEMS * NW) s_cnt[lane] = base + incl - c;
        if (lane == 0) s_cnt[ITEMS * NW] = base + total;
    }
    __syncthreads();
    *tile_end = s_cnt[ITEMS * NW];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) rank[k] = s_cnt[k * NW + warp] + in_warp[k];
}

// ---------------------------------------------------------------------------
// per-frame bounding boxes (frame culling in the rasteriser)
// ---------------------------------------------------------------------------
// order-preserving map double -> uint64, so atomicMin / atomicMax work on doubles
__device__ __forceinline__ unsigned long long ord_encode(double d) {
    long long b = __double_as_longlong(d);
    return b >= 0 ? ((unsigned long long)b ^ 0x8000000000000000ull) : ~(unsigned long long)b;
}
__device__ __forceinline__ double ord_decode(unsigned long long u) {
    long long b = (u & 0x8000000000000000ull) ? (long long)(u ^ 0x8000000000000000ull) : (long long)~u;
    return __longlong_as_double(b);
}
#define AABB_EMPTY_MIN 0xffffffffffffffffull
#define AABB_EMPTY_MAX 0ull

// Block-wide min / max of the kept points' stored coordinates -> 6 atomics per block.
// Must be called by every thread of the block. NaN coordinates are ignored (fmin / fmax).
// v = this thread's (min x, min y, min z, max x, max y, max z); +-inf when it kept nothing.
template <int BLOCK>
__device__ __forceinline__ void aabb_update_v(unsigned long long *aabb, double (&v)[6]);

template <int BLOCK>
__device__ __forceinline__ void aabb_update(unsigned long long *aabb, bool keep, double x, double y,
                                            double z) {
    double v[6];
    v[0] = keep ? x : INFINITY;
    v[1] = keep ? y : INFINITY;
    v[2] = keep ? z : INFINITY;
    v[3] = keep ? x : -INFINITY;
    v[4] = keep ? y : -INFINITY;
    v[5] = keep ? z : -INFINITY;
    aabb_update_v<BLOCK>(aabb, v);
}

template <int BLOCK>
__device__ __forceinline__ void aabb_update_v(unsigned long long *aabb, double (&v)[6]) {
    __shared__ double s_bb[BLOCK / 32][6];
    // warp min / max on the order-encoded halves with the REDUX unit (two reductions per
    // value) instead of five 64-bit shuffle + fmin steps: this reduction was a third of the
    // integrate kernels' instructions
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double x = v[k];
        if (x != x) x = k < 3 ? INFINITY : -INFINITY;   // NaN: the identity, as fmin / fmax had it
        const unsigned long long u = ord_encode(x);
        const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
        unsigned mh, ml;
        if (k < 3) {
            mh = __reduce_min_sync(0xffffffffu, hi);
            ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
        } else {
            mh = __reduce_max_sync(0xffffffffu, hi);
            ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        }
        v[k] = ord_decode(((unsigned long long)mh << 32) | ml);
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) s_bb[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        double r = s_bb[0][k];
        for (int w = 1; w < BLOCK / 32; w++) r = k < 3 ? fmin(r, s_bb[w][k]) : fmax(r, s_bb[w][k]);
        if (k < 3) {
            if (r != INFINITY) atomicMin(&aabb[k], ord_encode(r));
        } else {
            if (r != -INFINITY) atomicMax(&aabb[k], ord_encode(r));
        }
    }
}

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct FrameHost {
    int64_t n_in;     // upper bound of the kept count
    int64_t off_ub;   // upper bound of the ring offset (exact when `exact`)
    int64_t off;      // exact offset  (valid when exact)
    int64_t cnt;      // exact count   (valid when exact)
    int64_t epoch;    // number of re-base transforms recorded before this frame
    bool exact;
};

struct RingDev {
    double *x, *y, *z;
    float *inten;
    uint32_t *rgbs;  // r | g<<8 | b<<16 | sem<<24
    int32_t *inst;
    uint8_t *dyn;
};

struct pcacc_s {
    int device;
    int64_t capacity;
    int max_frames;
    RingDev ring;
    // device frame table, slot = frame id % max_frames
    int64_t *d_frame_off;
    int64_t *d_frame_cnt;
    int64_t *d_frame_epoch;
    double *d_comp;    // max_frames x 12: composed lazy matrix of each frame
    double *d_cull;    // max_frames x 12: every transform since insertion (never reset): frame culling
    unsigned long long *d_aabb;  // max_frames x 6: order-encoded min xyz / max xyz at insertion
    double *d_chain;   // max_frames x 12: re-base transforms, slot = epoch % max_frames
    uint32_t *d_flags;
    unsigned long long *d_tile_state;
    int64_t tile_cap;
    uint32_t *d_ticket;
    uint32_t launch_epoch;
    // host mirror
    int64_t first_id, next_id;
    int64_t rebase_epoch;        // number of transforms recorded so far
    int64_t chain_floor;         // smallest epoch still stored in d_chain
    std::vector<FrameHost> frames;  // indexed by slot
    bool wrapped;
    bool any_lazy;
    // pinned mailbox for table sync
    int64_t *h_mail;
    // parameter arena (host pinned + device): a ring of PCACC_ARENA_SEGS segments; a segment
    // is reused only after the event recorded when the allocator left it has completed
    char *h_arena, *d_arena;
    size_t arena_size, arena_pos;   // arena_pos: offset inside the current segment
    int arena_seg;
    cudaEvent_t arena_ev[PCACC_ARENA_SEGS];
    bool arena_ev_set[PCACC_ARENA_SEGS];
    int n_sm;            // multiprocessors of the device (grid sizing)
    // pinned staging ring of the host-buffer entry points (sparse staging mode): a slot is
    // rewritten only after the event recorded behind the kernel that read it has completed
    struct StageSlot { char *host; size_t cap; cudaEvent_t ev; bool busy; };
    StageSlot stage[PCACC_STAGE_SLOTS];
    int stage_turn;
    std::vector<uint32_t> stage_idx;   // scratch of the sparse staging pass
    std::vector<int64_t> stage_pix;
    // raster workspace (grow-only)
    void *d_ws;
    size_t ws_size;
    int64_t *d_rstats;  // [1] binned [2] replays of the last rasterise
    double *d_rgb_lut;   // (m2 * 0.5) / 255. for m2 = 0..510
    int64_t last_visit_ub;
    double inten_div;    // stored intensity / inten_div = reference intensity; 0 = not set yet
    uint32_t pending_flags;
    int cls_mult, bin_mult;   // grid size of k_bev_classify / k_bev_bin in multiples of the SM count
    bool classify_single;     // A/B switch: 1 = k_bev_classify for every variant count
    bool reduce_strips;   // debugging / A-B switch: 1 = the 32-cell strip kernel for float16 output too
    // launch accounting + optional per-kernel CUDA-event timing
    int64_t launches[PCACC_N_KERNELS];
    uint32_t prof_mask;   // bit k: launches of kernel class k are bracketed by events
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used;
    struct ProfSpan { int kernel; size_t ev0, ev1; };
    std::vector<ProfSpan> prof_spans;
    double prof_ms[PCACC_N_KERNELS];
    int64_t prof_n[PCACC_N_KERNELS];
    char err[512];
};

// RAII-less span helpers: count every launch; bracket it with events when profiling
size_t pcacc_prof_begin(pcacc_t h, int kernel, cudaStream_t st);
void pcacc_prof_end(pcacc_t h, int kernel, size_t ev0, cudaStream_t st);
int pcacc_prof_flush(pcacc_t h);
int pcacc_init_tables(pcacc_t h);

int pcacc_fail(pcacc_t h, int status, const char *fmt, ...);
int pcacc_cuda_check(pcacc_t h, cudaError_t e, const char *what);
uint32_t pcacc_next_epoch(pcacc_t h);
int pcacc_ensure_tiles(pcacc_t h, int64_t n_tiles);
// copies `bytes` of host parameters into the arena; returns the device address
int pcacc_arena_put(pcacc_t h, const void *src, size_t bytes, void **dev, cudaStream_t st);
FrameHost *pcacc_frame(pcacc_t h, int64_t frame_id);

#define PCACC_CUDA(h, call)                                   \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return pcacc_cuda_check(h, e__, #call); \
    } while (0)
