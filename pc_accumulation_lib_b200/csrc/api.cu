// api.cu — handle lifetime, frame-table bookkeeping and error plumbing of libpcacc.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";

int pcacc_fail(pcacc_t h, int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (h) memcpy(h->err, g_err, sizeof(h->err));
    return status;
}

int pcacc_cuda_check(pcacc_t h, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return PCACC_OK;
    return pcacc_fail(h, PCACC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

uint32_t pcacc_next_epoch(pcacc_t h) {
    h->launch_epoch = (h->launch_epoch + 1u) & 0x3fffffffu;
    if (h->launch_epoch == 0) h->launch_epoch = 1;
    return h->launch_epoch;
}

int pcacc_ensure_tiles(pcacc_t h, int64_t n_tiles) {
    if (n_tiles <= h->tile_cap) return PCACC_OK;
    int64_t want = n_tiles * 2;
    if (want < 4096) want = 4096;
    if (h->d_tile_state) {
        PCACC_CUDA(h, cudaDeviceSynchronize());
        cudaFree(h->d_tile_state);
        h->d_tile_state = nullptr;
        h->tile_cap = 0;
    }
    PCACC_CUDA(h, cudaMalloc(&h->d_tile_state, (size_t)want * 8));
    PCACC_CUDA(h, cudaMemset(h->d_tile_state, 0, (size_t)want * 8));
    h->tile_cap = want;
    return PCACC_OK;
}

int pcacc_arena_put(pcacc_t h, const void *src, size_t bytes, void **dev, cudaStream_t st) {
    const size_t seg = PCACC_ARENA_SEG_BYTES;
    size_t need = (bytes + 255) / 256 * 256;
    if (need > seg) return pcacc_fail(h, PCACC_ERR_ARG, "parameter block of %zu bytes too large", bytes);
    if (h->arena_pos + need > seg) {
        // leave the current segment: everything enqueued so far (on this stream) consumes its
        // parameters before this event; the segment is handed out again PCACC_ARENA_SEGS - 1
        // segments later, after waiting for that event only — no device-wide synchronisation
        PCACC_CUDA(h, cudaEventRecord(h->arena_ev[h->arena_seg], st));
        h->arena_ev_set[h->arena_seg] = true;
        h->arena_seg = (h->arena_seg + 1) % PCACC_ARENA_SEGS;
        if (h->arena_ev_set[h->arena_seg]) {
            PCACC_CUDA(h, cudaEventSynchronize(h->arena_ev[h->arena_seg]));
            h->arena_ev_set[h->arena_seg] = false;
        }
        h->arena_pos = 0;
    }
    const size_t at = (size_t)h->arena_seg * seg + h->arena_pos;
    memcpy(h->h_arena + at, src, bytes);
    PCACC_CUDA(h, cudaMemcpyAsync(h->d_arena + at, h->h_arena + at, bytes, cudaMemcpyHostToDevice, st));
    *dev = h->d_arena + at;
    h->arena_pos += need;
    return PCACC_OK;
}

size_t pcacc_prof_begin(pcacc_t h, int kernel, cudaStream_t st) {
    h->launches[kernel]++;
    if (!((h->prof_mask >> kernel) & 1u)) return 0;
    if (h->prof_used + 2 > h->prof_events.size()) {
        size_t old = h->prof_events.size();
        h->prof_events.resize(old + 1024);
        for (size_t k = old; k < h->prof_events.size(); k++) cudaEventCreate(&h->prof_events[k]);
    }
    size_t ev0 = h->prof_used;
    h->prof_used += 2;
    cudaEventRecord(h->prof_events[ev0], st);
    return ev0;
}

void pcacc_prof_end(pcacc_t h, int kernel, size_t ev0, cudaStream_t st) {
    if (!((h->prof_mask >> kernel) & 1u)) return;
    cudaEventRecord(h->prof_events[ev0 + 1], st);
    h->prof_spans.push_back({kernel, ev0, ev0 + 1});
    if (h->prof_spans.size() >= 16384) pcacc_prof_flush(h);
}

int pcacc_prof_flush(pcacc_t h) {
    for (auto &sp : h->prof_spans) {
        cudaEventSynchronize(h->prof_events[sp.ev1]);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->prof_events[sp.ev0], h->prof_events[sp.ev1]) == cudaSuccess) {
            h->prof_ms[sp.kernel] += ms;
            h->prof_n[sp.kernel]++;
        }
    }
    h->prof_spans.clear();
    h->prof_used = 0;
    return PCACC_OK;
}

extern "C" int pcacc_profile(pcacc_t h, int class_mask) {
    if (!h) return PCACC_ERR_ARG;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    if (h->prof_mask) pcacc_prof_flush(h);
    h->prof_mask = (uint32_t)class_mask & ((1u << PCACC_N_KERNELS) - 1u);
    return PCACC_OK;
}

extern "C" int pcacc_profile_read(pcacc_t h, double ms[PCACC_N_KERNELS], int64_t launches[PCACC_N_KERNELS]) {
    if (!h) return PCACC_ERR_ARG;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    PCACC_CUDA(h, cudaDeviceSynchronize());
    pcacc_prof_flush(h);
    for (int k = 0; k < PCACC_N_KERNELS; k++) {
        if (ms) ms[k] = h->prof_ms[k];
        if (launches) launches[k] = h->launches[k];
        h->prof_ms[k] = 0.0;
        h->prof_n[k] = 0;
        h->launches[k] = 0;
    }
    return PCACC_OK;
}

FrameHost *pcacc_frame(pcacc_t h, int64_t frame_id) {
    if (frame_id < h->first_id || frame_id >= h->next_id) return nullptr;
    return &h->frames[(int)(frame_id % h->max_frames)];
}

// ---------------------------------------------------------------------------
// bare events for the host-side staging ring (an Event.record() through PyTorch costs a
// current_stream() lookup of ~25 us per call; the staging path records one per integrate)
// ---------------------------------------------------------------------------
extern "C" int pcacc_event_create(void **event) {
    if (!event) return PCACC_ERR_ARG;
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        return PCACC_ERR_CUDA;
    }
    *event = (void *)e;
    return PCACC_OK;
}
extern "C" int pcacc_event_record(void *event, void *stream) {
    if (!event) return PCACC_ERR_ARG;
    return cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream) == cudaSuccess ? PCACC_OK : PCACC_ERR_CUDA;
}
extern "C" int pcacc_event_sync(void *event) {
    if (!event) return PCACC_OK;
    return cudaEventSynchronize((cudaEvent_t)event) == cudaSuccess ? PCACC_OK : PCACC_ERR_CUDA;
}
extern "C" int pcacc_event_destroy(void *event) {
    if (event) cudaEventDestroy((cudaEvent_t)event);
    return PCACC_OK;
}

extern "C" const char *pcacc_strerror(int status) {
    switch (status) {
        case PCACC_OK: return "ok";
        case PCACC_ERR_ARG: return "invalid argument";
        case PCACC_ERR_CUDA: return "CUDA error";
        case PCACC_ERR_CAPACITY: return "ring or frame table capacity exceeded";
        case PCACC_ERR_NOMEM: return "device memory allocation failed";
        case PCACC_ERR_STATE: return "call not valid in the current state";
        default: return "unknown status";
    }
}

extern "C" const char *pcacc_last_error(pcacc_t h) { return h ? h->err : g_err; }

extern "C" int pcacc_abi_version(void) { return PCACC_ABI_VERSION; }

static void free_all(pcacc_t h) {
    cudaFree(h->ring.x);
    cudaFree(h->ring.y);
    cudaFree(h->ring.z);
    cudaFree(h->ring.inten);
    cudaFree(h->ring.rgbs);
    cudaFree(h->ring.inst);
    cudaFree(h->ring.dyn);
    cudaFree(h->d_frame_off);
    cudaFree(h->d_frame_cnt);
    cudaFree(h->d_frame_epoch);
    cudaFree(h->d_comp);
    cudaFree(h->d_cull);
    cudaFree(h->d_aabb);
    cudaFree(h->d_chain);
    cudaFree(h->d_flags);
    cudaFree(h->d_tile_state);
    cudaFree(h->d_ticket);
    cudaFree(h->d_arena);
    cudaFree(h->d_ws);
    cudaFree(h->d_rstats);
    cudaFree(h->d_rgb_lut);
    for (auto &e : h->prof_events) cudaEventDestroy(e);
    for (int k = 0; k < PCACC_ARENA_SEGS; k++)
        if (h->arena_ev[k]) cudaEventDestroy(h->arena_ev[k]);
    for (int k = 0; k < PCACC_STAGE_SLOTS; k++) {
        if (h->stage[k].ev) cudaEventDestroy(h->stage[k].ev);
        if (h->stage[k].host) cudaFreeHost(h->stage[k].host);
    }
    if (h->h_mail) cudaFreeHost(h->h_mail);
    if (h->h_arena) cudaFreeHost(h->h_arena);
}

extern "C" int pcacc_create(int device, int64_t capacity_pts, int max_frames, pcacc_t *out) {
    if (!out || capacity_pts <= 0 || max_frames < 4 || capacity_pts > 0xffffff00ll)
        return pcacc_fail(nullptr, PCACC_ERR_ARG, "pcacc_create: capacity_pts > 0 and max_frames >= 4 required");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return pcacc_fail(nullptr, PCACC_ERR_CUDA, "no CUDA device (%s): libpcacc has no CPU fallback",
                          e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n_dev)
        return pcacc_fail(nullptr, PCACC_ERR_ARG, "device %d out of range (%d devices)", device, n_dev);
    pcacc_t h = new (std::nothrow) pcacc_s();
    if (!h) return PCACC_ERR_NOMEM;
    h->device = device;
    // records are addressed in groups of PCACC_ALIGN_PTS
    h->capacity = (capacity_pts + PCACC_ALIGN_PTS - 1) / PCACC_ALIGN_PTS * PCACC_ALIGN_PTS;
    h->max_frames = max_frames;
    h->frames.resize((size_t)max_frames);
    h->launch_epoch = 0;
    h->arena_size = (size_t)PCACC_ARENA_SEGS * PCACC_ARENA_SEG_BYTES;
    size_t cap = (size_t)h->capacity;
#define TRY(call)                                                           \
    do {                                                                    \
        cudaError_t e2 = (call);                                            \
        if (e2 != cudaSuccess) {                                            \
            int rc = pcacc_fail(nullptr, e2 == cudaErrorMemoryAllocation ? PCACC_ERR_NOMEM : PCACC_ERR_CUDA, \
                                "%s: %s", #call, cudaGetErrorString(e2));   \
            cudaGetLastError();                                             \
            free_all(h);                                                    \
            delete h;                                                       \
            return rc;                                                      \
        }                                                                   \
    } while (0)
    TRY(cudaSetDevice(device));
    TRY(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device));
    for (int k = 0; k < PCACC_ARENA_SEGS; k++)
        TRY(cudaEventCreateWithFlags(&h->arena_ev[k], cudaEventDisableTiming));
    TRY(cudaMalloc(&h->ring.x, cap * 8));
    TRY(cudaMalloc(&h->ring.y, cap * 8));
    TRY(cudaMalloc(&h->ring.z, cap * 8));
    TRY(cudaMalloc(&h->ring.inten, cap * 4));
    TRY(cudaMalloc(&h->ring.rgbs, cap * 4));
    TRY(cudaMalloc(&h->ring.inst, cap * 4));
    TRY(cudaMalloc(&h->ring.dyn, cap));
    TRY(cudaMalloc(&h->d_frame_off, (size_t)max_frames * 8));
    TRY(cudaMalloc(&h->d_frame_cnt, (size_t)max_frames * 8));
    TRY(cudaMalloc(&h->d_frame_epoch, (size_t)max_frames * 8));
    TRY(cudaMalloc(&h->d_comp, (size_t)max_frames * 12 * 8));
    TRY(cudaMalloc(&h->d_chain, (size_t)max_frames * 12 * 8));
    TRY(cudaMalloc(&h->d_cull, (size_t)max_frames * 12 * 8));
    TRY(cudaMalloc(&h->d_aabb, (size_t)max_frames * 6 * 8));
    TRY(cudaMemset(h->d_cull, 0, (size_t)max_frames * 12 * 8));
    {   // every box starts empty: min = +max code, max = 0
        std::vector<unsigned long long> init((size_t)max_frames * 6);
        for (size_t k = 0; k < init.size(); k++) init[k] = (k % 6) < 3 ? ~0ull : 0ull;
        TRY(cudaMemcpy(h->d_aabb, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    }
    TRY(cudaMalloc(&h->d_flags, 4));
    TRY(cudaMalloc(&h->d_ticket, 4));
    TRY(cudaMalloc(&h->d_rstats, 4 * 8));
    TRY(cudaMalloc(&h->d_arena, h->arena_size));
    TRY(cudaMemset(h->d_frame_off, 0, (size_t)max_frames * 8));
    TRY(cudaMemset(h->d_frame_cnt, 0, (size_t)max_frames * 8));
    TRY(cudaMemset(h->d_frame_epoch, 0, (size_t)max_frames * 8));
    TRY(cudaMemset(h->d_comp, 0, (size_t)max_frames * 12 * 8));
    TRY(cudaMemset(h->d_chain, 0, (size_t)max_frames * 12 * 8));
    TRY(cudaMemset(h->d_flags, 0, 4));
    TRY(cudaMemset(h->d_ticket, 0, 4));
    TRY(cudaMemset(h->d_rstats, 0, 4 * 8));
    TRY(cudaMallocHost(&h->h_mail, (size_t)(3 * max_frames + 8) * 8));
    TRY(cudaMallocHost(&h->h_arena, h->arena_size));
#undef TRY
    {   // A/B switch for profiling runs (see pcacc_set_option)
        const char *e = getenv("PCACC_REDUCE_STRIPS");
        h->reduce_strips = e && e[0] == '1';
        const char *c = getenv("PCACC_CLS_MULT"), *b = getenv("PCACC_BIN_MULT");
        h->cls_mult = c && atoi(c) > 0 ? atoi(c) : 4;
        h->bin_mult = b && atoi(b) > 0 ? atoi(b) : 0;   // 0 = by launch shape (raster.cu): 6 blocks per SM for batches of variants, 9 otherwise
        const char *cs = getenv("PCACC_CLASSIFY_SINGLE");
        h->classify_single = cs && cs[0] == '1';
    }
    int rc = pcacc_ensure_tiles(h, 4096);
    if (!rc) rc = pcacc_init_tables(h);
    if (rc) {
        free_all(h);
        delete h;
        return rc;
    }
    *out = h;
    return PCACC_OK;
}

extern "C" int pcacc_destroy(pcacc_t h) {
    if (!h) return PCACC_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_all(h);
    delete h;
    return PCACC_OK;
}

extern "C" int pcacc_reset(pcacc_t h, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    (void)stream;
    h->first_id = h->next_id;  // ids keep counting; nothing is live
    h->wrapped = false;
    h->any_lazy = false;
    h->inten_div = 0.0;
    return PCACC_OK;
}

extern "C" int pcacc_set_option(pcacc_t h, int option, int value) {
    if (!h) return PCACC_ERR_ARG;
    switch (option) {
        case PCACC_OPT_REDUCE_STRIPS: h->reduce_strips = value != 0; return PCACC_OK;
        case PCACC_OPT_CLASSIFY_SINGLE: h->classify_single = value != 0; return PCACC_OK;
        default: return pcacc_fail(h, PCACC_ERR_ARG, "unknown option %d", option);
    }
}

extern "C" int pcacc_evict(pcacc_t h, int n_frames) {
    if (!h) return PCACC_ERR_ARG;
    if (n_frames < 0 || n_frames > (int)(h->next_id - h->first_id))
        return pcacc_fail(h, PCACC_ERR_ARG, "cannot evict %d of %d live frames", n_frames,
                          (int)(h->next_id - h->first_id));
    h->first_id += n_frames;
    return PCACC_OK;
}

extern "C" int pcacc_sync(pcacc_t h, uint32_t *flags, void *stream) {
    if (!h) return PCACC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    PCACC_CUDA(h, cudaSetDevice(h->device));
    const int mf = h->max_frames;
    PCACC_CUDA(h, cudaMemcpyAsync(h->h_mail, h->d_frame_off, (size_t)mf * 8, cudaMemcpyDeviceToHost, st));
    PCACC_CUDA(h, cudaMemcpyAsync(h->h_mail + mf, h->d_frame_cnt, (size_t)mf * 8, cudaMemcpyDeviceToHost, st));
    PCACC_CUDA(h, cudaMemcpyAsync(h->h_mail + 2 * mf, h->d_flags, 4, cudaMemcpyDeviceToHost, st));
    PCACC_CUDA(h, cudaMemsetAsync(h->d_flags, 0, 4, st));
    PCACC_CUDA(h, cudaStreamSynchronize(st));
    for (int64_t id = h->first_id; id < h->next_id; id++) {
        int slot = (int)(id % mf);
        FrameHost &f = h->frames[slot];
        f.off = h->h_mail[slot];
        f.cnt = h->h_mail[mf + slot];
        f.off_ub = f.off;
        f.n_in = f.cnt;
        f.exact = true;
    }
    uint32_t fl = *(uint32_t *)(h->h_mail + 2 * mf);
    h->pending_flags |= fl;
    if (flags) {
        *flags = h->pending_flags;
        h->pending_flags = 0;
    }
    return PCACC_OK;
}

extern "C" int pcacc_num_frames(pcacc_t h, int64_t *first_frame_id, int *n_live) {
    if (!h) return PCACC_ERR_ARG;
    if (first_frame_id) *first_frame_id = h->first_id;
    if (n_live) *n_live = (int)(h->next_id - h->first_id);
    return PCACC_OK;
}

extern "C" int pcacc_frame_count(pcacc_t h, int64_t frame_id, int64_t *count) {
    if (!h || !count) return PCACC_ERR_ARG;
    FrameHost *f = pcacc_frame(h, frame_id);
    if (!f) return pcacc_fail(h, PCACC_ERR_ARG, "frame %lld is not live", (long long)frame_id);
    if (!f->exact) return pcacc_fail(h, PCACC_ERR_STATE, "pcacc_sync() needed after integrate");
    *count = f->cnt;
    return PCACC_OK;
}

extern "C" int pcacc_frame_offset(pcacc_t h, int64_t frame_id, int64_t *offset) {
    if (!h || !offset) return PCACC_ERR_ARG;
    FrameHost *f = pcacc_frame(h, frame_id);
    if (!f) return pcacc_fail(h, PCACC_ERR_ARG, "frame %lld is not live", (long long)frame_id);
    if (!f->exact) return pcacc_fail(h, PCACC_ERR_STATE, "pcacc_sync() needed after integrate");
    *offset = f->off;
    return PCACC_OK;
}

extern "C" int64_t pcacc_resident_points(pcacc_t h) {
    if (!h) return 0;
    int64_t s = 0;
    for (int64_t id = h->first_id; id < h->next_id; id++) {
        FrameHost &f = h->frames[(int)(id % h->max_frames)];
        s += f.exact ? f.cnt : f.n_in;
    }
    return s;
}
