/* pcacc.h — C ABI of libpcacc.so: the sm_100a implementation of the per-frame
 * semantic fusion + BEV rasterisation path of robin-karlsson0/pc-accumulation-lib.
 *
 * The reference has no FFI of its own (it is pure Python/numpy);
 * its boundary for this path is the Python class API.  Each entry point below
 * names the reference function(s) it replaces (paths relative to the reference
 * root).  The Python mirror of that class API (the .py files of pc_accumulation_lib_b200/)
 * binds these symbols with ctypes; INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every bulk pointer named *_dev is a DEVICE pointer; small parameter
 *     blocks (matrices, filter lists, pcacc_bev_params) are HOST pointers and
 *     are consumed before the call returns;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     nothing synchronises unless the comment says so;
 *   - return value: PCACC_OK (0) or a negative pcacc_status; pcacc_strerror()
 *     gives text, pcacc_last_error() the detail of the last failure;
 *   - there is no CPU fallback: without a CUDA device every call fails with
 *     PCACC_ERR_CUDA.
 */
#ifndef PCACC_H_
#define PCACC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCACC_ABI_VERSION 1

typedef struct pcacc_s *pcacc_t;

typedef enum {
    PCACC_OK = 0,
    PCACC_ERR_ARG = -1,       /* bad argument */
    PCACC_ERR_CUDA = -2,      /* CUDA runtime error (see pcacc_last_error) */
    PCACC_ERR_CAPACITY = -3,  /* ring or frame table full */
    PCACC_ERR_NOMEM = -4,     /* device allocation failed */
    PCACC_ERR_STATE = -5      /* call not valid in the current state */
} pcacc_status;

/* device-side data error flags, OR-ed into the word read by pcacc_sync() */
#define PCACC_FLAG_UV_OUT_OF_IMAGE 1u  /* pts_feat_from_img assertion, datasets/nuscenes_utils.py:194-195 */
#define PCACC_FLAG_INTENSITY_F32   2u  /* informational: an intensity was rounded to float32 (6e-8 relative; the ring stores f32) */
#define PCACC_FLAG_ATTR_RANGE      4u  /* r/g/b/sem outside 0..255 or inst outside int32 */
#define PCACC_FLAG_CELL_OVERFLOW   8u  /* internal: scatter cursor ran past its segment */

/* dtype codes for semantic maps */
#define PCACC_SEM_U8 0
#define PCACC_SEM_I32 1
#define PCACC_SEM_I64 2
#define PCACC_SEM_I16 4
#define PCACC_SEM_F32_PROB 3 /* (H,W,K) float32 class probabilities: class = first argmax over K */
/* image dtypes accepted by pcacc_pts_feat_from_img in addition to the integer codes above */
#define PCACC_IMG_F32 6
#define PCACC_IMG_F64 7

const char *pcacc_strerror(int status);
const char *pcacc_last_error(pcacc_t h);
int pcacc_abi_version(void);

/* ---- lifetime ----------------------------------------------------------
 * The accumulated cloud (the reference's `self.sem_pcs` list of (M,10) float64
 * arrays, sem_pc_accum.py:100) lives in a device-resident SoA ring of
 * `capacity_pts` records: x,y,z float64 | intensity float32 | r,g,b,sem uint8x4
 * | inst int32 | dyn uint8 (37 B/point), plus a table of `max_frames` frame
 * segments.  `device` is the CUDA ordinal. */
int pcacc_create(int device, int64_t capacity_pts, int max_frames, pcacc_t *out);
int pcacc_destroy(pcacc_t h);
/* drop all frames (new scene) */
int pcacc_reset(pcacc_t h, void *stream);

/* ---- stand-alone operators (helpers the reference exposes) --------------
 * velo2frame + velo2img, sem_pc_accum.py:347-402: for ALL n points writes
 * u, v (int32; value of np.round(...).astype(int) when it fits) and the
 * in-image mask (1/0).  pts_dev: (n, pts_stride) float32, columns 0..2 = xyz.
 * P: host (3,4) float64 row-major. */
int pcacc_project(const float *pts_dev, int64_t n, int pts_stride, const double *P,
                  int img_h, int img_w, double max_depth,
                  int32_t *u_dev, int32_t *v_dev, uint8_t *mask_dev, void *stream);

/* velo2img as the reference returns it, sem_pc_accum.py:367-402: out_dev (n_kept, pts_stride + 2)
 * float64 = the rows [pc_velo, u, v] of the points inside the image (and 0 < depth < max_depth), in
 * input order; *n_kept_dev = their number (out_dev must hold n rows). */
int pcacc_velo2img(pcacc_t h, const float *pts_dev, int64_t n, int pts_stride, const double *P,
                   int img_h, int img_w, double max_depth, double *out_dev, int64_t *n_kept_dev,
                   void *stream);

/* gen_semantic_pc, sem_pc_accum.py:323-345: projection, order-preserving
 * compaction of in-image points and the K-channel gather map[v,u,:].
 * out_dev: (n, 4+K) float64 (only the first *n_kept rows are written),
 * n_kept_dev: device int64.  map_dtype: PCACC_SEM_U8 / I32 / I64 or
 * PCACC_SEM_F32_PROB (plain float32 copy here, no argmax). */
int pcacc_gen_semantic_pc(pcacc_t h, const float *pts_dev, int64_t n, const double *P,
                          const void *map_dev, int map_dtype, int img_h, int img_w, int K,
                          double *out_dev, int64_t *n_kept_dev, void *stream);

/* ---- integrate: append one frame to the ring ----------------------------
 * All three are asynchronous; the frame's kept count lands in the device
 * frame table and is fetched by pcacc_sync().  Each returns the new frame's
 * absolute id in *frame_id (ids count up from 0 after create/reset).
 *
 * KITTI-360 frustum path — Kitti360SemanticPointCloudAccumulator
 * .obs2sem_vec_space with sem_gt=None, kitti360_sem_pc_accum.py:129-156 over
 * sem_pc_accum.py:317-402: projection, in-image mask, RGB + class gather,
 * class filter, inst=0, dyn=0, order-preserving append.
 *   pts_dev (n,4) float32 [x,y,z,intensity]; rgb_dev (H,W,3) uint8;
 *   sem_dev (H,W) class indices of sem_dtype, or (H,W,K) float32 probabilities
 *   (PCACC_SEM_F32_PROB, class = argmax); filters: host list of class ids. */
int pcacc_integrate_frustum(pcacc_t h, const float *pts_dev, int64_t n, const double *P,
                            const uint8_t *rgb_dev, const void *sem_dev, int sem_dtype, int K,
                            int img_h, int img_w, double max_depth,
                            const int32_t *filters, int n_filters,
                            int64_t *frame_id, void *stream);

/* KITTI-360 `use_gt_sem` path, kitti360_sem_pc_accum.py:138-156: no
 * projection, rgb = 0, class from sem_gt (n,) int16, class filter. */
int pcacc_integrate_gt(pcacc_t h, const float *pts_dev, int64_t n, const int16_t *sem_gt_dev,
                       const int32_t *filters, int n_filters,
                       int64_t *frame_id, void *stream);

/* nuScenes oracle-pose path — NuScenesOracleSemanticPointCloudAccumulator
 * .obs2sem_vec_space, nuscenes_oracle_sem_pc_accum.py:454-501 with
 * pts_feat_from_img(method='nearest') (datasets/nuscenes_utils.py:181-214)
 * and homo_transform (datasets/nuscenes_utils.py:46-60).
 *   pc_dev (n,7) float64 [x,y,z,intensity,u,v,inst] (the dataloader's layout,
 *   obs_dataloaders/nuscenes_obs_dataloader.py:103-123); cam_idx_dev (n,) int64
 *   (-1 = seen by no camera); rgb_maps/sem_maps: host arrays of n_cams device
 *   pointers ((H,W,3) uint8 and (H,W) sem_dtype); T_ego_world: host (4,4)
 *   float64; intensity is stored raw (float32) and divided by intensity_div
 *   (255.) in float64 wherever it is used; one ring holds one divisor
 *   (PCACC_ERR_STATE otherwise; the frustum / gt / cloud paths use 1). */
int pcacc_integrate_records(pcacc_t h, const double *pc_dev, const int64_t *cam_idx_dev, int64_t n,
                            const uint8_t *const *rgb_maps, const void *const *sem_maps,
                            int n_cams, int sem_dtype, int img_h, int img_w,
                            const double *T_ego_world, double intensity_div,
                            const int32_t *filters, int n_filters,
                            int64_t *frame_id, void *stream);

/* The same for n_sweeps observations in ONE launch (offline dataset generation:
 * all sweeps of a scene are on the device already).  Arrays of n_sweeps host
 * entries; rgb_maps / sem_maps hold n_sweeps * n_cams device pointers
 * (sweep-major); T_ego_world holds n_sweeps (4,4) matrices.  Every sweep becomes
 * its own frame (ids first_frame_id .. +n_sweeps-1), placed at fixed upper-bound
 * offsets so no sweep waits for another sweep's kept count. */
int pcacc_integrate_records_batch(pcacc_t h, int n_sweeps, const double *const *pc_dev,
                                  const int64_t *const *cam_idx_dev, const int64_t *n,
                                  const uint8_t *const *rgb_maps, const void *const *sem_maps,
                                  int n_cams, int sem_dtype, int img_h, int img_w,
                                  const double *T_ego_world, double intensity_div,
                                  const int32_t *filters, int n_filters,
                                  int64_t *first_frame_id, void *stream);

/* Host-buffer variant of pcacc_integrate_records — the end-to-end boundary: every bulk
 * pointer is a HOST pointer (the arrays the dataloader hands to
 * NuScenesOracleSemanticPointCloudAccumulator.integrate, nuscenes_oracle_sem_pc_accum.py:
 * 139-160).  rgb_maps_host / sem_maps_host: host arrays of n_cams host pointers.
 *   PCACC_STAGE_DIRECT  every array is page-locked (cudaHostAlloc / cudaHostRegister, e.g.
 *                       torch pin_memory): the kernel reads them in place over PCIe — it
 *                       touches cam_idx, the u,v of visible points and the rows of kept
 *                       points, a fraction of the arrays — and no host copy is made.  The
 *                       arrays must stay alive and unchanged until the stream has passed
 *                       this call (pcacc_sync, or any later synchronisation of `stream`).
 *   PCACC_STAGE_SPARSE  pageable arrays: the host walks cam_idx once and packs the rows a
 *                       camera sees plus their (rgb, class) samples — 64 B per visible
 *                       point, in input order — into a pinned ring slot which the same
 *                       kernel reads (map dtype "sampled"); the arrays are consumed before
 *                       the call returns.  Identical records by construction (tested).
 *   PCACC_STAGE_AUTO    DIRECT if every array is page-locked, else SPARSE.
 * *staging_used (optional) reports the mode taken. */
#define PCACC_STAGE_AUTO 0
#define PCACC_STAGE_DIRECT 1
#define PCACC_STAGE_SPARSE 2
int pcacc_integrate_records_host(pcacc_t h, const double *pc_host, const int64_t *cam_idx_host,
                                 int64_t n, const uint8_t *const *rgb_maps_host,
                                 const void *const *sem_maps_host, int n_cams, int sem_dtype,
                                 int img_h, int img_w, const double *T_ego_world,
                                 double intensity_div, const int32_t *filters, int n_filters,
                                 int staging, int *staging_used, int64_t *frame_id, void *stream);
/* 1 if `p` is page-locked host memory (or device / managed memory) a kernel may dereference */
int pcacc_host_is_pinned(const void *p);
/* cudaMemcpyAsync device -> (pinned) host on `stream`: the result planes of pcacc_rasterise */
int pcacc_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes, void *stream);

/* Append an already-built (n,10) float64 cloud [x,y,z,i,r,g,b,sem,inst,dyn]
 * as one frame (used by BEVGenerator.generate(pcs, ...) when it is handed host
 * clouds, bev_generator/bev_generator.py:63-125).  Rows are stored unchanged. */
int pcacc_integrate_cloud(pcacc_t h, const double *rec_dev, int64_t n,
                          int64_t *frame_id, void *stream);

/* ---- ego-motion re-basing: update_sem_pcs, sem_pc_accum.py:167-183 ------
 * T_new_prev: host (4,4) float64 applied to every stored point of every live
 * frame.  eager=1 rewrites xyz in place (exactly what the reference does);
 * eager=0 records T in the per-frame chain and folds it into the frame's
 * composed matrix — points stay in their source frame and the rasteriser
 * applies the composed matrix, replaying the exact chain for any point within
 * a guard band of a cell / crop / height boundary (DESIGN.md "lazy re-base"). */
int pcacc_rebase(pcacc_t h, const double *T_new_prev, int eager, void *stream);

/* horizon eviction, remove_observations sem_pc_accum.py:185-209: drop the
 * n_frames oldest live frames (host bookkeeping only). */
int pcacc_evict(pcacc_t h, int n_frames);

/* fake-tracker dynamic flags, nuscenes_oracle_sem_pc_accum.py:223-230,243-250:
 * for each pair k: dyn = 1 on points of frame frame_ids[k] whose inst equals
 * inst_idx[k].  Host arrays. */
int pcacc_mark_dynamic(pcacc_t h, const int64_t *frame_ids, const int32_t *inst_idx, int n_pairs,
                       void *stream);

/* ---- table / export -------------------------------------------------------
 * Synchronises `stream`, fetches the frame table and the data-error flags.
 * *flags (optional) receives the OR of PCACC_FLAG_* raised since the last
 * sync and clears them. */
int pcacc_sync(pcacc_t h, uint32_t *flags, void *stream);
int pcacc_num_frames(pcacc_t h, int64_t *first_frame_id, int *n_live);
/* kept-point count of a live frame (requires a pcacc_sync after its integrate) */
int pcacc_frame_count(pcacc_t h, int64_t frame_id, int64_t *count);
int64_t pcacc_resident_points(pcacc_t h); /* sum over live frames (after sync) */
/* one frame as the reference's (M,10) float64 record array, `self.sem_pcs[i]`;
 * lazy re-base chains are replayed exactly. out_dev: (count,10) float64. */
int pcacc_export_frame(pcacc_t h, int64_t frame_id, double *out_dev, void *stream);

/* ---- BEV rasterisation ----------------------------------------------------
 * One pcacc_bev_params per BEV variant; a call rasterises n_variants BEVs of
 * the resident cloud in one batch of launches.  Replaces
 *   accumulators' generate_bev window split + origin shift
 *     (kitti360_sem_pc_accum.py:179-213, nuscenes_oracle_sem_pc_accum.py:521-560),
 *   BEVGenerator.preprocess_pc_and_trajs / geometric_transform / crop_view /
 *     pos2grid for the point cloud (bev_generator/bev_generator.py:127-160,
 *     207-255,737-747),
 *   SemBEVGenerator.generate_bev's grids (bev_generator/sem_bev.py:54-118,
 *     196-257): partition_semantic_pc, gen_sem_probmap, gen_intensity_map,
 *     road_marking_transform, get_elevation_map, get_rgb_maps, astype(float16).
 * Trajectories stay on the host (pcacc_preprocess_trajectories). */
typedef struct {
    int64_t frame_begin;   /* absolute ids: present = [frame_begin, frame_split) */
    int64_t frame_split;   /*               future  = [frame_split, frame_end)   */
    int64_t frame_end;     /*               full    = [frame_begin, frame_end)   */
    double origin[3];      /* poses[present_idx], subtracted from xyz */
    double R[9];           /* rotation_matrix_3d(rot_ang), computed on the host with numpy */
    double trans_dx, trans_dy;
    double view;           /* zoom_scalar * view_size */
    double height_filter;  /* NaN = none */
    double int_scaler, int_sep_scaler, int_mid_threshold;
    double rgb_fill;
    int32_t road_cls;      /* sem_idxs['road'] */
    int32_t veh_cls[4];    /* car, truck, bus, motorcycle */
    int32_t elevation_max; /* 0 = reference (per-cell min z); 1 = north-star variant (max z) */
} pcacc_bev_params;

#define PCACC_BEV_PLANES 7 /* road, intensity, r, g, b, dynamic, elevation */
#define PCACC_BEV_WINDOWS 3 /* present, future, full */

/* out_f16_dev: (n_variants, 3 windows, 7 planes, P, P) float16.
 * out_f64_dev: optional, same shape in float64 (values before the cast).
 * dbg_cell_dev: optional (ring capacity,) int32 indexed by ring position: for
 *   variant 0, row*P+col of each point that survives crop/height/static
 *   filtering, -1 otherwise (parity tests). */
int pcacc_rasterise(pcacc_t h, const pcacc_bev_params *params, int n_variants, int P,
                    void *out_f16_dev, double *out_f64_dev, int32_t *dbg_cell_dev,
                    void *stream);

/* Polynomial warp of finished planes — BEVGenerator.warp_dense_probmaps,
 * bev_generator/bev_generator.py:482-525: out[b, p, jw, iw] = in[b, p, jmap[b][jw], imap[b][iw]].
 * The warp is a pure gather, so applying it to the float16 planes equals casting the warped
 * float64 maps (sem_bev.py:121-257).  imap / jmap: HOST arrays (n_bevs x P) of source indices,
 * already clamped to [0, P-1] (the Python mirror computes them with the reference's
 * rint(c1*k + c2*k^2)).  in / out: (n_bevs, n_planes, P, P) float16, distinct buffers. */
int pcacc_warp_planes(pcacc_t h, const void *in_f16_dev, void *out_f16_dev, int n_bevs, int n_planes,
                      int P, const int32_t *imap, const int32_t *jmap, void *stream);

/* Implementation switches for A/B measurements and equivalence tests (results are bit-identical
 * either way).  PCACC_OPT_REDUCE_STRIPS = 1: the per-cell reduction of float16-only output runs
 * the 32-cell strip kernel (k_bev_reduce) instead of the 128-cell chunk kernel
 * (k_bev_reduce_chunk, the default whenever P*P is a multiple of 128). */
#define PCACC_OPT_REDUCE_STRIPS 1
/* PCACC_OPT_CLASSIFY_SINGLE = 1: candidate selection runs k_bev_classify (block per variant) for
 * every batch size instead of k_bev_classify_mv (warp per variant, batches of >= 8 variants). */
#define PCACC_OPT_CLASSIFY_SINGLE 2
/* (Profiling runs can set the same switches, and the grid sizes of k_bev_classify / k_bev_bin in
 * multiples of the SM count, through the environment read by pcacc_create: PCACC_REDUCE_STRIPS=1,
 * PCACC_CLASSIFY_SINGLE=1, PCACC_CLS_MULT=<n>, PCACC_BIN_MULT=<n>.) */
int pcacc_set_option(pcacc_t h, int option, int value);

/* ring position (record index) of a live frame's first point, for dbg_cell_dev */
int pcacc_frame_offset(pcacc_t h, int64_t frame_id, int64_t *offset);

/* counters of the last pcacc_rasterise on this handle (after a sync):
 * [0] points visited, [1] points binned (in view, static), [2] exact-chain replays */
int pcacc_raster_stats(pcacc_t h, int64_t stats[3], void *stream);

/* ---- host helper: trajectory crop ------------------------------------------
 * BEVGenerator.crop_trajectory + cal_intersec_pnt, bev_generator/bev_generator.py:257-371:
 * polyline traj (n,3) against the open box (-view/2, view/2)^2.  A vertex that starts an
 * edge and lies inside is kept; an edge that crosses the boundary adds its midpoint-
 * bisection point (refined until the replaced end moved by <= thresh).  Pure host code
 * (a few dozen points per BEV; the Python loop was the end-to-end bottleneck).
 * out: at least 2*n rows of 3 doubles; *n_out rows are written. */
int pcacc_crop_trajectory(const double *traj, int n, double view, double thresh, double *out,
                          int *n_out);

/* ---- host helper: all trajectories of one generate call ---------------------
 * BEVGenerator.preprocess_pc_and_trajs for the trajectory lists (bev_generator.py:127-160:
 * geometric_transform 207-238 -> crop_trajectory 257-316 -> pos2grid 737-747), for n_var
 * augmentation variants at once.  pts: the polylines back to back, (traj_off[n_traj], 3)
 * doubles; polyline t = rows traj_off[t] .. traj_off[t+1].  variants: n_var x 12 doubles
 * {R[9] row-major, trans_dx, trans_dy, view}.  Per variant and polyline: R @ p as numpy's
 * float64 matmul computes it (FMA chain over k), += (dx, dy), crop against the open view
 * box, x,y -> floor(q / view * P + 0.5 P).  Polylines of fewer than 2 points give 0 rows.
 * out: n_var x 2*traj_off[n_traj] rows of 3 doubles; the rows of (v, t) start at row
 * v*2*traj_off[n_traj] + 2*traj_off[t]; out_cnt[v*n_traj + t] rows are valid. */
int pcacc_preprocess_trajectories(const double *pts, const int32_t *traj_off, int n_traj,
                                  const double *variants, int n_var, int P, double thresh,
                                  double *out, int32_t *out_cnt);

/* ---- bare events for host-side staging rings -------------------------------
 * The binding stages every observation through reusable pinned buffers and must know when
 * the kernels that read a buffer are done before it overwrites it: one event recorded per
 * integrate on the caller's stream (timing disabled), waited for two integrates later. */
int pcacc_event_create(void **event);
int pcacc_event_record(void *event, void *stream);
int pcacc_event_sync(void *event);   /* NULL: nothing to wait for */
int pcacc_event_destroy(void *event);

/* ---- the dataloader's input side ----------------------
 * What produces the (N,7) rows and pc_cam_idx that pcacc_integrate_records consumes.
 *
 * pcacc_assign_boxes — the box loop of inst_centric_get_sweeps, datasets/nuscenes_utils.py:
 * 412-470, with find_points_in_box :317-329 and apply_tf :233-243: for every box in order,
 * box_points = ([xyz 1] @ inv(target_from_box).T)[:, :3] (float64 FMA chain),
 * inside = all(|box_points / dxdydz| < 0.5 + tolerance); a later box overwrites an earlier
 * one.  The 4x4 inverses are computed by the caller (numpy LA.inv, as the reference does) and
 * passed as box_from_target (host, n_boxes x 16, row-major); dxdydz: host n_boxes x 3.
 * pts: device, (n, stride) float32 (pts_f32 != 0) or float64, x y z in the first three columns.
 * out_box: device (n,) int32 = index of the LAST box that contains the point, -1 if none.
 * out_count: device (n_boxes,) int32 = points inside each box (the loop skips boxes with none,
 * which decides the instance numbering on the host). */
int pcacc_assign_boxes(pcacc_t h, const void *pts_dev, int pts_f32, int64_t n, int64_t stride,
                       const double *box_from_target, const double *dxdydz, int n_boxes,
                       double tolerance, int32_t *out_box_dev, int32_t *out_count_dev, void *stream);

/* pcacc_project_cameras — obs_dataloaders/nuscenes_obs_dataloader.py:176-198:
 * pc_in_glob = homo_transform(glob_from_ego, pc_in_ego); for every camera j in order:
 * pc_in_cam = homo_transform(inv(cam.glob_from_self), pc_in_glob); NuScenesCamera.project_pts3d
 * (datasets/nuscenes_utils.py:112-136: valid = z > depth_thres; uv = (viewpad(K) [p 1])[:2] /
 * (...)[2], nuscenes-devkit view_points with normalize=True; inside = 1 < uv < img_wh - 1);
 * pc_uv[inside] = uv, pc_cam_idx[inside] = j (a later camera overwrites an earlier one).
 * pc_ego: device (n, stride) float64.  Host arrays: glob_from_ego 16 doubles; cam_from_glob
 * n_cams x 16 (the inverses, computed by the caller); cam_K n_cams x 9; img_wh n_cams x 2
 * (width, height as doubles).  out_uv: device (n,2) float64, zeros where no camera sees the
 * point; out_cam_idx: device (n,) int64, -1 there.  At most PCACC_MAX_CAMS cameras. */
int pcacc_project_cameras(pcacc_t h, const double *pc_ego_dev, int64_t n, int64_t stride,
                          const double *glob_from_ego, const double *cam_from_glob,
                          const double *cam_K, const double *img_wh, int n_cams, double depth_thres,
                          double *out_uv_dev, int64_t *out_cam_idx_dev, void *stream);

/* pts_feat_from_img, datasets/nuscenes_utils.py:181-214, both branches: for n points with pixel
 * coordinates uv_dev (n,2) float64 and an image img_dev (H,W,C) of img_dtype, out_dev (n,C)
 * float64 = img[rint(v), rint(u)] (bilinear = 0, half-to-even like np.round) or the four-neighbour
 * blend with the reference's weights in its operation order (bilinear = 1; an integral u or v
 * gives the reference's 0/0).  The reference's bounds assertion (1 < uv < wh - 1) comes back as
 * PCACC_FLAG_UV_OUT_OF_IMAGE from pcacc_sync.  The reference multiplies (N,) weights with (N,C)
 * features, which numpy only broadcasts for a 2-D image; every channel gets the per-point weights
 * here. */
int pcacc_pts_feat_from_img(pcacc_t h, const double *uv_dev, int64_t n, const void *img_dev,
                            int img_dtype, int img_h, int img_w, int channels, int bilinear,
                            double *out_dev, void *stream);

/* static_obj_partitioning_by_elev, bev_generator/sem_bev.py:556-591: pc_dev (n,10) float64 rows
 * whose columns 0, 1 are grid coordinates; per-cell minimum z (row P-1-j, column i), then column 8
 * of every point more than elev_thresh above its cell's minimum is set to 1 IN PLACE.
 * elevmap_dev (P,P) float64 (0 where unobserved), obs_mask_dev (P,P) uint8, scratch_dev: P*P
 * uint64.  Grid coordinates outside [-P, P) (the reference's IndexError) or a NaN z raise
 * PCACC_FLAG_ATTR_RANGE. */
int pcacc_static_obj_partitioning(pcacc_t h, double *pc_dev, int64_t n, int P, double elev_thresh,
                                  double *elevmap_dev, uint8_t *obs_mask_dev,
                                  unsigned long long *scratch_dev, void *stream);

/* ---- the rasteriser's steps as stand-alone operators --------------------------------------
 * pcacc_rasterise fuses these; the reference exposes each as a public method, so each has an
 * entry point with the method's argument meaning.  Clouds are (n, cols) float64 rows on the
 * device, x, y, z in columns 0..2. */

/* get_elevation_map, bev_generator/sem_bev.py:535-553: per-cell minimum z (row P-1-j, column i;
 * 0 where unobserved) and the observed mask; index rules and flags of
 * pcacc_static_obj_partitioning. */
int pcacc_elevation_map(pcacc_t h, const double *pc_dev, int64_t n, int cols, int P,
                        double *elevmap_dev, uint8_t *obs_mask_dev, unsigned long long *scratch_dev,
                        void *stream);

/* velo2frame, sem_pc_accum.py:347-366: out_dev (n,3) float64 = rows of the (3,4) host matrix P
 * times [x y z 1] for the first three columns of pts_dev (n, stride) float32 (pts_f64 = 0) or
 * float64 (1). */
int pcacc_velo2frame(pcacc_t h, const void *pts_dev, int pts_f64, int64_t n, int stride, const double *P,
                     double *out_dev, void *stream);

/* preprocess_pc_and_trajs (cloud part), geometric_transform and crop_view,
 * bev_generator/bev_generator.py:127-160, 207-256, 737-747, in the reference's order: rotation by the
 * (3,3) host matrix rot and translation (skipped when rot is NULL), strict crop of x then y to
 * +-crop_view/2 (NaN = no crop), z < height_filter (NaN = off), pos2grid of columns 0, 1 over
 * grid_view metres and P pixels (NaN = stay metric).  Kept rows go to out_dev in input order with
 * their other columns untouched; *n_kept_dev = their number. */
int pcacc_preprocess_pc(pcacc_t h, const double *pc_dev, int64_t n, int cols, const double *rot,
                        double trans_dx, double trans_dy, double crop_view, double height_filter,
                        double grid_view, int P, double *out_dev, int64_t *n_kept_dev, void *stream);

/* partition_semantic_pc + gen_gridmap_count_map (+ gen_sem_probmap / gen_intensity_map),
 * bev_generator/bev_generator.py:373-453: a point is "selected" when column sem_col equals one of
 * sems[0..n_sems) (n_sems < 0: every point; at most PCACC_MAX_SEMS = 32 classes).
 * np.histogram2d's binning over [0,P]x[0,P] of the grid coordinates in columns 1, 0, flipped along
 * the first axis.  Any of the three (P,P) float64
 * outputs may be NULL: count_sel_dev / count_rest_dev = number of selected / other points per cell,
 * wsum_sel_dev = sum over the selected points of column weight_col (or of weights_dev[point] when
 * weight_col < 0; order-dependent rounding, ~1e-16 relative).  finish: 0 = raw maps; 1 = the
 * Dirichlet expectation with a uniform prior of the two counts, in place (gen_sem_probmap);
 * 2 = wsum / (count_sel + 1) in place (gen_intensity_map). */
#define PCACC_MAX_SEMS 32
int pcacc_cell_stats(pcacc_t h, const double *pc_dev, int64_t n, int cols, int P, int sem_col,
                     const int32_t *sems, int n_sems, int weight_col, const double *weights_dev,
                     int finish, double *count_sel_dev, double *count_rest_dev, double *wsum_sel_dev,
                     void *stream);

/* partition_semantic_pc, bev_generator/bev_generator.py:411-432: rows whose column sem_col equals
 * one of sems[0..n_sems) go to out_sel_dev, the others to out_rest_dev (both (n, cols), input
 * order); *n_sel_dev = number of selected rows (the others number n - that). */
int pcacc_partition_semantic_pc(pcacc_t h, const double *pc_dev, int64_t n, int cols, int sem_col,
                                const int32_t *sems, int n_sems, double *out_sel_dev,
                                double *out_rest_dev, int64_t *n_sel_dev, void *stream);

/* dirichlet_dist_expectation, bev_generator/bev_generator.py:455-481, in place on maps_dev
 * (n_maps, cells) float64. */
int pcacc_dirichlet_expectation(pcacc_t h, double *maps_dev, int n_maps, int64_t cells,
                                double obs_weight, void *stream);

/* road_marking_transform (sigmoid_only = 0) and sigmoid (1), bev_generator/sem_bev.py:593-617,
 * elementwise on n float64 values. */
int pcacc_road_marking(pcacc_t h, const double *in_dev, int64_t n, double int_scaler,
                       double int_sep_scaler, double int_mid_threshold, int sigmoid_only,
                       double *out_dev, void *stream);

/* ---- accounting / profiling (bench.py: gpu_launches, roofline) -------------
 * Kernel classes of this library. */
#define PCACC_K_INTEGRATE 0 /* k_integrate_frustum / _gt / _records / _cloud, k_gen_semantic_pc, k_project */
#define PCACC_K_REBASE 1    /* k_rebase_lazy, k_materialise */
#define PCACC_K_MARK 2      /* k_mark_dynamic */
#define PCACC_K_BIN 3       /* k_bev_bin (exact per-candidate work) */
#define PCACC_K_SCAN 4      /* k_scan */
#define PCACC_K_SCATTER 5   /* k_bev_scatter */
#define PCACC_K_REDUCE 6    /* k_bev_consts + k_bev_reduce (empty and small cells) */
#define PCACC_K_EXPORT 7    /* k_export_frame, k_warp_planes */
#define PCACC_K_REDUCE_BIG 8 /* k_bev_reduce_big (queued large cells) */
#define PCACC_K_CLASSIFY 9  /* k_bev_classify (streaming crop test over the ring) */
#define PCACC_N_KERNELS 10

/* class_mask: bit k set = every launch of kernel class k of this handle is bracketed by CUDA
 * events on its launch stream (two cudaEventRecord calls per launch: time only the classes
 * you need inside a throughput measurement).  0 = off, PCACC_PROFILE_ALL = every class. */
#define PCACC_PROFILE_ALL ((1 << PCACC_N_KERNELS) - 1)
int pcacc_profile(pcacc_t h, int class_mask);
/* Synchronises, then returns and clears the accumulated per-class device time
 * (ms, from the events) and launch counts since the last read.  launches[] is
 * counted whether or not event timing is enabled. */
int pcacc_profile_read(pcacc_t h, double ms[PCACC_N_KERNELS], int64_t launches[PCACC_N_KERNELS]);

#ifdef __cplusplus
}
#endif
#endif /* PCACC_H_ */
