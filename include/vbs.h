/*
 * vbs.h - C ABI of libvbs_b200.so: the B200-native per-frame marker pipeline.
 *
 * The reference (UPM-ROB-Lab/Vision-basedSensor) has no FFI: its boundary for this path
 * is a set of Python callables.  Each entry point below names the reference callable(s)
 * it replaces (paths under /root/reference/code):
 *   MD = Marker_Tracking/marker_detection.py
 *   R3 = Marker_Calibration/3d_reconstruction.py
 *   FD = ForceDistribution/ForceDistribution.py
 *
 * Conventions
 *   - plain C: pointers, sizes, no torch / C++ types.
 *   - every function returns VBS_OK (0) or a negative vbs_status; vbs_last_error() gives text.
 *   - the caller owns every frame / output buffer; the context owns only its scratch.
 *   - one context = one GPU + one stream; not thread-safe; calls are asynchronous on the
 *     context's stream unless stated otherwise, vbs_sync() waits and reports device-side errors.
 *   - no CPU fallback anywhere: without a CUDA device vbs_create() fails.
 */
#ifndef VBS_H_
#define VBS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vbs_status {
    VBS_OK = 0,
    VBS_ERR_BAD_ARG = -1,   /* ValueError in the Python mirror (MD:38, R3:95,117)            */
    VBS_ERR_CAPACITY = -2,  /* more labels / contours / rechecks than the context was sized for */
    VBS_ERR_CUDA = -3,      /* CUDA runtime failure (no device, launch error, ...)            */
    VBS_ERR_STATE = -4,     /* call order: e.g. 3D stage requested before vbs_set_camera       */
    VBS_ERR_INTERNAL = -5   /* an invariant the parity argument relies on was violated         */
} vbs_status;

typedef struct vbs_ctx vbs_ctx;

typedef struct vbs_config {
    int32_t device;       /* CUDA ordinal                                                     */
    int32_t height;       /* processed (cropped) frame height in pixels                        */
    int32_t width;        /* processed (cropped) frame width in pixels                         */
    int32_t channels;     /* 1 = gray, 3 = BGR (converted like cvtColor, MD:114)               */
    int32_t max_batch;    /* frames per vbs_process_* call (scratch is sized for this)         */
    int32_t max_markers;  /* capacity for labels / contours / markers per frame (<= 8192)      */
    int32_t max_refs;     /* capacity of the reference-state array                             */
} vbs_config;

/* Output block of one batch.  Every pointer may be NULL (that output is skipped).  For
 * vbs_process_device they are device pointers, for vbs_process_host host pointers.
 * B = batch, M = max_markers, R = number of reference entries set by vbs_set_reference. */
typedef struct vbs_outputs {
    int32_t *n_labels;    /* [B]        components of the ring-maxima image (MD:176)            */
    double  *centres;     /* [B][M][2]  (row, col) centroids in label order (MD:181)            */
    int32_t *n_markers;   /* [B]        markers returned by _marker_center (MD:249)             */
    double  *marker_xy;   /* [B][M][2]  'center' = (x=col, y=row), reference output order       */
    double  *marker_axes; /* [B][M][3]  major_axis, minor_axis (float32 values) and angle, which the
                                        reference forms in float64 as angle32 + 90 (MD:213-217,238-243)   */
    int32_t *row_det;     /* [B][R]     index into the marker list, or -1 (MD:369-373)          */
    double  *row_cxy;     /* [B][R][2]  Cx, Cy of the tracking row (MD:386-387)                 */
    double  *row_axes;    /* [B][R][3]  major_axis, minor_axis, angle of the row (MD:388-390)   */
    double  *pos3d;       /* [B][R][7]  X,Y,Z,dX,dY,dZ,displacement (R3:296-307)                */
    uint8_t *pos_flags;   /* [B][R]     bit0: observation enters R3 (row present, major>=min),
                                        bit1: 3D position valid, bit2: displacement row emitted */
    double  *plane;       /* [B][4]     a, b, c, tilt_deg of the contact plane (FD:141-159)     */
    int32_t *plane_n;     /* [B]        points that entered the plane fit                       */
} vbs_outputs;

/* stage images of the most recent batch, for stage-level parity checks (vbs_debug_stage) */
typedef enum vbs_stage {
    VBS_STAGE_AREA_MASK = 0,  /* uint8 [B][H][W] {0,255}  area_mask (MD:129)                   */
    VBS_STAGE_MASK = 1,       /* uint8 [B][H][W] {0,1}    mask = ncc > 0.1 (MD:133)            */
    VBS_STAGE_MAXIMA = 2,     /* uint8 [B][H][W] {0,1}    maxima (MD:172-174)                  */
    VBS_STAGE_LABELS = 3,     /* int32 [B][H][W]          labeled (MD:176)                     */
    VBS_STAGE_OPENED = 4,     /* uint8 [B][H][W] {0,255}  area_mask after the 5x5 open (MD:195)*/
    VBS_STAGE_RECHECKS = 5,   /* int32 [B]                pixels re-decided in float64         */
    VBS_STAGE_ELLIPSES = 6,   /* float64 [B][M][6]        per external contour of the opened mask, in cv2.findContours order (MD:196-220):
                                                          cx, cy, major, minor, angle (+90 as MD:213-217), 1 = kept (>= 5 points, minor >= 5) */
    VBS_STAGE_NCONTOURS = 7   /* int32 [B]                external contours of the opened mask  */
} vbs_stage;

/* lifetime ---------------------------------------------------------------------------------- */
int  vbs_create(vbs_ctx **out, const vbs_config *cfg);      /* replaces MarkerTracker.__init__ (MD:15-31) */
void vbs_destroy(vbs_ctx *ctx);                             /* replaces _cleanup (MD:470-474)             */
const char *vbs_last_error(const vbs_ctx *ctx);
int  vbs_set_stream(vbs_ctx *ctx, void *cuda_stream);       /* run on the caller's stream; 0 = the context's own
                                                                non-blocking stream, cudaStreamLegacy (0x1) = the legacy default stream */
int  vbs_sync(vbs_ctx *ctx);                                /* wait; returns device-side status of all work so far */
const char *vbs_version(void);

/* reference state ----------------------------------------------------------------------------
 * replaces self.first_frame_markers (MD:31,289-347) + config['min_marker_distance'] (MD:359).
 * Entries are matched in array order (= the reference's dict order).                          */
int vbs_set_reference(vbs_ctx *ctx, int32_t n, const int32_t *row, const int32_t *col,
                      const double *ox, const double *oy, double min_marker_distance);

/* camera: replaces MarkerAnalysis.load_parameters results (R3:87-124) and Config (R3:21-24).
 * K, R row-major 3x3; D = k1,k2,p1,p2,k3; all float32 like the reference stores them.         */
int vbs_set_camera(vbs_ctx *ctx, const float K[9], const float D[5], const float R[9],
                   const float T[3], double marker_diameter_mm, double min_marker_size_px,
                   double max_displacement, int32_t warmup_frames);

/* plane fit inputs: replaces df_ref / d_vert of process_marker_data (FD:173-204) and the
 * 'plane' / 'shell' switch (FD:15,222).  n must equal the reference-array length.
 *   ref_xyz  [n][3]  theoretical marker coordinates (FD:29-95)
 *   start_xyz[n][3]  P_start of the tilted run (displacement = P(frame) - P_start)
 *   d_vert   [n][3]  displacement of the vertical baseline run; NULL = zeros
 *   use      [n]     1 = marker takes part (common_ids, FD:184); NULL = all                    */
int vbs_set_plane(vbs_ctx *ctx, int32_t n, const double *ref_xyz, const double *start_xyz,
                  const double *d_vert, const uint8_t *use, int32_t shell_mode, double scale);

/* forget the last-seen observations (R3:252 marker_dict) and the first-frame number (R3:255). */
int vbs_reset_sequence(vbs_ctx *ctx);
/* last-seen table exchange for frame-sharded multi-GPU runs: [R][4] = u, v, diameter, frame (-1 = never) */
int vbs_get_last_seen(vbs_ctx *ctx, double *host_table);
int vbs_set_last_seen(vbs_ctx *ctx, const double *host_table);
/* frame-sharded runs: a shard is processed with an empty last-seen table; once the table arriving
 * from the preceding shards is known, this emits the one displacement row per reference entry that
 * was missing (its first observation in the shard).  pos3d [N][R][7] / pos_flags [N][R] are the
 * shard's device-resident outputs, incoming [R][4] is a host table as vbs_get_last_seen returns it. */
int vbs_fix_displacement(vbs_ctx *ctx, double *pos3d_device, uint8_t *pos_flags_device, int64_t nframes,
                         const double *incoming_host);
/* frame-sharded runs: the warm-up window (R3:255-256) counts from the GLOBAL first frame number */
int vbs_set_first_frame(vbs_ctx *ctx, int64_t first_frame);

/* the hot path --------------------------------------------------------------------------------
 * One call = MD:440-449 (crop view, _find_markers, _marker_center, _track_markers) for every
 * frame of the batch, then R3:259-307 (undistort, 3D position, last-seen displacement) and
 * FD:141-159 (plane tilt) when camera / plane inputs are set.
 *   frames       first byte of the (cropped) first frame
 *   frame_stride bytes between consecutive frames
 *   row_pitch    bytes between consecutive rows
 *   frameno0     frame number of the first frame of the batch (CSV column 'frameno')
 * vbs_process_device: frames and outputs live in device memory, nothing is copied.
 * vbs_process_host  : frames live in host memory (pinned for full speed), outputs in host or device
 *                     memory; the H2D / D2H copies are part of the call, which returns after the
 *                     results have landed.                                                      */
int vbs_process_device(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride,
                       int64_t row_pitch, int64_t frameno0, const vbs_outputs *out);
int vbs_process_host(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride,
                     int64_t row_pitch, int64_t frameno0, const vbs_outputs *out);

/* optional for vbs_process_device (default off; vbs_process_host always does it): cut a batch of >= 64
 * frames into chunks and run the long detection kernels of chunk c+1 beside the short latency-bound
 * kernels of chunk c on a second stream */
int vbs_set_overlap(vbs_ctx *ctx, int32_t enable);
/* frames per chunk of the host entry points' copy/compute overlap.  vbs_process_host: 0 = default 64.
 * vbs_submit_host: 0 = copy each batch in one piece into one of two max_batch-sized staging slots and run the
 * unchunked pipeline (fastest on one GPU); > 0 = chunked like vbs_process_host, the chunk rotation continuing
 * across batches (less host memory traffic in flight per rank: faster when several ranks share one host).
 * A batch always fits: the chunk is raised to ceil(batch / 8) when needed.  Not while batches are in flight. */
int vbs_set_host_chunk(vbs_ctx *ctx, int32_t frames_per_chunk);

/* Asynchronous host entry point for streams of batches (config 5): vbs_submit_host enqueues the H2D copies of
 * the batch on a copy stream, the pipeline and the copies of the results, and returns at once; up to two
 * batches may be in flight, so batch i+1 crosses PCIe while batch i is processed.  vbs_wait_host blocks until
 * the oldest batch in flight has landed in its `out` arrays and returns its status.  Frames must be pinned
 * host memory; the `out` pointers of the host entry points may be pinned host memory OR device memory (e.g. to
 * gather the records of several GPUs with NCCL before one copy to the host); both must stay untouched until the
 * matching vbs_wait_host. */
int vbs_submit_host(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch,
                    int64_t frameno0, const vbs_outputs *out);
int vbs_wait_host(vbs_ctx *ctx);

/* optional lens correction of the cropped frame, MarkerTracker._undistort_frame (MD:93-109), applied by
 * _preprocess_frame (MD:88-89) when config['calibration_params'] is present:
 *   new_K = cv2.getOptimalNewCameraMatrix(K, D, (w,h), 0, (w,h)); maps = cv2.initUndistortRectifyMap(K, D, None,
 *   new_K, (w,h), CV_16SC2); frame = cv2.remap(frame, maps, INTER_LINEAR)            (bit-exact, BORDER_CONSTANT 0)
 * K: 3x3 row-major float64, D: nd = 4, 5 or 8 float64 coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6]]).  Once set, every
 * vbs_process_* / vbs_find_markers call corrects its frames first; K == NULL switches it off.  The maps depend on
 * (K, D, size) only and are built once here (the reference rebuilds them for every frame).
 * vbs_get_undistort_maps: new_K (host, 9 doubles) and the maps in OpenCV's layout (device: map1 int16 [H][W][2],
 *                         map2 uint16 [H][W]); any pointer may be NULL (maps: both or neither)
 * vbs_undistort_frames  : the corrected frames themselves, [batch][H][W*C] uint8, device                         */
int vbs_set_undistort(vbs_ctx *ctx, const double *K, const double *D, int32_t nd);
int vbs_get_undistort_maps(vbs_ctx *ctx, double *new_camera_matrix, int16_t *map1_device, uint16_t *map2_device);
int vbs_undistort_frames(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch,
                         uint8_t *out_device);

/* stage-level entry points (same kernels, for the static-method mirrors) -----------------------
 * vbs_find_markers : MarkerTracker._find_markers (MD:111-135): frames -> area_mask, mask (kept in ctx)
 * vbs_marker_center: MarkerTracker._marker_center (MD:166-249) on masks supplied by the caller
 *                    (uint8, device, [B][H][W]; mask nonzero = 1, area nonzero = 255)            */
int vbs_find_markers(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch);
int vbs_marker_center(vbs_ctx *ctx, const uint8_t *mask, const uint8_t *area_mask, int32_t batch,
                      const vbs_outputs *out);
/* vbs_ncc_mask     : the second half of _find_markers alone (MD:132-133 with MarkerTracker._normxcorr2, MD:146-164):
 *                    mask = normxcorr2(gkern, area_mask) > 0.1 for area masks supplied by the caller (uint8, device,
 *                    [B][H][W], nonzero = 255).  The mask stays in the context: read it with
 *                    vbs_debug_stage(VBS_STAGE_MASK), the float64 re-decisions with VBS_STAGE_RECHECKS.            */
int vbs_ncc_mask(vbs_ctx *ctx, const uint8_t *area_mask, int32_t batch);

/* table-level entry points: host arrays in and out, synchronous.  They run the same kernels as
 * vbs_process_* on rows / points supplied by the caller, for the per-call mirrors of the reference:
 * vbs_track_markers    : MarkerTracker._track_markers (MD:349-396) on one marker list
 *                        xy [n][2], axes [n][3] -> row_det [R], row_cxy [R][2], row_axes [R][3]
 * vbs_reconstruct_rows : MarkerAnalysis._track_markers (R3:240-316) + plane fit on tracking rows
 *                        row_det [B][R] (>= 0: row present), row_cxy [B][R][2], row_axes [B][R][3]
 * vbs_undistort_points : MarkerAnalysis._undistort_points (R3:185-193), uv [n][2]
 * vbs_position_3d      : MarkerAnalysis._calculate_3d_position (R3:195-238), uvd [n][3] = u, v, diameter;
 *                        ok[i] = 0 where the reference raises
 * vbs_fit_plane        : fit_plane_least_squares (FD:141-159) -> a, b, c, tilt_deg                   */
int vbs_track_markers(vbs_ctx *ctx, int32_t n, const double *marker_xy, const double *marker_axes, int32_t *row_det,
                      double *row_cxy, double *row_axes);
int vbs_reconstruct_rows(vbs_ctx *ctx, int32_t batch, int64_t frameno0, const int32_t *row_det, const double *row_cxy,
                         const double *row_axes, double *pos3d, uint8_t *pos_flags, double *plane, int32_t *plane_n);
int vbs_undistort_points(vbs_ctx *ctx, int32_t n, const double *uv, double *out);
int vbs_position_3d(vbs_ctx *ctx, int32_t n, const double *uvd, double *P, uint8_t *ok);
int vbs_fit_plane(vbs_ctx *ctx, int32_t n, const double *X, const double *Y, const double *Z, double out[4]);

/* copy the stage images of the most recent batch into dst (device memory): the first frames of the batch, as many as
 * `bytes` holds (at least one whole frame) */
int vbs_debug_stage(vbs_ctx *ctx, int32_t stage, void *dst_device, size_t bytes);

/* per-stage device timing for bench.py's roofline: CUDA events on the context's stream around
 * each stage of vbs_process_*.  ms[7] = accumulated milliseconds of blur+DoG, NCC, morphology,
 * components, contours+ellipse, tracking+3D+plane, output copies; *calls = batches accumulated. */
int vbs_set_profiling(vbs_ctx *ctx, int32_t enable);
int vbs_get_stage_ms(vbs_ctx *ctx, double ms[7], int64_t *calls);

/* launch accounting for bench.py ("gpu_launches"): kernels launched by this context so far */
int64_t vbs_kernel_launches(const vbs_ctx *ctx);
/* how many of the blur launches staged their input tiles with TMA (cp.async.bulk.tensor); the rest used
 * the generic loader (BGR input, unaligned crop views, or VBS_NO_TMA=1 in the environment) */
int64_t vbs_tma_launches(const vbs_ctx *ctx);

/* OPT-IN experiment (off by default, also switched on by VBS_BLUR_TC=1 in the environment): run the
 * two fixed-point Gaussian blurs of MD:117-126 as banded-Toeplitz int8 GEMMs on the tensor cores (tcgen05.mma kind::i8,
 * accumulators in TMEM, TMA-staged operands) instead of the integer-dot-product kernel.  Same bits out (area_mask is
 * tested equal); gray frames the TMA unit can describe only (16-byte aligned base and pitches, W >= 128, H >= 64) -
 * anything else silently takes the default kernel.  vbs_tc_launches counts the blur launches that took this path. */
int vbs_set_blur_tc(vbs_ctx *ctx, int32_t enable);
int64_t vbs_tc_launches(const vbs_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* VBS_H_ */
