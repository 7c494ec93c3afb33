// vbs_ctx.h - internal context and launcher declarations of libvbs_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/vbs.h"
#include "vbs_geom.h"
#include <nvtx3/nvToolsExt.h>

// NVTX range around the launches of one stage (header-only NVTX 3: a no-op unless a profiler is attached;
// `ncu --nvtx --nvtx-include "vbs:blur/"` or an nsys timeline then groups the kernels by stage)
struct VbsRange {
    explicit VbsRange(const char *name) { nvtxRangePushA(name); }
    ~VbsRange() { nvtxRangePop(); }
    VbsRange(const VbsRange &) = delete;
    VbsRange &operator=(const VbsRange &) = delete;
};

// device-side status bits (OR-ed into ctx->d_status by kernels)
enum : uint32_t {
    VBS_DEV_LABEL_OVERFLOW = 1u << 0,    // more ring components than max_markers
    VBS_DEV_CONTOUR_OVERFLOW = 1u << 1,  // more opened blobs than max_markers
    VBS_DEV_RECHECK_OVERFLOW = 1u << 2,  // float64 recheck list full
    VBS_DEV_TRACE_GUARD = 1u << 3,       // border following did not close
    VBS_DEV_MATCH_CONFLICT = 1u << 4,    // one centroid claimed by two contours
    VBS_DEV_TMA_TIMEOUT = 1u << 5,       // a TMA tile never arrived (bounded wait gave up)
};

struct VbsBranch {       // constants that switch on the frame height (MD:117-126,129,170)
    int ks, kl;          // blur sizes: 21/35 or 39/101
    int tl;              // template length 33 / 80
    double tsigma;       // 7.4 / 13
    int lo, hi;          // inRange bounds 35..180 / 20..200
    int nb;              // max/min filter size 8 / 14
};

struct vbs_ctx {
    vbs_config cfg;
    int H, W, WW, C, B, M, Rcap;
    int big;                         // 1: height > 480 branch
    VbsBranch br;
    cudaStream_t stream, own_stream;
    std::string err;
    int64_t launches;
    int no_tma; int64_t tma_launches;      // VBS_NO_TMA=1 forces the generic loader; launches that used the TMA path
    // opt-in tensor-core blur (k_blur_tc.cu): VBS_BLUR_TC=1 or vbs_set_blur_tc
    int blur_tc; int64_t tc_launches;
    uint8_t *tc_a1, *tc_a2;                // operator matrices [nstrips][128][256], [2][128][256]


    // optional lens correction before K1 (MD:93-109)
    int undist_on; double new_k[4];            // fx', fy', cx', cy' of getOptimalNewCameraMatrix(alpha = 0)
    int2 *undist_map; uint8_t *d_undist;       // [H][W] source position in 1/32 px; [B][H][W*C] corrected frames

    // frame staging for the host entry point
    uint8_t *d_frames; size_t frames_bytes;      // two staging buffers of host_chunk frames
    int host_chunk; cudaStream_t copy_stream; cudaEvent_t ev_copied[2], ev_consumed[2];
    // asynchronous host path: two whole-batch staging slots
    uint8_t *d_slots; cudaEvent_t ev_slot_in[2], ev_slot_free[2], ev_slot_done[2]; uint32_t *h_slot_status;
    int64_t submitted; int inflight; int slot_used[2];
    int last_chunk_frames;                       // frames per chunk of the latest chunked batch (scratch ranges of chunk c depend on it)
    int64_t chunk_seq;                           // chunks uploaded so far (staging buffer = chunk_seq & 1, across calls)
    cudaEvent_t ev_bchunk[8]; int bchunk_live[8];    // end of stage B of chunk c of the latest chunked batch
    // bit images [B][H][WW]
    uint32_t *area_bits, *mask_bits, *max_bits, *open_bits;
    uint32_t *area_count;            // [B] set pixels of area_mask
    // NCC tables
    float *thr_lut;                  // [tl*tl+1] interior threshold on G as a function of S
    double *d_n64;                   // [tl] template factor n
    double *d_cn64;                  // [tl+1+16] guarded prefix sums of n
    double st2;                      // (sum n^2)^2 - 1/L^2
    int *d_cnfix;                    // [4][112] shifted fixed-point (2^30) copies of the prefix sums
    int2 *recheck; uint32_t *recheck_n; int recheck_cap;   // float64 recheck list [B][cap]
    // union-find scratch
    int32_t *parent, *parent2;       // [B][H*W] ring maxima / opened image (fg + bg)
    int32_t *nroots, *rootlist, *slot2label;   // [2B] roots per (frame, image), [2B][M] their pixel indices, [B][M] slot -> label
    int32_t *nrec; int4 *recs; int rcap;       // [2B] tile-local components per (frame, image), [2B][rcap] {start pixel, count, sum x, sum y}
    uint8_t *rowflag;                // [2B][ceil(WW/32)][H] 1: the labelling tile of this 1024-px band was closed early above this row
    int32_t *d_nlabels, *d_ncont;    // [B]
    // ring components
    uint32_t *lab_cnt; unsigned long long *lab_sx, *lab_sy;   // [B][M]
    double *centres;                 // [B][M][2] (row, col)
    // contours / ellipses, slot = contour order (descending start pixel)
    int32_t *croot;                  // [B][M] start pixel index
    double *cell;                    // [B][M][6] cx, cy, major, minor, angle, valid
    int32_t *claim;                  // [B][M] how many contours matched each centroid
    int32_t *cmatch;                 // [B][M] matched label index or -1
    uint32_t *cpts; int32_t *cpn;    // [B][M][128] stored contour vertices (x | y<<16), [B][M] vertex counts
    int32_t *euler4, *holes;         // [B] 4 x Euler number of the opened image; holes = blobs - Euler number
    // outputs kept on device
    int32_t *d_nmarkers;             // [B]
    double *marker_xy;               // [B][M][2]
    double *marker_axes;             // [B][M][3]
    // reference state
    int R; double min_dist;
    int32_t *ref_row, *ref_col; double *ref_xy;
    int32_t *row_det; double *row_cxy; double *row_axes;  // [B][R]...
    int32_t *cell_start, *cell_items;                     // [B][8193] first item of every cell, [B][M] marker indices sorted by cell (k_track3d.cu)
    int32_t *cbin_start, *cbin_items;                     // the same for the ring centroids (centre <-> ellipse match, k_contour.cu)
    double *obs;                     // [B][R][3] undistorted u, v and diameter of each observation
    // camera / 3D
    int have_cam; vbs::CameraF64 cam; int warmup; int64_t first_frame; int have_first;
    double *pos3d; uint8_t *pos_flags;      // [B][R][7], [B][R]
    double *last_seen;                       // [R][4] u, v, diameter, frame
    // plane
    int have_plane; int shell; double pscale;
    double *pl_ref, *pl_start, *pl_dvert; uint8_t *pl_use;
    double *plane; int32_t *plane_n;         // [B][4], [B]
    uint32_t *d_status; uint32_t *h_status;  // device flag word, pinned host mirror
    int32_t *d_nrecheck;                     // [B] recheck counts (debug)
    int last_batch;
    int image_ready;                         // per-pixel scratch allocated (lazily: table-only contexts never pay for it)
    int track_cap;                           // frames the marker cell grids (cell_start / cell_items) are sized for
    // optional per-stage timing (events on the context's stream)
    int profiling, prof_pending, prof_chunks;
    cudaEvent_t pev[72];                        // 8 chunks x 9 stage boundaries
    // chunked two-stream pipeline
    cudaStream_t stream_b; cudaEvent_t ev_a[8], ev_b_done; int overlap_device;
    int no_branch_overlap;                      // 1: run the open-mask branch after the NCC instead of beside it
    int sm_count;                               // multiprocessors of the context's device (grid planning)
    int seg_plan;                               // VBS_SEG_PLAN=0: every (frame, strip) split into the same number of row segments (round 1)
    int blur_variant;                           // VBS_BLUR_VARIANT=0: every vertical tap an integer dot product (blur_area_kernel); default 1: column sums for the 101-tap pass (blur_area_cs_kernel)
    int ncc_variant;                            // VBS_NCC_VARIANT: 0 = round 1's kernel (two tap-half threads per column), 1 = thread per column, default 2 = thread per column with horizontal / vertical warp roles
    double stage_ms[7]; int64_t stage_calls;
};
enum { VBS_NSTAGES = 7 };   // blur, ncc, morph, components, contours, track3d, output copies

#define VBS_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return VBS_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#include "vbs_segplan.h"    // VbsSegPlan, vbs_seg_plan: row-segment plan of the strip-marching kernels (blur, NCC)

// launchers (each returns a cudaError_t from cudaGetLastError after the launches)
cudaError_t vbs_launch_blur(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch);
cudaError_t vbs_undistort_setup(vbs_ctx *ctx, const double *K, const double *D, int nd);
cudaError_t vbs_launch_export_maps(vbs_ctx *ctx, int16_t *map1, uint16_t *map2);
cudaError_t vbs_launch_remap(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch, uint8_t *out);
cudaError_t vbs_launch_ncc(vbs_ctx *ctx, int batch);
cudaError_t vbs_ncc_setup(vbs_ctx *ctx);
cudaError_t vbs_launch_prepare(vbs_ctx *ctx, int batch);
cudaError_t vbs_launch_morph(vbs_ctx *ctx, int batch, int which);
cudaError_t vbs_launch_components(vbs_ctx *ctx, int batch, int which);
cudaError_t vbs_launch_contours(vbs_ctx *ctx, int batch, int which);
cudaError_t vbs_launch_track(vbs_ctx *ctx, int batch, int64_t frameno0);
cudaError_t vbs_launch_reconstruct(vbs_ctx *ctx, int batch, int64_t frameno0);
cudaError_t vbs_launch_fix_displacement(vbs_ctx *ctx, double *pos3d, uint8_t *flags, const double *incoming_dev, long long nframes);
cudaError_t vbs_launch_undistort(vbs_ctx *ctx, const double *uv, double *out, int n);
cudaError_t vbs_launch_position(vbs_ctx *ctx, const double *uvd, double *P, uint8_t *ok, int n);
cudaError_t vbs_launch_plane_points(vbs_ctx *ctx, const double *X, const double *Y, const double *Z, int n, double *out);
cudaError_t vbs_launch_pack_masks(vbs_ctx *ctx, const uint8_t *mask, const uint8_t *area, int batch);
cudaError_t vbs_launch_pack_area(vbs_ctx *ctx, const uint8_t *area, int batch);
cudaError_t vbs_launch_unpack(vbs_ctx *ctx, int stage, void *dst, int batch);
int vbs_check_taps(std::string &err);       // baked integer taps == host recipe
void vbs_host_taps(int ksize, double sigma, int *out);   // OpenCV's 8.8 fixed-point Gaussian kernel (k_blur.cu)
cudaError_t vbs_blur_tc_setup(vbs_ctx *ctx);
cudaError_t vbs_launch_blur_tc(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch);
