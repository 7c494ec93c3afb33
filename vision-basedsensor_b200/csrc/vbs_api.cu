// vbs_api.cu - the C ABI of include/vbs.h: context lifetime, state setters, and the batch
// entry points that chain the kernels on the context's stream.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "vbs_ctx.h"

namespace {

// zero-filled device allocation: padding entries of the per-frame tables (beyond n_markers / n_labels) read as 0
template <class T> cudaError_t dalloc(T **p, size_t n) {
    *p = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void **)p, bytes);
    return e != cudaSuccess ? e : cudaMemset(*p, 0, bytes);
}

int fail(vbs_ctx *ctx, int code, const char *msg) { ctx->err = msg; return code; }

// Every exported call runs on the context's GPU, whatever device is current in the calling thread, and
// leaves the caller's current device as it found it (two contexts on different GPUs in one process,
// or a caller that switches devices between calls, must not allocate or launch on the wrong GPU).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
        if (cur != dev) {
            ok = cudaSetDevice(dev) == cudaSuccess;
            prev = cur;
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define VBS_ON_DEVICE(ctx)                                                  \
    DeviceGuard guard_((ctx)->cfg.device);                                  \
    if (!guard_.ok) return fail(ctx, VBS_ERR_CUDA, "cudaSetDevice failed for the context's device")

// cell grids of the nearest-marker match (k_track3d.cu), sized for `frames` frames per call
int ensure_track(vbs_ctx *ctx, int frames) {
    if (ctx->track_cap >= frames) return VBS_OK;
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->cell_start) { cudaFree(ctx->cell_start); ctx->cell_start = nullptr; }
    if (ctx->cell_items) { cudaFree(ctx->cell_items); ctx->cell_items = nullptr; }
    ctx->track_cap = 0;
    VBS_CUDA(dalloc(&ctx->cell_start, (size_t)frames * 8193)); VBS_CUDA(dalloc(&ctx->cell_items, (size_t)frames * ctx->M));
    ctx->track_cap = frames;
    return VBS_OK;
}

// Per-pixel scratch of the image pipeline.  Allocated on the first call that processes frames or masks, so a
// context that only serves the table-level entry points (vbs_reconstruct_rows, vbs_undistort_points, ...)
// with a large max_batch never pays for it (recheck lists alone are 128 KiB per frame).
int ensure_image(vbs_ctx *ctx) {
    if (ctx->image_ready) return VBS_OK;
    const size_t B = ctx->B, H = ctx->H, W = ctx->W, WW = ctx->WW, M = ctx->M;
    const size_t nbits = B * H * WW;
    VBS_CUDA(dalloc(&ctx->area_bits, nbits)); VBS_CUDA(dalloc(&ctx->mask_bits, nbits)); VBS_CUDA(dalloc(&ctx->max_bits, nbits));
    VBS_CUDA(dalloc(&ctx->open_bits, nbits));
    VBS_CUDA(dalloc(&ctx->area_count, B));
    VBS_CUDA(dalloc(&ctx->recheck, B * ctx->recheck_cap)); VBS_CUDA(dalloc(&ctx->recheck_n, B));
    VBS_CUDA(dalloc(&ctx->parent, B * H * W)); VBS_CUDA(dalloc(&ctx->parent2, B * H * W));
    VBS_CUDA(dalloc(&ctx->nroots, 2 * B)); VBS_CUDA(dalloc(&ctx->rootlist, 2 * B * M)); VBS_CUDA(dalloc(&ctx->slot2label, B * M));
    VBS_CUDA(dalloc(&ctx->rowflag, 2 * B * ((WW + 31) / 32) * H));
    VBS_CUDA(dalloc(&ctx->nrec, 2 * B)); VBS_CUDA(dalloc(&ctx->recs, 2 * B * (size_t)ctx->rcap));
    VBS_CUDA(dalloc(&ctx->lab_cnt, B * M)); VBS_CUDA(dalloc(&ctx->lab_sx, B * M)); VBS_CUDA(dalloc(&ctx->lab_sy, B * M));
    VBS_CUDA(dalloc(&ctx->centres, B * M * 2));
    VBS_CUDA(dalloc(&ctx->croot, B * M)); VBS_CUDA(dalloc(&ctx->cell, B * M * 6)); VBS_CUDA(dalloc(&ctx->claim, B * M));
    VBS_CUDA(dalloc(&ctx->cmatch, B * M));
    VBS_CUDA(dalloc(&ctx->cpts, B * M * 128)); VBS_CUDA(dalloc(&ctx->cpn, B * M));
    VBS_CUDA(dalloc(&ctx->euler4, B)); VBS_CUDA(dalloc(&ctx->holes, B));
    VBS_CUDA(dalloc(&ctx->cbin_start, B * 8193)); VBS_CUDA(dalloc(&ctx->cbin_items, B * M));
    int rc = ensure_track(ctx, ctx->B);
    if (rc != VBS_OK) return rc;
    ctx->image_ready = 1;
    return VBS_OK;
}

// operator matrices of the opt-in tensor-core blur: built on the context itself (the launchers work on value copies)
int ensure_blur_tc(vbs_ctx *ctx) {
    if (ctx->blur_tc && !ctx->tc_a1) VBS_CUDA(vbs_blur_tc_setup(ctx));
    return VBS_OK;
}

struct Scratch {                       // small device staging buffer, freed on scope exit
    void *p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
};

int map_status(vbs_ctx *ctx, uint32_t st) {
    if (!st) return VBS_OK;
    std::string m = "device status:";
    if (st & VBS_DEV_LABEL_OVERFLOW) m += " ring components exceed max_markers;";
    if (st & VBS_DEV_CONTOUR_OVERFLOW) m += " opened blobs exceed max_markers;";
    if (st & VBS_DEV_RECHECK_OVERFLOW) m += " float64 recheck list overflow;";
    if (st & VBS_DEV_TRACE_GUARD) m += " border following did not close;";
    if (st & VBS_DEV_MATCH_CONFLICT) m += " a centroid was matched by two contours;";
    if (st & VBS_DEV_TMA_TIMEOUT) m += " a TMA tile load timed out;";
    ctx->err = m;
    if (st & (VBS_DEV_TRACE_GUARD | VBS_DEV_MATCH_CONFLICT | VBS_DEV_TMA_TIMEOUT)) return VBS_ERR_INTERNAL;
    return VBS_ERR_CAPACITY;
}

void free_all(vbs_ctx *c) {
    void *ptrs[] = {c->tc_a1, c->tc_a2, c->d_frames, c->undist_map, c->d_undist, c->area_bits, c->mask_bits, c->max_bits, c->open_bits, c->area_count, c->thr_lut,
                    c->d_n64, c->d_cn64, c->d_cnfix, c->recheck, c->recheck_n, c->parent, c->parent2, c->nroots, c->rootlist, c->slot2label, c->nrec, c->recs, c->rowflag, c->d_nlabels,
                    c->d_ncont, c->lab_cnt, c->lab_sx, c->lab_sy, c->centres, c->croot, c->cell, c->claim, c->cmatch, c->cpts, c->cpn, c->euler4, c->holes, c->d_nmarkers,
                    c->marker_xy, c->marker_axes, c->ref_row, c->ref_col, c->ref_xy, c->row_det, c->row_cxy, c->row_axes, c->cell_start, c->cell_items, c->cbin_start, c->cbin_items, c->obs,
                    c->pos3d, c->pos_flags, c->last_seen, c->pl_ref, c->pl_start, c->pl_dvert, c->pl_use, c->plane, c->plane_n,
                    c->d_status};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->h_slot_status) cudaFreeHost(c->h_slot_status);
    if (c->d_slots) cudaFree(c->d_slots);
    for (int i = 0; i < 2; ++i) { if (c->ev_slot_in[i]) cudaEventDestroy(c->ev_slot_in[i]); if (c->ev_slot_free[i]) cudaEventDestroy(c->ev_slot_free[i]); if (c->ev_slot_done[i]) cudaEventDestroy(c->ev_slot_done[i]); }
    for (cudaEvent_t e : c->pev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_a) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_bchunk) if (e) cudaEventDestroy(e);
    if (c->ev_b_done) cudaEventDestroy(c->ev_b_done);
    if (c->stream_b) cudaStreamDestroy(c->stream_b);
    for (int i = 0; i < 2; ++i) { if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
}

// run the detection chain on frames already in device memory
int run_detection(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch) {
    VBS_CUDA(vbs_launch_blur(ctx, frames, batch, frame_stride, row_pitch));
    VBS_CUDA(vbs_launch_ncc(ctx, batch));
    return VBS_OK;
}
int run_centres(vbs_ctx *ctx, int batch) {
    VBS_CUDA(vbs_launch_prepare(ctx, batch));
    VBS_CUDA(vbs_launch_morph(ctx, batch, 3));
    VBS_CUDA(vbs_launch_components(ctx, batch, 3));
    VBS_CUDA(vbs_launch_contours(ctx, batch, 6));
    return VBS_OK;
}

struct CopyPlan { void *dst; const void *src; size_t bytes; };

int plan_outputs(vbs_ctx *ctx, const vbs_outputs *out, int batch, CopyPlan *plan) {
    int n = 0;
    const size_t B = batch, M = ctx->M, R = ctx->R;
    if (!out) return 0;
    auto add = [&](void *dst, const void *src, size_t bytes) { if (dst && bytes) plan[n++] = CopyPlan{dst, src, bytes}; };
    add(out->n_labels, ctx->d_nlabels, B * sizeof(int32_t));
    add(out->centres, ctx->centres, B * M * 2 * sizeof(double));
    add(out->n_markers, ctx->d_nmarkers, B * sizeof(int32_t));
    add(out->marker_xy, ctx->marker_xy, B * M * 2 * sizeof(double));
    add(out->marker_axes, ctx->marker_axes, B * M * 3 * sizeof(double));
    if (R > 0) {
        add(out->row_det, ctx->row_det, B * R * sizeof(int32_t));
        add(out->row_cxy, ctx->row_cxy, B * R * 2 * sizeof(double));
        add(out->row_axes, ctx->row_axes, B * R * 3 * sizeof(double));
        if (ctx->have_cam) {
            add(out->pos3d, ctx->pos3d, B * R * 7 * sizeof(double));
            add(out->pos_flags, ctx->pos_flags, B * R);
            if (ctx->have_plane) {
                add(out->plane, ctx->plane, B * 4 * sizeof(double));
                add(out->plane_n, ctx->plane_n, B * sizeof(int32_t));
            }
        }
    }
    return n;
}

// ---- chunked two-stream pipeline -------------------------------------------------------------------
// Stage A (blur + NCC: long, issue-bound kernels) of chunk c+1 runs on the caller's stream while stage B
// (morphology, components, contours, tracking, 3D, plane, output copies: many short latency-bound
// kernels) of chunk c runs on a second, higher-priority stream; chunks use disjoint frame ranges of the
// context's scratch, so a "view" (pointer-offset copy of the context) is all a launcher needs.
constexpr int VBS_MAX_CHUNKS = 8;
constexpr int VBS_EV_PER_CHUNK = 9;       // A: blur start, ncc start, ncc end; B: morph, comp, contour, track, copies, end

int prof_collect(vbs_ctx *ctx) {               // fold the previous batch's events into the totals
    if (!ctx->prof_pending) return VBS_OK;
    for (int c = 0; c < ctx->prof_chunks; ++c) {
        cudaEvent_t *e = ctx->pev + c * VBS_EV_PER_CHUNK;
        VBS_CUDA(cudaEventSynchronize(e[8]));
        const int lo[VBS_NSTAGES] = {0, 1, 3, 4, 5, 6, 7}, hi[VBS_NSTAGES] = {1, 2, 4, 5, 6, 7, 8};
        for (int i = 0; i < VBS_NSTAGES; ++i) {
            float ms = 0.f;
            VBS_CUDA(cudaEventElapsedTime(&ms, e[lo[i]], e[hi[i]]));
            ctx->stage_ms[i] += ms;
        }
    }
    ctx->stage_calls += 1;
    ctx->prof_pending = 0;
    return VBS_OK;
}

#define VBS_MARKC(c, i, st) do { if (ctx->profiling) VBS_CUDA(cudaEventRecord(ctx->pev[(c) * VBS_EV_PER_CHUNK + (i)], st)); } while (0)

// pointer-offset copy of the context for frames [off, off + n) on stream st
vbs_ctx make_view(const vbs_ctx *c, int off, cudaStream_t st) {
    vbs_ctx v = *c;
    const size_t o = (size_t)off, HW = (size_t)c->H * c->W, HWW = (size_t)c->H * c->WW, M = c->M, R = c->Rcap;
    v.stream = st;
    v.area_bits += o * HWW; v.mask_bits += o * HWW; v.max_bits += o * HWW; v.open_bits += o * HWW;
    v.area_count += o; v.recheck += o * c->recheck_cap; v.recheck_n += o;
    v.parent += o * HW; v.parent2 += o * HW;
    if (v.d_undist) v.d_undist += o * HW * c->C;
    v.nroots += 2 * o; v.rootlist += 2 * o * M; v.slot2label += o * M;
    v.nrec += 2 * o; v.recs += 2 * o * (size_t)c->rcap; v.rowflag += 2 * o * (size_t)((c->WW + 31) / 32) * c->H;
    v.d_nlabels += o; v.d_ncont += o;
    v.lab_cnt += o * M; v.lab_sx += o * M; v.lab_sy += o * M; v.centres += o * M * 2;
    v.croot += o * M; v.cell += o * M * 6; v.claim += o * M; v.cmatch += o * M; v.cpts += o * M * 128; v.cpn += o * M;
    v.euler4 += o; v.holes += o;
    v.d_nmarkers += o; v.marker_xy += o * M * 2; v.marker_axes += o * M * 3;
    v.cell_start += o * 8193; v.cell_items += o * M; v.cbin_start += o * 8193; v.cbin_items += o * M;
    v.row_det += o * R; v.row_cxy += o * R * 2; v.row_axes += o * R * 3; v.obs += o * R * 3;
    v.pos3d += o * R * 7; v.pos_flags += o * R; v.plane += o * 4; v.plane_n += o;
    return v;
}
void fold_view(vbs_ctx *c, const vbs_ctx &v) {   // state a launcher may have changed
    c->launches = v.launches; c->tma_launches = v.tma_launches; c->tc_launches = v.tc_launches; c->have_first = v.have_first; c->first_frame = v.first_frame;
    if (!v.err.empty()) c->err = v.err;
}

vbs_outputs offset_outputs(const vbs_ctx *ctx, const vbs_outputs *out, int off) {
    vbs_outputs o;
    std::memset(&o, 0, sizeof(o));
    if (!out) return o;
    o = *out;
    const size_t M = ctx->M, R = ctx->R, f = (size_t)off;
    if (o.n_labels) o.n_labels += f;
    if (o.centres) o.centres += f * M * 2;
    if (o.n_markers) o.n_markers += f;
    if (o.marker_xy) o.marker_xy += f * M * 2;
    if (o.marker_axes) o.marker_axes += f * M * 3;
    if (o.row_det) o.row_det += f * R;
    if (o.row_cxy) o.row_cxy += f * R * 2;
    if (o.row_axes) o.row_axes += f * R * 3;
    if (o.pos3d) o.pos3d += f * R * 7;
    if (o.pos_flags) o.pos_flags += f * R;
    if (o.plane) o.plane += f * 4;
    if (o.plane_n) o.plane_n += f;
    return o;
}

// output copies of a view: the view's scratch pointers are already offset, the destination is not
int copy_outputs(vbs_ctx *ctx, vbs_ctx &v, const vbs_outputs *out, int off, int n, cudaMemcpyKind kind) {
    if (!out) return VBS_OK;
    const vbs_outputs o = offset_outputs(ctx, out, off);
    CopyPlan plan[16];
    const int np = plan_outputs(&v, &o, n, plan);
    for (int i = 0; i < np; ++i) VBS_CUDA(cudaMemcpyAsync(plan[i].dst, plan[i].src, plan[i].bytes, kind, v.stream));
    return VBS_OK;
}

int ensure_pipeline(vbs_ctx *ctx) {
    if (ctx->stream_b) return VBS_OK;
    int lo = 0, hi = 0;
    VBS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VBS_CUDA(cudaStreamCreateWithPriority(&ctx->stream_b, cudaStreamNonBlocking, hi));
    for (int i = 0; i < VBS_MAX_CHUNKS; ++i) VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_a[i], cudaEventDisableTiming));
    VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_b_done, cudaEventDisableTiming));
    return VBS_OK;
}

// stage A of one chunk on the context's stream
int stage_a(vbs_ctx *ctx, int c, int off, int n, const uint8_t *frames, int64_t frame_stride, int64_t row_pitch) {
    vbs_ctx v = make_view(ctx, off, ctx->stream);
    VBS_MARKC(c, 0, v.stream);
    cudaError_t e = vbs_launch_blur(&v, frames, n, frame_stride, row_pitch);
    VBS_MARKC(c, 1, v.stream);
    if (e == cudaSuccess) e = vbs_launch_ncc(&v, n);
    VBS_MARKC(c, 2, v.stream);
    fold_view(ctx, v);
    VBS_CUDA(e);
    return VBS_OK;
}
// stage B of one chunk on stream st
int stage_b(vbs_ctx *ctx, int c, int off, int n, int64_t frameno0, const vbs_outputs *out, cudaMemcpyKind kind, cudaStream_t st) {
    vbs_ctx v = make_view(ctx, off, st);
    VBS_MARKC(c, 3, st);
    cudaError_t e = vbs_launch_prepare(&v, n);
    if (e == cudaSuccess) e = vbs_launch_morph(&v, n, 3);
    VBS_MARKC(c, 4, st);
    if (e == cudaSuccess) e = vbs_launch_components(&v, n, 3);
    VBS_MARKC(c, 5, st);
    if (e == cudaSuccess) e = vbs_launch_contours(&v, n, 6);
    VBS_MARKC(c, 6, st);
    if (e == cudaSuccess) e = vbs_launch_track(&v, n, frameno0 + off);
    VBS_MARKC(c, 7, st);
    fold_view(ctx, v);
    VBS_CUDA(e);
    int rc = copy_outputs(ctx, v, out, off, n, kind);
    VBS_MARKC(c, 8, st);
    return rc;
}

// chunk plan of a batch: up to 4 chunks of >= 32 frames (smaller batches run as one chunk on one stream)
int plan_chunks(int batch, int *chunk) {
    int nch = batch / 32;
    if (nch > 4) nch = 4;
    if (nch < 1) nch = 1;
    *chunk = (batch + nch - 1) / nch;
    return (batch + *chunk - 1) / *chunk;
}

int process_common(vbs_ctx *ctx, const uint8_t *d_frames, int batch, int64_t frame_stride, int64_t row_pitch, int64_t frameno0,
                   const vbs_outputs *out, cudaMemcpyKind kind) {
    int rc;
    if (ctx->profiling && (rc = prof_collect(ctx)) != VBS_OK) return rc;
    int chunk = batch;
    // measured on B200 (1080p, batch 256): chunking costs more (shorter column segments in the marching
    // kernels, 4x the launches, SM contention) than the overlap wins - 12.3 vs 10.7 ms - so device-resident
    // batches run unchunked unless asked; the host path below always pipelines (it is PCIe-bound)
    const int nch = ctx->overlap_device ? plan_chunks(batch, &chunk) : 1;
    if (nch <= 1 && !ctx->no_branch_overlap && ensure_pipeline(ctx) == VBS_OK) {
        // The blobs of the opened AREA mask (5x5 open, labelling, border following, ellipse fits) depend on
        // K1 only, so that branch runs on the second stream beside the NCC; the branches join at the
        // centroid <-> ellipse matching.  Stage marks stay on the caller's stream (their sum = the step).
        vbs_ctx va = make_view(ctx, 0, ctx->stream);
        VBS_MARKC(0, 0, ctx->stream);
        cudaError_t e = vbs_launch_prepare(&va, batch);
        if (e == cudaSuccess) e = vbs_launch_blur(&va, d_frames, batch, frame_stride, row_pitch);
        VBS_MARKC(0, 1, ctx->stream);
        fold_view(ctx, va);
        VBS_CUDA(e);
        VBS_CUDA(cudaEventRecord(ctx->ev_a[0], ctx->stream));
        VBS_CUDA(cudaStreamWaitEvent(ctx->stream_b, ctx->ev_a[0], 0));
        vbs_ctx vb = make_view(ctx, 0, ctx->stream_b);
        e = vbs_launch_morph(&vb, batch, 2);
        if (e == cudaSuccess) e = vbs_launch_components(&vb, batch, 2);
        if (e == cudaSuccess) e = vbs_launch_contours(&vb, batch, 2);
        fold_view(ctx, vb);
        VBS_CUDA(e);
        VBS_CUDA(cudaEventRecord(ctx->ev_b_done, ctx->stream_b));
        va.launches = ctx->launches; va.tma_launches = ctx->tma_launches; va.tc_launches = ctx->tc_launches;
        e = vbs_launch_ncc(&va, batch);
        VBS_MARKC(0, 2, ctx->stream); VBS_MARKC(0, 3, ctx->stream);
        if (e == cudaSuccess) e = vbs_launch_morph(&va, batch, 1);
        VBS_MARKC(0, 4, ctx->stream);
        if (e == cudaSuccess) e = vbs_launch_components(&va, batch, 1);
        VBS_MARKC(0, 5, ctx->stream);
        fold_view(ctx, va);
        VBS_CUDA(e);
        VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b_done, 0));       // join: ellipses are ready
        e = vbs_launch_contours(&va, batch, 4);
        VBS_MARKC(0, 6, ctx->stream);
        if (e == cudaSuccess) e = vbs_launch_track(&va, batch, frameno0);
        VBS_MARKC(0, 7, ctx->stream);
        fold_view(ctx, va);
        VBS_CUDA(e);
        if ((rc = copy_outputs(ctx, va, out, 0, batch, kind)) != VBS_OK) return rc;
        VBS_MARKC(0, 8, ctx->stream);
    } else if (nch <= 1) {                              // everything in order on the caller's stream
        if ((rc = stage_a(ctx, 0, 0, batch, d_frames, frame_stride, row_pitch)) != VBS_OK) return rc;
        if ((rc = stage_b(ctx, 0, 0, batch, frameno0, out, kind, ctx->stream)) != VBS_OK) return rc;
    } else {
        if ((rc = ensure_pipeline(ctx)) != VBS_OK) return rc;
        for (int c = 0; c < nch; ++c) {
            const int off = c * chunk, n = batch - off < chunk ? batch - off : chunk;
            if ((rc = stage_a(ctx, c, off, n, d_frames + (size_t)frame_stride * off, frame_stride, row_pitch)) != VBS_OK) return rc;
            VBS_CUDA(cudaEventRecord(ctx->ev_a[c], ctx->stream));
            VBS_CUDA(cudaStreamWaitEvent(ctx->stream_b, ctx->ev_a[c], 0));
            if ((rc = stage_b(ctx, c, off, n, frameno0, out, kind, ctx->stream_b)) != VBS_OK) return rc;
        }
        VBS_CUDA(cudaEventRecord(ctx->ev_b_done, ctx->stream_b));
        VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b_done, 0));      // the caller's stream sees finished results
    }
    if (ctx->profiling) { ctx->prof_pending = 1; ctx->prof_chunks = nch; }
    ctx->last_batch = batch;
    return VBS_OK;
}

int check_batch(vbs_ctx *ctx, const void *frames, int batch) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    if (!frames) return fail(ctx, VBS_ERR_BAD_ARG, "frames is NULL");
    if (batch < 1 || batch > ctx->B) return fail(ctx, VBS_ERR_BAD_ARG, "batch must be in [1, max_batch]");
    return VBS_OK;
}

}  // namespace

extern "C" {

const char *vbs_version(void) { return "vbs_b200 0.1.0 (sm_100a)"; }

int vbs_create(vbs_ctx **out, const vbs_config *cfg) {
    if (!out || !cfg) return VBS_ERR_BAD_ARG;
    *out = nullptr;
    if (cfg->height < 8 || cfg->width < 8 || (cfg->channels != 1 && cfg->channels != 3) || cfg->max_batch < 1 ||
        cfg->max_markers < 1 || cfg->max_markers > 8192 || cfg->max_refs < 0)
        return VBS_ERR_BAD_ARG;
    vbs_ctx *ctx = new (std::nothrow) vbs_ctx();
    if (!ctx) return VBS_ERR_CUDA;
    ctx->cfg = *cfg;
    ctx->H = cfg->height; ctx->W = cfg->width; ctx->C = cfg->channels;
    ctx->WW = ((cfg->width + 31) / 32 + 3) / 4 * 4;      // bit-image rows padded to whole 128-bit quads
    ctx->B = cfg->max_batch; ctx->M = cfg->max_markers; ctx->Rcap = cfg->max_refs;
    ctx->big = cfg->height > 480;                                             // MD:117
    ctx->br = ctx->big ? VbsBranch{39, 101, 80, 13.0, 20, 200, 14} : VbsBranch{21, 35, 33, 7.4, 35, 180, 8};
    ctx->min_dist = 20.0;
    { const char *e = getenv("VBS_NO_TMA"); ctx->no_tma = (e && e[0] == '1') ? 1 : 0; }
    { const char *e = getenv("VBS_NCC_VARIANT"); ctx->ncc_variant = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }
    { const char *e = getenv("VBS_BLUR_VARIANT"); ctx->blur_variant = (e && e[0] == '0') ? 0 : 1; }
    { const char *e = getenv("VBS_SEG_PLAN"); ctx->seg_plan = (e && e[0] == '0') ? 0 : 1; }
    ctx->sm_count = 148;
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && v > 0) ctx->sm_count = v; }
    { const char *e = getenv("VBS_BLUR_TC"); ctx->blur_tc = (e && e[0] == '1') ? 1 : 0; }      // opt-in tensor-core blur
    // The open-mask branch (5x5 open, blobs, border following, ellipse fits) needs K1 only and runs on a second,
    // high-priority stream beside the NCC: the NCC stretches from 3.0 to 3.7 ms but the 0.9 ms of small latency-bound
    // kernels disappear behind it (8.41 -> 8.26 ms per 256 frames on B200).  VBS_BRANCH_OVERLAP=0 runs them in sequence.
    { const char *e = getenv("VBS_BRANCH_OVERLAP"); ctx->no_branch_overlap = (e && e[0] == '0') ? 1 : 0; }
    ctx->first_frame = 0; ctx->have_first = 0;
    *out = ctx;                                    // returned even on failure so vbs_last_error works; caller destroys
    if (vbs_check_taps(ctx->err) != 0) return VBS_ERR_INTERNAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ctx, VBS_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    VBS_ON_DEVICE(ctx);                              // the caller's current device is restored on return
    VBS_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    const size_t B = ctx->B, H = ctx->H, W = ctx->W, M = ctx->M, R = ctx->Rcap;
    VBS_CUDA(dalloc(&ctx->thr_lut, (size_t)ctx->br.tl * ctx->br.tl + 1));
    VBS_CUDA(dalloc(&ctx->d_n64, 96)); VBS_CUDA(dalloc(&ctx->d_cn64, 160)); VBS_CUDA(dalloc(&ctx->d_cnfix, 4 * 112));
    ctx->recheck_cap = 16384;
    ctx->rcap = (int)(8 * M + 64 * ((W + 1023) / 1024) * ((H + 63) / 64));      // tile-local components: 64 x 1024 px tiles (k_ccl.cu)
    // per-pixel scratch and the cell grids come later (ensure_image / ensure_track): table-only contexts skip them
    VBS_CUDA(dalloc(&ctx->d_nlabels, B)); VBS_CUDA(dalloc(&ctx->d_ncont, B));
    VBS_CUDA(dalloc(&ctx->d_nmarkers, B)); VBS_CUDA(dalloc(&ctx->marker_xy, B * M * 2)); VBS_CUDA(dalloc(&ctx->marker_axes, B * M * 3));
    VBS_CUDA(dalloc(&ctx->ref_row, R)); VBS_CUDA(dalloc(&ctx->ref_col, R)); VBS_CUDA(dalloc(&ctx->ref_xy, R * 2));
    VBS_CUDA(dalloc(&ctx->row_det, B * R)); VBS_CUDA(dalloc(&ctx->row_cxy, B * R * 2)); VBS_CUDA(dalloc(&ctx->row_axes, B * R * 3));
    VBS_CUDA(dalloc(&ctx->obs, B * R * 3)); VBS_CUDA(dalloc(&ctx->pos3d, B * R * 7)); VBS_CUDA(dalloc(&ctx->pos_flags, B * R));
    VBS_CUDA(dalloc(&ctx->last_seen, R * 4));
    VBS_CUDA(dalloc(&ctx->pl_ref, R * 3)); VBS_CUDA(dalloc(&ctx->pl_start, R * 3)); VBS_CUDA(dalloc(&ctx->pl_dvert, R * 3));
    VBS_CUDA(dalloc(&ctx->pl_use, R)); VBS_CUDA(dalloc(&ctx->plane, B * 4)); VBS_CUDA(dalloc(&ctx->plane_n, B));
    VBS_CUDA(dalloc(&ctx->d_status, 4));             // [0]: synchronous calls, [1], [2]: the two batches vbs_submit_host keeps in flight
    VBS_CUDA(cudaHostAlloc((void **)&ctx->h_status, sizeof(uint32_t), cudaHostAllocDefault));
    *ctx->h_status = 0;
    VBS_CUDA(vbs_ncc_setup(ctx));
    return vbs_reset_sequence(ctx);
}

void vbs_destroy(vbs_ctx *ctx) {
    if (!ctx) return;
    {
        DeviceGuard guard(ctx->cfg.device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        free_all(ctx);
    }
    delete ctx;
}

const char *vbs_last_error(const vbs_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vbs_set_stream(vbs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return VBS_OK;
}

int vbs_sync(vbs_ctx *ctx) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    VBS_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream));
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t st = *ctx->h_status;
    *ctx->h_status = 0;
    return map_status(ctx, st);
}

int vbs_set_reference(vbs_ctx *ctx, int32_t n, const int32_t *row, const int32_t *col, const double *ox, const double *oy,
                      double min_marker_distance) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    if (n < 0 || n > ctx->Rcap) return fail(ctx, VBS_ERR_BAD_ARG, "reference count exceeds max_refs");
    if (n > 0 && (!row || !col || !ox || !oy)) return fail(ctx, VBS_ERR_BAD_ARG, "NULL reference array");
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    double *xy = new double[2 * (size_t)(n ? n : 1)];
    for (int i = 0; i < n; ++i) { xy[2 * i] = ox[i]; xy[2 * i + 1] = oy[i]; }
    cudaError_t e = cudaMemcpy(ctx->ref_xy, xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice);
    delete[] xy;
    VBS_CUDA(e);
    VBS_CUDA(cudaMemcpy(ctx->ref_row, row, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    VBS_CUDA(cudaMemcpy(ctx->ref_col, col, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    ctx->R = n; ctx->min_dist = min_marker_distance;
    ctx->have_plane = 0;
    return vbs_reset_sequence(ctx);
}

int vbs_set_camera(vbs_ctx *ctx, const float K[9], const float D[5], const float R[9], const float T[3], double marker_diameter_mm,
                   double min_marker_size_px, double max_displacement, int32_t warmup_frames) {
    if (!ctx || !K || !D || !R || !T) return VBS_ERR_BAD_ARG;
    if (!(K[0] > 0) || !(K[4] > 0)) return fail(ctx, VBS_ERR_BAD_ARG, "Focal lengths must be positive");   // R3:94-95
    vbs::CameraF64 c;
    c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
    c.k1 = D[0]; c.k2 = D[1]; c.p1 = D[2]; c.p2 = D[3]; c.k3 = D[4];
    for (int i = 0; i < 9; ++i) c.R[i] = R[i];
    for (int i = 0; i < 3; ++i) c.T[i] = T[i];
    // NumPy-2 scalar promotion at R3:211,219: float32 + float32, / python int -> float32;
    // python float / float32 -> float32
    volatile float favg = (K[0] + K[4]) / 2.0f;
    volatile float ratio = (float)marker_diameter_mm / favg;
    volatile float favg_sq = favg * favg;                    // np.float32 ** 2 -> float32 (R3:219)
    c.f_avg = favg; c.ratio = ratio; c.f_avg_sq = favg_sq;
    c.min_size = min_marker_size_px; c.max_disp = max_displacement;
    ctx->cam = c; ctx->have_cam = 1; ctx->warmup = warmup_frames;
    return vbs_reset_sequence(ctx);
}

int vbs_set_plane(vbs_ctx *ctx, int32_t n, const double *ref_xyz, const double *start_xyz, const double *d_vert, const uint8_t *use,
                  int32_t shell_mode, double scale) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    if (n != ctx->R || n <= 0) return fail(ctx, VBS_ERR_BAD_ARG, "plane arrays must match the reference array length");
    if (!ref_xyz || !start_xyz) return fail(ctx, VBS_ERR_BAD_ARG, "NULL plane array");
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    VBS_CUDA(cudaMemcpy(ctx->pl_ref, ref_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    VBS_CUDA(cudaMemcpy(ctx->pl_start, start_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    if (d_vert) VBS_CUDA(cudaMemcpy(ctx->pl_dvert, d_vert, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    else VBS_CUDA(cudaMemset(ctx->pl_dvert, 0, sizeof(double) * 3 * n));
    if (use) VBS_CUDA(cudaMemcpy(ctx->pl_use, use, n, cudaMemcpyHostToDevice));
    else VBS_CUDA(cudaMemset(ctx->pl_use, 1, n));
    ctx->shell = shell_mode ? 1 : 0; ctx->pscale = scale; ctx->have_plane = 1;
    return VBS_OK;
}

int vbs_reset_sequence(vbs_ctx *ctx) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    const int R = ctx->Rcap;
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (R > 0) {
        double *t = new double[4 * (size_t)R];
        for (int i = 0; i < R; ++i) { t[4 * i] = t[4 * i + 1] = t[4 * i + 2] = 0.0; t[4 * i + 3] = -1.0; }
        cudaError_t e = cudaMemcpy(ctx->last_seen, t, sizeof(double) * 4 * R, cudaMemcpyHostToDevice);
        delete[] t;
        VBS_CUDA(e);
    }
    ctx->have_first = 0; ctx->first_frame = 0;
    return VBS_OK;
}

int vbs_get_last_seen(vbs_ctx *ctx, double *host_table) {
    if (!ctx || !host_table) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    VBS_CUDA(cudaMemcpy(host_table, ctx->last_seen, sizeof(double) * 4 * ctx->R, cudaMemcpyDeviceToHost));
    return VBS_OK;
}

int vbs_set_last_seen(vbs_ctx *ctx, const double *host_table) {
    if (!ctx || !host_table) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    VBS_CUDA(cudaMemcpy(ctx->last_seen, host_table, sizeof(double) * 4 * ctx->R, cudaMemcpyHostToDevice));
    return VBS_OK;
}

int vbs_fix_displacement(vbs_ctx *ctx, double *pos3d_device, uint8_t *pos_flags_device, int64_t nframes, const double *incoming_host) {
    if (!ctx || !pos3d_device || !pos_flags_device || !incoming_host || nframes < 0) return VBS_ERR_BAD_ARG;
    if (ctx->R <= 0 || !ctx->have_cam) return fail(ctx, VBS_ERR_STATE, "reference array and camera must be set first");
    VBS_ON_DEVICE(ctx);
    Scratch s;
    VBS_CUDA(cudaMalloc(&s.p, sizeof(double) * 4 * ctx->R));
    VBS_CUDA(cudaMemcpyAsync(s.p, incoming_host, sizeof(double) * 4 * ctx->R, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(vbs_launch_fix_displacement(ctx, pos3d_device, pos_flags_device, (const double *)s.p, nframes));
    return vbs_sync(ctx);
}

int vbs_set_first_frame(vbs_ctx *ctx, int64_t first_frame) {       // frame-sharded runs: warm-up counts from the global first frame
    if (!ctx) return VBS_ERR_BAD_ARG;
    ctx->first_frame = first_frame; ctx->have_first = 1;
    return VBS_OK;
}

int vbs_process_device(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch, int64_t frameno0,
                       const vbs_outputs *out) {
    int rc = check_batch(ctx, frames, batch);
    if (rc != VBS_OK) return rc;
    if (row_pitch < (int64_t)ctx->W * ctx->C) return fail(ctx, VBS_ERR_BAD_ARG, "row_pitch smaller than a row");
    VBS_ON_DEVICE(ctx);
    if ((rc = ensure_image(ctx)) != VBS_OK || (rc = ensure_blur_tc(ctx)) != VBS_OK) return rc;
    return process_common(ctx, frames, batch, frame_stride, row_pitch, frameno0, out, cudaMemcpyDefault);
}

}  // extern "C"

namespace {

// ---- host entry points -----------------------------------------------------------------------------------
// Frames live in host memory (pinned for full speed); outputs may be host OR device pointers (the copies are
// cudaMemcpyDefault, the driver tells them apart).  Two ways to overlap the PCIe copy with the kernels:
//   chunked : the batch is cut into chunks; chunk c+1 crosses PCIe on a copy stream while chunk c is processed
//             (two rotating staging buffers of one chunk each, events both ways).  The rotation continues
//             ACROSS calls, so with vbs_submit_host the first chunk of batch i+1 flies beside the tail of batch i.
//   whole   : two staging slots of max_batch frames; batch i+1 is copied in one piece beside the unchunked
//             pipeline of batch i (fewer, longer kernels: faster on the device, but every rank of a multi-GPU
//             job then pulls max_batch frames through the host at once).
// vbs_process_host is always chunked; vbs_submit_host is whole-batch unless vbs_set_host_chunk(> 0) was called.
int ensure_host_streams(vbs_ctx *ctx) {
    if (!ctx->copy_stream) {
        VBS_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < VBS_MAX_CHUNKS; ++i) VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_bchunk[i], cudaEventDisableTiming));
    }
    if (!ctx->h_slot_status) {
        for (int i = 0; i < 2; ++i) {
            VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_slot_in[i], cudaEventDisableTiming));
            VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_slot_free[i], cudaEventDisableTiming));
            VBS_CUDA(cudaEventCreateWithFlags(&ctx->ev_slot_done[i], cudaEventDisableTiming));
        }
        VBS_CUDA(cudaHostAlloc((void **)&ctx->h_slot_status, 2 * sizeof(uint32_t), cudaHostAllocDefault));
        ctx->h_slot_status[0] = ctx->h_slot_status[1] = 0;
    }
    return ensure_pipeline(ctx);
}

// frames per chunk of a chunked host call: the caller's setting (or 64), raised so that the batch needs at most
// VBS_MAX_CHUNKS chunks (any batch <= max_batch is accepted)
int host_chunk_frames(const vbs_ctx *ctx, int batch, int setting) {
    int ch = setting > 0 ? setting : 64;
    const int floor_ch = (batch + VBS_MAX_CHUNKS - 1) / VBS_MAX_CHUNKS;
    if (ch < floor_ch) ch = floor_ch;
    if (ch > ctx->B) ch = ctx->B;
    return ch;
}

int upload_frames(vbs_ctx *ctx, uint8_t *dst, const uint8_t *src, int n, int64_t frame_stride, int64_t row_pitch) {
    const size_t rowb = (size_t)ctx->W * ctx->C, fb = rowb * ctx->H;
    if ((size_t)row_pitch == rowb && (size_t)frame_stride == fb) {
        VBS_CUDA(cudaMemcpyAsync(dst, src, fb * n, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {                                           // crop view (MD:85): one strided copy per frame
        for (int f = 0; f < n; ++f)
            VBS_CUDA(cudaMemcpy2DAsync(dst + fb * f, rowb, src + (size_t)frame_stride * f, (size_t)row_pitch, rowb, ctx->H,
                                       cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    return VBS_OK;
}

// Enqueue one batch on the chunked pipeline.  On return the last stage-B work (and the output copies) of the
// batch sit on ctx->stream_b; nothing has been waited for.
int enqueue_host_chunks(vbs_ctx *ctx, const uint8_t *frames, int batch, int64_t frame_stride, int64_t row_pitch, int64_t frameno0,
                        const vbs_outputs *out, int CH) {
    int rc;
    const size_t fb = (size_t)ctx->W * ctx->C * ctx->H;
    if (!ctx->d_frames || ctx->frames_bytes < 2 * fb * CH) {
        if (ctx->d_frames) { VBS_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->d_frames); ctx->d_frames = nullptr; }
        ctx->frames_bytes = 2 * fb * CH;
        ctx->chunk_seq = 0;                            // fresh buffers: no previous consumer to wait for
        VBS_CUDA(cudaMalloc((void **)&ctx->d_frames, ctx->frames_bytes));
    }
    const size_t half = ctx->frames_bytes / 2;
    const int nchunks = (batch + CH - 1) / CH;
    if (ctx->profiling && (rc = prof_collect(ctx)) != VBS_OK) return rc;
    if (CH != ctx->last_chunk_frames) {
        // another chunk size than the batch before: chunk c no longer covers the scratch range chunk c covered, so
        // the per-index events below do not order the reuse - wait for every chunk that may still be in stage B
        for (int c = 0; c < VBS_MAX_CHUNKS; ++c)
            if (ctx->bchunk_live[c]) { VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_bchunk[c], 0)); ctx->bchunk_live[c] = 0; }
        ctx->last_chunk_frames = CH;
    }
    // the copy stream runs one chunk ahead of stage A; stage B of the previous chunk runs beside stage A
    auto upload = [&](int c) -> int {
        const int off = c * CH, n = batch - off < CH ? batch - off : CH;
        const int buf = (int)((ctx->chunk_seq + c) & 1);
        if (ctx->chunk_seq + c >= 2) VBS_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[buf], 0));
        int r = upload_frames(ctx, ctx->d_frames + (size_t)buf * half, frames + (size_t)frame_stride * off, n, frame_stride, row_pitch);
        if (r != VBS_OK) return r;
        VBS_CUDA(cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream));
        return VBS_OK;
    };
    if ((rc = upload(0)) != VBS_OK) return rc;
    for (int c = 0; c < nchunks; ++c) {
        const int off = c * CH, n = batch - off < CH ? batch - off : CH;
        const int buf = (int)((ctx->chunk_seq + c) & 1);
        uint8_t *dst = ctx->d_frames + (size_t)buf * half;
        if (c + 1 < nchunks && (rc = upload(c + 1)) != VBS_OK) return rc;
        VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[buf], 0));
        // the scratch range of chunk c was last read by stage B of chunk c of the previous batch (stream_b)
        if (ctx->bchunk_live[c]) VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_bchunk[c], 0));
        if ((rc = stage_a(ctx, c, off, n, dst, (int64_t)fb, (int64_t)ctx->W * ctx->C)) != VBS_OK) return rc;
        VBS_CUDA(cudaEventRecord(ctx->ev_consumed[buf], ctx->stream));          // frames are dead after the blur
        VBS_CUDA(cudaEventRecord(ctx->ev_a[c], ctx->stream));
        VBS_CUDA(cudaStreamWaitEvent(ctx->stream_b, ctx->ev_a[c], 0));
        if ((rc = stage_b(ctx, c, off, n, frameno0, out, cudaMemcpyDefault, ctx->stream_b)) != VBS_OK) return rc;
        VBS_CUDA(cudaEventRecord(ctx->ev_bchunk[c], ctx->stream_b));
        ctx->bchunk_live[c] = 1;
    }
    ctx->chunk_seq += nchunks;
    if (ctx->profiling) { ctx->prof_pending = 1; ctx->prof_chunks = nchunks; }
    ctx->last_batch = batch;
    return VBS_OK;
}

}  // namespace

extern "C" {

int vbs_process_host(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch, int64_t frameno0,
                     const vbs_outputs *out) {
    int rc = check_batch(ctx, frames, batch);
    if (rc != VBS_OK) return rc;
    if (row_pitch < (int64_t)ctx->W * ctx->C) return fail(ctx, VBS_ERR_BAD_ARG, "row_pitch smaller than a row");
    if (ctx->inflight > 0) return fail(ctx, VBS_ERR_STATE, "batches submitted with vbs_submit_host are still in flight: call vbs_wait_host first");
    VBS_ON_DEVICE(ctx);
    if ((rc = ensure_image(ctx)) != VBS_OK || (rc = ensure_blur_tc(ctx)) != VBS_OK) return rc;
    if ((rc = ensure_host_streams(ctx)) != VBS_OK) return rc;
    if ((rc = enqueue_host_chunks(ctx, frames, batch, frame_stride, row_pitch, frameno0, out, host_chunk_frames(ctx, batch, ctx->host_chunk))) != VBS_OK)
        return rc;
    VBS_CUDA(cudaEventRecord(ctx->ev_b_done, ctx->stream_b));
    VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b_done, 0));
    return vbs_sync(ctx);
}

// ---- asynchronous host entry point: batch i+1 crosses PCIe while batch i is being processed ------------
// Up to two batches in flight.  submit: enqueue the copies, the pipeline and the output copies, then a per-slot
// "done" event (device-side status travels in a per-slot word, so a batch never reports its neighbour's flags).
// wait: blocks on the oldest batch in flight.
int vbs_submit_host(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch, int64_t frameno0,
                    const vbs_outputs *out) {
    int rc = check_batch(ctx, frames, batch);
    if (rc != VBS_OK) return rc;
    const size_t rowb = (size_t)ctx->W * ctx->C, fb = rowb * ctx->H;
    if (row_pitch < (int64_t)rowb) return fail(ctx, VBS_ERR_BAD_ARG, "row_pitch smaller than a row");
    if (ctx->inflight >= 2) return fail(ctx, VBS_ERR_STATE, "two batches already in flight: call vbs_wait_host first");
    VBS_ON_DEVICE(ctx);
    if ((rc = ensure_image(ctx)) != VBS_OK || (rc = ensure_blur_tc(ctx)) != VBS_OK) return rc;
    if ((rc = ensure_host_streams(ctx)) != VBS_OK) return rc;
    const int slot = (int)(ctx->submitted & 1);
    uint32_t *const status_all = ctx->d_status;
    uint32_t *const word = status_all + 1 + slot;
    cudaStream_t tail;                                   // the stream the batch finishes on
    ctx->d_status = word;                                // kernels of this batch flag into the slot's own word
    if (ctx->host_chunk > 0) {
        rc = enqueue_host_chunks(ctx, frames, batch, frame_stride, row_pitch, frameno0, out, host_chunk_frames(ctx, batch, ctx->host_chunk));
        tail = ctx->stream_b;
    } else {
        rc = VBS_OK;
        if (!ctx->d_slots) {
            cudaError_t e = cudaMalloc((void **)&ctx->d_slots, 2 * fb * ctx->B);
            if (e != cudaSuccess) { ctx->d_status = status_all; VBS_CUDA(e); }
        }
        uint8_t *dst = ctx->d_slots + (size_t)slot * fb * ctx->B;
        cudaError_t e = cudaSuccess;
        if (ctx->submitted >= 2 && ctx->slot_used[slot]) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_slot_free[slot], 0);
        if (e == cudaSuccess) rc = upload_frames(ctx, dst, frames, batch, frame_stride, row_pitch);
        if (e == cudaSuccess && rc == VBS_OK) e = cudaEventRecord(ctx->ev_slot_in[slot], ctx->copy_stream);
        if (e == cudaSuccess && rc == VBS_OK) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_slot_in[slot], 0);
        // chunked batches of earlier calls may still be in stage B on stream_b and share the scratch
        for (int c = 0; c < VBS_MAX_CHUNKS && e == cudaSuccess && rc == VBS_OK; ++c)
            if (ctx->bchunk_live[c]) { e = cudaStreamWaitEvent(ctx->stream, ctx->ev_bchunk[c], 0); ctx->bchunk_live[c] = 0; }
        if (e == cudaSuccess && rc == VBS_OK) rc = process_common(ctx, dst, batch, (int64_t)fb, (int64_t)rowb, frameno0, out, cudaMemcpyDefault);
        if (e == cudaSuccess && rc == VBS_OK) e = cudaEventRecord(ctx->ev_slot_free[slot], ctx->stream);   // (frames are dead after the blur; the end of the batch is a safe bound)
        if (e != cudaSuccess) { ctx->d_status = status_all; VBS_CUDA(e); }
        ctx->slot_used[slot] = 1;
        tail = ctx->stream;
    }
    ctx->d_status = status_all;
    if (rc != VBS_OK) return rc;
    VBS_CUDA(cudaMemcpyAsync(&ctx->h_slot_status[slot], word, sizeof(uint32_t), cudaMemcpyDeviceToHost, tail));
    VBS_CUDA(cudaMemsetAsync(word, 0, sizeof(uint32_t), tail));
    VBS_CUDA(cudaEventRecord(ctx->ev_slot_done[slot], tail));
    ctx->submitted += 1;
    ctx->inflight += 1;
    return VBS_OK;
}

int vbs_wait_host(vbs_ctx *ctx) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    if (ctx->inflight <= 0) return fail(ctx, VBS_ERR_STATE, "no batch in flight");
    VBS_ON_DEVICE(ctx);
    const int slot = (int)((ctx->submitted - ctx->inflight) & 1);
    VBS_CUDA(cudaEventSynchronize(ctx->ev_slot_done[slot]));
    ctx->inflight -= 1;
    if (ctx->inflight == 0) {
        // nothing in flight: later calls on the context's stream (vbs_process_device, setters) see finished work
        VBS_CUDA(cudaEventRecord(ctx->ev_b_done, ctx->stream_b));
        VBS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_b_done, 0));
    }
    const uint32_t st = ctx->h_slot_status[slot];
    ctx->h_slot_status[slot] = 0;
    return map_status(ctx, st);
}

int vbs_set_overlap(vbs_ctx *ctx, int32_t enable) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->overlap_device = enable ? 1 : 0;
    return VBS_OK;
}

int vbs_set_host_chunk(vbs_ctx *ctx, int32_t frames_per_chunk) {
    if (!ctx || frames_per_chunk < 0) return VBS_ERR_BAD_ARG;
    if (ctx->inflight > 0) return fail(ctx, VBS_ERR_STATE, "batches are in flight: call vbs_wait_host first");
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->host_chunk = frames_per_chunk;
    return VBS_OK;
}

// MD:93-109.  K == NULL switches the correction off again.
int vbs_set_undistort(vbs_ctx *ctx, const double *K, const double *D, int32_t nd) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    if (!K) { ctx->undist_on = 0; return VBS_OK; }
    if (!D || (nd != 4 && nd != 5 && nd != 8)) return fail(ctx, VBS_ERR_BAD_ARG, "dist_coeffs must hold 4, 5 or 8 values");
    for (int i = 0; i < 9; ++i) if (!std::isfinite(K[i])) return fail(ctx, VBS_ERR_BAD_ARG, "camera_matrix is not finite");
    for (int i = 0; i < nd; ++i) if (!std::isfinite(D[i])) return fail(ctx, VBS_ERR_BAD_ARG, "dist_coeffs are not finite");
    if (K[0] == 0.0 || K[4] == 0.0) return fail(ctx, VBS_ERR_BAD_ARG, "focal lengths must not be zero");
    VBS_ON_DEVICE(ctx);
    const size_t HW = (size_t)ctx->H * ctx->W;
    if (!ctx->undist_map) VBS_CUDA(dalloc(&ctx->undist_map, HW));
    if (!ctx->d_undist) VBS_CUDA(dalloc(&ctx->d_undist, (size_t)ctx->B * HW * ctx->C));
    VBS_CUDA(vbs_undistort_setup(ctx, K, D, nd));
    ctx->undist_on = 1;
    return VBS_OK;
}

// the maps in OpenCV's CV_16SC2 layout (device pointers, either may be NULL) and the new camera matrix (host, 3x3 row-major)
int vbs_get_undistort_maps(vbs_ctx *ctx, double *new_camera_matrix, int16_t *map1_device, uint16_t *map2_device) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    if (!ctx->undist_on) return fail(ctx, VBS_ERR_STATE, "vbs_set_undistort has not been called");
    VBS_ON_DEVICE(ctx);
    if (new_camera_matrix) {
        const double m[9] = {ctx->new_k[0], 0.0, ctx->new_k[2], 0.0, ctx->new_k[1], ctx->new_k[3], 0.0, 0.0, 1.0};
        for (int i = 0; i < 9; ++i) new_camera_matrix[i] = m[i];
    }
    if (map1_device && map2_device) VBS_CUDA(vbs_launch_export_maps(ctx, map1_device, map2_device));
    else if (map1_device || map2_device) return fail(ctx, VBS_ERR_BAD_ARG, "pass both maps or neither");
    return VBS_OK;
}

// stage entry: corrected frames [batch][H][W*C] into out_device (what _preprocess_frame returns, MD:88-91)
int vbs_undistort_frames(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch, uint8_t *out_device) {
    int rc = check_batch(ctx, frames, batch);
    if (rc != VBS_OK) return rc;
    if (!out_device) return fail(ctx, VBS_ERR_BAD_ARG, "out_device is NULL");
    if (!ctx->undist_on) return fail(ctx, VBS_ERR_STATE, "vbs_set_undistort has not been called");
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(vbs_launch_remap(ctx, frames, batch, frame_stride, row_pitch, out_device));
    return VBS_OK;
}

int vbs_find_markers(vbs_ctx *ctx, const uint8_t *frames, int32_t batch, int64_t frame_stride, int64_t row_pitch) {
    int rc = check_batch(ctx, frames, batch);
    if (rc != VBS_OK) return rc;
    VBS_ON_DEVICE(ctx);
    if ((rc = ensure_image(ctx)) != VBS_OK || (rc = ensure_blur_tc(ctx)) != VBS_OK) return rc;
    ctx->last_batch = batch;
    return run_detection(ctx, frames, batch, frame_stride, row_pitch);
}

int vbs_marker_center(vbs_ctx *ctx, const uint8_t *mask, const uint8_t *area_mask, int32_t batch, const vbs_outputs *out) {
    int rc = check_batch(ctx, mask, batch);
    if (rc != VBS_OK) return rc;
    if (!area_mask) return fail(ctx, VBS_ERR_BAD_ARG, "area_mask is NULL");
    VBS_ON_DEVICE(ctx);
    if ((rc = ensure_image(ctx)) != VBS_OK || (rc = ensure_blur_tc(ctx)) != VBS_OK) return rc;
    VBS_CUDA(vbs_launch_pack_masks(ctx, mask, area_mask, batch));
    if ((rc = run_centres(ctx, batch)) != VBS_OK) return rc;
    CopyPlan plan[16];
    vbs_outputs o;
    std::memset(&o, 0, sizeof(o));
    if (out) o = *out;
    o.row_det = nullptr; o.row_cxy = nullptr; o.row_axes = nullptr; o.pos3d = nullptr; o.pos_flags = nullptr; o.plane = nullptr; o.plane_n = nullptr;
    const int n = plan_outputs(ctx, &o, batch, plan);
    for (int i = 0; i < n; ++i) VBS_CUDA(cudaMemcpyAsync(plan[i].dst, plan[i].src, plan[i].bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->last_batch = batch;
    return VBS_OK;
}

// K2 alone on area masks supplied by the caller: mask = normxcorr2(gkern, area_mask) > 0.1 (MD:132-133, 146-164)
int vbs_ncc_mask(vbs_ctx *ctx, const uint8_t *area_mask, int32_t batch) {
    int rc = check_batch(ctx, area_mask, batch);
    if (rc != VBS_OK) return rc;
    VBS_ON_DEVICE(ctx);
    if ((rc = ensure_image(ctx)) != VBS_OK || (rc = ensure_blur_tc(ctx)) != VBS_OK) return rc;
    VBS_CUDA(vbs_launch_pack_area(ctx, area_mask, batch));
    VBS_CUDA(vbs_launch_ncc(ctx, batch));
    ctx->last_batch = batch;
    return VBS_OK;
}

int vbs_debug_stage(vbs_ctx *ctx, int32_t stage, void *dst_device, size_t bytes) {
    if (!ctx || !dst_device) return VBS_ERR_BAD_ARG;
    int batch = ctx->last_batch;
    if (batch <= 0 || !ctx->image_ready) return fail(ctx, VBS_ERR_STATE, "no batch processed yet");
    VBS_ON_DEVICE(ctx);
    // the first frames of the most recent batch, as many as `bytes` holds (at least one)
    const size_t per_frame = (stage == VBS_STAGE_RECHECKS || stage == VBS_STAGE_NCONTOURS) ? sizeof(int32_t)
                             : stage == VBS_STAGE_ELLIPSES ? sizeof(double) * 6 * (size_t)ctx->M
                                                           : (size_t)ctx->H * ctx->W * (stage == VBS_STAGE_LABELS ? 4 : 1);
    if (bytes < per_frame) return fail(ctx, VBS_ERR_BAD_ARG, "destination too small");
    if ((size_t)batch > bytes / per_frame) batch = (int)(bytes / per_frame);
    if (stage == VBS_STAGE_RECHECKS || stage == VBS_STAGE_NCONTOURS || stage == VBS_STAGE_ELLIPSES) {
        const void *src = stage == VBS_STAGE_RECHECKS ? (const void *)ctx->recheck_n : stage == VBS_STAGE_NCONTOURS ? (const void *)ctx->d_ncont : (const void *)ctx->cell;
        VBS_CUDA(cudaMemcpyAsync(dst_device, src, per_frame * batch, cudaMemcpyDeviceToDevice, ctx->stream));
        return VBS_OK;
    }
    VBS_CUDA(vbs_launch_unpack(ctx, stage, dst_device, batch));
    return VBS_OK;
}

// ---- table-level entry points (host arrays in / out; synchronous) --------------------------------
int vbs_track_markers(vbs_ctx *ctx, int32_t n, const double *marker_xy, const double *marker_axes, int32_t *row_det, double *row_cxy,
                      double *row_axes) {
    if (!ctx || n < 0 || (n > 0 && (!marker_xy || !marker_axes))) return VBS_ERR_BAD_ARG;
    if (n > ctx->M) return fail(ctx, VBS_ERR_CAPACITY, "marker list exceeds max_markers");
    if (ctx->R <= 0) return fail(ctx, VBS_ERR_STATE, "no reference array set");
    VBS_ON_DEVICE(ctx);
    { int rc = ensure_track(ctx, 1); if (rc != VBS_OK) return rc; }
    VBS_CUDA(cudaMemcpyAsync(ctx->marker_xy, marker_xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(ctx->marker_axes, marker_axes, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(ctx->d_nmarkers, &n, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    const int save_cam = ctx->have_cam;
    ctx->have_cam = 0;                                  // tracking only: leave the 3D sequence state alone
    cudaError_t e = vbs_launch_track(ctx, 1, 0);
    ctx->have_cam = save_cam;
    VBS_CUDA(e);
    const size_t R = ctx->R;
    if (row_det) VBS_CUDA(cudaMemcpyAsync(row_det, ctx->row_det, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, ctx->stream));
    if (row_cxy) VBS_CUDA(cudaMemcpyAsync(row_cxy, ctx->row_cxy, sizeof(double) * 2 * R, cudaMemcpyDeviceToHost, ctx->stream));
    if (row_axes) VBS_CUDA(cudaMemcpyAsync(row_axes, ctx->row_axes, sizeof(double) * 3 * R, cudaMemcpyDeviceToHost, ctx->stream));
    return vbs_sync(ctx);
}

int vbs_reconstruct_rows(vbs_ctx *ctx, int32_t batch, int64_t frameno0, const int32_t *row_det, const double *row_cxy,
                         const double *row_axes, double *pos3d, uint8_t *pos_flags, double *plane, int32_t *plane_n) {
    if (!ctx || !row_det || !row_cxy || !row_axes) return VBS_ERR_BAD_ARG;
    if (batch < 1 || batch > ctx->B) return fail(ctx, VBS_ERR_BAD_ARG, "batch must be in [1, max_batch]");
    if (ctx->R <= 0 || !ctx->have_cam) return fail(ctx, VBS_ERR_STATE, "reference array and camera must be set first");
    VBS_ON_DEVICE(ctx);
    const size_t n = (size_t)batch * ctx->R;
    VBS_CUDA(cudaMemcpyAsync(ctx->row_det, row_det, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(ctx->row_cxy, row_cxy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(ctx->row_axes, row_axes, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(vbs_launch_reconstruct(ctx, batch, frameno0));
    if (pos3d) VBS_CUDA(cudaMemcpyAsync(pos3d, ctx->pos3d, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (pos_flags) VBS_CUDA(cudaMemcpyAsync(pos_flags, ctx->pos_flags, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->have_plane) {
        if (plane) VBS_CUDA(cudaMemcpyAsync(plane, ctx->plane, sizeof(double) * 4 * batch, cudaMemcpyDeviceToHost, ctx->stream));
        if (plane_n) VBS_CUDA(cudaMemcpyAsync(plane_n, ctx->plane_n, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return vbs_sync(ctx);
}

int vbs_undistort_points(vbs_ctx *ctx, int32_t n, const double *uv, double *out) {
    if (!ctx || n < 0 || (n > 0 && (!uv || !out))) return VBS_ERR_BAD_ARG;
    if (!ctx->have_cam) return fail(ctx, VBS_ERR_STATE, "camera not set");
    if (n == 0) return VBS_OK;
    VBS_ON_DEVICE(ctx);
    Scratch s;
    VBS_CUDA(cudaMalloc(&s.p, sizeof(double) * 4 * n));
    double *d_in = (double *)s.p, *d_out = d_in + 2 * (size_t)n;
    VBS_CUDA(cudaMemcpyAsync(d_in, uv, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(vbs_launch_undistort(ctx, d_in, d_out, n));
    VBS_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    return vbs_sync(ctx);
}

int vbs_position_3d(vbs_ctx *ctx, int32_t n, const double *uvd, double *P, uint8_t *ok) {
    if (!ctx || n < 0 || (n > 0 && (!uvd || !P || !ok))) return VBS_ERR_BAD_ARG;
    if (!ctx->have_cam) return fail(ctx, VBS_ERR_STATE, "camera not set");
    if (n == 0) return VBS_OK;
    VBS_ON_DEVICE(ctx);
    Scratch s;
    VBS_CUDA(cudaMalloc(&s.p, sizeof(double) * 6 * n + n));
    double *d_in = (double *)s.p, *d_P = d_in + 3 * (size_t)n;
    uint8_t *d_ok = (uint8_t *)(d_P + 3 * (size_t)n);
    VBS_CUDA(cudaMemcpyAsync(d_in, uvd, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(vbs_launch_position(ctx, d_in, d_P, d_ok, n));
    VBS_CUDA(cudaMemcpyAsync(P, d_P, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    return vbs_sync(ctx);
}

int vbs_fit_plane(vbs_ctx *ctx, int32_t n, const double *X, const double *Y, const double *Z, double out[4]) {
    if (!ctx || n < 1 || !X || !Y || !Z || !out) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    Scratch s;
    VBS_CUDA(cudaMalloc(&s.p, sizeof(double) * (3 * (size_t)n + 4)));
    double *dX = (double *)s.p, *dY = dX + n, *dZ = dY + n, *dO = dZ + n;
    VBS_CUDA(cudaMemcpyAsync(dX, X, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(dY, Y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(cudaMemcpyAsync(dZ, Z, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    VBS_CUDA(vbs_launch_plane_points(ctx, dX, dY, dZ, n, dO));
    VBS_CUDA(cudaMemcpyAsync(out, dO, sizeof(double) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return vbs_sync(ctx);
}

int64_t vbs_kernel_launches(const vbs_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t vbs_tma_launches(const vbs_ctx *ctx) { return ctx ? ctx->tma_launches : 0; }
int64_t vbs_tc_launches(const vbs_ctx *ctx) { return ctx ? ctx->tc_launches : 0; }

int vbs_set_blur_tc(vbs_ctx *ctx, int32_t enable) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    VBS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->blur_tc = enable ? 1 : 0;
    return VBS_OK;
}

int vbs_set_profiling(vbs_ctx *ctx, int32_t enable) {
    if (!ctx) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    if (enable && !ctx->pev[0])
        for (int i = 0; i < VBS_MAX_CHUNKS * VBS_EV_PER_CHUNK; ++i) VBS_CUDA(cudaEventCreate(&ctx->pev[i]));
    if (!enable) { int rc = prof_collect(ctx); if (rc != VBS_OK) return rc; }
    ctx->profiling = enable ? 1 : 0;
    return VBS_OK;
}

int vbs_get_stage_ms(vbs_ctx *ctx, double ms[7], int64_t *calls) {
    if (!ctx || !ms) return VBS_ERR_BAD_ARG;
    VBS_ON_DEVICE(ctx);
    int rc = prof_collect(ctx);
    if (rc != VBS_OK) return rc;
    for (int i = 0; i < VBS_NSTAGES; ++i) ms[i] = ctx->stage_ms[i];
    if (calls) *calls = ctx->stage_calls;
    return VBS_OK;
}

}  // extern "C"
