// k_ccl.cu - K3b/K4b: connected components on bit images by union-find over WORD SEGMENTS.
//
// A segment is a maximal run of set bits inside one 32-bit word; its id is the pixel index of
// its first bit, so ids live in a dense [H*W] int32 array that is only ever touched at segment
// starts (sparse traffic, no clearing pass).  The foreground is labelled in two levels: per
// 64 x 1024-px tile in shared memory by one warp (fg_strip_kernel), then across tile edges in the
// global array (fg_border_kernel).  Links always point to the smaller index, hence a
// component's root is its first pixel in raster order - exactly what the reference needs:
//   * scipy.ndimage.label numbers 4-connected components by first pixel (MD:176)  -> rank roots
//   * cv2.findContours(RETR_EXTERNAL) starts each outer border at the blob's topmost-leftmost
//     pixel and lists contours in descending start order (MD:196)     -> slot = n-1-rank
// The opened image is labelled twice in the same array: foreground 8-connected and background
// 4-connected (disjoint index sets); background touching the frame is united with a virtual
// "outside" (-1), so a blob is external iff the background left of its start pixel reaches -1.
#include "vbs_ctx.h"

namespace {

constexpr int OUTSIDE = -1;

__device__ __forceinline__ uint32_t valid_mask(int wx, int W) {
    const int rem = W - 32 * wx;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}
template <bool INV>
__device__ __forceinline__ uint32_t get_bits(const uint32_t *img, int y, int wx, int W, int WW) {
    const uint32_t w = __ldg(img + (size_t)y * WW + wx);
    return INV ? (~w & valid_mask(wx, W)) : w;
}
// first bit of the run of set bits that contains set bit b
__device__ __forceinline__ int seg_start(uint32_t w, int b) {
    const uint32_t t = ~w & ((2u << b) - 1u);
    return t ? 32 - __clz(t) : 0;
}
__device__ __forceinline__ uint32_t run_mask(uint32_t w, int s) {      // run of set bits starting at s
    const uint32_t t = ~w >> s;
    const int len = t ? __ffs(t) - 1 : 32 - s;
    return (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
}
__device__ __forceinline__ int find_root(const int32_t *par, int i) {
    while (i >= 0) {
        const int p = par[i];
        if (p == i) break;
        i = p;
    }
    return i;          // root index, or a negative terminal (OUTSIDE / encoded label)
}
// find with path halving (every visited node is re-pointed at its grandparent; links only ever
// move towards smaller indices, so concurrent atomicMin links stay consistent)
__device__ __forceinline__ int find_root_halving(int32_t *par, int i) {
    if (i < 0) return i;
    int p = par[i];
    while (p != i) {
        if (p < 0) return p;
        const int g = par[p];
        if (g == p) return p;
        if (g < 0) return g;
        par[i] = g;
        i = g;
        p = par[i];
    }
    return i;
}
__device__ void unite(int32_t *par, int a, int b) {
    for (;;) {
        a = find_root_halving(par, a);
        b = find_root_halving(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(par + a, b);      // a >= 0 here (a > b >= -1)
        if (old == a) return;
        a = old;
    }
}

// word kernels (background pass, only frames with holes do anything): block (64, 4) = 64 words x 4 rows, looped
// BG_ROWS / 4 times; grid (ceil(WW/64), ceil(H/BG_ROWS), batch)
constexpr int BG_ROWS = 64;
#define VBS_BG_LOOP                                                                  \
    const int wx = blockIdx.x * 64 + threadIdx.x;                                   \
    const int f = blockIdx.z;                                                       \
    if (holes[f] == 0 || wx >= WW) return;                                          \
    for (int y = blockIdx.y * BG_ROWS + threadIdx.y; y < min(H, (int)(blockIdx.y + 1) * BG_ROWS); y += 4)

// ---- background of the opened image (4-connected, frames with holes only) -----------------------------
// 1. every background segment start becomes its own root
__global__ void __launch_bounds__(256) bg_init_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                       const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_BG_LOOP {
        const uint32_t w = ~__ldg(bits + ((size_t)f * H + y) * WW + wx) & valid_mask(wx, W);
        int32_t *par = parent + (size_t)f * H * W;
        const int base = y * W + 32 * wx;
        uint32_t starts = w & ~(w << 1);
        while (starts) { const int s = __ffs(starts) - 1; starts &= starts - 1; par[base + s] = base + s; }
    }
}

// 2. link background segments that touch; segments on the frame are united with OUTSIDE
__global__ void __launch_bounds__(256) bg_merge_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                        const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_BG_LOOP {
        const uint32_t *img = bits + (size_t)f * H * WW;
        const uint32_t w = get_bits<true>(img, y, wx, W, WW);
        if (!w) continue;
        int32_t *par = parent + (size_t)f * H * W;
        const int base = y * W + 32 * wx;
        if ((w & 1u) && wx > 0) {
            const uint32_t pw = get_bits<true>(img, y, wx - 1, W, WW);
            if (pw >> 31) unite(par, base, base - 32 + seg_start(pw, 31));
        }
        const uint32_t up = y > 0 ? get_bits<true>(img, y - 1, wx, W, WW) : 0u;
        const int last_bit = (W - 1) - 32 * wx;         // position of pixel W-1 in this word (may be >= 32)
        uint32_t rem = w;
        while (rem) {
            const int s = __ffs(rem) - 1;
            const uint32_t seg = run_mask(w, s);
            rem &= ~seg;
            const int id = base + s;
            uint32_t ov = up & seg;
            while (ov) {
                const int us = seg_start(up, __ffs(ov) - 1);
                ov &= ~run_mask(up, us);
                unite(par, id, base - W + us);
            }
            const bool touches = y == 0 || y == H - 1 || (wx == 0 && (seg & 1u)) ||
                                 (last_bit >= 0 && last_bit < 32 && ((seg >> last_bit) & 1u));
            if (touches) unite(par, id, OUTSIDE);
        }
    }
}

// 3. point every background segment straight at its root (or at OUTSIDE)
__global__ void __launch_bounds__(256) bg_flatten_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                          const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_BG_LOOP {
        const uint32_t bg = ~__ldg(bits + ((size_t)f * H + y) * WW + wx) & valid_mask(wx, W);
        int32_t *par = parent + (size_t)f * H * W;
        const int base = y * W + 32 * wx;
        uint32_t bs = bg & ~(bg << 1);
        while (bs) {
            const int s = __ffs(bs) - 1;
            bs &= bs - 1;
            const int r = find_root(par, base + s);
            if (r != base + s) par[base + s] = r;
        }
    }
}


// ---- foreground, both images in one launch ---------------------------------------------------------
// Two levels.  (1) fg_strip_kernel: ONE WARP marches down a tile of SH rows x 32 words (1024 px), lane =
// word column, the classic row-scan labelling done warp-synchronously: a segment that touches
// segments of the row above inherits their (united) provisional label, a segment that touches
// nothing opens a new label.  Labels, their equivalences and their moments live in shared memory;
// global memory sees one store per segment (parent[segment] = first pixel of its label), one
// store per merged label and one 16-byte record per tile-local component.  (2) fg_border_kernel:
// the links that cross a tile edge are united in the global array.  Everything after that (root
// discovery, moments, ranking) walks the few records instead of the many segments.
// A tile whose label table would overflow is closed early at that row and continued as a new tile;
// rowflag marks such rows so that the border kernel links them like a regular tile edge.  One row
// can open at most 512 labels (1024 px / 2), hence SP = 512 always makes progress.
struct FgImages {
    const uint32_t *bits[2];     // max_bits, open_bits
    int32_t *parent[2];          // parent, parent2
    int img0, nimg;              // images handled by this launch: img0 .. img0 + nimg - 1 (grid.z = batch * nimg)
};
constexpr int SH = 64;           // rows per tile
constexpr int SPX = 1024;        // tile width in pixels (32 words, one per lane)
constexpr int SP = 512;          // provisional labels per tile

// links of one word to its left / upper neighbours in the global array.  Neighbour words that must
// not be linked by this caller are passed as 0.
template <bool CONN8>
__device__ __forceinline__ void merge_word(int32_t *par, uint32_t w, uint32_t prev, uint32_t up, uint32_t upl, uint32_t upr, int base, int stride) {
    if ((w & 1u) && (prev >> 31)) unite(par, base, base - 32 + seg_start(prev, 31));
    if (!(up | upl | upr)) return;
    uint32_t rem = w;
    while (rem) {
        const int s = __ffs(rem) - 1;
        const uint32_t seg = run_mask(w, s);
        rem &= ~seg;
        const int id = base + s;
        uint32_t nb = seg;
        if (CONN8) nb |= (seg << 1) | (seg >> 1);
        uint32_t ov = up & nb;
        while (ov) {
            const int b = __ffs(ov) - 1;
            const int us = seg_start(up, b);
            ov &= ~run_mask(up, us);
            unite(par, id, base - stride + us);
        }
        if (CONN8) {
            if ((seg & 1u) && (upl >> 31)) unite(par, id, base - stride - 32 + seg_start(upl, 31));
            if ((seg >> 31) && (upr & 1u)) unite(par, id, base - stride + 32);
        }
    }
}

// unite two labels, return the surviving (smaller) root
__device__ __forceinline__ int unite_root(int32_t *par, int a, int b) {
    for (;;) {
        a = find_root_halving(par, a);
        b = find_root_halving(par, b);
        if (a == b) return a;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(par + a, b);
        if (old == a) return b;
        a = old;
    }
}

template <bool MOMENTS>
struct StripTables {
    int16_t rl[2][SPX / 2];      // label of every segment of the current / previous row, index = (local x of its start) >> 1
    int32_t lp[SP];              // label equivalences (union-find, root = smallest = earliest label)
    int32_t pg[SP];              // first pixel of the label (global pixel index of the frame)
    uint32_t ac[MOMENTS ? SP : 1], ax[MOMENTS ? SP : 1], ay[MOMENTS ? SP : 1];   // ring image only: pixel count, sum of local x, sum of local y
};

// end of a (sub-)tile: fold merged labels into their roots, emit one record per root
template <bool MOMENTS>
__device__ __forceinline__ void strip_close(StripTables<MOMENTS> &T, int next, int lane, int32_t *par, int x0, int y0, int z,
                                            int32_t *nrec, int4 *recs, int RCAP, uint32_t *status, uint32_t overflow_bit) {
    __syncwarp();
    for (int L = lane; L < next; L += 32) {
        const int r = find_root(T.lp, L);
        if (r != L) {
            par[T.pg[L]] = T.pg[r];
            if (MOMENTS) { atomicAdd(&T.ac[r], T.ac[L]); atomicAdd(&T.ax[r], T.ax[L]); atomicAdd(&T.ay[r], T.ay[L]); }
        }
    }
    __syncwarp();
    int total = 0;
    for (int L0 = 0; L0 < next; L0 += 32) {
        const int L = L0 + lane;
        total += __popc(__ballot_sync(0xffffffffu, L < next && T.lp[L] == L));
    }
    int off = 0;
    if (lane == 0 && total) off = atomicAdd(nrec + z, total);
    off = __shfl_sync(0xffffffffu, off, 0);
    for (int L0 = 0; L0 < next; L0 += 32) {
        const int L = L0 + lane;
        const bool root = L < next && T.lp[L] == L;
        const uint32_t m = __ballot_sync(0xffffffffu, root);
        if (root) {
            const int slot = off + __popc(m & ((1u << lane) - 1u));
            const uint32_t c = MOMENTS ? T.ac[L] : 0u;
            if (slot < RCAP) recs[(size_t)z * RCAP + slot] = make_int4(T.pg[L], (int)c, MOMENTS ? (int)(T.ax[L] + c * (uint32_t)x0) : 0,
                                                                       MOMENTS ? (int)(T.ay[L] + c * (uint32_t)y0) : 0);
            else atomicOr(status, overflow_bit);
        }
        off += __popc(m);
    }
    __syncwarp();
}

// CONN8 = opened area mask (8-connected, no moments); !CONN8 = ring maxima (4-connected, moments)
// block 32 (one warp); grid (ceil(WW / 32), ceil(H / SH), batch)
template <bool CONN8>
__global__ void __launch_bounds__(32) fg_strip_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent, int32_t *__restrict__ nrec,
                                                       int4 *__restrict__ recs, uint8_t *__restrict__ rowflag, int H, int W, int WW, int RCAP, uint32_t *status) {
    constexpr bool MOMENTS = !CONN8;
    constexpr uint32_t overflow_bit = CONN8 ? VBS_DEV_CONTOUR_OVERFLOW : VBS_DEV_LABEL_OVERFLOW;
    __shared__ StripTables<MOMENTS> T;
    const int f = blockIdx.z, z = 2 * f + (CONN8 ? 1 : 0);
    int32_t *par = parent + (size_t)f * H * W;
    const int lane = threadIdx.x;
    const int band = blockIdx.x, nbands = gridDim.x;
    const int x0 = band * SPX, wx = band * 32 + lane, lx0 = 32 * lane;
    const int y0 = blockIdx.y * SH, rows = min(SH, H - y0);
    const uint32_t *src = bits + ((size_t)f * H + y0) * WW + wx;
    const bool inw = wx < WW;
    int next = 0;
    unsigned long long closed_rows = 0ull;
    uint32_t up = 0u;
    uint32_t wnext = inw ? __ldg(src) : 0u;                  // rows are fetched two steps ahead
    uint32_t wnext2 = (inw && rows > 1) ? __ldg(src + WW) : 0u;
    int gbase = y0 * W + x0 + lx0;                           // pixel index of this lane's bit 0 in the current row
    for (int ly = 0; ly < rows; ++ly, gbase += W) {
        const uint32_t w = wnext;
        wnext = wnext2;
        wnext2 = (inw && ly + 2 < rows) ? __ldg(src + (size_t)(ly + 2) * WW) : 0u;
        if (!__any_sync(0xffffffffu, w != 0u)) { up = 0u; continue; }
        uint32_t upl = 0u, upr = 0u;
        if (CONN8) {
            upl = __shfl_up_sync(0xffffffffu, up, 1); upr = __shfl_down_sync(0xffffffffu, up, 1);
            if (lane == 0) upl = 0u;
            if (lane == 31) upr = 0u;
        }
        uint32_t prevw = __shfl_up_sync(0xffffffffu, w, 1);
        if (lane == 0) prevw = 0u;
        int16_t *cur = T.rl[ly & 1];
        const int16_t *prv = T.rl[(ly & 1) ^ 1];
        const uint32_t starts = w & ~(w << 1);
        uint32_t touch;
        int my_base;
        for (;;) {
            // bits of this word that are adjacent to a pixel of the row above
            touch = w & (CONN8 ? (up | (up << 1) | (up >> 1) | (upl >> 31) | ((upr & 1u) << 31)) : up);
            // a carry started at the first bit of a run climbs out of the run iff no touch bit stops it
            const int fresh = __popcll(((unsigned long long)(w & ~touch) + starts) & ~(unsigned long long)w);
            my_base = next;
            if (!__any_sync(0xffffffffu, fresh != 0)) break;
            int incl = fresh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            my_base = next + incl - fresh;
            if (next + total <= SP) { next += total; break; }
            // label table full: close the tile above this row and start a new one here
            strip_close<MOMENTS>(T, next, lane, par, x0, y0, z, nrec, recs, RCAP, status, overflow_bit);
            next = 0; up = 0u; upl = 0u; upr = 0u; closed_rows |= 1ull << ly;
        }
        for (uint32_t rem = w; rem;) {
            const int s = __ffs(rem) - 1;
            const uint32_t seg = run_mask(w, s);
            rem &= ~seg;
            const int gid = gbase + s;
            uint32_t len = 0, xs = 0, ys = 0;
            if (MOMENTS) { len = __popc(seg); xs = len * (uint32_t)(lx0 + s) + len * (len - 1) / 2; ys = len * (uint32_t)ly; }
            int L;
            if (!(seg & touch)) {
                L = my_base++;
                T.lp[L] = L; T.pg[L] = gid;
                if (MOMENTS) { T.ac[L] = len; T.ax[L] = xs; T.ay[L] = ys; }
                par[gid] = gid;
            } else {
                L = -1;
                const uint32_t nb = CONN8 ? (seg | (seg << 1) | (seg >> 1)) : seg;
                uint32_t ov = up & nb;
                while (ov) {
                    const int us = seg_start(up, __ffs(ov) - 1);
                    ov &= ~run_mask(up, us);
                    const int Lu = prv[(lx0 + us) >> 1];
                    L = L < 0 ? find_root_halving(T.lp, Lu) : unite_root(T.lp, L, Lu);
                }
                if (CONN8) {
                    if ((seg & 1u) && (upl >> 31)) {
                        const int Lu = prv[(lx0 - 32 + seg_start(upl, 31)) >> 1];
                        L = L < 0 ? find_root_halving(T.lp, Lu) : unite_root(T.lp, L, Lu);
                    }
                    if ((seg >> 31) && (upr & 1u)) {
                        const int Lu = prv[(lx0 + 32) >> 1];
                        L = L < 0 ? find_root_halving(T.lp, Lu) : unite_root(T.lp, L, Lu);
                    }
                }
                if (MOMENTS) { atomicAdd(&T.ac[L], len); atomicAdd(&T.ax[L], xs); atomicAdd(&T.ay[L], ys); }
                par[gid] = T.pg[L];
            }
            cur[(lx0 + s) >> 1] = (int16_t)L;
        }
        __syncwarp();
        if ((w & 1u) && (prevw >> 31)) {
            const int a = cur[lx0 >> 1], b = cur[(lx0 - 32 + seg_start(prevw, 31)) >> 1];
            if (a != b) unite(T.lp, a, b);       // usually both already hold the same root
        }
        __syncwarp();
        up = w;
    }
    strip_close<MOMENTS>(T, next, lane, par, x0, y0, z, nrec, recs, RCAP, status, overflow_bit);
    uint8_t *flag = rowflag + ((size_t)z * nbands + band) * H + y0;
    for (int r = lane; r < rows; r += 32) flag[r] = (uint8_t)((closed_rows >> r) & 1ull);
}

// links across tile edges, in the global array.  One thread per candidate; grid (ceil(n / 256), nimg * batch) with
// n = nA + nB + nC:  A = words of the fixed tile-top rows (y = SH, 2 SH, ...), B = the (row, band edge) pairs, all
// links across a vertical tile edge, C = (row, band) pairs, the rows that fg_strip_kernel flagged as early tile tops.
// A and C link across the horizontal edge inside one band only (up, and the diagonals that stay in the band).
template <bool CONN8>
__device__ __forceinline__ void link_tile_top(int32_t *par, const uint32_t *img_bits, int y, int wx, int W, int WW) {
    const uint32_t w = __ldg(img_bits + (size_t)y * WW + wx);
    if (!w) return;
    const uint32_t *urow = img_bits + (size_t)(y - 1) * WW;
    const uint32_t up = __ldg(urow + wx);
    const uint32_t upl = (CONN8 && wx % 32 != 0) ? __ldg(urow + wx - 1) : 0u;
    const uint32_t upr = (CONN8 && wx % 32 != 31 && wx + 1 < WW) ? __ldg(urow + wx + 1) : 0u;
    merge_word<CONN8>(par, w, 0u, up, upl, upr, y * W + 32 * wx, W);
}

__global__ void __launch_bounds__(256) fg_border_kernel(FgImages im, const uint8_t *__restrict__ rowflag, int H, int W, int WW, int nbands) {
    const int f = blockIdx.y / im.nimg, img = im.img0 + blockIdx.y % im.nimg, z = 2 * f + img;
    const uint32_t *img_bits = im.bits[img] + (size_t)f * H * WW;
    int32_t *par = im.parent[img] + (size_t)f * H * W;
    const int nA = ((H - 1) / SH) * WW, nB = H * (nbands - 1), nC = H * nbands;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nA) {
        const int y = (t / WW + 1) * SH, wx = t % WW;
        if (img) link_tile_top<true>(par, img_bits, y, wx, W, WW);
        else link_tile_top<false>(par, img_bits, y, wx, W, WW);
        return;
    }
    t -= nA;
    if (t < nB) {
        const int y = t % H, wx = 32 * (1 + t / H);
        const uint32_t w = __ldg(img_bits + (size_t)y * WW + wx), prev = __ldg(img_bits + (size_t)y * WW + wx - 1);
        const int base = y * W + 32 * wx;
        if ((w & 1u) && (prev >> 31)) unite(par, base, base - 32 + seg_start(prev, 31));
        if (img && y > 0) {
            const uint32_t upl = __ldg(img_bits + (size_t)(y - 1) * WW + wx - 1), upw = __ldg(img_bits + (size_t)(y - 1) * WW + wx);
            if ((w & 1u) && (upl >> 31)) unite(par, base, base - W - 32 + seg_start(upl, 31));
            if ((prev >> 31) && (upw & 1u)) unite(par, base - 32 + seg_start(prev, 31), base - W);
        }
        return;
    }
    t -= nB;
    if (t < nC) {
        const int y = t % H, band = t / H;
        if (y % SH == 0 || !rowflag[((size_t)z * nbands + band) * H + y]) return;
        for (int wx = 32 * band; wx < min(32 * band + 32, WW); ++wx) {
            if (img) link_tile_top<true>(par, img_bits, y, wx, W, WW);
            else link_tile_top<false>(par, img_bits, y, wx, W, WW);
        }
    }
}

// root discovery over the records: a record whose pixel is still its own root is a component; it takes
// the next slot of its (frame, image) list and encodes it in place (parent[root] = -2 - slot)
__global__ void __launch_bounds__(256) fg_roots_kernel(FgImages im, const int32_t *__restrict__ nrec, const int4 *__restrict__ recs,
                                                        int32_t *__restrict__ nroots, int32_t *__restrict__ rootlist,
                                                        int H, int W, int M, int RCAP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y / im.nimg, img = im.img0 + blockIdx.y % im.nimg, z = 2 * f + img;
    if (i >= min(nrec[z], RCAP)) return;
    int32_t *par = im.parent[img] + (size_t)f * H * W;
    const int pix = recs[(size_t)z * RCAP + i].x;
    if (par[pix] != pix) return;
    const int slot = atomicAdd(nroots + z, 1);
    if (slot < M) { rootlist[(size_t)z * M + slot] = pix; par[pix] = -2 - slot; }
}

// ring components: integer moments per root slot (MD:181 center_of_mass on a 0/1 mask), one atomic
// triple per record
__global__ void __launch_bounds__(256) moments_kernel(const int32_t *__restrict__ parent, const int32_t *__restrict__ nrec,
                                                       const int4 *__restrict__ recs, uint32_t *__restrict__ cnt,
                                                       unsigned long long *__restrict__ sx, unsigned long long *__restrict__ sy,
                                                       int H, int W, int M, int RCAP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y, z = 2 * f;
    if (i >= min(nrec[z], RCAP)) return;
    const int4 rec = recs[(size_t)z * RCAP + i];
    const int p = find_root(parent + (size_t)f * H * W, rec.x);
    if (p >= 0) return;                         // root beyond capacity (flagged by the sort kernel)
    const int slot = -2 - p;
    if (slot < 0 || slot >= M) return;
    atomicAdd(cnt + (size_t)f * M + slot, (uint32_t)rec.y);
    atomicAdd(sx + (size_t)f * M + slot, (unsigned long long)(uint32_t)rec.z);
    atomicAdd(sy + (size_t)f * M + slot, (unsigned long long)(uint32_t)rec.w);
}

// ---- Euler number of the 8-connected foreground by bit quads ----------------------------------------
// 4 E = #Q1 - #Q3 - 2 #QD over all 2x2 windows of the zero-padded image (Gray's formula).  The
// number of holes of the whole image is (#blobs - E); when it is 0 no blob can lie inside
// another one, so the background labelling that RETR_EXTERNAL would need is skipped for the frame.
constexpr int EU_ROWS = 16;      // quad rows per thread (consecutive, so every image row is loaded once per thread)
__global__ void __launch_bounds__(256) euler_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ euler4, int H, int W, int WW) {
    const int wx = blockIdx.x * 64 + threadIdx.x;            // 0 .. WW (one extra all-zero column)
    const int f = blockIdx.z;
    int v = 0;
    if (wx <= WW) {
        const uint32_t *img = bits + (size_t)f * H * WW;
        auto ld = [&](int y, int w) -> uint32_t { return (y >= 0 && y < H && w >= 0 && w < WW) ? __ldg(img + (size_t)y * WW + w) : 0u; };
        int yy = (int)(blockIdx.y * 4 + threadIdx.y) * EU_ROWS - 1;      // top row of the quad: -1 .. H-1
        uint32_t b = ld(yy, wx), bl = ld(yy, wx - 1);
        for (int i = 0; i < EU_ROWS && yy < H; ++i, ++yy) {
            const uint32_t d = ld(yy + 1, wx), dl = ld(yy + 1, wx - 1);
            if (b | d | ((bl | dl) >> 31)) {
                const uint32_t a = (b << 1) | (bl >> 31), c = (d << 1) | (dl >> 31);
                const uint32_t x1 = a ^ b, x2 = c ^ d, n1 = a & b, n2 = c & d;
                const uint32_t q1 = (x1 & ~x2 & ~n2) | (x2 & ~x1 & ~n1);
                const uint32_t q3 = (x1 & n2) | (x2 & n1);
                const uint32_t qd = x1 & x2 & ~(a ^ d);
                v += __popc(q1) - __popc(q3) - 2 * __popc(qd);
            }
            b = d; bl = dl;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (((threadIdx.y * 64 + threadIdx.x) & 31) == 0 && v) atomicAdd(euler4 + f, v);
}

// ---- rank the roots of one (frame, image) in raster order: bitonic sort in shared memory -----------
// image 0: label = rank -> slot2label, centroids in label order.  image 1: contour slot = n-1-rank -> croot, holes.
__global__ void __launch_bounds__(1024) rank_kernel(const int32_t *__restrict__ nroots, const int32_t *__restrict__ rootlist,
                                                     const int32_t *__restrict__ parent, const uint32_t *__restrict__ cnt,
                                                     const unsigned long long *__restrict__ sx, const unsigned long long *__restrict__ sy,
                                                     const int32_t *__restrict__ euler4, int32_t *__restrict__ slot2label,
                                                     double *__restrict__ centres, int32_t *__restrict__ nlabels, int32_t *__restrict__ croot,
                                                     int32_t *__restrict__ ncont, int32_t *__restrict__ holes, int H, int W, int M, int P2,
                                                     int img0, uint32_t *status) {
    extern __shared__ int32_t key[];
    const int f = blockIdx.x, img = img0 + blockIdx.y, z = 2 * f + img, tid = threadIdx.x;
    const int total = nroots[z];
    const int n = min(total, M);
    if (total > M && tid == 0) atomicOr(status, img == 0 ? VBS_DEV_LABEL_OVERFLOW : VBS_DEV_CONTOUR_OVERFLOW);
    int P = 1;
    while (P < n) P <<= 1;                                  // P <= P2 (shared memory is sized for P2)
    for (int i = tid; i < P; i += blockDim.x) key[i] = i < n ? rootlist[(size_t)z * M + i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const int a = key[i], b = key[l];
                    if (((i & k) == 0) == (a > b)) { key[i] = b; key[l] = a; }
                }
            }
            __syncthreads();
        }
    if (img == 0) {
        const int32_t *par = parent + (size_t)f * H * W;
        for (int i = tid; i < n; i += blockDim.x) {
            const int slot = -2 - par[key[i]];
            slot2label[(size_t)f * M + slot] = i;
            const size_t s = (size_t)f * M + slot;
            const double c = (double)cnt[s];
            centres[((size_t)f * M + i) * 2 + 0] = (double)sy[s] / c;     // row: exact integer sum, one float64 division
            centres[((size_t)f * M + i) * 2 + 1] = (double)sx[s] / c;     // col
        }
        if (tid == 0) nlabels[f] = total;
    } else {
        for (int i = tid; i < n; i += blockDim.x) croot[(size_t)f * M + (n - 1 - i)] = key[i];
        if (tid == 0) { ncont[f] = total; holes[f] = total - euler4[f] / 4; }
    }
}

}  // namespace

// zero the per-batch accumulators once, before the two branches fork
cudaError_t vbs_launch_prepare(vbs_ctx *ctx, int batch) {
    const size_t M = ctx->M;
    cudaStream_t st = ctx->stream;
    cudaError_t e;
    if ((e = cudaMemsetAsync(ctx->lab_cnt, 0, sizeof(uint32_t) * batch * M, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->lab_sx, 0, sizeof(unsigned long long) * batch * M, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->lab_sy, 0, sizeof(unsigned long long) * batch * M, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->euler4, 0, sizeof(int32_t) * batch, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->nroots, 0, sizeof(int32_t) * 2 * batch, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ctx->nrec, 0, sizeof(int32_t) * 2 * batch, st)) != cudaSuccess) return e;
    return cudaMemsetAsync(ctx->claim, 0, sizeof(int32_t) * batch * M, st);
}

// which: bit 0 = ring-maxima image (labels, centroids), bit 1 = opened area mask (blobs, contour order, holes,
// background pass).  The two are independent, so they may run on different streams after vbs_launch_prepare.
cudaError_t vbs_launch_components(vbs_ctx *ctx, int batch, int which) {
    VbsRange range("vbs:components");
    const int H = ctx->H, W = ctx->W, WW = ctx->WW, M = ctx->M;
    const dim3 wb(64, 4);
    const dim3 eg((WW + 1 + 63) / 64, (H + 1 + 4 * EU_ROWS - 1) / (4 * EU_ROWS), batch);
    cudaStream_t st = ctx->stream;
    FgImages im;
    im.bits[0] = ctx->max_bits; im.bits[1] = ctx->open_bits; im.parent[0] = ctx->parent; im.parent[1] = ctx->parent2;
    im.img0 = (which & 1) ? 0 : 1;
    im.nimg = (which == 3) ? 2 : 1;
    const int RCAP = ctx->rcap;
    const int nbands = (WW + 31) / 32;
    const dim3 sg(nbands, (H + SH - 1) / SH, batch);
    if (which & 2) { fg_strip_kernel<true><<<sg, 32, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->nrec, ctx->recs, ctx->rowflag, H, W, WW, RCAP, ctx->d_status); ctx->launches += 1; }
    if (which & 1) { fg_strip_kernel<false><<<sg, 32, 0, st>>>(ctx->max_bits, ctx->parent, ctx->nrec, ctx->recs, ctx->rowflag, H, W, WW, RCAP, ctx->d_status); ctx->launches += 1; }
    const int nborder = ((H - 1) / SH) * WW + H * (nbands - 1) + H * nbands;
    fg_border_kernel<<<dim3((nborder + 255) / 256, im.nimg * batch), 256, 0, st>>>(im, ctx->rowflag, H, W, WW, nbands);
    fg_roots_kernel<<<dim3((RCAP + 255) / 256, im.nimg * batch), 256, 0, st>>>(im, ctx->nrec, ctx->recs, ctx->nroots, ctx->rootlist, H, W, M, RCAP);
    ctx->launches += 2;
    if (which & 1) { moments_kernel<<<dim3((RCAP + 255) / 256, batch), 256, 0, st>>>(ctx->parent, ctx->nrec, ctx->recs, ctx->lab_cnt, ctx->lab_sx, ctx->lab_sy, H, W, M, RCAP); ctx->launches += 1; }
    if (which & 2) { euler_kernel<<<eg, wb, 0, st>>>(ctx->open_bits, ctx->euler4, H, W, WW); ctx->launches += 1; }
    int P2 = 1;
    while (P2 < M) P2 <<= 1;
    rank_kernel<<<dim3(batch, im.nimg), 1024, sizeof(int32_t) * P2, st>>>(ctx->nroots, ctx->rootlist, ctx->parent, ctx->lab_cnt, ctx->lab_sx, ctx->lab_sy,
                                                                         ctx->euler4, ctx->slot2label, ctx->centres, ctx->d_nlabels, ctx->croot,
                                                                         ctx->d_ncont, ctx->holes, H, W, M, P2, im.img0, ctx->d_status);
    ctx->launches += 1;
    if (which & 2) {
        // background pass of the opened image: every thread of a hole-free frame returns at once
        const dim3 bg((WW + 63) / 64, (H + BG_ROWS - 1) / BG_ROWS, batch);
        bg_init_kernel<<<bg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
        bg_merge_kernel<<<bg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
        bg_flatten_kernel<<<bg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
        ctx->launches += 3;
    }
    return cudaGetLastError();
}
