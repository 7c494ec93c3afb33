// k_ccl.cu - K3b/K4b: connected components on bit images by union-find over WORD SEGMENTS.
//
// A segment is a maximal run of set bits inside one 32-bit word; its id is the pixel index of
// its first bit, so ids live in a dense [H*W] int32 array that is only ever touched at segment
// starts (sparse traffic, no clearing pass).  Links always point to the smaller index, hence a
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

// word kernels (background pass): block (64, 4) = 64 words x 4 rows, grid (ceil(WW/64), ceil(H/4), batch)
#define VBS_WORD_COORDS                                          \
    const int wx = blockIdx.x * 64 + threadIdx.x;               \
    const int y = blockIdx.y * 4 + threadIdx.y;                 \
    const int f = blockIdx.z;                                   \
    if (wx >= WW || y >= H) return;

// ---- 1. every segment start becomes its own root ------------------------------------------------
// FG: set bits.  BG (conditional on holes[f] != 0): cleared bits inside the image.
template <bool FG, bool BG>
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                        const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_WORD_COORDS
    const uint32_t raw = __ldg(bits + ((size_t)f * H + y) * WW + wx);
    int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    if (FG) {
        uint32_t starts = raw & ~(raw << 1);
        while (starts) { const int s = __ffs(starts) - 1; starts &= starts - 1; par[base + s] = base + s; }
    }
    if (BG && holes[f] != 0) {
        const uint32_t w = ~raw & valid_mask(wx, W);
        uint32_t starts = w & ~(w << 1);
        while (starts) { const int s = __ffs(starts) - 1; starts &= starts - 1; par[base + s] = base + s; }
    }
}

// ---- 2. link segments that touch -----------------------------------------------------------------
template <bool CONN8, bool INV, bool BORDER>
__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                         const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_WORD_COORDS
    if (INV && holes[f] == 0) return;          // no hole anywhere in this frame: background labels are not needed
    const uint32_t *img = bits + (size_t)f * H * WW;
    const uint32_t w = get_bits<INV>(img, y, wx, W, WW);
    if (!w) return;
    int32_t *par = parent + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    if ((w & 1u) && wx > 0) {
        const uint32_t pw = get_bits<INV>(img, y, wx - 1, W, WW);
        if (pw >> 31) unite(par, base, base - 32 + seg_start(pw, 31));
    }
    const uint32_t up = y > 0 ? get_bits<INV>(img, y - 1, wx, W, WW) : 0u;
    const uint32_t upl = (CONN8 && y > 0 && wx > 0) ? get_bits<INV>(img, y - 1, wx - 1, W, WW) : 0u;
    const uint32_t upr = (CONN8 && y > 0 && wx + 1 < WW) ? get_bits<INV>(img, y - 1, wx + 1, W, WW) : 0u;
    const int last_bit = (W - 1) - 32 * wx;         // position of pixel W-1 in this word (may be >= 32)
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
            unite(par, id, base - W + us);
        }
        if (CONN8) {
            if ((seg & 1u) && (upl >> 31)) unite(par, id, base - W - 32 + seg_start(upl, 31));
            if ((seg >> 31) && (upr & 1u)) unite(par, id, base - W + 32);
        }
        if (BORDER) {
            const bool touches = y == 0 || y == H - 1 || (wx == 0 && (seg & 1u)) ||
                                 (last_bit >= 0 && last_bit < 32 && ((seg >> last_bit) & 1u));
            if (touches) unite(par, id, OUTSIDE);
        }
    }
}

// background segments of frames with holes: point straight at the root (or at OUTSIDE)
__global__ void __launch_bounds__(256) ccl_flatten_bg_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ parent,
                                                              const int32_t *__restrict__ holes, int H, int W, int WW) {
    VBS_WORD_COORDS
    if (holes[f] == 0) return;
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


// ---- foreground, both images in one launch ---------------------------------------------------------
// Two levels.  (1) fg_tile_kernel: a CTA owns a tile of TLH rows x TLW words, runs the union-find of the
// segments inside it in SHARED memory, sums the moments of every tile-local component there, and
// only then touches global memory: parent[segment] = tile-local root (one write per segment) plus
// one 16-byte record per tile-local root.  (2) fg_border_kernel: the links that cross a tile edge
// are united in the global array.  Everything after that (root discovery, moments, ranking) walks
// the few records instead of the many segments.
struct FgImages {
    const uint32_t *bits[2];     // max_bits, open_bits
    int32_t *parent[2];          // parent, parent2
    int img0, nimg;              // images handled by this launch: img0 .. img0 + nimg - 1 (grid.z = batch * nimg)
};
constexpr int TLH = 64, TLW = 4;              // tile: 64 rows x 4 words (128 px)
constexpr int TPX = 32 * TLW;                 // tile width in pixels
constexpr int TN = TLH * TPX / 2;             // node slots: two segment starts are never adjacent, so (local pixel >> 1) is unique

// links of one word to its left / upper neighbours.  Node id of the segment that starts at bit s of a
// word whose first pixel has index `base` is (base + s) >> SHIFT; `stride` is the index distance of a row.
// Neighbour words that must not be linked by this caller are passed as 0.
template <bool CONN8, int SHIFT>
__device__ __forceinline__ void merge_word(int32_t *par, uint32_t w, uint32_t prev, uint32_t up, uint32_t upl, uint32_t upr, int base, int stride) {
    if ((w & 1u) && (prev >> 31)) unite(par, base >> SHIFT, (base - 32 + seg_start(prev, 31)) >> SHIFT);
    if (!(up | upl | upr)) return;
    uint32_t rem = w;
    while (rem) {
        const int s = __ffs(rem) - 1;
        const uint32_t seg = run_mask(w, s);
        rem &= ~seg;
        const int id = (base + s) >> SHIFT;
        uint32_t nb = seg;
        if (CONN8) nb |= (seg << 1) | (seg >> 1);
        uint32_t ov = up & nb;
        while (ov) {
            const int b = __ffs(ov) - 1;
            const int us = seg_start(up, b);
            ov &= ~run_mask(up, us);
            unite(par, id, (base - stride + us) >> SHIFT);
        }
        if (CONN8) {
            if ((seg & 1u) && (upl >> 31)) unite(par, id, (base - stride - 32 + seg_start(upl, 31)) >> SHIFT);
            if ((seg >> 31) && (upr & 1u)) unite(par, id, (base - stride + 32) >> SHIFT);
        }
    }
}

// block (TLW, TLH); grid (ceil(WW / TLW), ceil(H / TLH), nimg * batch)
__global__ void __launch_bounds__(TLW * TLH) fg_tile_kernel(FgImages im, int32_t *__restrict__ nrec, int4 *__restrict__ recs,
                                                             int H, int W, int WW, int RCAP, uint32_t *status) {
    extern __shared__ int32_t tile_smem[];
    int32_t *lp = tile_smem;                                          // [TN] tile-local union-find
    uint32_t *acc_cy = reinterpret_cast<uint32_t *>(tile_smem) + TN;  // [TN] per local root: pixel count (14 bits) | sum of local y << 14
    uint32_t *acc_x = acc_cy + TN;                                    // [TN] per local root: sum of local x
    uint32_t(*sw)[TLW + 2] = reinterpret_cast<uint32_t(*)[TLW + 2]>(acc_x + TN);   // [TLH + 1][TLW + 2] the tile's words, zero frame above / left / right
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int wx = blockIdx.x * TLW + lx, y = blockIdx.y * TLH + ly;
    const int f = blockIdx.z / im.nimg, img = im.img0 + blockIdx.z % im.nimg, z = 2 * f + img;
    const bool in = wx < WW && y < H;
    const uint32_t w = in ? __ldg(im.bits[img] + ((size_t)f * H + y) * WW + wx) : 0u;
    sw[ly + 1][lx + 1] = w;
    if (ly == 0) sw[0][lx + 1] = 0u;
    if (lx == 0) { sw[ly + 1][0] = 0u; sw[ly + 1][TLW + 1] = 0u; if (ly == 0) { sw[0][0] = 0u; sw[0][TLW + 1] = 0u; } }
    const int lbase = ly * TPX + 32 * lx;
    {
        uint32_t starts = w & ~(w << 1);
        while (starts) {
            const int n = (lbase + __ffs(starts) - 1) >> 1;
            starts &= starts - 1;
            lp[n] = n; acc_cy[n] = 0u; acc_x[n] = 0u;
        }
    }
    __syncthreads();
    if (w) {
        const uint32_t prev = sw[ly + 1][lx], up = sw[ly][lx + 1];
        if (img) merge_word<true, 1>(lp, w, prev, up, sw[ly][lx], sw[ly][lx + 2], lbase, TPX);
        else merge_word<false, 1>(lp, w, prev, up, 0u, 0u, lbase, TPX);
    }
    __syncthreads();
    {   // flatten, and add every segment's moments to its local root
        uint32_t rem = w;
        while (rem) {
            const int s = __ffs(rem) - 1;
            const uint32_t seg = run_mask(w, s);
            rem &= ~seg;
            const int n = (lbase + s) >> 1;
            const int r = find_root(lp, n);
            if (r != n) lp[n] = r;
            const uint32_t len = __popc(seg);
            atomicAdd(acc_cy + r, len | ((len * (uint32_t)ly) << 14));
            atomicAdd(acc_x + r, len * (uint32_t)(32 * lx + s) + len * (len - 1) / 2);
        }
    }
    __syncthreads();
    if (!w) return;
    int32_t *par = im.parent[img] + (size_t)f * H * W;
    const int x0 = blockIdx.x * TPX, y0 = blockIdx.y * TLH;
    uint32_t starts = w & ~(w << 1);
    while (starts) {
        const int s = __ffs(starts) - 1;
        starts &= starts - 1;
        const int n = (lbase + s) >> 1;
        const int r = lp[n];
        const int gid = y * W + 32 * wx + s;
        if (r != n) {
            // the root node covers local pixels 2r and 2r+1; its start is 2r if that bit is set, else 2r+1
            const int rp = 2 * r, ry = rp / TPX, rx = rp % TPX;
            const int rs = rx + (((sw[ry + 1][(rx >> 5) + 1] >> (rx & 31)) & 1u) ? 0 : 1);
            par[gid] = (y0 + ry) * W + x0 + rs;
            continue;
        }
        par[gid] = gid;
        const uint32_t cy = acc_cy[n], cnt = cy & 0x3fffu, sy = cy >> 14;
        const int slot = atomicAdd(nrec + z, 1);
        if (slot < RCAP) recs[(size_t)z * RCAP + slot] = make_int4(gid, (int)cnt, (int)(acc_x[n] + cnt * (uint32_t)x0), (int)(sy + cnt * (uint32_t)y0));
        else atomicOr(status, img == 0 ? VBS_DEV_LABEL_OVERFLOW : VBS_DEV_CONTOUR_OVERFLOW);
    }
}

// links across tile edges, in the global array.  block (64, 4) like the word kernels.
__global__ void __launch_bounds__(256) fg_border_kernel(FgImages im, int H, int W, int WW) {
    const int wx = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int f = blockIdx.z / im.nimg, img = im.img0 + blockIdx.z % im.nimg;
    if (wx >= WW || y >= H) return;
    const bool top = y % TLH == 0, left = wx % TLW == 0, right = wx % TLW == TLW - 1;
    if (!(top | left | right)) return;
    const uint32_t *img_bits = im.bits[img] + (size_t)f * H * WW;
    const uint32_t w = __ldg(img_bits + (size_t)y * WW + wx);
    if (!w) return;
    const uint32_t prev = (left && (w & 1u) && wx > 0) ? __ldg(img_bits + (size_t)y * WW + wx - 1) : 0u;
    uint32_t up = 0, upl = 0, upr = 0;
    if (y > 0) {
        const uint32_t *urow = img_bits + (size_t)(y - 1) * WW;
        if (top) up = __ldg(urow + wx);
        if (img) {                               // diagonal neighbours (8-connectivity)
            if ((top | left) && wx > 0) upl = __ldg(urow + wx - 1);
            if ((top | right) && wx + 1 < WW) upr = __ldg(urow + wx + 1);
        }
    }
    if (!(prev | up | upl | upr)) return;
    int32_t *par = im.parent[img] + (size_t)f * H * W;
    const int base = y * W + 32 * wx;
    if (img) merge_word<true, 0>(par, w, prev, up, upl, upr, base, W);
    else merge_word<false, 0>(par, w, prev, up, 0u, 0u, base, W);
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
__global__ void __launch_bounds__(256) euler_kernel(const uint32_t *__restrict__ bits, int32_t *__restrict__ euler4, int H, int W, int WW) {
    const int wx = blockIdx.x * 64 + threadIdx.x;            // 0 .. WW (one extra all-zero column)
    const int yy = (int)(blockIdx.y * 4 + threadIdx.y) - 1;  // top row of the quad: -1 .. H-1
    const int f = blockIdx.z;
    int v = 0;
    if (wx <= WW && yy < H) {
        const uint32_t *img = bits + (size_t)f * H * WW;
        auto ld = [&](int y, int w) -> uint32_t { return (y >= 0 && y < H && w >= 0 && w < WW) ? __ldg(img + (size_t)y * WW + w) : 0u; };
        const uint32_t b = ld(yy, wx), d = ld(yy + 1, wx);
        if (b | d | (wx > 0 ? 1u : 0u)) {
            const uint32_t a = (b << 1) | (ld(yy, wx - 1) >> 31), c = (d << 1) | (ld(yy + 1, wx - 1) >> 31);
            const uint32_t x1 = a ^ b, x2 = c ^ d, n1 = a & b, n2 = c & d;
            const uint32_t q1 = (x1 & ~x2 & ~n2) | (x2 & ~x1 & ~n1);
            const uint32_t q3 = (x1 & n2) | (x2 & n1);
            const uint32_t qd = x1 & x2 & ~(a ^ d);
            v = __popc(q1) - __popc(q3) - 2 * __popc(qd);
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
    const int H = ctx->H, W = ctx->W, WW = ctx->WW, M = ctx->M;
    const dim3 wb(64, 4);
    const dim3 wg((WW + 63) / 64, (H + 3) / 4, batch);
    const dim3 eg((WW + 1 + 63) / 64, (H + 1 + 3) / 4, batch);
    cudaStream_t st = ctx->stream;
    FgImages im;
    im.bits[0] = ctx->max_bits; im.bits[1] = ctx->open_bits; im.parent[0] = ctx->parent; im.parent[1] = ctx->parent2;
    im.img0 = (which & 1) ? 0 : 1;
    im.nimg = (which == 3) ? 2 : 1;
    const int RCAP = ctx->rcap;
    constexpr int TILE_SMEM = (3 * TN + (TLH + 1) * (TLW + 2)) * 4;
    static const cudaError_t attr = cudaFuncSetAttribute(fg_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
    if (attr != cudaSuccess) return attr;
    fg_tile_kernel<<<dim3((WW + TLW - 1) / TLW, (H + TLH - 1) / TLH, im.nimg * batch), dim3(TLW, TLH), TILE_SMEM, st>>>(im, ctx->nrec, ctx->recs, H, W, WW, RCAP, ctx->d_status);
    fg_border_kernel<<<dim3((WW + 63) / 64, (H + 3) / 4, im.nimg * batch), wb, 0, st>>>(im, H, W, WW);
    fg_roots_kernel<<<dim3((RCAP + 255) / 256, im.nimg * batch), 256, 0, st>>>(im, ctx->nrec, ctx->recs, ctx->nroots, ctx->rootlist, H, W, M, RCAP);
    ctx->launches += 3;
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
        ccl_init_kernel<false, true><<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
        ccl_merge_kernel<false, true, true><<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
        ccl_flatten_bg_kernel<<<wg, wb, 0, st>>>(ctx->open_bits, ctx->parent2, ctx->holes, H, W, WW);
        ctx->launches += 3;
    }
    return cudaGetLastError();
}
