// vbs_segplan.h - grid plan of the strip-marching kernels (blur, NCC).  Plain C++ without CUDA types, so that the CPU
// tests can compile it (tests/hostcheck) and check that every plan covers every row exactly once.
#pragma once

// Row-segment plan of the two strip-marching kernels (blur, NCC).  A CTA marches down one 128-px strip of one frame;
// every CTA pays `lead` halo steps before its first output row, so tall CTAs are cheaper per row, but whole-height CTAs
// alone leave the last wave of the grid partly empty.  The first `n_full` (frame, strip) items therefore run whole-height
// (a multiple of the resident-CTA slots: full waves), the rest are cut into `vsegs` row segments that fill the tail
// (256 1080p frames: blur 3.91 -> 3.80 ms, NCC 2.32 -> 2.21 ms against two segments for every item).
// CTA b < n_full: item b, all rows; otherwise item n_full + (b - n_full) / vsegs, segment (b - n_full) % vsegs.
struct VbsSegPlan { int n_full, vsegs, seg_rows, ctas; };
inline VbsSegPlan vbs_seg_plan(int H, long long items, int slots, int lead, int rb, double lead_cost, bool mixed) {
    auto seg_rows_of = [&](int vs) { return ((H + vs - 1) / vs + rb - 1) / rb * rb; };
    auto steps_of = [&](int rows) { return (double)((rows + rb - 1) / rb) + lead_cost * lead; };
    VbsSegPlan best{0, 1, seg_rows_of(1), 0};
    double best_cost = -1.0;
    // one wave's worth of items always stays segmented: whatever the real number of resident CTAs is (another stream's
    // kernels take slots too), short CTAs are left to fill the gaps
    // while the last tall ones finish.  Short grids (< 3 waves) stay uniform: a wave of tall NCC CTAs would also keep the
    // high-priority open-mask branch off the SMs for its whole lifetime (batch 64: 0.83 -> 0.93 ms with it).
    const long long kmax = mixed && items / slots >= 3 ? items / slots - 1 : 0;
    for (long long k = kmax; k >= 0 && k + 3 > kmax; --k) {
        const long long rest = items - k * slots;
        for (int vs = 1; vs <= 16; ++vs) {
            if (vs > 1 && (H + vs - 1) / vs < 64) break;            // at least 64 output rows per segment
            const int sr = seg_rows_of(vs), nv = (H + sr - 1) / sr;
            if (nv != vs) continue;
            const long long waves = (rest * vs + slots - 1) / slots;
            const double cost = (double)k * steps_of(H) + (double)waves * steps_of(sr);
            if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best = {(int)(k * slots), vs, sr, 0}; }
            if (rest == 0) break;
        }
    }
    best.ctas = (int)(best.n_full + (items - best.n_full) * best.vsegs);
    return best;
}
