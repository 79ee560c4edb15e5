// Host-side tile planner for the SWT kernel (shared by the CUDA launcher and the CPU host simulator used in tests).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>

#include "swt2_core.cuh"

namespace b200 {

inline bool swt_fast_path(int F, int level) { return F >= 2 && F <= 10 && (F % 2) == 0 && level >= 1 && level <= 3; }

inline size_t swt_smem_bytes(const SwtGeom &g) {
    return (static_cast<size_t>(g.nbuf) * (static_cast<size_t>(g.RH) * g.RWp + 2 * kSwtGuard)) * sizeof(float);
}

// Chooses the output tile.  Full-width row bands when they fit (no column halo traffic, contiguous rows);
// otherwise the width is split.  Cost = work amplification (RH*RWp)/(TH*TW) x (1 + SMs/CTAs): halo overhead against
// load balance (the block scheduler leaves about one CTA of imbalance per SM).  Tiles that let two CTAs share an
// SM's shared memory are tried first.
inline int swt_plan(SwtGeom &g, int B, int C, int H, int W, int F, int level, int in_is_u8, const float *lo, const float *hi,
                    int num_sms) {
    if (B < 1 || C < 1 || H < 1 || W < 1 || F < 2 || (F & 1) || F > 20 || level < 1 || level > 4) return -1;
    if (H % (1 << level) || W % (1 << level)) return -1;
    g = SwtGeom{};
    g.B = B, g.C = C, g.H = H, g.W = W, g.level = level, g.F = F, g.in_is_u8 = in_is_u8;
    for (int i = 0; i < 20; ++i) g.lo[i] = i < F ? lo[i] : 0.f, g.hi[i] = i < F ? hi[i] : 0.f;
    const bool fast = swt_fast_path(F, level);
    g.vec = fast ? ((W % 4 == 0) ? 4 : 2) : 1;
    const int span = (1 << level) - 1;
    g.top = g.left = span * (F / 2 - 1);
    g.bot = g.right = span * (F / 2);
    g.padL = (g.left + 3) / 4 * 4;
    g.nbuf = fast ? (level > 1 ? 2 : 1) : 3;
    const long long planes = static_cast<long long>(B) * C;
    double best = 1e30;
    int bth = 0, btw = 0;
    for (int nx = 1; nx <= 64; ++nx) {
        int tw = (W + nx - 1) / nx;
        tw = (tw + 3) / 4 * 4;
        if (nx == 1) tw = W;
        if (nx > 1 && tw < 32) break;
        const int rwp = (g.padL + tw + g.right + 3) / 4 * 4;
        for (int th : {128, 64, 56, 48, 32, 28, 24, 16, 14, 8, 4, 2}) {
            if (th > H) continue;
            const int rh = th + g.top + g.bot;
            const size_t bytes = (static_cast<size_t>(g.nbuf) * (static_cast<size_t>(rh) * rwp + 2 * kSwtGuard)) * 4;
            if (bytes > 200 * 1024) continue;
            const long long ctas = planes * ((H + th - 1) / th) * nx;
            double cost = (static_cast<double>(rh) * rwp) / (static_cast<double>(th) * tw);   // halo amplification
            cost *= static_cast<double>(((H + th - 1) / th) * th) / H;                         // ragged last band
            cost *= 1.0 + static_cast<double>(num_sms) / static_cast<double>(ctas);            // SM load imbalance ~ one CTA
            const double per_sm = std::min(8.0, std::floor(227.0 * 1024 / static_cast<double>(bytes + 1024)));
            cost *= 1.0 + 0.15 / per_sm;                                                      // CTAs per SM hide the load phase
            if (cost < best) best = cost, bth = th, btw = tw;
        }
    }
    if (const char *ov = std::getenv("B200_SWT_TILE")) {      // tuning override: "TH,TW"
        int th = 0, tw = 0;
        if (std::sscanf(ov, "%d,%d", &th, &tw) == 2 && th > 0 && tw > 0 && th <= H && tw <= W && tw % 4 == 0) {
            const int rwp = (g.padL + tw + g.right + 3) / 4 * 4;
            const size_t bytes = (static_cast<size_t>(g.nbuf) * (static_cast<size_t>(th + g.top + g.bot) * rwp + 2 * kSwtGuard)) * 4;
            if (bytes <= 220 * 1024) bth = th, btw = tw;
        }
    }
    if (bth == 0) return -2;
    g.TH = bth, g.TW = btw;
    g.tiles_y = (H + g.TH - 1) / g.TH;
    g.tiles_x = (W + g.TW - 1) / g.TW;
    g.RH = g.TH + g.top + g.bot;
    g.RWp = (g.padL + g.TW + g.right + 3) / 4 * 4;
    // threads: one work unit (column group x residue x run) per thread in the final level where possible
    const int S = 1 << (level - 1);
    const int vec = fast ? g.vec : 1;
    const int ncg = (g.TW + vec - 1) / vec;
    const int per_class = (g.TH + S - 1) / S;
    int run = per_class;
    const int min_run = std::max(4, 2 * F);      // amortise the F-1 warm-up rows of the sliding window
    while (run / 2 >= min_run && static_cast<long long>(ncg) * S * ((per_class + run / 2 - 1) / (run / 2)) <= 512) run /= 2;
    g.run = std::max(run, 1);
    const long long units = static_cast<long long>(ncg) * S * ((per_class + g.run - 1) / g.run);
    g.threads = static_cast<int>(std::min<long long>(512, std::max<long long>(64, (units + 31) / 32 * 32)));
    if (!fast) g.threads = 256;
    return 0;
}

}  // namespace b200
