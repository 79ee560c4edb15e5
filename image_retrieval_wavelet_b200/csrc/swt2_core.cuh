// HP-SWT core: undecimated separable wavelet stencil over one shared-memory tile.
//
// Replaces pywt.swt2 as called by SWTTransform._apply_wavelet
// (/root/reference/main/transforms/custom_transforms.py:163-166): level l (dilation s = 2^(l-1)) is the periodised FIR
//      y[n] = sum_t h[t] * x[(n + s * (F/2 - t)) mod N]
// along axis -2 and then axis -1 with (dec_lo | dec_hi); level l+1 consumes level l's LL; only the level-L bands
// (cA, cH, cV, cD) are kept.  The two 1-D passes commute, so the tile is filtered along W first (taps read from
// shared memory into registers) and along H second from a register sliding window: one pass over the tile per
// level, no intermediate plane ever leaves the SM, and the four sub-bands are written once with vector stores.
//
// Everything here is __host__ __device__ and expressed per (tid, nthreads) so that tests/ can run the very same
// indexing code on the CPU (csrc/hostsim.cu) — there is no CPU product path.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace b200 {

constexpr int kSwtGuard = 32;   // floats of slack before/after each tile buffer (garbage column groups may touch it)

struct SwtGeom {
    int B, C, H, W;          // planes = B*C
    int level, F;
    int vec;                 // 4 when W % 4 == 0 else 2
    int TH, TW;              // output tile
    int tiles_y, tiles_x;
    int top, left;           // rows/cols needed before the tile = (2^L - 1) * (F/2 - 1)
    int bot, right;          // rows/cols needed after the tile  = (2^L - 1) * F/2
    int padL;                // left halo rounded up to a multiple of 4: tile column 0 sits at buffer column padL
    int RH, RWp;             // buffer rows, padded row stride in floats (multiple of 4)
    int run;                 // output rows per thread and sliding-window run
    int threads;             // CTA size
    int in_is_u8;
    int nbuf;                // 1 (level 1) or 2 (ping-pong for the intermediate levels)
    float lo[20], hi[20];
};

// aligned VEC-wide shared-memory access (p is VEC*4-byte aligned by construction: RWp % 4 == 0, j0 % VEC == 0)
template <int VEC>
__host__ __device__ __forceinline__ void swt_ld_vec(const float *p, float *d) {
#ifdef __CUDA_ARCH__
    if constexpr (VEC == 4) {
        const float4 v = *reinterpret_cast<const float4 *>(p);
        d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    } else {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        d[0] = v.x, d[1] = v.y;
    }
#else
    for (int e = 0; e < VEC; ++e) d[e] = p[e];
#endif
}
template <int VEC>
__host__ __device__ __forceinline__ void swt_st_vec(float *p, const float *s) {
#ifdef __CUDA_ARCH__
    if constexpr (VEC == 4)
        *reinterpret_cast<float4 *>(p) = make_float4(s[0], s[1], s[2], s[3]);
    else
        *reinterpret_cast<float2 *>(p) = make_float2(s[0], s[1]);
#else
    for (int e = 0; e < VEC; ++e) p[e] = s[e];
#endif
}

__host__ __device__ __forceinline__ int swt_wrap(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// ---------------------------------------------------------------------------------------------- tile load
// Region rows [ty*TH - top, ty*TH + TH + bot) x cols [tx*TW - left, tx*TW + TW + right), wrapped, converted to
// float32 (/255 exactly as custom_transforms.py:147) into buf[row][padL - left + col].
template <typename LoadU8x4, typename LoadF32x4>
__host__ __device__ __forceinline__ void swt_load_tile(const SwtGeom &g, const void *in_plane, float *buf, int ty, int tx,
                                                       int tid, int nthreads, LoadU8x4 ld_u8x4, LoadF32x4 ld_f32x4) {
    const int r_first = ty * g.TH - g.top;
    const int c_first = tx * g.TW - g.left;
    const int ncols = g.left + g.TW + g.right;
    const int col0 = g.padL - g.left;                 // buffer column of region column 0
    const uint8_t *in8 = static_cast<const uint8_t *>(in_plane);
    const float *in32 = static_cast<const float *>(in_plane);
    // body: aligned groups of 4 columns that neither wrap nor straddle the row end; everything else scalar.
    // A group starts at buffer column 4m (global column c_first - col0 + 4m); vector loads need W % 4 == 0.
    const int groups = g.RWp / 4;
    const bool can_vec = (g.W % 4) == 0;
    for (int idx = tid; idx < g.RH * groups; idx += nthreads) {
        const int i = idx / groups;
        const int m = idx - i * groups;
        const int gr = swt_wrap(r_first + i, g.H);
        const int jb = 4 * m;                          // buffer column
        const int rc = jb - col0;                      // region column (may be < 0 or >= ncols: padding)
        float *dst = buf + i * g.RWp + jb;
        const int gc = c_first + rc;                   // unwrapped global column
        if (can_vec && rc >= 0 && rc + 4 <= ncols && gc >= 0 && gc + 4 <= g.W && (gc & 3) == 0) {
            float v[4];
            if (g.in_is_u8)
                ld_u8x4(in8 + static_cast<size_t>(gr) * g.W + gc, v);
            else
                ld_f32x4(in32 + static_cast<size_t>(gr) * g.W + gc, v);
            swt_st_vec<4>(dst, v);
        } else {
            float v1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float x = 0.f;
                if (rc + e >= 0 && rc + e < ncols) {
                    const int c = swt_wrap(gc + e, g.W);
                    x = g.in_is_u8 ? static_cast<float>(in8[static_cast<size_t>(gr) * g.W + c]) / 255.0f
                                   : in32[static_cast<size_t>(gr) * g.W + c];
                }
                v1[e] = x;
            }
            swt_st_vec<4>(dst, v1);
        }
    }
}

// ---------------------------------------------------------------------------------------------- one level
// Horizontal taps of VEC adjacent outputs at buffer position (row, j0): loads the aligned window once, then
// hl[v] = sum_t lo[t] * x[j0 + v + S*(F/2 - t)] (and hh with hi when BOTH).
template <int F, int S, int VEC, bool BOTH>
__host__ __device__ __forceinline__ void swt_hrow(const SwtGeom &g, const float *row_j0, float (&hl)[VEC], float (&hh)[VEC]) {
    constexpr int omin = -S * (F / 2 - 1);
    constexpr int omax = S * (F / 2);
    constexpr int amin = (omin >= 0) ? (omin / VEC) * VEC : -(((-omin) + VEC - 1) / VEC) * VEC;   // floor to VEC
    constexpr int nvec = (VEC - 1 + omax - amin) / VEC + 1;
    float w[nvec * VEC];
#pragma unroll
    for (int i = 0; i < nvec; ++i) swt_ld_vec<VEC>(row_j0 + amin + i * VEC, w + i * VEC);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int t = 0; t < F; ++t) {
            const float x = w[v + S * (F / 2 - t) - amin];
            a += g.lo[t] * x;
            if (BOTH) d += g.hi[t] * x;
        }
        hl[v] = a;
        hh[v] = d;
    }
}

// One level over the tile.  Output rows [oy0, oy1), output column groups [floor(ox0), ceil(ox1)) in buffer
// coordinates.  FINAL: the four bands go to global memory (out_plane: [4][H][W]) through `store`; otherwise LL goes
// to `dst` (same geometry as src).  A work unit is (column group, row residue mod S, run of `g.run` rows); the thread
// slides an F-deep window of horizontally filtered rows down its run.
template <int F, int S, int VEC, bool FINAL, typename Store>
__host__ __device__ __forceinline__ void swt_level(const SwtGeom &g, const float *src, float *dst, float *out_plane,
                                                   int oy0, int oy1, int ox0, int ox1, int row_g0, int col_g0, int tid,
                                                   int nthreads, Store store) {
    const int gx0 = ox0 / VEC, gx1 = (ox1 + VEC - 1) / VEC;
    const int ncg = gx1 - gx0;
    const int rows = oy1 - oy0;
    const int per_class = (rows + S - 1) / S;             // outputs per residue class (upper bound)
    const int nrun = (per_class + g.run - 1) / g.run;
    const int units = ncg * S * nrun;
    for (int u = tid; u < units; u += nthreads) {
        const int cg = u % ncg;
        const int rest = u / ncg;
        const int rho = rest % S;
        const int ru = rest / S;
        const int j0 = (gx0 + cg) * VEC;
        const int first = oy0 + rho + S * (ru * g.run);   // first output row of this unit
        if (first >= oy1) continue;
        float wl[F][VEC] = {}, wh[F][VEC] = {};
        // warm-up: horizontally filtered rows first - S*(F/2-1) ... first + S*(F/2 - 1), oldest first
#pragma unroll
        for (int m = 0; m < F - 1; ++m) {
            const int r = first - S * (F / 2 - 1) + S * m;
#pragma unroll
            for (int t = F - 1; t > 0; --t) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) wl[t][v] = wl[t - 1][v], wh[t][v] = wh[t - 1][v];
            }
            swt_hrow<F, S, VEC, FINAL>(g, src + r * g.RWp + j0, wl[0], wh[0]);
        }
        for (int m = 0; m < g.run; ++m) {
            const int i = first + S * m;
            if (i >= oy1) break;
#pragma unroll
            for (int t = F - 1; t > 0; --t) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) wl[t][v] = wl[t - 1][v], wh[t][v] = wh[t - 1][v];
            }
            swt_hrow<F, S, VEC, FINAL>(g, src + (i + S * (F / 2)) * g.RWp + j0, wl[0], wh[0]);
            // vertical taps: tap t multiplies the row i + S*(F/2 - t) = window slot t
            float ll[VEC], lh[VEC], hl[VEC], hh[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
#pragma unroll
                for (int t = 0; t < F; ++t) {
                    a += g.lo[t] * wl[t][v];               // lo along H, lo along W : cA (LL)
                    if (FINAL) {
                        b += g.hi[t] * wl[t][v];           // hi along H, lo along W : cH = 'da' (LH)
                        c += g.lo[t] * wh[t][v];           // lo along H, hi along W : cV = 'ad' (HL)
                        d += g.hi[t] * wh[t][v];           // hi along H, hi along W : cD (HH)
                    }
                }
                ll[v] = a, lh[v] = b, hl[v] = c, hh[v] = d;
            }
            if (FINAL) {
                const int gr = row_g0 + i, gc = col_g0 + j0;      // global row / column of this vector
                if (gr < g.H && gc < g.W) {
                    const size_t plane = static_cast<size_t>(g.H) * g.W;
                    float *o = out_plane + static_cast<size_t>(gr) * g.W + gc;
                    store(o, ll), store(o + plane, lh), store(o + 2 * plane, hl), store(o + 3 * plane, hh);
                }
            } else {
                swt_st_vec<VEC>(dst + i * g.RWp + j0, ll);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- tile programs
// The per-CTA program is written once, as a sequence of phases handed to `exec`: on the device a phase runs as
// phase(threadIdx.x, blockDim.x) followed by __syncthreads(); the CPU simulator runs it for tid = 0..nthreads-1.
struct SwtTileId {
    int plane, ty, tx;
};
__host__ __device__ __forceinline__ SwtTileId swt_tile_id(const SwtGeom &g, long long bid) {
    SwtTileId t;
    t.tx = static_cast<int>(bid % g.tiles_x);
    t.ty = static_cast<int>((bid / g.tiles_x) % g.tiles_y);
    t.plane = static_cast<int>(bid / (static_cast<long long>(g.tiles_x) * g.tiles_y));
    return t;
}

template <int F, int VEC, int LEVEL, typename Exec, typename Store, typename LdU8, typename LdF32>
__host__ __device__ __forceinline__ void swt_tile_program(const SwtGeom &g, const void *in, float *out, long long bid,
                                                          float *smem, Exec exec, Store store, LdU8 ld_u8, LdF32 ld_f32) {
    const SwtTileId id = swt_tile_id(g, bid);
    const size_t plane_px = static_cast<size_t>(g.H) * g.W;
    const void *in_plane = g.in_is_u8 ? static_cast<const void *>(static_cast<const uint8_t *>(in) + id.plane * plane_px)
                                      : static_cast<const void *>(static_cast<const float *>(in) + id.plane * plane_px);
    float *out_plane = out + static_cast<size_t>(id.plane) * 4 * plane_px;
    const size_t buf_floats = static_cast<size_t>(g.RH) * g.RWp + 2 * kSwtGuard;
    float *a = smem + kSwtGuard;
    float *b = a + buf_floats;
    exec([&](int tid, int n) { swt_load_tile(g, in_plane, a, id.ty, id.tx, tid, n, ld_u8, ld_f32); });
    constexpr int hb = F / 2 - 1, ha = F / 2;          // halo of a dilation-1 step
    int vb = 0, ve = g.RH;                             // valid rows of the current approximation
    int cb = g.padL - g.left, ce = g.padL + g.TW + g.right;
    if constexpr (LEVEL >= 2) {
        vb += hb, ve -= ha, cb += hb, ce -= ha;
        exec([&](int tid, int n) { swt_level<F, 1, VEC, false>(g, a, b, nullptr, vb, ve, cb, ce, 0, 0, tid, n, store); });
        float *t = a; a = b; b = t;
    }
    if constexpr (LEVEL >= 3) {
        vb += 2 * hb, ve -= 2 * ha, cb += 2 * hb, ce -= 2 * ha;
        exec([&](int tid, int n) { swt_level<F, 2, VEC, false>(g, a, b, nullptr, vb, ve, cb, ce, 0, 0, tid, n, store); });
        float *t = a; a = b; b = t;
    }
    exec([&](int tid, int n) {
        swt_level<F, (1 << (LEVEL - 1)), VEC, true>(g, a, nullptr, out_plane, g.top, g.top + g.TH, g.padL, g.padL + g.TW,
                                                    id.ty * g.TH - g.top, id.tx * g.TW - g.padL, tid, n, store);
    });
}

// Generic fallback (any even F <= 20, level <= 4): runtime taps, one pixel per work item, three buffers.
__host__ __device__ __forceinline__ void swt_gen_hpass(const SwtGeom &g, const float *src, float *dlo, float *dhi, int S,
                                                       int r0, int r1, int c0, int c1, int tid, int n) {
    const int w = c1 - c0;
    for (int idx = tid; idx < (r1 - r0) * w; idx += n) {
        const int i = r0 + idx / w, j = c0 + idx % w;
        float a = 0.f, d = 0.f;
        for (int t = 0; t < g.F; ++t) {
            const float x = src[i * g.RWp + j + S * (g.F / 2 - t)];
            a += g.lo[t] * x;
            d += g.hi[t] * x;
        }
        dlo[i * g.RWp + j] = a;
        if (dhi) dhi[i * g.RWp + j] = d;
    }
}
__host__ __device__ __forceinline__ void swt_gen_vpass_ll(const SwtGeom &g, const float *tlo, float *dst, int S, int r0,
                                                          int r1, int c0, int c1, int tid, int n) {
    const int w = c1 - c0;
    for (int idx = tid; idx < (r1 - r0) * w; idx += n) {
        const int i = r0 + idx / w, j = c0 + idx % w;
        float a = 0.f;
        for (int t = 0; t < g.F; ++t) a += g.lo[t] * tlo[(i + S * (g.F / 2 - t)) * g.RWp + j];
        dst[i * g.RWp + j] = a;
    }
}
__host__ __device__ __forceinline__ void swt_gen_vpass_final(const SwtGeom &g, const float *tlo, const float *thi,
                                                             float *out_plane, int S, int row_g0, int col_g0, int tid,
                                                             int n) {
    const size_t plane = static_cast<size_t>(g.H) * g.W;
    for (int idx = tid; idx < g.TH * g.TW; idx += n) {
        const int i = g.top + idx / g.TW, j = g.padL + idx % g.TW;
        const int gr = row_g0 + i, gc = col_g0 + j;
        if (gr >= g.H || gc >= g.W) continue;
        float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
        for (int t = 0; t < g.F; ++t) {
            const float xl = tlo[(i + S * (g.F / 2 - t)) * g.RWp + j], xh = thi[(i + S * (g.F / 2 - t)) * g.RWp + j];
            a += g.lo[t] * xl, b += g.hi[t] * xl, c += g.lo[t] * xh, d += g.hi[t] * xh;
        }
        float *o = out_plane + static_cast<size_t>(gr) * g.W + gc;
        o[0] = a, o[plane] = b, o[2 * plane] = c, o[3 * plane] = d;
    }
}

template <typename Exec, typename LdU8, typename LdF32>
__host__ __device__ __forceinline__ void swt_generic_program(const SwtGeom &g, const void *in, float *out, long long bid,
                                                             float *smem, Exec exec, LdU8 ld_u8, LdF32 ld_f32) {
    const SwtTileId id = swt_tile_id(g, bid);
    const size_t plane_px = static_cast<size_t>(g.H) * g.W;
    const void *in_plane = g.in_is_u8 ? static_cast<const void *>(static_cast<const uint8_t *>(in) + id.plane * plane_px)
                                      : static_cast<const void *>(static_cast<const float *>(in) + id.plane * plane_px);
    float *out_plane = out + static_cast<size_t>(id.plane) * 4 * plane_px;
    const size_t buf_floats = static_cast<size_t>(g.RH) * g.RWp + 2 * kSwtGuard;
    float *x = smem + kSwtGuard, *tlo = x + buf_floats, *thi = tlo + buf_floats;
    exec([&](int tid, int n) { swt_load_tile(g, in_plane, x, id.ty, id.tx, tid, n, ld_u8, ld_f32); });
    const int hb = g.F / 2 - 1, ha = g.F / 2;
    int vb = 0, ve = g.RH, cb = g.padL - g.left, ce = g.padL + g.TW + g.right;
    for (int lv = 1; lv < g.level; ++lv) {
        const int S = 1 << (lv - 1);
        cb += S * hb, ce -= S * ha;
        exec([&](int tid, int n) { swt_gen_hpass(g, x, tlo, nullptr, S, vb, ve, cb, ce, tid, n); });
        vb += S * hb, ve -= S * ha;
        exec([&](int tid, int n) { swt_gen_vpass_ll(g, tlo, x, S, vb, ve, cb, ce, tid, n); });
    }
    const int S = 1 << (g.level - 1);
    exec([&](int tid, int n) { swt_gen_hpass(g, x, tlo, thi, S, vb, ve, g.padL, g.padL + g.TW, tid, n); });
    exec([&](int tid, int n) {
        swt_gen_vpass_final(g, tlo, thi, out_plane, S, id.ty * g.TH - g.top, id.tx * g.TW - g.padL, tid, n);
    });
}

}  // namespace b200
