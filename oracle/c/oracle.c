/*
 * CPU oracle in plain C (TEST INFRASTRUCTURE — see oracle/__init__.py).
 *
 * A scalar/OpenMP restatement of the two reference hot paths, used (a) to cross-check
 * the numpy oracle, (b) as the "port" CPU baseline timed by bench.py on the GPU box's
 * host cores.  Never linked into, or called from, the product library.
 *
 *  oracle_swt2        <- /root/reference/main/transforms/custom_transforms.py:145-166
 *                        (np.array(img)/255, per-channel pywt.swt2, keep coeffs[0], stack);
 *                        pywt.swt2 restated from PyWavelets' published algorithm: per level
 *                        l (dilation 2^(l-1)), periodised FIR along axis -2 then axis -1,
 *                        y[n] = sum_j h[j] * x[(n + 2^(l-1) * (F/2 - j)) mod N], float32
 *                        accumulate in ascending j; next level consumes 'aa'.
 *  oracle_maphashing  <- /root/reference/main/engine/accuracy_calculator.py:203-231
 *                        (per query: relevance row :31-37, hamming row :183-186, sort,
 *                        top-k, AP) with the tie order fixed to (distance, index).
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int wrap(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

/* one periodised, dilated analysis step along rows (axis -2) or columns (axis -1) */
static void swt_axis(const float *src, float *dst, int H, int W, const float *h, int F, int dil, int along_rows) {
    if (along_rows) {
        for (int r = 0; r < H; ++r) {
            float *o = dst + (size_t)r * W;
            for (int c = 0; c < W; ++c) o[c] = 0.f;
            for (int j = 0; j < F; ++j) {
                const float *s = src + (size_t)wrap(r + dil * (F / 2 - j), H) * W;
                const float hj = h[j];
                if (hj == 0.f) continue;
                for (int c = 0; c < W; ++c) o[c] += hj * s[c];
            }
        }
    } else {
        for (int r = 0; r < H; ++r) {
            const float *s = src + (size_t)r * W;
            float *o = dst + (size_t)r * W;
            for (int c = 0; c < W; ++c) {
                float acc = 0.f;
                for (int j = 0; j < F; ++j) acc += h[j] * s[wrap(c + dil * (F / 2 - j), W)];
                o[c] = acc;
            }
        }
    }
}

/* in: [B,C,H,W] uint8 (scaled by 1/255 like custom_transforms.py:147) or float32.
 * out: [B,C,4,H,W] float32, bands (cA, cH, cV, cD) of the coarsest level. */
int oracle_swt2(const void *in, int in_is_u8, float *out, int B, int C, int H, int W, const float *dec_lo,
                const float *dec_hi, int F, int level, int nthreads) {
    if (level < 1 || F < 2 || (F & 1) || H % (1 << level) || W % (1 << level)) return -1;
    const size_t plane = (size_t)H * W;
    const int planes = B * C;
    int failed = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < planes; ++p) {
        float *a = (float *)malloc(plane * sizeof(float));
        float *ra = (float *)malloc(plane * sizeof(float));
        float *rd = (float *)malloc(plane * sizeof(float));
        float *aa = (float *)malloc(plane * sizeof(float));
        if (!a || !ra || !rd || !aa) {
            failed = 1;
        } else {
            if (in_is_u8) {
                const uint8_t *s = (const uint8_t *)in + (size_t)p * plane;
                for (size_t i = 0; i < plane; ++i) a[i] = (float)s[i] / 255.0f;
            } else {
                memcpy(a, (const float *)in + (size_t)p * plane, plane * sizeof(float));
            }
            float *o = out + (size_t)p * 4 * plane;
            for (int lv = 1; lv <= level; ++lv) {
                const int dil = 1 << (lv - 1);
                swt_axis(a, ra, H, W, dec_lo, F, dil, 1);
                swt_axis(a, rd, H, W, dec_hi, F, dil, 1);
                if (lv < level) {
                    swt_axis(ra, aa, H, W, dec_lo, F, dil, 0);
                    float *t = a; a = aa; aa = t;
                } else {
                    swt_axis(ra, o + 0 * plane, H, W, dec_lo, F, dil, 0); /* aa = cA */
                    swt_axis(rd, o + 1 * plane, H, W, dec_lo, F, dil, 0); /* da = cH */
                    swt_axis(ra, o + 2 * plane, H, W, dec_hi, F, dil, 0); /* ad = cV */
                    swt_axis(rd, o + 3 * plane, H, W, dec_hi, F, dil, 0); /* dd = cD */
                }
            }
        }
        free(a); free(ra); free(rd); free(aa);
    }
    return failed ? -2 : 0;
}

static int cmp_u64(const void *x, const void *y) {
    const uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return (a > b) - (a < b);
}

/* codes: float32 +-1 [Q,B] / [N,B].  labels: label_mode 0 -> float32 multi-hot [.,L] (relevant iff
 * dot > 0); label_mode 1 -> float32 [.,1] compared for equality.  topk < 0 means "all".
 * ap_out[Q] (float64), tsum_out[Q]; returns mean AP over all Q queries via *map_out. */
int oracle_maphashing(const float *q, const float *ql, const float *r, const float *rl, int Q, int N, int B, int L,
                      int label_mode, long topk, double *ap_out, long *tsum_out, double *map_out, int nthreads) {
    if (Q < 0 || N < 0 || B < 1) return -1;
    long k = (topk < 0 || topk > N) ? N : topk;
    int failed = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        uint64_t *key = (uint64_t *)malloc((size_t)(N > 0 ? N : 1) * sizeof(uint64_t));
        unsigned char *rel = (unsigned char *)malloc((size_t)(N > 0 ? N : 1));
        if (!key || !rel) failed = 1;
#pragma omp for schedule(dynamic, 4)
        for (int i = 0; i < Q; ++i) {
            if (failed) continue;
            const float *qi = q + (size_t)i * B;
            for (int j = 0; j < N; ++j) {
                const float *rj = r + (size_t)j * B;
                float dot = 0.f;
                for (int b = 0; b < B; ++b) dot += qi[b] * rj[b];
                const float hamm = 0.5f * ((float)B - dot);          /* accuracy_calculator.py:185 */
                key[j] = ((uint64_t)(uint32_t)lrintf(hamm) << 32) | (uint32_t)j;
                if (label_mode == 0) {
                    float s = 0.f;
                    for (int l = 0; l < L; ++l) s += ql[(size_t)i * L + l] * rl[(size_t)j * L + l];
                    rel[j] = s > 0.f;                                   /* :34 */
                } else {
                    rel[j] = ql[i] == rl[j];                            /* :37 */
                }
            }
            qsort(key, (size_t)N, sizeof(uint64_t), cmp_u64);           /* (distance, index) order */
            long hits = 0;
            double acc = 0.0;
            for (long p = 0; p < k; ++p) {
                if (rel[(uint32_t)key[p]]) {
                    ++hits;
                    acc += (double)hits / (double)(p + 1);              /* :227-229 */
                }
            }
            ap_out[i] = hits ? acc / (double)hits : 0.0;
            if (tsum_out) tsum_out[i] = hits;
        }
        free(key);
        free(rel);
    }
    if (failed) return -2;
    double s = 0.0;
    for (int i = 0; i < Q; ++i) s += ap_out[i];
    if (map_out) *map_out = Q ? s / Q : NAN;                           /* :231 */
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
