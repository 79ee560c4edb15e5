// the decimated transform next to the SWT.
//
//   b200_dwt2_fwd  <- DWTTransform._apply_wavelet   /root/reference/main/transforms/custom_transforms.py:196-200
//                     pywt.wavedec2(channel, wavelet, level=level) with PyWavelets' default mode 'symmetric', keeping
//                     coeffs[0] and coeffs[1]: (cA, cH, cV, cD) of the coarsest level (config/transform/cifar_dwt.yaml).
//
// PyWavelets' published algorithm (the package is absent here, like for the SWT): one level along an axis of length N is
//      y[o] = sum_j h[j] * xe[2o + 1 - j],   o = 0 .. floor((N + F - 1) / 2) - 1,
// xe = x extended by half-sample symmetry (... x1 x0 | x0 x1 ... x(N-1) | x(N-1) x(N-2) ...); dwt2 filters axis -2 first,
// then axis -1; keys aa = cA, da = cH, ad = cV, dd = cD; the next level consumes cA.  Legacy CNN configs on 32 x 32
// images use this transform, so the kernels are the plain two-pass form (one thread per output coefficient, the
// intermediate planes in a caller-provided workspace), not the tiled shared-memory pipeline of the SWT.
#include <algorithm>

#include "common.cuh"

namespace b200 {

struct DwtTaps {
    float lo[B200_SWT_MAX_FILTER], hi[B200_SWT_MAX_FILTER];
    int F;
};

__device__ __forceinline__ int dwt_reflect(int n, int N) {      // half-sample symmetric extension, any distance
    while (n < 0 || n >= N) n = n < 0 ? -1 - n : 2 * N - 1 - n;
    return n;
}

// along H: in [P][Hin][W] (plane stride in_ps elements; uint8 scaled by 1/255 or float32) -> tmp [P][2][Hout][W]
template <bool U8>
__global__ void __launch_bounds__(256) dwt_rows_kernel(const void *__restrict__ in, long long in_ps, float *__restrict__ tmp,
                                                       long long planes, int Hin, int Hout, int W, const __grid_constant__ DwtTaps t) {
    const long long total = planes * Hout * W;
    for (long long u = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; u < total;
         u += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(u % W);
        const long long r = u / W;
        const int yo = static_cast<int>(r % Hout);
        const long long p = r / Hout;
        float a = 0.f, d = 0.f;
        for (int j = 0; j < t.F; ++j) {
            const size_t idx = static_cast<size_t>(p * in_ps) + static_cast<size_t>(dwt_reflect(2 * yo + 1 - j, Hin)) * W + x;
            float v;
            if (U8) {
                v = __fdiv_rn(static_cast<float>(static_cast<const uint8_t *>(in)[idx]), 255.0f);      // custom_transforms.py:147
            } else {
                v = static_cast<const float *>(in)[idx];
            }
            a += t.lo[j] * v, d += t.hi[j] * v;
        }
        float *o = tmp + (static_cast<size_t>(p) * 2 * Hout + yo) * W + x;
        o[0] = a, o[static_cast<size_t>(Hout) * W] = d;
    }
}

// along W: tmp [P][2][Hout][W] -> out [P][bands][Hout][Wout]; bands = 4: (aa, da, ad, dd) = (cA, cH, cV, cD); bands = 1: cA only
__global__ void __launch_bounds__(256) dwt_cols_kernel(const float *__restrict__ tmp, float *__restrict__ out, long long planes, int Hout,
                                                       int W, int Wout, int bands, const __grid_constant__ DwtTaps t) {
    const long long total = planes * Hout * Wout;
    for (long long u = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; u < total;
         u += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xo = static_cast<int>(u % Wout);
        const long long r = u / Wout;
        const int y = static_cast<int>(r % Hout);
        const long long p = r / Hout;
        const float *tl = tmp + (static_cast<size_t>(p) * 2 * Hout + y) * W, *th = tl + static_cast<size_t>(Hout) * W;
        float aa = 0.f, ad = 0.f, da = 0.f, dd = 0.f;
        for (int j = 0; j < t.F; ++j) {
            const int n = dwt_reflect(2 * xo + 1 - j, W);
            const float l = tl[n], h = th[n];
            aa += t.lo[j] * l, ad += t.hi[j] * l, da += t.lo[j] * h, dd += t.hi[j] * h;
        }
        const size_t plane = static_cast<size_t>(Hout) * Wout;
        float *o = out + static_cast<size_t>(p) * bands * plane + static_cast<size_t>(y) * Wout + xo;
        o[0] = aa;
        if (bands == 4) o[plane] = da, o[2 * plane] = ad, o[3 * plane] = dd;
    }
}

struct DwtLayout {
    size_t off_tmp, off_a, off_b, bytes;
};
static DwtLayout dwt_layout(long long planes, int H, int W, int F, int level) {
    DwtLayout l{};
    const int h1 = (H + F - 1) / 2, w1 = (W + F - 1) / 2, h2 = (h1 + F - 1) / 2, w2 = (w1 + F - 1) / 2;
    size_t o = 0;
    auto take = [&](size_t n) {
        const size_t at = o;
        o = round_up<size_t>(o + n, 256);
        return at;
    };
    l.off_tmp = take(static_cast<size_t>(planes) * 2 * h1 * W * sizeof(float));
    l.off_a = take(level > 1 ? static_cast<size_t>(planes) * h1 * w1 * sizeof(float) : 0);      // cA of levels 1, 3, ...
    l.off_b = take(level > 2 ? static_cast<size_t>(planes) * h2 * w2 * sizeof(float) : 0);      // cA of levels 2, 4, ...
    l.bytes = o;
    return l;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_dwt2_workspace_bytes(long long planes, int H, int W, int F, int level) {
    if (planes < 1 || H < 1 || W < 1 || F < 2 || (F & 1) || F > B200_SWT_MAX_FILTER || level < 1) return 0;
    return dwt_layout(planes, H, W, F, level).bytes;
}

int b200_dwt2_fwd(const void *in, int in_is_u8, float *out, long long planes, int H, int W, const float *dec_lo,
                  const float *dec_hi, int F, int level, void *workspace, size_t workspace_bytes, b200_stream_t stream) {
    if (!in || !out || !dec_lo || !dec_hi || planes < 1 || H < 1 || W < 1 || level < 1) return B200_ERR_INVALID_ARG;
    if (F < 2 || (F & 1)) return B200_ERR_INVALID_ARG;
    if (F > B200_SWT_MAX_FILTER || level > 16) return B200_ERR_UNSUPPORTED;
    const DwtLayout l = dwt_layout(planes, H, W, F, level);
    if (!workspace || workspace_bytes < l.bytes) return B200_ERR_WORKSPACE;
    DwtTaps t{};
    t.F = F;
    for (int i = 0; i < F; ++i) t.lo[i] = dec_lo[i], t.hi[i] = dec_hi[i];
    cudaStream_t st = as_stream(stream);
    unsigned char *w = static_cast<unsigned char *>(workspace);
    float *tmp = reinterpret_cast<float *>(w + l.off_tmp);
    float *ll[2] = {reinterpret_cast<float *>(w + l.off_a), reinterpret_cast<float *>(w + l.off_b)};
    const void *src = in;
    int h = H, wd = W;
    const int max_grid = sm_count() * 32;
    for (int lv = 1; lv <= level; ++lv) {
        const int ho = (h + F - 1) / 2, wo = (wd + F - 1) / 2;
        const bool last = lv == level;
        const long long n_rows = planes * ho * wd, n_cols = planes * ho * wo;
        const int g_rows = static_cast<int>(std::min<long long>(ceil_div<long long>(n_rows, 256), max_grid));
        const int g_cols = static_cast<int>(std::min<long long>(ceil_div<long long>(n_cols, 256), max_grid));
        if (lv == 1 && in_is_u8)
            dwt_rows_kernel<true><<<g_rows, 256, 0, st>>>(src, static_cast<long long>(h) * wd, tmp, planes, h, ho, wd, t);
        else
            dwt_rows_kernel<false><<<g_rows, 256, 0, st>>>(src, static_cast<long long>(h) * wd, tmp, planes, h, ho, wd, t);
        B200_LAUNCH_CHECK("dwt_rows_kernel");
        float *dst = last ? out : ll[(lv - 1) & 1];
        dwt_cols_kernel<<<g_cols, 256, 0, st>>>(tmp, dst, planes, ho, wd, wo, last ? 4 : 1, t);
        B200_LAUNCH_CHECK("dwt_cols_kernel");
        src = dst, h = ho, wd = wo;
    }
    return B200_OK;
}

}  // extern "C"
