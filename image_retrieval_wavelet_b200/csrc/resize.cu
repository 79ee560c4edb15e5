// the pixel step in front of the SWT, on the device.
//
//   b200_resize_u8  <- BaseWaveletTransform.fix_size            /root/reference/main/transforms/custom_transforms.py:132-139
//                      (PIL `image.resize((new_w, new_h), resample=Image.BICUBIC)` up to a multiple of 2^level: 518 -> 520
//                      for levels 2-3) and the PIL resize behind torchvision's Resize in the eval transforms
//                      (config/transform/NAME.yaml: Resize(256) -> CenterCrop(224), bilinear with antialiasing).
//
// The arithmetic is Pillow's (third-party, pinned `pillow==8.2.0` in requirements.txt:5; src/libImaging/Resample.c, the
// 8-bit path, unchanged through the 12.x release installed here), restated:
//   * per output coordinate, the window [xmin, xmin + n) of input samples and its weights
//       w(x) = filter((x + xmin - center + 0.5) / filterscale), center = (xx + 0.5) * in/out, support = filter support *
//       max(in/out, 1), normalised to sum 1 in double precision (precompute_coeffs), then rounded to 22-bit fixed point
//       (normalize_coeffs_8bpc);
//   * a horizontal pass into a uint8 intermediate, then a vertical pass (ImagingResampleInner: a pass is skipped when
//       its size does not change), each output = clip8((2^21 + sum pixel * weight) >> 22).
// The tables are built on the host in the same double arithmetic (a few hundred entries) and copied to the caller's
// workspace; the two kernels are one thread per 4 output bytes, HBM-bound at ~2 bytes moved per pixel and pass.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace b200 {

constexpr int kResizePrecisionBits = 32 - 8 - 2;
constexpr int kResizeMaxTaps = 64;

static double resize_filter(int filter, double x) {
    if (x < 0.0) x = -x;
    if (filter == B200_RESIZE_BILINEAR) return x < 1.0 ? 1.0 - x : 0.0;
    const double a = -0.5;                                   // bicubic, Keys a = -0.5 (Resample.c: bicubic_filter)
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static double resize_support(int filter) { return filter == B200_RESIZE_BILINEAR ? 1.0 : 2.0; }

static int resize_ksize(int in_size, int out_size, int filter) {
    double filterscale = static_cast<double>(in_size) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    return static_cast<int>(std::ceil(resize_support(filter) * filterscale)) * 2 + 1;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc: bounds[2*xx] = first input sample, bounds[2*xx+1] = count,
// kk[xx*ksize + x] = 22-bit fixed-point weights (zero beyond the count).
static void resize_coeffs(int in_size, int out_size, int filter, int ksize, std::vector<int> &bounds, std::vector<int> &kk) {
    const double scale = static_cast<double>(in_size) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = resize_support(filter) * filterscale;
    bounds.assign(static_cast<size_t>(out_size) * 2, 0);
    kk.assign(static_cast<size_t>(out_size) * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        const double ss = 1.0 / filterscale;
        double ww = 0.0;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = resize_filter(filter, (x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            const double v = k[x] * (1 << kResizePrecisionBits);
            kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
        }
        bounds[2 * xx] = xmin, bounds[2 * xx + 1] = xmax;
    }
}

__device__ __forceinline__ uint32_t resize_clip8(int acc) {
    const int v = acc >> kResizePrecisionBits;
    return static_cast<uint32_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[p][y][xx] from in[p][y][xmin .. xmin + n): one thread = 4 adjacent output columns of one row
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long rows,
                                                       int W, int Wout, int ksize, const int2 *__restrict__ bounds,
                                                       const int *__restrict__ kk) {
    const int groups = (Wout + 3) / 4;
    const long long total = rows * groups;
    for (long long u = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; u < total;
         u += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = u / groups;
        const int x0 = static_cast<int>(u - row * groups) * 4;
        const uint8_t *src = in + row * W;
        uint32_t px[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            px[e] = 0;
            const int xx = x0 + e;
            if (xx < Wout) {
                const int2 b = __ldg(bounds + xx);
                const int *k = kk + static_cast<size_t>(xx) * ksize;
                int acc = 1 << (kResizePrecisionBits - 1);
                for (int x = 0; x < b.y; ++x) acc += static_cast<int>(__ldg(src + b.x + x)) * __ldg(k + x);
                px[e] = resize_clip8(acc);
            }
        }
        uint8_t *dst = out + row * Wout + x0;
        if (x0 + 3 < Wout && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
            *reinterpret_cast<uint32_t *>(dst) = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
        } else {
            for (int e = 0; e < 4 && x0 + e < Wout; ++e) dst[e] = static_cast<uint8_t>(px[e]);
        }
    }
}

// out[p][yy][x] from in[p][ymin .. ymin + n)[x]: one thread = 4 adjacent columns of one output row
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long planes,
                                                       int H, int Hout, int W, int ksize, const int2 *__restrict__ bounds,
                                                       const int *__restrict__ kk) {
    const int groups = (W + 3) / 4;
    const long long total = planes * Hout * groups;
    for (long long u = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; u < total;
         u += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long orow = u / groups;                       // p * Hout + yy
        const int x0 = static_cast<int>(u - orow * groups) * 4;
        const long long p = orow / Hout;
        const int yy = static_cast<int>(orow - p * Hout);
        const int2 b = __ldg(bounds + yy);
        const int *k = kk + static_cast<size_t>(yy) * ksize;
        const uint8_t *src = in + (p * H + b.x) * W + x0;
        const bool vec = x0 + 3 < W && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (W & 3) == 0;
        int acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = 1 << (kResizePrecisionBits - 1);
        for (int y = 0; y < b.y; ++y) {
            const int w = __ldg(k + y);
            const uint8_t *r = src + static_cast<size_t>(y) * W;
            if (vec) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(r));
                acc[0] += static_cast<int>(v & 0xffu) * w, acc[1] += static_cast<int>((v >> 8) & 0xffu) * w;
                acc[2] += static_cast<int>((v >> 16) & 0xffu) * w, acc[3] += static_cast<int>(v >> 24) * w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (x0 + e < W) acc[e] += static_cast<int>(__ldg(r + e)) * w;
            }
        }
        uint8_t *dst = out + orow * W + x0;
        if (vec) {
            *reinterpret_cast<uint32_t *>(dst) =
                resize_clip8(acc[0]) | (resize_clip8(acc[1]) << 8) | (resize_clip8(acc[2]) << 16) | (resize_clip8(acc[3]) << 24);
        } else {
            for (int e = 0; e < 4 && x0 + e < W; ++e) dst[e] = static_cast<uint8_t>(resize_clip8(acc[e]));
        }
    }
}

struct ResizeLayout {
    int ksh, ksv;
    size_t off_tmp, off_hb, off_hk, off_vb, off_vk, bytes;
};
static ResizeLayout resize_layout(long long planes, int H, int W, int Hout, int Wout, int filter) {
    ResizeLayout l{};
    const bool need_h = Wout != W, need_v = Hout != H;
    l.ksh = need_h ? resize_ksize(W, Wout, filter) : 0;
    l.ksv = need_v ? resize_ksize(H, Hout, filter) : 0;
    size_t o = 0;
    auto take = [&](size_t n) {
        const size_t at = o;
        o = round_up<size_t>(o + n, 256);
        return at;
    };
    l.off_tmp = take(need_h && need_v ? static_cast<size_t>(planes) * H * Wout : 0);
    l.off_hb = take(need_h ? static_cast<size_t>(Wout) * 8 : 0);
    l.off_hk = take(need_h ? static_cast<size_t>(Wout) * l.ksh * 4 : 0);
    l.off_vb = take(need_v ? static_cast<size_t>(Hout) * 8 : 0);
    l.off_vk = take(need_v ? static_cast<size_t>(Hout) * l.ksv * 4 : 0);
    l.bytes = o;
    return l;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_resize_workspace_bytes(long long planes, int H, int W, int Hout, int Wout, int filter) {
    if (planes < 1 || H < 1 || W < 1 || Hout < 1 || Wout < 1) return 0;
    if (filter != B200_RESIZE_BICUBIC && filter != B200_RESIZE_BILINEAR) return 0;
    return resize_layout(planes, H, W, Hout, Wout, filter).bytes;
}

int b200_resize_u8(const uint8_t *in, uint8_t *out, long long planes, int H, int W, int Hout, int Wout, int filter,
                   void *workspace, size_t workspace_bytes, b200_stream_t stream) {
    if (!in || !out || planes < 1 || H < 1 || W < 1 || Hout < 1 || Wout < 1) return B200_ERR_INVALID_ARG;
    if (filter != B200_RESIZE_BICUBIC && filter != B200_RESIZE_BILINEAR) return B200_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    const bool need_h = Wout != W, need_v = Hout != H;
    if (!need_h && !need_v) {                                   // PIL returns a copy
        B200_CUDA_TRY(cudaMemcpyAsync(out, in, static_cast<size_t>(planes) * H * W, cudaMemcpyDeviceToDevice, st));
        return B200_OK;
    }
    const ResizeLayout l = resize_layout(planes, H, W, Hout, Wout, filter);
    if (l.ksh > kResizeMaxTaps || l.ksv > kResizeMaxTaps) return B200_ERR_UNSUPPORTED;      // > ~15x reduction
    if (!workspace || workspace_bytes < l.bytes) return B200_ERR_WORKSPACE;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    std::vector<int> bounds, kk;
    // pageable-source cudaMemcpyAsync stages the host data before it returns: the vectors may die at scope exit
    if (need_h) {
        resize_coeffs(W, Wout, filter, l.ksh, bounds, kk);
        B200_CUDA_TRY(cudaMemcpyAsync(w + l.off_hb, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, st));
        B200_CUDA_TRY(cudaMemcpyAsync(w + l.off_hk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, st));
    }
    if (need_v) {
        resize_coeffs(H, Hout, filter, l.ksv, bounds, kk);
        B200_CUDA_TRY(cudaMemcpyAsync(w + l.off_vb, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, st));
        B200_CUDA_TRY(cudaMemcpyAsync(w + l.off_vk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, st));
    }
    const uint8_t *src = in;
    const int max_grid = sm_count() * 16;
    if (need_h) {
        uint8_t *dst = need_v ? w + l.off_tmp : out;
        const long long rows = planes * H;
        const long long blocks = ceil_div<long long>(rows * ((Wout + 3) / 4), 256);
        resize_h_kernel<<<static_cast<int>(blocks < max_grid ? blocks : max_grid), 256, 0, st>>>(
            src, dst, rows, W, Wout, l.ksh, reinterpret_cast<const int2 *>(w + l.off_hb), reinterpret_cast<const int *>(w + l.off_hk));
        B200_LAUNCH_CHECK("resize_h_kernel");
        src = dst;
    }
    if (need_v) {
        const long long blocks = ceil_div<long long>(planes * Hout * ((Wout + 3) / 4), 256);
        resize_v_kernel<<<static_cast<int>(blocks < max_grid ? blocks : max_grid), 256, 0, st>>>(
            src, out, planes, H, Hout, Wout, l.ksv, reinterpret_cast<const int2 *>(w + l.off_vb),
            reinterpret_cast<const int *>(w + l.off_vk));
        B200_LAUNCH_CHECK("resize_v_kernel");
    }
    return B200_OK;
}

}  // extern "C"
