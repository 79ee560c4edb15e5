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
#include <cstdlib>
#include <cstring>
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

// Launch shape of both passes: blockDim = (column groups of 4 output bytes, kResizeRows rows); one CTA covers
// kResizeRows consecutive rows so that the weight rows it reads stay in L1; no division anywhere on the device.
constexpr int kResizeRows = 4;

// out[row][xx] from in[row][xmin .. xmin + n): one thread = 4 adjacent output columns of one row
__global__ void __launch_bounds__(1024) resize_h_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long rows,
                                                        int W, int Wout, int ksize, const int2 *__restrict__ bounds,
                                                        const int *__restrict__ kk) {
    const long long row = static_cast<long long>(blockIdx.x) * kResizeRows + threadIdx.y;
    if (row >= rows) return;
    const uint8_t *src = in + row * W;
    uint8_t *dst_row = out + row * Wout;
    for (int x0 = 4 * threadIdx.x; x0 < Wout; x0 += 4 * blockDim.x) {
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
        uint8_t *dst = dst_row + x0;
        if (x0 + 3 < Wout && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
            *reinterpret_cast<uint32_t *>(dst) = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
        } else {
            for (int e = 0; e < 4 && x0 + e < Wout; ++e) dst[e] = static_cast<uint8_t>(px[e]);
        }
    }
}

// out[p][yy][x] from in[p][ymin .. ymin + n)[x]: one thread = 4 adjacent columns of one output row; grid = (row blocks, planes)
__global__ void __launch_bounds__(1024) resize_v_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int H, int Hout,
                                                        int W, int ksize, const int2 *__restrict__ bounds,
                                                        const int *__restrict__ kk) {
    const int yy = static_cast<int>(blockIdx.x) * kResizeRows + threadIdx.y;
    if (yy >= Hout) return;
    const size_t p = blockIdx.y;
    const int2 b = __ldg(bounds + yy);
    const int *k = kk + static_cast<size_t>(yy) * ksize;
    const uint8_t *src_row = in + (p * H + b.x) * W;
    uint8_t *dst_row = out + (p * Hout + yy) * W;
    for (int x0 = 4 * threadIdx.x; x0 < W; x0 += 4 * blockDim.x) {
        const uint8_t *src = src_row + x0;
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
        uint8_t *dst = dst_row + x0;
        if (vec && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
            *reinterpret_cast<uint32_t *>(dst) =
                resize_clip8(acc[0]) | (resize_clip8(acc[1]) << 8) | (resize_clip8(acc[2]) << 16) | (resize_clip8(acc[3]) << 24);
        } else {
            for (int e = 0; e < 4 && x0 + e < W; ++e) dst[e] = static_cast<uint8_t>(resize_clip8(acc[e]));
        }
    }
}

static dim3 resize_block(int width) {
    const int groups = (width + 3) / 4;
    const int tx = groups >= 256 ? 256 : (groups + 31) / 32 * 32;
    return dim3(tx, kResizeRows);
}

// ---- fused form: one CTA per 32 x 128 output tile of one plane.  The input window of the tile is staged once in shared
// memory (aligned 32-bit loads; a row of a 518-wide plane starts on any byte, so every staged row keeps its own byte
// offset), the horizontal pass writes the uint8 intermediate of exactly the rows the tile's vertical windows need into
// shared memory, the vertical pass writes the tile.  Same integer arithmetic per pixel as the two-kernel form (and as
// Pillow), but the intermediate never reaches HBM: 1 byte read + 1 byte written per pixel.  An axis whose size does not
// change runs with identity tables (window 1, weight 2^22: (2^21 + p * 2^22) >> 22 == p exactly).
constexpr int kRzTH = 32, kRzTW = 128, kRzThreads = 256;

struct ResizeFusedArgs {
    const uint8_t *in;
    uint8_t *out;
    const int2 *hb, *vb;
    const int *hk, *vk;
    int H, W, Hout, Wout, ksh, ksv;
    int max_rows, max_words;          // staged rows / 32-bit words per staged row (upper bounds over all tiles)
    long long plane0;                 // first plane of this launch (gridDim.z <= 65535 planes per launch)
};

// KSH / KSV > 0: the window length is a compile-time constant (tables are zero beyond each window's count, so the
// fixed-length sums are the same integers), taps unrolled and the thread's weights held in registers; 0: run-time counts.
template <int KSH, int KSV>
__global__ void __launch_bounds__(kRzThreads) resize_fused_kernel(const __grid_constant__ ResizeFusedArgs a) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    int2 *s_hb = reinterpret_cast<int2 *>(rz_smem);                                          // [kRzTW]
    int2 *s_vb = s_hb + kRzTW;                                                               // [kRzTH]
    uint32_t *s_in = reinterpret_cast<uint32_t *>(s_vb + kRzTH);                             // [max_rows][max_words]
    int *s_hk = reinterpret_cast<int *>(s_in + a.max_rows * a.max_words);                    // [kRzTW][ksh]
    int *s_vk = s_hk + kRzTW * a.ksh;                                                        // [kRzTH][ksv]
    int *s_off = s_vk + kRzTH * a.ksv;                                                       // [max_rows] byte offset of column xlo in the staged row
    uint8_t *s_tmp = reinterpret_cast<uint8_t *>(s_off + a.max_rows);                        // [max_rows + pad][kRzTW]
    const int tid = static_cast<int>(threadIdx.x);
    const int x0 = static_cast<int>(blockIdx.x) * kRzTW, y0 = static_cast<int>(blockIdx.y) * kRzTH;
    const size_t p = static_cast<size_t>(a.plane0) + blockIdx.z;
    const int nx = a.Wout - x0 < kRzTW ? a.Wout - x0 : kRzTW, ny = a.Hout - y0 < kRzTH ? a.Hout - y0 : kRzTH;
    // windows are monotone in the output coordinate: the tile's input window is [first.x, last.x + last.y)
    const int2 bx0 = __ldg(a.hb + x0), bx1 = __ldg(a.hb + x0 + nx - 1), by0 = __ldg(a.vb + y0), by1 = __ldg(a.vb + y0 + ny - 1);
    const int xlo = bx0.x, cols = bx1.x + bx1.y - bx0.x, ylo = by0.x, rows = by1.x + by1.y - by0.x;
    const int ksh = KSH ? KSH : a.ksh, ksv = KSV ? KSV : a.ksv;
    for (int i = tid; i < kRzTW * ksh; i += kRzThreads) s_hk[i] = i < nx * ksh ? __ldg(a.hk + static_cast<size_t>(x0) * ksh + i) : 0;
    for (int i = tid; i < kRzTH * ksv; i += kRzThreads) s_vk[i] = i < ny * ksv ? __ldg(a.vk + static_cast<size_t>(y0) * ksv + i) : 0;
    for (int i = tid; i < kRzTW; i += kRzThreads) s_hb[i] = i < nx ? __ldg(a.hb + x0 + i) : make_int2(xlo, 0);
    for (int i = tid; i < kRzTH; i += kRzThreads) s_vb[i] = i < ny ? __ldg(a.vb + y0 + i) : make_int2(ylo, 0);
    // stage: warp w takes rows w, w + 8, ...; its lanes the aligned words of the row
    const size_t plane_off = p * a.H * a.W;
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < rows; r += kRzThreads / 32) {
        const size_t base = plane_off + static_cast<size_t>(ylo + r) * a.W + xlo;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.in + (base & ~static_cast<size_t>(3)));
        const int words = (static_cast<int>(base & 3) + cols + 3) >> 2;
        for (int w = lane; w < words; w += 32) s_in[r * a.max_words + w] = __ldg(src + w);
        if (lane == 0) s_off[r] = static_cast<int>(base & 3) - xlo;
    }
    __syncthreads();
    // horizontal pass: thread = one output column (weights in registers when KSH > 0), rows 2 apart
    {
        const int xx = tid & (kRzTW - 1);
        const int2 b = s_hb[xx];
        const int *k = s_hk + xx * ksh;
        int kr[KSH ? KSH : 1];
        if constexpr (KSH > 0) {
#pragma unroll
            for (int x = 0; x < KSH; ++x) kr[x] = k[x];
        }
        if (xx < nx) {
            for (int r = tid / kRzTW; r < rows; r += kRzThreads / kRzTW) {
                const uint8_t *row = reinterpret_cast<const uint8_t *>(s_in + r * a.max_words) + s_off[r] + b.x;
                int acc = 1 << (kResizePrecisionBits - 1);
                if constexpr (KSH > 0) {
#pragma unroll
                    for (int x = 0; x < KSH; ++x) acc += static_cast<int>(row[x]) * kr[x];
                } else {
                    for (int x = 0; x < b.y; ++x) acc += static_cast<int>(row[x]) * k[x];
                }
                s_tmp[r * kRzTW + xx] = static_cast<uint8_t>(resize_clip8(acc));
            }
        }
    }
    __syncthreads();
    // vertical pass: thread = 4 adjacent columns, rows 8 apart
    {
        const int g4 = 4 * (tid & 31);
        if (g4 < nx) {
            for (int yy = tid >> 5; yy < ny; yy += kRzThreads / 32) {
                const int2 b = s_vb[yy];
                const int *k = s_vk + yy * ksv;
                const uint8_t *col = s_tmp + (b.x - ylo) * kRzTW + g4;
                int acc[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] = 1 << (kResizePrecisionBits - 1);
                auto tap = [&](int y) {
                    const uint32_t v = *reinterpret_cast<const uint32_t *>(col + y * kRzTW);
                    const int w = k[y];
                    acc[0] += static_cast<int>(v & 0xffu) * w, acc[1] += static_cast<int>((v >> 8) & 0xffu) * w;
                    acc[2] += static_cast<int>((v >> 16) & 0xffu) * w, acc[3] += static_cast<int>(v >> 24) * w;
                };
                if constexpr (KSV > 0) {
#pragma unroll
                    for (int y = 0; y < KSV; ++y) tap(y);
                } else {
                    for (int y = 0; y < b.y; ++y) tap(y);
                }
                uint8_t *dst = a.out + (p * a.Hout + y0 + yy) * a.Wout + x0 + g4;
                if (g4 + 3 < nx && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
                    *reinterpret_cast<uint32_t *>(dst) =
                        resize_clip8(acc[0]) | (resize_clip8(acc[1]) << 8) | (resize_clip8(acc[2]) << 16) | (resize_clip8(acc[3]) << 24);
                } else {
                    for (int e = 0; e < 4 && g4 + e < nx; ++e) dst[e] = static_cast<uint8_t>(resize_clip8(acc[e]));
                }
            }
        }
    }
}

using resize_fused_fn = void (*)(const ResizeFusedArgs);
static resize_fused_fn resize_pick_fused(int ksh, int ksv) {
    if (ksh == 5 && ksv == 5) return resize_fused_kernel<5, 5>;      // fix_size: bicubic, both axes grow
    if (ksh == 1 && ksv == 5) return resize_fused_kernel<1, 5>;      // one axis unchanged (identity window)
    if (ksh == 5 && ksv == 1) return resize_fused_kernel<5, 1>;
    return resize_fused_kernel<0, 0>;
}

// upper bounds of the input window of any tile: max over tiles of (last.x + last.y - first.x)
static int resize_max_window(const std::vector<int> &bounds, int out_size, int tile) {
    int m = 1;
    for (int t0 = 0; t0 < out_size; t0 += tile) {
        const int t1 = (t0 + tile < out_size ? t0 + tile : out_size) - 1;
        const int span = bounds[2 * t1] + bounds[2 * t1 + 1] - bounds[2 * t0];
        if (span > m) m = span;
    }
    return m;
}
static void resize_identity(int size, std::vector<int> &bounds, std::vector<int> &kk) {
    bounds.resize(static_cast<size_t>(size) * 2);
    kk.assign(size, 1 << kResizePrecisionBits);
    for (int i = 0; i < size; ++i) bounds[2 * i] = i, bounds[2 * i + 1] = 1;
}

struct ResizeLayout {
    int ksh, ksv;
    size_t off_tmp, off_hb, off_hk, off_vb, off_vk, bytes;
};
static ResizeLayout resize_layout(long long planes, int H, int W, int Hout, int Wout, int filter) {
    ResizeLayout l{};
    const bool need_h = Wout != W, need_v = Hout != H;
    l.ksh = need_h ? resize_ksize(W, Wout, filter) : 1;         // an unchanged axis runs with identity tables in the fused form
    l.ksv = need_v ? resize_ksize(H, Hout, filter) : 1;
    size_t o = 0;
    auto take = [&](size_t n) {
        const size_t at = o;
        o = round_up<size_t>(o + n, 256);
        return at;
    };
    l.off_tmp = take(need_h && need_v ? static_cast<size_t>(planes) * H * Wout : 0);      // two-kernel form only
    l.off_hb = take(static_cast<size_t>(Wout) * 8);
    l.off_hk = take(static_cast<size_t>(Wout) * l.ksh * 4);
    l.off_vb = take(static_cast<size_t>(Hout) * 8);
    l.off_vk = take(static_cast<size_t>(Hout) * l.ksv * 4);
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
    std::vector<int> hbnd, hkk, vbnd, vkk;
    if (need_h) resize_coeffs(W, Wout, filter, l.ksh, hbnd, hkk); else resize_identity(W, hbnd, hkk);
    if (need_v) resize_coeffs(H, Hout, filter, l.ksv, vbnd, vkk); else resize_identity(H, vbnd, vkk);
    // one H2D copy for the four tables (they are adjacent in the workspace); a pageable-source cudaMemcpyAsync stages
    // the host data before it returns, so the staging vector may die at scope exit
    {
        std::vector<unsigned char> stage(l.bytes - l.off_hb, 0);
        std::memcpy(stage.data(), hbnd.data(), hbnd.size() * 4);
        std::memcpy(stage.data() + (l.off_hk - l.off_hb), hkk.data(), hkk.size() * 4);
        std::memcpy(stage.data() + (l.off_vb - l.off_hb), vbnd.data(), vbnd.size() * 4);
        std::memcpy(stage.data() + (l.off_vk - l.off_hb), vkk.data(), vkk.size() * 4);
        B200_CUDA_TRY(cudaMemcpyAsync(w + l.off_hb, stage.data(), stage.size(), cudaMemcpyHostToDevice, st));
    }
    {   // fused form whenever the largest tile window fits shared memory and the input allows aligned 32-bit loads
        ResizeFusedArgs fa;
        fa.in = in, fa.out = out;
        fa.hb = reinterpret_cast<const int2 *>(w + l.off_hb), fa.vb = reinterpret_cast<const int2 *>(w + l.off_vb);
        fa.hk = reinterpret_cast<const int *>(w + l.off_hk), fa.vk = reinterpret_cast<const int *>(w + l.off_vk);
        fa.H = H, fa.W = W, fa.Hout = Hout, fa.Wout = Wout, fa.ksh = l.ksh, fa.ksv = l.ksv;
        // fixed-length windows read up to ksize - 1 samples past a window's own count (zero weights): pad rows and columns
        fa.max_rows = resize_max_window(vbnd, Hout, kRzTH);
        fa.max_words = (resize_max_window(hbnd, Wout, kRzTW) + 3 + 3 + l.ksh) / 4 + 1;
        const size_t smem = static_cast<size_t>(fa.max_rows) * fa.max_words * 4 + static_cast<size_t>(fa.max_rows + l.ksv) * kRzTW +
                            static_cast<size_t>(kRzTW) * l.ksh * 4 + static_cast<size_t>(kRzTH) * l.ksv * 4 + (kRzTW + kRzTH) * 8 +
                            static_cast<size_t>(fa.max_rows) * 4;
        const char *force = std::getenv("B200_RESIZE_FUSED");
        const bool want = !(force && force[0] == '0');
        if (want && smem <= 160 * 1024 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 && ceil_div(Hout, kRzTH) <= 65535) {
            resize_fused_fn fn = resize_pick_fused(l.ksh, l.ksv);
            B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(smem)));
            for (long long p0 = 0; p0 < planes; p0 += 65535) {
                const unsigned np = static_cast<unsigned>(planes - p0 < 65535 ? planes - p0 : 65535);
                fa.plane0 = p0;
                fn<<<dim3(ceil_div(Wout, kRzTW), ceil_div(Hout, kRzTH), np), kRzThreads, smem, st>>>(fa);
                B200_LAUNCH_CHECK("resize_fused_kernel");
            }
            return B200_OK;
        }
    }
    const uint8_t *src = in;
    if (need_h) {
        uint8_t *dst = need_v ? w + l.off_tmp : out;
        const long long rows = planes * H;
        const long long blocks = ceil_div<long long>(rows, kResizeRows);
        if (blocks > 0x7fffffffLL) return B200_ERR_UNSUPPORTED;
        resize_h_kernel<<<static_cast<unsigned>(blocks), resize_block(Wout), 0, st>>>(
            src, dst, rows, W, Wout, l.ksh, reinterpret_cast<const int2 *>(w + l.off_hb), reinterpret_cast<const int *>(w + l.off_hk));
        B200_LAUNCH_CHECK("resize_h_kernel");
        src = dst;
    }
    if (need_v) {
        const size_t in_plane = static_cast<size_t>(H) * Wout, out_plane = static_cast<size_t>(Hout) * Wout;
        for (long long p0 = 0; p0 < planes; p0 += 65535) {
            const unsigned np = static_cast<unsigned>(planes - p0 < 65535 ? planes - p0 : 65535);
            resize_v_kernel<<<dim3(ceil_div(Hout, kResizeRows), np), resize_block(Wout), 0, st>>>(
                src + p0 * in_plane, out + p0 * out_plane, H, Hout, Wout, l.ksv, reinterpret_cast<const int2 *>(w + l.off_vb),
                reinterpret_cast<const int *>(w + l.off_vk));
            B200_LAUNCH_CHECK("resize_v_kernel");
        }
    }
    return B200_OK;
}

}  // extern "C"
