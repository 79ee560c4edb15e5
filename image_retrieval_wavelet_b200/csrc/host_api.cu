// Host-buffer entry points: what a caller without CUDA buffers binds (the reference's evaluator works on CPU
// tensors — main/engine/evaluate.py:42,79 — and its transform runs on a PIL image in a DataLoader worker —
// main/transforms/custom_transforms.py:145-157).  Each call stages host memory to the device on a private stream
// with stream-ordered allocations (pool-cached after the first call), runs the same kernels as the device API and
// copies the results back.  Copies are DMA'd at full PCIe rate when the caller's buffers are pinned.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "hamming_plan.h"

namespace b200 {

int swt2_launch(const void *in, int in_is_u8, float *out, int B, int C, int H, int W, const float *lo, const float *hi, int F,
                int level, cudaStream_t st);
// hamming_select.cu / hamming_map.cu: the select pipeline in phases (a database that arrives in row chunks)
int select_begin(const b200_map_plan *p, const uint64_t *qc, const uint64_t *dc, const uint64_t *sample_codes, void *ws, cudaStream_t st);
int select_segments(const b200_map_plan *p, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc, const uint64_t *dl, void *ws,
                    uint32_t *status, int seg0, int seg1, cudaStream_t st);
int hamming_map_after_select(const b200_map_plan *plan, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc, const uint64_t *dl,
                             void *ws, double *ap, uint32_t *tsum, double *map_out, uint32_t *status, cudaStream_t st);

static cudaStream_t copy_stream() {
    static thread_local cudaStream_t st = nullptr;
    if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) st = nullptr;
    return st;
}

static cudaStream_t host_stream() {
    static thread_local cudaStream_t st = nullptr;
    if (!st) {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
            st = nullptr;
            return st;
        }
        // keep freed staging buffers in the device's default pool across calls: with the default release threshold (0)
        // every synchronising call hands its ~100 MB of staging back to the driver and the next call maps it again
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return st;
}

struct AsyncArena {                    // frees everything it handed out, stream-ordered, on scope exit
    cudaStream_t st;
    void *ptrs[24];
    int n = 0;
    explicit AsyncArena(cudaStream_t s) : st(s) {}
    ~AsyncArena() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st);
    }
    template <typename T>
    int get(T **p, size_t count) {
        void *q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, (count ? count : 1) * sizeof(T), st);
        if (e != cudaSuccess) {
            set_last_cuda_error(e, "cudaMallocAsync");
            return B200_ERR_CUDA;
        }
        ptrs[n++] = q;
        *p = static_cast<T *>(q);
        return B200_OK;
    }
};

// [H][W][C] uint8 (np.array(PIL image)) -> planar [C][H][W]
__global__ void __launch_bounds__(256) hwc_to_chw_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int H, int W,
                                                         int C) {
    const long long px = static_cast<long long>(H) * W;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < px * C;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long p = i / C;
        const int c = static_cast<int>(i - p * C);
        out[c * px + p] = in[i];
    }
}

}  // namespace b200

using namespace b200;

#define B200_TRY(expr)             \
    do {                           \
        int _rc = (expr);          \
        if (_rc != B200_OK) return _rc; \
    } while (0)

extern "C" {

int b200_swt2_fwd_host(const void *in_host, int in_is_u8, int in_is_hwc, float *out_host, int B, int C, int H, int W,
                       const float *dec_lo, const float *dec_hi, int F, int level) {
    if (!in_host || !out_host || !dec_lo || !dec_hi || B < 1 || C < 1 || H < 1 || W < 1) return B200_ERR_INVALID_ARG;
    if (F < 2 || (F & 1) || level < 1 || H % (1 << level) || W % (1 << level)) return B200_ERR_INVALID_ARG;
    if (F > B200_SWT_MAX_FILTER || level > B200_SWT_MAX_LEVEL) return B200_ERR_UNSUPPORTED;
    if (in_is_hwc && (!in_is_u8 || B != 1)) return B200_ERR_INVALID_ARG;
    cudaStream_t st = host_stream();
    if (!st) return B200_ERR_NO_DEVICE;
    AsyncArena arena(st);
    const size_t px = static_cast<size_t>(B) * C * H * W;
    const size_t in_bytes = px * (in_is_u8 ? 1 : 4);
    uint8_t *d_in = nullptr, *d_planar = nullptr;
    float *d_out = nullptr;
    B200_TRY(arena.get(&d_in, in_bytes + 16));
    B200_TRY(arena.get(&d_out, px * 4));
    B200_CUDA_TRY(cudaMemcpyAsync(d_in, in_host, in_bytes, cudaMemcpyHostToDevice, st));
    const void *src = d_in;
    if (in_is_hwc) {
        B200_TRY(arena.get(&d_planar, in_bytes + 16));
        hwc_to_chw_kernel<<<static_cast<unsigned>(ceil_div<size_t>(px, 256)), 256, 0, st>>>(d_in, d_planar, H, W, C);
        B200_LAUNCH_CHECK("hwc_to_chw_kernel");
        src = d_planar;
    }
    B200_TRY(swt2_launch(src, in_is_u8, d_out, B, C, H, W, dec_lo, dec_hi, F, level, st));
    B200_CUDA_TRY(cudaMemcpyAsync(out_host, d_out, px * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200_OK;
}

int b200_maphashing_host(const float *q_codes, const float *q_labels, const float *db_codes, const float *db_labels, int Q,
                         long long N, int B, int L, int label_mode, long long k, double *ap_out, uint32_t *tsum_out,
                         double *map_out, int *n_invalid) {
    if (!q_codes || !q_labels || !map_out || Q < 1 || N < 0 || B < 1 || L < 1 || k < 1) return B200_ERR_INVALID_ARG;
    if (N > 0 && (!db_codes || !db_labels)) return B200_ERR_INVALID_ARG;
    if (label_mode != B200_LABELS_OVERLAP && label_mode != B200_LABELS_EQUAL) return B200_ERR_INVALID_ARG;
    if (label_mode == B200_LABELS_EQUAL && L != 1) return B200_ERR_INVALID_ARG;
    if (B > B200_MAX_CODE_BITS || L > B200_MAX_LABEL_BITS) return B200_ERR_UNSUPPORTED;
    if (n_invalid) *n_invalid = 0;
    if (N == 0) {                       // empty database: every AP is 0 (accuracy_calculator.py:226)
        if (ap_out) for (int i = 0; i < Q; ++i) ap_out[i] = 0.0;
        if (tsum_out) for (int i = 0; i < Q; ++i) tsum_out[i] = 0;
        *map_out = 0.0;
        return B200_OK;
    }
    cudaStream_t st = host_stream();
    if (!st) return B200_ERR_NO_DEVICE;
    b200_map_plan plan;
    const int cw = b200_code_words(B);
    const int lw = label_mode == B200_LABELS_EQUAL ? 1 : b200_label_words(L);
    constexpr int kMaxChunks = 8;
    // The database crosses PCIe in row chunks on a second stream; chunk c is bit-packed while chunk c+1 is still in
    // flight (pinned caller buffers make the copies asynchronous).  Select plans are STREAMED: the sample rows (every
    // stride-th 32-row group, codes only: one strided 2-D copy of a few MB) cross first and give the bound; every chunk's
    // segments are then scored by the select kernel as soon as the chunk is packed, so that behind the last byte of the
    // transfer only one chunk's select pass, the rank kernel and the mean are left.  The three stages (no select plan)
    // start when the whole packed database is there.
    const size_t h2d_bytes = static_cast<size_t>(N) * (static_cast<size_t>(B) + L) * sizeof(float);
    int kChunks = h2d_bytes >= (64u << 20) ? 8 : (h2d_bytes >= (16u << 20) ? 4 : 1);
    if (const char *e = getenv("B200_HOST_CHUNKS")) kChunks = atoi(e) < 1 ? 1 : (atoi(e) > kMaxChunks ? kMaxChunks : atoi(e));
    const bool trace = getenv("B200_HOST_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    B200_TRY(b200_map_plan_init(&plan, Q, N, N, B, lw, label_mode, k));
    AsyncArena arena(st);
    float *f_qc, *f_ql, *f_dc, *f_dl;
    uint64_t *p_qc, *p_ql, *p_dc, *p_dl;
    unsigned char *ws;
    double *d_ap, *d_map;
    uint32_t *d_tsum;
    int *d_bad;
    const size_t Qp = round_up<size_t>(Q, 2), Np = round_up<size_t>(N, 2);
    B200_TRY(arena.get(&f_qc, static_cast<size_t>(Q) * B));
    B200_TRY(arena.get(&f_ql, static_cast<size_t>(Q) * L));
    B200_TRY(arena.get(&f_dc, static_cast<size_t>(N) * B));
    B200_TRY(arena.get(&f_dl, static_cast<size_t>(N) * L));
    B200_TRY(arena.get(&p_qc, Qp * cw));
    B200_TRY(arena.get(&p_ql, Qp * lw));
    B200_TRY(arena.get(&p_dc, Np * cw));
    B200_TRY(arena.get(&p_dl, Np * lw));
    B200_TRY(arena.get(&ws, plan.workspace_bytes));
    B200_TRY(arena.get(&d_ap, static_cast<size_t>(Q)));
    B200_TRY(arena.get(&d_tsum, static_cast<size_t>(Q)));
    B200_TRY(arena.get(&d_map, 1));
    B200_TRY(arena.get(&d_bad, 4));                            // [0..1] invalid entries, [2] status of the optimistic select round
    B200_CUDA_TRY(cudaMemsetAsync(d_bad, 0, 4 * sizeof(int), st));
    uint32_t *d_status = reinterpret_cast<uint32_t *>(d_bad + 2);
    B200_CUDA_TRY(cudaMemcpyAsync(f_qc, q_codes, sizeof(float) * Q * B, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(f_ql, q_labels, sizeof(float) * Q * L, cudaMemcpyHostToDevice, st));
    B200_TRY(b200_pack_codes(f_qc, Q, B, p_qc, d_bad, st));
    if (label_mode == B200_LABELS_EQUAL)
        B200_TRY(b200_pack_labels_scalar(f_ql, 0, Q, p_ql, d_bad + 1, st));
    else
        B200_TRY(b200_pack_labels(f_ql, Q, L, p_ql, d_bad + 1, st));
    cudaStream_t cs = copy_stream();
    if (!cs) return B200_ERR_NO_DEVICE;
    cudaEvent_t ready[kMaxChunks], start, smp_ready;
    bool streamed = plan.select != 0 && kChunks > 1;
    if (const char *e = getenv("B200_HOST_STREAMED")) streamed = streamed && atoi(e) != 0;       // A/B
    long long per = round_up<long long>(ceil_div<long long>(N, kChunks), 2);      // even: packed chunks stay 16-byte aligned
    if (streamed) per = round_up<long long>(per, plan.sel_seg_len);              // whole segments per chunk (sel_seg_len is even)
    const int chunks = static_cast<int>(ceil_div<long long>(N, per));
    float *f_smp = nullptr;
    uint64_t *p_smp = nullptr;
    if (streamed) {
        B200_TRY(arena.get(&f_smp, static_cast<size_t>(plan.smp_rows) * B));
        B200_TRY(arena.get(&p_smp, round_up<size_t>(plan.smp_rows, 2) * cw));
    }
    B200_CUDA_TRY(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    for (int c = 0; c < chunks; ++c) B200_CUDA_TRY(cudaEventCreateWithFlags(&ready[c], cudaEventDisableTiming));
    B200_CUDA_TRY(cudaEventRecord(start, st));                 // the staging buffers are this call's once `st` gets here
    B200_CUDA_TRY(cudaStreamWaitEvent(cs, start, 0));
    int rc = B200_OK;
    if (streamed) {
        // sample row r = database row 32 stride (r / 32) + r % 32: smp_rows / 32 pieces of 32 rows, 32 stride rows apart
        B200_CUDA_TRY(cudaEventCreateWithFlags(&smp_ready, cudaEventDisableTiming));
        const size_t piece = static_cast<size_t>(32) * B * sizeof(float);
        cudaMemcpy2DAsync(f_smp, piece, db_codes, piece * plan.sel_stride, piece, static_cast<size_t>(plan.smp_rows / 32),
                          cudaMemcpyHostToDevice, cs);
        cudaEventRecord(smp_ready, cs);
        cudaStreamWaitEvent(st, smp_ready, 0);
        rc = b200_pack_codes(f_smp, plan.smp_rows, B, p_smp, nullptr, st);
        if (rc == B200_OK) rc = select_begin(&plan, p_qc, p_dc, p_smp, ws, st);
        cudaEventDestroy(smp_ready);
    }
    int seg_done = 0;
    for (int c = 0; c < chunks && rc == B200_OK; ++c) {
        const long long r0 = c * per, r1 = r0 + per < N ? r0 + per : N, rows = r1 - r0;
        cudaMemcpyAsync(f_dc + r0 * B, db_codes + r0 * B, sizeof(float) * rows * B, cudaMemcpyHostToDevice, cs);
        cudaMemcpyAsync(f_dl + r0 * L, db_labels + r0 * L, sizeof(float) * rows * L, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(ready[c], cs);
        cudaStreamWaitEvent(st, ready[c], 0);
        rc = b200_pack_codes(f_dc + r0 * B, rows, B, p_dc + r0 * cw, d_bad, st);
        if (rc == B200_OK)
            rc = label_mode == B200_LABELS_EQUAL ? b200_pack_labels_scalar(f_dl + r0 * L, 0, rows, p_dl + r0 * lw, d_bad + 1, st)
                                                 : b200_pack_labels(f_dl + r0 * L, rows, L, p_dl + r0 * lw, d_bad + 1, st);
        if (streamed && rc == B200_OK) {                       // the segments this chunk completed
            const int seg_end = r1 == N ? plan.sel_S : static_cast<int>(r1 / plan.sel_seg_len);
            rc = select_segments(&plan, p_qc, p_ql, p_dc, p_dl, ws, d_status, seg_done, seg_end, st);
            seg_done = seg_end;
        }
    }
    if (rc == B200_OK)
        rc = streamed ? hamming_map_after_select(&plan, p_qc, p_ql, p_dc, p_dl, ws, d_ap, d_tsum, d_map, d_status, st)
                      : b200_hamming_map(&plan, p_qc, p_ql, p_dc, p_dl, ws, d_ap, d_tsum, d_map, st);
    if (rc != B200_OK) cudaStreamSynchronize(cs);              // the arena frees on `st`: no copy may still be in flight
    cudaEventDestroy(start);
    for (int c = 0; c < chunks; ++c) cudaEventDestroy(ready[c]);
    if (rc != B200_OK) return rc;
    const auto t_enqueued = std::chrono::steady_clock::now();
    int bad[4] = {0, 0, 0, 0};
    B200_CUDA_TRY(cudaMemcpyAsync(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaMemcpyAsync(map_out, d_map, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (ap_out) B200_CUDA_TRY(cudaMemcpyAsync(ap_out, d_ap, sizeof(double) * Q, cudaMemcpyDeviceToHost, st));
    if (tsum_out) B200_CUDA_TRY(cudaMemcpyAsync(tsum_out, d_tsum, sizeof(uint32_t) * Q, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    if (streamed && bad[2] && !(bad[0] || bad[1])) {
        // the optimistic round was not enough (a list shorter than k, or the pool overflowed): the complete sequence on
        // the packed database, which is all on the device by now
        B200_TRY(b200_hamming_map(&plan, p_qc, p_ql, p_dc, p_dl, ws, d_ap, d_tsum, d_map, st));
        B200_CUDA_TRY(cudaMemcpyAsync(map_out, d_map, sizeof(double), cudaMemcpyDeviceToHost, st));
        if (ap_out) B200_CUDA_TRY(cudaMemcpyAsync(ap_out, d_ap, sizeof(double) * Q, cudaMemcpyDeviceToHost, st));
        if (tsum_out) B200_CUDA_TRY(cudaMemcpyAsync(tsum_out, d_tsum, sizeof(uint32_t) * Q, cudaMemcpyDeviceToHost, st));
        B200_CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (trace) {
        const auto t_done = std::chrono::steady_clock::now();
        fprintf(stderr, "b200_maphashing_host: chunks=%d select=%d stash=%d enqueue %.3f ms, total %.3f ms\n", chunks, plan.select,
                plan.stash, std::chrono::duration<double, std::milli>(t_enqueued - t_begin).count(),
                std::chrono::duration<double, std::milli>(t_done - t_begin).count());
    }
    if (bad[0] || bad[1]) {
        if (n_invalid) *n_invalid = bad[0] + bad[1];
        return B200_ERR_INVALID_ARG;
    }
    return B200_OK;
}

// The same evaluation for a caller that keeps PACKED codes and labels (b200_pack_* layout) in host memory: what crosses
// PCIe is 24-48 bytes per row instead of 4 (B + L) — the hand-over format once the evaluator glue packs on the device
// right behind the model (main/engine/evaluate.py:26-64 keeps float32 codes on the CPU today).
int b200_maphashing_host_packed(const uint64_t *q_codes, const uint64_t *q_labels, const uint64_t *db_codes,
                                const uint64_t *db_labels, int Q, long long N, int B, int LW, int label_mode, long long k,
                                double *ap_out, uint32_t *tsum_out, double *map_out) {
    if (!q_codes || !q_labels || !map_out || Q < 1 || N < 0 || B < 1 || k < 1) return B200_ERR_INVALID_ARG;
    if (N > 0 && (!db_codes || !db_labels)) return B200_ERR_INVALID_ARG;
    if (label_mode != B200_LABELS_OVERLAP && label_mode != B200_LABELS_EQUAL) return B200_ERR_INVALID_ARG;
    if (B > B200_MAX_CODE_BITS || !(LW == 1 || LW == 2 || LW == 4) || (label_mode == B200_LABELS_EQUAL && LW != 1)) return B200_ERR_UNSUPPORTED;
    if (N == 0) {
        if (ap_out) for (int i = 0; i < Q; ++i) ap_out[i] = 0.0;
        if (tsum_out) for (int i = 0; i < Q; ++i) tsum_out[i] = 0;
        *map_out = 0.0;
        return B200_OK;
    }
    cudaStream_t st = host_stream();
    if (!st) return B200_ERR_NO_DEVICE;
    b200_map_plan plan;
    const int cw = b200_code_words(B);
    B200_TRY(b200_map_plan_init(&plan, Q, N, N, B, LW, label_mode, k));
    AsyncArena arena(st);
    uint64_t *p_qc, *p_ql, *p_dc, *p_dl;
    unsigned char *ws;
    double *d_ap, *d_map;
    uint32_t *d_tsum;
    const size_t Qp = round_up<size_t>(Q, 2), Np = round_up<size_t>(N, 2);
    B200_TRY(arena.get(&p_qc, Qp * cw));
    B200_TRY(arena.get(&p_ql, Qp * LW));
    B200_TRY(arena.get(&p_dc, Np * cw));
    B200_TRY(arena.get(&p_dl, Np * LW));
    B200_TRY(arena.get(&ws, plan.workspace_bytes));
    B200_TRY(arena.get(&d_ap, static_cast<size_t>(Q)));
    B200_TRY(arena.get(&d_tsum, static_cast<size_t>(Q)));
    B200_TRY(arena.get(&d_map, 1));
    // the kernels read whole 16-byte units: the padding row of an odd row count must exist (and be zero) on the device
    B200_CUDA_TRY(cudaMemsetAsync(p_qc + (Qp - 1) * cw, 0, sizeof(uint64_t) * cw, st));
    B200_CUDA_TRY(cudaMemsetAsync(p_ql + (Qp - 1) * LW, 0, sizeof(uint64_t) * LW, st));
    B200_CUDA_TRY(cudaMemsetAsync(p_dc + (Np - 1) * cw, 0, sizeof(uint64_t) * cw, st));
    B200_CUDA_TRY(cudaMemsetAsync(p_dl + (Np - 1) * LW, 0, sizeof(uint64_t) * LW, st));
    B200_CUDA_TRY(cudaMemcpyAsync(p_qc, q_codes, sizeof(uint64_t) * Q * cw, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(p_ql, q_labels, sizeof(uint64_t) * Q * LW, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(p_dc, db_codes, sizeof(uint64_t) * N * cw, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(p_dl, db_labels, sizeof(uint64_t) * N * LW, cudaMemcpyHostToDevice, st));
    // select plans: the optimistic round first (b200_hamming_map_try), the complete sequence only if its status word is set
    uint32_t *d_status = nullptr, status = 0;
    if (plan.select) {
        B200_TRY(arena.get(&d_status, 4));
        B200_CUDA_TRY(cudaMemsetAsync(d_status, 0, 4 * sizeof(uint32_t), st));
    }
    for (int pass = 0; pass < 2; ++pass) {
        if (d_status && pass == 0) {
            B200_TRY(b200_hamming_map_try(&plan, p_qc, p_ql, p_dc, p_dl, ws, d_ap, d_tsum, d_status, st));
            B200_TRY(b200_mean_f64(d_ap, nullptr, Q, d_map, st));
            B200_CUDA_TRY(cudaMemcpyAsync(&status, d_status, sizeof(status), cudaMemcpyDeviceToHost, st));
        } else {
            B200_TRY(b200_hamming_map(&plan, p_qc, p_ql, p_dc, p_dl, ws, d_ap, d_tsum, d_map, st));
        }
        B200_CUDA_TRY(cudaMemcpyAsync(map_out, d_map, sizeof(double), cudaMemcpyDeviceToHost, st));
        if (ap_out) B200_CUDA_TRY(cudaMemcpyAsync(ap_out, d_ap, sizeof(double) * Q, cudaMemcpyDeviceToHost, st));
        if (tsum_out) B200_CUDA_TRY(cudaMemcpyAsync(tsum_out, d_tsum, sizeof(uint32_t) * Q, cudaMemcpyDeviceToHost, st));
        B200_CUDA_TRY(cudaStreamSynchronize(st));
        if (!(d_status && pass == 0 && status != 0)) break;
    }
    return B200_OK;
}

}  // extern "C"
