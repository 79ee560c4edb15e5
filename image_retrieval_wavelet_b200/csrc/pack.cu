// Bit packing of hash codes / label sets and the per-bit population counts.
//
// Reference hand-over format (what these kernels consume): float32 +-1 codes [N][B] and float32 multi-hot labels
// [N][L] as assembled by compute_all_embeddings, /root/reference/main/engine/evaluate.py:26-64, i.e. the
// arguments of calc_hamming_dist (accuracy_calculator.py:183-186) and label_comparison_fn (:31-37).
//
// HBM-bound streaming kernels: one warp turns 64 consecutive floats (two coalesced 128-byte reads) into one
// uint64 with two ballots.  Algorithmic bytes per packed row: B*4 read + ceil(B/64)*8 written.
#include "common.cuh"

namespace b200 {

enum class PackMode { kCodes, kLabels };

// every packed word goes to all n destinations: one for a local buffer, one per rank when the packed shard is written
// straight into every peer's copy of the database (comm.cu) — pack + all-gather in one pass over the floats
struct PackDst {
    uint64_t *p[B200_COMM_MAX_RANKS];
    int n;
};

template <PackMode MODE>
__global__ void __launch_bounds__(256) pack_rows_kernel(const float *__restrict__ src, long long rows, int cols, int words,
                                                        long long rows_padded, const PackDst dst,
                                                        int *__restrict__ n_invalid) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long tasks = rows_padded * words;
    int bad = 0;
    // kU (row, word) tasks per iteration: all their loads are issued before the first ballot (the kernel is a pure
    // stream of 4-byte reads; one task at a time leaves a single load in flight per lane)
    constexpr int kU = 4;
    for (long long t0 = warp; t0 < tasks; t0 += nwarps * kU) {
        float x[kU][2];
        bool in[kU][2];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const long long t = t0 + static_cast<long long>(u) * nwarps;
            const long long row = t / words;
            const int w = static_cast<int>(t - row * words);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = w * 64 + h * 32 + lane;
                in[u][h] = t < tasks && row < rows && col < cols;
                x[u][h] = in[u][h] ? __ldg(src + row * cols + col) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const long long t = t0 + static_cast<long long>(u) * nwarps;
            if (t >= tasks) break;                       // warp-uniform
            uint32_t half[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float v = x[u][h];
                bool bit, ok;
                if (MODE == PackMode::kCodes) {
                    bit = v > 0.f;
                    ok = !in[u][h] || v == 1.f || v == -1.f;
                } else {
                    bit = v != 0.f;
                    ok = !in[u][h] || v == 0.f || v == 1.f;
                }
                half[h] = __ballot_sync(0xffffffffu, bit && in[u][h]);
                bad += __popc(__ballot_sync(0xffffffffu, !ok));
            }
            // lane r stores to destination r (stores to peers are fire-and-forget over NVLink)
            if (lane < dst.n) dst.p[lane][t] = (static_cast<uint64_t>(half[1]) << 32) | half[0];
        }
    }
    if (lane == 0 && bad && n_invalid) atomicAdd(n_invalid, bad);
}

// 1-D labels: canonical 64-bit pattern so that equal values <=> equal words, whatever the dtype they arrive in (the
// reference's `==` promotes): integral values map to their two's-complement int64, other floats to the bits of the
// double (-0.0 folded onto +0.0); NaN is invalid.  KIND: 0 float32, 1 int64, 2 float64.
template <int KIND>
__global__ void __launch_bounds__(256) pack_scalar_labels_kernel(const void *__restrict__ src, long long rows,
                                                                 long long rows_padded, uint64_t *__restrict__ dst,
                                                                 int *__restrict__ n_invalid) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows_padded; i += stride) {
        uint64_t v = 0ull;    // padding row (never read as a label)
        if (i < rows) {
            if (KIND == 1) {
                v = static_cast<uint64_t>(static_cast<const long long *>(src)[i]);
            } else {
                const double x = KIND == 2 ? static_cast<const double *>(src)[i]
                                           : static_cast<double>(static_cast<const float *>(src)[i]);
                if (x != x) {
                    if (n_invalid) atomicAdd(n_invalid, 1);
                }
                if (x == rint(x) && fabs(x) < 9.2e18)
                    v = static_cast<uint64_t>(static_cast<long long>(x));
                else
                    v = static_cast<uint64_t>(__double_as_longlong(x + 0.0));
            }
        }
        dst[i] = v;
    }
}

// ones[b] = number of rows whose bit b is set.  Each warp owns a contiguous slab of rows; lane l counts bits l and
// l+32 of every word (the word is a warp-wide broadcast load), block-level smem reduction, integer atomics.
__global__ void __launch_bounds__(256) bit_counts_kernel(const uint64_t *__restrict__ codes, long long rows, int words,
                                                         int bits, uint32_t *__restrict__ ones) {
    __shared__ uint32_t acc[B200_MAX_CODE_BITS];
    for (int i = threadIdx.x; i < B200_MAX_CODE_BITS; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long per = ceil_div(rows, nwarps);
    const long long r0 = warp * per;
    const long long r1 = r0 + per < rows ? r0 + per : rows;
    for (int w = 0; w < words; ++w) {
        uint32_t lo = 0, hi = 0;
        for (long long r = r0; r < r1; ++r) {
            const uint64_t x = __ldg(codes + r * words + w);
            lo += static_cast<uint32_t>(x >> lane) & 1u;
            hi += static_cast<uint32_t>(x >> (lane + 32)) & 1u;
        }
        if (lo) atomicAdd(&acc[w * 64 + lane], lo);
        if (hi) atomicAdd(&acc[w * 64 + 32 + lane], hi);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bits; i += blockDim.x)
        if (acc[i]) atomicAdd(&ones[i], acc[i]);
}

// Dense utilities (API parity with calc_hamming_dist :183-186 and label_comparison_fn :31-37; the fused evaluator
// never materialises these matrices).  One thread per (query, row) pair, rows fastest => coalesced stores.
template <int CW>
__global__ void __launch_bounds__(256) hamming_dist_kernel(const uint64_t *__restrict__ q, const uint64_t *__restrict__ db, int Q,
                                                           long long N, float *__restrict__ out) {
    const long long total = static_cast<long long>(Q) * N;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long qi = i / N, n = i - qi * N;
        int d = 0;
#pragma unroll
        for (int w = 0; w < CW; ++w) d += __popcll(q[qi * CW + w] ^ __ldg(db + n * CW + w));
        out[i] = static_cast<float>(d);
    }
}
template <int LW, bool EQ>
__global__ void __launch_bounds__(256) label_rel_kernel(const uint64_t *__restrict__ q, const uint64_t *__restrict__ db, int Q,
                                                        long long N, uint8_t *__restrict__ out) {
    const long long total = static_cast<long long>(Q) * N;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long qi = i / N, n = i - qi * N;
        bool rel;
        if (EQ) {
            rel = q[qi] == __ldg(db + n);
        } else {
            uint64_t any = 0;
#pragma unroll
            for (int w = 0; w < LW; ++w) any |= q[qi * LW + w] & __ldg(db + n * LW + w);
            rel = any != 0;
        }
        out[i] = rel;
    }
}

static int pack_grid(long long tasks_in_warps) {
    const long long blocks = ceil_div<long long>(tasks_in_warps, 8);
    const long long cap = static_cast<long long>(sm_count()) * 8;
    return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace b200

using namespace b200;

extern "C" {

static int pack_rows_launch(bool codes, const float *src, long long N, int cols, const PackDst &dst, int *n_invalid, cudaStream_t st) {
    const int words = codes ? b200_code_words(cols) : b200_label_words(cols);
    const long long padded = round_up<long long>(N, 2);
    if (codes)
        pack_rows_kernel<PackMode::kCodes><<<pack_grid(padded * words), 256, 0, st>>>(src, N, cols, words, padded, dst, n_invalid);
    else
        pack_rows_kernel<PackMode::kLabels><<<pack_grid(padded * words), 256, 0, st>>>(src, N, cols, words, padded, dst, n_invalid);
    B200_LAUNCH_CHECK(codes ? "pack_codes" : "pack_labels");
    return B200_OK;
}

int b200_pack_codes(const float *codes, long long N, int B, uint64_t *packed, int *n_invalid, b200_stream_t stream) {
    if (N < 0 || B < 1 || (N > 0 && (!codes || !packed))) return B200_ERR_INVALID_ARG;
    if (B > B200_MAX_CODE_BITS) return B200_ERR_UNSUPPORTED;
    if (N == 0) return B200_OK;
    PackDst dst = {};
    dst.p[0] = packed, dst.n = 1;
    return pack_rows_launch(true, codes, N, B, dst, n_invalid, as_stream(stream));
}

int b200_pack_labels(const float *labels, long long N, int L, uint64_t *packed, int *n_invalid, b200_stream_t stream) {
    if (N < 0 || L < 1 || (N > 0 && (!labels || !packed))) return B200_ERR_INVALID_ARG;
    if (L > B200_MAX_LABEL_BITS) return B200_ERR_UNSUPPORTED;
    if (N == 0) return B200_OK;
    PackDst dst = {};
    dst.p[0] = packed, dst.n = 1;
    return pack_rows_launch(false, labels, N, L, dst, n_invalid, as_stream(stream));
}

// Pack this rank's shard and write it at byte offset dst_offset of EVERY rank's exchange region (b200_comm_*).
int b200_pack_to_ranks(const float *src, int is_codes, long long N, int cols, b200_comm *comm, size_t dst_offset, int *n_invalid,
                       b200_stream_t stream) {
    if (!comm || N < 0 || cols < 1 || (N > 0 && !src) || (dst_offset & 15)) return B200_ERR_INVALID_ARG;
    if (cols > (is_codes ? B200_MAX_CODE_BITS : B200_MAX_LABEL_BITS)) return B200_ERR_UNSUPPORTED;
    if (N == 0) return B200_OK;
    const int words = is_codes ? b200_code_words(cols) : b200_label_words(cols);
    if (dst_offset + static_cast<size_t>(round_up<long long>(N, 2)) * words * 8 > b200_comm_bytes(comm)) return B200_ERR_INVALID_ARG;
    PackDst dst = {};
    dst.n = b200_comm_world(comm);
    for (int r = 0; r < dst.n; ++r) {
        unsigned char *base = static_cast<unsigned char *>(b200_comm_buffer(comm, r));
        if (!base) return B200_ERR_INVALID_ARG;
        dst.p[r] = reinterpret_cast<uint64_t *>(base + dst_offset);
    }
    return pack_rows_launch(is_codes != 0, src, N, cols, dst, n_invalid, as_stream(stream));
}

int b200_pack_labels_scalar(const void *labels, int is_int64, long long N, uint64_t *packed, int *n_invalid,
                            b200_stream_t stream) {
    if (N < 0 || (N > 0 && (!labels || !packed))) return B200_ERR_INVALID_ARG;
    if (N == 0) return B200_OK;
    const long long padded = round_up<long long>(N, 2);
    const int grid = pack_grid(ceil_div<long long>(padded, 32));
    if (is_int64 == 1)
        pack_scalar_labels_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(labels, N, padded, packed, n_invalid);
    else if (is_int64 == 2)
        pack_scalar_labels_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(labels, N, padded, packed, n_invalid);
    else if (is_int64 == 0)
        pack_scalar_labels_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(labels, N, padded, packed, n_invalid);
    else
        return B200_ERR_INVALID_ARG;
    B200_LAUNCH_CHECK("pack_labels_scalar");
    return B200_OK;
}

int b200_bit_counts(const uint64_t *packed_codes, long long N, int B, uint32_t *ones, b200_stream_t stream) {
    if (N < 0 || B < 1 || !ones || (N > 0 && !packed_codes)) return B200_ERR_INVALID_ARG;
    if (B > B200_MAX_CODE_BITS) return B200_ERR_UNSUPPORTED;
    B200_CUDA_TRY(cudaMemsetAsync(ones, 0, sizeof(uint32_t) * B, as_stream(stream)));
    if (N == 0) return B200_OK;
    const int grid = pack_grid(ceil_div<long long>(N, 64));
    bit_counts_kernel<<<grid, 256, 0, as_stream(stream)>>>(packed_codes, N, b200_code_words(B), B, ones);
    B200_LAUNCH_CHECK("bit_counts");
    return B200_OK;
}

int b200_hamming_dist(const uint64_t *q_codes, const uint64_t *db_codes, int Q, long long N, int B, float *dist,
                      b200_stream_t stream) {
    if (Q < 0 || N < 0 || B < 1 || (Q > 0 && N > 0 && (!q_codes || !db_codes || !dist))) return B200_ERR_INVALID_ARG;
    if (B > B200_MAX_CODE_BITS) return B200_ERR_UNSUPPORTED;
    if (Q == 0 || N == 0) return B200_OK;
    const int grid = pack_grid(ceil_div<long long>(static_cast<long long>(Q) * N, 32));
    switch (b200_code_words(B)) {
        case 1: hamming_dist_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(q_codes, db_codes, Q, N, dist); break;
        case 2: hamming_dist_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(q_codes, db_codes, Q, N, dist); break;
        default: hamming_dist_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(q_codes, db_codes, Q, N, dist); break;
    }
    B200_LAUNCH_CHECK("hamming_dist_kernel");
    return B200_OK;
}

int b200_label_relevance(const uint64_t *q_labels, const uint64_t *db_labels, int Q, long long N, int LW, int label_mode,
                         uint8_t *rel, b200_stream_t stream) {
    if (Q < 0 || N < 0 || (Q > 0 && N > 0 && (!q_labels || !db_labels || !rel))) return B200_ERR_INVALID_ARG;
    if (label_mode != B200_LABELS_OVERLAP && label_mode != B200_LABELS_EQUAL) return B200_ERR_INVALID_ARG;
    if (Q == 0 || N == 0) return B200_OK;
    const int grid = pack_grid(ceil_div<long long>(static_cast<long long>(Q) * N, 32));
    cudaStream_t st = as_stream(stream);
    if (label_mode == B200_LABELS_EQUAL) {
        if (LW != 1) return B200_ERR_INVALID_ARG;
        label_rel_kernel<1, true><<<grid, 256, 0, st>>>(q_labels, db_labels, Q, N, rel);
    } else if (LW == 1) {
        label_rel_kernel<1, false><<<grid, 256, 0, st>>>(q_labels, db_labels, Q, N, rel);
    } else if (LW == 2) {
        label_rel_kernel<2, false><<<grid, 256, 0, st>>>(q_labels, db_labels, Q, N, rel);
    } else if (LW == 4) {
        label_rel_kernel<4, false><<<grid, 256, 0, st>>>(q_labels, db_labels, Q, N, rel);
    } else {
        return B200_ERR_UNSUPPORTED;
    }
    B200_LAUNCH_CHECK("label_rel_kernel");
    return B200_OK;
}

}  // extern "C"
