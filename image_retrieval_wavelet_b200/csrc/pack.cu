// Bit packing of hash codes / label sets and the per-bit population counts.
//
// Reference hand-over format (what these kernels consume): float32 +-1 codes [N][B] and float32 multi-hot labels
// [N][L] as assembled by compute_all_embeddings, /root/reference/main/engine/evaluate.py:26-64, i.e. the
// arguments of calc_hamming_dist (accuracy_calculator.py:183-186) and label_comparison_fn (:31-37).
//
// HBM-bound streaming kernels: one warp turns 64 consecutive floats (two coalesced 128-byte reads) into one
// uint64 with two ballots.  Algorithmic bytes per packed row: B*4 read + ceil(B/64)*8 written.
#include <cstdlib>

#include "common.cuh"

#ifndef B200_PACK_KU
#define B200_PACK_KU 16
#endif

namespace b200 {

enum class PackMode { kCodes, kLabels };

// every packed word goes to all n destinations: one for a local buffer, one per rank when the packed shard is written
// straight into every peer's copy of the database (comm.cu) — pack + all-gather in one pass over the floats
struct PackDst {
    uint64_t *p[B200_COMM_MAX_RANKS];
    int n;
};

// A warp owns 32 consecutive (row, word) tasks per batch: lane j keeps the word of task j, so the batch leaves as ONE
// coalesced 256-byte store per destination (peers over NVLink get 256-byte writes instead of 8-byte ones) and the
// 64-bit division task -> (row, word) is paid once per batch, not per task (round 2, first form: one task per warp and
// iteration — the division made the kernel issue-bound at 1.6-2.3 TB/s on the 1 M-row database of c5).
template <PackMode MODE>
__global__ void __launch_bounds__(256) pack_rows_kernel(const float *__restrict__ src, long long rows, int cols, int words,
                                                        long long rows_padded, const PackDst dst,
                                                        int *__restrict__ n_invalid) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long tasks = rows_padded * words;
    const long long batches = (tasks + 31) >> 5;
    int bad = 0;
    // kU tasks per round: all their loads are issued before the first ballot (the kernel is a pure stream of 4-byte
    // reads; one task at a time leaves a single load in flight per lane)
    constexpr int kU = 16;
    for (long long b = warp; b < batches; b += nwarps) {
        const long long t0 = b << 5;
        long long row = t0 / words;
        int w = static_cast<int>(t0 - row * words);
        uint64_t mine = 0ull;
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += kU) {
            float x[kU][2];
            bool in[kU][2];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const float *prow = src + row * cols;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = w * 64 + h * 32 + lane;
                    in[u][h] = row < rows && col < cols;
                    x[u][h] = in[u][h] ? __ldg(prow + col) : 0.f;
                }
                if (++w == words) w = 0, ++row;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
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
                    bad += ok ? 0 : 1;
                }
                if (lane == j0 + u) mine = (static_cast<uint64_t>(half[1]) << 32) | half[0];
            }
        }
        // rows past `rows` (the padding rows and what lies past the last task) are zero words; stores to peers are
        // fire-and-forget over NVLink
        if (t0 + lane < tasks)
            for (int r = 0; r < dst.n; ++r) dst.p[r][t0 + lane] = mine;
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (lane == 0 && bad && n_invalid) atomicAdd(n_invalid, bad);
}

// 128-bit form (cols % 4 == 0, cols <= 128, 16-byte aligned source — every bench / reference shape): a row is read as
// float4s by G = 8 / 16 / 32 adjacent lanes (32 / G rows per warp-wide load), each lane turns its 4 values into a
// nibble, an 8-lane OR butterfly assembles 32 consecutive bits, and two shuffles hand word j of the batch to lane j — about
// 32 instructions per 320-512 bytes instead of ~100 per 256 bytes (the scalar kernel above stays issue-bound below half of
// the HBM rate; it remains the general path for odd widths, cols > 128 and unaligned sources).
template <PackMode MODE, int G>
__global__ void __launch_bounds__(256) pack_rows_v4_kernel(const float *__restrict__ src, long long rows, int cols, long long rows_padded,
                                                           const PackDst dst, int *__restrict__ n_invalid) {
    constexpr int R = 32 / G;                        // rows per load
    constexpr int W = G == 32 ? 2 : 1;               // words per row
    constexpr int TPL = R * W;                       // (row, word) tasks per load: 2, 2, 4
    constexpr int LOADS = 32 / TPL;                  // loads per batch of 32 tasks: 16, 16, 8
    constexpr int kRowsPerBatch = LOADS * R;
    constexpr int kU = B200_PACK_KU < LOADS ? B200_PACK_KU : LOADS;      // loads in flight per lane
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long tasks = rows_padded * W;
    const long long batches = (rows_padded + kRowsPerBatch - 1) / kRowsPerBatch;
    const int sub = lane / G, c4 = lane % G, c4s = cols >> 2;
    const bool col_ok = c4 < c4s;
    // word j of a batch = load i = j / TPL, 8-lane groups 2 (j % TPL) [low half] and 2 (j % TPL) + 1 [high half] (G = 8:
    // group j % 4, no high half).  Lane s keeps the piece of load i when s % 8 == i % 8 (one register per 8 loads), so
    // the whole batch is handed over with 2 shuffles per 8 loads instead of 2 per load.
    const int want = lane / TPL;                                     // the load whose word this lane stores
    const int src_lo = (G == 8 ? (lane & 3) * 8 : (lane & 1) * 16) + (want & 7);
    int bad = 0;
    for (long long b = warp; b < batches; b += nwarps) {
        const long long row0 = b * kRowsPerBatch;
        const float4 *base = reinterpret_cast<const float4 *>(src + row0 * cols);
        const int left = static_cast<int>(rows - row0 < kRowsPerBatch ? rows - row0 : kRowsPerBatch);      // rows of this batch that exist
        uint32_t keep[LOADS / 8];                                       // [i / 8]: the piece of load i in the lanes with lane % 8 == i % 8
#pragma unroll
        for (int k = 0; k < LOADS / 8; ++k) keep[k] = 0u;
        uint32_t off = 0u;                                              // != 0 (sign bit aside): an invalid entry in this lane's loads
#pragma unroll
        for (int k = 0; k < LOADS / kU; ++k) {
            float4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int rr = (k * kU + u) * R + sub;
                v[u] = __ldg(base + ((col_ok && rr < left) ? rr * c4s + c4 : 0));      // (the batch's first float4 stands in: always readable)
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                uint32_t nib = 0u, o = 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // x x - 1 == 0 <=> x in {+1, -1};  x x - x == 0 <=> x in {0, 1}  (inf, NaN: never 0)
                    if (MODE == PackMode::kCodes) {
                        nib |= e[c] > 0.f ? 1u << c : 0u;
                        o |= __float_as_uint(fmaf(e[c], e[c], -1.f));
                    } else {
                        nib |= e[c] != 0.f ? 1u << c : 0u;
                        o |= __float_as_uint(fmaf(e[c], e[c], -e[c]));
                    }
                }
                if (!(col_ok && (k * kU + u) * R + sub < left)) nib = 0u, o = 0u;      // a stand-in value
                off |= o;
                uint32_t piece = nib << (4 * (lane & 7));                // OR over the 8 lanes of a group: 32 consecutive bits
                piece |= __shfl_xor_sync(0xffffffffu, piece, 1);          // (REDUX with a per-group mask compiles to a loop
                piece |= __shfl_xor_sync(0xffffffffu, piece, 2);          //  over the distinct masks)
                piece |= __shfl_xor_sync(0xffffffffu, piece, 4);
                const int i = k * kU + u;
                if ((lane & 7) == (i & 7)) keep[i >> 3] = piece;
            }
        }
        if (__any_sync(0xffffffffu, (off & 0x7fffffffu) != 0u)) {      // rare: read the batch again and count the entries exactly
            for (int i = 0; i < LOADS; ++i) {
                const int rr = i * R + sub;
                if (!(col_ok && rr < left)) continue;
                const float4 x = __ldg(base + rr * c4s + c4);
                const float e[4] = {x.x, x.y, x.z, x.w};
                for (int c = 0; c < 4; ++c) {
                    if (MODE == PackMode::kCodes)
                        bad += fabsf(e[c]) == 1.f ? 0 : 1;
                    else
                        bad += (e[c] == 0.f || e[c] == 1.f) ? 0 : 1;
                }
            }
        }
        uint32_t mlo = 0u, mhi = 0u;
#pragma unroll
        for (int k = 0; k < LOADS / 8; ++k) {
            const uint32_t lo = __shfl_sync(0xffffffffu, keep[k], src_lo);
            const uint32_t hi = G == 8 ? 0u : __shfl_sync(0xffffffffu, keep[k], src_lo + 8);
            if ((want >> 3) == k) mlo = lo, mhi = hi;
        }
        const long long t0 = row0 * W;
        if (t0 + lane < tasks) {
            const uint64_t word = (static_cast<uint64_t>(mhi) << 32) | mlo;
            for (int r = 0; r < dst.n; ++r) dst.p[r][t0 + lane] = word;
        }
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
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

template <PackMode MODE>
static void pack_rows_v4_launch(const float *src, long long N, int cols, long long padded, const PackDst &dst, int *n_invalid, cudaStream_t st) {
    if (cols > 64) {
        pack_rows_v4_kernel<MODE, 32><<<pack_grid(ceil_div<long long>(padded, 16)), 256, 0, st>>>(src, N, cols, padded, dst, n_invalid);
    } else if (cols > 32) {
        pack_rows_v4_kernel<MODE, 16><<<pack_grid(ceil_div<long long>(padded, 32)), 256, 0, st>>>(src, N, cols, padded, dst, n_invalid);
    } else {
        pack_rows_v4_kernel<MODE, 8><<<pack_grid(ceil_div<long long>(padded, 32)), 256, 0, st>>>(src, N, cols, padded, dst, n_invalid);
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

static int pack_rows_launch(bool codes, const float *src, long long N, int cols, const PackDst &dst, int *n_invalid, cudaStream_t st) {
    const int words = codes ? b200_code_words(cols) : b200_label_words(cols);
    const long long padded = round_up<long long>(N, 2);
    const char *v4_env = std::getenv("B200_PACK_V4");         // A/B and tests: 0 = the scalar kernel everywhere
    const bool scalar_only = v4_env && v4_env[0] == '0';
    if (!scalar_only && cols % 4 == 0 && cols <= 128 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        if (codes)
            pack_rows_v4_launch<PackMode::kCodes>(src, N, cols, padded, dst, n_invalid, st);
        else
            pack_rows_v4_launch<PackMode::kLabels>(src, N, cols, padded, dst, n_invalid, st);
    } else if (codes) {
        pack_rows_kernel<PackMode::kCodes><<<pack_grid(ceil_div<long long>(padded * words, 32)), 256, 0, st>>>(src, N, cols, words, padded, dst, n_invalid);
    } else {
        pack_rows_kernel<PackMode::kLabels><<<pack_grid(ceil_div<long long>(padded * words, 32)), 256, 0, st>>>(src, N, cols, words, padded, dst, n_invalid);
    }
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
