// the other Hamming metrics on the same scan.
//
//   radius_counts_kernel   <- DSCH pr_curve                          /root/reference/main/engine/DSCH/_utils.py:467-492
//                             get_precision_recall_by_Hamming_Radius  _utils.py:577-594
//                             Both only need, per query and radius r, (#rows, #relevant rows) with distance <= r: the
//                             running sum over distance of the stage-A shard totals (no second pass over the database,
//                             no [Q, N] float matrix).
//   ranked_cumhits_kernel  <- CustomCalculator.calculate_pr_rc_hashing   main/engine/accuracy_calculator.py:235-273
//                             (gnd[argsort(hamm)] -> cumsum) and DSCH p_topK (_utils.py:495-512, gnd[sort(hamm)[:K]].sum()):
//                             relevance along the ranked list as a running hit count.  One warp per query, 32 ranks at a
//                             time: packed-label AND != 0 (or equality) per lane, ballot -> prefix count.
//   curve_accumulate_kernel<- the tail of calculate_pr_rc_hashing (:251-265): sum over the selected queries of
//                             float32(cum / rank) and float32(cum / total) per rank, one thread per rank (coalesced
//                             over ranks, queries in index order => bit-reproducible), accumulated into float64 so that
//                             the caller can stream the queries in chunks.
#include "common.cuh"
#include "hamming_core.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) radius_counts_kernel(const U32x2 *__restrict__ tot, int bins, int Qpad, int Q,
                                                            U32x2 *__restrict__ cum) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    uint32_t rows = 0, rel = 0;
    for (int d = 0; d < bins; ++d) {
        const U32x2 c = tot[static_cast<size_t>(d) * Qpad + q];
        rows += c.x, rel += c.y;
        cum[static_cast<size_t>(d) * Q + q] = U32x2{rows, rel};
    }
}

template <int LW, bool EQ>
__global__ void __launch_bounds__(256) ranked_cumhits_kernel(const uint32_t *__restrict__ idx, int Q, long long k,
                                                             const uint64_t *__restrict__ q_labels,
                                                             const uint64_t *__restrict__ db_labels,
                                                             uint32_t *__restrict__ cum) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    uint64_t ql[LW];
#pragma unroll
    for (int i = 0; i < LW; ++i) ql[i] = q_labels[static_cast<size_t>(q) * LW + i];
    const uint32_t *row = idx + static_cast<size_t>(q) * k;
    uint32_t *out = cum + static_cast<size_t>(q) * k;
    uint32_t hits = 0;
    for (long long base = 0; base < k; base += 32) {
        const long long p = base + lane;
        bool rel = false;
        if (p < k) {
            const uint32_t id = row[p];
            if (id != 0xffffffffu) {
                const uint64_t *l = db_labels + static_cast<size_t>(id) * LW;
                if (EQ) {
                    rel = __ldg(l) == ql[0];
                } else {
                    uint64_t any = 0;
#pragma unroll
                    for (int i = 0; i < LW; ++i) any |= __ldg(l + i) & ql[i];
                    rel = any != 0;
                }
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, rel);
        if (p < k) out[p] = hits + __popc(ballot & (0xffffffffu >> (31 - lane)));
        hits += __popc(ballot);
    }
}

__global__ void __launch_bounds__(256) curve_accumulate_kernel(const uint32_t *__restrict__ cum, int Q, long long k,
                                                               const uint8_t *__restrict__ mask, double *__restrict__ prec,
                                                               double *__restrict__ rec, uint32_t *__restrict__ n_used) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= k) return;
    const float rank = static_cast<float>(p + 1);
    double sp = 0.0, sr = 0.0;
    uint32_t used = 0;
    for (int q = 0; q < Q; ++q) {
        const uint32_t total = __ldg(cum + static_cast<size_t>(q) * k + (k - 1));      // same address for the whole warp
        if (total == 0 || (mask && !mask[q])) continue;
        const float c = static_cast<float>(cum[static_cast<size_t>(q) * k + p]);
        sp += static_cast<double>(__fdiv_rn(c, rank));                                  // prec_sum / return_images  (:255)
        sr += static_cast<double>(__fdiv_rn(c, static_cast<float>(total)));             // prec_sum / all_sim_num    (:256)
        ++used;
    }
    prec[p] += sp, rec[p] += sr;
    if (p == 0) n_used[0] += used;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_hamming_radius_counts(const b200_map_plan *plan, const void *workspace, uint32_t *cum, b200_stream_t stream) {
    if (!plan || !workspace || !cum || plan->Q < 1 || plan->bins != plan->B + 1) return B200_ERR_INVALID_ARG;
    const unsigned char *w = static_cast<const unsigned char *>(workspace);
    radius_counts_kernel<<<ceil_div(plan->Q, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const U32x2 *>(w + plan->off_tot), plan->bins, plan->Qpad, plan->Q, reinterpret_cast<U32x2 *>(cum));
    B200_LAUNCH_CHECK("radius_counts_kernel");
    return B200_OK;
}

int b200_ranked_cumhits(const uint32_t *idx, int Q, long long k, const uint64_t *q_labels, const uint64_t *db_labels, int LW,
                        int label_mode, uint32_t *cum, b200_stream_t stream) {
    if (Q < 1 || k < 1 || !idx || !q_labels || !db_labels || !cum) return B200_ERR_INVALID_ARG;
    if (label_mode != B200_LABELS_OVERLAP && label_mode != B200_LABELS_EQUAL) return B200_ERR_INVALID_ARG;
    const bool eq = label_mode == B200_LABELS_EQUAL;
    if (eq && LW != 1) return B200_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    const int grid = ceil_div(Q, 8);
    if (eq) {
        ranked_cumhits_kernel<1, true><<<grid, 256, 0, st>>>(idx, Q, k, q_labels, db_labels, cum);
    } else if (LW == 1) {
        ranked_cumhits_kernel<1, false><<<grid, 256, 0, st>>>(idx, Q, k, q_labels, db_labels, cum);
    } else if (LW == 2) {
        ranked_cumhits_kernel<2, false><<<grid, 256, 0, st>>>(idx, Q, k, q_labels, db_labels, cum);
    } else if (LW == 4) {
        ranked_cumhits_kernel<4, false><<<grid, 256, 0, st>>>(idx, Q, k, q_labels, db_labels, cum);
    } else {
        return B200_ERR_UNSUPPORTED;
    }
    B200_LAUNCH_CHECK("ranked_cumhits_kernel");
    return B200_OK;
}

int b200_curve_accumulate(const uint32_t *cum, int Q, long long k, const uint8_t *query_mask, double *prec_sum,
                          double *rec_sum, uint32_t *n_used, b200_stream_t stream) {
    if (Q < 1 || k < 1 || !cum || !prec_sum || !rec_sum || !n_used) return B200_ERR_INVALID_ARG;
    const long long grid = ceil_div<long long>(k, 256);
    if (grid > 0x7fffffffLL) return B200_ERR_UNSUPPORTED;
    curve_accumulate_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(cum, Q, k, query_mask, prec_sum, rec_sum,
                                                                                        n_used);
    B200_LAUNCH_CHECK("curve_accumulate_kernel");
    return B200_OK;
}

}  // extern "C"
