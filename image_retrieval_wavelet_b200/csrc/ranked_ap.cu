// K3: relevance + AP over a ranked list, and the k-way merge of per-shard ranked lists.
//
//   ranked_ap_kernel  <- CustomCalculator.calculate_map, /root/reference/main/engine/accuracy_calculator.py:156-167
//                        (label_comparison_fn(query_labels[:, None], knn_labels) + torchmetrics RetrievalMAP over the
//                        knn list: AP = mean over hits of hit-ordinal / rank, 0 for a query without hits), and the
//                        tail of calculate_maphashing (:221-229) when the ranked list is materialised.
//   merge_topk_kernel <- the host-side merge faiss does for a database sharded over all GPUs
//                        (/root/reference/main/engine/get_knn.py:41-44, co.shards = True).
//
// ranked_ap: one warp per query walks the list 32 ranks at a time: packed-label AND != 0 (or equality) per lane, a
// ballot gives the segmented prefix count of hits, each hit adds ordinal/rank.  HBM traffic per query: k indices
// (4 or 8 B) + k gathered label rows (LW*8 B, random access into an L2-resident table).
#include "common.cuh"

namespace b200 {

int launch_mean(const double *ap, const uint8_t *mask, int Q, double *out, cudaStream_t st);

template <int LW, bool EQ, bool IDX64>
__global__ void __launch_bounds__(256) ranked_ap_kernel(const void *__restrict__ idx_, int Q, long long k,
                                                        const uint64_t *__restrict__ q_labels,
                                                        const uint64_t *__restrict__ db_labels, double *__restrict__ ap,
                                                        uint32_t *__restrict__ hits_out) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    uint64_t ql[LW];
#pragma unroll
    for (int i = 0; i < LW; ++i) ql[i] = q_labels[static_cast<size_t>(q) * LW + i];
    uint32_t hits = 0;
    double sum = 0.0;
    for (long long base = 0; base < k; base += 32) {
        const long long p = base + lane;
        bool rel = false;
        if (p < k) {
            long long id;
            if (IDX64) {
                id = static_cast<const long long *>(idx_)[static_cast<size_t>(q) * k + p];
            } else {
                const uint32_t v = static_cast<const uint32_t *>(idx_)[static_cast<size_t>(q) * k + p];
                id = v == 0xffffffffu ? -1 : static_cast<long long>(v);
            }
            if (id >= 0) {
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
        if (rel) {
            const uint32_t ordinal = hits + __popc(ballot & (0xffffffffu >> (31 - lane)));
            sum += static_cast<double>(__fdiv_rn(static_cast<float>(ordinal), static_cast<float>(p + 1)));
        }
        hits += __popc(ballot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) {
        ap[q] = hits ? sum / static_cast<double>(hits) : 0.0;
        if (hits_out) hits_out[q] = hits;
    }
}

// One CTA per query.  Every shard list is sorted by (distance, index) and shards hold ascending index ranges, so
// the merged order is (distance, shard, position): a counting merge — run boundaries give the per-(shard,
// distance) counts, a scan over (distance, shard) gives each run's output offset, every element is scattered once.
__global__ void __launch_bounds__(512) merge_topk_kernel(const uint32_t *__restrict__ in_idx,
                                                         const uint16_t *__restrict__ in_dist, int n_shards, int Q,
                                                         long long k, int bins, uint32_t *__restrict__ out_idx,
                                                         uint16_t *__restrict__ out_dist) {
    extern __shared__ uint32_t s_merge[];
    uint32_t *s_start = s_merge;                         // [n_shards][bins] first position of the run
    uint32_t *s_end = s_start + n_shards * bins;         // [n_shards][bins] one past the run
    uint32_t *s_base = s_end + n_shards * bins;          // [n_shards][bins] output offset of the run
    uint32_t *s_tot = s_base + n_shards * bins;          // [512] scan scratch over distances
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * n_shards * bins; i += blockDim.x) s_merge[i] = 0;
    __syncthreads();
    const size_t plane = static_cast<size_t>(Q) * k;
    for (int r = 0; r < n_shards; ++r) {
        const uint16_t *dist = in_dist + r * plane + static_cast<size_t>(q) * k;
        for (long long p = tid; p < k; p += blockDim.x) {
            const uint16_t d = dist[p];
            const uint16_t prev = p ? dist[p - 1] : 0xffff;
            if (p == 0 || prev != d) {
                if (d != 0xffff) s_start[r * bins + d] = static_cast<uint32_t>(p);
                if (p && prev != 0xffff) s_end[r * bins + prev] = static_cast<uint32_t>(p);
            }
            if (p == k - 1 && d != 0xffff) s_end[r * bins + d] = static_cast<uint32_t>(k);
        }
    }
    __syncthreads();
    // exclusive scan over distances of the all-shard bucket sizes (bins <= 257 <= 512 threads)
    uint32_t mine = 0;
    if (tid < bins)
        for (int r = 0; r < n_shards; ++r) mine += s_end[r * bins + tid] - s_start[r * bins + tid];
    s_tot[tid] = mine;
    __syncthreads();
    for (int off = 1; off < 512; off <<= 1) {
        const uint32_t v = tid >= off ? s_tot[tid - off] : 0;
        __syncthreads();
        s_tot[tid] += v;
        __syncthreads();
    }
    if (tid < bins) {
        uint32_t run = s_tot[tid] - mine;
        for (int r = 0; r < n_shards; ++r) {
            s_base[r * bins + tid] = run;
            run += s_end[r * bins + tid] - s_start[r * bins + tid];
        }
    }
    const uint32_t total = s_tot[511];
    __syncthreads();
    for (int r = 0; r < n_shards; ++r) {
        const uint16_t *dist = in_dist + r * plane + static_cast<size_t>(q) * k;
        const uint32_t *idx = in_idx + r * plane + static_cast<size_t>(q) * k;
        for (long long p = tid; p < k; p += blockDim.x) {
            const uint16_t d = dist[p];
            if (d == 0xffff) continue;
            const uint32_t o = s_base[r * bins + d] + (static_cast<uint32_t>(p) - s_start[r * bins + d]);
            if (o < k) {
                out_idx[static_cast<size_t>(q) * k + o] = idx[p];
                out_dist[static_cast<size_t>(q) * k + o] = d;
            }
        }
    }
    for (long long p = total + tid; p < k; p += blockDim.x) {
        out_idx[static_cast<size_t>(q) * k + p] = 0xffffffffu;
        out_dist[static_cast<size_t>(q) * k + p] = 0xffff;
    }
}

template <bool IDX64>
static int launch_ranked(const void *idx, int Q, long long k, const uint64_t *ql, const uint64_t *dl, int LW, bool eq,
                         double *ap, uint32_t *hits, cudaStream_t st) {
    const int grid = ceil_div(Q, 8);
    if (eq) {
        ranked_ap_kernel<1, true, IDX64><<<grid, 256, 0, st>>>(idx, Q, k, ql, dl, ap, hits);
    } else if (LW == 1) {
        ranked_ap_kernel<1, false, IDX64><<<grid, 256, 0, st>>>(idx, Q, k, ql, dl, ap, hits);
    } else if (LW == 2) {
        ranked_ap_kernel<2, false, IDX64><<<grid, 256, 0, st>>>(idx, Q, k, ql, dl, ap, hits);
    } else if (LW == 4) {
        ranked_ap_kernel<4, false, IDX64><<<grid, 256, 0, st>>>(idx, Q, k, ql, dl, ap, hits);
    } else {
        return B200_ERR_UNSUPPORTED;
    }
    B200_LAUNCH_CHECK("ranked_ap_kernel");
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_ranked_ap(const void *idx, int is_int64, int Q, long long k, const uint64_t *q_labels, const uint64_t *db_labels,
                   int LW, int label_mode, const uint8_t *query_mask, double *ap, uint32_t *hits, double *map_out,
                   b200_stream_t stream) {
    if (Q < 1 || k < 0 || !ap || !q_labels || (k > 0 && (!idx || !db_labels))) return B200_ERR_INVALID_ARG;
    if (label_mode != B200_LABELS_OVERLAP && label_mode != B200_LABELS_EQUAL) return B200_ERR_INVALID_ARG;
    const bool eq = label_mode == B200_LABELS_EQUAL;
    if (eq && LW != 1) return B200_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    int rc = is_int64 ? launch_ranked<true>(idx, Q, k, q_labels, db_labels, LW, eq, ap, hits, st)
                      : launch_ranked<false>(idx, Q, k, q_labels, db_labels, LW, eq, ap, hits, st);
    if (rc) return rc;
    if (map_out) return launch_mean(ap, query_mask, Q, map_out, st);
    return B200_OK;
}

int b200_merge_topk(const uint32_t *in_idx, const uint16_t *in_dist, int n_shards, int Q, long long k, int B,
                    uint32_t *out_idx, uint16_t *out_dist, b200_stream_t stream) {
    if (n_shards < 1 || Q < 1 || k < 1 || B < 1 || !in_idx || !in_dist || !out_idx || !out_dist) return B200_ERR_INVALID_ARG;
    if (B > B200_MAX_CODE_BITS || n_shards > 64) return B200_ERR_UNSUPPORTED;
    const int bins = B + 1;
    const size_t smem = (static_cast<size_t>(3) * n_shards * bins + 512) * sizeof(uint32_t);
    if (smem > 227 * 1024) return B200_ERR_UNSUPPORTED;
    B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(merge_topk_kernel),
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    merge_topk_kernel<<<Q, 512, smem, as_stream(stream)>>>(in_idx, in_dist, n_shards, Q, k, bins, out_idx, out_dist);
    B200_LAUNCH_CHECK("merge_topk_kernel");
    return B200_OK;
}

}  // extern "C"
