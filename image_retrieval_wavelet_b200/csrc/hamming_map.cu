// Hamming mAP@k as a counting sort — the B200 replacement of the per-query Python loop
//   CustomCalculator.calculate_maphashing   /root/reference/main/engine/accuracy_calculator.py:203-231
//   (calc_hamming_dist :183-186, label_comparison_fn :31-37, torch.argsort :220, AP :223-229).
//
// Why a counting sort.  For +-1 codes of B bits the distance takes only B+1 values, so "argsort, take the first k,
// average hit-ordinal/rank over the hits" never needs a comparison sort: the rank of database row j for query q is
//       rank(q,j) = #{i : d(q,i) < d(q,j)}  +  #{i <= j : d(q,i) = d(q,j)}            (index tie-break)
// and the hit ordinal is the same expression restricted to relevant rows.  Both are prefix counts:
//   stage A  hist[s][d][q]  = (#rows, #relevant rows) of segment s at distance d          (hamming_hist_kernel)
//   stage S  exclusive scan over (d, shard, s)  -> rank / ordinal base of every (s, d, q)  (hamming_scan_kernel)
//   stage B  re-walk each segment in index order, bump the per-distance counters from their bases and add
//            ordinal/rank for every relevant row with rank <= k                            (hamming_ap_kernel)
//
// Mapping.  One THREAD owns one query; the CTA's threads walk the same segment of the database in index order, so a
// database row is one shared-memory broadcast read for the whole warp, the query code/labels live in registers and
// the per-distance counters are a private shared-memory column cnt[d][t] (bank = t mod 32: conflict-free).  The
// packed database (8..32 B/row) is L2-resident; the kernels are bound by the INT/POPC issue rate, not by HBM
// (DESIGN.md §4).
//
// Counter width.  Narrow mode (k <= 65534, segment <= 65534 rows) packs (rows | relevant << 16) in one uint32 so a
// row costs one LDS + one STS; wide mode uses uint64 (rows | relevant << 32).
#include "common.cuh"
#include "hamming_core.cuh"
#include "hamming_plan.h"

namespace b200 {

// hamming_select.cu
constexpr int kSelFlagFallback = 1;
int hamming_select_run(const b200_map_plan *p, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc, const uint64_t *dl,
                       void *ws, double *ap, uint32_t *tsum, uint32_t *rank_idx, uint16_t *rank_dist, uint32_t *status,
                       cudaStream_t st);
int select_finish(const b200_map_plan *p, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc, const uint64_t *dl, void *ws,
                  double *ap, uint32_t *tsum, uint32_t *rank_idx, uint16_t *rank_dist, uint32_t *status, bool round0_selected,
                  cudaStream_t st);

struct MapDeviceExec {
    template <typename Fn>
    __device__ __forceinline__ void operator()(Fn fn) const {
        fn(static_cast<int>(threadIdx.y * blockDim.x + threadIdx.x), static_cast<int>(blockDim.x * blockDim.y));
        __syncthreads();
    }
    template <typename State>
    __device__ __forceinline__ State *state(State *local) const { return local; }
    __device__ __forceinline__ int slot(int) const { return 0; }
};

// 16-byte streaming copy of a database tile into shared memory
struct DevLoadTile {
    __device__ __forceinline__ void operator()(uint32_t *dst, const uint64_t *src, int n16, int t, int nt) const {
        const uint4 *g = reinterpret_cast<const uint4 *>(src);
        for (int i = t; i < n16; i += nt) reinterpret_cast<uint4 *>(dst)[i] = ldg_stream_u4(g + i);
    }
};

template <int CW, int LW, bool EQ, bool WIDE>
__global__ void hamming_hist_kernel(const __grid_constant__ MapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.gate && *a.gate == 0u) return;
    hamming_hist_program<CW, LW, EQ, WIDE>(a, blockIdx.x, a.seg_base + blockIdx.y, blockDim.x, smem_raw, MapDeviceExec{}, DevLoadTile{});
}

template <int CW, int LW, bool EQ, bool WIDE, bool ALL>
__global__ void hamming_walk_kernel(const __grid_constant__ MapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.gate && *a.gate == 0u) return;
    hamming_walk_program<CW, LW, EQ, WIDE, ALL>(a, blockIdx.x, blockIdx.y, blockDim.x, smem_raw, MapDeviceExec{}, DevLoadTile{});
}

template <bool WIDE, bool ALL>
__global__ void hamming_rank_kernel(const __grid_constant__ MapArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.gate && *a.gate == 0u) return;
    hamming_rank_program<WIDE, ALL>(a, blockIdx.x, blockIdx.y, blockDim.x, smem_raw, MapDeviceExec{});
}

template <bool WIDE>
__global__ void __launch_bounds__(256) hamming_totals_kernel(const void *hist, int S, int bins, int Qpad, U32x2 *tot) {
    const size_t n = static_cast<size_t>(bins) * Qpad;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        hamming_totals_item<WIDE>(hist, S, n, i, tot);
}

template <bool WIDE>
__global__ void __launch_bounds__(kScanQ *kScanY) hamming_scan_kernel(void *hist, int S, int bins, int Qpad, uint32_t k,
                                                                       const U32x2 *__restrict__ ext, int n_shards,
                                                                       int shard, uint32_t *__restrict__ dstar,
                                                                       const uint32_t *__restrict__ gate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (gate && *gate == 0u) return;
    hamming_scan_program<WIDE>(hist, S, bins, Qpad, k, ext, n_shards, shard, dstar, blockIdx.x, smem_raw, MapDeviceExec{});
}

__global__ void __launch_bounds__(256) ap_finalize_kernel(const unsigned long long *__restrict__ psum, const uint32_t *__restrict__ phits,
                                                          int parts, long long stride, int Q, double *__restrict__ ap,
                                                          uint32_t *__restrict__ tsum, const uint32_t *__restrict__ gate) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (gate && *gate == 0u) return;
    if (q < Q) ap_finalize_item(psum, phits, parts, stride, q, ap, tsum);
}

__global__ void __launch_bounds__(256) ap_reduce_kernel(const unsigned long long *__restrict__ psum, const uint32_t *__restrict__ phits,
                                                        int S, int Qpad, int Q, unsigned long long *__restrict__ sum_q,
                                                        uint32_t *__restrict__ hits_q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) ap_reduce_item(psum, phits, S, Qpad, q, sum_q, hits_q);
}

// mean over queries (optionally masked), one CTA, fixed-shape tree => bit-reproducible
__global__ void __launch_bounds__(1024) mean_kernel(const double *__restrict__ ap, const uint8_t *__restrict__ mask, int Q,
                                                    double *__restrict__ out) {
    __shared__ double s_sum[1024];
    __shared__ unsigned s_cnt[1024];
    double s = 0.0;
    unsigned c = 0;
    for (int i = threadIdx.x; i < Q; i += 1024)
        if (!mask || mask[i]) s += ap[i], ++c;
    s_sum[threadIdx.x] = s, s_cnt[threadIdx.x] = c;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) s_sum[threadIdx.x] += s_sum[threadIdx.x + w], s_cnt[threadIdx.x] += s_cnt[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s_cnt[0] ? s_sum[0] / static_cast<double>(s_cnt[0]) : 0.0;
}

int launch_mean(const double *ap, const uint8_t *mask, int Q, double *out, cudaStream_t st) {
    mean_kernel<<<1, 1024, 0, st>>>(ap, mask, Q, out);
    B200_LAUNCH_CHECK("mean_kernel");
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ dispatch
using walk_fn = void (*)(const MapArgs);

// phase 0: stage A; phase 1: stage B by scoring again (`all`: k covers the whole database)
template <int CW, int LW, bool EQ>
static walk_fn pick3(bool wide, int phase, bool all) {
    if (!phase) return wide ? hamming_hist_kernel<CW, LW, EQ, true> : hamming_hist_kernel<CW, LW, EQ, false>;
    if (wide) return all ? hamming_walk_kernel<CW, LW, EQ, true, true> : hamming_walk_kernel<CW, LW, EQ, true, false>;
    return all ? hamming_walk_kernel<CW, LW, EQ, false, true> : hamming_walk_kernel<CW, LW, EQ, false, false>;
}
template <int CW>
static walk_fn pick2(int lw, bool eq, bool wide, int phase, bool all) {
    if (eq) return pick3<CW, 1, true>(wide, phase, all);
    switch (lw) {
        case 1: return pick3<CW, 1, false>(wide, phase, all);
        case 2: return pick3<CW, 2, false>(wide, phase, all);
        case 4: return pick3<CW, 4, false>(wide, phase, all);
    }
    return nullptr;
}
static walk_fn pick(int cw, int lw, bool eq, bool wide, int phase, bool all) {
    switch (cw) {
        case 1: return pick2<1>(lw, eq, wide, phase, all);
        case 2: return pick2<2>(lw, eq, wide, phase, all);
        case 4: return pick2<4>(lw, eq, wide, phase, all);
    }
    return nullptr;
}

// stage A always counts in 16|16-bit shared-memory counters (a segment has <= 65534 rows)
static size_t walk_smem(const b200_map_plan *p, int phase) {
    const int cw = b200_code_words(p->B);
    const bool all = p->k >= p->N_total;      // all-rows stage B keeps 16|16-bit running counts, bases stay in global memory
    return static_cast<size_t>(p->bins) * p->T * ((phase && p->wide && !all) ? 8 : 4) + static_cast<size_t>(p->tile) * (cw + p->LW) * 8;
}

static int map_threads_per_query() {
    const char *e = std::getenv("B200_MAP_TPQ");
    const int v = e ? std::atoi(e) : 2;
    return v >= 1 && v <= 4 ? v : 2;
}

static int check_plan(const b200_map_plan *p) {
    if (!p || p->Q < 1 || p->N < 0 || p->B < 1 || p->k < 1 || p->bins != p->B + 1 || p->T < 32 || p->S < 1) return B200_ERR_INVALID_ARG;
    return B200_OK;
}

// seg0 / nseg: stage A only — the range of database segments this launch covers (nseg < 0: all of them)
static int launch_walk(const b200_map_plan *p, int phase, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc,
                       const uint64_t *dl, void *ws, uint32_t *rank_idx, uint16_t *rank_dist, long long index_base,
                       cudaStream_t st, int seg0 = 0, int nseg = -1, const uint32_t *gate = nullptr) {
    const int cw = b200_code_words(p->B);
    const bool all = p->k >= p->N_total;
    walk_fn fn = pick(cw, p->LW, p->label_mode == B200_LABELS_EQUAL, p->wide != 0, phase, all);
    if (!fn) return B200_ERR_UNSUPPORTED;
    const size_t smem = walk_smem(p, phase);
    B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    unsigned char *w = static_cast<unsigned char *>(ws);
    MapArgs a;
    a.gate = gate;
    a.q_codes = qc, a.q_labels = ql, a.db_codes = dc, a.db_labels = dl;
    a.hist = w + p->off_hist;
    a.dstar = reinterpret_cast<const uint32_t *>(w + p->off_dstar);
    a.psum = reinterpret_cast<unsigned long long *>(w + p->off_psum);
    a.phits = reinterpret_cast<uint32_t *>(w + p->off_phits);
    a.rank_idx = rank_idx, a.rank_dist = rank_dist, a.index_base = index_base;
    a.seg_base = phase == 0 ? seg0 : 0;
    const int grid_y = (phase == 0 && nseg >= 0) ? nseg : p->S;
    if (grid_y == 0) return B200_OK;
    a.stash_d = p->stash ? reinterpret_cast<U32x4 *>(w + p->off_stash_d) : nullptr;
    a.stash_r = p->stash ? reinterpret_cast<uint32_t *>(w + p->off_stash_r) : nullptr;
    a.Q = p->Q, a.N = static_cast<int>(p->N), a.bins = p->bins, a.seg_len = p->seg_len, a.tile = p->tile, a.Qpad = p->Qpad;
    a.k = static_cast<uint32_t>(p->k);
    if (phase == 1 && p->stash) {          // stage B from the stash: counters only, no database tile in shared memory
        walk_fn rf = p->wide ? (all ? hamming_rank_kernel<true, true> : hamming_rank_kernel<true, false>)
                             : (all ? hamming_rank_kernel<false, true> : hamming_rank_kernel<false, false>);
        const size_t rsmem = static_cast<size_t>(p->bins) * p->T * ((p->wide && !all) ? 8 : 4);
        B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(rf), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(rsmem)));
        rf<<<dim3(p->groups, p->S), p->T, rsmem, st>>>(a);
        B200_LAUNCH_CHECK("hamming_rank_kernel");
        return B200_OK;
    }
    // stage A: two threads per query (blockDim = (T, 2)); B200_MAP_TPQ=1 for the one-thread form (A/B)
    fn<<<dim3(p->groups, grid_y), dim3(p->T, phase == 0 ? map_threads_per_query() : 1), smem, st>>>(a);
    B200_LAUNCH_CHECK(phase ? "hamming_ap_kernel" : "hamming_hist_kernel");
    return B200_OK;
}

static int launch_scan(const b200_map_plan *p, void *ws, const uint32_t *ext, int n_shards, int shard, cudaStream_t st,
                       const uint32_t *gate = nullptr) {
    unsigned char *w = static_cast<unsigned char *>(ws);
    const size_t smem = static_cast<size_t>(2) * p->bins * kScanQ * sizeof(U32x2);
    const dim3 block(kScanQ, kScanY);
    const int grid = p->Qpad / kScanQ;
    uint32_t *dstar = reinterpret_cast<uint32_t *>(w + p->off_dstar);
    if (p->wide) {
        B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(hamming_scan_kernel<true>),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        hamming_scan_kernel<true><<<grid, block, smem, st>>>(w + p->off_hist, p->S, p->bins, p->Qpad, static_cast<uint32_t>(p->k),
                                                            reinterpret_cast<const U32x2 *>(ext), n_shards, shard, dstar, gate);
    } else {
        B200_CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void *>(hamming_scan_kernel<false>),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        hamming_scan_kernel<false><<<grid, block, smem, st>>>(w + p->off_hist, p->S, p->bins, p->Qpad,
                                                             static_cast<uint32_t>(p->k), reinterpret_cast<const U32x2 *>(ext),
                                                             n_shards, shard, dstar, gate);
    }
    B200_LAUNCH_CHECK("hamming_scan_kernel");
    return B200_OK;
}

static int ap_finalize_launch(const uint64_t *sums, const uint32_t *hits, int n_parts, long long stride, int Q, double *ap,
                              uint32_t *tsum, const uint32_t *gate, cudaStream_t st) {
    ap_finalize_kernel<<<ceil_div(Q, 256), 256, 0, st>>>(reinterpret_cast<const unsigned long long *>(sums), hits, n_parts, stride,
                                                        Q, ap, tsum, gate);
    B200_LAUNCH_CHECK("ap_finalize_kernel");
    return B200_OK;
}

// stage A, S, B and the per-query finalize of an unsharded database (no mean)
static int three_stage_map(const b200_map_plan *p, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc, const uint64_t *dl,
                           void *ws, double *ap, uint32_t *tsum, const uint32_t *gate, cudaStream_t st) {
    if (int rc = launch_walk(p, 0, qc, ql, dc, dl, ws, nullptr, nullptr, 0, st, 0, -1, gate)) return rc;
    stage_mark(gate ? "gated_hist" : "hist", st);
    if (int rc = launch_scan(p, ws, nullptr, 1, 0, st, gate)) return rc;
    stage_mark(gate ? "gated_scan" : "scan", st);
    if (int rc = launch_walk(p, 1, qc, ql, dc, dl, ws, nullptr, nullptr, 0, st, 0, -1, gate)) return rc;
    stage_mark(gate ? "gated_ap" : "ap", st);
    unsigned char *w = static_cast<unsigned char *>(ws);
    return ap_finalize_launch(reinterpret_cast<const uint64_t *>(w + p->off_psum), reinterpret_cast<const uint32_t *>(w + p->off_phits),
                              p->S, p->Qpad, p->Q, ap, tsum, gate, st);
}

// Tail of b200_hamming_map for a caller that ran the select pipeline's first phases itself (select_begin, then
// select_segments over every segment as the database arrived: host_api.cu): rank, retry round, the gated three stages, mean.
// status != null: the optimistic form (round 0 only, b200_hamming_map_try's contract) — *status != 0 afterwards means the
// caller has to run b200_hamming_map on the complete packed database.
int hamming_map_after_select(const b200_map_plan *plan, const uint64_t *qc, const uint64_t *ql, const uint64_t *dc, const uint64_t *dl,
                             void *ws, double *ap, uint32_t *tsum, double *map_out, uint32_t *status, cudaStream_t st) {
    if (int rc = select_finish(plan, qc, ql, dc, dl, ws, ap, tsum, nullptr, nullptr, status, true, st)) return rc;
    if (!status) {
        const uint32_t *gate = reinterpret_cast<const uint32_t *>(static_cast<unsigned char *>(ws) + plan->off_sel_flags) + kSelFlagFallback;
        if (int rc = three_stage_map(plan, qc, ql, dc, dl, ws, ap, tsum, gate, st)) return rc;
    }
    if (map_out) return launch_mean(ap, nullptr, plan->Q, map_out, st);
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_map_plan_init(b200_map_plan *plan, int Q, long long N, long long N_total, int B, int LW, int label_mode,
                       long long k) {
    return map_plan_init(plan, Q, N, N_total, B, LW, label_mode, k, sm_count(), 1, true);
}

int b200_hamming_hist(const b200_map_plan *plan, const uint64_t *q_codes, const uint64_t *q_labels,
                      const uint64_t *db_codes, const uint64_t *db_labels, void *workspace, b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!q_codes || !q_labels || !workspace || (plan->N > 0 && (!db_codes || !db_labels))) return B200_ERR_INVALID_ARG;
    if (int rc = launch_walk(plan, 0, q_codes, q_labels, db_codes, db_labels, workspace, nullptr, nullptr, 0, as_stream(stream)))
        return rc;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    const int grid = static_cast<int>(ceil_div<size_t>(static_cast<size_t>(plan->bins) * plan->Qpad, 256));
    if (plan->wide)
        hamming_totals_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(w + plan->off_hist, plan->S, plan->bins, plan->Qpad,
                                                                        reinterpret_cast<U32x2 *>(w + plan->off_tot));
    else
        hamming_totals_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(w + plan->off_hist, plan->S, plan->bins, plan->Qpad,
                                                                         reinterpret_cast<U32x2 *>(w + plan->off_tot));
    B200_LAUNCH_CHECK("hamming_totals_kernel");
    return B200_OK;
}

int b200_hamming_scan(const b200_map_plan *plan, void *workspace, const uint32_t *tot_all_shards, int n_shards, int shard,
                      b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!workspace) return B200_ERR_WORKSPACE;
    if (tot_all_shards && (n_shards < 1 || shard < 0 || shard >= n_shards)) return B200_ERR_INVALID_ARG;
    return launch_scan(plan, workspace, tot_all_shards, n_shards, shard, as_stream(stream));
}

int b200_hamming_ap(const b200_map_plan *plan, const uint64_t *q_codes, const uint64_t *q_labels, const uint64_t *db_codes,
                    const uint64_t *db_labels, void *workspace, uint32_t *rank_idx, uint16_t *rank_dist, long long index_base,
                    b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!q_codes || !q_labels || !workspace || (plan->N > 0 && (!db_codes || !db_labels))) return B200_ERR_INVALID_ARG;
    return launch_walk(plan, 1, q_codes, q_labels, db_codes, db_labels, workspace, rank_idx, rank_dist, index_base,
                       as_stream(stream));
}

int b200_ap_reduce(const b200_map_plan *plan, void *workspace, uint64_t *sum_q, uint32_t *hits_q, b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!workspace || !sum_q || !hits_q) return B200_ERR_INVALID_ARG;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    ap_reduce_kernel<<<ceil_div(plan->Q, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const unsigned long long *>(w + plan->off_psum), reinterpret_cast<const uint32_t *>(w + plan->off_phits),
        plan->S, plan->Qpad, plan->Q, reinterpret_cast<unsigned long long *>(sum_q), hits_q);
    B200_LAUNCH_CHECK("ap_reduce_kernel");
    return B200_OK;
}

int b200_ap_finalize(const uint64_t *sums, const uint32_t *hits, int n_parts, long long stride, int Q, double *ap,
                     uint32_t *tsum, double *map_out, b200_stream_t stream) {
    if (!sums || !hits || !ap || n_parts < 1 || Q < 1 || stride < Q) return B200_ERR_INVALID_ARG;
    if (int rc = ap_finalize_launch(sums, hits, n_parts, stride, Q, ap, tsum, nullptr, as_stream(stream))) return rc;
    if (map_out) return launch_mean(ap, nullptr, Q, map_out, as_stream(stream));
    return B200_OK;
}

int b200_mean_f64(const double *ap, const uint8_t *query_mask, int Q, double *out, b200_stream_t stream) {
    if (!ap || !out || Q < 1) return B200_ERR_INVALID_ARG;
    return launch_mean(ap, query_mask, Q, out, as_stream(stream));
}

int b200_hamming_map(const b200_map_plan *plan, const uint64_t *q_codes, const uint64_t *q_labels, const uint64_t *db_codes,
                     const uint64_t *db_labels, void *workspace, double *ap, uint32_t *tsum, double *map_out,
                     b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!q_codes || !q_labels || !workspace || !ap || (plan->N > 0 && (!db_codes || !db_labels))) return B200_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    const uint32_t *gate = nullptr;
    if (plan->select) {       // candidate-list pipeline first; the three stages below then only run if it gave up (gate != 0)
        if (int rc = hamming_select_run(plan, q_codes, q_labels, db_codes, db_labels, workspace, ap, tsum, nullptr, nullptr, nullptr, st))
            return rc;
        gate = reinterpret_cast<const uint32_t *>(static_cast<unsigned char *>(workspace) + plan->off_sel_flags) + kSelFlagFallback;
    }
    if (int rc = three_stage_map(plan, q_codes, q_labels, db_codes, db_labels, workspace, ap, tsum, gate, st)) return rc;
    if (map_out) return launch_mean(ap, nullptr, plan->Q, map_out, st);
    return B200_OK;
}

// b200_hamming_map with a CUDA event after every stage: device time per stage, for bench.py's roofline object (the
// stages of a graph replay cannot be timed one by one).  Synchronises.  names_out[i] points to static strings.
int b200_hamming_map_stage_ms(const b200_map_plan *plan, const uint64_t *q_codes, const uint64_t *q_labels, const uint64_t *db_codes,
                              const uint64_t *db_labels, void *workspace, double *ap, uint32_t *tsum, int max_stages,
                              float *ms_out, const char **names_out, int *n_stages, b200_stream_t stream) {
    if (!ms_out || !names_out || !n_stages || max_stages < 1) return B200_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    StageMarks &m = stage_marks();
    for (int i = 0; i < StageMarks::kMax; ++i) B200_CUDA_TRY(cudaEventCreate(&m.ev[i]));
    m.n = 0, m.on = true;
    stage_mark("begin", st);
    const int rc = b200_hamming_map(plan, q_codes, q_labels, db_codes, db_labels, workspace, ap, tsum, nullptr, stream);
    stage_mark("finalize", st);
    m.on = false;
    int out = 0;
    if (rc == B200_OK && cudaStreamSynchronize(st) == cudaSuccess) {
        for (int i = 1; i < m.n && out < max_stages; ++i, ++out) {
            cudaEventElapsedTime(&ms_out[out], m.ev[i - 1], m.ev[i]);
            names_out[out] = m.name[i];
        }
    }
    *n_stages = out;
    for (int i = 0; i < StageMarks::kMax; ++i) cudaEventDestroy(m.ev[i]);
    return rc;
}

int b200_hamming_map_try(const b200_map_plan *plan, const uint64_t *q_codes, const uint64_t *q_labels, const uint64_t *db_codes,
                         const uint64_t *db_labels, void *workspace, double *ap, uint32_t *tsum, uint32_t *status,
                         b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!q_codes || !q_labels || !workspace || !ap || !status || (plan->N > 0 && (!db_codes || !db_labels))) return B200_ERR_INVALID_ARG;
    cudaStream_t st = as_stream(stream);
    if (plan->select)
        return hamming_select_run(plan, q_codes, q_labels, db_codes, db_labels, workspace, ap, tsum, nullptr, nullptr, status, st);
    return three_stage_map(plan, q_codes, q_labels, db_codes, db_labels, workspace, ap, tsum, nullptr, st);
}

// out[0] = mean of ap[0..Q), out[1] = 1.0 when any of the n_status words (status_stride bytes apart) is set
__global__ void __launch_bounds__(1024) map_final_kernel(const double *__restrict__ ap, int Q, const unsigned char *__restrict__ status,
                                                         int n_status, long long status_stride, double *__restrict__ out) {
    __shared__ double s_sum[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < Q; i += 1024) s += ap[i];
    s_sum[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) s_sum[threadIdx.x] += s_sum[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t any = 0;
        for (int r = 0; r < n_status; ++r) any |= *reinterpret_cast<const uint32_t *>(status + static_cast<long long>(r) * status_stride);
        out[0] = s_sum[0] / static_cast<double>(Q);
        out[1] = any ? 1.0 : 0.0;
    }
}

int b200_map_final(const double *ap, int Q, const void *status, int n_status, long long status_stride, double *out2,
                   b200_stream_t stream) {
    if (!ap || !out2 || Q < 1 || n_status < 0 || (n_status > 0 && !status)) return B200_ERR_INVALID_ARG;
    map_final_kernel<<<1, 1024, 0, as_stream(stream)>>>(ap, Q, static_cast<const unsigned char *>(status), n_status, status_stride, out2);
    B200_LAUNCH_CHECK("map_final_kernel");
    return B200_OK;
}

int b200_hamming_topk(const b200_map_plan *plan, const uint64_t *q_codes, const uint64_t *db_codes, void *workspace,
                      uint32_t *idx, uint16_t *dist, b200_stream_t stream) {
    if (int rc = check_plan(plan)) return rc;
    if (!q_codes || !workspace || (!idx && !dist) || (plan->N > 0 && !db_codes)) return B200_ERR_INVALID_ARG;
    if (plan->LW != 1) return B200_ERR_INVALID_ARG;   // plan for top-k is label-free: LW = 1, any label mode
    cudaStream_t st = as_stream(stream);
    const uint32_t *gate = nullptr;
    // relevance is irrelevant here: the code words double as (ignored) label words
    if (plan->select) {
        if (int rc = hamming_select_run(plan, q_codes, q_codes, db_codes, db_codes, workspace, nullptr, nullptr, idx, dist, nullptr, st))
            return rc;
        gate = reinterpret_cast<const uint32_t *>(static_cast<unsigned char *>(workspace) + plan->off_sel_flags) + kSelFlagFallback;
    }
    if (int rc = launch_walk(plan, 0, q_codes, q_codes, db_codes, db_codes, workspace, nullptr, nullptr, 0, st, 0, -1, gate)) return rc;
    if (int rc = launch_scan(plan, workspace, nullptr, 1, 0, st, gate)) return rc;
    return launch_walk(plan, 1, q_codes, q_codes, db_codes, db_codes, workspace, idx, dist, 0, st, 0, -1, gate);
}

}  // extern "C"
