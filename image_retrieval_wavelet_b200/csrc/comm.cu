// Peer-memory exchange between the ranks of one box (one process per GPU): the collectives of the sharded evaluator
// as plain stores over NVLink / NVSwitch instead of NCCL launches.
//
// What it replaces: the reference's only multi-GPU retrieval, faiss' index_cpu_to_all_gpus(shards=True)
// (/root/reference/main/engine/get_knn.py:41-44), merges per-GPU results on the host.  Here every rank owns one
// device region of the same size (cudaMalloc + CUDA IPC, mapped into every peer), laid out identically, and
//   * producers write their part straight into EVERY peer's copy from inside the producing kernel (bit-packing
//     writes the packed shard to all ranks: pack + all-gather in one pass over the float codes),
//   * a barrier kernel (one release store per peer + an acquire spin per peer, ~one NVLink round trip) replaces the
//     rendezvous of a collective; it is a normal stream-ordered kernel, so the whole evaluation step — packing,
//     exchange, evaluation, result exchange, mean — is ONE CUDA graph with no host involvement.
// Exchanged volumes are small (a packed COCO database is 5.6 MB, the per-query results 60 KB): the cost of a
// collective here is its launch + rendezvous latency, which is exactly what this removes.
#include <cstdio>
#include <cstring>

#include "common.cuh"

struct b200_comm {
    int rank, world;
    size_t bytes;                 // payload bytes of a region
    unsigned char *local;         // this rank's region (payload + control page)
    unsigned char *peer[B200_COMM_MAX_RANKS];      // every rank's region as mapped here (peer[rank] == local)
    bool opened[B200_COMM_MAX_RANKS];
};

namespace b200 {

constexpr size_t kCtlBytes = 4096;      // control page behind the payload: arrival flags, epoch, status
// control page layout (uint32 words): [0..15] arrival epoch of rank r (written by rank r), [16] this rank's epoch,
// [17] status (1: a barrier timed out)
constexpr int kCtlEpoch = 16, kCtlStatus = 17;

struct PeerPtrs {
    void *p[B200_COMM_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One CTA, one thread per rank.  Thread t announces this rank's new epoch in rank t's control page and waits until
// rank t has announced the same epoch here.  Everything this rank stored to peer memory earlier in the stream is
// ordered before the announcement (kernel boundary + system fence + release), everything a peer stored before ITS
// announcement is visible after the acquire.
__global__ void __launch_bounds__(B200_COMM_MAX_RANKS) comm_barrier_kernel(PeerPtrs ctl, int rank, int world, long long spin_limit) {
    uint32_t *mine = static_cast<uint32_t *>(ctl.p[rank]);
    const int t = threadIdx.x;
    const uint32_t e = mine[kCtlEpoch] + 1u;
    __syncthreads();
    if (t < world) {
        __threadfence_system();
        st_release_sys(static_cast<uint32_t *>(ctl.p[t]) + rank, e);
        const long long t0 = clock64();
        while (static_cast<int>(ld_acquire_sys(mine + t) - e) < 0) {
            if (clock64() - t0 > spin_limit) {          // a peer died or never launched: give up instead of hanging the GPU
                mine[kCtlStatus] = 1u;
                break;
            }
        }
    }
    __syncthreads();
    if (t == 0) mine[kCtlEpoch] = e;
}

// region[r][off_k + i] = src_k[i] for every rank r and up to 4 segments k (16-byte units): the small result vectors of a
// step (AP slice, hit counts, status word) reach all ranks in one launch
struct PutSegs {
    const uint4 *src[4];
    long long off16[4], n16[4];
    int n;
};
__global__ void __launch_bounds__(256) comm_put_kernel(PutSegs segs, PeerPtrs region, int world) {
    for (int k = 0; k < segs.n; ++k) {
        const uint4 *src = segs.src[k];
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < segs.n16[k];
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const uint4 v = src[i];
            for (int r = 0; r < world; ++r) static_cast<uint4 *>(region.p[r])[segs.off16[k] + i] = v;
        }
    }
}

// Last kernel of a multi-GPU evaluation step: put + barrier + mean in ONE single-CTA launch (three graph nodes less on a
// step whose fixed cost is launch latency).  (1) the segments go to every rank's region, (2) the barrier of
// comm_barrier_kernel (every rank's results have landed here once it is passed), (3) out2[0] = mean of the Q values at
// ap_off16 of the LOCAL region in the fixed order of map_final_kernel (hamming_map.cu: bit-identical to it, so a mAP
// does not depend on the world size), out2[1] = 1.0 when any rank's status word is set.
__global__ void __launch_bounds__(1024) comm_put_barrier_final_kernel(PutSegs segs, PeerPtrs region, PeerPtrs ctl, int rank, int world,
                                                                      long long spin_limit, long long ap_off16, int Q,
                                                                      long long status_off16, double *__restrict__ out2) {
    __shared__ double s_sum[1024];
    const int t = threadIdx.x;
    for (int k = 0; k < segs.n; ++k) {
        const uint4 *src = segs.src[k];
        for (long long i = t; i < segs.n16[k]; i += 1024) {
            const uint4 v = src[i];
            for (int r = 0; r < world; ++r) static_cast<uint4 *>(region.p[r])[segs.off16[k] + i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    uint32_t *mine = static_cast<uint32_t *>(ctl.p[rank]);
    const uint32_t e = mine[kCtlEpoch] + 1u;
    __syncthreads();
    if (t < world) {
        __threadfence_system();
        st_release_sys(static_cast<uint32_t *>(ctl.p[t]) + rank, e);
        const long long t0 = clock64();
        while (static_cast<int>(ld_acquire_sys(mine + t) - e) < 0) {
            if (clock64() - t0 > spin_limit) {
                mine[kCtlStatus] = 1u;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0) mine[kCtlEpoch] = e;
    // what the peers stored is read past L1 (this SM may hold lines of an earlier step)
    const double *ap = reinterpret_cast<const double *>(static_cast<const uint4 *>(region.p[rank]) + ap_off16);
    double s = 0.0;
    for (int i = t; i < Q; i += 1024) s += __ldcg(ap + i);
    s_sum[t] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (t < w) s_sum[t] += s_sum[t + w];
        __syncthreads();
    }
    if (t == 0) {
        const uint4 *st = static_cast<const uint4 *>(region.p[rank]) + status_off16;
        uint32_t any = 0;
        for (int r = 0; r < world; ++r) any |= __ldcg(reinterpret_cast<const uint32_t *>(st + r));
        out2[0] = s_sum[0] / static_cast<double>(Q);
        out2[1] = any ? 1.0 : 0.0;
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_comm_create(int rank, int world, size_t bytes, b200_comm **out) {
    if (!out || world < 1 || world > B200_COMM_MAX_RANKS || rank < 0 || rank >= world) return B200_ERR_INVALID_ARG;
    b200_comm *c = new b200_comm();
    c->rank = rank, c->world = world;
    c->bytes = round_up<size_t>(bytes ? bytes : 16, 256);
    for (int r = 0; r < B200_COMM_MAX_RANKS; ++r) c->peer[r] = nullptr, c->opened[r] = false;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, c->bytes + kCtlBytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, c->bytes + kCtlBytes);
    if (e != cudaSuccess) {
        set_last_cuda_error(e, "b200_comm_create");
        if (p) cudaFree(p);
        delete c;
        return B200_ERR_CUDA;
    }
    c->local = static_cast<unsigned char *>(p);
    c->peer[rank] = c->local;
    *out = c;
    return B200_OK;
}

int b200_comm_export(b200_comm *c, void *handle64) {
    if (!c || !handle64) return B200_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == B200_COMM_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    B200_CUDA_TRY(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle64, &h, sizeof(h));
    return B200_OK;
}

int b200_comm_open(b200_comm *c, const void *handles) {
    if (!c || !handles) return B200_ERR_INVALID_ARG;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char *>(handles) + static_cast<size_t>(r) * B200_COMM_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        B200_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<unsigned char *>(p);
        c->opened[r] = true;
    }
    return B200_OK;
}

void *b200_comm_buffer(b200_comm *c, int peer) {
    if (!c || peer < 0 || peer >= c->world) return nullptr;
    return c->peer[peer];
}

size_t b200_comm_bytes(b200_comm *c) { return c ? c->bytes : 0; }
int b200_comm_world(b200_comm *c) { return c ? c->world : 0; }
int b200_comm_rank(b200_comm *c) { return c ? c->rank : -1; }

int b200_comm_barrier(b200_comm *c, b200_stream_t stream) {
    if (!c) return B200_ERR_INVALID_ARG;
    PeerPtrs ctl;
    for (int r = 0; r < B200_COMM_MAX_RANKS; ++r) ctl.p[r] = r < c->world ? c->peer[r] + c->bytes : nullptr;
    for (int r = 0; r < c->world; ++r)
        if (!ctl.p[r]) return B200_ERR_INVALID_ARG;          // b200_comm_open has not run
    comm_barrier_kernel<<<1, B200_COMM_MAX_RANKS, 0, as_stream(stream)>>>(ctl, c->rank, c->world, 4000000000ll);   // ~2 s
    B200_LAUNCH_CHECK("comm_barrier_kernel");
    return B200_OK;
}

int b200_comm_put(b200_comm *c, int n_segments, const void *const *src, const size_t *dst_offset, const size_t *bytes,
                  b200_stream_t stream) {
    if (!c || n_segments < 1 || n_segments > 4 || !src || !dst_offset || !bytes) return B200_ERR_INVALID_ARG;
    PutSegs segs = {};
    long long most = 0;
    for (int k = 0; k < n_segments; ++k) {
        if (!src[k] || (dst_offset[k] & 15) || (bytes[k] & 15) || (reinterpret_cast<uintptr_t>(src[k]) & 15)) return B200_ERR_ALIGNMENT;
        if (dst_offset[k] + bytes[k] > c->bytes) return B200_ERR_INVALID_ARG;
        segs.src[k] = static_cast<const uint4 *>(src[k]);
        segs.off16[k] = static_cast<long long>(dst_offset[k] / 16), segs.n16[k] = static_cast<long long>(bytes[k] / 16);
        most = segs.n16[k] > most ? segs.n16[k] : most;
    }
    segs.n = n_segments;
    if (most == 0) return B200_OK;
    PeerPtrs region;
    for (int r = 0; r < B200_COMM_MAX_RANKS; ++r) region.p[r] = r < c->world ? c->peer[r] : nullptr;
    const int grid = static_cast<int>(ceil_div<long long>(most, 256) < 2ll * sm_count() ? ceil_div<long long>(most, 256) : 2ll * sm_count());
    comm_put_kernel<<<grid, 256, 0, as_stream(stream)>>>(segs, region, c->world);
    B200_LAUNCH_CHECK("comm_put_kernel");
    return B200_OK;
}

int b200_comm_put_barrier_final(b200_comm *c, int n_segments, const void *const *src, const size_t *dst_offset, const size_t *bytes,
                                size_t ap_offset, int Q, size_t status_offset, double *out2, b200_stream_t stream) {
    if (!c || n_segments < 0 || n_segments > 4 || (n_segments > 0 && (!src || !dst_offset || !bytes)) || !out2 || Q < 1)
        return B200_ERR_INVALID_ARG;
    if ((ap_offset & 15) || (status_offset & 15)) return B200_ERR_ALIGNMENT;
    if (ap_offset + static_cast<size_t>(Q) * 8 > c->bytes || status_offset + 16 * static_cast<size_t>(c->world) > c->bytes)
        return B200_ERR_INVALID_ARG;
    PutSegs segs = {};
    for (int k = 0; k < n_segments; ++k) {
        if (!src[k] || (dst_offset[k] & 15) || (bytes[k] & 15) || (reinterpret_cast<uintptr_t>(src[k]) & 15)) return B200_ERR_ALIGNMENT;
        if (dst_offset[k] + bytes[k] > c->bytes) return B200_ERR_INVALID_ARG;
        segs.src[k] = static_cast<const uint4 *>(src[k]);
        segs.off16[k] = static_cast<long long>(dst_offset[k] / 16), segs.n16[k] = static_cast<long long>(bytes[k] / 16);
    }
    segs.n = n_segments;
    PeerPtrs region, ctl;
    for (int r = 0; r < B200_COMM_MAX_RANKS; ++r) {
        region.p[r] = r < c->world ? c->peer[r] : nullptr;
        ctl.p[r] = r < c->world ? c->peer[r] + c->bytes : nullptr;
    }
    for (int r = 0; r < c->world; ++r)
        if (!region.p[r]) return B200_ERR_INVALID_ARG;       // b200_comm_open has not run
    comm_put_barrier_final_kernel<<<1, 1024, 0, as_stream(stream)>>>(segs, region, ctl, c->rank, c->world, 4000000000ll,
                                                                     static_cast<long long>(ap_offset / 16), Q,
                                                                     static_cast<long long>(status_offset / 16), out2);
    B200_LAUNCH_CHECK("comm_put_barrier_final_kernel");
    return B200_OK;
}

const void *b200_comm_status_word(b200_comm *c) {
    return c ? c->local + c->bytes + kCtlStatus * sizeof(uint32_t) : nullptr;
}

int b200_comm_status(b200_comm *c, int *timed_out) {
    if (!c || !timed_out) return B200_ERR_INVALID_ARG;
    uint32_t v = 0;
    B200_CUDA_TRY(cudaMemcpy(&v, c->local + c->bytes + kCtlStatus * sizeof(uint32_t), sizeof(v), cudaMemcpyDeviceToHost));
    *timed_out = static_cast<int>(v);
    return B200_OK;
}

int b200_comm_destroy(b200_comm *c) {
    if (!c) return B200_OK;
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    return B200_OK;
}

}  // extern "C"
