// HP-EVAL core: Hamming mAP@k as a counting sort (see hamming_map.cu for the algorithm and the reference lines).
//
// Per-CTA programs written once as phases handed to `exec` (device: run + __syncthreads(); CPU simulator used by
// tests/: run for tid = 0..nthreads-1).  Per-thread values that live across phases sit in a `State` reached through
// exec.state(): one register-resident struct on the device, an array indexed by tid in the simulator.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace b200 {

constexpr int kScanQ = 32;   // queries per scan CTA (one warp-width: coalesced rows of hist[.][.][q])
constexpr int kScanY = 32;  // bucket lanes per query: 32 x 32 = 1024 threads per CTA (the scan is latency-bound: more lanes, shorter serial loops)

struct alignas(8) U32x2 {
    uint32_t x, y;
};
struct alignas(16) U32x4 {
    uint32_t x, y, z, w;
};

struct MapArgs {
    const uint64_t *q_codes, *q_labels, *db_codes, *db_labels;
    void *hist;              // Ctr [S][bins][Qpad]
    const uint32_t *dstar;   // [Qpad]
    unsigned long long *psum; // [S][Qpad] fixed-point partial sums (ap_term)
    uint32_t *phits;         // [S][Qpad]
    uint32_t *rank_idx;      // [Q][k] or null
    uint16_t *rank_dist;     // [Q][k] or null
    U32x4 *stash_d;          // [ceil(N/16)][Qpad]: distance bytes of rows 16g..16g+15, one 128-bit word per query (stash mode) or null
    uint32_t *stash_r;       // [ceil(N/32)][Qpad]: relevance bits of rows 32g..32g+31 (stash mode) or null
    const uint32_t *gate;    // device word or null: when given and zero, the kernel returns at once (the three-stage path as the
                             // select pipeline's fallback, hamming_select.cu)
    long long index_base;
    int seg_base;            // stage A launched over a range of segments: segment = seg_base + blockIdx.y
    int Q, N, bins, seg_len, tile, Qpad;
    uint32_t k;
};

__host__ __device__ __forceinline__ uint32_t popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return static_cast<uint32_t>(__builtin_popcount(x));
#endif
}
__host__ __device__ __forceinline__ float div_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

// One AP summand, hit-ordinal / rank, as an exact 2^-40 fixed-point integer: the float32 quotient (IEEE division, the
// same bits on the device and in the CPU simulator) scaled by 2^40 converts without rounding, and integer addition is
// associative — so per-query AP sums do not depend on how the database was cut into segments or shards.
constexpr int kApFracBits = 40;
constexpr int kWalkBatch = 4;
__host__ __device__ __forceinline__ unsigned long long ap_term(uint32_t ordinal, uint32_t rank) {
    const float q = div_rn(static_cast<float>(ordinal), static_cast<float>(rank));
#ifdef __CUDA_ARCH__
    return __float2ull_rz(q * 1099511627776.0f);
#else
    return static_cast<unsigned long long>(q * 1099511627776.0f);
#endif
}

// (rows | relevant) counter pair.  Narrow: one uint32, 16 bits each (k <= 65534 and segments <= 65534 rows).
template <bool WIDE>
struct Ctr;
template <>
struct Ctr<false> {
    using type = uint32_t;
    static constexpr int kShift = 16;
    __host__ __device__ static __forceinline__ uint32_t lo(type c) { return c & 0xffffu; }
    __host__ __device__ static __forceinline__ uint32_t hi(type c) { return c >> 16; }
    __host__ __device__ static __forceinline__ type make(uint32_t a, uint32_t r) { return (a & 0xffffu) | (r << 16); }
};
template <>
struct Ctr<true> {
    using type = unsigned long long;
    static constexpr int kShift = 32;
    __host__ __device__ static __forceinline__ uint32_t lo(type c) { return static_cast<uint32_t>(c); }
    __host__ __device__ static __forceinline__ uint32_t hi(type c) { return static_cast<uint32_t>(c >> 32); }
    __host__ __device__ static __forceinline__ type make(uint32_t a, uint32_t r) {
        return static_cast<type>(a) | (static_cast<type>(r) << 32);
    }
};


// distance and relevance of staged database row j against the thread's query (row = one broadcast smem read)
template <int CW, int LW, bool EQ>
__host__ __device__ __forceinline__ void score_row(const uint32_t *s_codes, const uint32_t *s_labs, int j,
                                                   const uint32_t *qc, const uint32_t *ql, uint32_t &d, bool &rel) {
    uint32_t c[2 * CW], l[2 * LW];
    if constexpr (CW == 1) {
        const U32x2 v = reinterpret_cast<const U32x2 *>(s_codes)[j];
        c[0] = v.x, c[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
            const U32x4 v = reinterpret_cast<const U32x4 *>(s_codes)[j * (CW / 2) + i];
            c[4 * i] = v.x, c[4 * i + 1] = v.y, c[4 * i + 2] = v.z, c[4 * i + 3] = v.w;
        }
    }
    if constexpr (LW == 1) {
        const U32x2 v = reinterpret_cast<const U32x2 *>(s_labs)[j];
        l[0] = v.x, l[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < LW / 2; ++i) {
            const U32x4 v = reinterpret_cast<const U32x4 *>(s_labs)[j * (LW / 2) + i];
            l[4 * i] = v.x, l[4 * i + 1] = v.y, l[4 * i + 2] = v.z, l[4 * i + 3] = v.w;
        }
    }
    d = 0;
#pragma unroll
    for (int i = 0; i < 2 * CW; ++i) d += popc32(qc[i] ^ c[i]);
    if constexpr (EQ) {
        rel = (ql[0] == l[0]) && (ql[1] == l[1]);
    } else {
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < 2 * LW; ++i) any |= ql[i] & l[i];
        rel = any != 0;
    }
}

template <int CW, int LW>
struct WalkState {
    uint32_t qc[2 * CW], ql[2 * LW];
    uint32_t dstar, hits;
    unsigned long long sum;      // sum of ap_term(): exact 2^-40 fixed point
};

// Shared-memory counter updates.  Device: native shared atomics (fire-and-forget RED for stage A; with return for the
// all-rows walk) — a thread only ever touches its own column, the atomics are there to take the read-modify-write
// chain off the dependency path, not for exclusion.  Simulator: plain arithmetic.
__host__ __device__ __forceinline__ void ctr_add32(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
__host__ __device__ __forceinline__ uint32_t ctr_fetch_add32(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const uint32_t o = *p;
    *p = o + v;
    return o;
#endif
}

// A thread's private counter column cnt[.][t] of stage A: on the device the shared-window byte address and the byte
// stride between distances, so that a bump is one IMAD (FMA pipe) + one RED.shared — the ALU pipe is the busy one.
struct CtrColumn {
#ifdef __CUDA_ARCH__
    uint32_t base, stride;
    __device__ __forceinline__ CtrColumn(uint32_t *cnt, int T, int t)
        : base(static_cast<uint32_t>(__cvta_generic_to_shared(cnt + t))), stride(static_cast<uint32_t>(T) * 4u) {}
    __device__ __forceinline__ void add(uint32_t d, uint32_t v) const {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(base + d * stride), "r"(v) : "memory");
    }
#else
    uint32_t *p;
    int T;
    CtrColumn(uint32_t *cnt, int T_, int t) : p(cnt + t), T(T_) {}
    void add(uint32_t d, uint32_t v) const { p[static_cast<size_t>(d) * T] += v; }
#endif
};

// Stage A: (rows | relevant << 16) histogram of one database segment (<= 65534 rows, so 16-bit halves always suffice
// in shared memory; the global histogram entry is widened on the way out when the plan uses wide counters), and —
// stash mode — the (distance, relevance) of every (row, query) pair for stage B.
// Grid (query group gx, segment gy); T = queries per CTA, run by nt = T or 2T threads: with two threads per query the
// halves take alternate 32-row groups of every tile and bump the same counter column (the bumps are shared-memory
// atomics already), which doubles the resident warps at the same shared-memory footprint — ncu had stage A at 17 %
// occupancy (the counters cap it at 3 CTAs of 128 queries per SM) with fixed-latency dependency stalls on top.
// smem: cnt u32 [bins][T] | codes[tile] | labels[tile].
template <int CW, int LW, bool EQ, bool WIDE, typename Exec, typename LoadTile>
__host__ __device__ __forceinline__ void hamming_hist_program(const MapArgs &a, int gx, int gy, int T, unsigned char *smem,
                                                              Exec exec, LoadTile load_tile) {
    using C = Ctr<WIDE>;
    using ctr_t = typename C::type;
    using State = WalkState<CW, LW>;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_codes = cnt + static_cast<size_t>(a.bins) * T;
    uint32_t *s_labs = s_codes + static_cast<size_t>(a.tile) * 2 * CW;
    const int seg_begin = gy * a.seg_len;
    const int seg_end = seg_begin + a.seg_len < a.N ? seg_begin + a.seg_len : a.N;
    ctr_t *hist_seg = static_cast<ctr_t *>(a.hist) + static_cast<size_t>(gy) * a.bins * a.Qpad + static_cast<size_t>(gx) * T;
    State local;
    State *states = exec.state(&local);

    exec([&](int tt, int nt) {
        State &st = states[exec.slot(tt)];
        const int tpq = nt / T, half = tt / T, t = tt - half * T;      // half: which of the tpq threads of the query
        const int q = gx * T + t;
        const int qq = q < a.Q ? q : a.Q - 1;     // padding threads replay the last query; their outputs land in padding
        const uint32_t *pc = reinterpret_cast<const uint32_t *>(a.q_codes) + static_cast<size_t>(qq) * 2 * CW;
        const uint32_t *pl = reinterpret_cast<const uint32_t *>(a.q_labels) + static_cast<size_t>(qq) * 2 * LW;
#pragma unroll
        for (int i = 0; i < 2 * CW; ++i) st.qc[i] = pc[i];
#pragma unroll
        for (int i = 0; i < 2 * LW; ++i) st.ql[i] = pl[i];
        for (int d = half; d < a.bins; d += tpq) cnt[d * T + t] = 0;
    });

    for (int tile0 = seg_begin; tile0 < seg_end; tile0 += a.tile) {
        const int n = a.tile < seg_end - tile0 ? a.tile : seg_end - tile0;
        // rows tile0 .. tile0+n are contiguous in both arrays; tile0 is even and the buffers are padded to even rows
        exec([&](int t, int nt) {
            load_tile(s_codes, a.db_codes + static_cast<size_t>(tile0) * CW, (n * CW + 1) / 2, t, nt);
            load_tile(s_labs, a.db_labels + static_cast<size_t>(tile0) * LW, (n * LW + 1) / 2, t, nt);
        });
        exec([&](int tt, int nt) {
            State &st = states[exec.slot(tt)];
            const int tpq = nt / T, half = tt / T, t = tt - half * T;      // half: which of the tpq threads of the query
            const int q = gx * T + t;
            const bool stash = a.stash_d != nullptr;
            const CtrColumn col(cnt, T, t);
            int j = 32 * half;
            // whole 32-row groups, fully unrolled: the relevance bits of the group land at compile-time positions
            for (; j + 32 <= n; j += 32 * tpq) {
                uint32_t relw = 0, dw[4];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    uint32_t d[4];
                    bool rel[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) score_row<CW, LW, EQ>(s_codes, s_labs, j + 4 * b + i, st.qc, st.ql, d[i], rel[i]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) col.add(d[i], rel[i] ? 0x10001u : 1u);
                    if (stash) {
                        dw[b & 3] = d[0] | (d[1] << 8) | (d[2] << 16) | (d[3] << 24);
                        if ((b & 3) == 3)      // 16 rows = one 128-bit store, consecutive queries = consecutive words
                            a.stash_d[static_cast<size_t>((tile0 + j + 4 * b) >> 4) * a.Qpad + q] = U32x4{dw[0], dw[1], dw[2], dw[3]};
#pragma unroll
                        for (int i = 0; i < 4; ++i) relw |= rel[i] ? (1u << (4 * b + i)) : 0u;
                    }
                }
                if (stash) a.stash_r[static_cast<size_t>((tile0 + j) >> 5) * a.Qpad + q] = relw;
            }
            // the last, partial group of the database: the half whose turn it would be
            if (((n >> 5) % tpq) != half) return;
            uint32_t relw = 0, tw[4] = {0u, 0u, 0u, 0u};
            for (j = n & ~31; j < n; ++j) {
                uint32_t d;
                bool rel;
                score_row<CW, LW, EQ>(s_codes, s_labs, j, st.qc, st.ql, d, rel);
                col.add(d, rel ? 0x10001u : 1u);
                tw[(j >> 2) & 3] |= d << (8 * (j & 3));
                relw |= static_cast<uint32_t>(rel) << (j & 31);
                if (stash && (j & 15) == 15) {
                    a.stash_d[static_cast<size_t>((tile0 + j) >> 4) * a.Qpad + q] = U32x4{tw[0], tw[1], tw[2], tw[3]};
                    tw[0] = tw[1] = tw[2] = tw[3] = 0u;
                }
            }
            if (stash) {
                if (n & 15) a.stash_d[static_cast<size_t>((tile0 + n) >> 4) * a.Qpad + q] = U32x4{tw[0], tw[1], tw[2], tw[3]};
                if (n & 31) a.stash_r[static_cast<size_t>((tile0 + n) >> 5) * a.Qpad + q] = relw;
            }
        });
    }

    exec([&](int tt, int nt) {
        const int tpq = nt / T, half = tt / T, t = tt - half * T;      // half: which of the tpq threads of the query
        for (int d = half; d < a.bins; d += tpq) {
            const uint32_t c = cnt[d * T + t];
            hist_seg[static_cast<size_t>(d) * a.Qpad + t] = C::make(c & 0xffffu, c >> 16);
        }
    });
}

// Stage B by scoring again (no stash): AP partials + optional ranked-list emission.  ALL (k >= database size): every
// row counts, so the counter bump is unconditional — one shared atomic with return per row, no dependency chain;
// otherwise only rows with distance <= d* and rank <= k bump their counter (4 rows per batch, same-counter collisions
// resolved in registers).  smem: cnt[bins][T] | codes[tile] | labels[tile].
template <int CW, int LW, bool EQ, bool WIDE, bool ALL, typename Exec, typename LoadTile>
__host__ __device__ __forceinline__ void hamming_walk_program(const MapArgs &a, int gx, int gy, int T, unsigned char *smem,
                                                              Exec exec, LoadTile load_tile) {
    using C = Ctr<WIDE>;
    using ctr_t = typename C::type;
    using State = WalkState<CW, LW>;
    // ALL with wide bases (k > 65534): the shared counters are the 16|16-bit running counts of this segment (as in
    // stage A) and the rank / ordinal bases stay in the global histogram (L2-resident), fetched for relevant rows only;
    // otherwise the shared counters start at the bases.
    ctr_t *cnt = reinterpret_cast<ctr_t *>(smem);
    uint32_t *run = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_codes = ALL ? run + static_cast<size_t>(a.bins) * T : reinterpret_cast<uint32_t *>(cnt + static_cast<size_t>(a.bins) * T);
    uint32_t *s_labs = s_codes + static_cast<size_t>(a.tile) * 2 * CW;
    const int seg_begin = gy * a.seg_len;
    const int seg_end = seg_begin + a.seg_len < a.N ? seg_begin + a.seg_len : a.N;
    ctr_t *hist_seg = static_cast<ctr_t *>(a.hist) + static_cast<size_t>(gy) * a.bins * a.Qpad + static_cast<size_t>(gx) * T;
    State local;
    State *states = exec.state(&local);

    exec([&](int t, int) {
        State &st = states[exec.slot(t)];
        const int q = gx * T + t;
        const int qq = q < a.Q ? q : a.Q - 1;     // padding threads replay the last query; their outputs land in padding
        const uint32_t *pc = reinterpret_cast<const uint32_t *>(a.q_codes) + static_cast<size_t>(qq) * 2 * CW;
        const uint32_t *pl = reinterpret_cast<const uint32_t *>(a.q_labels) + static_cast<size_t>(qq) * 2 * LW;
#pragma unroll
        for (int i = 0; i < 2 * CW; ++i) st.qc[i] = pc[i];
#pragma unroll
        for (int i = 0; i < 2 * LW; ++i) st.ql[i] = pl[i];
        st.sum = 0ull, st.hits = 0;
        st.dstar = a.dstar[q];
        if (ALL && WIDE) {
            for (int d = 0; d < a.bins; ++d) run[d * T + t] = 0u;
        } else {       // narrow counters (k <= 65534) hold base + running count themselves, also in the all-rows walk
            for (int d = 0; d < a.bins; ++d) cnt[d * T + t] = hist_seg[static_cast<size_t>(d) * a.Qpad + t];
        }
    });

    for (int tile0 = seg_begin; tile0 < seg_end; tile0 += a.tile) {
        const int n = a.tile < seg_end - tile0 ? a.tile : seg_end - tile0;
        exec([&](int t, int nt) {
            load_tile(s_codes, a.db_codes + static_cast<size_t>(tile0) * CW, (n * CW + 1) / 2, t, nt);
            load_tile(s_labs, a.db_labels + static_cast<size_t>(tile0) * LW, (n * LW + 1) / 2, t, nt);
        });
        exec([&](int t, int) {
            State &st = states[exec.slot(t)];
            const int q = gx * T + t;
            const uint32_t k = a.k;
            const bool emit = (a.rank_idx != nullptr || a.rank_dist != nullptr) && q < a.Q;
            int j = 0;
            if (ALL) {
                // wide: 8 rows per batch so that the base fetches (L2 latency) of a batch overlap
                constexpr int kAllBatch = WIDE ? 2 * kWalkBatch : kWalkBatch;
                for (; j + kAllBatch <= n; j += kAllBatch) {
                    uint32_t d[kAllBatch], rank[kAllBatch], ordinal[kAllBatch];
                    bool rel[kAllBatch];
#pragma unroll
                    for (int i = 0; i < kAllBatch; ++i) score_row<CW, LW, EQ>(s_codes, s_labs, j + i, st.qc, st.ql, d[i], rel[i]);
#pragma unroll
                    for (int i = 0; i < kAllBatch; ++i) {
                        const uint32_t o = ctr_fetch_add32(run + d[i] * T + t, 1u + (static_cast<uint32_t>(rel[i]) << 16));
                        rank[i] = (o & 0xffffu) + 1u, ordinal[i] = (o >> 16) + 1u;
                        if (WIDE && (rel[i] || emit)) {
                            const ctr_t b = hist_seg[static_cast<size_t>(d[i]) * a.Qpad + t];
                            rank[i] += C::lo(b), ordinal[i] += C::hi(b);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < kAllBatch; ++i) {
                        if (rel[i]) {
                            st.sum += ap_term(ordinal[i], rank[i]);
                            ++st.hits;
                        }
                        if (emit) {
                            const size_t o = static_cast<size_t>(q) * k + (rank[i] - 1u);
                            if (a.rank_idx) a.rank_idx[o] = static_cast<uint32_t>(a.index_base + tile0 + j + i);
                            if (a.rank_dist) a.rank_dist[o] = static_cast<uint16_t>(d[i]);
                        }
                    }
                }
            } else {
                // kWalkBatch rows per iteration: their scores and counter loads are independent, so the shared-memory
                // latency of the read-modify-write chain is paid once per batch; rows of the batch that hit the same
                // counter are resolved in registers (later rows see the earlier rows' updates) and stored in row order.
                for (; j + kWalkBatch <= n; j += kWalkBatch) {
                    uint32_t d[kWalkBatch], addr[kWalkBatch];
                    bool rel[kWalkBatch], take[kWalkBatch];
                    ctr_t c[kWalkBatch];
#pragma unroll
                    for (int i = 0; i < kWalkBatch; ++i) {
                        score_row<CW, LW, EQ>(s_codes, s_labs, j + i, st.qc, st.ql, d[i], rel[i]);
                        addr[i] = d[i] * T + t;
                    }
#pragma unroll
                    for (int i = 0; i < kWalkBatch; ++i) {
                        take[i] = d[i] <= st.dstar;
                        c[i] = take[i] ? cnt[addr[i]] : static_cast<ctr_t>(0);
                    }
#pragma unroll
                    for (int i = 0; i < kWalkBatch; ++i) {
#pragma unroll
                        for (int e = 0; e < i; ++e)
                            if (addr[e] == addr[i]) c[i] = c[e];
                        const uint32_t rank = C::lo(c[i]) + 1u;
                        if (take[i] && rank <= k) {
                            c[i] += static_cast<ctr_t>(1) + (static_cast<ctr_t>(rel[i]) << C::kShift);
                            if (rel[i]) {
                                st.sum += ap_term(C::hi(c[i]), rank);
                                ++st.hits;
                            }
                            if (emit) {
                                const size_t o = static_cast<size_t>(q) * k + (rank - 1u);
                                if (a.rank_idx) a.rank_idx[o] = static_cast<uint32_t>(a.index_base + tile0 + j + i);
                                if (a.rank_dist) a.rank_dist[o] = static_cast<uint16_t>(d[i]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < kWalkBatch; ++i)
                        if (take[i]) cnt[addr[i]] = c[i];
                }
            }
            for (; j < n; ++j) {
                uint32_t d;
                bool rel;
                score_row<CW, LW, EQ>(s_codes, s_labs, j, st.qc, st.ql, d, rel);
                if (ALL) {
                    const uint32_t o = ctr_fetch_add32(run + d * T + t, 1u + (static_cast<uint32_t>(rel) << 16));
                    const ctr_t b = WIDE ? hist_seg[static_cast<size_t>(d) * a.Qpad + t] : static_cast<ctr_t>(0);
                    const uint32_t rank = C::lo(b) + (o & 0xffffu) + 1u;
                    if (rel) {
                        st.sum += ap_term(C::hi(b) + (o >> 16) + 1u, rank);
                        ++st.hits;
                    }
                    if (emit) {
                        const size_t oo = static_cast<size_t>(q) * k + (rank - 1u);
                        if (a.rank_idx) a.rank_idx[oo] = static_cast<uint32_t>(a.index_base + tile0 + j);
                        if (a.rank_dist) a.rank_dist[oo] = static_cast<uint16_t>(d);
                    }
                } else if (d <= st.dstar) {
                    ctr_t c = cnt[d * T + t];
                    const uint32_t rank = C::lo(c) + 1u;
                    if (rank <= k) {
                        c += static_cast<ctr_t>(1) + (static_cast<ctr_t>(rel) << C::kShift);
                        cnt[d * T + t] = c;
                        if (rel) {
                            st.sum += ap_term(C::hi(c), rank);
                            ++st.hits;
                        }
                        if (emit) {
                            const size_t o = static_cast<size_t>(q) * k + (rank - 1u);
                            if (a.rank_idx) a.rank_idx[o] = static_cast<uint32_t>(a.index_base + tile0 + j);
                            if (a.rank_dist) a.rank_dist[o] = static_cast<uint16_t>(d);
                        }
                    }
                }
            }
        });
    }

    exec([&](int t, int) {
        State &st = states[exec.slot(t)];
        const int q = gx * T + t;
        a.psum[static_cast<size_t>(gy) * a.Qpad + q] = st.sum;
        a.phits[static_cast<size_t>(gy) * a.Qpad + q] = st.hits;
    });
}

// Stage B from the stash (plan->stash): no scoring — every thread re-reads its query's distance bytes / relevance bits
// written by stage A (coalesced: consecutive queries are consecutive words), 32 rows per iteration, and walks only the
// rows that can be in the top k (distance <= d*): their bit mask is built first and consumed lowest-bit-first, so a
// warp iterates max-over-lanes(popcount) times instead of once per row.  ALL (k >= database size): every row counts,
// the walk is a straight unrolled loop.  Grid / shared-memory counters exactly as hamming_walk_program PHASE 1.
__host__ __device__ __forceinline__ uint32_t stash_byte(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int i) {
    const uint32_t lo = (i & 8) ? w2 : w0, hi = (i & 8) ? w3 : w1;
#ifdef __CUDA_ARCH__
    return __byte_perm(lo, hi, static_cast<uint32_t>(i & 7)) & 0xffu;
#else
    return (((i & 4) ? hi : lo) >> (8 * (i & 3))) & 0xffu;
#endif
}
__host__ __device__ __forceinline__ int lowest_bit(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs(static_cast<int>(m)) - 1;
#else
    return __builtin_ctz(m);
#endif
}

template <bool WIDE, bool ALL, typename Exec>
__host__ __device__ __forceinline__ void hamming_rank_program(const MapArgs &a, int gx, int gy, int T, unsigned char *smem,
                                                              Exec exec) {
    using C = Ctr<WIDE>;
    using ctr_t = typename C::type;
    ctr_t *cnt = reinterpret_cast<ctr_t *>(smem);
    const int seg_begin = gy * a.seg_len;                 // multiple of 32 in stash mode
    const int seg_end = seg_begin + a.seg_len < a.N ? seg_begin + a.seg_len : a.N;
    const ctr_t *hist_seg = static_cast<const ctr_t *>(a.hist) + static_cast<size_t>(gy) * a.bins * a.Qpad + static_cast<size_t>(gx) * T;
    exec([&](int t, int) {
        const int q = gx * T + t;
        const uint32_t k = a.k;
        const uint32_t dstar = a.dstar[q];
        const bool emit = (a.rank_idx != nullptr || a.rank_dist != nullptr) && q < a.Q;
        // ALL: 16|16-bit running counts of this segment bumped by one shared atomic per row, bases fetched from the
        // (L2-resident) global histogram for relevant rows only — as in hamming_walk_program<ALL = true>
        uint32_t *run = reinterpret_cast<uint32_t *>(smem);
        if (ALL) {
            for (int d = 0; d < a.bins; ++d) run[d * T + t] = 0u;
        } else {
            for (int d = 0; d < a.bins; ++d) cnt[d * T + t] = hist_seg[static_cast<size_t>(d) * a.Qpad + t];
        }
        unsigned long long sum = 0;
        uint32_t hits = 0;
        auto visit = [&](uint32_t d, bool rel, int row) {
            if (ALL) {
                const uint32_t o = ctr_fetch_add32(run + d * T + t, 1u + (static_cast<uint32_t>(rel) << 16));
                if (rel || emit) {
                    const ctr_t b = hist_seg[static_cast<size_t>(d) * a.Qpad + t];
                    const uint32_t rank = C::lo(b) + (o & 0xffffu) + 1u;
                    if (rel) {
                        sum += ap_term(C::hi(b) + (o >> 16) + 1u, rank);
                        ++hits;
                    }
                    if (emit) {
                        const size_t oo = static_cast<size_t>(q) * k + (rank - 1u);
                        if (a.rank_idx) a.rank_idx[oo] = static_cast<uint32_t>(a.index_base + row);
                        if (a.rank_dist) a.rank_dist[oo] = static_cast<uint16_t>(d);
                    }
                }
                return;
            }
            ctr_t c = cnt[d * T + t];
            const uint32_t rank = C::lo(c) + 1u;
            if (rank <= k) {
                c += static_cast<ctr_t>(1) + (static_cast<ctr_t>(rel) << C::kShift);
                cnt[d * T + t] = c;
                if (rel) {
                    sum += ap_term(C::hi(c), rank);
                    ++hits;
                }
                if (emit) {
                    const size_t o = static_cast<size_t>(q) * k + (rank - 1u);
                    if (a.rank_idx) a.rank_idx[o] = static_cast<uint32_t>(a.index_base + row);
                    if (a.rank_dist) a.rank_dist[o] = static_cast<uint16_t>(d);
                }
            }
        };
        const size_t Qp = static_cast<size_t>(a.Qpad);
        // the stash words of the next 32-row group are requested before the current group is walked (the reads come
        // from DRAM / L2 at 600+ cycles; the walk itself is short)
        auto fetch = [&](int row0, uint32_t (&w)[8], uint32_t &relw) {
            const int nvalid = seg_end - row0;
            const U32x4 *pd = a.stash_d + static_cast<size_t>(row0 >> 4) * Qp + q;
            const U32x4 none{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            const U32x4 v0 = nvalid > 0 ? pd[0] : none, v1 = nvalid > 16 ? pd[Qp] : none;      // bytes past the end are never visited
            w[0] = v0.x, w[1] = v0.y, w[2] = v0.z, w[3] = v0.w, w[4] = v1.x, w[5] = v1.y, w[6] = v1.z, w[7] = v1.w;
            relw = nvalid > 0 ? a.stash_r[static_cast<size_t>(row0 >> 5) * Qp + q] : 0u;
        };
        uint32_t wn[8], relwn;
        fetch(seg_begin, wn, relwn);
        for (int row0 = seg_begin; row0 < seg_end; row0 += 32) {
            const int nvalid = seg_end - row0 < 32 ? seg_end - row0 : 32;
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = wn[i];
            const uint32_t relw = relwn;
            fetch(row0 + 32, wn, relwn);
            if (ALL) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nvalid) visit((w[i >> 2] >> (8 * (i & 3))) & 0xffu, (relw >> i) & 1u, row0 + i);
            } else {
                uint32_t mask = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) mask |= static_cast<uint32_t>(((w[i >> 2] >> (8 * (i & 3))) & 0xffu) <= dstar) << i;
                if (nvalid < 32) mask &= (1u << nvalid) - 1u;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t m = (mask >> (16 * half)) & 0xffffu;
                    while (m) {
                        const int i = lowest_bit(m);
                        m &= m - 1u;
                        const uint32_t d = stash_byte(w[4 * half], w[4 * half + 1], w[4 * half + 2], w[4 * half + 3], i);
                        visit(d, (relw >> (16 * half + i)) & 1u, row0 + 16 * half + i);
                    }
                }
            }
        }
        a.psum[static_cast<size_t>(gy) * a.Qpad + q] = sum;
        a.phits[static_cast<size_t>(gy) * a.Qpad + q] = hits;
    });
}

// Stage S.  CTA = 32 queries (tx) x kScanY bucket lanes (ty); thread id t = ty * 32 + tx.
//  (1) bucket totals: own segments, or every shard's gathered totals (+ the rows of earlier shards in the bucket)
//  (2) the ty == 0 warp scans the distances per query and finds d* (first bucket whose cumulative size reaches k)
//  (3) exclusive scan over the segments inside each bucket, stored in place as saturated (rank base, ordinal base)
template <bool WIDE, typename Exec>
__host__ __device__ __forceinline__ void hamming_scan_program(void *hist_, int S, int bins, int Qpad, uint32_t k,
                                                              const U32x2 *ext, int n_shards, int shard, uint32_t *dstar,
                                                              int gx, unsigned char *smem, Exec exec) {
    using C = Ctr<WIDE>;
    using ctr_t = typename C::type;
    U32x2 *s_tot = reinterpret_cast<U32x2 *>(smem);                 // [bins][32] bucket totals
    U32x2 *s_start = s_tot + static_cast<size_t>(bins) * kScanQ;    // [bins][32] earlier-shard rows, then bucket start
    ctr_t *hist = static_cast<ctr_t *>(hist_);
    const size_t plane = static_cast<size_t>(bins) * Qpad;

    exec([&](int t, int) {
        const int tx = t % kScanQ, ty = t / kScanQ;
        const int q = gx * kScanQ + tx;
        for (int d = ty; d < bins; d += kScanY) {
            uint32_t a = 0, r = 0, ba = 0, br = 0;
            if (ext == nullptr) {
                const ctr_t *col = hist + static_cast<size_t>(d) * Qpad + q;
                int s = 0;
                for (; s + 8 <= S; s += 8) {          // 8 independent loads in flight (planes are megabytes apart)
                    ctr_t c[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) c[i] = col[static_cast<size_t>(s + i) * plane];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a += C::lo(c[i]), r += C::hi(c[i]);
                }
                for (; s < S; ++s) {
                    const ctr_t c = col[static_cast<size_t>(s) * plane];
                    a += C::lo(c), r += C::hi(c);
                }
            } else {
                for (int p = 0; p < n_shards; ++p) {
                    const U32x2 v = ext[static_cast<size_t>(p) * plane + static_cast<size_t>(d) * Qpad + q];
                    a += v.x, r += v.y;
                    if (p < shard) ba += v.x, br += v.y;
                }
            }
            s_tot[d * kScanQ + tx] = U32x2{a, r};
            s_start[d * kScanQ + tx] = U32x2{ba, br};
        }
    });
    exec([&](int t, int) {
        const int tx = t % kScanQ, ty = t / kScanQ;
        if (ty != 0) return;
        uint32_t A = 0, R = 0, ds = static_cast<uint32_t>(bins - 1);
        bool found = false;
        for (int d = 0; d < bins; ++d) {
            const U32x2 tot = s_tot[d * kScanQ + tx];
            const U32x2 before = s_start[d * kScanQ + tx];
            s_start[d * kScanQ + tx] = U32x2{A + before.x, R + before.y};
            A += tot.x, R += tot.y;
            if (!found && A >= k) ds = static_cast<uint32_t>(d), found = true;
        }
        dstar[gx * kScanQ + tx] = ds;
    });
    exec([&](int t, int) {
        const int tx = t % kScanQ, ty = t / kScanQ;
        const int q = gx * kScanQ + tx;
        for (int d = ty; d < bins; d += kScanY) {
            U32x2 run = s_start[d * kScanQ + tx];
            ctr_t *col = hist + static_cast<size_t>(d) * Qpad + q;
            int s = 0;
            for (; s + 8 <= S; s += 8) {              // load 8 segments' counts, then overwrite them with their bases
                ctr_t c[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = col[static_cast<size_t>(s + i) * plane];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // the part of a bucket that starts at rank base >= k is dead: rank = base + 1 > k
                    col[static_cast<size_t>(s + i) * plane] = C::make(run.x < k ? run.x : k, run.x < k ? run.y : 0u);
                    run.x += C::lo(c[i]), run.y += C::hi(c[i]);
                }
            }
            for (; s < S; ++s) {
                ctr_t *p = col + static_cast<size_t>(s) * plane;
                const ctr_t c = *p;
                *p = C::make(run.x < k ? run.x : k, run.x < k ? run.y : 0u);
                run.x += C::lo(c), run.y += C::hi(c);
            }
        }
    });
}

// Shard totals tot[d][q] = sum_s hist[s][d][q] as (rows, relevant) pairs: the block a sharded run all-gathers.
template <bool WIDE>
__host__ __device__ __forceinline__ void hamming_totals_item(const void *hist_, int S, size_t plane, size_t i, U32x2 *tot) {
    using C = Ctr<WIDE>;
    const typename C::type *hist = static_cast<const typename C::type *>(hist_);
    uint32_t a = 0, r = 0;
    for (int s = 0; s < S; ++s) {
        const typename C::type c = hist[static_cast<size_t>(s) * plane + i];
        a += C::lo(c), r += C::hi(c);
    }
    tot[i] = U32x2{a, r};
}

// AP_q = (sum of partials) / hits; 0 without a hit (accuracy_calculator.py:226-229).  Integer sums: order-free.
__host__ __device__ __forceinline__ void ap_finalize_item(const unsigned long long *psum, const uint32_t *phits, int parts,
                                                          long long stride, int q, double *ap, uint32_t *tsum) {
    unsigned long long s = 0;
    uint32_t h = 0;
    for (int i = 0; i < parts; ++i) {
        s += psum[static_cast<size_t>(i) * stride + q];
        h += phits[static_cast<size_t>(i) * stride + q];
    }
    ap[q] = h ? (static_cast<double>(s) / 1099511627776.0) / static_cast<double>(h) : 0.0;
    if (tsum) tsum[q] = h;
}
__host__ __device__ __forceinline__ void ap_reduce_item(const unsigned long long *psum, const uint32_t *phits, int S, int Qpad,
                                                        int q, unsigned long long *sum_q, uint32_t *hits_q) {
    unsigned long long s = 0;
    uint32_t h = 0;
    for (int i = 0; i < S; ++i) {
        s += psum[static_cast<size_t>(i) * Qpad + q];
        h += phits[static_cast<size_t>(i) * Qpad + q];
    }
    sum_q[q] = s, hits_q[q] = h;
}

}  // namespace b200
