// Throughput of __match_any_sync against the ballot loop that replaces it (rank kernel of the select pipeline):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_match tools/ubench_match.cu && tools/ubench_match
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_match(unsigned *out, int iters, unsigned seed) {
    unsigned v = (threadIdx.x * 2654435761u + seed) >> 28, acc = 0;          // ~16 distinct values per warp
    for (int i = 0; i < iters; ++i) {
        acc += __match_any_sync(0xffffffffu, v);
        v = (v * 5u + acc) & 15u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_ballot_loop(unsigned *out, int iters, unsigned seed) {
    unsigned v = (threadIdx.x * 2654435761u + seed) >> 28, acc = 0;
    const unsigned lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
        unsigned remaining = 0xffffffffu, peers = 0;
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const unsigned dv = __shfl_sync(0xffffffffu, v, leader);
            const unsigned grp = __ballot_sync(0xffffffffu, v == dv);
            if (v == dv) peers = grp;
            remaining &= ~grp;
        }
        acc += peers;
        v = (v * 5u + acc + lane) & 15u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    unsigned *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    const int iters = 4096;
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (which == 0) k_match<<<148 * 8, 256>>>(out, iters, rep); else k_ballot_loop<<<148 * 8, 256>>>(out, iters, rep);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            const double warp_ops = 148.0 * 8 * 8 * iters;
            if (rep) printf("%s: %.3f ms, %.1f cycles per warp-op per SM-scheduler slot (at 1.965 GHz, 64 warps/SM resident)\n",
                            which ? "ballot loop (16 values)" : "match_any", ms, ms * 1e-3 * 1.965e9 * 148 * 4 / warp_ops);
        }
    }
    return 0;
}
