// Micro-benchmark of the integer pipes the Hamming kernels lean on (run on the B200 box via gpurun; not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int ubench_int.cu && ./ubench_int
// Prints lane-operations per clock per SM for POPC, LOP3, IADD3, IMAD, shared-memory RMW (LDS+STS) and RED.shared.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <int OP>
__global__ void __launch_bounds__(512) k_alu(uint32_t *out, uint32_t seed, long long *clk) {
    uint32_t r[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) r[i] = seed + threadIdx.x * 17 + i * 101;
    const uint32_t a = seed * 3 + 1, b = seed ^ 0x55aa55aa;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (OP == 0) r[i] = __popc(r[i]) + a;                     // POPC + IADD (IADD on another pipe)
            if (OP == 1) r[i] = (r[i] & a) ^ b;                       // LOP3
            if (OP == 2) r[i] = r[i] + a + b;                         // IADD3
            if (OP == 3) r[i] = r[i] * a + b;                         // IMAD
            if (OP == 4) r[i] = __popc(r[i] ^ a) + (__popc(r[i] ^ b) << 1);   // 2 POPC + 2 LOP3 + LEA
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// shared-memory counter update, private column per thread (bank = tid % 32): MODE 0 = LDS+IADD+STS, 1 = atomicAdd (RED)
template <int MODE>
__global__ void __launch_bounds__(512) k_smem(uint32_t *out, uint32_t seed, long long *clk, int bins) {
    extern __shared__ uint32_t cnt[];
    const int T = blockDim.x, t = threadIdx.x;
    for (int d = 0; d < bins; ++d) cnt[d * T + t] = 0;
    uint32_t x = seed + t * 2654435761u;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x = x * 1664525u + 1013904223u;
            const uint32_t d = (x >> 20) % bins;
            if (MODE == 0) cnt[d * T + t] += 1u + ((x & 1u) << 16);
            else atomicAdd(&cnt[d * T + t], 1u + ((x & 1u) << 16));
        }
    }
    long long t1 = clock64();
    __syncthreads();
    uint32_t s = 0;
    for (int d = 0; d < bins; ++d) s += cnt[d * T + t];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out;
    long long *clk, h[4096];
    cudaMalloc(&out, sizeof(uint32_t) * sms * 4 * 512);
    cudaMalloc(&clk, sizeof(long long) * sms * 4);
    const char *names[] = {"POPC(+IADD)", "LOP3", "IADD3", "IMAD", "2xPOPC+2xLOP3+LEA"};
    const double ops_per[] = {1, 1, 1, 1, 2};
    for (int op = 0; op < 5; ++op) {
        for (int warps = 4; warps <= 16; warps *= 2) {
            const int threads = warps * 32, blocks = sms * 4;      // 4 CTAs per SM
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
                    case 0: k_alu<0><<<blocks, threads>>>(out, 12345u + rep, clk); break;
                    case 1: k_alu<1><<<blocks, threads>>>(out, 12345u + rep, clk); break;
                    case 2: k_alu<2><<<blocks, threads>>>(out, 12345u + rep, clk); break;
                    case 3: k_alu<3><<<blocks, threads>>>(out, 12345u + rep, clk); break;
                    case 4: k_alu<4><<<blocks, threads>>>(out, 12345u + rep, clk); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < blocks; ++i) avg += h[i];
            avg /= blocks;
            const double lane_ops = 4.0 * threads * kIters * kChains * ops_per[op];   // per SM (4 CTAs resident)
            printf("%-20s warps/SM=%3d  %.1f lane-ops/clk/SM\n", names[op], warps * 4, lane_ops / avg);
        }
    }
    for (int mode = 0; mode < 2; ++mode) {
        for (int threads = 64; threads <= 256; threads *= 2) {
            const int bins = 65, blocks = sms * 2;
            const size_t smem = sizeof(uint32_t) * bins * threads;
            if (mode == 0) {
                cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                for (int rep = 0; rep < 2; ++rep) { k_smem<0><<<blocks, threads, smem>>>(out, 777u, clk, bins); cudaDeviceSynchronize(); }
            } else {
                cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                for (int rep = 0; rep < 2; ++rep) { k_smem<1><<<blocks, threads, smem>>>(out, 777u, clk, bins); cudaDeviceSynchronize(); }
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < blocks; ++i) avg += h[i];
            avg /= blocks;
            printf("%-20s warps/SM=%3d  %.2f counter updates/clk/SM\n", mode ? "RED.shared" : "LDS+STS RMW", threads / 32 * 2,
                   2.0 * threads * kIters * 4 / avg);
        }
    }
    return 0;
}
