// Library-level C-ABI plumbing: version, error strings, launch counter.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace b200 {

static thread_local char g_last_cuda_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_last_cuda_error(cudaError_t e, const char *where) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", where, cudaGetErrorString(e),
             cudaGetErrorName(e));
    cudaGetLastError();   // clear the sticky-free error state so the next call starts clean
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

StageMarks &stage_marks() {
    static thread_local StageMarks m;
    return m;
}

int sm_count() {
    static std::atomic<int> cached{0};
    int v = cached.load(std::memory_order_relaxed);
    if (v == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            v <= 0)
            v = kNumSMs;
        cached.store(v, std::memory_order_relaxed);
    }
    return v;
}

}  // namespace b200

extern "C" {

int b200_version(void) { return 200; }   // 0.2.0

size_t b200_sizeof_map_plan(void) { return sizeof(b200_map_plan); }

const char *b200_error_string(int code) {
    switch (code) {
        case B200_OK: return "ok";
        case B200_ERR_INVALID_ARG: return "invalid argument";
        case B200_ERR_UNSUPPORTED: return "unsupported configuration";
        case B200_ERR_CUDA: return "CUDA runtime error";
        case B200_ERR_WORKSPACE: return "workspace missing or too small";
        case B200_ERR_ALIGNMENT: return "misaligned pointer";
        case B200_ERR_NO_DEVICE: return "no usable sm_100 device";
        default: return "unknown error code";
    }
}

const char *b200_last_cuda_error(void) { return b200::g_last_cuda_error; }

unsigned long long b200_launch_count(void) { return b200::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
