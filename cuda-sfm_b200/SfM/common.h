// Source-compatible replacement of the reference's SfM/common.h (index macros,
// error check, ilog2, PerformanceTimer) for code that includes it next to the
// B200 facade.  Reference: SfM/common.h:17-132, SfM/common.cu:3-15.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>

#define FILENAME (strrchr(__FILE__, '/') ? strrchr(__FILE__, '/') + 1 : __FILE__)
// The reference declares a one-argument macro but calls it with two
// (sfm.h:13); variadic here so both spellings compile on every compiler.
#define checkCUDAError(msg, ...) checkCUDAErrorFn(msg, FILENAME, __LINE__)
// Row-major element (i, j) of a matrix with `col` columns, and element (i, j)
// of matrix k in a batch of row x col matrices (SfM/common.h:19-20), with the
// arguments parenthesised.
#define access2(i, j, col) ((i) * (col) + (j))
#define access3(i, j, k, row, col) ((k) * (row) * (col) + (i) * (col) + (j))

// Prints and exits like the reference (common.cu:3-15); the C ABI underneath
// returns status codes instead.
inline void checkCUDAErrorFn(const char* msg, const char* file = NULL, int line = -1) {
    cudaError_t err = cudaGetLastError();
    if (cudaSuccess == err) return;
    fprintf(stderr, "CUDA error");
    if (file) fprintf(stderr, " (%s:%d)", file, line);
    fprintf(stderr, ": %s: %s\n", msg, cudaGetErrorString(err));
    exit(EXIT_FAILURE);
}

inline int ilog2(int x) {
    int lg = 0;
    while (x >>= 1) ++lg;
    return lg;
}
inline int ilog2ceil(int x) { return x == 1 ? 0 : ilog2(x - 1) + 1; }

namespace Common {
// Stopwatch pair with the reference's method names (SfM/common.h:48-132): a host
// clock and a pair of CUDA events on the legacy default stream.  Misuse (start
// twice, stop without start) throws std::runtime_error like the reference.
class PerformanceTimer {
    using clock = std::chrono::steady_clock;
    enum Which { CPU = 0, GPU = 1 };
    bool running_[2] = {false, false};
    float last_ms_[2] = {0.f, 0.f};
    clock::time_point cpu_begin_;
    cudaEvent_t gpu_begin_ = nullptr, gpu_end_ = nullptr;

    void arm(Which w, const char* what) {
        if (running_[w]) throw std::runtime_error(std::string(what) + " timer already started");
        running_[w] = true;
    }
    void disarm(Which w, const char* what) {
        if (!running_[w]) throw std::runtime_error(std::string(what) + " timer not started");
        running_[w] = false;
    }

public:
    PerformanceTimer() { cudaEventCreate(&gpu_begin_); cudaEventCreate(&gpu_end_); }
    ~PerformanceTimer() { cudaEventDestroy(gpu_begin_); cudaEventDestroy(gpu_end_); }
    PerformanceTimer(const PerformanceTimer&) = delete;
    PerformanceTimer& operator=(const PerformanceTimer&) = delete;

    void startCpuTimer() { arm(CPU, "CPU"); cpu_begin_ = clock::now(); }
    void endCpuTimer() {
        const auto stop = clock::now();
        disarm(CPU, "CPU");
        last_ms_[CPU] = std::chrono::duration<float, std::milli>(stop - cpu_begin_).count();
    }
    void startGpuTimer() { arm(GPU, "GPU"); cudaEventRecord(gpu_begin_); }
    void endGpuTimer() {
        cudaEventRecord(gpu_end_);
        cudaEventSynchronize(gpu_end_);
        disarm(GPU, "GPU");
        cudaEventElapsedTime(&last_ms_[GPU], gpu_begin_, gpu_end_);
    }
    float getCpuElapsedTimeForPreviousOperation() { return last_ms_[CPU]; }
    float getGpuElapsedTimeForPreviousOperation() { return last_ms_[GPU]; }
};
}  // namespace Common
