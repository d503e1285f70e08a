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
// CPU (chrono) and GPU (cudaEvent) stopwatch with the reference's method names.
class PerformanceTimer {
public:
    PerformanceTimer() {
        cudaEventCreate(&event_start);
        cudaEventCreate(&event_end);
    }
    ~PerformanceTimer() {
        cudaEventDestroy(event_start);
        cudaEventDestroy(event_end);
    }
    void startCpuTimer() {
        if (cpu_started) throw std::runtime_error("CPU timer already started");
        cpu_started = true;
        t0 = std::chrono::high_resolution_clock::now();
    }
    void endCpuTimer() {
        auto t1 = std::chrono::high_resolution_clock::now();
        if (!cpu_started) throw std::runtime_error("CPU timer not started");
        cpu_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
        cpu_started = false;
    }
    void startGpuTimer() {
        if (gpu_started) throw std::runtime_error("GPU timer already started");
        gpu_started = true;
        cudaEventRecord(event_start);
    }
    void endGpuTimer() {
        cudaEventRecord(event_end);
        cudaEventSynchronize(event_end);
        if (!gpu_started) throw std::runtime_error("GPU timer not started");
        cudaEventElapsedTime(&gpu_ms, event_start, event_end);
        gpu_started = false;
    }
    float getCpuElapsedTimeForPreviousOperation() { return cpu_ms; }
    float getGpuElapsedTimeForPreviousOperation() { return gpu_ms; }
    PerformanceTimer(const PerformanceTimer&) = delete;
    PerformanceTimer& operator=(const PerformanceTimer&) = delete;

private:
    cudaEvent_t event_start = nullptr, event_end = nullptr;
    std::chrono::high_resolution_clock::time_point t0;
    bool cpu_started = false, gpu_started = false;
    float cpu_ms = 0.f, gpu_ms = 0.f;
};
}  // namespace Common
