// common.cuh -- shared helpers for the sm_100a kernels of libprfdd_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../include/prfdd_b200.h"

namespace prfdd
{
extern long long g_launch_count;
extern double g_algorithmic_bytes;

inline cudaStream_t S(prfdd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// every launcher ends with this: counts the launch and the ALGORITHMIC bytes of the kernel (what it must move at least: every
// operand once), reports launch-configuration errors.  The byte count feeds bench.py's whole-solve roofline figure.
inline int launched(double algorithmic_bytes = 0.0)
{
    ++g_launch_count;
    g_algorithmic_bytes += algorithmic_bytes;
    return (int)cudaGetLastError();
}

constexpr int kNumSMsB200 = 148;

inline int num_sms()
{
    static int n = 0;
    if (n == 0)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsB200;
    }
    return n;
}

// grid for a streaming element-wise kernel: whole waves of the SM count, capped by the work
inline int stream_grid(long long n, int threads, int per_thread, int waves_per_sm)
{
    long long need = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
    long long cap = (long long)num_sms() * waves_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
} // namespace prfdd

struct prfdd_reduce_ws
{
    double *partials;        // [kMaxK][kMaxBlocks]
    unsigned int *counter;   // ticket for "last block finishes the sum"
};
