// common.cuh -- shared helpers for the sm_100a kernels of libprfdd_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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

// Programmatic dependent launch (PDL).  A kernel launched with launch_pdl may be scheduled while its predecessor on the stream is
// still running; it must call pdl_wait() before its first global-memory access (the call returns when the predecessor has
// completed and its writes are visible), after which it lets ITS successor start launching (pdl_wait does both).  In a kernel
// launched the ordinary way both instructions are no-ops.  Inside a captured graph the dependency becomes a programmatic edge.
// What it buys: the launch latency and the ramp-up of a kernel overlap the tail of the one before it -- the V-cycle is 26
// dependent launches of 10-45 us each.  PRFDD_PDL=0 turns it off (ordinary launches).
__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled()
{
    static const bool on = !(getenv("PRFDD_PDL") && atoi(getenv("PRFDD_PDL")) == 0);
    return on;
}

template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
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
