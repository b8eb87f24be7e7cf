// globals.cpp -- definitions of the process-wide globals of config.hpp and the small library-level
// C ABI (version, device selection, buffers).
#include "config.hpp"
#include "../../../include/prfdd_b200.h"
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace prfdd_host
{
int dim = 0;
int proc_id = 0;
int num_procs = 1;
int verbose = 0;
dev::device device;
Timer<double> timer;
Comm comm_world;
FILE *pstdout_file = nullptr;
} // namespace prfdd_host

extern "C" {

const char *prfdd_version(void) { return "prfdd_b200 0.1 (sm_100a)"; }

const char *prfdd_error_string(int code)
{
    switch (code)
    {
    case 0: return "success";
    case -1: return "exception in host code (see stderr)";
    case -2: return "missing reduction workspace";
    case -3: return "too many vectors in a fused operation (max 32)";
    case -4: return "unsupported number of GLL points (2..16)";
    case -5: return "could not read the derivative matrix from device memory";
    case -6: return "threads_per_row must be one of 1,2,4,8,16,32";
    case -7: return "row_end < row_start";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown prfdd error";
    }
}

int prfdd_halo_build_lists(int me, int nprocs, const long long *ids, const long long *offsets, int *peers, int *peer_count, int *peer_offset, int *idx,
                           long long idx_capacity, long long *total)
{
    std::unordered_map<long long, int> mine;
    for (long long i = offsets[me]; i < offsets[me + 1]; i++) mine[ids[i]] = (int)(i - offsets[me]);
    int np = 0;
    long long n = 0;
    for (int p = 0; p < nprocs; p++)
    {
        if (p == me) continue;
        std::vector<std::pair<long long, int>> shared;
        for (long long i = offsets[p]; i < offsets[p + 1]; i++)
        {
            auto it = mine.find(ids[i]);
            if (it != mine.end()) shared.push_back({ids[i], it->second});
        }
        if (shared.empty()) continue;
        std::sort(shared.begin(), shared.end());
        if (n + (long long)shared.size() > idx_capacity) return -8;
        peers[np] = p;
        peer_offset[np] = (int)n;
        peer_count[np] = (int)shared.size();
        for (auto &s : shared) idx[n++] = s.second;
        np++;
    }
    *total = n;
    return np;
}

int prfdd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int prfdd_set_device(int d) { return (int)cudaSetDevice(d); }
int prfdd_malloc(void **dptr, size_t bytes) { return (int)cudaMalloc(dptr, bytes); }
int prfdd_free(void *dptr) { return (int)cudaFree(dptr); }
int prfdd_memcpy_h2d(void *dst, const void *src, size_t bytes, prfdd_stream_t s) { return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s); }
int prfdd_memcpy_d2h(void *dst, const void *src, size_t bytes, prfdd_stream_t s) { return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s); }
int prfdd_memcpy_d2d(void *dst, const void *src, size_t bytes, prfdd_stream_t s) { return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s); }
int prfdd_stream_synchronize(prfdd_stream_t s) { return (int)cudaStreamSynchronize((cudaStream_t)s); }

} // extern "C"
