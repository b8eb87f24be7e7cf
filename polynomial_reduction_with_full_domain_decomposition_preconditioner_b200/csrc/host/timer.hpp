// timer.hpp -- named phase accumulators with the reference's keys and report
// (/root/reference/timer.hpp:14-38, timer.tpp:48-125).
//
// The reference's start()/stop() each do device.finish() + MPI_Barrier / MPI_Allreduce(MAX)
// (timer.tpp:48-68), i.e. every timed phase of the hot loop is fenced.  Here timing is OFF by
// default (start/stop are no-ops and nothing synchronises); when enabled each phase is bracketed
// by a stream synchronise so the per-phase seconds are real device time, as in the reference.
// Keys are compared by VALUE (the reference keys its maps by `const char*` pointer identity).
#pragma once
#include <chrono>
#include <cuda_runtime.h>
#include <string>
#include <unordered_map>
#include <vector>

template <typename DType = double>
class Timer
{
    std::unordered_map<std::string, std::chrono::high_resolution_clock::time_point> t_start;
    std::unordered_map<std::string, DType> t_total;

  public:
    bool enabled = false;
    cudaStream_t stream = nullptr;

    void initialize() { t_total.clear(); }

    void start(const char *name, bool = true)
    {
        if (!enabled) return;
        cudaStreamSynchronize(stream);
        t_start[name] = std::chrono::high_resolution_clock::now();
    }

    void stop(const char *name, bool = true)
    {
        if (!enabled) return;
        cudaStreamSynchronize(stream);
        auto now = std::chrono::high_resolution_clock::now();
        t_total[name] += std::chrono::duration_cast<std::chrono::duration<DType>>(now - t_start[name]).count();
    }

    void reset(const char *name) { t_total[name] = 0.0; }

    DType total(const char *name) const
    {
        auto it = t_total.find(name);
        return it == t_total.end() ? (DType)(-1.0) : it->second;
    }

    std::vector<std::string> keys() const
    {
        std::vector<std::string> k;
        for (auto &kv : t_total) k.push_back(kv.first);
        return k;
    }
};
