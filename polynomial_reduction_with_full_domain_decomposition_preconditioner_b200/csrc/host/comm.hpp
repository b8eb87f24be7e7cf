// comm.hpp -- the communicator that stands where MPI + GSLib stand in the reference.
//
// One process per GPU; collectives are NCCL over NVLink/NVSwitch on the solver's stream, entirely
// in device memory (the reference stages every message through the host: domain.tpp:590-594,
// subdomain.tpp:4615-4635).  NCCL is resolved with dlopen at run time so a single-GPU run needs no
// NCCL at all, and a multi-GPU run uses whichever libnccl.so.2 the process already loaded (torch's).
//   MPI_Allreduce(SUM, 1-2 doubles)   domain.tpp:929,946,969,995  -> allreduce_sum
//   MPI_Allgatherv                    subdomain.tpp:4620-4621     -> allgather
//   gslib_gs(gs_add)                  domain.tpp:592              -> Domain's halo lists + sendrecv
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <vector>

class Comm
{
  public:
    int rank = 0;
    int size = 1;
    cudaStream_t stream = nullptr;

    void init(int rank_, int size_, const void *nccl_unique_id, cudaStream_t stream_);
    void finalize();
    bool active() const { return size > 1; }

    // in-place sum over ranks of `count` doubles in device memory
    void allreduce_sum(double *dptr, int count);
    // every rank contributes `count` doubles (same count everywhere); recv has size*count
    void allgather(const double *send, double *recv, size_t count);
    // one grouped exchange: for each peer p, send send_count[p] doubles from send_ptr[p] and receive
    // recv_count[p] doubles into recv_ptr[p]
    void sendrecv(const std::vector<int> &peers, const std::vector<const double *> &send_ptr, const std::vector<size_t> &send_count,
                  const std::vector<double *> &recv_ptr, const std::vector<size_t> &recv_count);
    // host-side helpers for setup (staged through device buffers)
    void allgather_host(const void *send, void *recv, size_t bytes_per_rank);
    long long allreduce_sum_host(long long v);
    long long allreduce_max_host(long long v);
    void barrier();

  private:
    void *comm_ = nullptr; // ncclComm_t
};
