// comm.cpp -- NCCL-backed communicator (see comm.hpp).  NCCL entry points are resolved with
// dlopen/dlsym; types come from the system <nccl.h>.
#include "comm.hpp"
#include "device.hpp"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <stdexcept>
#include <string>

namespace
{
struct Nccl
{
    void *handle = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;

    void load()
    {
        if (handle) return;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
        {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) throw std::runtime_error("prfdd: num_procs > 1 needs NCCL, but libnccl.so.2 could not be loaded");
#define SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(handle, "nccl" #f)); if (!f) throw std::runtime_error("prfdd: NCCL symbol nccl" #f " missing");
        SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(AllGather) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    }
    void check(ncclResult_t r, const char *what)
    {
        if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + GetErrorString(r));
    }
} nccl;
} // namespace

void Comm::init(int rank_, int size_, const void *nccl_unique_id, cudaStream_t stream_)
{
    rank = rank_;
    size = size_;
    stream = stream_;
    comm_ = nullptr;
    if (size <= 1) return;
    if (!nccl_unique_id) throw std::runtime_error("prfdd: num_procs > 1 requires an ncclUniqueId");
    nccl.load();
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(&id, nccl_unique_id, sizeof(id));
    ncclComm_t c;
    nccl.check(nccl.CommInitRank(&c, size, id, rank), "ncclCommInitRank");
    comm_ = c;
}

void Comm::finalize()
{
    if (comm_) nccl.CommDestroy((ncclComm_t)comm_);
    comm_ = nullptr;
}

void Comm::allreduce_sum(double *dptr, int count)
{
    if (size <= 1) return;
    nccl.check(nccl.AllReduce(dptr, dptr, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)comm_, stream), "ncclAllReduce");
}

void Comm::allgather(const double *send, double *recv, size_t count)
{
    if (size <= 1)
    {
        if (send != recv) dev::check(cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, stream), "allgather/self");
        return;
    }
    nccl.check(nccl.AllGather(send, recv, count, ncclDouble, (ncclComm_t)comm_, stream), "ncclAllGather");
}

void Comm::sendrecv(const std::vector<int> &peers, const std::vector<const double *> &send_ptr, const std::vector<size_t> &send_count,
                    const std::vector<double *> &recv_ptr, const std::vector<size_t> &recv_count)
{
    if (size <= 1 || peers.empty()) return;
    nccl.check(nccl.GroupStart(), "ncclGroupStart");
    for (size_t k = 0; k < peers.size(); k++)
    {
        if (send_count[k]) nccl.check(nccl.Send(send_ptr[k], send_count[k], ncclDouble, peers[k], (ncclComm_t)comm_, stream), "ncclSend");
        if (recv_count[k]) nccl.check(nccl.Recv(recv_ptr[k], recv_count[k], ncclDouble, peers[k], (ncclComm_t)comm_, stream), "ncclRecv");
    }
    nccl.check(nccl.GroupEnd(), "ncclGroupEnd");
}

void Comm::allgather_host(const void *send, void *recv, size_t bytes_per_rank)
{
    if (size <= 1)
    {
        memcpy(recv, send, bytes_per_rank);
        return;
    }
    size_t padded = (bytes_per_rank + 7) / 8 * 8;
    void *ds = nullptr, *dr = nullptr;
    dev::check(cudaMalloc(&ds, padded), "allgather_host");
    dev::check(cudaMalloc(&dr, padded * size), "allgather_host");
    dev::check(cudaMemcpyAsync(ds, send, bytes_per_rank, cudaMemcpyHostToDevice, stream), "allgather_host");
    nccl.check(nccl.AllGather(ds, dr, padded, ncclChar, (ncclComm_t)comm_, stream), "ncclAllGather(host)");
    dev::check(cudaStreamSynchronize(stream), "allgather_host");
    for (int p = 0; p < size; p++)
        dev::check(cudaMemcpy((char *)recv + (size_t)p * bytes_per_rank, (char *)dr + (size_t)p * padded, bytes_per_rank, cudaMemcpyDeviceToHost), "allgather_host");
    cudaFree(ds);
    cudaFree(dr);
}

long long Comm::allreduce_sum_host(long long v)
{
    if (size <= 1) return v;
    std::vector<long long> all(size);
    allgather_host(&v, all.data(), sizeof(long long));
    long long s = 0;
    for (auto x : all) s += x;
    return s;
}

long long Comm::allreduce_max_host(long long v)
{
    if (size <= 1) return v;
    std::vector<long long> all(size);
    allgather_host(&v, all.data(), sizeof(long long));
    long long s = all[0];
    for (auto x : all) s = x > s ? x : s;
    return s;
}

void Comm::barrier()
{
    if (size <= 1) return;
    allreduce_sum_host(0);
}
