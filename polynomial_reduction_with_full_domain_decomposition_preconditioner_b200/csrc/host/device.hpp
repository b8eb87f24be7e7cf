// device.hpp -- the thin CUDA buffer / device layer that stands where OCCA stands in the reference.
//
// Mirrors exactly the occa::device / occa::memory calls the reference makes:
// device.malloc<T>(count), memory.copyFrom / copyTo (host pointer or memory), slice(offset, count)
// in ELEMENTS, ptr(), free(), device.finish().  Slices alias their parent (ref-counted), like
// occa::memory.  Everything runs on ONE explicit stream per GPU.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>

namespace dev
{
inline void check(cudaError_t e, const char *what)
{
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

inline void check_rc(int rc, const char *what)
{
    if (rc != 0)
        throw std::runtime_error(std::string("prfdd kernel launch failed in ") + what + " (code " + std::to_string(rc) + ")");
}

class memory
{
    struct Block
    {
        void *p = nullptr;
        ~Block() { if (p) cudaFree(p); }
    };
    std::shared_ptr<Block> block_;
    char *ptr_ = nullptr;
    size_t bytes_ = 0;
    size_t elem_ = 1;
    cudaStream_t stream_ = nullptr;

  public:
    memory() {}
    memory(size_t count, size_t elem_size, cudaStream_t stream) : bytes_(count * elem_size), elem_(elem_size), stream_(stream)
    {
        block_ = std::make_shared<Block>();
        if (bytes_ > 0)
        {
            check(cudaMalloc(&block_->p, bytes_), "device.malloc");
            check(cudaMemsetAsync(block_->p, 0, bytes_, stream_), "device.malloc/memset");
        }
        ptr_ = (char *)block_->p;
    }

    bool is_initialized() const { return (bool)block_; }
    size_t size() const { return bytes_; }
    size_t length() const { return elem_ ? bytes_ / elem_ : 0; }
    void *ptr() const { return ptr_; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr_); }
    void free() { block_.reset(); ptr_ = nullptr; bytes_ = 0; }

    // offsets and counts in elements of the allocation's type, like occa::memory::slice
    memory slice(size_t offset, size_t count) const
    {
        memory m(*this);
        m.ptr_ = ptr_ + offset * elem_;
        m.bytes_ = count * elem_;
        return m;
    }

    void copyFrom(const void *host, size_t bytes) const
    {
        if (bytes == 0) return;
        check(cudaMemcpyAsync(ptr_, host, bytes, cudaMemcpyHostToDevice, stream_), "memory.copyFrom(host)");
        check(cudaStreamSynchronize(stream_), "memory.copyFrom(host)/sync");
    }
    void copyFrom(const memory &src, size_t bytes) const
    {
        if (bytes == 0) return;
        check(cudaMemcpyAsync(ptr_, src.ptr_, bytes, cudaMemcpyDeviceToDevice, stream_), "memory.copyFrom(memory)");
    }
    void copyTo(void *host, size_t bytes) const
    {
        if (bytes == 0) return;
        check(cudaMemcpyAsync(host, ptr_, bytes, cudaMemcpyDeviceToHost, stream_), "memory.copyTo(host)");
        check(cudaStreamSynchronize(stream_), "memory.copyTo(host)/sync");
    }
    void copyTo(const memory &dst, size_t bytes) const
    {
        if (bytes == 0) return;
        check(cudaMemcpyAsync(dst.ptr_, ptr_, bytes, cudaMemcpyDeviceToDevice, stream_), "memory.copyTo(memory)");
    }
};

class device
{
  public:
    cudaStream_t stream = nullptr;
    int id = 0;

    void setup(int device_id, cudaStream_t s)
    {
        id = device_id;
        stream = s;
    }
    template <typename T> memory malloc(size_t count) const { return memory(count, sizeof(T), stream); }
    void finish() const { check(cudaStreamSynchronize(stream), "device.finish"); }
};
} // namespace dev
