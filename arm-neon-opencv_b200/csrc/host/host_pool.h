// Host storage for vision::Tensor in libvacv.so.  The reference allocates every Tensor with cudaHostAlloc when built with
// USE_CUDA (src/common/va_allocator.cpp:12-31, va_cuda_allocator.cu:8-34) and otherwise with malloc; page-locked memory is
// what lets a host<->device copy run at full PCIe speed, but cudaHostAlloc / cudaFreeHost themselves cost 0.1-1 ms.  This
// pool keeps freed page-locked blocks in power-of-two size classes and hands them out again, so a steady-state caller
// (one frame after another, same shapes) never reaches the driver.  Small tensors (statistics, matrices) use malloc.
// All CUDA access goes through the C-ABI runtime helpers (vacv_cuda_host_alloc / vacv_cuda_host_free).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

#include "vacv_cuda.h"

namespace vacv_host {

class HostPool {
public:
    static HostPool& instance() { static HostPool* p = new HostPool; return *p; }   // leaked on purpose: tensors may outlive static destructors

    // returns a pointer to `bytes` usable bytes, 64-byte aligned
    void* allocate(size_t bytes) {
        const size_t need = bytes + kHeader;
        if (need >= kPinnedMin) {
            const int cls = size_class(need);
            {
                std::lock_guard<std::mutex> g(mu_);
                auto& fl = free_[cls];
                if (!fl.empty()) { void* b = fl.back(); fl.pop_back(); cached_ -= (size_t)1 << cls; return payload(b); }
            }
            void* b = nullptr;
            if (vacv_cuda_host_alloc(&b, (size_t)1 << cls) == VACV_OK && b) {
                header(b)->cls = cls;
                return payload(b);
            }
            // no device / out of page-locked memory: pageable storage still works (copies are just slower)
        }
        void* b = std::malloc(need);
        if (!b) throw std::bad_alloc();
        header(b)->cls = 0;
        return payload(b);
    }

    void release(void* p) {
        if (!p) return;
        void* b = static_cast<unsigned char*>(p) - kHeader;
        const int cls = header(b)->cls;
        if (cls == 0) { std::free(b); return; }
        {
            std::lock_guard<std::mutex> g(mu_);
            if (cached_ + ((size_t)1 << cls) <= kMaxCached) { free_[cls].push_back(b); cached_ += (size_t)1 << cls; return; }
        }
        vacv_cuda_host_free(b);
    }

private:
    struct Header { int cls; };                               // 0 = malloc, else log2 of the page-locked block size
    static constexpr size_t kHeader = 64;                     // keeps the payload 64-byte aligned
    static constexpr size_t kPinnedMin = 64 * 1024;           // below this the copy is latency-bound anyway
    static constexpr size_t kMaxCached = (size_t)1 << 30;     // page-locked bytes kept for reuse
    static Header* header(void* b) { return static_cast<Header*>(b); }
    static void* payload(void* b) { return static_cast<unsigned char*>(b) + kHeader; }
    static int size_class(size_t n) { int c = 16; while (((size_t)1 << c) < n) ++c; return c; }
    HostPool() : free_(48) {}
    std::mutex mu_;
    std::vector<std::vector<void*>> free_;
    size_t cached_ = 0;
};

}  // namespace vacv_host
