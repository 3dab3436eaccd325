// Per-thread GPU context of the C++ host layer: one stream and a few grow-only device scratch buffers, all obtained
// through the C-ABI runtime helpers (include/vacv_cuda.h) -- the host layer includes no CUDA header.
// Every failure raises std::runtime_error with the C-ABI's message: there is no CPU fallback.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>

#include "vacv_cuda.h"

namespace vacv_host {

class DeviceContext {
public:
    static DeviceContext& current() {
        static thread_local DeviceContext ctx;
        return ctx;
    }
    void* stream() { return stream_; }
    void check(int status) {
        if (status != VACV_OK) throw std::runtime_error(std::string("vacv: ") + vacv_cuda_last_error());
    }
    // device scratch slot `i` of at least `bytes` (contents undefined)
    void* scratch(int i, size_t bytes) {
        Slot& s = slots_[i];
        if (s.cap < bytes) {
            check(vacv_cuda_stream_sync(stream_));
            if (s.p) check(vacv_cuda_free(s.p));
            s.p = nullptr; s.cap = 0;
            const size_t want = bytes + bytes / 4 + 256;
            check(vacv_cuda_malloc(&s.p, want));
            s.cap = want;
        }
        return s.p;
    }
    void* upload(int i, const void* host, size_t bytes) {
        void* d = scratch(i, bytes);
        check(vacv_cuda_memcpy_h2d(d, host, bytes, stream_));
        return d;
    }
    // copy back and wait: the reference API is synchronous
    void download(void* host, const void* dev, size_t bytes) {
        check(vacv_cuda_memcpy_d2h(host, dev, bytes, stream_));
        check(vacv_cuda_stream_sync(stream_));
    }
    void sync() { check(vacv_cuda_stream_sync(stream_)); }

private:
    struct Slot { void* p = nullptr; size_t cap = 0; };
    static constexpr int kSlots = 6;
    DeviceContext() {
        int n = 0;
        if (vacv_cuda_device_count(&n) != VACV_OK || n <= 0)
            throw std::runtime_error("vacv: no CUDA device available (libvacv has no CPU fallback)");
        check(vacv_cuda_stream_create(&stream_));
    }
    ~DeviceContext() {
        for (Slot& s : slots_) if (s.p) vacv_cuda_free(s.p);
        if (stream_) vacv_cuda_stream_destroy(stream_);
    }
    DeviceContext(const DeviceContext&) = delete;
    void* stream_ = nullptr;
    Slot slots_[kSlots];
};

}  // namespace vacv_host
