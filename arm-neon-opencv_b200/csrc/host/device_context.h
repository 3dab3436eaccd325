// Per-thread, per-device GPU context of the C++ host layer: one stream and a few grow-only device scratch buffers, all
// obtained through the C-ABI runtime helpers (include/vacv_cuda.h) -- the host layer includes no CUDA header.
// A stream and device memory belong to the device they were created on, so the context is keyed by the CURRENT device:
// a thread that switches GPUs with vacv_cuda_set_device (the reference's CudaDevice::set_device, src/cv/cuda_device.cu:15-18)
// transparently gets that GPU's own stream and scratch.
// Every failure raises std::runtime_error with the C-ABI's message: there is no CPU fallback.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>

#include "vacv_cuda.h"

namespace vacv_host {

class DeviceContext {
public:
    static DeviceContext& current();
    static DeviceContext& current_impl(DeviceContext* table) {
        int n = 0, dev = 0;
        if (vacv_cuda_device_count(&n) != VACV_OK || n <= 0)
            throw std::runtime_error("vacv: no CUDA device available (libvacv has no CPU fallback)");
        if (vacv_cuda_get_device(&dev) != VACV_OK || dev < 0 || dev >= kMaxDevices)
            throw std::runtime_error(std::string("vacv: ") + vacv_cuda_last_error());
        DeviceContext& ctx = table[dev];
        if (!ctx.stream_) {
            ctx.device_ = dev;
            ctx.check(vacv_cuda_stream_create(&ctx.stream_));
        }
        return ctx;
    }
    void* stream() { return stream_; }
    int device() const { return device_; }
    // Failure: first wait for what is already queued on this stream -- async copies from / to the caller's (pooled, pinned)
    // tensor memory must not outlive the call that throws.
    void check(int status) {
        if (status == VACV_OK) return;
        const std::string msg = std::string("vacv: ") + vacv_cuda_last_error();
        if (stream_) vacv_cuda_stream_sync(stream_);
        throw std::runtime_error(msg);
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
    static constexpr int kMaxDevices = 16;
    DeviceContext() = default;
    DeviceContext(const DeviceContext&) = delete;
    void destroy() {   // with device_ current
        if (!stream_) return;
        vacv_cuda_stream_sync(stream_);
        for (Slot& s : slots_) if (s.p) vacv_cuda_free(s.p);
        vacv_cuda_stream_destroy(stream_);
        stream_ = nullptr;
    }
    friend struct DeviceContextTable;
    void* stream_ = nullptr;
    int device_ = -1;
    Slot slots_[kSlots];
};

struct DeviceContextTable {   // one per host thread; thread exit releases what the thread created
    DeviceContext ctx[DeviceContext::kMaxDevices];
    ~DeviceContextTable() {
        int prev = 0;
        if (vacv_cuda_get_device(&prev) != VACV_OK) return;
        for (DeviceContext& c : ctx)
            if (c.stream_ && vacv_cuda_set_device(c.device_) == VACV_OK) c.destroy();
        vacv_cuda_set_device(prev);
    }
};

inline DeviceContext& DeviceContext::current() {
    static thread_local DeviceContextTable table;
    return current_impl(table.ctx);
}

}  // namespace vacv_host
