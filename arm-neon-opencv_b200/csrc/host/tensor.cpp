// vision::Tensor for libvacv.so -- same observable behaviour as the reference's src/common/tensor.cpp (cited per
// function), written for this library: pooled page-locked host storage (host_pool.h; the reference's USE_CUDA allocator is
// cudaHostAlloc, va_allocator.cpp:12-31) with an intrusive reference count placed behind the (4-byte aligned) payload, and change_layout / change_dtype executed on the GPU through the C-ABI.
#include "common/tensor.h"

#include <atomic>
#include <cstdlib>
#include <stdexcept>

#include "device_context.h"
#include "host_pool.h"
#include "vacv_cuda.h"

namespace vision {

namespace {
inline size_t dtype_bytes(DType t) { return t == FP64 ? 8 : t == FP32 ? 4 : t == FP16 ? 2 : 1; }   // tensor.cpp:575-591
inline std::atomic<int>* counter(int* p) { return reinterpret_cast<std::atomic<int>*>(p); }
}  // namespace

Tensor::Tensor()
    : w(0), h(0), c(0), stride(0), dims(0), data(nullptr), dtype(FP32), layout(NCHW), _name(), _ref_count(nullptr) {}

// Owning constructors.  (The reference calls create() with w/h/c uninitialised, tensor.cpp:54-58 / SURVEY C-7;
// here the members are zeroed first so create() always allocates.)
Tensor::Tensor(int w_, DType dtype_, DLayout layout_) : Tensor() { create(w_, 1, 1, dtype_, layout_); }
Tensor::Tensor(int w_, int h_, DType dtype_, DLayout layout_) : Tensor() { create(w_, h_, 1, dtype_, layout_); }
Tensor::Tensor(int w_, int h_, int c_, DType dtype_, DLayout layout_) : Tensor() { create(w_, h_, c_, dtype_, layout_); }
Tensor::Tensor(int w_, DLayout layout_, DType dtype_) : Tensor(w_, dtype_, layout_) {}
Tensor::Tensor(int w_, int h_, DLayout layout_, DType dtype_) : Tensor(w_, h_, dtype_, layout_) {}
Tensor::Tensor(int w_, int h_, int c_, DLayout layout_, DType dtype_) : Tensor(w_, h_, c_, dtype_, layout_) {}

// Borrowing constructors (tensor.cpp:72-97): no reference count, never freed.
Tensor::Tensor(int w_, void* p, DType dtype_, DLayout layout_)
    : w(w_), h(1), c(1), stride(w_), dims(1), data(p), dtype(dtype_), layout(layout_), _name(), _ref_count(nullptr) {}
Tensor::Tensor(int w_, int h_, void* p, DType dtype_, DLayout layout_)
    : w(w_), h(h_), c(1), stride(w_ * h_), dims(2), data(p), dtype(dtype_), layout(layout_), _name(), _ref_count(nullptr) {}
Tensor::Tensor(int w_, int h_, int c_, void* p, DType dtype_, DLayout layout_)
    : w(w_), h(h_), c(c_), stride(w_ * h_), dims(3), data(p), dtype(dtype_), layout(layout_), _name(), _ref_count(nullptr) {}
Tensor::Tensor(int w_, void* p, DLayout layout_, DType dtype_) : Tensor(w_, p, dtype_, layout_) {}
Tensor::Tensor(int w_, int h_, void* p, DLayout layout_, DType dtype_) : Tensor(w_, h_, p, dtype_, layout_) {}
Tensor::Tensor(int w_, int h_, int c_, void* p, DLayout layout_, DType dtype_) : Tensor(w_, h_, c_, p, dtype_, layout_) {}

Tensor::Tensor(const Tensor& t)
    : w(t.w), h(t.h), c(t.c), stride(t.stride), dims(t.dims), data(t.data), dtype(t.dtype), layout(t.layout),
      _name(t._name), _ref_count(t._ref_count) {
    add_ref();
}

Tensor::~Tensor() { release(); }

Tensor& Tensor::operator=(const Tensor& t) {   // tensor.cpp:103-127: take the new reference before dropping ours
    if (this == &t) return *this;
    t.add_ref();
    release();
    w = t.w; h = t.h; c = t.c; stride = t.stride; dims = t.dims;
    data = t.data; dtype = t.dtype; layout = t.layout;
    _name = t._name;
    _ref_count = t._ref_count;
    return *this;
}

Tensor Tensor::clone() const {   // tensor.cpp:146-158
    Tensor t;
    if (empty()) return t;
    t.create(w, h, c, dtype, layout);
    if (size() > 0) std::memcpy(t.data, data, len());
    return t;
}

void Tensor::create(int w_, DType dtype_, DLayout layout_) { create(w_, 1, 1, dtype_, layout_); }
void Tensor::create(int w_, int h_, DType dtype_, DLayout layout_) { create(w_, h_, 1, dtype_, layout_); }
void Tensor::create(int w_, DLayout layout_, DType dtype_) { create(w_, 1, 1, dtype_, layout_); }
void Tensor::create(int w_, int h_, DLayout layout_, DType dtype_) { create(w_, h_, 1, dtype_, layout_); }
void Tensor::create(int w_, int h_, int c_, DLayout layout_, DType dtype_) { create(w_, h_, c_, dtype_, layout_); }

void Tensor::create(int w_, int h_, int c_, DType dtype_, DLayout layout_) {   // tensor.cpp:512-541
    if (data && w == w_ && h == h_ && c == c_ && dtype == dtype_ && layout == layout_) return;   // reuse
    release();
    dtype = dtype_; layout = layout_;
    w = w_; h = h_; c = c_;
    stride = w * h;
    dims = (h == 1 && c == 1) ? 1 : (c == 1 ? 2 : 3);
    const size_t bytes = len();
    if (bytes == 0) return;
    const size_t padded = (bytes + 3) & ~size_t(3);   // allocation = align4(len) + sizeof(int) (:534-539)
    data = vacv_host::HostPool::instance().allocate(padded + sizeof(int));
    _ref_count = reinterpret_cast<int*>(static_cast<unsigned char*>(data) + padded);
    *_ref_count = 1;
}

void Tensor::release() {   // tensor.cpp:552-568
    if (_ref_count && counter(_ref_count)->fetch_sub(1, std::memory_order_acq_rel) == 1) vacv_host::HostPool::instance().release(data);
    data = nullptr; _ref_count = nullptr;
    dtype = FP32; layout = NCHW;
    stride = 0; w = 0; h = 0; c = 0;
    _name.clear();
}

bool Tensor::empty() const { return data == nullptr || size() == 0; }
size_t Tensor::size() const { return static_cast<size_t>(stride) * c; }
size_t Tensor::len() const { return size() * dtype_bytes(dtype); }
void Tensor::set_name(const std::string& name) { _name = name; }
std::string Tensor::get_name() const { return _name; }
int Tensor::get_ref_count() const { return _ref_count ? *_ref_count : 0; }
void Tensor::add_ref() const { if (_ref_count) counter(_ref_count)->fetch_add(1, std::memory_order_acq_rel); }

// change_layout (tensor.cpp:393-457) on the GPU: H2D, vacv_cuda_layout_change, D2H.
Tensor Tensor::change_layout(DLayout to) {
    if (empty()) return Tensor();
    if (c == 1 || to == layout) return clone();
    if (dtype != FP32 && dtype != FP16 && dtype != INT8)
        throw std::runtime_error("vacv: change_layout supports FP32 / FP16 / INT8");
    Tensor t;
    t.create(w, h, c, dtype, to);
    vacv_host::DeviceContext& ctx = vacv_host::DeviceContext::current();
    void* d_in = ctx.upload(0, data, len());
    void* d_out = ctx.scratch(1, len());
    ctx.check(vacv_cuda_layout_change(d_in, d_out, 1, w, h, c, dtype, layout, to, ctx.stream()));
    ctx.download(t.data, d_out, len());
    return t;
}

// change_dtype (tensor.cpp:459-502) on the GPU.  Unsupported pairs raise instead of returning uninitialised memory.
Tensor Tensor::change_dtype(DType to) {
    if (empty()) return Tensor();
    if (to == dtype) return clone();
    if (!((dtype == INT8 && to == FP32) || (dtype == FP32 && to == INT8)))
        throw std::runtime_error("vacv: change_dtype supports INT8 <-> FP32 only (reference: tensor.cpp:474-499)");
    Tensor t;
    t.create(w, h, c, to, layout);
    vacv_host::DeviceContext& ctx = vacv_host::DeviceContext::current();
    void* d_in = ctx.upload(0, data, len());
    void* d_out = ctx.scratch(1, t.len());
    ctx.check(vacv_cuda_dtype_change(d_in, d_out, size(), dtype, to, ctx.stream()));
    ctx.download(t.data, d_out, t.len());
    return t;
}

// VRect (vision_structs.cpp:72-90)
}  // namespace vision

#include "common/vision_structs.h"
namespace vision {
void VRect::set(float l, float t, float r, float b) { left = l; top = t; right = r; bottom = b; }
float VRect::width() const { return right - left; }
float VRect::height() const { return bottom - top; }
bool VRect::contains(float x, float y) { return left < right && top < bottom && x >= left && x < right && y >= top && y < bottom; }
}  // namespace vision
