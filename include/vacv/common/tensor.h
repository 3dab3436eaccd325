// vision::Tensor -- host-side container of the vacv drop-in library (libvacv.so).
//
// Binary- and source-compatible with the reference's src/common/tensor.h:27-84: same namespace, same public data
// members in the same order (callers and the reference's tests read w/h/c/stride/dims/data/dtype/layout directly,
// e.g. src/test/src/impl/test_resize.cpp:41-43), same constructors and methods, followed by the two private members
// (name string, intrusive ref-count pointer).  Build against the same libstdc++ C++11 ABI.
//
// Semantics kept from src/common/tensor.cpp: dense storage, stride = w*h elements per channel plane (:524);
// create() is a no-op when shape, dtype and layout already match (:512-516); buffers are ref-counted and freed by
// the last owner (:554-557); the (w,h,c,void*) constructors borrow memory (:82-85).
// change_layout / change_dtype (:393-502) run on the GPU through the C-ABI (include/vacv_cuda.h).
#ifndef VISION_TENSOR_H
#define VISION_TENSOR_H

#include <cstddef>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace vision {

enum DType { FP32 = 0, FP16 = 1, INT8 = 2, FP64 = 3, DTYPE_UNKNOWN };   // tensor.h:12-18
enum DLayout { NCHW = 0, NHWC = 1 };                                     // tensor.h:21-24

class Tensor {
public:
    // ---- data: public, read directly by callers; this order IS the binary interface (reference tensor.h:71-78)
    int w, h, c;        // width, height, channels
    int stride;         // elements per channel plane = w * h
    int dims;           // 1, 2 or 3
    void* data;         // host pointer, dense
    DType dtype;
    DLayout layout;

    // ---- lifetime
    Tensor();
    Tensor(const Tensor& other);
    Tensor& operator=(const Tensor& other);
    ~Tensor();

    // owning: allocate width [x height [x channels]] elements; the reference offers both (layout, dtype) and
    // (dtype, layout) argument orders and both are kept
    explicit Tensor(int width, DLayout arrangement = NCHW, DType element = FP32);
    explicit Tensor(int width, DType element = FP32, DLayout arrangement = NCHW);
    Tensor(int width, int height, DLayout arrangement = NCHW, DType element = FP32);
    Tensor(int width, int height, DType element = FP32, DLayout arrangement = NCHW);
    Tensor(int width, int height, int channels, DLayout arrangement = NCHW, DType element = FP32);
    Tensor(int width, int height, int channels, DType element = FP32, DLayout arrangement = NCHW);

    // borrowing: wrap caller memory, never freed here
    Tensor(int width, void* memory, DType element = FP32, DLayout arrangement = NCHW);
    Tensor(int width, void* memory, DLayout arrangement = NCHW, DType element = FP32);
    Tensor(int width, int height, void* memory, DType element = FP32, DLayout arrangement = NCHW);
    Tensor(int width, int height, void* memory, DLayout arrangement = NCHW, DType element = FP32);
    Tensor(int width, int height, int channels, void* memory, DType element = FP32, DLayout arrangement = NCHW);
    Tensor(int width, int height, int channels, void* memory, DLayout arrangement = NCHW, DType element = FP32);

    // (re)allocate; a no-op when shape, dtype and layout already match
    void create(int width, int height, int channels, DType element = FP32, DLayout arrangement = NCHW);
    void create(int width, int height, int channels, DLayout arrangement = NCHW, DType element = FP32);
    void create(int width, int height, DType element = FP32, DLayout arrangement = NCHW);
    void create(int width, int height, DLayout arrangement = NCHW, DType element = FP32);
    void create(int width, DType element = FP32, DLayout arrangement = NCHW);
    void create(int width, DLayout arrangement = NCHW, DType element = FP32);
    void release();

    // ---- queries
    bool empty() const;
    size_t size() const;                      // elements
    size_t len() const;                       // bytes
    int get_ref_count() const;
    std::string get_name() const;
    void set_name(const std::string& name);

    // ---- conversions (run on the GPU)
    Tensor clone() const;
    Tensor change_dtype(DType element);       // INT8 <-> FP32 (fp32 -> u8 truncates); same dtype -> clone
    Tensor change_layout(DLayout arrangement);   // HWC <-> CHW; c == 1 or same layout -> clone

private:
    void add_ref() const;
    std::string _name;
    int* _ref_count;
};

using TensorArray = std::vector<Tensor>;
using TensorPtr = std::shared_ptr<Tensor>;

}  // namespace vision

#endif  // VISION_TENSOR_H
