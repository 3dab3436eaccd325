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
    // -- owning constructors (allocate); two argument orders exist in the reference and both are kept
    Tensor();
    explicit Tensor(int w, DLayout layout = NCHW, DType dtype = FP32);
    Tensor(int w, int h, DLayout layout = NCHW, DType dtype = FP32);
    Tensor(int w, int h, int c, DLayout layout = NCHW, DType type = FP32);
    explicit Tensor(int w, DType dtype = FP32, DLayout layout = NCHW);
    Tensor(int w, int h, DType dtype = FP32, DLayout layout = NCHW);
    Tensor(int w, int h, int c, DType type = FP32, DLayout layout = NCHW);
    // -- borrowing constructors (wrap caller memory, never freed here)
    Tensor(int w, void* data, DType dtype = FP32, DLayout layout = NCHW);
    Tensor(int w, int h, void* data, DType dtype = FP32, DLayout layout = NCHW);
    Tensor(int w, int h, int c, void* data, DType type = FP32, DLayout layout = NCHW);
    Tensor(int w, void* data, DLayout layout = NCHW, DType dtype = FP32);
    Tensor(int w, int h, void* data, DLayout layout = NCHW, DType dtype = FP32);
    Tensor(int w, int h, int c, void* data, DLayout layout = NCHW, DType type = FP32);

    Tensor(const Tensor& t);
    ~Tensor();
    Tensor& operator=(const Tensor& t);

    Tensor clone() const;
    Tensor change_layout(DLayout layout);   // HWC <-> CHW; c == 1 or same layout -> clone
    Tensor change_dtype(DType dtype);       // INT8 <-> FP32 (fp32 -> u8 truncates); same dtype -> clone

    void create(int w, DType dtype = FP32, DLayout layout = NCHW);
    void create(int w, int h, DType dtype = FP32, DLayout layout = NCHW);
    void create(int w, int h, int c, DType dtype = FP32, DLayout layout = NCHW);
    void create(int w, DLayout layout = NCHW, DType dtype = FP32);
    void create(int w, int h, DLayout layout = NCHW, DType dtype = FP32);
    void create(int w, int h, int c, DLayout layout = NCHW, DType dtype = FP32);
    void release();

    bool empty() const;
    size_t size() const;   // elements
    size_t len() const;    // bytes
    void set_name(const std::string& name);
    std::string get_name() const;
    int get_ref_count() const;

    // public data, order fixed by the ABI
    int w;
    int h;
    int c;
    int stride;
    int dims;
    void* data;
    DType dtype;
    DLayout layout;

private:
    void add_ref() const;
    std::string _name;
    int* _ref_count;
};

using TensorArray = std::vector<Tensor>;
using TensorPtr = std::shared_ptr<Tensor>;

}  // namespace vision

#endif  // VISION_TENSOR_H
