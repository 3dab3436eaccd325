// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- never linked into the product.
//
// Translation unit that compiles the reference's own src/common/tensor.cpp *where it lies*
// under /root/reference (it is #included, not copied).
//
// Why a wrapper: the reference HEAD does not compile.  tensor.cpp:536 calls
//     data = VaAllocator::allocate(data_len + (int)sizeof(*_ref_count));
// but va_allocator.h:8 declares  static void allocate(void** data, int len);
// (SURVEY.md section 0 / App. B shim 1).  Instead of patching a copy of the source we rename the
// class for the duration of that one file to an adapter with the call shape tensor.cpp expects,
// forwarding to the reference's real allocator (va_allocator.cpp, compiled unmodified).
#include "common/va_allocator.h"   // include guard is now set; tensor.cpp's own include is a no-op

namespace vision {
struct VaAllocatorCallAdapter {
    static void* allocate(int len) {
        void* p = nullptr;
        VaAllocator::allocate(&p, len);
        return p;
    }
    static void deallocate(void* p) { VaAllocator::deallocate(p); }
    static int align_size(int len) { return VaAllocator::align_size(len); }
    static int align_size(int len, int n) { return VaAllocator::align_size(len, n); }
};
}  // namespace vision

#define VaAllocator VaAllocatorCallAdapter
#include "common/tensor.cpp"
#undef VaAllocator
