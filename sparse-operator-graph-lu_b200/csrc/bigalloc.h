// Allocator for the large host arrays of the front-end (op lists, task graphs: 10^8 entries).
// Big requests are anonymous mappings advised to use 2 MiB pages (page faults dominate the cost of
// first touching several GB); elements are default-initialised, i.e. vector(n) / resize(n) of a trivial
// type does NOT zero -- pass an explicit value where zeros are needed.  Small requests use malloc.
#pragma once
#include <sys/mman.h>

#include <cstddef>
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>

namespace soglu {

constexpr size_t BIG_ALLOC_MIN = size_t(8) << 20;

inline void* big_alloc(size_t bytes) {
    if (bytes < BIG_ALLOC_MIN) {
        void* p = std::malloc(bytes ? bytes : 1);
        if (!p) throw std::bad_alloc();
        return p;
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
    madvise(p, bytes, MADV_HUGEPAGE);
#endif
    return p;
}
inline void big_free(void* p, size_t bytes) {
    if (!p) return;
    if (bytes < BIG_ALLOC_MIN) std::free(p);
    else munmap(p, bytes);
}

template <class T>
struct BigAlloc {
    using value_type = T;
    BigAlloc() = default;
    template <class U> BigAlloc(const BigAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(big_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t n) { big_free(p, n * sizeof(T)); }
    template <class U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
    template <class U> bool operator==(const BigAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const BigAlloc<U>&) const { return false; }
};
template <class T>
using BigVec = std::vector<T, BigAlloc<T>>;

}  // namespace soglu
