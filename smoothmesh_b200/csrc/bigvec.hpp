// bigvec.hpp -- the vector type of the mesh-sized host arrays.
//
// sm::Vec<T> is std::vector<T> with an allocator that (a) leaves trivially constructible elements uninitialised on
// resize(n) / Vec(n), so that the OpenMP loop which fills the array is also the one that first touches its pages
// (a serial zero-fill of a 768 MB table costs more than building it), and (b) places large blocks on 2 MB
// boundaries and asks for transparent huge pages, which cuts the page-fault count of the one-time set-up by 512.
// Elements that must start at a value are set explicitly (parFill, or assign(n, v) for small arrays).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>
#ifdef __linux__
#include <sys/mman.h>
#endif

namespace sm
{

template <class T> struct BigAlloc
{
    using value_type = T;
    BigAlloc() = default;
    template <class U> BigAlloc(const BigAlloc<U> &) {}
    T *allocate(size_t n)
    {
        const size_t bytes = n * sizeof(T), huge = (size_t)2 << 20;
        void *p = nullptr;
        if (bytes >= 2 * huge)
        {
            p = std::aligned_alloc(huge, (bytes + huge - 1) / huge * huge);
#ifdef __linux__
            if (p)
                madvise(p, bytes, MADV_HUGEPAGE);
#endif
        }
        else
            p = std::malloc(bytes ? bytes : 1);
        if (!p)
            throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_t) { std::free(p); }
    template <class U, class... A> void construct(U *p, A &&...a)
    {
        if constexpr (sizeof...(A) == 0)
            ::new ((void *)p) U; // default-initialisation: nothing for arithmetic types
        else
            ::new ((void *)p) U(std::forward<A>(a)...);
    }
    template <class U> bool operator==(const BigAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const BigAlloc<U> &) const { return false; }
};

template <class T> using Vec = std::vector<T, BigAlloc<T>>;

// v = n copies of value, written by all threads (first touch)
template <class T> inline void parFill(Vec<T> &v, size_t n, T value)
{
    v.resize(n);
    T *p = v.data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        p[i] = value;
}

// v = src[0..n), copied by all threads
template <class T> inline void parCopy(Vec<T> &v, const T *src, size_t n)
{
    v.resize(n);
    T *p = v.data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        p[i] = src[i];
}

} // namespace sm
