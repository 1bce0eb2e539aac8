// Shared definitions for the libroar_sup kernels.  Everything marked HD is plain C++ that also
// compiles for the host, so the CPU harness under tests/hostemu can run the *same* per-thread
// code the kernels run (thread loops replace the grid; test infrastructure only).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace roar {

struct alignas(8) cf32 { float x, y; };
struct alignas(16) cf64 { double x, y; };

template <class C> struct real_of;
template <> struct real_of<cf32> { typedef float type; };
template <> struct real_of<cf64> { typedef double type; };

template <class C> HD C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <class C> HD C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <class C> HD C cmul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <class C> HD C cconj(C a) { C r; r.x = a.x; r.y = -a.y; return r; }
// multiply by -i (forward) or +i (inverse)
template <bool INV, class C> HD C cmul_mi(C a) { C r; if (INV) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; } return r; }

}  // namespace roar
