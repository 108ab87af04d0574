// What the per-lane code of the tolerance update (tol_update.cuh) needs from its environment, for
// the device (the product) and for the host (the CPU simulation under oracle/hostsim, test
// infrastructure only): the shared-memory array and 32-bit offsets into it, atomics, fast math.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "philox.cuh"

#define NXB_INF_F64_BITS 0x7ff0000000000000LL

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------ device
#include <cuda_runtime.h>
#define NXB_DEV __device__ __forceinline__
#define NXB_INF (__longlong_as_double(NXB_INF_F64_BITS))
#define TOL_ANY(cond) __any_sync(0xffffffffu, (cond))

namespace nxb {
// Everything a warp touches in shared memory is addressed as a 32-bit offset from the one dynamic
// shared array, so that the compiler keeps shared-space (LDS/STS) addressing with 32-bit
// arithmetic instead of rebuilding 64-bit generic pointers inside the hot loops.
extern __shared__ __align__(16) unsigned char nxb_smem[];

// single MUFU forms (flush-to-zero variants: no range fix-ups around the instruction).  Their
// callers keep the arguments in range: the logarithm sees (2^-24, 1), the exponential is clamped
// from below afterwards, the divisor is a normalised message sum in [0.5, 2).
NXB_DEV float nxb_fast_expf(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f)); return y; }
NXB_DEV float nxb_fast_log2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
NXB_DEV float nxb_fast_exp2f(float x) { return exp2f(x); }
NXB_DEV float nxb_fast_divf(float a, float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return a * r; }
NXB_DEV int nxb_float_as_int(float x) { return __float_as_int(x); }
NXB_DEV float nxb_int_as_float(int x) { return __int_as_float(x); }
NXB_DEV long long nxb_double_as_ll(double x) { return __double_as_longlong(x); }
NXB_DEV double nxb_ll_as_double(long long x) { return __longlong_as_double(x); }
NXB_DEV int nxb_clzll(unsigned long long x) { return __clzll((long long)x); }
NXB_DEV void nxb_atomic_or_shared(unsigned* a, unsigned v) { atomicOr(a, v); }
NXB_DEV void nxb_atomic_add_shared(unsigned* a, unsigned v) { atomicAdd(a, v); }
NXB_DEV void nxb_atomic_add_f64(double* a, double v) { atomicAdd(a, v); }
// One out-of-line copy: the sweep kernels are far larger than the instruction cache, so code size
// (not call overhead) is what matters.
__device__ __noinline__ NxbU4 nxb_philox_call(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                              unsigned k0, unsigned k1) {
    return nxb_philox4x32_10(c0, c1, c2, c3, k0, k1);
}
}  // namespace nxb

#else
// -------------------------------------------------------------------------------------- host
#define NXB_DEV inline
#define NXB_INF (INFINITY)
#define TOL_ANY(cond) (cond)
#ifndef __align__
#define __align__(n) __attribute__((aligned(n)))
#endif
struct uint4 { unsigned x, y, z, w; };

namespace nxb {
extern unsigned char nxb_smem[];

NXB_DEV float nxb_fast_expf(float x) { return expf(x); }
NXB_DEV float nxb_fast_log2f(float x) { return log2f(x); }
NXB_DEV float nxb_fast_exp2f(float x) { return exp2f(x); }
NXB_DEV float nxb_fast_divf(float a, float b) { return a / b; }
NXB_DEV int nxb_float_as_int(float x) { int i; memcpy(&i, &x, 4); return i; }
NXB_DEV float nxb_int_as_float(int i) { float x; memcpy(&x, &i, 4); return x; }
NXB_DEV long long nxb_double_as_ll(double x) { long long i; memcpy(&i, &x, 8); return i; }
NXB_DEV double nxb_ll_as_double(long long i) { double x; memcpy(&x, &i, 8); return x; }
NXB_DEV int nxb_clzll(unsigned long long x) { return x ? __builtin_clzll(x) : 64; }
NXB_DEV void nxb_atomic_or_shared(unsigned* a, unsigned v) { *a |= v; }
NXB_DEV void nxb_atomic_add_shared(unsigned* a, unsigned v) { *a += v; }
NXB_DEV void nxb_atomic_add_f64(double* a, double v) { *a += v; }
NXB_DEV NxbU4 nxb_philox_call(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
    return nxb_philox4x32_10(c0, c1, c2, c3, k0, k1);
}
}  // namespace nxb
#endif

namespace nxb {
template <typename T> struct SPtr {
    unsigned off;
    NXB_DEV T& operator[](int i) const { return reinterpret_cast<T*>(nxb_smem + off)[i]; }
    NXB_DEV T* ptr() const { return reinterpret_cast<T*>(nxb_smem + off); }
};
template <typename T> NXB_DEV T* sm_ptr(unsigned off) { return reinterpret_cast<T*>(nxb_smem + off); }
NXB_DEV float fmin_msg(float a, float b) { return a < b ? a : b; }
NXB_DEV double fmin_msg(double a, double b) { return a < b ? a : b; }
}  // namespace nxb
