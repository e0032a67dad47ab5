// Shared device helpers for the marlpde_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>

namespace mpde {

template <typename T> struct Cx { T re, im; };

template <typename T> __device__ __forceinline__ Cx<T> cx(T re, T im) { Cx<T> r; r.re = re; r.im = im; return r; }
template <typename T> __device__ __forceinline__ Cx<T> operator+(Cx<T> a, Cx<T> b) { return cx<T>(a.re + b.re, a.im + b.im); }
template <typename T> __device__ __forceinline__ Cx<T> operator-(Cx<T> a, Cx<T> b) { return cx<T>(a.re - b.re, a.im - b.im); }
// The library is compiled with -fmad=false: the compiler never fuses a multiply into an add on its own, so the same
// source expression rounds identically in every template instantiation (the team-size variants must agree bitwise).
// Fused multiply-adds are written out where they are wanted.
template <typename T> __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
    return cx<T>(fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re));
}
// a * conj(b)
template <typename T> __device__ __forceinline__ Cx<T> cmulc(Cx<T> a, Cx<T> b) {
    return cx<T>(fma(a.re, b.re, a.im * b.im), fma(a.im, b.re, -(a.re * b.im)));
}
template <typename T> __device__ __forceinline__ Cx<T> conj(Cx<T> a) { return cx<T>(a.re, -a.im); }

// vector type with the same layout as Cx<T> for 8/16-byte global accesses
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename T> __device__ __forceinline__ Cx<T> ldcx(const Cx<T>* p) {
    typename Vec2<T>::type v = *reinterpret_cast<const typename Vec2<T>::type*>(p);
    return cx<T>(v.x, v.y);
}
template <typename T> __device__ __forceinline__ void stcx(Cx<T>* p, Cx<T> a) {
    typename Vec2<T>::type v; v.x = a.re; v.y = a.im;
    *reinterpret_cast<typename Vec2<T>::type*>(p) = v;
}

__device__ __forceinline__ double shfl(double v, int src, unsigned mask) { return __shfl_sync(mask, v, src); }
__device__ __forceinline__ float shfl(float v, int src, unsigned mask) { return __shfl_sync(mask, v, src); }
__device__ __forceinline__ double shfl_xor(double v, int m, unsigned mask) { return __shfl_xor_sync(mask, v, m); }
__device__ __forceinline__ float shfl_xor(float v, int m, unsigned mask) { return __shfl_xor_sync(mask, v, m); }
template <typename T> __device__ __forceinline__ Cx<T> shfl_xor(Cx<T> v, int m, unsigned mask) {
    return cx<T>(shfl_xor(v.re, m, mask), shfl_xor(v.im, m, mask));
}
template <typename T> __device__ __forceinline__ Cx<T> shfl(Cx<T> v, int src, unsigned mask) {
    return cx<T>(shfl(v.re, src, mask), shfl(v.im, src, mask));
}

// The reference traps overflow when it casts v to complex64 for its history
// (Burger.py:8,498 / KS.py:7,273): an env is "blown up" as soon as a component of v
// is not representable in float32 (this also catches NaN/inf).
template <typename T> __device__ __forceinline__ bool blown(Cx<T> v) {
    return !(fabs((double)v.re) <= (double)FLT_MAX && fabs((double)v.im) <= (double)FLT_MAX);
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor in the stream is still running; it must not touch memory the predecessor
// writes before pdl_wait() returns (= predecessor complete and flushed).  Without the attribute both are no-ops.
// a / b through y = 1 / b (the correctly rounded reciprocal, computed earlier and off the critical path): q0 = a y,
// q = q0 + (a - q0 b) y is the correctly rounded quotient -- the bits of a / b -- as long as nothing under- or overflows
// (Markstein's correction step).  rcp_ok() is that safe range; callers fall back to the division itself outside it (also
// for non-finite operands, which fail every comparison).  float kernels keep the division: it is cheap.
__device__ __forceinline__ bool rcp_ok(double a, double b) {
    const double aa = fabs(a), bb = fabs(b);
    return (aa == 0.0 || (aa > 1e-250 && aa < 1e250)) && bb > 1e-250 && bb < 1e250;
}
__device__ __forceinline__ bool rcp_ok(float, float) { return false; }
__device__ __forceinline__ double div_rcp_fast(double a, double b, double y) {
    const double q0 = a * y;
    return fma(fma(-q0, b, a), y, q0);
}
__device__ __forceinline__ float div_rcp_fast(float a, float b, float) { return a / b; }
template <typename T> __device__ __forceinline__ T div_by_rcp(T a, T b, T y) { return rcp_ok(a, b) ? div_rcp_fast(a, b, y) : a / b; }

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Stores to an NVSwitch multicast address (multimem.st, sm_90+): the switch replicates the store into the
// buffer of every GPU bound to the multicast object.  Plain bit moves (no reduction).
__device__ __forceinline__ void st_multicast(double* p, double v) { asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st_multicast(float* p, float v) { asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void st_multicast(Cx<double>* p, Cx<double> v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__int_as_float(__double2loint(v.re))),
                 "f"(__int_as_float(__double2hiint(v.re))), "f"(__int_as_float(__double2loint(v.im))),
                 "f"(__int_as_float(__double2hiint(v.im)))
                 : "memory");
}
__device__ __forceinline__ void st_multicast(Cx<float>* p, Cx<float> v) {
    asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.re), "f"(v.im) : "memory");
}

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

}  // namespace mpde
