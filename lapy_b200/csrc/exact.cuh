// exact.cuh - IEEE round-to-nearest arithmetic that the compiler may not contract into FMAs.
//
// NumPy evaluates the reference's element formulas with separate multiplies and adds
// (np.sum(x*y,axis=1) == (x0*y0 + x1*y1) + x2*y2, np.cross unfused).  Using the
// *_rn intrinsics makes the CUDA kernels reproduce those local entries bit for bit independent
// of -fmad, including the exact zeros of right angles that the reference stores explicitly.
#pragma once
#include <cuda_runtime.h>

namespace lb {

// 32-byte aligned 4 x fp64: exactly one DRAM/L2 sector per gathered vertex or element record
struct __align__(32) D4 {
    double x, y, z, w;
};
__device__ __forceinline__ D4 ldg_d4(const D4 *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void st_d4(D4 *p, double x, double y, double z, double w) {
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(x, y);
    q[1] = make_double2(z, w);
}

template <class T>
struct Ex;

template <>
struct Ex<double> {
    using V4 = D4;
    static __device__ __forceinline__ V4 ldg(const V4 *p) { return ldg_d4(p); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; }
};

template <>
struct Ex<float> {
    using V4 = float4;
    static __device__ __forceinline__ V4 ldg(const V4 *p) { return __ldg(p); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    // python float eps compared against float32 arrays is cast to float32 (NEP 50)
    static __device__ __forceinline__ float eps() { return (float)2.220446049250313e-16; }
};

template <class T>
struct Vec3 {
    T x, y, z;
};

template <class T>
__device__ __forceinline__ Vec3<T> vsub(const Vec3<T> &a, const Vec3<T> &b) {
    return {Ex<T>::sub(a.x, b.x), Ex<T>::sub(a.y, b.y), Ex<T>::sub(a.z, b.z)};
}
template <class T>
__device__ __forceinline__ Vec3<T> vneg(const Vec3<T> &a) {
    return {-a.x, -a.y, -a.z};
}
// (x0*y0 + x1*y1) + x2*y2
template <class T>
__device__ __forceinline__ T vdot(const Vec3<T> &a, const Vec3<T> &b) {
    using E = Ex<T>;
    return E::add(E::add(E::mul(a.x, b.x), E::mul(a.y, b.y)), E::mul(a.z, b.z));
}
// np.cross: (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
template <class T>
__device__ __forceinline__ Vec3<T> vcross(const Vec3<T> &a, const Vec3<T> &b) {
    using E = Ex<T>;
    return {E::sub(E::mul(a.y, b.z), E::mul(a.z, b.y)), E::sub(E::mul(a.z, b.x), E::mul(a.x, b.z)),
            E::sub(E::mul(a.x, b.y), E::mul(a.y, b.x))};
}
template <class T>
__device__ __forceinline__ Vec3<double> vwiden(const Vec3<T> &a) {
    return {(double)a.x, (double)a.y, (double)a.z};
}

template <class T>
__device__ __forceinline__ Vec3<T> load_vertex(const typename Ex<T>::V4 *__restrict__ v4, int idx) {
    typename Ex<T>::V4 p = Ex<T>::ldg(v4 + idx);
    return {p.x, p.y, p.z};
}
}  // namespace lb
