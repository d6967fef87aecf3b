// lanes.cuh -- two rays per thread: the value types the surface steps are written over.
//
// Why: on the B200 the FP64 pipe's throughput depends on how many INDEPENDENT FP64 instructions a warp can issue back
// to back, not on how many warps are resident (tools/ubench/fp64_issue.cu, profiles/r01_fp64_issue_ubench.txt: one
// dependent chain per warp tops out at 66 % of the pipe with 4 or with 8 warps per scheduler, two chains at 79 %,
// four at 92-95 %).  One ray through one surface is mostly serial (Newton chains of / and sqrt, left-associated sums),
// so the trace kernels carry TWO rays per thread and run every step on both in the same basic block: ptxas then
// interleaves the two independent instruction streams.
//
// The steps (surface_steps.cuh) are templates over the value type V: `double` (one ray; the out-of-line Careful
// fall-backs and the helper kernels) or `D2` (a pair).  Everything here is lane-wise; a lane never sees the other's data,
// so a pair computes exactly what two scalar evaluations compute.
#pragma once

#include <cuda_runtime.h>

namespace rtb {

struct D2 {
    double a, b;
};
struct B2 {
    bool a, b;
};

template <class V>
struct LaneTraits;
template <>
struct LaneTraits<double> {
    using Bool = bool;
    static constexpr int kLanes = 1;
};
template <>
struct LaneTraits<D2> {
    using Bool = B2;
    static constexpr int kLanes = 2;
};
template <class V>
using bool_of = typename LaneTraits<V>::Bool;

// ---- arithmetic (each operator is one rounded operation per lane; the translation unit is built with -fmad=false) ----
#define RTB_D2_BINOP(op)                                                                                              \
    __device__ __forceinline__ D2 operator op(D2 x, D2 y) { return D2{x.a op y.a, x.b op y.b}; }                      \
    __device__ __forceinline__ D2 operator op(D2 x, double y) { return D2{x.a op y, x.b op y}; }                      \
    __device__ __forceinline__ D2 operator op(double x, D2 y) { return D2{x op y.a, x op y.b}; }
RTB_D2_BINOP(+)
RTB_D2_BINOP(-)
RTB_D2_BINOP(*)
#undef RTB_D2_BINOP
__device__ __forceinline__ D2 operator-(D2 x) { return D2{-x.a, -x.b}; }

#define RTB_D2_CMP(op)                                                                                                \
    __device__ __forceinline__ B2 operator op(D2 x, D2 y) { return B2{x.a op y.a, x.b op y.b}; }                      \
    __device__ __forceinline__ B2 operator op(D2 x, double y) { return B2{x.a op y, x.b op y}; }                      \
    __device__ __forceinline__ B2 operator op(double x, D2 y) { return B2{x op y.a, x op y.b}; }
RTB_D2_CMP(<)
RTB_D2_CMP(<=)
RTB_D2_CMP(>)
RTB_D2_CMP(>=)
RTB_D2_CMP(==)
RTB_D2_CMP(!=)
#undef RTB_D2_CMP

__device__ __forceinline__ B2 operator&(B2 x, B2 y) { return B2{x.a && y.a, x.b && y.b}; }
__device__ __forceinline__ B2 operator|(B2 x, B2 y) { return B2{x.a || y.a, x.b || y.b}; }
__device__ __forceinline__ B2 operator&(B2 x, bool y) { return B2{x.a && y, x.b && y}; }
__device__ __forceinline__ B2 operator|(B2 x, bool y) { return B2{x.a || y, x.b || y}; }
__device__ __forceinline__ B2 operator!(B2 x) { return B2{!x.a, !x.b}; }

// ---- selects and reductions over the lanes --------------------------------------------------------------------------
__device__ __forceinline__ double vsel(bool c, double x, double y) { return c ? x : y; }
__device__ __forceinline__ D2 vsel(B2 c, D2 x, D2 y) { return D2{c.a ? x.a : y.a, c.b ? x.b : y.b}; }
__device__ __forceinline__ D2 vsel(B2 c, D2 x, double y) { return D2{c.a ? x.a : y, c.b ? x.b : y}; }
__device__ __forceinline__ D2 vsel(B2 c, double x, D2 y) { return D2{c.a ? x : y.a, c.b ? x : y.b}; }
__device__ __forceinline__ D2 vsel(B2 c, double x, double y) { return D2{c.a ? x : y, c.b ? x : y}; }
__device__ __forceinline__ bool vall(bool c) { return c; }
__device__ __forceinline__ bool vall(B2 c) { return c.a && c.b; }
__device__ __forceinline__ bool vany(bool c) { return c; }
__device__ __forceinline__ bool vany(B2 c) { return c.a || c.b; }
__device__ __forceinline__ double vabs(double x) { return fabs(x); }
__device__ __forceinline__ D2 vabs(D2 x) { return D2{fabs(x.a), fabs(x.b)}; }

// ---- lane access -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lane_of(double v, int) { return v; }
__device__ __forceinline__ double lane_of(D2 v, int j) { return j == 0 ? v.a : v.b; }
__device__ __forceinline__ bool lane_of(bool v, int) { return v; }
__device__ __forceinline__ bool lane_of(B2 v, int j) { return j == 0 ? v.a : v.b; }
__device__ __forceinline__ void set_lane(double &v, int, double x) { v = x; }
__device__ __forceinline__ void set_lane(D2 &v, int j, double x)
{
    if (j == 0) v.a = x;
    else v.b = x;
}
__device__ __forceinline__ void set_lane(bool &v, int, bool x) { v = x; }
__device__ __forceinline__ void set_lane(B2 &v, int j, bool x)
{
    if (j == 0) v.a = x;
    else v.b = x;
}

template <class V>
__device__ __forceinline__ V splat(double x);
template <>
__device__ __forceinline__ double splat<double>(double x) { return x; }
template <>
__device__ __forceinline__ D2 splat<D2>(double x) { return D2{x, x}; }

} // namespace rtb
