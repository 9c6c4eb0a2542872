// Double-double arithmetic (an unevaluated sum hi + lo of two fp64 numbers, ~106 bits) for the Newton systems of the
// interior-point solver (ipm.cu).  The normal matrix G' W^-2 G of a late interior-point iteration mixes weights 1/slack^2 that
// span > 1e16, so the curvature contributed by the inactive rows is below the rounding unit of the active rows' and is lost
// when the sum is FORMED in fp64; measured on the N = 256 fir_ap_cvx problem at obj = 1e4 (tools/ipm_proto.py): fp64 ends
// 3e-4 from the optimum with a dual residual that stops decreasing, 64-bit-mantissa arithmetic ends 2e-6 from it.
//
// Only error-free transformations built from IEEE adds and FMAs: every operation uses the _rn intrinsics so that nvcc neither
// contracts nor reassociates them.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define DD_FN __host__ __device__ __forceinline__
#else
#define DD_FN inline
#endif
#include <cmath>

namespace mbrf {

// rounded-to-nearest primitives that the compiler must not contract: device intrinsics, plain IEEE operations on the host
// (the host build is only used by the CPU unit test of this header, compiled with -ffp-contract=off)
#ifdef __CUDA_ARCH__
#define DD_ADD(a, b) __dadd_rn(a, b)
#define DD_SUB(a, b) __dsub_rn(a, b)
#define DD_MUL(a, b) __dmul_rn(a, b)
#define DD_FMA(a, b, c) __fma_rn(a, b, c)
#define DD_DIV(a, b) __ddiv_rn(a, b)
#define DD_SQRT(a) __dsqrt_rn(a)
#else
#define DD_ADD(a, b) ((a) + (b))
#define DD_SUB(a, b) ((a) - (b))
#define DD_MUL(a, b) ((a) * (b))
#define DD_FMA(a, b, c) std::fma(a, b, c)
#define DD_DIV(a, b) ((a) / (b))
#define DD_SQRT(a) std::sqrt(a)
#endif

struct dd {
    double hi, lo;
};

DD_FN dd dd_make(double a) { return dd{a, 0.0}; }

DD_FN dd two_sum(double a, double b)
{
    const double s = DD_ADD(a, b);
    const double bb = DD_SUB(s, a);
    const double e = DD_ADD(DD_SUB(a, DD_SUB(s, bb)), DD_SUB(b, bb));
    return dd{s, e};
}
DD_FN dd quick_two_sum(double a, double b)   // |a| >= |b|
{
    const double s = DD_ADD(a, b);
    return dd{s, DD_SUB(b, DD_SUB(s, a))};
}
DD_FN dd two_prod(double a, double b)
{
    const double p = DD_MUL(a, b);
    return dd{p, DD_FMA(a, b, -p)};
}
DD_FN dd dd_add(dd a, dd b)     // accurate (IEEE-style) sum
{
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = DD_ADD(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = DD_ADD(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
DD_FN dd dd_add_d(dd a, double b)
{
    dd s = two_sum(a.hi, b);
    s.lo = DD_ADD(s.lo, a.lo);
    return quick_two_sum(s.hi, s.lo);
}
DD_FN dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
DD_FN dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
DD_FN dd dd_mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo = DD_FMA(a.hi, b.lo, p.lo);
    p.lo = DD_FMA(a.lo, b.hi, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
DD_FN dd dd_mul_d(dd a, double b)
{
    dd p = two_prod(a.hi, b);
    p.lo = DD_FMA(a.lo, b, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
// c - a*b, the inner operation of the factorisation: one renormalisation less than dd_sub(c, dd_mul(a, b))
DD_FN dd dd_fnma(dd a, dd b, dd c)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo = DD_FMA(a.hi, b.lo, p.lo);
    p.lo = DD_FMA(a.lo, b.hi, p.lo);
    dd s = two_sum(c.hi, -p.hi);
    s.lo = DD_ADD(s.lo, DD_SUB(c.lo, p.lo));
    return quick_two_sum(s.hi, s.lo);
}
// the same inside a long accumulation: the sum is kept UN-normalised (hi exact partial sums, lo the collected errors) and
// renormalised once at the end (dd_renorm) -- 12 instead of 15 FP64 instructions per term
DD_FN dd dd_fnma_acc(dd a, dd b, dd c)
{
    const double p = DD_MUL(a.hi, b.hi);
    double e = DD_FMA(a.hi, b.hi, -p);
    e = DD_FMA(a.hi, b.lo, e);
    e = DD_FMA(a.lo, b.hi, e);
    const dd s = two_sum(c.hi, -p);
    return dd{s.hi, DD_ADD(c.lo, DD_SUB(s.lo, e))};
}
DD_FN dd dd_renorm(dd a) { return two_sum(a.hi, a.lo); }
DD_FN dd dd_div(dd a, dd b)
{
    const double q1 = DD_DIV(a.hi, b.hi);
    dd r = dd_sub(a, dd_mul_d(b, q1));
    const double q2 = DD_DIV(r.hi, b.hi);
    r = dd_sub(r, dd_mul_d(b, q2));
    const double q3 = DD_DIV(r.hi, b.hi);
    dd q = quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
DD_FN dd dd_sqrt(dd a)          // a > 0 (Karp's trick: one Newton step on the fp64 root)
{
    if (!(a.hi > 0.0)) return dd{a.hi == 0.0 ? 0.0 : NAN, 0.0};
    const double x = DD_DIV(1.0, DD_SQRT(a.hi));
    const double ax = DD_MUL(a.hi, x);
    const dd r = dd_sub(a, two_prod(ax, ax));
    return two_sum(ax, DD_MUL(r.hi, DD_MUL(x, 0.5)));
}
DD_FN double dd_to_double(dd a) { return DD_ADD(a.hi, a.lo); }

// ---- scalar-type traits so that the factorisation / moment kernels are written once for double and dd ----
template <typename T> struct Num;
template <> struct Num<double> {
    static DD_FN double from(double a) { return a; }
    static DD_FN double from_dd(dd a) { return a.hi; }
    static DD_FN double to_double(double a) { return a; }
    static DD_FN double add(double a, double b) { return a + b; }
    static DD_FN double sub(double a, double b) { return a - b; }
    static DD_FN double mul(double a, double b) { return a * b; }
    static DD_FN double mul_d(double a, double b) { return a * b; }
    static DD_FN double fnma(double a, double b, double c) { return DD_FMA(-a, b, c); }
    static DD_FN double fnma_acc(double a, double b, double c) { return DD_FMA(-a, b, c); }
    static DD_FN double renorm(double a) { return a; }
    static DD_FN double div(double a, double b) { return a / b; }
    static DD_FN double sqrt_(double a) { return sqrt(a); }
    static DD_FN double neg(double a) { return -a; }
    static DD_FN bool positive(double a) { return a > 0.0; }
    static DD_FN double zero() { return 0.0; }
};
template <> struct Num<dd> {
    static DD_FN dd from(double a) { return dd{a, 0.0}; }
    static DD_FN dd from_dd(dd a) { return a; }
    static DD_FN double to_double(dd a) { return dd_to_double(a); }
    static DD_FN dd add(dd a, dd b) { return dd_add(a, b); }
    static DD_FN dd sub(dd a, dd b) { return dd_sub(a, b); }
    static DD_FN dd mul(dd a, dd b) { return dd_mul(a, b); }
    static DD_FN dd mul_d(dd a, double b) { return dd_mul_d(a, b); }
    static DD_FN dd fnma(dd a, dd b, dd c) { return dd_fnma(a, b, c); }
    static DD_FN dd fnma_acc(dd a, dd b, dd c) { return dd_fnma_acc(a, b, c); }
    static DD_FN dd renorm(dd a) { return dd_renorm(a); }
    static DD_FN dd div(dd a, dd b) { return dd_div(a, b); }
    static DD_FN dd sqrt_(dd a) { return dd_sqrt(a); }
    static DD_FN dd neg(dd a) { return dd_neg(a); }
    static DD_FN bool positive(dd a) { return a.hi > 0.0; }
    static DD_FN dd zero() { return dd{0.0, 0.0}; }
};

// sin and cos of a double-double argument |x| <= 4: scale down by 2^6, Taylor series, six double-angle steps on
// (sin, versine) -- the versine form loses no digits near cos = 1.
DD_FN void dd_sincos(dd x, dd *s_out, dd *c_out)
{
    const dd r = dd_mul_d(x, 1.0 / 64.0);
    const dd r2 = dd_mul(r, r);
    // sin r = r (1 - r2/6 (1 - r2/20 (1 - ...)));   vers r = 1 - cos r = r2/2 (1 - r2/12 (1 - r2/30 (...)))
    dd ps = dd_make(1.0), pv = dd_make(1.0);
#pragma unroll 1
    for (int k = 12; k >= 1; --k) {
        ps = dd_sub(dd_make(1.0), dd_mul(dd_mul_d(r2, 1.0), dd_div(ps, dd_make((double)((2 * k) * (2 * k + 1))))));
        pv = dd_sub(dd_make(1.0), dd_mul(r2, dd_div(pv, dd_make((double)((2 * k + 1) * (2 * k + 2))))));
    }
    dd s = dd_mul(r, ps);
    dd v = dd_mul(dd_mul_d(r2, 0.5), pv);
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
        const dd c = dd_sub(dd_make(1.0), v);
        const dd s2 = dd_mul_d(dd_mul(s, c), 2.0);      // sin 2a = 2 sin a cos a
        v = dd_mul_d(dd_mul(s, s), 2.0);                // vers 2a = 2 sin^2 a
        s = s2;
    }
    *s_out = s;
    *c_out = dd_sub(dd_make(1.0), v);
}

}  // namespace mbrf
