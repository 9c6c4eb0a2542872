// Rotation coefficients shared by the Bloch and forward-SLR kernels:
//     w  = cos(phi/2)            s2 = 2 sin(phi/2)/phi          from u = phi^2.
// Both are entire functions of u (rot_poly.h), so calcrotmat's / abrot's sqrt, divide, sin
// and cos (blochC.c:180,201,202; abrx.c:93-99) become two Horner chains.  The coefficients
// live in __constant__ memory so every DFMA takes its coefficient as a c[bank][offset]
// operand: no register, no per-iteration re-materialisation.
#pragma once
#include "rot_poly.h"

namespace mbrf {

enum { TIER_TINY = 0, TIER_SMALL = 1, TIER_MED = 2, TIER_BIG = 3, TIER_ANY = 4 };

__constant__ double kc_tiny_c[] = ROT_C_TINY;
__constant__ double kc_tiny_s[] = ROT_S_TINY;
__constant__ double kc_small_c[] = ROT_C_SMALL;
__constant__ double kc_small_s[] = ROT_S_SMALL;
__constant__ double kc_med_c[] = ROT_C_MED;
__constant__ double kc_med_s[] = ROT_S_MED;
__constant__ double kc_big_c[] = ROT_C_BIG;
__constant__ double kc_big_s[] = ROT_S_BIG;

template <int N>
__device__ __forceinline__ double horner_c(const double (&c)[N], double u)
{
    double acc = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; --i) acc = fma(acc, u, c[i]);
    return acc;
}

__device__ __forceinline__ int rot_tier(double u_bound)
{
    // NaN compares false everywhere and lands in TIER_ANY, whose sincos returns NaN like the reference
    return u_bound <= ROT_U_TINY ? TIER_TINY : u_bound <= ROT_U_SMALL ? TIER_SMALL : u_bound <= ROT_U_MED ? TIER_MED : u_bound <= ROT_U_BIG ? TIER_BIG : TIER_ANY;
}

template <int TIER>
__device__ __forceinline__ void rot_coeffs(double u, double &w, double &s2)
{
    if (TIER == TIER_TINY) {
        w = horner_c(kc_tiny_c, u);
        s2 = horner_c(kc_tiny_s, u);
    } else if (TIER == TIER_SMALL) {
        w = horner_c(kc_small_c, u);
        s2 = horner_c(kc_small_s, u);
    } else if (TIER == TIER_MED) {
        w = horner_c(kc_med_c, u);
        s2 = horner_c(kc_med_s, u);
    } else if (TIER == TIER_BIG || u <= ROT_U_BIG) {
        w = horner_c(kc_big_c, u);
        s2 = horner_c(kc_big_s, u);
    } else {  // more than a full turn in one sample (or NaN): the reference's own formula
        const double phi = sqrt(u);
        double s, c;
        sincos(0.5 * phi, &s, &c);
        w = c;
        s2 = 2.0 * s / phi;
    }
}

}  // namespace mbrf
