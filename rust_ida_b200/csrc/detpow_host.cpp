// detpow_host.cpp -- host build of detpow.h, used ONLY by tests to check det_pow against the
// platform libm pow (the reference's f64::powf) and against the device build.  Not a CPU fallback:
// nothing in the engine calls it.
#include <math.h>
#include <stdint.h>

#include "detpow.h"

extern "C" {

double detpow_host(double x, double y) { return detpow::det_pow(x, y); }

// splitmix64 -> uniform double in [0,1)
static inline uint64_t sm64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// Compare det_pow with libm pow on n samples: x log-uniform in [xlo, xhi], y cycling through ys[].
// Returns the number of bit mismatches; the first few mismatching (x, y) are stored in bad[2*k].
int64_t detpow_compare_libm(int64_t n, uint64_t seed, double xlo, double xhi, const double* ys, int ny,
                            double* bad, int maxbad, int64_t* n_fallback) {
    uint64_t s = seed;
    int64_t mism = 0, fb = 0;
    const double l0 = log(xlo), l1 = log(xhi);
    for (int64_t i = 0; i < n; ++i) {
        const double u = (double)(sm64(s) >> 11) * 0x1.0p-53;
        const double x = exp(l0 + (l1 - l0) * u);
        const double y = ys[i % ny];
        double r;
        const bool main_path = detpow::pow_main(x, y, &r);
        if (!main_path) { fb++; continue; }
        const double ref = pow(x, y);
        if (detpow::d2u(r) != detpow::d2u(ref)) {
            if (mism < maxbad) { bad[2 * mism] = x; bad[2 * mism + 1] = y; }
            mism++;
        }
    }
    if (n_fallback) *n_fallback = fb;
    return mism;
}

void detpow_host_batch(const double* x, const double* y, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = detpow::det_pow(x[i], y[i]);
}
}
