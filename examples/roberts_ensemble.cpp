// roberts_ensemble.cpp -- the reference's examples/roberts.rs (single Roberts instance, tout = 0.4*10^i) run as a
// small ensemble through the C++ mirror (include/b200ida.hpp) of the reference API.  Prints the reference's table
// for instance 0 and the final run statistics; exits non-zero if instance 0 does not reproduce 362 steps.
//   g++ -std=c++17 -Iinclude examples/roberts_ensemble.cpp -Lrust_ida_b200/csrc -lb200ida -o roberts_ensemble
#include <cstdio>
#include <vector>

#include "b200ida.hpp"

int main() {
    using namespace b200ida_host;
    const int64_t B = 256;
    const double rtol = 1e-4, atol[3] = {1e-8, 1e-6, 1e-6};
    std::vector<double> params(3 * B), yy0(3 * B, 0.0), yp0(3 * B, 0.0);
    for (int64_t b = 0; b < B; ++b) {
        const double s = 1.0 + 0.1 * ((double)b / (double)B - 0.5) * (b > 0);  // instance 0 = the reference's rates
        params[b] = 0.04 * s;
        params[B + b] = 1.0e4;
        params[2 * B + b] = 3.0e7;
        yy0[b] = 1.0;
        yp0[b] = -params[b];
        yp0[B + b] = params[b];
    }
    try {
        EnsembleIda ida(B200IDA_PROBLEM_ROBERTS, 3, B, 3, params.data(), yy0.data(), yp0.data(), rtol, atol);
        std::printf("%-12s %-14s %-14s %-14s %5s %2s %-12s\n", "t", "y1", "y2", "y3", "nst", "k", "h");
        double tout = 0.4;
        for (int iout = 0; iout < 12; ++iout, tout *= 10.0) {
            auto st = ida.solve(tout);
            auto yy = ida.get_yy();
            auto nst = ida.get_num_steps();
            std::printf("%-12.5e %-14.5e %-14.5e %-14.5e %5ld %2d %-12.5e status %d\n", tout, yy[0], yy[B], yy[2 * B], (long)nst[0],
                        ida.get_last_order()[0], ida.get_last_step()[0], st[0]);
        }
        // dense output on the last step: y' at the final output time from get_dky(k = 1) next to the solver's own yp
        std::vector<double> d1;
        const auto dst = ida.get_dky(4.0e10, 1, d1);
        const auto yp = ida.get_yp();
        std::printf("dky1 %.6e %.6e %.6e dky_status %d yp %.6e %.6e %.6e\n", d1[0], d1[B], d1[2 * B], dst[0], yp[0],
                    yp[B], yp[2 * B]);
        auto nst = ida.get_num_steps();
        std::printf("steps %ld  res evals %ld  jac evals %ld  nonlin iters %ld  err test fails %ld  conv fails %ld\n", (long)nst[0],
                    (long)ida.get_num_res_evals()[0], (long)ida.get_num_jac_evals()[0], (long)ida.get_num_nonlin_solv_iters()[0],
                    (long)ida.get_num_err_test_fails()[0], (long)ida.get_num_nonlin_solv_conv_fails()[0]);
        return nst[0] == 362 ? 0 : 1;
    } catch (const Error& e) {
        std::fprintf(stderr, "b200ida error %d: %s\n", e.code, e.what());
        return 2;
    }
}
