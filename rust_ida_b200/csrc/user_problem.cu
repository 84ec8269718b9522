// user_problem.cu -- compiles ONE user-supplied problem functor into the engine, without touching the engine sources.
//
//     make -C rust_ida_b200/csrc user PROBLEM_HDR=/path/to/my_problem.cuh [USER_LIB=libb200ida_user.so]
//
// PROBLEM_HDR defines `struct UserProblem` in namespace b200ida: the reference's Residual / Jacobian / Root traits
// (src/traits.rs:28-89) as a device functor -- the contract is spelled out at the top of ida_thread_launch.cuh, and
// examples/user_problem_lorenz_root.cuh is a worked example.  This translation unit instantiates the thread-per-instance
// kernels for it (production path, and the rootfinding instantiation when NR > 0) and registers them with the engine
// under B200IDA_PROBLEM_USER when the library is loaded; it is linked with the engine's prebuilt objects into a library
// that exports the same C ABI.
#ifndef B200IDA_USER_PROBLEM_HDR
#error "build with -DB200IDA_USER_PROBLEM_HDR=\"/path/to/problem.cuh\" (make user PROBLEM_HDR=...)"
#endif
#include "ida_thread_launch.cuh"
#include B200IDA_USER_PROBLEM_HDR

namespace b200ida {

static_assert(UserProblem::N >= 1 && UserProblem::N <= 8, "the thread-per-instance kernels keep n-vectors in registers: n <= 8");

static int user_launch(const ThreadLaunch& a) {
    return launch_thread_family<UserProblem, (UserProblem::NR > 0 ? CAP_ROOTS : 0)>(a);
}
static const UserProblemHooks g_hooks = {UserProblem::N, UserProblem::NP, UserProblem::NR, user_launch};
static const int g_registered = (register_user_problem(&g_hooks), 0);

}  // namespace b200ida
