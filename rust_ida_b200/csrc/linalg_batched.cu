// linalg_batched.cu -- C ABI "Level 1" (include/b200ida.h): batched dense LU (getrf), triangular solves
// (getrs) and WRMS norms for ensembles of small and medium systems.
//
// Reference semantics, reproduced bit-for-bit (crates/linear/src/dense.rs:86-206, src/norm_rms.rs:31-38):
//   getrf: pivot = first row with strictly largest |a_ik|; whole-row swap; l_ik = a_ik * (1/a_kk);
//          a_ij -= a_kj * l_ik, column j skipped when a_kj == 0; zero pivot -> info = k+1 and stop.
//   getrs: apply pivots to b; unit-lower forward sweep; backward sweep dividing by u_kk.
//   wrms : sqrt( fold_{i=0..n-1}(acc + (x_i w_i)^2) / n ).
// Every a_ij update of elimination step k is a single mul + sub that depends only on step k-1, so any
// parallelisation over (i, j) inside one step is bit-exact; only the k loop is sequential.  No FMA
// (-fmad=false).
//
// Kernels
//   *_thread_soa<N>   n <= 8, SoA: one thread per system, matrix in registers, perfectly coalesced
//                     [n][n][batch] loads/stores -> HBM-bound (16 n^2 + 8n B/system for getrf).
//   *_cta             any n <= 256, either layout: one CTA per system; the matrix is staged in shared
//                     memory when it fits (n <= 160), else factored in place in global memory / L2.
#include <cstdlib>
#include <mutex>
#include <vector>
#include "host_util.h"
#include "ida_common.cuh"
#include "lu_team.cuh"
#include "lu_blocked.cuh"

// Kernel-selection overrides used by tools/ while tuning (B200IDA_LUBG_THREADS, B200IDA_GETRS_NO_TMA): compiled in only
// with -DB200IDA_EXPERIMENTS, so that the shipped library's kernel choice never depends on the environment.
static inline const char* experiment_env(const char* name) {
#ifdef B200IDA_EXPERIMENTS
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

namespace b200ida {

// element (i, j) of system b
template <bool SOA>
__device__ __forceinline__ size_t mat_idx(int i, int j, int n, int64_t b, int64_t B) {
    return SOA ? ((size_t)(i * n + j) * B + b) : ((size_t)b * n * n + (size_t)j * n + i);
}
template <bool SOA>
__device__ __forceinline__ size_t vec_idx(int i, int n, int64_t b, int64_t B) {
    return SOA ? ((size_t)i * B + b) : ((size_t)b * n + i);
}

// ---- thread-per-system, registers, SoA --------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) getrf_thread_soa(double* __restrict__ A, int64_t* __restrict__ piv,
                                                        int32_t* __restrict__ info, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double a[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) a[i][j] = A[(size_t)(i * N + j) * B + b];
    int p[N];
#pragma unroll
    for (int k = 0; k < N; ++k) p[k] = 0;
    int inf = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (inf == 0) {
            int l = k;
            double best = fabs(a[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double v = fabs(a[i][k]);
                if (v > best) {
                    best = v;
                    l = i;
                }
            }
            p[k] = l;
            if (best == 0.0) {
                inf = k + 1;
            } else {
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    if (l == i) {
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            const double t = a[k][c];
                            a[k][c] = a[i][c];
                            a[i][c] = t;
                        }
                    }
                }
                const double mult = 1.0 / a[k][k];
#pragma unroll
                for (int i = k + 1; i < N; ++i) a[i][k] *= mult;
#pragma unroll
                for (int j = k + 1; j < N; ++j) {
                    const double a_kj = a[k][j];
                    if (a_kj != 0.0) {
#pragma unroll
                        for (int i = k + 1; i < N; ++i) a[i][j] -= a_kj * a[i][k];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) A[(size_t)(i * N + j) * B + b] = a[i][j];
#pragma unroll
    for (int k = 0; k < N; ++k) piv[(size_t)k * B + b] = p[k];
    info[b] = inf;
}

template <int N>
__global__ void __launch_bounds__(128) getrs_thread_soa(const double* __restrict__ A, const int64_t* __restrict__ piv,
                                                        double* __restrict__ bx, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double a[N][N], x[N];
    int p[N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) a[i][j] = A[(size_t)(i * N + j) * B + b];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        x[i] = bx[(size_t)i * B + b];
        p[i] = (int)piv[(size_t)i * B + b];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            if (p[k] == i) {
                const double t = x[k];
                x[k] = x[i];
                x[i] = t;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
        const double bk = x[k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) x[i] -= a[i][k] * bk;
    }
#pragma unroll
    for (int k = N - 1; k >= 1; --k) {
        x[k] /= a[k][k];
        const double bk = x[k];
#pragma unroll
        for (int i = 0; i < k; ++i) x[i] -= a[i][k] * bk;
    }
    x[0] /= a[0][0];
#pragma unroll
    for (int i = 0; i < N; ++i) bx[(size_t)i * B + b] = x[i];
}

// ---- warp-per-system, REGISTER resident, AoS, n == N in {16, 32} -------------------------------------------------
// Lane i holds row i of the matrix in registers for the whole factorisation; N = 16 packs TWO systems into a warp
// (lanes 0..15 and 16..31, every collective on the half-warp).  Rows are never moved: each lane tracks the LOGICAL
// position of its row (what the reference's whole-row swaps would have produced) and writes the row there at the
// end; the arithmetic on each element is unchanged.  Per elimination step: arg-max of |a_ik| by three integer
// redux.sync on the IEEE bit pattern (NaN below every candidate, ties to the lowest logical position, key carries the
// lane), exact reciprocal of the pivot in every lane, pivot row broadcast through shared memory, N-1-k independent
// mul/sub.  The throughput bound is instructions issued per system (6.5 k at N = 32, half of them the update), not bytes.
// one elimination step with compile-time k (template recursion forces the full unroll that keeps row[] in registers)
template <int N, int K>
__device__ __forceinline__ void lu_regs_step(double (&row)[N], double* prow, unsigned gmask, int gl, int& pos, bool& done, int& inf,
                                             int& mypiv) {
    if (inf == 0) {  // uniform over the group
        const double x = row[K];
        const unsigned pp = ((unsigned)pos << 8) | (unsigned)gl;
        const bool atk = pos == K;  // the lane that holds logical row K (not yet a pivot row: positions < K are final)
        // Fast path (the common case): the high words of |x| alone -- sign stripped, they order like the values for finite
        // x -- name a UNIQUE, finite maximum: one reduction and a ballot instead of three reductions.  A unique strictly
        // largest high word is a strictly largest |x|, so the choice is the reference's; equal high words (possible ties)
        // and any inf / NaN candidate (a NaN's high word is >= 0x7ff00000, so it is the maximum) take the full key below.
        const int key = done ? -1 : (__double2hiint(x) & 0x7fffffff);
        const int mkey = __reduce_max_sync(gmask, key);
        const unsigned wm = __ballot_sync(gmask, key == mkey) & gmask;
        unsigned mpp;
        if (__popc(wm) == 1 && mkey >= 0 && mkey < 0x7ff00000) {  // uniform over the group
            mpp = __shfl_sync(gmask, pp, __ffs(wm) - 1);
        } else {
            mpp = lu_argmax_full(gmask, x, done, pp, !done && atk);
        }
        const int l = (int)(mpp >> 8), lw = (int)(mpp & 0xffu);
        const double pv = __shfl_sync(gmask, x, lw, N);
        if (gl == K) mypiv = l;
        if (pv == 0.0) {
            inf = K + 1;
        } else {
            const double mult = __drcp_rn(pv);  // == 1.0 / pv (both correctly rounded)
            if (gl == lw) {
                pos = K;
                done = true;
            } else if (atk) {
                pos = l;
            }
            const double lik = x * mult;
            if (!done) row[K] = lik;
            // pivot row through shared memory: the owner lane stores row[K+1..N) as 16-byte pairs, everybody reads them
            // back as broadcasts -- one store + one load per PAIR instead of four 32-bit shuffles.  The a_kj == 0 skip
            // of the reference stays a per-element test: dropping it is only wrong when a_ij is -0.0 or the multiplier
            // is not finite, but then it IS wrong, and every cheaper formulation tried loses one of those corners.
            constexpr int J0 = (K + 1) & ~1;
            if (gl == lw) {
#pragma unroll
                for (int j = J0; j < N; j += 2) *reinterpret_cast<double2*>(prow + j) = make_double2(row[j], row[j + 1]);
            }
            __syncwarp(gmask);
#pragma unroll
            for (int j = J0; j < N; j += 2) {
                const double2 a = *reinterpret_cast<const double2*>(prow + j);
                if (j > K && !done && a.x != 0.0) row[j] -= a.x * lik;
                if (!done && a.y != 0.0) row[j + 1] -= a.y * lik;
            }
            __syncwarp(gmask);
        }
    }
    if constexpr (K + 1 < N) lu_regs_step<N, K + 1>(row, prow, gmask, gl, pos, done, inf, mypiv);
}

template <int N>
__global__ void __launch_bounds__(128) getrf_warp_regs(double* __restrict__ A, int64_t* __restrict__ piv,
                                                       int32_t* __restrict__ info, int64_t B) {
    constexpr int G = 32 / N;  // systems per warp
    __shared__ __align__(16) double prow_all[4 * 32];  // pivot-row staging: 32 doubles per warp (N per system)
    const int lane = threadIdx.x & 31, gl = lane & (N - 1), g = lane / N;
    double* prow = prow_all + (threadIdx.x >> 5) * 32 + g * N;
    const unsigned gmask = (N == 32) ? 0xffffffffu : (0xffffu << (16 * g));
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t nwork = (B + G - 1) / G;  // warp-sized work items
    for (int64_t wk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wk < nwork; wk += wpg) {
        const int64_t b = wk * G + g;
        const bool live = b < B;  // uniform over the group (a trailing odd system leaves the second half idle)
        double* Ab = A + (size_t)(live ? b : 0) * N * N;
        double row[N];
#pragma unroll
        for (int j = 0; j < N; ++j) row[j] = live ? Ab[j * N + gl] : (j == gl ? 1.0 : 0.0);
        int pos = gl;       // logical row index of the data this lane holds
        bool done = false;  // this lane's row has become a pivot row
        int inf = 0;
        int mypiv = 0;      // lane k keeps piv[k]
        lu_regs_step<N, 0>(row, prow, gmask, gl, pos, done, inf, mypiv);
        if (live) {
#pragma unroll
            for (int j = 0; j < N; ++j) Ab[j * N + pos] = row[j];
            piv[(size_t)b * N + gl] = (inf == 0 || gl < inf) ? mypiv : 0;
            if (gl == 0) info[b] = inf;
        }
    }
}

template <int N>
__global__ void __launch_bounds__(128) getrs_warp_regs(const double* __restrict__ A, const int64_t* __restrict__ piv,
                                                       double* __restrict__ bx, int64_t B) {
    // N = 16 packs TWO systems into a warp (lanes 0..15 and 16..31), like getrf_warp_regs: every shuffle has width N
    constexpr int G = 32 / N;
    const int lane = threadIdx.x & 31, gl = lane & (N - 1), g = lane / N;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t nwork = (B + G - 1) / G;
    for (int64_t wk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wk < nwork; wk += wpg) {
        const int64_t b = wk * G + g;
        const bool live = b < B;  // a trailing odd system leaves the second half idle (it computes on system 0, stores nothing)
        const double* Ab = A + (size_t)(live ? b : 0) * N * N;
        double row[N];
#pragma unroll
        for (int j = 0; j < N; ++j) row[j] = Ab[j * N + gl];
        double x = bx[(size_t)(live ? b : 0) * N + gl];
        const int pl = (int)piv[(size_t)(live ? b : 0) * N + gl];
#pragma unroll
        for (int k = 0; k < N; ++k) {  // apply the pivots in order: swap x_k <-> x_{p_k}
            const int pk = __shfl_sync(0xffffffffu, pl, k, N);
            const double xk = __shfl_sync(0xffffffffu, x, k, N), xq = __shfl_sync(0xffffffffu, x, pk, N);
            if (pk != k) {
                if (gl == k) x = xq;
                else if (gl == pk) x = xk;
            }
        }
#pragma unroll
        for (int k = 0; k < N - 1; ++k) {
            const double bk = __shfl_sync(0xffffffffu, x, k, N);
            if (gl > k) x -= row[k] * bk;
        }
        // backward sweep.  The division by u_kk is straight-line code: every lane forms the exact-reciprocal quotient of its
        // own (x, u_ll, 1 / u_ll) each step (exact_div.cuh; 1 / u_ll computed up front, off the chain), only the pivot
        // lane's is broadcast, a zero numerator takes q0 (the zero of the right sign).  The range tests run once, on the
        // numerator each lane had at its pivot step; a failure redoes the sweep with real divisions.
        const double dg = Ab[gl * N + gl];
        const double rd = 1.0 / dg;
        const double x_fwd = x;
        double xp = 1.0;
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            const double q0 = __dmul_rn(x, rd);
            const double rr = __fma_rn(-dg, q0, x);
            const double q1 = __fma_rn(rr, rd, q0);
            const double qq = (x == 0.0) ? q0 : q1;
            const bool pv = gl == k;
            xp = pv ? x : xp;
            const double bk = __shfl_sync(0xffffffffu, pv ? qq : x, k, N);
            const double t = x - row[k] * bk;
            x = pv ? bk : ((gl < k) ? t : x);
        }
        {
            const unsigned ex = ((unsigned)__double2hiint(xp) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(dg) >> 20) & 0x7ffu;
            const bool ok = (xp == 0.0) ? (ed - 1u < 0x7feu) : ((ex - 543u) <= 960u && (ed - 543u) <= 960u);
            if (!__all_sync(0xffffffffu, ok)) {  // rare; both halves redo (the real divisions give the same bits)
                x = x_fwd;
                for (int k = N - 1; k >= 1; --k) {
                    if (gl == k) x /= dg;
                    const double bk = __shfl_sync(0xffffffffu, x, k, N);
                    const double c = Ab[k * N + gl];
                    if (gl < k) x -= c * bk;
                }
                if (gl == 0) x /= dg;
            }
        }
        if (live) bx[(size_t)b * N + gl] = x;
    }
}

// ---- group-per-system kernels on lu_team.cuh: one warp (n <= 32) or one CTA (n <= 256) per matrix ----------
constexpr int CTA_THREADS = 256;
constexpr int WARPS_PER_CTA = 4;   // warp-per-system kernels
constexpr int SMEM_MAX_N = 160;    // CTA kernels stage the matrix in shared memory up to here (161*160*8 B)

__host__ __device__ inline int lu_ld(int n) { return n | 1; }  // odd leading dimension: conflict-free row swaps
__host__ __device__ inline size_t lu_group_smem_bytes(int n) {
    // matrix (ld x n doubles) + rhs/pivot staging (n doubles + n ints) + LuShared, 16-byte aligned
    size_t b = (size_t)lu_ld(n) * n * sizeof(double) + (size_t)n * sizeof(double) + ((size_t)n * sizeof(int) + 15) / 16 * 16 + sizeof(LuShared);
    return (b + 15) / 16 * 16;
}

// WARP = true : blockDim = 32*WARPS_PER_CTA, each warp factors its own matrices (grid-stride over the batch)
// WARP = false: the whole CTA factors one matrix at a time; SMEM = false works in place in global memory (AoS only)
template <bool SOA, bool WARP, bool SMEM>
__global__ void __launch_bounds__(CTA_THREADS) getrf_group(double* __restrict__ A, int64_t* __restrict__ piv,
                                                           int32_t* __restrict__ info, int n, int64_t B, unsigned group_bytes) {
    extern __shared__ __align__(16) unsigned char lu_smem[];
    using G = LuGroup<WARP>;
    const int gi = WARP ? (int)(threadIdx.x >> 5) : 0, ng = WARP ? (int)(blockDim.x >> 5) : 1;
    unsigned char* base = lu_smem + (size_t)gi * group_bytes;
    const int ld = SMEM ? lu_ld(n) : n;
    double* Ms = reinterpret_cast<double*>(base);
    int* pivs = reinterpret_cast<int*>(base + (SMEM ? (size_t)ld * n * sizeof(double) : 0) + (size_t)n * sizeof(double));
    LuShared* sh = reinterpret_cast<LuShared*>(reinterpret_cast<unsigned char*>(pivs) + ((size_t)n * sizeof(int) + 15) / 16 * 16);
    const int tid = G::tid(), nt = G::nthreads(), nw = G::nwarps(), w = G::warp(), lane = tid & 31;
    for (int64_t b = (int64_t)blockIdx.x * ng + gi; b < B; b += (int64_t)gridDim.x * ng) {
        double* M = Ms;
        if (SMEM) {
            for (int j = w; j < n; j += nw)
                for (int i = lane; i < n; i += 32) Ms[(size_t)j * ld + i] = A[mat_idx<SOA>(i, j, n, b, B)];
        } else {
            M = A + (size_t)b * n * n;
        }
        G::sync();
        lu_factor<WARP>(M, ld, n, sh, pivs);
        if (SMEM) {
            for (int j = w; j < n; j += nw)
                for (int i = lane; i < n; i += 32) A[mat_idx<SOA>(i, j, n, b, B)] = Ms[(size_t)j * ld + i];
        }
        const int inf = sh->info;
        for (int k = tid; k < n; k += nt) piv[SOA ? ((size_t)k * B + b) : ((size_t)b * n + k)] = (inf == 0 || k < inf) ? pivs[k] : 0;
        if (tid == 0) info[b] = inf;
        G::sync();
    }
}

template <bool SOA, bool WARP, bool SMEM>
__global__ void __launch_bounds__(CTA_THREADS) getrs_group(const double* __restrict__ A, const int64_t* __restrict__ piv,
                                                           double* __restrict__ bx, int n, int64_t B, unsigned group_bytes) {
    extern __shared__ __align__(16) unsigned char lu_smem[];
    using G = LuGroup<WARP>;
    const int gi = WARP ? (int)(threadIdx.x >> 5) : 0, ng = WARP ? (int)(blockDim.x >> 5) : 1;
    unsigned char* base = lu_smem + (size_t)gi * group_bytes;
    const int ld = SMEM ? lu_ld(n) : n;
    double* Ms = reinterpret_cast<double*>(base);
    double* xs = reinterpret_cast<double*>(base + (SMEM ? (size_t)ld * n * sizeof(double) : 0));
    int* pivs = reinterpret_cast<int*>(xs + n);
    const int tid = G::tid(), nt = G::nthreads(), nw = G::nwarps(), w = G::warp(), lane = tid & 31;
    for (int64_t b = (int64_t)blockIdx.x * ng + gi; b < B; b += (int64_t)gridDim.x * ng) {
        const double* M = Ms;
        if (SMEM) {
            for (int j = w; j < n; j += nw)
                for (int i = lane; i < n; i += 32) Ms[(size_t)j * ld + i] = A[mat_idx<SOA>(i, j, n, b, B)];
        } else {
            M = A + (size_t)b * n * n;
        }
        for (int i = tid; i < n; i += nt) {
            xs[i] = bx[vec_idx<SOA>(i, n, b, B)];
            pivs[i] = (int)piv[SOA ? ((size_t)i * B + b) : ((size_t)b * n + i)];
        }
        G::sync();
        lu_solve<WARP>(M, ld, n, pivs, xs);
        for (int i = tid; i < n; i += nt) bx[vec_idx<SOA>(i, n, b, B)] = xs[i];
        G::sync();
    }
}

template <bool SOA>
__global__ void __launch_bounds__(256) wrms_thread(const double* __restrict__ x, const double* __restrict__ w,
                                                   double* __restrict__ out, int n, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const double p = x[vec_idx<SOA>(i, n, b, B)] * w[vec_idx<SOA>(i, n, b, B)];
        acc = acc + p * p;
    }
    out[b] = sqrt(acc / (double)n);
}

// masked norm, src/norm_rms.rs:49-57: ((x_i w_i) m_i)^2 folded in index order, divided by n (not by the mask count)
template <bool SOA>
__global__ void __launch_bounds__(256) wrms_masked_thread(const double* __restrict__ x, const double* __restrict__ w,
                                                          const unsigned char* __restrict__ id, double* __restrict__ out, int n,
                                                          int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const double p = (x[vec_idx<SOA>(i, n, b, B)] * w[vec_idx<SOA>(i, n, b, B)]) * (id[i] ? 1.0 : 0.0);
        acc = acc + p * p;
    }
    out[b] = sqrt(acc / (double)n);
}

// SoA <-> AoS conversions used when a large-n SoA batch has to go through the in-place AoS kernel.
// Matrices: SoA element (i, j, b) at (i*n + j)*B + b  <->  AoS column-major b*n*n + j*n + i.
__global__ void mat_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int n, int64_t B) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * n * B) return;
    const int64_t r = e / B, b = e % B;
    const int i = (int)(r / n), j = (int)(r % n);
    aos[b * n * n + (int64_t)j * n + i] = soa[e];
}
__global__ void mat_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int n, int64_t B) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * n * B) return;
    const int64_t r = e / B, b = e % B;
    const int i = (int)(r / n), j = (int)(r % n);
    soa[e] = aos[b * n * n + (int64_t)j * n + i];
}
__global__ void vec_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int64_t rows, int64_t B) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * B) return;
    const int64_t r = e / B, b = e % B;
    aos[b * rows + r] = soa[e];
}
// vectors (any 8-byte element): SoA (i, b) at i*B + b  <->  AoS b*n + i
__global__ void vec_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int64_t rows, int64_t B) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * B) return;
    const int64_t r = e / B, b = e % B;
    soa[e] = aos[b * rows + r];
}

template <int N>
static void launch_getrf_thread(double* A, int64_t* piv, int32_t* info, int64_t B, cudaStream_t st) {
    getrf_thread_soa<N><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(A, piv, info, B);
}
template <int N>
static void launch_getrs_thread(const double* A, const int64_t* piv, double* bx, int64_t B, cudaStream_t st) {
    getrs_thread_soa<N><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(A, piv, bx, B);
}

static int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// launch the group-per-system getrf (FACTOR) or getrs kernel for any n <= 256
// The attribute / occupancy queries cost several microseconds of host time per call -- comparable to the short
// kernels they configure -- so their result is cached per (kernel, device, threads, shared-memory size).
struct LaunchCfgEntry {
    const void* fn;
    int dev, threads;
    size_t smem;
    int occ;
};
template <class K>
static int launch_cfg(K kernel, int threads, size_t smem, int64_t groups_needed, int groups_per_cta, unsigned* grid) {
    static std::mutex mu;
    static std::vector<LaunchCfgEntry> cache;
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    const void* fn = reinterpret_cast<const void*>(kernel);
    int occ = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const LaunchCfgEntry& e : cache)
            if (e.fn == fn && e.dev == dev && e.threads == threads && e.smem == smem) occ = e.occ;
    }
    if (occ == 0) {
        // the dynamic shared-memory limit of a kernel is only ever raised (a smaller request after a larger one must
        // not lower it under a cached configuration)
        size_t cur = 0;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (const LaunchCfgEntry& e : cache)
                if (e.fn == fn && e.dev == dev && e.smem > cur) cur = e.smem;
        }
        if (smem > cur) B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
        if (occ < 1) occ = 1;
        std::lock_guard<std::mutex> lk(mu);
        cache.push_back({fn, dev, threads, smem, occ});
    }
    const int64_t want = (groups_needed + groups_per_cta - 1) / groups_per_cta;
    const int64_t cap = (int64_t)sm_count() * occ;
    *grid = (unsigned)(want < cap ? want : cap);
    return B200IDA_OK;
}


// ---- mbarrier / bulk-copy (TMA) helpers shared by the CTA LU below and getrs_tma ------------------------------------------
__device__ __forceinline__ void gt_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gt_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GT_DONE;\n"
        "bra GT_WAIT;\n"
        "GT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void gt_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// shared -> global bulk store (the TMA unit again): writes by the generic proxy must be fenced before the async proxy
// reads them; completion through the bulk async-group
__device__ __forceinline__ void gt_bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- CTA-per-system blocked LU (lu_blocked.cuh), 33 <= n <= 160: matrix resident in shared memory ---------------------
// TMA staging (AoS, even n, 16-byte aligned batch): every column goes in and out as ONE bulk copy, which needs 16-byte
// aligned shared-memory columns, i.e. an EVEN leading dimension -- n + 2 when n is a multiple of 4 (so that ld / 2 stays
// odd and row accesses across columns are at worst 2-way bank conflicts), else n.  Other shapes keep the odd ld.
__host__ __device__ inline int lub_ld_tma(int n) { return (n & 3) ? n : n + 2; }
__host__ __device__ inline size_t lub_smem_bytes(int n) {
    const int ldmax = lu_ld(n) > lub_ld_tma(n) ? lu_ld(n) : lub_ld_tma(n);
    size_t b = (size_t)ldmax * n * sizeof(double);
    b = (b + 15) / 16 * 16 + (lub_ub_bytes(n) + 15) / 16 * 16 + ((size_t)n * sizeof(int) + 15) / 16 * 16 + sizeof(LubCtl);
    return (b + 15) / 16 * 16 + 16;  // + the staging mbarrier
}
constexpr int lub_threads(int RP) { return RP <= 2 ? 128 : 256; }
constexpr int lub_min_ctas(int RP) { return RP <= 2 ? 4 : (RP <= 4 ? 2 : 1); }

template <bool SOA, int RP, bool TEAM = false>
__global__ void __launch_bounds__(lub_threads(RP), lub_min_ctas(RP))
    getrf_blocked(double* __restrict__ A, int64_t* __restrict__ piv, int32_t* __restrict__ info, int n, int64_t B, int tma) {
    extern __shared__ __align__(16) unsigned char lu_smem[];
    const bool use_tma = !SOA && tma != 0;
    const int ld = use_tma ? lub_ld_tma(n) : lu_ld(n);
    double* Ms = reinterpret_cast<double*>(lu_smem);
    const int ldmax = lu_ld(n) > lub_ld_tma(n) ? lu_ld(n) : lub_ld_tma(n);
    size_t off = ((size_t)ldmax * n * sizeof(double) + 15) / 16 * 16;
    double* Ub = reinterpret_cast<double*>(lu_smem + off);
    off += (lub_ub_bytes(n) + 15) / 16 * 16;
    int* pivs = reinterpret_cast<int*>(lu_smem + off);
    off += ((size_t)n * sizeof(int) + 15) / 16 * 16;
    LubCtl* ctl = reinterpret_cast<LubCtl*>(lu_smem + off);
    off += (sizeof(LubCtl) + 15) / 16 * 16;
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(lu_smem + off);
    const int tid = threadIdx.x, nt = blockDim.x;
    if (use_tma) {
        if (tid == 0) {
            gt_mbar_init(mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    unsigned phase = 0;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        // stage: warp per column (4 columns in flight), lanes along the rows -> coalesced, no index divisions
        LUB_T(ts0);
        if (SOA) {
            const int w = tid >> 5, nw = nt >> 5, lane = tid & 31;
            for (int j = 4 * w; j < n; j += 4 * nw) {
                for (int i = lane; i < n; i += 32) {
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j + u < n) v[u] = A[mat_idx<SOA>(i, j + u, n, b, B)];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j + u < n) Ms[(j + u) * ld + i] = v[u];
                }
            }
        } else if (use_tma) {
            // AoS, TMA-staged: the whole matrix arrives as n bulk copies (cp.async.bulk, one per column, n * 8 bytes each)
            // issued by the threads of warp 0 and counted in bytes on one mbarrier every thread then waits on -- no
            // per-element copy instructions at all (the per-element LDGSTS staging below costs n^2 / threads of them)
            const double* Ab = A + (size_t)b * n * n;
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(Ms);
            if (tid == 0) gt_mbar_expect_tx(mbar, (unsigned)(n * n) * 8u);
            __syncthreads();  // the expected byte count is in place before any copy can complete
            if (tid < 32)
                for (int j = tid; j < n; j += 32) gt_bulk_g2s(sbase + (unsigned)(j * ld) * 8u, Ab + (size_t)j * n, (unsigned)n * 8u, mbar);
            gt_mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            // AoS: the matrix is n*n contiguous doubles; 8-byte asynchronous copies (LDGSTS) straight into the padded
            // shared-memory columns, all in flight at once, no registers and no per-load round trip
            const double* Ab = A + (size_t)b * n * n;
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(Ms);
            const int w = tid >> 5, nw = nt >> 5, lane = tid & 31;
            for (int j = w; j < n; j += nw) {
                const double* src = Ab + (size_t)j * n;
                const unsigned dst = sbase + (unsigned)(j * ld) * 8u;
                for (int i = lane; i < n; i += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (unsigned)i * 8u), "l"(src + i) : "memory");
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
        LUB_T(ts1);
        LUB_ACC(7, ts0, ts1);
        if (TEAM) lub_factor<RP>(Ms, ld, n, ctl, pivs, Ub);  // B200IDA_LU_TEAM_PANEL: the IDA team kernel's flavour
        else lub_factor_1w<RP>(Ms, ld, n, ctl, pivs, Ub);
        if (use_tma) {
            // factors back to global the same way: one bulk store per column (lub_factor ended with a CTA barrier)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            double* Ab = A + (size_t)b * n * n;
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(Ms);
            if (tid < 32) {
                for (int j = tid; j < n; j += 32) gt_bulk_s2g(Ab + (size_t)j * n, sbase + (unsigned)(j * ld) * 8u, (unsigned)n * 8u);
                gt_bulk_commit_wait();  // the shared-memory columns may be overwritten by the next system after this
            }
        } else {
            const int w = tid >> 5, nw = nt >> 5, lane = tid & 31;
            for (int j = 4 * w; j < n; j += 4 * nw) {
                for (int i = lane; i < n; i += 32) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j + u < n) A[mat_idx<SOA>(i, j + u, n, b, B)] = Ms[(j + u) * ld + i];
                }
            }
        }
        const int inf = ctl->info;
        for (int k = tid; k < n; k += nt) piv[SOA ? ((size_t)k * B + b) : ((size_t)b * n + k)] = (inf == 0 || k < inf) ? pivs[k] : 0;
        if (tid == 0) info[b] = inf;
        __syncthreads();
    }
}

template <bool SOA, int RP, bool TEAM = false>
static int launch_getrf_blocked(double* A, int64_t* piv, int32_t* info, int n, int64_t B, cudaStream_t st) {
    const size_t smem = lub_smem_bytes(n);
    unsigned grid = 1;
    int rc;
    // TMA-staged panels when every column segment is 16-byte aligned: AoS, even n, aligned batch
    const int tma = (!SOA && (n & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && !experiment_env("B200IDA_GETRF_NO_TMA")) ? 1 : 0;
    if ((rc = launch_cfg(getrf_blocked<SOA, RP, TEAM>, lub_threads(RP), smem, B, 1, &grid))) return rc;
    getrf_blocked<SOA, RP, TEAM><<<grid, lub_threads(RP), smem, st>>>(A, piv, info, n, B, tma);
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}
template <bool SOA>
static int launch_getrf_blocked_n(double* A, int64_t* piv, int32_t* info, int n, int64_t B, cudaStream_t st) {
    if (n <= 64) return launch_getrf_blocked<SOA, 2>(A, piv, info, n, B, st);
    if (n <= 96) return launch_getrf_blocked<SOA, 3>(A, piv, info, n, B, st);
    if (n <= 128) return launch_getrf_blocked<SOA, 4>(A, piv, info, n, B, st);
    return launch_getrf_blocked<SOA, 5>(A, piv, info, n, B, st);
}
// B200IDA_LU_TEAM_PANEL (AoS, 33 <= n <= SMEM_MAX_N): the same kernel with the CTA-team flavour of the factorisation
static int launch_getrf_blocked_team(double* A, int64_t* piv, int32_t* info, int n, int64_t B, cudaStream_t st) {
    if (n <= 64) return launch_getrf_blocked<false, 2, true>(A, piv, info, n, B, st);
    if (n <= 96) return launch_getrf_blocked<false, 3, true>(A, piv, info, n, B, st);
    if (n <= 128) return launch_getrf_blocked<false, 4, true>(A, piv, info, n, B, st);
    return launch_getrf_blocked<false, 5, true>(A, piv, info, n, B, st);
}

// ---- warp-per-system streaming getrs (AoS, n > 32): every element of the factors is used exactly once, so the
// matrix is never staged -- each warp streams its matrix from HBM/L2 straight into registers, one chunk of columns
// ahead of the substitution chain (register double buffer), and only the triangle each sweep needs is fetched
// (32-row slot granularity: ~n^2 + 32n doubles in total).  The right-hand side lives in registers (lane owns rows
// lane, lane+32, ..), a step of the chain is one shuffle + mul + sub; the divisions of the backward sweep use
// the exact reciprocal sequence (exact_div.cuh) with 1/u_kk computed as soon as the column arrives, off the chain.
template <int NS, int CB>
__device__ __forceinline__ void gs_load_fwd(const double* __restrict__ Mg, int n, int k0, int kend, int rbase, int lane,
                                            double (&c)[CB][NS]) {
#pragma unroll
    for (int q = 0; q < CB; ++q) {
        const int k = k0 + q;
        const double* col = Mg + (size_t)(k < kend ? k : kend - 1) * n;
#pragma unroll
        for (int r = 0; r < NS; ++r) {
            const int i = lane + 32 * (rbase + r);
            c[q][r] = (k < kend && i < n) ? __ldg(col + i) : 0.0;
        }
    }
}

template <int RMAX, int R0>
__device__ __forceinline__ void gs_fwd_block(const double* __restrict__ Mg, int n, int lane, double (&x)[RMAX]) {
    constexpr int NS = RMAX - R0;               // slots R0 .. RMAX-1 hold the rows below the pivots of this block
    constexpr int CB = (NS >= 5) ? 4 : 8;       // columns per chunk (register budget)
    const int kbeg = 32 * R0;
    const int kend = (kbeg + 32 < n - 1) ? kbeg + 32 : n - 1;
    if (kbeg < kend) {
        double cur[CB][NS], nxt[CB][NS];
        gs_load_fwd<NS, CB>(Mg, n, kbeg, kend, R0, lane, cur);
        for (int k0 = kbeg; k0 < kend; k0 += CB) {
            gs_load_fwd<NS, CB>(Mg, n, k0 + CB, kend, R0, lane, nxt);  // predicated off past the block
#pragma unroll
            for (int q = 0; q < CB; ++q) {
                const int k = k0 + q;
                if (k < kend) {  // warp-uniform
                    const int kk = k & 31;
                    const double bk = __shfl_sync(0xffffffffu, x[R0], kk);
                    const double t0 = x[R0] - cur[q][0] * bk;
                    x[R0] = (lane > kk) ? t0 : x[R0];
#pragma unroll
                    for (int r = 1; r < NS; ++r) {
                        const double t = x[R0 + r] - cur[q][r] * bk;
                        x[R0 + r] = (lane + 32 * (R0 + r) < n) ? t : x[R0 + r];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CB; ++q)
#pragma unroll
                for (int r = 0; r < NS; ++r) cur[q][r] = nxt[q][r];
        }
    }
    if constexpr (R0 + 1 < RMAX) gs_fwd_block<RMAX, R0 + 1>(Mg, n, lane, x);
}

// backward chunk: columns k0, k0-1, .., k0-CB+1 (descending), rows of slots 0..R0; rd = 1/u_kk of each column
template <int NS, int CB>
__device__ __forceinline__ void gs_load_bwd(const double* __restrict__ Mg, int n, int k0, int kbeg, int lane,
                                            double (&c)[CB][NS], double (&ud)[CB], double (&rd)[CB]) {
#pragma unroll
    for (int q = 0; q < CB; ++q) {
        const int k = k0 - q;
        const bool on = k >= kbeg;
        const double* col = Mg + (size_t)(on ? k : kbeg) * n;
#pragma unroll
        for (int r = 0; r < NS; ++r) {
            const int i = lane + 32 * r;
            c[q][r] = (on && i < k) ? __ldg(col + i) : 0.0;  // strictly above the diagonal
        }
        ud[q] = on ? __ldg(col + k) : 1.0;  // same address for the whole warp
        rd[q] = 1.0 / ud[q];
    }
}

template <int RMAX, int R0>
__device__ __forceinline__ void gs_bwd_block(const double* __restrict__ Mg, int n, int lane, double (&x)[RMAX]) {
    constexpr int NS = R0 + 1;                  // slots 0 .. R0 hold the rows above the pivots of this block
    constexpr int CB = (NS >= 5) ? 4 : 8;
    const int kbeg = 32 * R0;
    if (kbeg < n) {
        const int ktop = (kbeg + 31 < n - 1) ? kbeg + 31 : n - 1;
        double cur[CB][NS], nxt[CB][NS], udc[CB], rdc[CB], udn[CB], rdn[CB];
        gs_load_bwd<NS, CB>(Mg, n, ktop, kbeg, lane, cur, udc, rdc);
        for (int k0 = ktop; k0 >= kbeg; k0 -= CB) {
            gs_load_bwd<NS, CB>(Mg, n, k0 - CB, kbeg, lane, nxt, udn, rdn);
#pragma unroll
            for (int q = 0; q < CB; ++q) {
                const int k = k0 - q;
                if (k >= kbeg) {  // warp-uniform
                    const int kk = k & 31;
                    double xv = x[R0];
                    if (lane == kk) xv = div_by_recip(xv, udc[q], rdc[q]);
                    const double bk = __shfl_sync(0xffffffffu, xv, kk);
                    const double t0 = x[R0] - cur[q][R0] * bk;
                    x[R0] = (lane == kk) ? bk : ((lane < kk) ? t0 : x[R0]);
#pragma unroll
                    for (int r = 0; r < R0; ++r) x[r] -= cur[q][r] * bk;
                }
            }
#pragma unroll
            for (int q = 0; q < CB; ++q) {
#pragma unroll
                for (int r = 0; r < NS; ++r) cur[q][r] = nxt[q][r];
                udc[q] = udn[q];
                rdc[q] = rdn[q];
            }
        }
    }
    if constexpr (R0 > 0) gs_bwd_block<RMAX, R0 - 1>(Mg, n, lane, x);
}

constexpr int GS_WARPS = 4;  // warps (= systems in flight) per CTA
template <int RMAX>
__global__ void __launch_bounds__(32 * GS_WARPS) getrs_stream(const double* __restrict__ A, const int64_t* __restrict__ piv,
                                                               double* __restrict__ bx, int n, int64_t B) {
    extern __shared__ __align__(16) unsigned char lu_smem[];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    double* xs = reinterpret_cast<double*>(lu_smem) + (size_t)wi * n;
    int* ps = reinterpret_cast<int*>(lu_smem + (size_t)GS_WARPS * n * sizeof(double)) + (size_t)wi * n;
    for (int64_t b = (int64_t)blockIdx.x * GS_WARPS + wi; b < B; b += (int64_t)gridDim.x * GS_WARPS) {
        const double* Mg = A + (size_t)b * n * n;
        for (int i = lane; i < n; i += 32) {
            xs[i] = bx[(size_t)b * n + i];
            ps[i] = (int)piv[(size_t)b * n + i];
        }
        __syncwarp();
        if (lane == 0) {  // the row interchanges, in order (dense.rs:176-181)
            for (int k = 0; k < n; ++k) {
                const int pk = ps[k];
                if (pk != k) {
                    const double t = xs[k];
                    xs[k] = xs[pk];
                    xs[pk] = t;
                }
            }
        }
        __syncwarp();
        double x[RMAX];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int i = lane + 32 * r;
            x[r] = (i < n) ? xs[i] : 0.0;
        }
        gs_fwd_block<RMAX, 0>(Mg, n, lane, x);
        gs_bwd_block<RMAX, RMAX - 1>(Mg, n, lane, x);
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int i = lane + 32 * r;
            if (i < n) bx[(size_t)b * n + i] = x[r];
        }
        __syncwarp();
    }
}

template <int RMAX>
static int launch_getrs_stream(const double* A, const int64_t* piv, double* bx, int n, int64_t B, cudaStream_t st) {
    const size_t smem = (size_t)GS_WARPS * n * (sizeof(double) + sizeof(int));
    const int64_t want = (B + GS_WARPS - 1) / GS_WARPS, cap = (int64_t)sm_count() * 8;
    getrs_stream<RMAX><<<(unsigned)(want < cap ? want : cap), 32 * GS_WARPS, smem, st>>>(A, piv, bx, n, B);
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

// ---- CTA-per-system blocked LU for matrices too big for shared memory (161 <= n <= 256): matrix in place in global
// memory / L2, panels double-buffered in shared memory (lub_factor_global), AoS only ------------------------------------
__host__ __device__ inline size_t lubg_smem_bytes(int n) {
    size_t b = (size_t)2 * LUB_NB * (n | 1) * sizeof(double);
    b = (b + 15) / 16 * 16 + (lub_ub_bytes(n) + 15) / 16 * 16 + ((size_t)n * sizeof(int) + 15) / 16 * 16 + sizeof(LubCtl);
    return (b + 15) / 16 * 16;
}
template <int NT>
__global__ void __launch_bounds__(NT, NT <= 128 ? 4 : (NT <= 256 ? 2 : 1))
    getrf_blocked_global(double* __restrict__ A, int64_t* __restrict__ piv, int32_t* __restrict__ info, int n, int64_t B, int dmma) {
    extern __shared__ __align__(16) unsigned char lu_smem[];
    const int ldp = n | 1;
    double* Pb0 = reinterpret_cast<double*>(lu_smem);
    double* Pb1 = Pb0 + (size_t)LUB_NB * ldp;
    size_t off = ((size_t)2 * LUB_NB * ldp * sizeof(double) + 15) / 16 * 16;
    double* Ub = reinterpret_cast<double*>(lu_smem + off);
    off += (lub_ub_bytes(n) + 15) / 16 * 16;
    int* pivs = reinterpret_cast<int*>(lu_smem + off);
    off += ((size_t)n * sizeof(int) + 15) / 16 * 16;
    LubCtl* ctl = reinterpret_cast<LubCtl*>(lu_smem + off);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        lub_factor_global<8>(A + (size_t)b * n * n, n, ctl, pivs, Ub, Pb0, Pb1, dmma != 0);
        const int inf = ctl->info;
        for (int k = tid; k < n; k += nt) piv[(size_t)b * n + k] = (inf == 0 || k < inf) ? pivs[k] : 0;
        if (tid == 0) info[b] = inf;
        __syncthreads();
    }
}
static int launch_getrf_blocked_global(double* A, int64_t* piv, int32_t* info, int n, int64_t B, cudaStream_t st, int dmma = 0) {
    // the tensor-core update stores 16-byte pairs: even n and a 16-byte aligned batch, else the exact path
    if ((n & 1) || (reinterpret_cast<uintptr_t>(A) & 15)) dmma = 0;
    const size_t smem = lubg_smem_bytes(n);
    unsigned grid = 1;
    int rc;
    // few, large systems (one per SM): the trailing update, not the pivot chain, is the long pole -> 16 warps
    const char* force = experiment_env("B200IDA_LUBG_THREADS");
    if (force && atoi(force) == 128) {
        if ((rc = launch_cfg(getrf_blocked_global<128>, 128, smem, B, 1, &grid))) return rc;
        getrf_blocked_global<128><<<grid, 128, smem, st>>>(A, piv, info, n, B, dmma);
    } else if (force && atoi(force) == 256) {
        if ((rc = launch_cfg(getrf_blocked_global<256>, 256, smem, B, 1, &grid))) return rc;
        getrf_blocked_global<256><<<grid, 256, smem, st>>>(A, piv, info, n, B, dmma);
    } else if (n <= SMEM_MAX_N) {
        // 115..160: the shared-memory-resident kernel fits one CTA per SM only; four small CTAs per SM on L2-resident
        // matrices keep four pivot chains in flight instead (measured: n = 128 0.47 -> 0.37 ms)
        if ((rc = launch_cfg(getrf_blocked_global<128>, 128, smem, B, 1, &grid))) return rc;
        getrf_blocked_global<128><<<grid, 128, smem, st>>>(A, piv, info, n, B, dmma);
    } else if (B <= (int64_t)sm_count() && n >= 224) {
        if ((rc = launch_cfg(getrf_blocked_global<512>, 512, smem, B, 1, &grid))) return rc;
        getrf_blocked_global<512><<<grid, 512, smem, st>>>(A, piv, info, n, B, dmma);
    } else {
        if ((rc = launch_cfg(getrf_blocked_global<256>, 256, smem, B, 1, &grid))) return rc;
        getrf_blocked_global<256><<<grid, 256, smem, st>>>(A, piv, info, n, B, dmma);
    }
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

// ---- warp-per-system getrs with TMA-staged shared-memory panels (AoS, even n > 32, 16-byte aligned batch) -----------------
// Each warp streams its matrix through a private ring of GT_STAGES shared-memory panels of GT_CB columns filled by
// BULK asynchronous copies (cp.async.bulk, the TMA unit): one instruction per panel column, issued by the first
// GT_CB lanes in parallel, completion counted in bytes on a per-stage mbarrier the whole warp then waits on.  The
// copies run GT_STAGES-1 panels ahead of the substitution chain and fetch only the part of each column the sweep
// needs (rows >= k0 in the forward sweep, rows <= k in the backward sweep), i.e. ~n^2 + 8n doubles per system.
// Why TMA and not per-element copies: with a handful of warps per SM the kernel is bound by the number of
// instructions it issues per byte; the same ring filled by 8-byte LDGSTS copies was 1.7-2x slower than no ring at all
// (DESIGN.md section 5.3).  Even n keeps every column segment 16-byte aligned; other shapes take getrs_stream.
constexpr int GT_CB = 8;
constexpr int GT_STAGES = 4;
constexpr int GT_WARPS = 2;

struct GtPlan {  // chunk schedule of one system: forward chunks 0..nf-1, then the backward chunks
    int n, nf, ntop, nc, rbb;
    __device__ __forceinline__ explicit GtPlan(int n_) : n(n_) {
        nf = (n - 1 + GT_CB - 1) / GT_CB;                    // forward: columns 0 .. n-2
        rbb = (n + 31) >> 5;                                 // 32-column blocks
        ntop = (n - 32 * (rbb - 1) + GT_CB - 1) / GT_CB;      // backward chunks of the (ragged) top block
        nc = nf + ntop + (rbb - 1) * (32 / GT_CB);
    }
    __device__ __forceinline__ void bwd(int jb, int& k0, int& klo) const {  // first (highest) column, block start
        int R0, m;
        if (jb < ntop) {
            R0 = rbb - 1;
            m = jb;
        } else {
            R0 = rbb - 2 - (jb - ntop) / (32 / GT_CB);
            m = (jb - ntop) % (32 / GT_CB);
        }
        klo = 32 * R0;
        const int ktop = (klo + 31 < n - 1) ? klo + 31 : n - 1;
        k0 = ktop - GT_CB * m;
    }
};

// issue chunk j into ring stage j % GT_STAGES: the chunk's columns are contiguous in the column-major matrix, so one
// elected lane arms the stage's barrier and issues ONE bulk copy of whole columns (the kernel is bound by instructions
// issued per system, not by bytes: fetching the half of each column the sweep does not use costs less than issuing
// one trimmed copy per column -- cp.async.bulk is a uniform-datapath instruction, per-lane copies are serialised)
__device__ __forceinline__ void gt_issue(const GtPlan& P, const double* __restrict__ Mg, unsigned ring, unsigned bars, int j,
                                         int lane) {
    if (j >= P.nc || lane != 0) return;
    const int s = j % GT_STAGES, n = P.n;
    const unsigned st = ring + (unsigned)(s * GT_CB * n) * 8u, bar = bars + 8u * (unsigned)s;
    int kmin, ncols;
    if (j < P.nf) {
        kmin = j * GT_CB;
        ncols = (n - 1 - kmin < GT_CB) ? n - 1 - kmin : GT_CB;
    } else {
        int k0, klo;
        P.bwd(j - P.nf, k0, klo);
        kmin = (k0 - (GT_CB - 1) > klo) ? k0 - (GT_CB - 1) : klo;
        ncols = k0 - kmin + 1;
    }
    const unsigned bytes = (unsigned)(ncols * n) * 8u;
    gt_mbar_expect_tx(bar, bytes);
    gt_bulk_g2s(st, Mg + (size_t)kmin * n, bytes, bar);
}

template <int RMAX, int R0>
__device__ __forceinline__ void gt_fwd_block(const GtPlan& P, const double* __restrict__ Mg, unsigned ring, unsigned bars,
                                             const double* ringp, int lane, double (&x)[RMAX], unsigned& phases) {
    constexpr int NS = RMAX - R0;  // slots R0 .. RMAX-1 hold the rows below the pivots of this block
    const int n = P.n;
    const int kend = (32 * R0 + 32 < n - 1) ? 32 * R0 + 32 : n - 1;
    for (int k0 = 32 * R0; k0 < kend; k0 += GT_CB) {
        const int j = k0 / GT_CB, s = j % GT_STAGES;
        gt_issue(P, Mg, ring, bars, j + GT_STAGES - 1, lane);
        gt_mbar_wait(bars + 8u * (unsigned)s, (phases >> s) & 1u);
        phases ^= 1u << s;
        const double* st = ringp + s * GT_CB * n;
        const int qn = (kend - k0 < GT_CB) ? kend - k0 : GT_CB;
#pragma unroll 4
        for (int q = 0; q < qn; ++q) {
            const int kk = (k0 + q) & 31;
            double c[NS];
#pragma unroll
            for (int r = 0; r < NS; ++r) {
                const int i = lane + 32 * (R0 + r);
                c[r] = st[q * n + (i < n ? i : n - 1)];
            }
            const double bk = __shfl_sync(0xffffffffu, x[R0], kk);
            const double t0 = x[R0] - c[0] * bk;
            x[R0] = (lane > kk && lane + 32 * R0 < n) ? t0 : x[R0];
#pragma unroll
            for (int r = 1; r < NS; ++r) {
                const double t = x[R0 + r] - c[r] * bk;
                x[R0 + r] = (lane + 32 * (R0 + r) < n) ? t : x[R0 + r];
            }
        }
        __syncwarp();
    }
    if constexpr (R0 + 1 < RMAX) {
        if (32 * (R0 + 1) < n - 1) gt_fwd_block<RMAX, R0 + 1>(P, Mg, ring, bars, ringp, lane, x, phases);
    }
}

// SAFE = false: the division of the pivot lane is straight-line code -- EVERY lane forms the exact-reciprocal quotient of its
// own (x, u, 1/u) each step (three FP64 instructions, exact_div.cuh) and only the pivot lane's is broadcast; a zero numerator
// takes q0 = x * (1/u), the zero of the right sign; the range test that guards the sequence runs once per panel on the
// numerator each lane recorded when it was the pivot lane, and a failure sets `bad`: the caller then redoes the whole
// system with SAFE = true (the guarded division under a one-lane branch, which is what used to sit on every step's chain).
template <int RMAX, int R0, bool SAFE>
__device__ __forceinline__ void gt_bwd_block(const GtPlan& P, const double* __restrict__ Mg, unsigned ring, unsigned bars,
                                             const double* ringp, int lane, double (&x)[RMAX], unsigned& phases, int& j, bool& bad) {
    constexpr int NS = R0 + 1;  // slots 0 .. R0 hold the rows above the pivots of this block
    const int n = P.n;
    const int klo = 32 * R0;
    if (klo < n) {
        const int ktop = (klo + 31 < n - 1) ? klo + 31 : n - 1;
        for (int k0 = ktop; k0 >= klo; k0 -= GT_CB, ++j) {
            const int s = j % GT_STAGES;
            gt_issue(P, Mg, ring, bars, j + GT_STAGES - 1, lane);
            gt_mbar_wait(bars + 8u * (unsigned)s, (phases >> s) & 1u);
            phases ^= 1u << s;
            const double* st = ringp + s * GT_CB * n;
            const int qn = (k0 - klo + 1 < GT_CB) ? k0 - klo + 1 : GT_CB;
            // the lane that owns pivot row k (lane == k & 31) fetches u_kk and its reciprocal for its column of the
            // panel up front: one division per panel and lane, off the chain
            const int myq = (k0 - lane) & 31;  // column k0 - myq has k & 31 == lane
            double myud = 1.0;
            if (myq < qn) myud = st[(qn - 1 - myq) * n + (k0 - myq)];  // columns ascend in the stage: k0 - q at slot qn-1-q
            const double myrd = 1.0 / myud;
            double xp = 1.0;  // SAFE = false: this lane's numerator at its pivot step of the panel
#pragma unroll 4
            for (int q = 0; q < qn; ++q) {
                const int kk = (k0 - q) & 31;
                double c[NS];
#pragma unroll
                for (int r = 0; r < NS; ++r) c[r] = st[(qn - 1 - q) * n + lane + 32 * r];  // r < R0: always < k; r == R0: used if < k
                double xv = x[R0];
                if (SAFE) {
                    if (lane == kk) xv = div_by_recip(xv, myud, myrd);
                } else {
                    const double q0 = __dmul_rn(xv, myrd);
                    const double rr = __fma_rn(-myud, q0, xv);
                    const double q1 = __fma_rn(rr, myrd, q0);
                    const double qq = (xv == 0.0) ? q0 : q1;
                    xp = (lane == kk) ? xv : xp;
                    xv = (lane == kk) ? qq : xv;
                }
                const double bk = __shfl_sync(0xffffffffu, xv, kk);
                const double t0 = x[R0] - c[R0] * bk;
                x[R0] = (lane == kk) ? bk : ((lane < kk) ? t0 : x[R0]);
#pragma unroll
                for (int r = 0; r < R0; ++r) x[r] -= c[r] * bk;
            }
            if (!SAFE) {  // the range tests of exact_div.cuh, once per panel and lane
                const unsigned ex = ((unsigned)__double2hiint(xp) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(myud) >> 20) & 0x7ffu;
                const bool ok = (xp == 0.0) ? (ed - 1u < 0x7feu) : ((ex - 543u) <= 960u && (ed - 543u) <= 960u);
                bad = bad || !ok;
            }
            __syncwarp();
        }
    }
    if constexpr (R0 > 0) gt_bwd_block<RMAX, R0 - 1, SAFE>(P, Mg, ring, bars, ringp, lane, x, phases, j, bad);
}

__host__ __device__ inline size_t gt_warp_smem(int n) {  // ring (16-byte aligned: n even) + rhs + pivots + barriers
    return ((size_t)GT_STAGES * GT_CB * n * sizeof(double) + (size_t)n * sizeof(double) + (size_t)n * sizeof(int) + 8 * GT_STAGES + 15) / 16 * 16;
}

// one system by one warp; returns true (warp-uniform) when SAFE = false met a quotient outside the straight-line sequence's
// range -- nothing has been written then and the caller repeats the system with SAFE = true
template <int RMAX, bool SAFE>
__device__ __forceinline__ bool gt_solve_one(const GtPlan& P, const double* __restrict__ Mg, const int64_t* __restrict__ pivb,
                                             double* __restrict__ bxb, unsigned ring, unsigned bars, const double* ringp, double* xs,
                                             int* ps, int lane, unsigned& phases) {
    const int n = P.n;
    for (int j = 0; j < GT_STAGES - 1; ++j) gt_issue(P, Mg, ring, bars, j, lane);  // the factors start streaming first
    for (int i = lane; i < n; i += 32) {
        xs[i] = bxb[i];
        ps[i] = (int)pivb[i];
    }
    __syncwarp();
    if (lane == 0) {  // the row interchanges, in order (dense.rs:176-181)
        for (int k = 0; k < n; ++k) {
            const int pk = ps[k];
            if (pk != k) {
                const double t = xs[k];
                xs[k] = xs[pk];
                xs[pk] = t;
            }
        }
    }
    __syncwarp();
    double x[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const int i = lane + 32 * r;
        x[r] = (i < n) ? xs[i] : 0.0;
    }
    gt_fwd_block<RMAX, 0>(P, Mg, ring, bars, ringp, lane, x, phases);
    int j = P.nf;
    bool bad = false;
    gt_bwd_block<RMAX, RMAX - 1, SAFE>(P, Mg, ring, bars, ringp, lane, x, phases, j, bad);
    const bool redo = !SAFE && __any_sync(0xffffffffu, bad);
    if (!redo) {
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int i = lane + 32 * r;
            if (i < n) bxb[i] = x[r];
        }
    }
    __syncwarp();
    return redo;
}

template <int RMAX>
__global__ void __launch_bounds__(32 * GT_WARPS) getrs_tma(const double* __restrict__ A, const int64_t* __restrict__ piv,
                                                            double* __restrict__ bx, int n, int64_t B) {
    extern __shared__ __align__(16) unsigned char lu_smem[];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    unsigned char* base = lu_smem + (size_t)wi * gt_warp_smem(n);
    double* ringp = reinterpret_cast<double*>(base);
    double* xs = ringp + (size_t)GT_STAGES * GT_CB * n;
    unsigned long long* barp = reinterpret_cast<unsigned long long*>(xs + n);
    int* ps = reinterpret_cast<int*>(barp + GT_STAGES);
    const unsigned ring = (unsigned)__cvta_generic_to_shared(ringp);
    const unsigned bars = (unsigned)__cvta_generic_to_shared(barp);
    if (lane == 0) {
        for (int s = 0; s < GT_STAGES; ++s) gt_mbar_init(bars + 8u * (unsigned)s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phases = 0;  // parity to wait for, one bit per stage
    const GtPlan P(n);
    for (int64_t b = (int64_t)blockIdx.x * GT_WARPS + wi; b < B; b += (int64_t)gridDim.x * GT_WARPS) {
        const double* Mg = A + (size_t)b * n * n;
        if (gt_solve_one<RMAX, false>(P, Mg, piv + (size_t)b * n, bx + (size_t)b * n, ring, bars, ringp, xs, ps, lane, phases))
            gt_solve_one<RMAX, true>(P, Mg, piv + (size_t)b * n, bx + (size_t)b * n, ring, bars, ringp, xs, ps, lane, phases);
    }
}

template <int RMAX>
static int launch_getrs_tma(const double* A, const int64_t* piv, double* bx, int n, int64_t B, cudaStream_t st) {
    const size_t smem = gt_warp_smem(n) * GT_WARPS;
    unsigned grid = 1;
    int rc;
    if ((rc = launch_cfg(getrs_tma<RMAX>, 32 * GT_WARPS, smem, B, GT_WARPS, &grid))) return rc;
    getrs_tma<RMAX><<<grid, 32 * GT_WARPS, smem, st>>>(A, piv, bx, n, B);
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

template <bool FACTOR>
static int launch_group(double* A, int64_t* piv, int32_t* info, double* bx, int n, int64_t B, bool soa, cudaStream_t st) {
    const unsigned gb = (unsigned)lu_group_smem_bytes(n);
    unsigned grid = 1;
    int rc;
#define LAUNCH(SOA_, WARP_, SMEM_, threads, smem, gpc)                                                                   \
    do {                                                                                                               \
        if (FACTOR) {                                                                                                  \
            if ((rc = launch_cfg(getrf_group<SOA_, WARP_, SMEM_>, threads, smem, B, gpc, &grid))) return rc;             \
            getrf_group<SOA_, WARP_, SMEM_><<<grid, threads, smem, st>>>(A, piv, info, n, B, gb);                        \
        } else {                                                                                                       \
            if ((rc = launch_cfg(getrs_group<SOA_, WARP_, SMEM_>, threads, smem, B, gpc, &grid))) return rc;             \
            getrs_group<SOA_, WARP_, SMEM_><<<grid, threads, smem, st>>>(A, piv, bx, n, B, gb);                          \
        }                                                                                                              \
    } while (0)
    if (!soa && (n == 16 || n == 32)) {
        const int threads = 128, wpc = 4;
        const int64_t items = (n == 16) ? (B + 1) / 2 : B;  // n = 16: two systems per warp
        const int64_t want = (items + wpc - 1) / wpc, cap = (int64_t)sm_count() * 16;
        const unsigned g = (unsigned)(want < cap ? want : cap);
        if (FACTOR) {
            if (n == 16) getrf_warp_regs<16><<<g, threads, 0, st>>>(A, piv, info, B);
            else getrf_warp_regs<32><<<g, threads, 0, st>>>(A, piv, info, B);
        } else {
            if (n == 16) getrs_warp_regs<16><<<g, threads, 0, st>>>(A, piv, bx, B);
            else getrs_warp_regs<32><<<g, threads, 0, st>>>(A, piv, bx, B);
        }
    } else if (n <= 32) {
        const int threads = 32 * WARPS_PER_CTA;
        const size_t smem = (size_t)gb * WARPS_PER_CTA;
        if (soa) LAUNCH(true, true, true, threads, smem, WARPS_PER_CTA);
        else LAUNCH(false, true, true, threads, smem, WARPS_PER_CTA);
    } else if (n <= SMEM_MAX_N) {
        const int threads = n <= 64 ? 128 : CTA_THREADS;
        if (soa) LAUNCH(true, false, true, threads, (size_t)gb, 1);
        else LAUNCH(false, false, true, threads, (size_t)gb, 1);
    } else {
        // too big for shared memory: factor / solve in place in global memory (L2-resident), AoS only
        const size_t smem = (size_t)n * sizeof(double) + ((size_t)n * sizeof(int) + 15) / 16 * 16 + sizeof(LuShared) + 16;
        if (!soa) {
            if constexpr (FACTOR) return launch_getrf_blocked_global(A, piv, info, n, B, st);
            else LAUNCH(false, false, false, CTA_THREADS, smem, 1);
        } else {
            DevBuf tmp, ptmp, xtmp;
            B200_CUDA_TRY(tmp.alloc((size_t)n * n * B * sizeof(double)));
            B200_CUDA_TRY(ptmp.alloc((size_t)n * B * sizeof(int64_t)));
            const unsigned tg = (unsigned)(((int64_t)n * n * B + 255) / 256), vg = (unsigned)(((int64_t)n * B + 255) / 256);
            mat_soa_to_aos<<<tg, 256, 0, st>>>(A, tmp.as<double>(), n, B);
            double* A0 = A;
            int64_t* piv0 = piv;
            double* bx0 = bx;
            A = tmp.as<double>();
            if (FACTOR) {
                piv = ptmp.as<int64_t>();
                if ((rc = launch_getrf_blocked_global(A, piv, info, n, B, st))) return rc;
                mat_aos_to_soa<<<tg, 256, 0, st>>>(tmp.as<double>(), A0, n, B);
                vec_aos_to_soa<<<vg, 256, 0, st>>>(reinterpret_cast<const double*>(ptmp.as<int64_t>()), reinterpret_cast<double*>(piv0), n, B);
            } else {
                B200_CUDA_TRY(xtmp.alloc((size_t)n * B * sizeof(double)));
                vec_soa_to_aos<<<vg, 256, 0, st>>>(reinterpret_cast<const double*>(piv0), reinterpret_cast<double*>(ptmp.as<int64_t>()), n, B);
                vec_soa_to_aos<<<vg, 256, 0, st>>>(bx0, xtmp.as<double>(), n, B);
                piv = ptmp.as<int64_t>();
                bx = xtmp.as<double>();
                LAUNCH(false, false, false, CTA_THREADS, smem, 1);
                vec_aos_to_soa<<<vg, 256, 0, st>>>(xtmp.as<double>(), bx0, n, B);
            }
            B200_CUDA_TRY(cudaGetLastError());
            B200_CUDA_TRY(cudaStreamSynchronize(st));  // scratch freed on return
            return B200IDA_OK;
        }
    }
#undef LAUNCH
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

// device-pointer implementations -------------------------------------------------------------------------
static int getrf_dev(double* A, int64_t* piv, int32_t* info, int n, int64_t B, int layout, cudaStream_t st) {
    const int dmma = (layout & B200IDA_LU_DMMA) ? 1 : 0;  // tolerance mode: honoured by the global-memory blocked LU (n >= 115, AoS)
    const bool team = (layout & B200IDA_LU_TEAM_PANEL) != 0;
    layout &= ~(B200IDA_LU_DMMA | B200IDA_LU_TEAM_PANEL);
    const bool soa = layout == B200IDA_LAYOUT_SOA;
    if (team && !soa && n > 32 && n <= SMEM_MAX_N) return launch_getrf_blocked_team(A, piv, info, n, B, st);
    if (soa && n <= 8) {
        switch (n) {
        case 1: launch_getrf_thread<1>(A, piv, info, B, st); break;
        case 2: launch_getrf_thread<2>(A, piv, info, B, st); break;
        case 3: launch_getrf_thread<3>(A, piv, info, B, st); break;
        case 4: launch_getrf_thread<4>(A, piv, info, B, st); break;
        case 5: launch_getrf_thread<5>(A, piv, info, B, st); break;
        case 6: launch_getrf_thread<6>(A, piv, info, B, st); break;
        case 7: launch_getrf_thread<7>(A, piv, info, B, st); break;
        default: launch_getrf_thread<8>(A, piv, info, B, st); break;
        }
        B200_CUDA_TRY(cudaGetLastError());
        return B200IDA_OK;
    }
    if (!soa && n > 32 && (n >= 115 || experiment_env("B200IDA_LUBG_THREADS"))) return launch_getrf_blocked_global(A, piv, info, n, B, st, dmma);
    if (n > 32 && n <= SMEM_MAX_N)
        return soa ? launch_getrf_blocked_n<true>(A, piv, info, n, B, st) : launch_getrf_blocked_n<false>(A, piv, info, n, B, st);
    return launch_group<true>(A, piv, info, nullptr, n, B, soa, st);
}

static int getrs_dev(const double* LU, const int64_t* piv, double* bx, int n, int64_t B, int layout, cudaStream_t st) {
    const bool soa = layout == B200IDA_LAYOUT_SOA;
    if (soa && n <= 8) {
        switch (n) {
        case 1: launch_getrs_thread<1>(LU, piv, bx, B, st); break;
        case 2: launch_getrs_thread<2>(LU, piv, bx, B, st); break;
        case 3: launch_getrs_thread<3>(LU, piv, bx, B, st); break;
        case 4: launch_getrs_thread<4>(LU, piv, bx, B, st); break;
        case 5: launch_getrs_thread<5>(LU, piv, bx, B, st); break;
        case 6: launch_getrs_thread<6>(LU, piv, bx, B, st); break;
        case 7: launch_getrs_thread<7>(LU, piv, bx, B, st); break;
        default: launch_getrs_thread<8>(LU, piv, bx, B, st); break;
        }
        B200_CUDA_TRY(cudaGetLastError());
        return B200IDA_OK;
    }
    if (!soa && n > 32 && (n & 1) == 0 && (reinterpret_cast<uintptr_t>(LU) & 15) == 0 && !experiment_env("B200IDA_GETRS_NO_TMA")) {
        if (n <= 64) return launch_getrs_tma<2>(LU, piv, bx, n, B, st);
        if (n <= 128) return launch_getrs_tma<4>(LU, piv, bx, n, B, st);
        return launch_getrs_tma<8>(LU, piv, bx, n, B, st);
    }
    if (!soa && n > 32) {
        if (n <= 64) return launch_getrs_stream<2>(LU, piv, bx, n, B, st);
        if (n <= 128) return launch_getrs_stream<4>(LU, piv, bx, n, B, st);
        return launch_getrs_stream<8>(LU, piv, bx, n, B, st);
    }
    return launch_group<false>(const_cast<double*>(LU), const_cast<int64_t*>(piv), nullptr, bx, n, B, soa, st);
}

static int wrms_dev(const double* x, const double* w, double* out, int n, int64_t B, int layout, cudaStream_t st) {
    const unsigned grid = (unsigned)((B + 255) / 256);
    if (layout == B200IDA_LAYOUT_SOA) wrms_thread<true><<<grid, 256, 0, st>>>(x, w, out, n, B);
    else wrms_thread<false><<<grid, 256, 0, st>>>(x, w, out, n, B);
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

static int wrms_masked_dev(const double* x, const double* w, const unsigned char* id, double* out, int n, int64_t B, int layout,
                           cudaStream_t st) {
    const unsigned grid = (unsigned)((B + 255) / 256);
    if (layout == B200IDA_LAYOUT_SOA) wrms_masked_thread<true><<<grid, 256, 0, st>>>(x, w, id, out, n, B);
    else wrms_masked_thread<false><<<grid, 256, 0, st>>>(x, w, id, out, n, B);
    B200_CUDA_TRY(cudaGetLastError());
    return B200IDA_OK;
}

static int check_common(const void* a, const void* b, int n, int64_t batch, int layout, int mem, const char* who) {
    if (!a || !b || n < 1 || n > 256 || batch < 1 || (layout != B200IDA_LAYOUT_SOA && layout != B200IDA_LAYOUT_AOS) ||
        (mem != B200IDA_MEM_DEVICE && mem != B200IDA_MEM_HOST))
        return set_error(B200IDA_ERR_ARG, "%s: bad argument (1 <= n <= 256, batch >= 1, valid layout/mem, non-null)", who);
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess || c == 0) {
        cudaGetLastError();
        return set_error(B200IDA_ERR_NO_DEVICE, "%s: no CUDA device", who);
    }
    return B200IDA_OK;
}

}  // namespace b200ida

using namespace b200ida;

extern "C" {

#ifdef LUB_PROFILE
int b200ida_debug_lub_prof(unsigned long long* out16, int reset) {
    if (out16) cudaMemcpyFromSymbol(out16, b200ida::lub_prof, 32 * sizeof(unsigned long long));
    if (reset) {
        unsigned long long z[32] = {0};
        cudaMemcpyToSymbol(b200ida::lub_prof, z, sizeof(z));
    }
    return 0;
}
#endif

int b200ida_getrf_batched(double* A, int64_t* piv, int32_t* info, int n, int64_t batch, int layout, int mem,
                          void* stream) {
    int rc = check_common(A, piv, n, batch, layout & ~(B200IDA_LU_DMMA | B200IDA_LU_TEAM_PANEL), mem, "b200ida_getrf_batched");
    if (rc) return rc;
    if (!info) return set_error(B200IDA_ERR_ARG, "b200ida_getrf_batched: info is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mem == B200IDA_MEM_DEVICE) return getrf_dev(A, piv, info, n, batch, layout, st);
    DevBuf dA, dP, dI;
    const size_t ab = (size_t)n * n * batch * sizeof(double), pb = (size_t)n * batch * sizeof(int64_t),
                 ib = (size_t)batch * sizeof(int32_t);
    B200_CUDA_TRY(dA.alloc(ab));
    B200_CUDA_TRY(dP.alloc(pb));
    B200_CUDA_TRY(dI.alloc(ib));
    B200_CUDA_TRY(cudaMemcpyAsync(dA.p, A, ab, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemsetAsync(dP.p, 0, pb, st));
    if ((rc = getrf_dev(dA.as<double>(), dP.as<int64_t>(), dI.as<int32_t>(), n, batch, layout, st))) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(A, dA.p, ab, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaMemcpyAsync(piv, dP.p, pb, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaMemcpyAsync(info, dI.p, ib, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}

int b200ida_getrs_batched(const double* LU, const int64_t* piv, double* bx, int n, int64_t batch, int layout,
                          int mem, void* stream) {
    int rc = check_common(LU, piv, n, batch, layout, mem, "b200ida_getrs_batched");
    if (rc) return rc;
    if (!bx) return set_error(B200IDA_ERR_ARG, "b200ida_getrs_batched: bx is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mem == B200IDA_MEM_DEVICE) return getrs_dev(LU, piv, bx, n, batch, layout, st);
    DevBuf dA, dP, dX;
    const size_t ab = (size_t)n * n * batch * sizeof(double), pb = (size_t)n * batch * sizeof(int64_t),
                 xb = (size_t)n * batch * sizeof(double);
    B200_CUDA_TRY(dA.alloc(ab));
    B200_CUDA_TRY(dP.alloc(pb));
    B200_CUDA_TRY(dX.alloc(xb));
    B200_CUDA_TRY(cudaMemcpyAsync(dA.p, LU, ab, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dP.p, piv, pb, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dX.p, bx, xb, cudaMemcpyHostToDevice, st));
    if ((rc = getrs_dev(dA.as<double>(), dP.as<int64_t>(), dX.as<double>(), n, batch, layout, st))) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(bx, dX.p, xb, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}

int b200ida_wrms_batched(const double* x, const double* w, double* out, int n, int64_t batch, int layout, int mem,
                         void* stream) {
    int rc = check_common(x, w, n, batch, layout, mem, "b200ida_wrms_batched");
    if (rc) return rc;
    if (!out) return set_error(B200IDA_ERR_ARG, "b200ida_wrms_batched: out is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mem == B200IDA_MEM_DEVICE) return wrms_dev(x, w, out, n, batch, layout, st);
    DevBuf dX, dW, dO;
    const size_t vb = (size_t)n * batch * sizeof(double), ob = (size_t)batch * sizeof(double);
    B200_CUDA_TRY(dX.alloc(vb));
    B200_CUDA_TRY(dW.alloc(vb));
    B200_CUDA_TRY(dO.alloc(ob));
    B200_CUDA_TRY(cudaMemcpyAsync(dX.p, x, vb, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dW.p, w, vb, cudaMemcpyHostToDevice, st));
    if ((rc = wrms_dev(dX.as<double>(), dW.as<double>(), dO.as<double>(), n, batch, layout, st))) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(out, dO.p, ob, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}

int b200ida_wrms_masked_batched(const double* x, const double* w, const unsigned char* id, double* out, int n,
                                int64_t batch, int layout, int mem, void* stream) {
    int rc = check_common(x, w, n, batch, layout, mem, "b200ida_wrms_masked_batched");
    if (rc) return rc;
    if (!out || !id) return set_error(B200IDA_ERR_ARG, "b200ida_wrms_masked_batched: out or id is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mem == B200IDA_MEM_DEVICE) return wrms_masked_dev(x, w, id, out, n, batch, layout, st);
    DevBuf dX, dW, dI, dO;
    const size_t vb = (size_t)n * batch * sizeof(double), ob = (size_t)batch * sizeof(double);
    B200_CUDA_TRY(dX.alloc(vb));
    B200_CUDA_TRY(dW.alloc(vb));
    B200_CUDA_TRY(dI.alloc((size_t)n));
    B200_CUDA_TRY(dO.alloc(ob));
    B200_CUDA_TRY(cudaMemcpyAsync(dX.p, x, vb, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dW.p, w, vb, cudaMemcpyHostToDevice, st));
    B200_CUDA_TRY(cudaMemcpyAsync(dI.p, id, (size_t)n, cudaMemcpyHostToDevice, st));
    if ((rc = wrms_masked_dev(dX.as<double>(), dW.as<double>(), dI.as<unsigned char>(), dO.as<double>(), n, batch, layout, st))) return rc;
    B200_CUDA_TRY(cudaMemcpyAsync(out, dO.p, ob, cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return B200IDA_OK;
}

}  // extern "C"
