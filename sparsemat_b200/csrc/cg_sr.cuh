// Single-reduction CG for the row-block distributed path (Chronopoulos & Gear's rearrangement of the loop at
// linearsolver.rs:41-60): r.r and (A r).r are formed behind the SAME product, so an iteration has ONE all-reduce where the
// reference's formulation has two dependent ones (p.Ap, then r.r):
//
//   U  p = r + (p*beta); s = w + (s*beta); x += (p*alpha); r -= (s*alpha); gamma' = r.r (rank-local)      cg.cu
//   A  w = A r, delta = w.r  — halo exchange + SpMV with the dot fused (spmv.cu / dist.cu)
//   S  all-reduce {delta, gamma'}; stop test on sqrt(gamma'); beta = gamma'/gamma; alpha = gamma'/(delta - beta*gamma'/alpha)
//
// S runs in one thread of the fused dot's finalize kernel (one launch) or of cgsr_scalar_kernel.  Same vector traffic per
// iteration as the two-reduction loop (9 N values + the product); not the reference's arithmetic — the recurrences for s and
// alpha replace A p and p.Ap — so the iteration count may differ by a few and it is opt-in (smb200_dist_cg_solve_sr).
#pragma once
#include "halo.cuh"
#include "reduce.cuh"

namespace smb {

// scalar block slots beside cg.cu's (S_RR = 0 holds gamma, S_PAP = 1..3 the delta partials, S_RR_NEW = 4 the rank-local
// gamma', S_THRESH = 5, S_ITER = 6, S_DONE = 7, S_RES2 = 9 the last global gamma')
enum { SR_ALPHA = 10, SR_BETA = 11, SR_STARTED = 12 };

// Set by dist.cu around the product of a single-reduction iteration (`active`); with S != nullptr — the product is ONE
// launch — the fused dot's finalize kernel also runs step S.
struct CgSrLaunch { double* S = nullptr; double* history = nullptr; unsigned long long hist_cap = 0; bool active = false; };
extern thread_local CgSrLaunch g_cgsr;

#ifdef __CUDACC__
// Step S after the all-reduce, one thread.  delta, gamma_new: global sums (f64); scalars are kept rounded to T like the
// reference's (its alpha, beta, r.r are T).
template <class T>
__device__ __forceinline__ void cgsr_scalars(double* __restrict__ S, double delta, double gamma_new, double* __restrict__ history,
                                             unsigned long long hist_cap) {
    const T gn = (T)gamma_new, dl = (T)delta;
    S[9] = (double)gn;
    if (S[SR_STARTED] == 0.0) {                 // behind r0 = b - A x0, w0 = A r0: no update has happened yet, nothing to test
        S[SR_STARTED] = 1.0;
        S[SR_BETA] = 0.0;
        S[SR_ALPHA] = (double)div_rn(gn, dl);
        S[0] = (double)gn;
        return;
    }
    const double res = sqrt((double)gn);        // f64::sqrt(r_norm_squared.into()), linearsolver.rs:52
    const unsigned long long it = (unsigned long long)S[6];
    if (history && it < hist_cap) history[it] = res;
    S[6] = (double)(it + 1);
    if (res < S[5]) { S[7] = 1.0; return; }
    const T beta = div_rn(gn, (T)S[0]);
    const T den = sub_rn(dl, div_rn(mul_rn(beta, gn), (T)S[SR_ALPHA]));
    S[SR_BETA] = (double)beta;
    S[SR_ALPHA] = (double)div_rn(gn, den);
    S[0] = (double)gn;
}
#endif

}  // namespace smb
