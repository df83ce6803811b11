// wide_x.cu -- x passes of pipeline "wide" (fft_wide.cuh) for ONE factorisation N = WX_A x WX_B of the x axis; the build
// compiles this file once per supported factorisation (-DWX_A=.. -DWX_B=..) so that the translation units build in parallel.
#include <algorithm>

#include "fft_wide.cuh"
#include "sim.cuh"

#ifndef WX_A
#error "compile with -DWX_A=<A> -DWX_B=<B>"
#endif

namespace lifu {

// persistent launch: one CTA per resident slot, as many as the kernel's registers / shared memory allow
template <typename K, typename... Args>
static void wlaunch_persistent(lifu_sim* s, K kernel, int tpb, size_t sm, int nbatch, Args... args) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, tpb, sm) != cudaSuccess || occ < 1) { occ = 1; cudaGetLastError(); }
  dim3 grid((unsigned)std::min(nbatch, s->n_sm * occ));
  kernel<<<grid, tpb, sm, s->stream>>>(args...);
}

// TPB: threads per CTA of the kernels without / with the absorption operands (the latter keep a running sum in shared memory)
template <int A, int B, int TPB, int TPB_ABS, bool HOMOG, bool SM>
static void wide_x_run(lifu_sim* s, int op, int src) {
  const V2Params& Q = s->Q;
  const long long pairs = (long long)Q.Nz * (Q.Ny / 2);
  const int nbatch = (int)(pairs / (TPB / B)), nbatch_abs = (int)(pairs / (TPB_ABS / B));
  if (op == 0) {
    wlaunch_persistent(s, kw_x_u<A, B, TPB, HOMOG, SM>, TPB, WStageU<A, B, TPB, HOMOG, SM>::SMEM, nbatch, s->P, s->Q);
  } else if (op == 1) {
    if (s->absorbing) {
      constexpr size_t sma1 = WStageRho<A, B, TPB_ABS, HOMOG, SM, true, 1>::SMEM, sma = WStageRho<A, B, TPB_ABS, HOMOG, SM, true, 0>::SMEM;
      if (src) wlaunch_persistent(s, kw_x_rho_p<A, B, TPB_ABS, HOMOG, SM, 1, true>, TPB_ABS, sma1, nbatch_abs, s->P, s->Q);
      else wlaunch_persistent(s, kw_x_rho_p<A, B, TPB_ABS, HOMOG, SM, 0, true>, TPB_ABS, sma, nbatch_abs, s->P, s->Q);
    } else {
      constexpr size_t sm = WStageRho<A, B, TPB, HOMOG, SM, false, 0>::SMEM;
      if (src) wlaunch_persistent(s, kw_x_rho_p<A, B, TPB, HOMOG, SM, 1, false>, TPB, sm, nbatch, s->P, s->Q);
      else wlaunch_persistent(s, kw_x_rho_p<A, B, TPB, HOMOG, SM, 0, false>, TPB, sm, nbatch, s->P, s->Q);
    }
  } else if (op == 2) {
    const int use_tau = s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION, use_eta = s->alpha_mode != LIFU_ALPHA_NO_DISPERSION;
    wlaunch_persistent(s, kw_x_p<A, B, TPB, HOMOG, SM>, TPB, WStageU<A, B, TPB, HOMOG, SM>::SMEM, nbatch, s->P, s->Q, use_tau, use_eta);
  }
}

#define WX_NAME2(a, b) wide_x_launch_##a##_##b
#define WX_NAME(a, b) WX_NAME2(a, b)

// op: 0 velocity update, 1 density update (+ equation of state / absorption operands; src = filtered source in Z4[3]),
//     2 absorbing equation of state, 3 x transform of the dense source slab
void WX_NAME(WX_A, WX_B)(lifu_sim* s, int op, int src) {
  constexpr int A = WX_A, B = WX_B, N = A * B;
  if (op == 3) {
    constexpr int G = 256 / B;
    const V2Params& Q = s->Q;
    const int gs = (int)(((long long)Q.nzs * (Q.Ny / 2) + G - 1) / G);
    const size_t sm = (size_t)G * N * 8 + 16 * (N + 1);
    cudaFuncSetAttribute(kw_x_src<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    kw_x_src<A, B><<<gs, 256, sm, s->stream>>>(s->P, s->Q);
    return;
  }
  // long lines: 64-thread CTAs where a group's staging is large (medium rows, running sum of the absorbing kernel), so that
  // at least two CTAs stay resident per SM.  Not staging the medium rows (SM = false: prefetch into L2, load where used)
  // would fit six to eight groups per SM but measured slower -- 768^3: kw_x_u 8.4 -> 9.9 ms, kw_x_rho_abs 10.6 -> 18.7 ms
  // (profiles/r2_wide_summary.md) -- so the rows stay in the stages.
  constexpr bool LONG = N >= 512;
  if (s->homogeneous) wide_x_run<A, B, 128, LONG ? 64 : 128, true, true>(s, op, src);
  else wide_x_run<A, B, LONG ? 64 : 128, LONG ? 64 : 128, false, true>(s, op, src);
}

}  // namespace lifu
