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

template <int A, int B, int TPB, bool HOMOG>
static void wide_x_run(lifu_sim* s, int op, int src) {
  constexpr int G = TPB / B;
  const V2Params& Q = s->Q;
  const int nbatch = (int)((long long)Q.Nz * (Q.Ny / 2) / G);
  if (op == 0) {
    wlaunch_persistent(s, kw_x_u<A, B, TPB, HOMOG>, TPB, WStageU<A, B, TPB, HOMOG>::SMEM, nbatch, s->P, s->Q);
  } else if (op == 1) {
    constexpr size_t sma1 = WStageRho<A, B, TPB, HOMOG, true, 1>::SMEM, sma = WStageRho<A, B, TPB, HOMOG, true, 0>::SMEM;
    constexpr size_t sm = WStageRho<A, B, TPB, HOMOG, false, 0>::SMEM;
    if (s->absorbing) {
      if (src) wlaunch_persistent(s, kw_x_rho_p<A, B, TPB, HOMOG, 1, true>, TPB, sma1, nbatch, s->P, s->Q);
      else wlaunch_persistent(s, kw_x_rho_p<A, B, TPB, HOMOG, 0, true>, TPB, sma, nbatch, s->P, s->Q);
    } else {
      if (src) wlaunch_persistent(s, kw_x_rho_p<A, B, TPB, HOMOG, 1, false>, TPB, sm, nbatch, s->P, s->Q);
      else wlaunch_persistent(s, kw_x_rho_p<A, B, TPB, HOMOG, 0, false>, TPB, sm, nbatch, s->P, s->Q);
    }
  } else if (op == 2) {
    const int use_tau = s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION, use_eta = s->alpha_mode != LIFU_ALPHA_NO_DISPERSION;
    wlaunch_persistent(s, kw_x_p<A, B, TPB, HOMOG>, TPB, WStageU<A, B, TPB, HOMOG>::SMEM, nbatch, s->P, s->Q, use_tau, use_eta);
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
  constexpr int TPB_HET = N >= 512 ? 64 : 128;     // heterogeneous media stage 24 bytes per point: keep >= 2 CTAs per SM
  if (s->homogeneous) wide_x_run<A, B, 128, true>(s, op, src);
  else wide_x_run<A, B, TPB_HET, false>(s, op, src);
}

}  // namespace lifu
