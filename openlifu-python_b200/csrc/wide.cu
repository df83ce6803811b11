// wide.cu -- host side and strided (y, z) passes of pipeline "wide" (fft_wide.cuh): the fused FFT passes for axis
// lengths 64 / 128 / 256 / 512 / 768 / 1024.  The x passes live in wide_x.cu, one translation unit per factorisation.
// Shares its state (V2Params, tables, buffers) with pipeline v2: lifusim.cu::v2_setup allocates, enqueue_step dispatches.
#include <algorithm>
#include <functional>

#include "fft_wide.cuh"
#include "sim.cuh"

namespace lifu {

#define WIDE_X_DECL(a, b) void wide_x_launch_##a##_##b(lifu_sim* s, int op, int src);
WIDE_X_DECL(8, 8) WIDE_X_DECL(8, 16) WIDE_X_DECL(16, 16) WIDE_X_DECL(16, 32) WIDE_X_DECL(24, 32) WIDE_X_DECL(32, 32)

bool wide_ab(int n, int* A, int* B) {
  int a = 0, b = 0;
  switch (n) {
    case 64: a = 8; b = 8; break;
    case 128: a = 8; b = 16; break;
    case 256: a = 16; b = 16; break;
    case 512: a = 16; b = 32; break;
    case 768: a = 24; b = 32; break;
    case 1024: a = 32; b = 32; break;
    default: return false;
  }
  if (A) *A = a;
  if (B) *B = b;
  return true;
}

static void wide_x(lifu_sim* s, int op, int src) {
  switch (s->N[0]) {
    case 64: wide_x_launch_8_8(s, op, src); break;
    case 128: wide_x_launch_8_16(s, op, src); break;
    case 256: wide_x_launch_16_16(s, op, src); break;
    case 512: wide_x_launch_16_32(s, op, src); break;
    case 768: wide_x_launch_24_32(s, op, src); break;
    default: wide_x_launch_32_32(s, op, src); break;
  }
}

template <typename K, typename... Args>
static void wlaunch(K kernel, dim3 grid, int threads, size_t sm, cudaStream_t st, Args... args) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  kernel<<<grid, threads, sm, st>>>(args...);
}

// EXPR sees constexpr int WA, WB = the factorisation of axis length NAX
#define WIDE_AB(NAX, EXPR)                                                  \
  do {                                                                      \
    switch (NAX) {                                                          \
      case 64: { constexpr int WA = 8, WB = 8; EXPR; } break;               \
      case 128: { constexpr int WA = 8, WB = 16; EXPR; } break;             \
      case 256: { constexpr int WA = 16, WB = 16; EXPR; } break;            \
      case 512: { constexpr int WA = 16, WB = 32; EXPR; } break;            \
      case 768: { constexpr int WA = 24, WB = 32; EXPR; } break;            \
      default: { constexpr int WA = 32, WB = 32; EXPR; } break;             \
    }                                                                       \
  } while (0)

// ... and constexpr int WL = lanes parameter of the strided kernels: 16-lane tiles on 32-thread lines when W16 is set
#define WIDE_ABL(NAX, W16, EXPR)                                                         \
  do {                                                                                   \
    if (W16) WIDE_AB(NAX, { constexpr int WL = WB == 32 ? 16 : 8; EXPR; });              \
    else WIDE_AB(NAX, { constexpr int WL = 8; EXPR; });                                  \
  } while (0)

template <int A, int B, int L32> static dim3 wgrid(const V2Params& Q, int n_other, int nz) {
  return dim3((unsigned)(Q.Nx / (2 * Wide<A, B, L32>::LANES) + 1), (unsigned)n_other, (unsigned)nz);
}

// pressure-gradient z pass: two chains per tile, interleaved along grid.x
template <int A, int B, int L32> static dim3 wgrid2(const V2Params& Q, int n_other) {
  return dim3((unsigned)(2 * (Q.Nx / (2 * Wide<A, B, L32>::LANES) + 1)), (unsigned)n_other, 1u);
}

// z chains: one chain per CTA (kw_z).  The persistent prefetching variant (kw_zp, fft_wide.cuh) is built only with
// -DLIFU_WIDE_ZP and taken with LIFU_WIDE_ZPERSIST=1: parity-green but measured SLOWER (768^3: z_grad 4.8 -> 6.6 ms, z_div
// 9.4 -> 10.9 ms, z_absorb 3.3 -> 4.4 ms; profiles/r2_wide_summary.md) -- like the TMA-fed z passes of round 1, keeping the
// next tile in flight does not help a chain that is bound by its own dependent DFT / exchange / barrier sequence.
// returns the number of kernels launched
template <int A, int B, int OP, int L32>
static int wlaunch_z(lifu_sim* s, const V2Params& Q, int nky, int nchain, bool persist) {
  using W = Wide<A, B, L32>;
  cudaStream_t st = s->stream;
#ifndef LIFU_WIDE_ZP
  persist = false;
#endif
  if (!persist) {
    // Split chains (fft_wide.cuh, PART 1 / 2): forward transform + operator and the inverse transforms as two kernels.
    // Measured (profiles/r2_wide_summary.md): the pressure gradient gains (one forward transform instead of one per chain:
    // 768^3 4.84 -> 3.76 ms), divergence and absorption lose (twice the traffic: 9.4 -> 10.4, 3.3 -> 4.7 ms).  Default: the
    // gradient on one GPU only; LIFU_WIDE_ZSPLIT=0 none, =1 every z pass.
    static const int split_env = [] { const char* e = getenv("LIFU_WIDE_ZSPLIT"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    const bool split = split_env == 1 || (split_env < 0 && OP == 0 && Q.G == 0);
    if (split) {
      wlaunch(kw_z<A, B, OP, L32, 1>, wgrid<A, B, L32>(Q, nky, OP == 0 ? 1 : nchain), W::THREADS, W::SMEM, st, s->P, Q);
      if (OP == 0) wlaunch(kw_z<A, B, OP, L32, 2>, wgrid2<A, B, L32>(Q, nky), W::THREADS, W::SMEM, st, s->P, Q);
      else wlaunch(kw_z<A, B, OP, L32, 2>, wgrid<A, B, L32>(Q, nky, nchain), W::THREADS, W::SMEM, st, s->P, Q);
      return 2;
    }
    if (OP == 0) wlaunch(kw_z<A, B, OP, L32>, wgrid2<A, B, L32>(Q, nky), W::THREADS, W::SMEM, st, s->P, Q);
    else wlaunch(kw_z<A, B, OP, L32>, wgrid<A, B, L32>(Q, nky, nchain), W::THREADS, W::SMEM, st, s->P, Q);
    return 1;
  }
#ifdef LIFU_WIDE_ZP
  constexpr size_t sm = WideZP<A, B, L32>::SMEM;
  auto kern = kw_zp<A, B, OP, L32>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W::THREADS, sm) != cudaSuccess || occ < 1) { occ = 1; cudaGetLastError(); }
  const int nxt = Q.Nx / (2 * W::LANES);
  const long long nitem = ((long long)nxt * nky + (nky + W::LANES - 1) / W::LANES) * nchain;
  const int grid = (int)std::min<long long>(nitem, (long long)s->n_sm * occ);
  kern<<<grid, W::THREADS, sm, st>>>(s->P, Q, nchain);
#endif
  return 1;
}

// kind: 0 no source, 1 source active (filtered additive source).
// Slab decomposition (Q.G > 0, `barrier` given): the y-forward kernels store straight into the owners' transposed buffers
// and the z kernels straight back into the owners' plane buffers over NVLink, so an exchange is the store phase of a
// transform kernel plus ONE barrier -- 4 barriers per time step (6 for absorbing media), no pack / unpack pass, no
// library transform.  A phase touches {local Z / H, remote T} or {local T, remote H}: one barrier per phase is enough.
int wide_enqueue_step(lifu_sim* s, int kind, int* n_kernels, const std::function<void(const char*, double)>& mark,
                      const std::function<int()>& barrier) {
  V2Params Q = s->Q;
  cudaStream_t st = s->stream;
  const int Ny = s->N[1], Nz = s->N[2];
  const int src = kind != 0 ? 1 : 0;
  const bool slab = Q.G > 0;
  const int nky = slab ? Q.Nyl : Q.Ny;                          // ky rows of the z passes
  const double srcf = (double)(slab ? Q.gnzs : Q.nzs) / Nz;
  int nk = 0;
  // exchange-bearing kernels of a slab decomposition: 128-byte row segments for the NVLink stores (measured at 2 GPUs:
  // y-forward + exchange stages 30 % shorter); one GPU: 64-byte segments, two CTAs per SM (LIFU_WIDE_LANES=8|16 overrides)
  bool x16 = slab, zdiv16 = true;            // the divergence z pass is the one strided kernel that is faster on 16-lane tiles
  if (const char* e = getenv("LIFU_WIDE_LANES")) x16 = zdiv16 = atoi(e) == 16;
  bool zpersist = false;                       // experiment switch, see wlaunch_z
  if (const char* e = getenv("LIFU_WIDE_ZPERSIST")) zpersist = e[0] == '1';
  auto sync_ranks = [&]() -> int { if (slab && barrier) { ++nk; return barrier(); } return LIFU_OK; };
  // (1) pressure gradient
  WIDE_ABL(Ny, x16, (wlaunch(kw_y_fwd<WA, WB, 0, WL>, wgrid<WA, WB, WL>(Q, Q.Nz, 1), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
  LIFU_CHECK(sync_ranks());
  ++nk; mark(slab ? "kw_y_fwd_p+xchg" : "kw_y_fwd_p", 8);
  WIDE_ABL(Nz, x16, (nk += wlaunch_z<WA, WB, 0, WL>(s, Q, nky, 2, zpersist)));
  LIFU_CHECK(sync_ranks());
  mark(slab ? "kw_z_grad+xchg" : "kw_z_grad", 12);
  WIDE_ABL(Ny, false, (wlaunch(kw_y_inv<WA, WB, true, WL>, wgrid<WA, WB, WL>(Q, Q.Nz, 3), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
  ++nk; mark("kw_y_inv_grad", 20);
  // (2) velocity update + forward x transform of the new velocity
  wide_x(s, 0, 0);
  ++nk; mark("kw_x_u", s->homogeneous ? 48 : 60);
  WIDE_ABL(Ny, x16, (wlaunch(kw_y_fwd<WA, WB, 1, WL>, wgrid<WA, WB, WL>(Q, Q.Nz, 3), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
  ++nk; if (!slab) mark("kw_y_fwd_u", 24);
  // (3) source field on its slab (slab decomposition: on this rank's planes of it, if any)
  if (src) {
    if (Q.nzs > 0) {
      if (s->S.n_src > 0) { k2_source_scatter<<<grid_blocks(s, s->S.n_src, 128), 128, 0, st>>>(s->P, Q, s->S); ++nk; }
      if (!slab) mark("k2_source_scatter", 0);
      wide_x(s, 3, 0);
      ++nk; if (!slab) mark("kw_x_src", 8 * srcf);
      WIDE_ABL(Ny, x16, (wlaunch(kw_y_fwd<WA, WB, 2, WL>, wgrid<WA, WB, WL>(Q, Q.nzs, 1), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
      ++nk; if (!slab) mark("kw_y_fwd_src", 8 * srcf);
    }
  }
  LIFU_CHECK(sync_ranks());
  if (slab) mark("kw_y_fwd_u+src+xchg", 24 + (src ? 16 * srcf : 0));
  // (4) divergence (+ filtered source) through z and back through y
  const int ncomp = src ? 4 : 3;
  Q.comp0 = 0;
  WIDE_ABL(Nz, zdiv16, (nk += wlaunch_z<WA, WB, 1, WL>(s, Q, nky, 3, zpersist)));
  if (src) WIDE_ABL(Nz, zdiv16, (nk += wlaunch_z<WA, WB, 3, WL>(s, Q, nky, 1, zpersist)));
  LIFU_CHECK(sync_ranks());
  mark(slab ? "kw_z_div+xchg" : "kw_z_div", 24 + (src ? 4 + 4 * srcf : 0));
  WIDE_ABL(Ny, false, (wlaunch(kw_y_inv<WA, WB, false, WL>, wgrid<WA, WB, WL>(Q, Q.Nz, ncomp), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
  ++nk; mark("kw_y_inv", 8 * ncomp);
  // (5) density update, source, equation of state, sensor, forward x transform of p
  wide_x(s, 1, src);
  ++nk;
  const double sens = (double)s->n[1] * s->n[2] / ((double)s->N[1] * s->N[2]);
  if (!s->absorbing) {
    mark("kw_x_rho_p", 12 + 24 + 16 * sens + 4 + (s->homogeneous ? 0 : 8) + (src ? 4 : 0));
  } else {
    // (6) absorbing medium: the two fractional Laplacians, then the equation of state
    mark("kw_x_rho_abs", 12 + 24 + 4 + 8 + (s->homogeneous ? 0 : 8) + (src ? 4 : 0));
    WIDE_ABL(Ny, x16, (wlaunch(kw_y_fwd<WA, WB, 3, WL>, wgrid<WA, WB, WL>(Q, Q.Nz, 2), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
    LIFU_CHECK(sync_ranks());
    ++nk; mark(slab ? "kw_y_fwd_abs+xchg" : "kw_y_fwd_abs", 16);
    WIDE_ABL(Nz, x16, (nk += wlaunch_z<WA, WB, 2, WL>(s, Q, nky, 2, zpersist)));
    LIFU_CHECK(sync_ranks());
    mark(slab ? "kw_z_absorb+xchg" : "kw_z_absorb", 16);
    WIDE_ABL(Ny, false, (wlaunch(kw_y_inv<WA, WB, false, WL>, wgrid<WA, WB, WL>(Q, Q.Nz, 2), Wide<WA, WB, WL>::THREADS, Wide<WA, WB, WL>::SMEM, st, s->P, Q)));
    ++nk; mark("kw_y_inv_abs", 16);
    wide_x(s, 2, 0);
    ++nk; mark("kw_x_p", 8 + 4 + 16 * sens + 4 + (s->homogeneous ? 0 : 12));
  }
  LIFU_CUDA(cudaGetLastError());
  if (n_kernels) *n_kernels = nk;
  return LIFU_OK;
}

void wide_pm_crop(lifu_sim* s) {
  kw_pm_crop<<<grid_blocks(s, s->Vsens, 256), 256, 0, s->stream>>>(s->P, s->Q.pm, s->Vsens);
}

}  // namespace lifu
