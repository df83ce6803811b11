// fft_v2.cuh -- pipeline v2: hand-written FFT passes with every spectral / real-space operation of
// the k-space time step fused into them (no library transforms, no separate element-wise kernels).
//
// What this replaces: the 10-12 cuFFT 3-D transforms + 5 element-wise kernels of pipeline v1, i.e. the
// per-step body of kspaceFirstOrder3D reached from /root/reference/src/openlifu/sim/kwave_if.py:124-129.
//
// Arithmetic.  The passes are instruction-issue bound as much as HBM bound (ncu, profiles/r1_v2_*), so
// all complex arithmetic uses the sm_100 packed-FP32 pipe: one FADD2 / FMUL2 / FFMA2 works on a whole
// complex number (register pair).  A 16-point register DFT is 68 packed instructions instead of ~205
// scalar ones; twiddles are stored as (w, i*w) pairs so a complex multiply is FMUL2 + FFMA2.
//
// Transform structure.  A length-N line (N = R*R, R in {8, 16}) is transformed by R threads:
// thread t holds x[t + R*j] (j = 0..R-1) in registers, does a register DFT of size R over j, multiplies
// the inter-stage twiddle w_N^(t*k2), exchanges through shared memory, and does a second register DFT
// over t.  Outputs sit as X[k2 + R*k1] in (thread k2, register k1) -- the same "thread + R*register"
// pattern as the input, so a forward transform can be followed by its inverse without any reordering.
//
// Layouts (float2 = complex):
//   real field   R[z][y][x]                                       (x fastest)
//   Z layout     Z[z][m][kx], kx = 0..Nx-1                        x-spectrum of a ROW PAIR
//                row(y_lo) + i*row(y_lo + Ry), m = (y_lo / 2Ry)*Ry + y_lo % Ry: the two rows of a pair
//                are the ones that land in adjacent registers of ONE thread of the y pass, so splitting
//                and merging pairs needs no shuffles
//   H layout     H[z][ky][kx], kx = 0..Nx/2, row pitch PH         half spectrum
// The x passes therefore are plain complex transforms; the y passes split (on load) or merge (on
// store) the packed row pairs with the Hermitian pairing kx <-> Nx-kx.
//
// Tiles of the strided (y, z) passes: CTA = 16 lanes x R threads.  Lanes are 16 consecutive kx for the
// Nx/32 regular tiles; the single Nyquist column kx = Nx/2 is handled by extra CTAs whose 16 lanes run
// over the other in-plane index instead, so no lane is ever idle.
//
// Kernels (one time step, lossless medium):
//   k2_y_fwd      Z -> H        split row pairs, [x-derivative multiplier], FFT_y, [y-derivative mult.]
//   k2_z_grad     H -> HA, HB   FFT_z, kappa, IFFT_z twice (plain and with i kz e^{+i kz dz/2})
//   k2_y_inv_grad HA,HB -> ZA,ZC,ZB   IFFT_y x3 with the x / y derivative multipliers, merge row pairs
//   k2_x_u        ZA,ZC,ZB -> u, ZU   IFFT_x x3, velocity update with staggered PML, FFT_x of new u
//   k2_z_div      H(3|1) in place     FFT_z, kappa [z-derivative | source cos filter], IFFT_z
//   k2_y_inv      H(3|4) -> Z         IFFT_y, merge row pairs
//   k2_x_rho_p    Z(3|4), rho -> rho, p_max/p_min, ZP   IFFT_x, density update + source, p = c0^2 sum rho,
//                                     sensor reduction, FFT_x of the new pressure for the next step
#pragma once
#include "common.cuh"
#include "step_kernels.cuh"
#include "twiddles32.cuh"

namespace lifu {

// ------------------------------------------------------------------------------------------------
// packed complex helpers (FADD2 / FMUL2 / FFMA2; operand swaps, broadcasts and whole-pair negation
// are free operand modifiers in SASS)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 cadd_i(float2 a, float2 b) { return __ffma2_rn(cswap(b), make_float2(-1.f, 1.f), a); }   // a + i b
__device__ __forceinline__ float2 csub_i(float2 a, float2 b) { return __ffma2_rn(cswap(b), make_float2(1.f, -1.f), a); }   // a - i b
__device__ __forceinline__ float2 cadd_conj(float2 a, float2 b) { return __ffma2_rn(b, make_float2(1.f, -1.f), a); }      // a + conj b
__device__ __forceinline__ float2 csub_conj(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, 1.f), a); }      // a - conj b
__device__ __forceinline__ float2 cmul_mi(float2 a) { return __fmul2_rn(cswap(a), make_float2(1.f, -1.f)); }             // -i a
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a * w, with w given together with i*w:  q = (w.x, w.y, -w.y, w.x)
__device__ __forceinline__ float2 cmul4(float2 a, float4 q) {
  return __ffma2_rn(make_float2(a.x, a.x), make_float2(q.x, q.y), __fmul2_rn(make_float2(a.y, a.y), make_float2(q.z, q.w)));
}
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) { return cmul4(a, make_float4(b.x, b.y, -b.y, b.x)); }
__device__ __forceinline__ float4 with_i(float2 w) { return make_float4(w.x, w.y, -w.y, w.x); }

template <int N, bool INV>
__device__ __forceinline__ float2 cmul_const(float2 a, int m) {   // a * exp(-+2 pi i m / N), N | 32, m compile-time
  const int k = (m * (32 / N)) & 31;
  const float wr = w32_re(k), wi = INV ? -w32_im(k) : w32_im(k);
  return cmul4(a, make_float4(wr, wi, -wi, wr));
}

// Register DFT of size N (2, 4, 8, 16), natural order in and out, unnormalised.
template <int N, bool INV>
__device__ __forceinline__ void dft(float2 (&x)[N]) {
  if constexpr (N == 2) {
    float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  } else if constexpr (N == 4) {
    float2 s0 = cadd(x[0], x[2]), d0 = csub(x[0], x[2]);
    float2 s1 = cadd(x[1], x[3]), d1 = csub(x[1], x[3]);
    x[0] = cadd(s0, s1);
    x[2] = csub(s0, s1);
    if constexpr (!INV) { x[1] = csub_i(d0, d1); x[3] = cadd_i(d0, d1); }
    else                { x[1] = cadd_i(d0, d1); x[3] = csub_i(d0, d1); }
  } else {
    constexpr int A = 4, B = N / 4;   // n = b + B*a, k = ka + A*kb
    float2 y[B][A];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float2 s[A];
#pragma unroll
      for (int a = 0; a < A; ++a) s[a] = x[b + B * a];
      dft<A, INV>(s);
#pragma unroll
      for (int ka = 0; ka < A; ++ka) y[b][ka] = (b * ka == 0) ? s[ka] : cmul_const<N, INV>(s[ka], b * ka);
    }
#pragma unroll
    for (int ka = 0; ka < A; ++ka) {
      float2 c[B];
#pragma unroll
      for (int b = 0; b < B; ++b) c[b] = y[b][ka];
      dft<B, INV>(c);
#pragma unroll
      for (int kb = 0; kb < B; ++kb) x[ka + A * kb] = c[kb];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Strided-axis transforms (y and z passes).  CTA = 16 lanes x R threads per line; thread (l, t):
// l = tid & 15, t = tid >> 4.  `tw` is the shared-memory copy of the (w, i w) table with N + 1 entries
// (entry N = entry 0, so the inverse can index N - m).  `sm` is one exchange buffer of N*16 float2;
// callers alternate between two buffers, which makes a single barrier per transform sufficient.
template <int R, bool INV>
__device__ __forceinline__ void strided_fft(float2 (&v)[R], const float4* __restrict__ tw, float2* sm, int l, int t) {
  constexpr int N = R * R;
  dft<R, INV>(v);
  if constexpr (!INV) {
#pragma unroll
    for (int k2 = 1; k2 < R; ++k2) v[k2] = cmul4(v[k2], tw[t * k2]);
    float2* w = sm + t * (R * 16) + l;
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) w[k2 * 16] = v[k2];
    __syncthreads();
    const float2* r = sm + t * 16 + l;
#pragma unroll
    for (int tt = 0; tt < R; ++tt) v[tt] = r[tt * (R * 16)];      // this thread now owns k2 = t
  } else {
#pragma unroll
    for (int tt = 1; tt < R; ++tt) v[tt] = cmul4(v[tt], tw[N - tt * t]);
    float2* w = sm + t * 16 + l;
#pragma unroll
    for (int tt = 0; tt < R; ++tt) w[tt * (R * 16)] = v[tt];
    __syncthreads();
    const float2* r = sm + t * (R * 16) + l;
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) v[k2] = r[k2 * 16];
  }
  dft<R, INV>(v);
}

template <int R>
struct Strided {
  static constexpr int N = R * R;
  static constexpr int THREADS = 16 * R;
  static constexpr int XCH = N * 16 * 8;                       // bytes of one exchange buffer
  static constexpr int TW = (N + 1) * 16;                      // bytes of the twiddle table
  static constexpr int smem(int nbuf) { return nbuf * XCH + TW; }
  // Stage the twiddle table.  Call it AFTER the first data loads have been issued: the barrier would otherwise
  // serialise the table's global-load latency with theirs (a fifth of a short-lived CTA, profiles/r1_zdiv_source_stalls.txt).
  static __device__ __forceinline__ const float4* load_tw(unsigned char* smraw, int nbuf, const float4* __restrict__ g) {
    float4* s = reinterpret_cast<float4*>(smraw + nbuf * XCH);
    for (int i = threadIdx.x; i <= N; i += THREADS) s[i] = g[i];
    __syncthreads();
    return s;
  }
};

// lane -> (kx, other index) of a strided-pass CTA; false when the CTA has no work.
// grid.x = nxt regular tiles + 1 Nyquist slot; grid.y runs over the other index (regular) or its 16-blocks.
__device__ __forceinline__ bool lane_map(const V2Params& Q, int n_other, int l, int& kx, int& o, int by) {
  const int bx = (int)blockIdx.x + Q.bx0;
  if (bx < Q.nxt) { kx = bx * 16 + l; o = by; return true; }
  if (by * 16 >= n_other) return false;
  kx = Q.Nx >> 1;
  o = by * 16 + l;
  return true;
}
__device__ __forceinline__ bool lane_map(const V2Params& Q, int n_other, int l, int& kx, int& o) {
  return lane_map(Q, n_other, l, kx, o, (int)blockIdx.y);
}
// Batched y passes: grid (tiles, Nz, ncomp), or with Q.zmajor (tiles, ncomp, Nz) so that the components of one plane are
// scheduled together -- the order in which the neighbouring x passes produce / consume them.
__device__ __forceinline__ void comp_plane(const V2Params& Q, int& comp, int& by) {
  comp = Q.zmajor ? (int)blockIdx.y : (int)blockIdx.z;
  by = Q.zmajor ? (int)blockIdx.z : (int)blockIdx.y;
}

// Merge the two rows of a pair (adjacent registers A = row y_lo, B = row y_lo + R) into the packed
// x-spectrum: Z[kx] = A + iB, Z[Nx-kx] = conj(A) + i conj(B) = conj(A - iB).  Bins 0 and Nx/2 keep the
// real parts only, as a C2R transform would.
template <int R>
__device__ __forceinline__ void merge_store(float2* __restrict__ zp, const float2 (&v)[R], int kx, int Nx, bool live) {
  const bool selfm = (kx == 0) || (2 * kx == Nx);
  const int km = Nx - kx;
  const int qstep = R * Nx;
#pragma unroll
  for (int q = 0; q < R / 2; ++q) {
    const float2 A = v[2 * q], B = v[2 * q + 1];
    const float2 lo = selfm ? make_float2(A.x, B.x) : cadd_i(A, B);
    if (live) zp[q * qstep + kx] = lo;
    if (live && !selfm) zp[q * qstep + km] = cconj(csub_i(A, B));
  }
}

// ------------------------------------------------------------------------------------------------
// y forward: packed row pairs -> half spectrum.  grid (nxt+1, Nz | Nz/16.., ncomp)
// MODE 0: pressure (no multipliers)   MODE 1: velocity (comp 0: i kx e^{-i kx dx/2}; comp 1: i ky e^{-i ky dy/2})
// MODE 2: source slab (z index relative to the slab)      MODE 3: Z4[comp] without multipliers (absorption operands)
template <int R, int MODE>
__global__ void __launch_bounds__(16 * R, R == 16 ? 3 : 6) k2_y_fwd(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  int comp, by;
  comp_plane(Q, comp, by);
  const int nz = MODE == 2 ? Q.nzs : Q.Nz;
  int kx, z;
  if (!lane_map(Q, nz, l, kx, z, by)) return;
  const bool live = z < nz;                       // only a slab's Nyquist tile can run past the end
  const int zc = live ? z : nz - 1;
  float2* xa = reinterpret_cast<float2*>(smraw);
  const float2* Zin = MODE == 0 ? Q.ZP : ((MODE == 1 || MODE == 3) ? Q.Z4 + comp * Q.ZS : Q.ZSslab);
  float2* Hout = MODE == 2 ? Q.HSslab : Q.H4 + comp * Q.HS;
  if (MODE == 0 && Q.nws > 0 && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    // first kernel of the step: the steady-source coefficients of this step, read by k2_x_rho_p<SRC = 3>
    const int ts = *P.step - Q.t0s;
    if (ts >= 0 && ts < Q.nws) { Q.qcur[0] = Q.qsrc[ts]; Q.qcur[1] = Q.qsrc[Q.nws + ts]; }
  }
  const int km = (Q.Nx - kx) & (Q.Nx - 1);
  const float2* zp = Zin + ((long long)zc * (Q.Ny / 2) + t) * Q.Nx;     // packed line m = q*R + t
  const int qstep = R * Q.Nx;
  float2 v[R];
#pragma unroll
  for (int q = 0; q < R / 2; ++q) { v[2 * q] = zp[q * qstep + kx]; v[2 * q + 1] = zp[q * qstep + km]; }
  const float4* tw = S::load_tw(smraw, 1, Q.tw4y);
#pragma unroll
  for (int q = 0; q < R / 2; ++q) {
    const float2 d = v[2 * q], m = v[2 * q + 1];
    v[2 * q] = cadd_conj(d, m);                  // 2A = Z[k] + conj Z[-k]          (row t + R*2q)
    v[2 * q + 1] = cmul_mi(csub_conj(d, m));     // 2B = -i (Z[k] - conj Z[-k])     (row t + R*(2q+1))
  }
  if (MODE == 1 && comp == 0) {
    const float4 mx = with_i(P.dnx[kx]);
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = cmul4(v[j], mx);
  }
  strided_fft<R, false>(v, tw, xa, l, t);
  if (live) {
    float2* hp = Hout + ((long long)z * Q.Ny + t) * Q.PH + kx;
    const int kstep = R * Q.PH;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      float2 o = v[k1];
      if (MODE == 1 && comp == 1) o = cmul4(o, Q.dny4[t + R * k1]);
      hp[k1 * kstep] = o;
    }
  }
}

// z pass of the pressure gradient: H4[0] -> H4[0] (kappa p^) and H4[1] (i kz e^{+i kz dz/2} kappa p^),
// both already inverse transformed along z.  grid (nxt+1, Ny)
template <int R, int POLY>
__global__ void __launch_bounds__(16 * R, R == 16 ? 2 : 4) k2_z_grad(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  int kx, ky;
  if (!lane_map(Q, Q.Ny, l, kx, ky)) return;
  float2* xa = reinterpret_cast<float2*>(smraw);
  float2* xb = reinterpret_cast<float2*>(smraw + S::XCH);
  const int zs = Q.Ny * Q.PH, jstep = R * zs;
  float2* hp = Q.H4 + (long long)t * zs + ky * Q.PH + kx;               // z = t + R*j
  float2 v[R], w[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = hp[j * jstep];
  const float axy = P.ax2[kx] + P.ay2[ky];
  const float4* tw = S::load_tw(smraw, 2, Q.tw4z);
  strided_fft<R, false>(v, tw, xa, l, t);
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) {
    const int kz = t + R * k1;
    const float a2 = axy + P.az2[kz];
    const float kap = kappa_sel<POLY>(a2) * Q.norm;
    v[k1] = cscale(v[k1], kap);
    w[k1] = cmul4(v[k1], Q.dpz4[kz]);
  }
  strided_fft<R, true>(v, tw, xb, l, t);
#pragma unroll
  for (int j = 0; j < R; ++j) hp[j * jstep] = v[j];
  strided_fft<R, true>(w, tw, xa, l, t);
  float2* hq = hp + Q.HS;
#pragma unroll
  for (int j = 0; j < R; ++j) hq[j * jstep] = w[j];
}

// y inverse of the three gradient components + row-pair merge.  grid (nxt+1, Nz)
//   Z4[0] <- i kx e^{+i kx dx/2} * IFFT_y[H4[0]]     (d/dx)
//   Z4[1] <-                      IFFT_y[i ky e^{+i ky dy/2} H4[0]]   (d/dy)
//   Z4[2] <-                      IFFT_y[H4[1]]                      (d/dz)
template <int R>
__global__ void __launch_bounds__(16 * R, R == 16 ? 2 : 4) k2_y_inv_grad(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  int kx, z;
  if (!lane_map(Q, Q.Nz, l, kx, z)) return;
  float2* xa = reinterpret_cast<float2*>(smraw);
  float2* xb = reinterpret_cast<float2*>(smraw + S::XCH);
  const float2* hp = Q.H4 + ((long long)z * Q.Ny + t) * Q.PH + kx;      // ky = t + R*k1
  const int kstep = R * Q.PH;
  float2* zp = Q.Z4 + ((long long)z * (Q.Ny / 2) + t) * Q.Nx;
  float2 a[R], c[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) a[k1] = hp[k1 * kstep];
  const float4* tw = S::load_tw(smraw, 2, Q.tw4y);
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) c[k1] = cmul4(a[k1], Q.dpy4[t + R * k1]);
  strided_fft<R, true>(a, tw, xa, l, t);
  const float4 mx = with_i(P.dpx[kx]);
#pragma unroll
  for (int j = 0; j < R; ++j) a[j] = cmul4(a[j], mx);
  merge_store<R>(zp, a, kx, Q.Nx, true);
  strided_fft<R, true>(c, tw, xb, l, t);
  merge_store<R>(zp + Q.ZS, c, kx, Q.Nx, true);
  const float2* hq = hp + Q.HS;
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) a[k1] = hq[k1 * kstep];
  strided_fft<R, true>(a, tw, xa, l, t);
  merge_store<R>(zp + 2 * Q.ZS, a, kx, Q.Nx, true);
}

// The same three gradient components with one component per CTA (grid (nxt+1, Nz, 3)): 80 registers, three CTAs per SM,
// no serial chain of three transforms inside a CTA; components 0 and 1 both read H4[0] (the second read hits L2).
// LIFU_V2_YGRAD=split selects it (profiles/r2_ygrad_split.md).
template <int R>
__global__ void __launch_bounds__(16 * R, R == 16 ? 3 : 6) k2_y_inv_grad_split(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int comp = blockIdx.z;
  int kx, z;
  if (!lane_map(Q, Q.Nz, l, kx, z)) return;
  float2* xa = reinterpret_cast<float2*>(smraw);
  const float2* hp = Q.H4 + (comp == 2 ? Q.HS : 0) + ((long long)z * Q.Ny + t) * Q.PH + kx;      // ky = t + R*k1
  const int kstep = R * Q.PH;
  float2 a[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) a[k1] = hp[k1 * kstep];
  const float4* tw = S::load_tw(smraw, 1, Q.tw4y);
  if (comp == 1) {
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) a[k1] = cmul4(a[k1], Q.dpy4[t + R * k1]);
  }
  strided_fft<R, true>(a, tw, xa, l, t);
  if (comp == 0) {
    const float4 mx = with_i(P.dpx[kx]);
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = cmul4(a[j], mx);
  }
  merge_store<R>(Q.Z4 + comp * Q.ZS + ((long long)z * (Q.Ny / 2) + t) * Q.Nx, a, kx, Q.Nx, true);
}

// z pass of the velocity divergence (comp 0..2, in place) and of the source field (comp 3).
// grid (nxt+1, Ny); the CTA walks the components with the loads of component c+1 in flight while component
// c is transformed (register double buffer), so HBM stays busy through the FFT phases; kappa is evaluated
// once per (kx,ky,kz).  comp 2 additionally gets i kz e^{-i kz dz/2}; comp 3 reads the slab planes only and
// is filtered with cos(c_ref k dt/2).
template <int R, int POLY>
__global__ void __launch_bounds__(16 * R, R == 16 ? 2 : 4) k2_z_div(StepParams P, V2Params Q, int ncomp) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  int kx, ky;
  if (!lane_map(Q, Q.Ny, l, kx, ky)) return;
  float2* xa = reinterpret_cast<float2*>(smraw);
  float2* xb = reinterpret_cast<float2*>(smraw + S::XCH);
  const int zs = Q.Ny * Q.PH, jstep = R * zs;
  float2* hp = Q.H4 + (long long)t * zs + ky * Q.PH + kx;
  const float2* sp = Q.HSslab + (long long)(t - Q.z0s) * zs + ky * Q.PH + kx;
  float2 v[R], nx[R];
  if (Q.comp0 < 3) {
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = hp[Q.comp0 * Q.HS + j * jstep];
  } else {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int zr = t + R * j - Q.z0s;
      v[j] = (zr >= 0 && zr < Q.nzs) ? sp[j * jstep] : make_float2(0.f, 0.f);
    }
  }
  const float axy = P.ax2[kx] + P.ay2[ky];
  float kap[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) kap[k1] = P.az2[t + R * k1];
  const float4* tw = S::load_tw(smraw, 2, Q.tw4z);
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) kap[k1] = kappa_sel<POLY>(axy + kap[k1]) * Q.norm;
#pragma unroll 1
  for (int comp = Q.comp0; comp < ncomp; ++comp) {
    // prefetch the next component
    if (comp + 1 < 3) {
      const float2* np = hp + (comp + 1) * Q.HS;
#pragma unroll
      for (int j = 0; j < R; ++j) nx[j] = np[j * jstep];
    } else if (comp + 1 < ncomp) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int zr = t + R * j - Q.z0s;
        nx[j] = (zr >= 0 && zr < Q.nzs) ? sp[j * jstep] : make_float2(0.f, 0.f);
      }
    }
    strided_fft<R, false>(v, tw, xa, l, t);
    if (comp < 2) {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) v[k1] = cscale(v[k1], kap[k1]);
    } else if (comp == 2) {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) v[k1] = cmul4(cscale(v[k1], kap[k1]), Q.dnz4[t + R * k1]);
    } else {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) v[k1] = cscale(v[k1], cosk_sel<POLY>(axy + P.az2[t + R * k1]) * Q.norm);
    }
    strided_fft<R, true>(v, tw, xb, l, t);
    float2* op = hp + comp * Q.HS;
#pragma unroll
    for (int j = 0; j < R; ++j) op[j * jstep] = v[j];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = nx[j];
  }
}


// z pass of the absorption operands (in place): H4[0] <- IFFT_z[k^(y-2) FFT_z H4[0]], H4[1] <- IFFT_z[k^(y-1) FFT_z H4[1]]
// (the fractional Laplacians of the power-law absorption / dispersion terms).  grid (nxt+1, Ny)
template <int R>
__global__ void __launch_bounds__(16 * R, R == 16 ? 2 : 4) k2_z_absorb(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  int kx, ky;
  if (!lane_map(Q, Q.Ny, l, kx, ky)) return;
  float2* xa = reinterpret_cast<float2*>(smraw);
  float2* xb = reinterpret_cast<float2*>(smraw + S::XCH);
  const int zs = Q.Ny * Q.PH, jstep = R * zs;
  float2* hp = Q.H4 + (long long)t * zs + ky * Q.PH + kx;
  float2 v[R], nx[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = hp[j * jstep];
#pragma unroll
  for (int j = 0; j < R; ++j) nx[j] = hp[Q.HS + j * jstep];
  const float kxy = P.kx2[kx] + P.ky2[ky];
  const float4* tw = S::load_tw(smraw, 2, Q.tw4z);
#pragma unroll 1
  for (int comp = 0; comp < 2; ++comp) {
    strided_fft<R, false>(v, tw, xa, l, t);
    const float e = comp == 0 ? P.y_minus2_half : P.y_minus1_half;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      const float k2 = kxy + P.kz2[t + R * k1];
      const float m = k2 > 0.f ? __powf(k2, e) * Q.norm : 0.f;
      v[k1] = cscale(v[k1], m);
    }
    strided_fft<R, true>(v, tw, xb, l, t);
    float2* op = hp + comp * Q.HS;
#pragma unroll
    for (int j = 0; j < R; ++j) op[j * jstep] = v[j];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = nx[j];
  }
}

// y inverse + row-pair merge of H4[comp] -> Z4[comp].  grid (nxt+1, Nz, ncomp)
template <int R>
__global__ void __launch_bounds__(16 * R, R == 16 ? 3 : 6) k2_y_inv(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = Strided<R>;
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  int comp, by;
  comp_plane(Q, comp, by);
  comp += Q.comp0;
  int kx, z;
  if (!lane_map(Q, Q.Nz, l, kx, z, by)) return;
  float2* xa = reinterpret_cast<float2*>(smraw);
  const float2* hp = Q.H4 + comp * Q.HS + ((long long)z * Q.Ny + t) * Q.PH + kx;
  const int kstep = R * Q.PH;
  float2 a[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) a[k1] = hp[k1 * kstep];
  const float4* tw = S::load_tw(smraw, 1, Q.tw4y);
  strided_fft<R, true>(a, tw, xa, l, t);
  merge_store<R>(Q.Z4 + comp * Q.ZS + ((long long)z * (Q.Ny / 2) + t) * Q.Nx, a, kx, Q.Nx, true);
}

// ------------------------------------------------------------------------------------------------
// x passes.  Persistent CTAs of 128 threads = 128/R groups of R lanes; a group owns one row pair at a
// time and walks its work as a sequence of "items" (one packed spectrum line + the pair of real rows,
// or the pair of sensor rows).  Items are prefetched two deep with cp.async into the group's private
// shared-memory stages, so the memory latency of item q+1 overlaps the transforms of item q regardless
// of occupancy.  The FFT exchange runs inside the (already consumed) spectrum stage with an XOR swizzle
// instead of padding.
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPENDING) : "memory"); }

template <int R, bool INV>
__device__ __forceinline__ void line_fft_sw(float2 (&v)[R], const float4* __restrict__ tw, float2* sm, int t) {
  constexpr int N = R * R;
  dft<R, INV>(v);
  if constexpr (!INV) {
#pragma unroll
    for (int k2 = 1; k2 < R; ++k2) v[k2] = cmul4(v[k2], tw[t * k2]);
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) sm[t * R + (k2 ^ t)] = v[k2];
    __syncwarp();
#pragma unroll
    for (int tt = 0; tt < R; ++tt) v[tt] = sm[tt * R + (t ^ tt)];
    __syncwarp();
  } else {
#pragma unroll
    for (int tt = 1; tt < R; ++tt) v[tt] = cmul4(v[tt], tw[N - tt * t]);
#pragma unroll
    for (int tt = 0; tt < R; ++tt) sm[tt * R + (t ^ tt)] = v[tt];
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) v[k2] = sm[t * R + (k2 ^ t)];
    __syncwarp();
  }
  dft<R, INV>(v);
}

// Running maximum / minimum of one sensor voxel.  The pair is written back only when it changed: away from the passing
// burst neither extreme moves, so most of the 8 bytes per sensor voxel and step of write traffic never reaches DRAM
// (the stored values are the same either way).
__device__ __forceinline__ void sensor_update(float2* dst, float2 old, float p, int always) {
  const float2 nv = make_float2(fmaxf(old.x, p), fminf(old.y, p));
  if (always || nv.x != old.x || nv.y != old.y) *dst = nv;
}

// One group's staging: a stage is SB*N bytes = [ N float2 spectrum line | row lo (N floats) | row hi (N floats) ]
// or [ sensor row lo (N float2) | sensor row hi (N float2) ]; heterogeneous media append the two rows of the medium
// map the item needs (SB = 24) so that they arrive with the fields instead of being loaded after the transform.
// XB*N extra bytes per group sit behind the ring (double-buffered per-pair medium rows of k2_x_rho_p).
template <int R, int SB = 16, int XB = 0>
struct XStage {
  static constexpr int N = R * R;
  static constexpr int BYTES = SB * N;          // per stage
  static constexpr int GROUPS = 128 / R;
  static constexpr int GBYTES = 2 * BYTES + XB * N;                 // per group
  static constexpr int SMEM = GROUPS * GBYTES + 16 * (N + 1);       // + twiddle table
  // copy `bytes` (multiple of 16*R) from global to shared with the R lanes of the group
  static __device__ __forceinline__ void copy(char* sdst, const char* gsrc, int bytes, int t) {
#pragma unroll
    for (int o = 0; o < 8 * N; o += 16 * R)
      if (o < bytes) cp_async16(sdst + o + 16 * t, gsrc + o + 16 * t);
  }
  static __device__ __forceinline__ const float4* load_tw(unsigned char* smraw, const float4* __restrict__ g) {
    float4* s = reinterpret_cast<float4*>(smraw + GROUPS * GBYTES);
    for (int i = threadIdx.x; i <= N; i += 128) s[i] = g[i];
    __syncthreads();
    return s;
  }
};
template <int R, bool HOMOG> using XStageU = XStage<R, HOMOG ? 16 : 24, 0>;
template <int R, bool HOMOG> using XStageRho = XStage<R, 16, HOMOG ? 0 : 16>;

// Traversal order of the persistent x kernels.  With Q.xrev the batches are walked from the LAST row pair down: the
// kernel then starts on the part of its input that the preceding y pass wrote last (still in the 126 MB L2) and ends on
// the low-z planes, which the following y pass reads first.
__device__ __forceinline__ int phys_pair(const V2Params& Q, int pair) {
  return Q.xrev ? Q.Nz * (Q.Ny >> 1) - 1 - pair : pair;
}

// row pair index (z*Ny/2 + m) -> z and the lower row y_lo of the pair (Ny/2 and Ry are powers of two)
__device__ __forceinline__ void pair_rows(const V2Params& Q, int pair, int& z, int& ylo) {
  const int m = pair & ((Q.Ny >> 1) - 1);
  z = pair >> Q.hy_sh;
  ylo = ((m >> Q.ry_sh) << (Q.ry_sh + 1)) | (m & (Q.Ry - 1));
}
__device__ __forceinline__ long long pair_row_lo(const V2Params& Q, int z, int m) {
  const int ylo = ((m >> Q.ry_sh) << (Q.ry_sh + 1)) | (m & (Q.Ry - 1));
  return ((long long)z * Q.Ny + ylo) * Q.Nx;
}

// Work of one persistent CTA: row-pair batches blockIdx.x, blockIdx.x + gridDim.x, ...; inside a batch the
// items of a pair are unrolled at compile time, so the per-item index arithmetic folds away.
template <int R, bool HOMOG>
__global__ void __launch_bounds__(128, HOMOG ? 3 : 2) k2_x_u(StepParams P, V2Params Q) {
  using XS = XStageU<R, HOMOG>;
  constexpr int N = R * R, G = XS::GROUPS;
  extern __shared__ __align__(16) unsigned char smraw[];
  const float4* tw = XS::load_tw(smraw, Q.tw4x);
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  char* gbase = reinterpret_cast<char*>(smraw) + g * XS::GBYTES;
  const long long hi = (long long)Q.Ry * N;                       // offset of the pair's upper row
  const int nbatch = Q.Nz * (Q.Ny / 2) / G;
  const int iters = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int pstep = gridDim.x * G;

  auto issue = [&](int lpair, int c, int stage, bool valid) {
    if (valid) {
      const int pair = phys_pair(Q, lpair);
      int z, ylo;
      pair_rows(Q, pair, z, ylo);
      const long long r0 = ((long long)z * Q.Ny + ylo) * N;
      char* st = gbase + stage * XS::BYTES;
      XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + c * Q.ZS + (long long)pair * N), 8 * N, t);
      XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.u + c * P.RS + r0), 4 * N, t);
      XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.u + c * P.RS + r0 + hi), 4 * N, t);
      if constexpr (!HOMOG) {
        XS::copy(st + 16 * N, reinterpret_cast<const char*>(P.dt_rho0_sg + c * P.RS + r0), 4 * N, t);
        XS::copy(st + 20 * N, reinterpret_cast<const char*>(P.dt_rho0_sg + c * P.RS + r0 + hi), 4 * N, t);
      }
    }
    cp_async_commit();
  };
  int lpair = blockIdx.x * G + g;
  issue(lpair, 0, 0, iters > 0);
  issue(lpair, 1, 1, iters > 0);
  for (int it = 0; it < iters; ++it, lpair += pstep) {
    const int pair = phys_pair(Q, lpair);
    int z, ylo;
    pair_rows(Q, pair, z, ylo);
    const long long r0 = ((long long)z * Q.Ny + ylo) * N;
    const int par = it & 1;                               // item index = 3*it + c
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cp_async_wait<1>();
      __syncwarp();
      const int stage = (par + c) & 1;
      float2* zb = reinterpret_cast<float2*>(gbase + stage * XS::BYTES);
      const float* rb = reinterpret_cast<const float*>(gbase + stage * XS::BYTES + 8 * N);
      float2 v[R];
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = zb[t + R * j];
      __syncwarp();
      line_fft_sw<R, true>(v, tw, zb, t);
      float* u = P.u + c * P.RS + r0;
      float2 s;
      if (c == 1) s = make_float2(P.sgy[ylo], P.sgy[ylo + Q.Ry]);
      else if (c == 2) s.x = s.y = P.sgz[z];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int x = t + R * j;
        if (c == 0) s.x = s.y = P.sgx[x];
        float2 d;
        if constexpr (HOMOG) { d.x = d.y = -P.dt_rho0_sg_s; }
        else { d.x = -rb[2 * N + x]; d.y = -rb[3 * N + x]; }
        // u = s (s u - dt/rho0 dp)
        const float2 un = __fmul2_rn(s, __ffma2_rn(d, v[j], __fmul2_rn(s, make_float2(rb[x], rb[N + x]))));
        u[x] = un.x;
        u[hi + x] = un.y;
        v[j] = un;
      }
      line_fft_sw<R, false>(v, tw, zb, t);
      float2* zo = Q.Z4 + c * Q.ZS + (long long)pair * N;
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = v[k1];
      __syncwarp();
      if (c == 0) issue(lpair, 2, stage, true);
      else issue(lpair + pstep, c - 1, stage, it + 1 < iters);
    }
  }
  cp_async_wait<0>();
}

// SRC: 0 none, 1 filtered source in Z4[3], 2 unfiltered dense slab, 3 steady-state source q1 F1 + q2 F2 (real fields Q.FK)
// Items per row pair: [source spectrum], rho_x, rho_y, rho_z, sensor rows (pm = interleaved (p_max, p_min)
// on the expanded grid; rows in the PML are skipped).
// ABS (absorbing medium): no equation of state here.  The kernel forms the operands of the two fractional
// Laplacians, rho0 * sum_xi d_xi u_xi and sum_xi rho_xi, writes their packed x-spectra to Z4[0] / Z4[1] (the lines
// this pair has already consumed) and keeps sum rho in r1; k2_x_p finishes the step.
template <int R, bool HOMOG, int SRC, bool ABS = false>
__global__ void __launch_bounds__(128, HOMOG ? 3 : 2) k2_x_rho_p(StepParams P, V2Params Q) {
  using XS = XStageRho<R, HOMOG>;
  constexpr int N = R * R, G = XS::GROUPS;
  constexpr int NI = ((SRC == 1 || SRC == 3) ? 5 : 4) - (ABS ? 1 : 0);
  constexpr int C0 = (SRC == 1 || SRC == 3) ? 1 : 0;    // item index of rho_x
  extern __shared__ __align__(16) unsigned char smraw[];
  const float4* tw = XS::load_tw(smraw, Q.tw4x);
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  char* gbase = reinterpret_cast<char*>(smraw) + g * XS::GBYTES;
  char* mbase = gbase + 2 * XS::BYTES;                  // heterogeneous: 2 x [dt*rho0 row lo | row hi], one slot per pair
  const long long hi = (long long)Q.Ry * N;
  const int nbatch = Q.Nz * (Q.Ny / 2) / G;
  const int iters = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int pstep = gridDim.x * G;

  // c: -1 source item, 0..2 rho_x, rho_y, rho_z, 3 sensor rows; slot: parity of the pair's iteration
  auto issue = [&](int lpair, int c, int stage, bool valid, int slot) {
    if (valid) {
      const int pair = phys_pair(Q, lpair);
      int z, ylo;
      pair_rows(Q, pair, z, ylo);
      const long long r0 = ((long long)z * Q.Ny + ylo) * N;
      char* st = gbase + stage * XS::BYTES;
      if (SRC == 3 && c < 0) {
        // source item of the steady window: rows lo / hi of the two filtered basis fields
        XS::copy(st, reinterpret_cast<const char*>(Q.FK + r0), 4 * N, t);
        XS::copy(st + 4 * N, reinterpret_cast<const char*>(Q.FK + r0 + hi), 4 * N, t);
        XS::copy(st + 8 * N, reinterpret_cast<const char*>(Q.FK + P.RS + r0), 4 * N, t);
        XS::copy(st + 12 * N, reinterpret_cast<const char*>(Q.FK + P.RS + r0 + hi), 4 * N, t);
      } else if (c < 3) {
        XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + (c < 0 ? 3 : c) * Q.ZS + (long long)pair * N), 8 * N, t);
        if (c >= 0) {
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.rho + c * P.RS + r0), 4 * N, t);
          XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.rho + c * P.RS + r0 + hi), 4 * N, t);
        }
        if constexpr (!HOMOG) {
          if (c == 0) {     // the pair's dt*rho0 rows ride with its first density item and serve all three
            XS::copy(mbase + slot * 8 * N, reinterpret_cast<const char*>(P.dt_rho0 + r0), 4 * N, t);
            XS::copy(mbase + slot * 8 * N + 4 * N, reinterpret_cast<const char*>(P.dt_rho0 + r0 + hi), 4 * N, t);
          }
        }
      } else {
        const bool zin = (unsigned)(z - P.pz) < (unsigned)P.nz;
        if (zin && (unsigned)(ylo - P.py) < (unsigned)P.ny)
          XS::copy(st, reinterpret_cast<const char*>(Q.pm + r0), 8 * N, t);
        if (zin && (unsigned)(ylo + Q.Ry - P.py) < (unsigned)P.ny)
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(Q.pm + r0 + hi), 8 * N, t);
      }
    }
    cp_async_commit();
  };
  int lpair = blockIdx.x * G + g;
  issue(lpair, 0 - C0, 0, iters > 0, 0);
  issue(lpair, 1 - C0, 1, iters > 0, 0);
  float2 src[R], sum[R];
  float2 dsum[ABS ? R : 1];
  float2 q1 = make_float2(0.f, 0.f), q2 = q1;
  if (SRC == 3) { q1.x = q1.y = Q.qcur[0]; q2.x = q2.y = Q.qcur[1]; }
  for (int it = 0; it < iters; ++it, lpair += pstep) {
    const int pair = phys_pair(Q, lpair);
    int z, ylo;
    pair_rows(Q, pair, z, ylo);
    const long long r0 = ((long long)z * Q.Ny + ylo) * N;
    const int par = (it * NI) & 1;
#pragma unroll
    for (int ci = 0; ci < NI; ++ci) {
      const int c = ci - C0;
      cp_async_wait<1>();
      __syncwarp();
      const int stage = (par + ci) & 1;
      float2* zb = reinterpret_cast<float2*>(gbase + stage * XS::BYTES);
      const float* rb = reinterpret_cast<const float*>(gbase + stage * XS::BYTES + 8 * N);
      const float* mb = reinterpret_cast<const float*>(mbase + (it & 1) * 8 * N);
      if (SRC == 3 && c < 0) {
        const float* fb = reinterpret_cast<const float*>(zb);     // [F1 lo | F1 hi | F2 lo | F2 hi]
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int x = t + R * j;
          src[j] = __ffma2_rn(q2, make_float2(fb[2 * N + x], fb[3 * N + x]), __fmul2_rn(q1, make_float2(fb[x], fb[N + x])));
        }
      } else if (c < 3) {
        float2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = zb[t + R * j];
        __syncwarp();
        line_fft_sw<R, true>(v, tw, zb, t);
        if (c < 0) {
#pragma unroll
          for (int j = 0; j < R; ++j) src[j] = v[j];
        } else {
          if (SRC == 2 && c == 0) {
            const int zr = z - Q.z0s;
            const bool in = zr >= 0 && zr < Q.nzs;
            const long long so = ((long long)zr * Q.Ny + ylo) * N;
#pragma unroll
            for (int j = 0; j < R; ++j)
              src[j] = in ? make_float2(Q.Sslab[so + t + R * j], Q.Sslab[so + hi + t + R * j]) : make_float2(0.f, 0.f);
          }
          float* rho = P.rho + c * P.RS + r0;
          float2 a;
          if (c == 1) a = make_float2(P.pmly[ylo], P.pmly[ylo + Q.Ry]);
          else if (c == 2) a.x = a.y = P.pmlz[z];
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int x = t + R * j;
            if (c == 0) a.x = a.y = P.pmlx[x];
            float2 d;
            if constexpr (HOMOG) { d.x = d.y = -P.dt_rho0_s; }
            else { d.x = -mb[x]; d.y = -mb[N + x]; }
            // rho = a (a rho - dt rho0 du) [+ S]
            float2 rn = __fmul2_rn(a, __ffma2_rn(d, v[j], __fmul2_rn(a, make_float2(rb[x], rb[N + x]))));
            if constexpr (SRC != 0) rn = cadd(rn, src[j]);
            rho[x] = rn.x;
            rho[hi + x] = rn.y;
            sum[j] = c == 0 ? rn : cadd(sum[j], rn);           // (rho_x + rho_y) + rho_z
            if constexpr (ABS) dsum[j] = c == 0 ? v[j] : cadd(dsum[j], v[j]);   // (dux + duy) + duz
          }
          if (ABS && c == 2) {
            if constexpr (ABS) {
#pragma unroll
              for (int j = 0; j < R; ++j) {
                const int x = t + R * j;
                float2 r0v;      // rho0 = (dt rho0) / dt: the staged row serves the tau operand too
                if constexpr (HOMOG) { r0v.x = r0v.y = P.rho0_s; } else { r0v.x = mb[x] * P.inv_dt; r0v.y = mb[N + x] * P.inv_dt; }
                dsum[j] = __fmul2_rn(r0v, dsum[j]);
                P.r1[r0 + x] = sum[j].x;
                P.r1[r0 + hi + x] = sum[j].y;
              }
              __syncwarp();
              line_fft_sw<R, false>(dsum, tw, zb, t);
              float2* zo = Q.Z4 + (long long)pair * N;
#pragma unroll
              for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = dsum[k1];
              __syncwarp();
              line_fft_sw<R, false>(sum, tw, zb, t);
              zo += Q.ZS;
#pragma unroll
              for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = sum[k1];
            }
          } else if (c == 2) {
            // equation of state; the sensor item that follows consumes p from `sum`
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const int x = t + R * j;
              float2 c2;
              if constexpr (HOMOG) { c2.x = c2.y = P.c2_s; } else { c2.x = P.c2[r0 + x]; c2.y = P.c2[r0 + hi + x]; }
              sum[j] = __fmul2_rn(c2, sum[j]);
              if (Q.store_p) { P.p[r0 + x] = sum[j].x; P.p[r0 + hi + x] = sum[j].y; }
            }
          }
        }
      } else {
        // running max / min on the two sensor rows, then the forward transform of the new pressure
        const bool zin = (unsigned)(z - P.pz) < (unsigned)P.nz;
        const bool in0 = zin && (unsigned)(ylo - P.py) < (unsigned)P.ny;
        const bool in1 = zin && (unsigned)(ylo + Q.Ry - P.py) < (unsigned)P.ny;
        float2* pmg = Q.pm + r0;
        if (in0) {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int x = t + R * j;
            sensor_update(pmg + x, zb[x], sum[j].x, Q.pm_always);
          }
        }
        if (in1) {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int x = t + R * j;
            sensor_update(pmg + hi + x, zb[N + x], sum[j].y, Q.pm_always);
          }
        }
        __syncwarp();
        line_fft_sw<R, false>(sum, tw, zb, t);
        float2* zo = Q.ZP + (long long)pair * N;
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = sum[k1];
      }
      __syncwarp();
      if (ci + 2 < NI) issue(lpair, ci + 2 - C0, stage, true, it & 1);
      else issue(lpair + pstep, ci + 2 - NI - C0, stage, it + 1 < iters, (it + 1) & 1);
    }
  }
  cp_async_wait<0>();
  if (blockIdx.x == 0 && threadIdx.x == 0) *P.step = *P.step + 1;
}


// Absorbing medium, last pass of the step.  Items per row pair: [Z4[0] line + the r1 rows (sum rho)],
// [Z4[1] line], [sensor rows]:  p = c0^2 (sum rho + tau L1 - eta L2), running max/min, FFT_x of p -> ZP.
template <int R, bool HOMOG>
__global__ void __launch_bounds__(128, HOMOG ? 3 : 2) k2_x_p(StepParams P, V2Params Q, int use_tau, int use_eta) {
  using XS = XStageU<R, HOMOG>;
  constexpr int N = R * R, G = XS::GROUPS;
  constexpr int NI = 3;
  extern __shared__ __align__(16) unsigned char smraw[];
  const float4* tw = XS::load_tw(smraw, Q.tw4x);
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  char* gbase = reinterpret_cast<char*>(smraw) + g * XS::GBYTES;
  const long long hi = (long long)Q.Ry * N;
  const int nbatch = Q.Nz * (Q.Ny / 2) / G;
  const int iters = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int pstep = gridDim.x * G;
  auto issue = [&](int lpair, int c, int stage, bool valid) {
    if (valid) {
      const int pair = phys_pair(Q, lpair);
      int z, ylo;
      pair_rows(Q, pair, z, ylo);
      const long long r0 = ((long long)z * Q.Ny + ylo) * N;
      char* st = gbase + stage * XS::BYTES;
      if (c < 2) {
        XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + c * Q.ZS + (long long)pair * N), 8 * N, t);
        if (c == 0) {
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.r1 + r0), 4 * N, t);
          XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.r1 + r0 + hi), 4 * N, t);
          if constexpr (!HOMOG) {
            XS::copy(st + 16 * N, reinterpret_cast<const char*>(P.tau + r0), 4 * N, t);
            XS::copy(st + 20 * N, reinterpret_cast<const char*>(P.tau + r0 + hi), 4 * N, t);
          }
        } else if constexpr (!HOMOG) {
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.eta + r0), 4 * N, t);
          XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.eta + r0 + hi), 4 * N, t);
          XS::copy(st + 16 * N, reinterpret_cast<const char*>(P.c2 + r0), 4 * N, t);
          XS::copy(st + 20 * N, reinterpret_cast<const char*>(P.c2 + r0 + hi), 4 * N, t);
        }
      } else {
        const bool zin = (unsigned)(z - P.pz) < (unsigned)P.nz;
        if (zin && (unsigned)(ylo - P.py) < (unsigned)P.ny)
          XS::copy(st, reinterpret_cast<const char*>(Q.pm + r0), 8 * N, t);
        if (zin && (unsigned)(ylo + Q.Ry - P.py) < (unsigned)P.ny)
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(Q.pm + r0 + hi), 8 * N, t);
      }
    }
    cp_async_commit();
  };
  int lpair = blockIdx.x * G + g;
  issue(lpair, 0, 0, iters > 0);
  issue(lpair, 1, 1, iters > 0);
  float2 acc[R];
  for (int it = 0; it < iters; ++it, lpair += pstep) {
    const int pair = phys_pair(Q, lpair);
    int z, ylo;
    pair_rows(Q, pair, z, ylo);
    const long long r0 = ((long long)z * Q.Ny + ylo) * N;
    const int par = (it * NI) & 1;
#pragma unroll
    for (int c = 0; c < NI; ++c) {
      cp_async_wait<1>();
      __syncwarp();
      const int stage = (par + c) & 1;
      float2* zb = reinterpret_cast<float2*>(gbase + stage * XS::BYTES);
      const float* rb = reinterpret_cast<const float*>(gbase + stage * XS::BYTES + 8 * N);
      if (c < 2) {
        float2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = zb[t + R * j];
        __syncwarp();
        line_fft_sw<R, true>(v, tw, zb, t);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int x = t + R * j;
          if (c == 0) {
            float2 ta;
            if constexpr (HOMOG) { ta.x = ta.y = P.tau_s; } else { ta.x = rb[2 * N + x]; ta.y = rb[3 * N + x]; }
            const float2 s0 = make_float2(rb[x], rb[N + x]);
            acc[j] = use_tau ? __ffma2_rn(ta, v[j], s0) : s0;              // sum rho + tau L1
          } else {
            float2 et, c2;
            if constexpr (HOMOG) { et.x = et.y = -P.eta_s; c2.x = c2.y = P.c2_s; }
            else { et.x = -rb[x]; et.y = -rb[N + x]; c2.x = rb[2 * N + x]; c2.y = rb[3 * N + x]; }
            if (use_eta) acc[j] = __ffma2_rn(et, v[j], acc[j]);            // ... - eta L2
            acc[j] = __fmul2_rn(c2, acc[j]);
            if (Q.store_p) { P.p[r0 + x] = acc[j].x; P.p[r0 + hi + x] = acc[j].y; }
          }
        }
      } else {
        const bool zin = (unsigned)(z - P.pz) < (unsigned)P.nz;
        const bool in0 = zin && (unsigned)(ylo - P.py) < (unsigned)P.ny;
        const bool in1 = zin && (unsigned)(ylo + Q.Ry - P.py) < (unsigned)P.ny;
        float2* pmg = Q.pm + r0;
        if (in0) {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int x = t + R * j;
            sensor_update(pmg + x, zb[x], acc[j].x, Q.pm_always);
          }
        }
        if (in1) {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const int x = t + R * j;
            sensor_update(pmg + hi + x, zb[N + x], acc[j].y, Q.pm_always);
          }
        }
        __syncwarp();
        line_fft_sw<R, false>(acc, tw, zb, t);
        float2* zo = Q.ZP + (long long)pair * N;
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = acc[k1];
      }
      __syncwarp();
      if (c + 2 < NI) issue(lpair, c + 2, stage, true);
      else issue(lpair + pstep, c + 2 - NI, stage, it + 1 < iters);
    }
  }
  cp_async_wait<0>();
}

// x forward of the dense source slab (row pairs).  grid = ceil(nzs*(Ny/2)/G), G = 256/R
template <int R>
__global__ void __launch_bounds__(256) k2_x_src(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int G = 256 / R, N = R * R;
  float4* tws = reinterpret_cast<float4*>(smraw + G * N * 8);
  for (int i = threadIdx.x; i <= N; i += 256) tws[i] = Q.tw4x[i];
  __syncthreads();
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  float2* sm = reinterpret_cast<float2*>(smraw) + g * N;
  const int hy = Q.Ny / 2;
  const long long pair = (long long)blockIdx.x * G + g;          // zr*(Ny/2) + m
  if (pair >= (long long)Q.nzs * hy) return;
  const int m = (int)(pair & (hy - 1)), zr = (int)(pair >> Q.hy_sh);
  const long long r0 = pair_row_lo(Q, zr, m), hi = (long long)Q.Ry * N;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = make_float2(Q.Sslab[r0 + t + R * j], Q.Sslab[r0 + hi + t + R * j]);
  line_fft_sw<R, false>(v, tws, sm, t);
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) Q.ZSslab[pair * N + t + R * k1] = v[k1];
}

// Source scatter into the dense slab (same arithmetic as k_source_scatter of v1).
static __global__ void __launch_bounds__(128) k2_source_scatter(StepParams P, V2Params Q, SourceParams S) {
  const int t = *P.step;
  const long long slab0 = (long long)Q.z0s * P.Ny * P.Nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.n_src;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = S.row_ptr[i]; j < S.row_ptr[i + 1]; ++j) {
      int e = S.col[j];
      int tt = t - S.delay[e];
      if (tt >= 0 && tt < S.n_base) acc = fmaf(S.w[j] * S.gain[e], S.base[tt], acc);
    }
    Q.Sslab[S.lin_exp[i] - slab0] = acc * S.scale[i];
  }
}

// Steady-state source, set-up: dense slab of one spatial basis field, sum_e W[i,e] coef_e (coef_e = gain_e * c_{k,e}).
static __global__ void __launch_bounds__(128) k2_source_basis(StepParams P, V2Params Q, SourceParams S, const float* __restrict__ coef) {
  const long long slab0 = (long long)Q.z0s * P.Ny * P.Nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.n_src;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = S.row_ptr[i]; j < S.row_ptr[i + 1]; ++j) acc = fmaf(S.w[j], coef[S.col[j]], acc);
    Q.Sslab[S.lin_exp[i] - slab0] = acc * S.scale[i];
  }
}

// x inverse of one packed spectrum field back to a real field (row pairs -> two rows).  grid = Nz*(Ny/2)/G, G = 256/R
template <int R>
__global__ void __launch_bounds__(256) k2_x_inv_real(StepParams P, V2Params Q, const float2* __restrict__ Zin, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int G = 256 / R, N = R * R;
  float4* tws = reinterpret_cast<float4*>(smraw + G * N * 8);
  for (int i = threadIdx.x; i <= N; i += 256) tws[i] = Q.tw4x[i];
  __syncthreads();
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  float2* sm = reinterpret_cast<float2*>(smraw) + g * N;
  const int pair = blockIdx.x * G + g;
  int z, ylo;
  pair_rows(Q, pair, z, ylo);
  const long long r0 = ((long long)z * Q.Ny + ylo) * N, hi = (long long)Q.Ry * N;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = Zin[(long long)pair * N + t + R * j];
  line_fft_sw<R, true>(v, tws, sm, t);
#pragma unroll
  for (int j = 0; j < R; ++j) { out[r0 + t + R * j] = v[j].x; out[r0 + hi + t + R * j] = v[j].y; }
}

// sensor field helpers: pm = interleaved (p_max, p_min) on the expanded grid
static __global__ void k2_pm_init(float2* pm, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pm[i] = make_float2(-INFINITY, INFINITY);
}
static __global__ void k2_pm_crop(StepParams P, V2Params Q) {
  const long long n = (long long)P.nx * P.ny * P.nz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % P.nx);
    const long long r = i / P.nx;
    const int y = (int)(r % P.ny), z = (int)(r / P.ny);
    const float2 v = Q.pm[((long long)(z + P.pz) * P.Ny + (y + P.py)) * P.Nx + (x + P.px)];
    P.pmax[i] = v.x;
    P.pmin[i] = v.y;
  }
}

}  // namespace lifu
