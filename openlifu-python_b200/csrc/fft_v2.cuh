// fft_v2.cuh -- pipeline v2: hand-written FFT passes with every spectral / real-space operation of
// the k-space time step fused into them (no library transforms, no separate element-wise kernels).
//
// What this replaces: the 10-12 cuFFT 3-D transforms + 5 element-wise kernels of pipeline v1, i.e. the
// per-step body of kspaceFirstOrder3D reached from /root/reference/src/openlifu/sim/kwave_if.py:124-129.
//
// Transform structure.  A length-N line (N = R*R, R in {8, 16}) is transformed by R threads:
// thread t holds x[t + R*j] (j = 0..R-1) in registers, does a register DFT of size R over j, multiplies
// the inter-stage twiddle w_N^(t*k2), exchanges through shared memory, and does a second register DFT
// over t.  Outputs sit as X[k2 + R*k1] in (thread k2, register k1) -- the same "thread + R*register"
// pattern as the input, so a forward transform can be followed by its inverse without any reordering.
//
// Layouts (float2 = complex):
//   real field   R[z][y][x]                                       (x fastest)
//   Z layout     Z[z][m][kx], m = y/2, kx = 0..Nx-1               x-spectrum of the ROW PAIR
//                (row 2m) + i*(row 2m+1): two real lines share one complex transform
//   H layout     H[z][ky][kx], kx = 0..Nx/2, row pitch PH         half spectrum
// The x passes therefore are plain complex transforms; the y passes split (on load) or merge (on
// store) the packed row pairs with the Hermitian pairing kx <-> Nx-kx.
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
// complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

template <int N, bool INV>
__device__ __forceinline__ float2 tw_const(int m) {   // exp(-+2 pi i m / N), N | 32
  const int k = (m * (32 / N)) & 31;
  return make_float2(w32_re(k), INV ? -w32_im(k) : w32_im(k));
}

// Register DFT of size N (2, 4, 8, 16, 32), natural order in and out, unnormalised.
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
    if constexpr (!INV) {          // X1 = d0 - i d1, X3 = d0 + i d1
      x[1] = make_float2(d0.x + d1.y, d0.y - d1.x);
      x[3] = make_float2(d0.x - d1.y, d0.y + d1.x);
    } else {
      x[1] = make_float2(d0.x - d1.y, d0.y + d1.x);
      x[3] = make_float2(d0.x + d1.y, d0.y - d1.x);
    }
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
      for (int ka = 0; ka < A; ++ka) y[b][ka] = (b * ka == 0) ? s[ka] : cmul2(s[ka], tw_const<N, INV>(b * ka));
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
// Strided-axis transforms (y and z passes).  CTA = 16 lanes (consecutive kx) x R threads per line.
// Thread (l, t): l = tid & 15, t = tid >> 4.  Exchange buffer: R*R*16 float2.
template <int R, bool INV>
__device__ __forceinline__ void strided_fft(float2 (&v)[R], const float2* __restrict__ tw, float2* sm, int l, int t) {
  constexpr int N = R * R;
  if constexpr (!INV) {
    dft<R, false>(v);                                   // over j -> k2
#pragma unroll
    for (int k2 = 1; k2 < R; ++k2) v[k2] = cmul2(v[k2], tw[(t * k2) & (N - 1)]);
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) sm[(t * R + k2) * 16 + l] = v[k2];
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < R; ++tt) v[tt] = sm[(tt * R + t) * 16 + l];   // this thread now owns k2 = t
    __syncthreads();
    dft<R, false>(v);                                   // over t -> k1 ; X[t + R*k1] = v[k1]
  } else {
    dft<R, true>(v);                                    // over k1 -> tt, for k2 = t
#pragma unroll
    for (int tt = 1; tt < R; ++tt) v[tt] = cmulc(v[tt], tw[(tt * t) & (N - 1)]);
#pragma unroll
    for (int tt = 0; tt < R; ++tt) sm[(tt * R + t) * 16 + l] = v[tt];
    __syncthreads();
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) v[k2] = sm[(t * R + k2) * 16 + l];
    __syncthreads();
    dft<R, true>(v);                                    // over k2 -> j ; x[t + R*j] = v[j]
  }
}

// Contiguous-axis transforms (x passes).  R consecutive lanes of a warp own one line; exchange region of
// R*(R+1) float2 per line, warp-synchronous.
template <int R, bool INV>
__device__ __forceinline__ void line_fft(float2 (&v)[R], const float2* __restrict__ tw, float2* sm, int t) {
  constexpr int N = R * R;
  constexpr int P = R + 1;
  if constexpr (!INV) {
    dft<R, false>(v);
#pragma unroll
    for (int k2 = 1; k2 < R; ++k2) v[k2] = cmul2(v[k2], tw[(t * k2) & (N - 1)]);
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) sm[t * P + k2] = v[k2];
    __syncwarp();
#pragma unroll
    for (int tt = 0; tt < R; ++tt) v[tt] = sm[tt * P + t];
    __syncwarp();
    dft<R, false>(v);
  } else {
    dft<R, true>(v);
#pragma unroll
    for (int tt = 1; tt < R; ++tt) v[tt] = cmulc(v[tt], tw[(tt * t) & (N - 1)]);
#pragma unroll
    for (int tt = 0; tt < R; ++tt) sm[tt * P + t] = v[tt];
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) v[k2] = sm[t * P + k2];
    __syncwarp();
    dft<R, true>(v);
  }
}

__device__ __forceinline__ float2 shfl_xor16(float2 v) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, 16), __shfl_xor_sync(0xffffffffu, v.y, 16));
}

// Merge the row pair held by two adjacent half-warps (t even: row 2m, t odd: row 2m+1) into the packed
// x-spectrum and store it: Z[kx] = A + iB, Z[Nx-kx] = conj(A) + i conj(B).  Bins 0 and Nx/2 keep real
// parts only, as a C2R transform would.
__device__ __forceinline__ void store_row_pair(float2* __restrict__ zline, float2 own, int t, int kx, int Nx, bool active) {
  float2 other = shfl_xor16(own);
  if (!active) return;
  const bool self_mirror = (kx == 0) || (2 * kx == Nx);
  if ((t & 1) == 0) {
    float2 A = own, B = other;
    zline[kx] = self_mirror ? make_float2(A.x, B.x) : make_float2(A.x - B.y, A.y + B.x);
  } else if (!self_mirror) {
    float2 A = other, B = own;
    zline[Nx - kx] = make_float2(A.x + B.y, B.x - A.y);
  }
}

// ------------------------------------------------------------------------------------------------
// y forward: packed row pairs -> half spectrum.  grid (PH/16, Nz, ncomp)
// MODE 0: pressure (no multipliers)   MODE 1: velocity (comp 0: i kx e^{-i kx dx/2}; comp 1: i ky e^{-i ky dy/2})
// MODE 2: source slab (z index relative to the slab)
template <int R, int MODE>
__global__ void __launch_bounds__(16 * R) k2_y_fwd(StepParams P, V2Params Q) {
  extern __shared__ float2 smem[];
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int kx = blockIdx.x * 16 + l, z = blockIdx.y, comp = blockIdx.z;
  const bool active = kx < Q.Nxh;
  const float2* Zin = MODE == 0 ? Q.ZP : (MODE == 1 ? Q.Z4 + comp * Q.ZS : Q.ZSslab);
  float2* Hout = MODE == 2 ? Q.HSslab : Q.H4 + comp * Q.HS;
  const long long zb = (long long)z * (Q.Ny / 2) * Q.Nx;
  const int km = (Q.Nx - kx) & (Q.Nx - 1);
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 d = make_float2(0.f, 0.f), m = d;
    if (active) {
      const long long row = zb + (long long)((t >> 1) + (R / 2) * j) * Q.Nx;
      d = Zin[row + kx];
      m = Zin[row + km];
    }
    // 2A = Z[k] + conj Z[-k] ; 2B = -i (Z[k] - conj Z[-k])
    v[j] = (t & 1) == 0 ? make_float2(d.x + m.x, d.y - m.y) : make_float2(d.y + m.y, m.x - d.x);
  }
  if (MODE == 1 && comp == 0 && active) {
    const float2 mx = P.dnx[kx];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = cmul2(v[j], mx);
  }
  strided_fft<R, false>(v, Q.twy, smem, l, t);
  if (active) {
    const long long hb = (long long)z * Q.Ny * Q.PH + kx;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      const int ky = t + R * k1;
      float2 o = v[k1];
      if (MODE == 1 && comp == 1) o = cmul2(o, P.dny[ky]);
      Hout[hb + (long long)ky * Q.PH] = o;
    }
  }
}

// z pass of the pressure gradient: H4[0] -> H4[0] (kappa p^) and H4[1] (i kz e^{+i kz dz/2} kappa p^),
// both already inverse transformed along z.  grid (PH/16, Ny)
template <int R, bool POLY>
__global__ void __launch_bounds__(16 * R) k2_z_grad(StepParams P, V2Params Q) {
  extern __shared__ float2 smem[];
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int kx = blockIdx.x * 16 + l, ky = blockIdx.y;
  const bool active = kx < Q.Nxh;
  const long long zs = (long long)Q.Ny * Q.PH;
  const long long base = (long long)ky * Q.PH + kx;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = active ? Q.H4[base + (long long)(t + R * j) * zs] : make_float2(0.f, 0.f);
  strided_fft<R, false>(v, Q.twz, smem, l, t);
  float2 w[R];
  const float axy = active ? P.ax2[kx] + P.ay2[ky] : 0.f;
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) {
    const int kz = t + R * k1;
    const float a2 = axy + P.az2[kz];
    const float kap = (POLY ? sinc_sqrt_poly(a2) : kappa_of(a2)) * Q.norm;
    v[k1] = cscale(v[k1], kap);
    w[k1] = cmul2(v[k1], P.dpz[kz]);
  }
  strided_fft<R, true>(v, Q.twz, smem, l, t);
  if (active) {
#pragma unroll
    for (int j = 0; j < R; ++j) Q.H4[base + (long long)(t + R * j) * zs] = v[j];
  }
  strided_fft<R, true>(w, Q.twz, smem, l, t);
  if (active) {
#pragma unroll
    for (int j = 0; j < R; ++j) Q.H4[Q.HS + base + (long long)(t + R * j) * zs] = w[j];
  }
}

// y inverse of the three gradient components + row-pair merge.  grid (PH/16, Nz)
//   Z4[0] <- i kx e^{+i kx dx/2} * IFFT_y[H4[0]]     (d/dx)
//   Z4[1] <-                      IFFT_y[i ky e^{+i ky dy/2} H4[0]]   (d/dy)
//   Z4[2] <-                      IFFT_y[H4[1]]                      (d/dz)
template <int R>
__global__ void __launch_bounds__(16 * R) k2_y_inv_grad(StepParams P, V2Params Q) {
  extern __shared__ float2 smem[];
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int kx = blockIdx.x * 16 + l, z = blockIdx.y;
  const bool active = kx < Q.Nxh;
  const long long hb = (long long)z * Q.Ny * Q.PH + kx;
  const long long zb = (long long)z * (Q.Ny / 2) * Q.Nx;
  float2 a[R], c[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) {
    const int ky = t + R * k1;
    a[k1] = active ? Q.H4[hb + (long long)ky * Q.PH] : make_float2(0.f, 0.f);
    c[k1] = cmul2(a[k1], P.dpy[ky]);
  }
  strided_fft<R, true>(a, Q.twy, smem, l, t);
  const float2 mx = active ? P.dpx[kx] : make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2* zl = Q.Z4 + zb + (long long)((t >> 1) + (R / 2) * j) * Q.Nx;
    store_row_pair(zl, cmul2(a[j], mx), t, kx, Q.Nx, active);
  }
  strided_fft<R, true>(c, Q.twy, smem, l, t);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2* zl = Q.Z4 + Q.ZS + zb + (long long)((t >> 1) + (R / 2) * j) * Q.Nx;
    store_row_pair(zl, c[j], t, kx, Q.Nx, active);
  }
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) a[k1] = active ? Q.H4[Q.HS + hb + (long long)(t + R * k1) * Q.PH] : make_float2(0.f, 0.f);
  strided_fft<R, true>(a, Q.twy, smem, l, t);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2* zl = Q.Z4 + 2 * Q.ZS + zb + (long long)((t >> 1) + (R / 2) * j) * Q.Nx;
    store_row_pair(zl, a[j], t, kx, Q.Nx, active);
  }
}

// z pass of the velocity divergence (comp 0..2, in place) and of the source field (comp 3).
// grid (PH/16, Ny); the CTA walks the components so that kappa is evaluated once per (kx,ky,kz).
// comp 2 additionally gets i kz e^{-i kz dz/2}; comp 3 reads the slab planes only and is filtered with
// cos(c_ref k dt/2).
template <int R, bool POLY>
__global__ void __launch_bounds__(16 * R, 2) k2_z_div(StepParams P, V2Params Q, int ncomp) {
  extern __shared__ float2 smem[];
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int kx = blockIdx.x * 16 + l, ky = blockIdx.y;
  const bool active = kx < Q.Nxh;
  const long long zs = (long long)Q.Ny * Q.PH;
  const long long base = (long long)ky * Q.PH + kx;
  const float axy = active ? P.ax2[kx] + P.ay2[ky] : 0.f;
  float kap[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) {
    const float a2 = axy + P.az2[t + R * k1];
    kap[k1] = (POLY ? sinc_sqrt_poly(a2) : kappa_of(a2)) * Q.norm;
  }
#pragma unroll 1
  for (int comp = 0; comp < ncomp; ++comp) {
    float2* H = Q.H4 + comp * Q.HS;
    float2 v[R];
    if (comp < 3) {
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = active ? H[base + (long long)(t + R * j) * zs] : make_float2(0.f, 0.f);
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int zr = t + R * j - Q.z0s;
        v[j] = (active && zr >= 0 && zr < Q.nzs) ? Q.HSslab[base + (long long)zr * zs] : make_float2(0.f, 0.f);
      }
    }
    strided_fft<R, false>(v, Q.twz, smem, l, t);
    if (comp < 2) {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) v[k1] = cscale(v[k1], kap[k1]);
    } else if (comp == 2) {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) v[k1] = cmul2(cscale(v[k1], kap[k1]), P.dnz[t + R * k1]);
    } else {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        const float a2 = axy + P.az2[t + R * k1];
        v[k1] = cscale(v[k1], (POLY ? cos_sqrt_poly(a2) : cosf(sqrtf(a2))) * Q.norm);
      }
    }
    strided_fft<R, true>(v, Q.twz, smem, l, t);
    if (active) {
#pragma unroll
      for (int j = 0; j < R; ++j) H[base + (long long)(t + R * j) * zs] = v[j];
    }
  }
}

// y inverse + row-pair merge of H4[comp] -> Z4[comp].  grid (PH/16, Nz, ncomp)
template <int R>
__global__ void __launch_bounds__(16 * R) k2_y_inv(StepParams P, V2Params Q) {
  extern __shared__ float2 smem[];
  const int l = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int kx = blockIdx.x * 16 + l, z = blockIdx.y, comp = blockIdx.z;
  const bool active = kx < Q.Nxh;
  const long long hb = (long long)z * Q.Ny * Q.PH + kx;
  const long long zb = (long long)z * (Q.Ny / 2) * Q.Nx;
  const float2* H = Q.H4 + comp * Q.HS;
  float2 a[R];
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) a[k1] = active ? H[hb + (long long)(t + R * k1) * Q.PH] : make_float2(0.f, 0.f);
  strided_fft<R, true>(a, Q.twy, smem, l, t);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2* zl = Q.Z4 + comp * Q.ZS + zb + (long long)((t >> 1) + (R / 2) * j) * Q.Nx;
    store_row_pair(zl, a[j], t, kx, Q.Nx, active);
  }
}

// ------------------------------------------------------------------------------------------------
// x passes.  Persistent CTAs of 128 threads = 128/R groups of R lanes; a group owns one row pair at a
// time and walks its work as a sequence of "items" (one packed spectrum line + one pair of real rows).
// Items are prefetched two deep with cp.async into the group's private shared-memory stages, so the
// memory latency of item q+1 overlaps the transforms of item q regardless of occupancy.  The FFT
// exchange runs inside the (already consumed) spectrum stage with an XOR swizzle instead of padding.
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPENDING) : "memory"); }

template <int R, bool INV>
__device__ __forceinline__ void line_fft_sw(float2 (&v)[R], const float2* __restrict__ tw, float2* sm, int t) {
  constexpr int N = R * R;
  if constexpr (!INV) {
    dft<R, false>(v);
#pragma unroll
    for (int k2 = 1; k2 < R; ++k2) v[k2] = cmul2(v[k2], tw[(t * k2) & (N - 1)]);
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) sm[t * R + (k2 ^ t)] = v[k2];
    __syncwarp();
#pragma unroll
    for (int tt = 0; tt < R; ++tt) v[tt] = sm[tt * R + (t ^ tt)];
    __syncwarp();
    dft<R, false>(v);
  } else {
    dft<R, true>(v);
#pragma unroll
    for (int tt = 1; tt < R; ++tt) v[tt] = cmulc(v[tt], tw[(tt * t) & (N - 1)]);
#pragma unroll
    for (int tt = 0; tt < R; ++tt) sm[tt * R + (t ^ tt)] = v[tt];
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < R; ++k2) v[k2] = sm[t * R + (k2 ^ t)];
    __syncwarp();
    dft<R, true>(v);
  }
}

// One group's staging: stage s = [ N float2 spectrum line | 2N floats (row pair) ]
template <int R>
struct XStage {
  static constexpr int N = R * R;
  static constexpr int BYTES = 16 * N;          // per stage
  static constexpr int GROUPS = 128 / R;
  static constexpr int SMEM = GROUPS * 2 * BYTES + 8 * N;   // + twiddle table
  // copy `bytes` (multiple of 16*R) from global to shared with the R lanes of the group
  static __device__ __forceinline__ void copy(char* sdst, const char* gsrc, int bytes, int t) {
#pragma unroll
    for (int o = 0; o < 8 * N; o += 16 * R)
      if (o < bytes) cp_async16(sdst + o + 16 * t, gsrc + o + 16 * t);
  }
};

template <int R, bool HOMOG>
__global__ void __launch_bounds__(128, 3) k2_x_u(StepParams P, V2Params Q) {
  using XS = XStage<R>;
  constexpr int N = R * R, G = XS::GROUPS;
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw + G * 2 * XS::BYTES);
  for (int i = threadIdx.x; i < N; i += 128) tw[i] = Q.twx[i];
  __syncthreads();
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  char* gbase = reinterpret_cast<char*>(smraw) + g * 2 * XS::BYTES;
  const long long nbatch = (long long)Q.Nz * (Q.Ny / 2) / G;
  const long long my_iters = blockIdx.x < nbatch ? (nbatch - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long nitems = my_iters * 3;

  auto issue = [&](long long q) {
    if (q < nitems) {
      const long long pair = (blockIdx.x + (q / 3) * gridDim.x) * G + g;
      const int c = (int)(q % 3);
      const int m = (int)(pair % (Q.Ny / 2)), z = (int)(pair / (Q.Ny / 2));
      char* st = gbase + (q & 1) * XS::BYTES;
      XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + c * Q.ZS + pair * N), 8 * N, t);
      XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.u + c * P.RS + ((long long)z * Q.Ny + 2 * m) * N), 8 * N, t);
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);
  for (long long q = 0; q < nitems; ++q) {
    cp_async_wait<1>();
    __syncwarp();
    const long long pair = (blockIdx.x + (q / 3) * gridDim.x) * G + g;
    const int c = (int)(q % 3);
    const int m = (int)(pair % (Q.Ny / 2)), z = (int)(pair / (Q.Ny / 2));
    float2* zb = reinterpret_cast<float2*>(gbase + (q & 1) * XS::BYTES);
    const float* rb = reinterpret_cast<const float*>(gbase + (q & 1) * XS::BYTES + 8 * N);
    const long long r0 = ((long long)z * Q.Ny + 2 * m) * N;
    float2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = zb[t + R * j];
    __syncwarp();
    line_fft_sw<R, true>(v, tw, zb, t);
    float* u = P.u + c * P.RS;
    float s0, s1;
    if (c == 1) { s0 = P.sgy[2 * m]; s1 = P.sgy[2 * m + 1]; } else { s0 = s1 = P.sgz[z]; }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int x = t + R * j;
      if (c == 0) s0 = s1 = P.sgx[x];
      float d0, d1;
      if constexpr (HOMOG) { d0 = d1 = P.dt_rho0_sg_s; }
      else { d0 = P.dt_rho0_sg[c * P.RS + r0 + x]; d1 = P.dt_rho0_sg[c * P.RS + r0 + N + x]; }
      const float u0 = s0 * (s0 * rb[x] - d0 * v[j].x);
      const float u1 = s1 * (s1 * rb[N + x] - d1 * v[j].y);
      u[r0 + x] = u0;
      u[r0 + N + x] = u1;
      v[j] = make_float2(u0, u1);
    }
    line_fft_sw<R, false>(v, tw, zb, t);
    float2* zo = Q.Z4 + c * Q.ZS + pair * N;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = v[k1];
    __syncwarp();
    issue(q + 2);
  }
  cp_async_wait<0>();
}

// SRC: 0 none, 1 filtered source in Z4[3], 2 unfiltered dense slab
template <int R, bool HOMOG, int SRC>
__global__ void __launch_bounds__(128, 3) k2_x_rho_p(StepParams P, V2Params Q) {
  using XS = XStage<R>;
  constexpr int N = R * R, G = XS::GROUPS;
  constexpr int NI = SRC == 1 ? 4 : 3;                 // items per row pair: [source], rho_x, rho_y, rho_z
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* tw = reinterpret_cast<float2*>(smraw + G * 2 * XS::BYTES);
  for (int i = threadIdx.x; i < N; i += 128) tw[i] = Q.twx[i];
  __syncthreads();
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  char* gbase = reinterpret_cast<char*>(smraw) + g * 2 * XS::BYTES;
  const long long nbatch = (long long)Q.Nz * (Q.Ny / 2) / G;
  const long long my_iters = blockIdx.x < nbatch ? (nbatch - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long nitems = my_iters * NI;

  auto issue = [&](long long q) {
    if (q < nitems) {
      const long long pair = (blockIdx.x + (q / NI) * gridDim.x) * G + g;
      const int it = (int)(q % NI);
      const int c = SRC == 1 ? it - 1 : it;              // -1: the source item
      const int m = (int)(pair % (Q.Ny / 2)), z = (int)(pair / (Q.Ny / 2));
      char* st = gbase + (q & 1) * XS::BYTES;
      XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + (c < 0 ? 3 : c) * Q.ZS + pair * N), 8 * N, t);
      if (c >= 0)
        XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.rho + c * P.RS + ((long long)z * Q.Ny + 2 * m) * N), 8 * N, t);
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);
  float2 src[R], sum[R];
  for (long long q = 0; q < nitems; ++q) {
    cp_async_wait<1>();
    __syncwarp();
    const long long pair = (blockIdx.x + (q / NI) * gridDim.x) * G + g;
    const int it = (int)(q % NI);
    const int c = SRC == 1 ? it - 1 : it;
    const int m = (int)(pair % (Q.Ny / 2)), z = (int)(pair / (Q.Ny / 2));
    float2* zb = reinterpret_cast<float2*>(gbase + (q & 1) * XS::BYTES);
    const float* rb = reinterpret_cast<const float*>(gbase + (q & 1) * XS::BYTES + 8 * N);
    const long long r0 = ((long long)z * Q.Ny + 2 * m) * N;
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
        const long long so = ((long long)zr * Q.Ny + 2 * m) * N;
#pragma unroll
        for (int j = 0; j < R; ++j)
          src[j] = in ? make_float2(Q.Sslab[so + t + R * j], Q.Sslab[so + N + t + R * j]) : make_float2(0.f, 0.f);
      }
      float* rho = P.rho + c * P.RS;
      float a0, a1;
      if (c == 1) { a0 = P.pmly[2 * m]; a1 = P.pmly[2 * m + 1]; } else { a0 = a1 = P.pmlz[z]; }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int x = t + R * j;
        if (c == 0) a0 = a1 = P.pmlx[x];
        float d0, d1;
        if constexpr (HOMOG) { d0 = d1 = P.dt_rho0_s; }
        else { d0 = P.dt_rho0[r0 + x]; d1 = P.dt_rho0[r0 + N + x]; }
        float q0 = a0 * (a0 * rb[x] - d0 * v[j].x);
        float q1 = a1 * (a1 * rb[N + x] - d1 * v[j].y);
        if constexpr (SRC != 0) { q0 += src[j].x; q1 += src[j].y; }
        rho[r0 + x] = q0;
        rho[r0 + N + x] = q1;
        if (c == 0) sum[j] = make_float2(q0, q1);
        else { sum[j].x += q0; sum[j].y += q1; }          // (rho_x + rho_y) + rho_z
      }
      if (c == 2) {
        // equation of state, sensor reduction, forward transform of the new pressure
        const int jz = z - P.pz, jy0 = 2 * m - P.py, jy1 = jy0 + 1;
        const bool zin = (unsigned)jz < (unsigned)P.nz;
        const bool in0 = zin && (unsigned)jy0 < (unsigned)P.ny, in1 = zin && (unsigned)jy1 < (unsigned)P.ny;
        const long long s0 = ((long long)jz * P.ny + jy0) * P.nx - P.px, s1 = s0 + P.nx;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int x = t + R * j;
          float c0, c1;
          if constexpr (HOMOG) { c0 = c1 = P.c2_s; } else { c0 = P.c2[r0 + x]; c1 = P.c2[r0 + N + x]; }
          const float p0 = c0 * sum[j].x, p1 = c1 * sum[j].y;
          sum[j] = make_float2(p0, p1);
          if (Q.store_p) { P.p[r0 + x] = p0; P.p[r0 + N + x] = p1; }
          const bool xin = (unsigned)(x - P.px) < (unsigned)P.nx;
          if (xin && in0) { P.pmax[s0 + x] = fmaxf(P.pmax[s0 + x], p0); P.pmin[s0 + x] = fminf(P.pmin[s0 + x], p0); }
          if (xin && in1) { P.pmax[s1 + x] = fmaxf(P.pmax[s1 + x], p1); P.pmin[s1 + x] = fminf(P.pmin[s1 + x], p1); }
        }
        line_fft_sw<R, false>(sum, tw, zb, t);
        float2* zo = Q.ZP + pair * N;
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) zo[t + R * k1] = sum[k1];
      }
    }
    __syncwarp();
    issue(q + 2);
  }
  cp_async_wait<0>();
  if (blockIdx.x == 0 && threadIdx.x == 0) *P.step = *P.step + 1;
}

// x forward of the dense source slab (row pairs).  grid = nzs*(Ny/2)/G
template <int R>
__global__ void __launch_bounds__(256) k2_x_src(StepParams P, V2Params Q) {
  extern __shared__ float2 smem[];
  constexpr int G = 256 / R;
  const int g = threadIdx.x / R, t = threadIdx.x % R;
  float2* sm = smem + g * R * (R + 1);
  const long long pair = (long long)blockIdx.x * G + g;          // zr*(Ny/2) + m
  if (pair >= (long long)Q.nzs * (Q.Ny / 2)) return;
  const long long r0 = pair * 2 * Q.Nx;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = make_float2(Q.Sslab[r0 + t + R * j], Q.Sslab[r0 + Q.Nx + t + R * j]);
  line_fft<R, false>(v, Q.twx, sm, t);
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) Q.ZSslab[pair * Q.Nx + t + R * k1] = v[k1];
}

// Source scatter into the dense slab (same arithmetic as k_source_scatter of v1).
__global__ void __launch_bounds__(128) k2_source_scatter(StepParams P, V2Params Q, SourceParams S) {
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

}  // namespace lifu
