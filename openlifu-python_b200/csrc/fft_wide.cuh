// fft_wide.cuh -- pipeline "wide": the fused FFT passes of fft_v2.cuh for axis lengths N = A * B with A <= B,
// B in {8, 16, 32}: 64 (8 x 8), 128 (8 x 16), 256 (16 x 16), 512 (16 x 32), 768 (24 x 32), 1024 (32 x 32).
//
// Why: pml_auto turns BASELINE config C5 (728^3 inner) into 768^3 and the 472^3 grid of the slab leg into 512^3; pipeline
// v2 is specialised for square factorisations (64 = 8 x 8, 256 = 16 x 16) and sends every other grid to the library FFT
// (/root/reference/src/openlifu/sim/kwave_if.py:117-129 is the call all of this replaces).  Same data flow, same fusion,
// same Z / H layouts as v2; what is generalised is the line transform.
//
// Line transform.  A line of N = A * B points is transformed by B threads.  Real ("A") side: thread t < B holds
// x[t + B i], i < A.  A register DFT of size A over i, the twiddle w_N^(t ka), ONE exchange through shared memory, then
// threads u < A hold Y[t'][u] (t' < B) and do a register DFT of size B: X[u + A kb] sits in (thread u, register kb) --
// the spectral ("B") side.  With A < B the threads u >= A idle through the second DFT (768: 8 of 32; whole warps in the
// strided kernels, masked lanes in the x kernels, where a warp runs the DFT once whatever its mask) -- the price for keeping
// everything in registers with one exchange, the property that makes v2 fast.  The inverse runs the same steps backwards,
// so forward -> point-wise operator -> inverse chains need no reordering.
//
// Tiles of the strided (y, z) passes: CTA = LANES kx lanes x B threads (LANES = 8 when B = 32: 256 threads, 64-byte row
// segments; 16 otherwise).  Every strided kernel here is ONE transform chain per CTA (component / pass = blockIdx.z):
// with up to 32 complex registers per thread there is no room for the multi-component register pipelines of
// k2_z_grad / k2_z_div / k2_y_inv_grad.  The pressure-gradient z pass is therefore out of place: H4[0] -> H4[2] (plain)
// and H4[1] (z derivative) -- on one GPU as a split chain (forward + kappa once into field 3, then the two inverse chains),
// in a slab decomposition as two full chains whose stores are routed to the owning ranks.
//
// x passes: persistent CTAs, a group of B lanes owns one row pair at a time and prefetches its items two deep with
// cp.async into private stages exactly as in fft_v2.cuh; the exchange runs inside the consumed spectrum stage with a
// rotation swizzle (conflict degree <= 2 for every (A, B) above).
#pragma once
#include "fft_v2.cuh"

// kx lanes of a strided tile when a line takes 32 threads (512 / 768 / 1024-point axes), template parameter L32 of the
// strided kernels: 8 -> 256-thread CTAs, two per SM, 64-byte row segments (faster on one GPU); 16 -> 512-thread CTAs, one per
// SM, 128-byte row segments -- what the NVLink stores of a slab decomposition want (profiles/r2_wide_summary.md).

namespace lifu {

__host__ __device__ constexpr float w24_re(int m) {
  switch (m % 24) {
    case 0: return 1.f;
    case 1: return 0.9659258263f;   case 2: return 0.8660254038f;   case 3: return 0.7071067812f;
    case 4: return 0.5f;            case 5: return 0.2588190451f;   case 6: return 0.f;
    case 7: return -0.2588190451f;  case 8: return -0.5f;           case 9: return -0.7071067812f;
    case 10: return -0.8660254038f; case 11: return -0.9659258263f; case 12: return -1.f;
    case 13: return -0.9659258263f; case 14: return -0.8660254038f; case 15: return -0.7071067812f;
    case 16: return -0.5f;          case 17: return -0.2588190451f; case 18: return 0.f;
    case 19: return 0.2588190451f;  case 20: return 0.5f;           case 21: return 0.7071067812f;
    case 22: return 0.8660254038f;  default: return 0.9659258263f;
  }
}
__host__ __device__ constexpr float w24_im(int m) { return w24_re(m + 6); }   // -sin(2 pi m / 24) = cos(2 pi (m + 6) / 24)

template <bool INV> __device__ __forceinline__ void wdft3(float2& a, float2& b, float2& c) {
  const float S3 = 0.86602540378443864676f;
  const float2 s = cadd(b, c), d = csub(b, c);
  const float2 m = __ffma2_rn(s, make_float2(-0.5f, -0.5f), a);
  // forward: X1 = m - i S3 d, X2 = m + i S3 d
  const float2 e = __fmul2_rn(cswap(d), INV ? make_float2(-S3, S3) : make_float2(S3, -S3));
  a = cadd(a, s);
  b = cadd(m, e);
  c = csub(m, e);
}

// DFT-24 = 3 x 8 (n = b + 3 a, k = ka + 8 kb), natural order in and out
template <bool INV> __device__ __forceinline__ void wdft24(float2 (&x)[24]) {
  float2 y[3][8];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    float2 s[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) s[a] = x[b + 3 * a];
    dft<8, INV>(s);
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) {
      if (b * ka == 0) y[b][ka] = s[ka];
      else {
        const float wr = w24_re(b * ka), wi = INV ? -w24_im(b * ka) : w24_im(b * ka);
        y[b][ka] = cmul4(s[ka], make_float4(wr, wi, -wi, wr));
      }
    }
  }
#pragma unroll
  for (int ka = 0; ka < 8; ++ka) {
    float2 p = y[0][ka], q = y[1][ka], r = y[2][ka];
    wdft3<INV>(p, q, r);
    x[ka] = p; x[ka + 8] = q; x[ka + 16] = r;
  }
}
template <int N, bool INV> __device__ __forceinline__ void wdft(float2 (&x)[N]) {
  if constexpr (N == 24) wdft24<INV>(x);
  else dft<N, INV>(x);
}

// ------------------------------------------------------------------------------------------------
// geometry of a strided tile
template <int A, int B, int L32 = 8> struct Wide {
  static constexpr int N = A * B;
  static constexpr int LANES = B == 32 ? L32 : 16;
  static constexpr int THREADS = LANES * B;
  static constexpr int MINB = THREADS > 256 ? 1 : 2;             // CTAs per SM the register budget is set for
  static constexpr int PAD = LANES == 8 ? 8 : 0;                 // float2 per t row: keeps half-warp accesses conflict free
  static constexpr int PITCH = A * LANES + PAD;                  // float2 between consecutive t
  static constexpr int XCH = B * PITCH * 8;                      // bytes of the exchange buffer
  static constexpr int TW = (N + 1) * 16;                        // (w, i w) table, entry N = entry 0
  static constexpr int SMEM = XCH + TW;
  static __device__ __forceinline__ const float4* load_tw(unsigned char* smraw, const float4* __restrict__ g) {
    float4* s = reinterpret_cast<float4*>(smraw + XCH);
    for (int i = threadIdx.x; i <= N; i += THREADS) s[i] = g[i];
    __syncthreads();
    return s;
  }
};

// One strided line transform.  Forward: in v[0..A) = x[t + B i] on every thread, out v[0..B) = X[t + A kb] on threads
// t < A.  Inverse: the reverse.  One barrier inside; the caller puts a barrier between two transforms that share `sm`.
template <int A, int B, bool INV, int L32 = 8>
__device__ __forceinline__ void wfft_strided(float2 (&v)[B], const float4* __restrict__ tw, float2* sm, int l, int t) {
  using W = Wide<A, B, L32>;
  constexpr int N = A * B, L = W::LANES, PT = W::PITCH;
  const bool act = (A == B) || t < A;
  if constexpr (!INV) {
    float2 s[A];
#pragma unroll
    for (int i = 0; i < A; ++i) s[i] = v[i];
    wdft<A, false>(s);
#pragma unroll
    for (int ka = 1; ka < A; ++ka) s[ka] = cmul4(s[ka], tw[t * ka]);
    float2* w = sm + t * PT + l;
#pragma unroll
    for (int ka = 0; ka < A; ++ka) w[ka * L] = s[ka];
    __syncthreads();
    if (act) {
      const float2* r = sm + t * L + l;
#pragma unroll
      for (int tt = 0; tt < B; ++tt) v[tt] = r[tt * PT];
      wdft<B, false>(v);
    }
  } else {
    if (act) {
      wdft<B, true>(v);
#pragma unroll
      for (int tt = 1; tt < B; ++tt) v[tt] = cmul4(v[tt], tw[N - tt * t]);
      float2* w = sm + t * L + l;
#pragma unroll
      for (int tt = 0; tt < B; ++tt) w[tt * PT] = v[tt];
    }
    __syncthreads();
    float2 s[A];
    const float2* r = sm + t * PT + l;
#pragma unroll
    for (int ka = 0; ka < A; ++ka) s[ka] = r[ka * L];
    wdft<A, true>(s);
#pragma unroll
    for (int i = 0; i < A; ++i) v[i] = s[i];
  }
}

// lane -> (kx, other index) of a strided CTA; false when the CTA has no work.  grid.x = Nx / (2 LANES) regular tiles + 1
// Nyquist slot whose lanes run over the other in-plane index.
template <int LANES>
__device__ __forceinline__ bool wlane_map(const V2Params& Q, int n_other, int l, int& kx, int& o, int by, int bx = (int)blockIdx.x) {
  const int nxt = Q.Nx / (2 * LANES);
  if (bx < nxt) { kx = bx * LANES + l; o = by; return true; }
  if (by * LANES >= n_other) return false;
  kx = Q.Nx >> 1;
  o = by * LANES + l;
  return true;
}

// row pair index (z * Ny/2 + m) -> z and the lower row of the pair (rows ylo and ylo + By)
__device__ __forceinline__ void wpair_rows(const V2Params& Q, int pair, int& z, int& ylo) {
  const int hy = Q.Ny >> 1;
  z = pair / hy;
  const int m = pair - z * hy;
  ylo = ((m >> Q.ry_sh) << (Q.ry_sh + 1)) | (m & (Q.Ry - 1));
}

// merge the rows of the pairs held by one thread (adjacent registers = rows y, y + B) into the packed x-spectrum
template <int A, int B>
__device__ __forceinline__ void wmerge_store(float2* __restrict__ zp, const float2 (&v)[B], int kx, int Nx, bool live) {
  const bool selfm = (kx == 0) || (2 * kx == Nx);
  const int km = Nx - kx;
  const long long qstep = (long long)B * Nx;
#pragma unroll
  for (int q = 0; q < A / 2; ++q) {
    const float2 a = v[2 * q], b = v[2 * q + 1];
    const float2 lo = selfm ? make_float2(a.x, b.x) : cadd_i(a, b);
    if (live) zp[q * qstep + kx] = lo;
    if (live && !selfm) zp[q * qstep + km] = cconj(csub_i(a, b));
  }
}

// ------------------------------------------------------------------------------------------------
// y forward: packed row pairs -> half spectrum.  grid (tiles + 1, planes | plane blocks, ncomp)
// MODE 0 pressure; 1 velocity (comp 0: x multiplier, comp 1: y multiplier); 2 source slab; 3 absorption operands
template <int A, int B, int MODE, int L32 = 8>
__global__ void __launch_bounds__(Wide<A, B, L32>::THREADS, Wide<A, B, L32>::MINB) kw_y_fwd(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using W = Wide<A, B, L32>;
  constexpr int L = W::LANES;
  const int l = threadIdx.x % L, t = threadIdx.x / L;
  const int comp = blockIdx.z;
  const int nz = MODE == 2 ? Q.nzs : Q.Nz;
  int kx, z;
  if (!wlane_map<L>(Q, nz, l, kx, z, (int)blockIdx.y)) return;
  const bool live = z < nz;
  const int zc = live ? z : nz - 1;
  float2* xa = reinterpret_cast<float2*>(smraw);
  const float2* Zin = MODE == 0 ? Q.ZP : ((MODE == 1 || MODE == 3) ? Q.Z4 + comp * Q.ZS : Q.ZSslab);
  float2* Hout = MODE == 2 ? Q.HSslab : Q.H4 + comp * Q.HS;
  const int km = kx == 0 ? 0 : Q.Nx - kx;
  const float2* zp = Zin + ((long long)zc * (Q.Ny / 2) + t) * Q.Nx;      // packed line m = q*B + t
  const long long qstep = (long long)B * Q.Nx;
  float2 v[B];
#pragma unroll
  for (int q = 0; q < A / 2; ++q) { v[2 * q] = zp[q * qstep + kx]; v[2 * q + 1] = zp[q * qstep + km]; }
  const float4* tw = W::load_tw(smraw, Q.tw4y);
#pragma unroll
  for (int q = 0; q < A / 2; ++q) {
    const float2 d = v[2 * q], m = v[2 * q + 1];
    v[2 * q] = cadd_conj(d, m);                  // 2 (row t + B 2q)
    v[2 * q + 1] = cmul_mi(csub_conj(d, m));     // 2 (row t + B (2q+1))
  }
  if (MODE == 1 && comp == 0) {
    const float4 mx = with_i(P.dnx[kx]);
#pragma unroll
    for (int i = 0; i < A; ++i) v[i] = cmul4(v[i], mx);
  }
  wfft_strided<A, B, false, L32>(v, tw, xa, l, t);
  if (live && (A == B || t < A)) {
    if (Q.G == 0) {
      float2* hp = Hout + (long long)z * Q.zsH + (long long)t * Q.PH + kx;   // ky = t + A kb
      const long long kstep = (long long)A * Q.PH;
#pragma unroll
      for (int kb = 0; kb < B; ++kb) {
        float2 o = v[kb];
        if (MODE == 1 && comp == 1) o = cmul4(o, Q.dny4[t + A * kb]);
        hp[kb * kstep] = o;
      }
    } else {
      // slab decomposition: row ky of plane z0g + z belongs to rank ky / Nyl; store it into that rank's T4 over NVLink
      // (G divides B: ky = t + A kb with t < A lies in block kb / (B / G), no division needed)
      const int fld = MODE == 2 ? 3 : comp;
      const long long zrow = (long long)(Q.z0g + z + (MODE == 2 ? Q.z0s : 0)) * Q.Nyl;
      const int bg = B / Q.G;
      int q = 0, rem = 0;
      float2* dst = Q.peer[0] + Q.peerT + fld * Q.HS + (zrow + t) * Q.PH + kx;
#pragma unroll
      for (int kb = 0; kb < B; ++kb) {
        float2 o = v[kb];
        if (MODE == 1 && comp == 1) o = cmul4(o, Q.dny4[t + A * kb]);
        dst[(long long)rem * A * Q.PH] = o;
        if (++rem == bg) { rem = 0; ++q; if (q < Q.G) dst = Q.peer[q] + Q.peerT + fld * Q.HS + (zrow + t) * Q.PH + kx; }
      }
    }
  }
}

// y inverse + row-pair merge.  grid (tiles + 1, Nz, ncomp)
// GRAD: the three pressure-gradient components from the two z-pass outputs:
//   Z4[0] <- i kx e^{+i kx dx/2} IFFT_y[H4[2]];  Z4[1] <- IFFT_y[i ky e^{+i ky dy/2} H4[2]];  Z4[2] <- IFFT_y[H4[1]]
// otherwise Z4[comp0 + c] <- IFFT_y[H4[comp0 + c]]
template <int A, int B, bool GRAD, int L32 = 8>
__global__ void __launch_bounds__(Wide<A, B, L32>::THREADS, Wide<A, B, L32>::MINB) kw_y_inv(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using W = Wide<A, B, L32>;
  constexpr int L = W::LANES;
  const int l = threadIdx.x % L, t = threadIdx.x / L;
  const int comp = (int)blockIdx.z + (GRAD ? 0 : Q.comp0);
  int kx, z;
  if (!wlane_map<L>(Q, Q.Nz, l, kx, z, (int)blockIdx.y)) return;
  const bool live = z < Q.Nz;
  const int zc = live ? z : Q.Nz - 1;
  float2* xa = reinterpret_cast<float2*>(smraw);
  const int hsel = GRAD ? (comp == 2 ? 1 : 2) : comp;
  const float2* hp = Q.H4 + hsel * Q.HS + (long long)zc * Q.zsH + (long long)t * Q.PH + kx;
  const long long kstep = (long long)A * Q.PH;
  float2 a[B];
  if (A == B || t < A) {
#pragma unroll
    for (int kb = 0; kb < B; ++kb) a[kb] = hp[kb * kstep];
  }
  const float4* tw = W::load_tw(smraw, Q.tw4y);
  if (GRAD && comp == 1 && (A == B || t < A)) {
#pragma unroll
    for (int kb = 0; kb < B; ++kb) a[kb] = cmul4(a[kb], Q.dpy4[t + A * kb]);
  }
  wfft_strided<A, B, true, L32>(a, tw, xa, l, t);
  if (GRAD && comp == 0) {
    const float4 mx = with_i(P.dpx[kx]);
#pragma unroll
    for (int i = 0; i < A; ++i) a[i] = cmul4(a[i], mx);
  }
  wmerge_store<A, B>(Q.Z4 + comp * Q.ZS + ((long long)zc * (Q.Ny / 2) + t) * Q.Nx, a, kx, Q.Nx, live);
}

// exact kappa / source filter (sinf / cosf with their argument-reduction slow paths) out of line: inlined at the 32 unrolled
// sites of kw_z they made the kernel 10.8 K instructions long, a quarter of its stall samples sat in that region with
// instruction-fetch stalls although the polynomial branch is the one that runs (profiles/r2_wide512_zdiv_stalls.txt)
static __device__ __noinline__ float wkappa_exact(float a2) { return kappa_of(a2); }
static __device__ __noinline__ float wcos_exact(float a2) { return cosf(sqrtf(a2)); }

// z passes, one transform chain per CTA.  grid (tiles + 1, Ny | Ny blocks, chains)
// OP 0 pressure gradient: chain 0: H4[2] <- IFFT_z[kappa FFT_z H4[0]]; chain 1: H4[1] <- IFFT_z[i kz e^{+i kz dz/2} kappa FFT_z H4[0]]
// OP 1 divergence, in place: comps comp0 .. : kappa (0, 1), i kz e^{-i kz dz/2} kappa (2)
// OP 2 absorption operands, in place: k^(y-2) (0), k^(y-1) (1)
// OP 3 source field: source slab (or T4[3] of a slab decomposition) x cos(c_ref k dt/2) -> H4[3]; its own instantiation so
//      that the divergence chains do not carry the unrolled cosine code
// PART 0: the whole chain.  PART 1 / 2 (experiment, LIFU_WIDE_ZSPLIT=1): the chain as two kernels -- forward transform +
// operator, spectrum stored in natural kz order (gradient: into field 3, otherwise in place); then spectrum -> inverse
// transform -> result.  Twice the traffic of the z pass for half the chain length per CTA.
template <int A, int B, int OP, int L32 = 8, int PART = 0>
__global__ void __launch_bounds__(Wide<A, B, L32>::THREADS, Wide<A, B, L32>::MINB) kw_z(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using W = Wide<A, B, L32>;
  constexpr int L = W::LANES;
  const int l = threadIdx.x % L, t = threadIdx.x / L;
  // OP 0: the two chains of a tile sit in neighbouring CTAs (grid.x = 2 x tiles), so the second read of the column tile
  // finds it in L2 (with the chain in grid.z it came from DRAM again: 2.2 GB per launch instead of 1.6 GB at 512^3)
  const int chain = OP == 0 ? (PART == 1 ? 0 : ((int)blockIdx.x & 1)) : (OP == 3 ? 3 : (int)blockIdx.z + (OP == 1 ? Q.comp0 : 0));
  int kx, ky;
  const int nky = Q.G ? Q.Nyl : Q.Ny;                 // ky rows held here (all of them without a slab decomposition)
  if (!wlane_map<L>(Q, nky, l, kx, ky, (int)blockIdx.y, (OP == 0 && PART != 1) ? (int)blockIdx.x >> 1 : (int)blockIdx.x)) return;
  const bool live = ky < nky;
  const int kyc = live ? ky : nky - 1;
  const int kyg = kyc + (Q.G ? Q.ky0 : 0);           // global ky: index of the 1-D tables
  float2* xa = reinterpret_cast<float2*>(smraw);
  const long long zs = Q.G ? (long long)Q.Nyl * Q.PH : Q.zsH;
  const long long col = (long long)kyc * Q.PH + kx;
  const float2* base = Q.G ? Q.T4 : Q.H4;
  const float2* in = OP == 0 ? base + col : base + chain * Q.HS + col;
  const int fout = OP == 0 ? (chain == 0 ? 2 : 1) : chain;
  float2 v[B];
  // spectrum of a split chain: natural kz order, gradient in field 3, otherwise the chain's own field
  float2* spec = const_cast<float2*>(base) + (OP == 0 ? 3 : chain) * Q.HS + col;
  if (PART == 2) {
    if (A == B || t < A) {
#pragma unroll
      for (int kb = 0; kb < B; ++kb) {
        v[kb] = spec[(long long)(t + A * kb) * zs];
        if (OP == 0 && chain == 1) v[kb] = cmul4(v[kb], Q.dpz4[t + A * kb]);
      }
    }
  } else if (OP == 3) {
    if (Q.G) {
      // the owners of the source planes stored their rows into T4[3]; the other planes of that field are not defined
#pragma unroll
      for (int i = 0; i < A; ++i) {
        const int zr = t + B * i - Q.gz0s;
        v[i] = (zr >= 0 && zr < Q.gnzs) ? in[(long long)(t + B * i) * zs] : make_float2(0.f, 0.f);
      }
    } else {
      const float2* sp = Q.HSslab + col;
#pragma unroll
      for (int i = 0; i < A; ++i) {
        const int zr = t + B * i - Q.z0s;
        v[i] = (zr >= 0 && zr < Q.nzs) ? sp[(long long)zr * zs] : make_float2(0.f, 0.f);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < A; ++i) v[i] = in[(long long)(t + B * i) * zs];
  }
  const float4* tw = W::load_tw(smraw, Q.tw4z);
  if (PART != 2) wfft_strided<A, B, false, L32>(v, tw, xa, l, t);
  if (PART != 2 && (A == B || t < A)) {
    if (OP == 2) {
      const float kxy = P.kx2[kx] + P.ky2[kyg];
      const float e = chain == 0 ? P.y_minus2_half : P.y_minus1_half;
#pragma unroll
      for (int kb = 0; kb < B; ++kb) {
        const float k2 = kxy + P.kz2[t + A * kb];
        v[kb] = cscale(v[kb], k2 > 0.f ? __powf(k2, e) * Q.norm : 0.f);
      }
    } else {
      const float axy = P.ax2[kx] + P.ay2[kyg];
#pragma unroll
      for (int kb = 0; kb < B; ++kb) {
        const int kz = t + A * kb;
        const float a2 = axy + P.az2[kz];
        float m;
        if (OP == 3) m = P.poly_ok == 2 ? cos_sqrt_poly8(a2) : (P.poly_ok == 1 ? cos_sqrt_poly(a2) : wcos_exact(a2));
        else m = P.poly_ok == 2 ? sinc_sqrt_poly8(a2) : (P.poly_ok == 1 ? sinc_sqrt_poly(a2) : wkappa_exact(a2));
        v[kb] = cscale(v[kb], m * Q.norm);
        if (PART == 0 && OP == 0 && chain == 1) v[kb] = cmul4(v[kb], Q.dpz4[kz]);
        if (OP == 1 && chain == 2) v[kb] = cmul4(v[kb], Q.dnz4[kz]);
      }
    }
  }
  if (PART == 1) {
    if (live && (A == B || t < A)) {
#pragma unroll
      for (int kb = 0; kb < B; ++kb) spec[(long long)(t + A * kb) * zs] = v[kb];
    }
    return;
  }
  if (PART == 0) __syncthreads();                // the exchange buffer is read out before the inverse reuses it
  wfft_strided<A, B, true, L32>(v, tw, xa, l, t);
  if (live) {
    if (Q.G == 0) {
      float2* out = Q.H4 + fout * Q.HS + col;
#pragma unroll
      for (int i = 0; i < A; ++i) out[(long long)(t + B * i) * zs] = v[i];
    } else {
      // slab decomposition: plane z belongs to rank z / Nzl; store the row into that rank's H4 over NVLink
      // (G divides A: z = t + B i with t < B lies in block i / (A / G))
      const long long rowg = (long long)kyg * Q.PH + kx;
      const int ag = A / Q.G;
      int q = 0, rem = 0;
      float2* dst = Q.peer[0] + fout * Q.HS + (long long)t * Q.zsH + rowg;
#pragma unroll
      for (int i = 0; i < A; ++i) {
        dst[(long long)rem * B * Q.zsH] = v[i];
        if (++rem == ag) { rem = 0; ++q; if (q < Q.G) dst = Q.peer[q] + fout * Q.HS + (long long)t * Q.zsH + rowg; }
      }
    }
  }
}

// EXPERIMENT (instantiated only with -DLIFU_WIDE_ZP, measured slower than kw_z: profiles/r2_wide_summary.md).
// The same z chains as kw_z in PERSISTENT CTAs that keep the column tile of their next item in flight.  A chain is
// load -> DFT -> exchange -> DFT -> operator -> DFT -> exchange -> DFT -> store; with two short-lived CTAs per SM the loads of
// one CTA overlap only whatever the other happens to compute (kw_z: 1.3 TB/s of DRAM traffic against 4.3 TB/s for the
// one-transform y kernels of the same tile shape, profiles/r2_wide_summary.md).  Here every thread copies ITS OWN A values of
// the next item into thread-private shared-memory slots with cp.async (no barrier: a thread reads back only what it copied
// itself) as soon as the forward transform of the current item is through, so a tile is on its way during the rest of the chain.
// Items = (tile, chain) with the chain fastest: the two gradient chains of a tile run in neighbouring CTAs at the same time
// (second read from L2).  grid = resident CTAs (wide.cu), `nchain` = chains per tile.
__device__ __forceinline__ void cp_async8w(void* sdst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
template <int A, int B, int L32 = 8> struct WideZP {
  using W = Wide<A, B, L32>;
  static constexpr int STAGE = A * W::THREADS * 8;               // bytes: A complex values per thread
  static constexpr int SMEM = W::XCH + W::TW + STAGE;
};

template <int A, int B, int OP, int L32 = 8>
__global__ void __launch_bounds__(Wide<A, B, L32>::THREADS, Wide<A, B, L32>::MINB) kw_zp(StepParams P, V2Params Q, int nchain) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using W = Wide<A, B, L32>;
  constexpr int L = W::LANES, TH = W::THREADS;
  const int tid = threadIdx.x, l = tid % L, t = tid / L;
  float2* xa = reinterpret_cast<float2*>(smraw);
  float2* stage = reinterpret_cast<float2*>(smraw + W::XCH + W::TW) + tid;      // this thread's slots: stage[i * TH]
  const int nky = Q.G ? Q.Nyl : Q.Ny;
  const int nxt = Q.Nx / (2 * L);
  const int nreg = nxt * nky;                                                   // regular tiles, then the Nyquist column
  const int nitem = (nreg + (nky + L - 1) / L) * nchain;
  const long long zs = Q.G ? (long long)Q.Nyl * Q.PH : Q.zsH;
  const float2* base = Q.G ? Q.T4 : Q.H4;
  auto decode = [&](int it, int& chain, int& kx, int& ky) {
    const int tile = it / nchain;
    chain = OP == 3 ? 3 : (it - tile * nchain) + (OP == 1 ? Q.comp0 : 0);
    if (tile < nreg) { const int by = tile / nxt; kx = (tile - by * nxt) * L + l; ky = by; }
    else { kx = Q.Nx >> 1; ky = (tile - nreg) * L + l; }
  };
  auto issue = [&](int it) {
    if (it < nitem) {
      int chain, kx, ky;
      decode(it, chain, kx, ky);
      const int kyc = ky < nky ? ky : nky - 1;
      const long long col = (long long)kyc * Q.PH + kx;
      if (OP == 3) {
        const float2* in = Q.G ? base + 3 * Q.HS + col : Q.HSslab + col;
        const int z0 = Q.G ? Q.gz0s : Q.z0s, nz = Q.G ? Q.gnzs : Q.nzs;
#pragma unroll
        for (int i = 0; i < A; ++i) {
          const int z = t + B * i, zr = z - z0;
          if (zr >= 0 && zr < nz) cp_async8w(stage + i * TH, in + (long long)(Q.G ? z : zr) * zs);
          else stage[i * TH] = make_float2(0.f, 0.f);
        }
      } else {
        const float2* in = (OP == 0 ? base : base + chain * Q.HS) + col;
#pragma unroll
        for (int i = 0; i < A; ++i) cp_async8w(stage + i * TH, in + (long long)(t + B * i) * zs);
      }
    }
    cp_async_commit();
  };
  int it = blockIdx.x;
  issue(it);
  const float4* tw = W::load_tw(smraw, Q.tw4z);
  for (; it < nitem; it += gridDim.x) {
    int chain, kx, ky;
    decode(it, chain, kx, ky);
    const bool live = ky < nky;
    const int kyc = live ? ky : nky - 1;
    const int kyg = kyc + (Q.G ? Q.ky0 : 0);
    const long long col = (long long)kyc * Q.PH + kx;
    const int fout = OP == 0 ? (chain == 0 ? 2 : 1) : chain;
    cp_async_wait<0>();
    float2 v[B];
#pragma unroll
    for (int i = 0; i < A; ++i) v[i] = stage[i * TH];
    wfft_strided<A, B, false, L32>(v, tw, xa, l, t);
    issue(it + gridDim.x);                          // the slots were consumed by the first DFT: next tile on its way
    if (A == B || t < A) {
      if (OP == 2) {
        const float kxy = P.kx2[kx] + P.ky2[kyg];
        const float e = chain == 0 ? P.y_minus2_half : P.y_minus1_half;
#pragma unroll
        for (int kb = 0; kb < B; ++kb) {
          const float k2 = kxy + P.kz2[t + A * kb];
          v[kb] = cscale(v[kb], k2 > 0.f ? __powf(k2, e) * Q.norm : 0.f);
        }
      } else {
        const float axy = P.ax2[kx] + P.ay2[kyg];
#pragma unroll
        for (int kb = 0; kb < B; ++kb) {
          const int kz = t + A * kb;
          const float a2 = axy + P.az2[kz];
          float m;
          if (OP == 3) m = P.poly_ok == 2 ? cos_sqrt_poly8(a2) : (P.poly_ok == 1 ? cos_sqrt_poly(a2) : wcos_exact(a2));
          else m = P.poly_ok == 2 ? sinc_sqrt_poly8(a2) : (P.poly_ok == 1 ? sinc_sqrt_poly(a2) : wkappa_exact(a2));
          v[kb] = cscale(v[kb], m * Q.norm);
          if (OP == 0 && chain == 1) v[kb] = cmul4(v[kb], Q.dpz4[kz]);
          if (OP == 1 && chain == 2) v[kb] = cmul4(v[kb], Q.dnz4[kz]);
        }
      }
    }
    __syncthreads();                               // the exchange buffer is read out before the inverse reuses it
    wfft_strided<A, B, true, L32>(v, tw, xa, l, t);
    if (live) {
      if (Q.G == 0) {
        float2* out = Q.H4 + fout * Q.HS + col;
#pragma unroll
        for (int i = 0; i < A; ++i) out[(long long)(t + B * i) * zs] = v[i];
      } else {
        const long long rowg = (long long)kyg * Q.PH + kx;
        const int ag = A / Q.G;
        int q = 0, rem = 0;
        float2* dst = Q.peer[0] + fout * Q.HS + (long long)t * Q.zsH + rowg;
#pragma unroll
        for (int i = 0; i < A; ++i) {
          dst[(long long)rem * B * Q.zsH] = v[i];
          if (++rem == ag) { rem = 0; ++q; if (q < Q.G) dst = Q.peer[q] + fout * Q.HS + (long long)t * Q.zsH + rowg; }
        }
      }
    }
    __syncthreads();                               // ... and before the next item's forward transform writes it
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// x passes
// Line transform of a group of B lanes inside one warp; exchange in `sm` (N float2) with a rotation swizzle.
template <int A, int B, bool INV>
__device__ __forceinline__ void wline_fft(float2 (&v)[B], const float4* __restrict__ tw, float2* sm, int t) {
  constexpr int N = A * B;
  const bool act = (A == B) || t < A;
  const int r = t >= A ? t - A : t;                                // t mod A (B <= 2 A for every supported pair ... or A == B)
  auto rot = [](int c) { return c >= A ? c - A : c; };
  if constexpr (!INV) {
    float2 s[A];
#pragma unroll
    for (int i = 0; i < A; ++i) s[i] = v[i];
    wdft<A, false>(s);
#pragma unroll
    for (int ka = 1; ka < A; ++ka) s[ka] = cmul4(s[ka], tw[t * ka]);
#pragma unroll
    for (int ka = 0; ka < A; ++ka) sm[t * A + rot(ka + r)] = s[ka];
    __syncwarp();
    if (act) {
#pragma unroll
      for (int tt = 0; tt < B; ++tt) v[tt] = sm[tt * A + rot(t + (tt % A))];
    }
    __syncwarp();
    if (act) wdft<B, false>(v);
  } else {
    if (act) {
      wdft<B, true>(v);
#pragma unroll
      for (int tt = 1; tt < B; ++tt) v[tt] = cmul4(v[tt], tw[N - tt * t]);
#pragma unroll
      for (int tt = 0; tt < B; ++tt) sm[tt * A + rot(t + (tt % A))] = v[tt];
    }
    __syncwarp();
    float2 s[A];
#pragma unroll
    for (int ka = 0; ka < A; ++ka) s[ka] = sm[t * A + rot(ka + r)];
    __syncwarp();
    wdft<A, true>(s);
#pragma unroll
    for (int i = 0; i < A; ++i) v[i] = s[i];
  }
}

// staging of one group (see XStage in fft_v2.cuh): stage = SB * N bytes, XB * N extra bytes behind the ring
template <int A, int B, int TPB, int SB = 16, int XB = 0>
struct WStage {
  static constexpr int N = A * B;
  static constexpr int BYTES = SB * N;
  static constexpr int GROUPS = TPB / B;
  static constexpr int GBYTES = 2 * BYTES + XB * N;
  static constexpr int SMEM = GROUPS * GBYTES + 16 * (N + 1);
  static __device__ __forceinline__ void copy(char* sdst, const char* gsrc, int bytes, int t) {
#pragma unroll
    for (int o = 0; o < 8 * N; o += 16 * B)
      if (o < bytes) cp_async16(sdst + o + 16 * t, gsrc + o + 16 * t);
  }
  static __device__ __forceinline__ const float4* load_tw(unsigned char* smraw, const float4* __restrict__ g) {
    float4* s = reinterpret_cast<float4*>(smraw + GROUPS * GBYTES);
    for (int i = threadIdx.x; i <= N; i += TPB) s[i] = g[i];
    __syncthreads();
    return s;
  }
};
// SM ("stage the medium"): heterogeneous media append the rows of the medium maps an item needs to its stage, as in
// fft_v2.cuh.  On long lines (N >= 512) that costs more in residency than it saves in latency -- a 768-point item with medium
// rows is 18 KB, four groups per SM -- so there the rows are prefetched into L2 with the item and loaded where they are used.
template <int A, int B, int TPB, bool HOMOG, bool SM> using WStageU = WStage<A, B, TPB, (HOMOG || !SM) ? 16 : 24, 0>;
template <int A, int B, int TPB, bool HOMOG, bool SM, bool ABS, int SRC>
using WStageRho = WStage<A, B, TPB, 16, ((HOMOG || !SM) ? 0 : 16) + (ABS ? 8 : 0) + ((ABS && SRC == 1) ? 8 : 0)>;
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the B lanes of a group prefetch `bytes` (a row of N floats) into L2, one 128-byte line per lane and step
template <int B> __device__ __forceinline__ void prefetch_row(const float* row, int bytes, int t) {
  for (int o = 128 * t; o < bytes; o += 128 * B) prefetch_l2(reinterpret_cast<const char*>(row) + o);
}

template <int A, int B, int TPB, bool HOMOG, bool SM>
__global__ void __launch_bounds__(TPB) kw_x_u(StepParams P, V2Params Q) {
  using XS = WStageU<A, B, TPB, HOMOG, SM>;
  constexpr int N = A * B, G = XS::GROUPS;
  extern __shared__ __align__(16) unsigned char smraw[];
  const float4* tw = XS::load_tw(smraw, Q.tw4x);
  const int g = threadIdx.x / B, t = threadIdx.x % B;
  const bool act = (A == B) || t < A;
  char* gbase = reinterpret_cast<char*>(smraw) + g * XS::GBYTES;
  const long long hi = (long long)Q.Ry * N;                       // offset of the pair's upper row
  const int nbatch = Q.Nz * (Q.Ny / 2) / G;
  const int iters = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int pstep = gridDim.x * G;

  auto issue = [&](int lpair, int c, int stage, bool valid) {
    if (valid) {
      const int pair = phys_pair(Q, lpair);
      int z, ylo;
      wpair_rows(Q, pair, z, ylo);
      const long long r0 = ((long long)z * Q.Ny + ylo) * N;
      char* st = gbase + stage * XS::BYTES;
      XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + c * Q.ZS + (long long)pair * N), 8 * N, t);
      XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.u + c * P.RS + r0), 4 * N, t);
      XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.u + c * P.RS + r0 + hi), 4 * N, t);
      if constexpr (!HOMOG && SM) {
        XS::copy(st + 16 * N, reinterpret_cast<const char*>(P.dt_rho0_sg + c * P.RS + r0), 4 * N, t);
        XS::copy(st + 20 * N, reinterpret_cast<const char*>(P.dt_rho0_sg + c * P.RS + r0 + hi), 4 * N, t);
      }
      if constexpr (!HOMOG && !SM) {
        prefetch_row<B>(P.dt_rho0_sg + c * P.RS + r0, 4 * N, t);
        prefetch_row<B>(P.dt_rho0_sg + c * P.RS + r0 + hi, 4 * N, t);
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
    wpair_rows(Q, pair, z, ylo);
    const long long r0 = ((long long)z * Q.Ny + ylo) * N;
    const int par = it & 1;                               // item index = 3*it + c
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cp_async_wait<1>();
      __syncwarp();
      const int stage = (par + c) & 1;
      float2* zb = reinterpret_cast<float2*>(gbase + stage * XS::BYTES);
      const float* rb = reinterpret_cast<const float*>(gbase + stage * XS::BYTES + 8 * N);
      float2 v[B];
      if (act) {
#pragma unroll
        for (int kb = 0; kb < B; ++kb) v[kb] = zb[t + A * kb];
      }
      __syncwarp();
      wline_fft<A, B, true>(v, tw, zb, t);
      float* u = P.u + c * P.RS + r0;
      const float* mr = P.dt_rho0_sg + c * P.RS + r0;
      float2 s;
      if (c == 1) s = make_float2(P.sgy[ylo], P.sgy[ylo + Q.Ry]);
      else if (c == 2) s.x = s.y = P.sgz[z];
#pragma unroll
      for (int i = 0; i < A; ++i) {
        const int x = t + B * i;
        if (c == 0) s.x = s.y = P.sgx[x];
        float2 d;
        if constexpr (HOMOG) { d.x = d.y = -P.dt_rho0_sg_s; }
        else if constexpr (SM) { d.x = -rb[2 * N + x]; d.y = -rb[3 * N + x]; }
        else { d.x = -mr[x]; d.y = -mr[hi + x]; }
        const float2 un = __fmul2_rn(s, __ffma2_rn(d, v[i], __fmul2_rn(s, make_float2(rb[x], rb[N + x]))));
        u[x] = un.x;
        u[hi + x] = un.y;
        v[i] = un;
      }
      wline_fft<A, B, false>(v, tw, zb, t);
      float2* zo = Q.Z4 + c * Q.ZS + (long long)pair * N;
      if (act) {
#pragma unroll
        for (int kb = 0; kb < B; ++kb) zo[t + A * kb] = v[kb];
      }
      __syncwarp();
      if (c == 0) issue(lpair, 2, stage, true);
      else issue(lpair + pstep, c - 1, stage, it + 1 < iters);
    }
  }
  cp_async_wait<0>();
}

// SRC: 0 none, 1 filtered source in Z4[3].  ABS: absorbing medium (operands of the fractional Laplacians go to Z4[0], Z4[1],
// sum rho waits in r1; kw_x_p finishes the step).
template <int A, int B, int TPB, bool HOMOG, bool SM, int SRC, bool ABS>
__global__ void __launch_bounds__(TPB) kw_x_rho_p(StepParams P, V2Params Q) {
  using XS = WStageRho<A, B, TPB, HOMOG, SM, ABS, SRC>;
  constexpr int N = A * B, G = XS::GROUPS;
  constexpr int NI = (SRC == 1 ? 5 : 4) - (ABS ? 1 : 0);
  constexpr int C0 = SRC == 1 ? 1 : 0;
  extern __shared__ __align__(16) unsigned char smraw[];
  const float4* tw = XS::load_tw(smraw, Q.tw4x);
  const int g = threadIdx.x / B, t = threadIdx.x % B;
  const bool act = (A == B) || t < A;
  char* gbase = reinterpret_cast<char*>(smraw) + g * XS::GBYTES;
  char* mbase = gbase + 2 * XS::BYTES;
  // absorbing medium: the running sum of the velocity gradients lives in thread-private shared-memory slots (x = t + B i)
  float2* dsl = reinterpret_cast<float2*>(mbase + ((HOMOG || !SM) ? 0 : 16 * N));
  float2* srl = dsl + N;                               // ... and with it the source rows (register budget)
  constexpr bool SRC_SM = ABS && SRC == 1;
  const long long hi = (long long)Q.Ry * N;
  const int nbatch = Q.Nz * (Q.Ny / 2) / G;
  const int iters = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int pstep = gridDim.x * G;

  auto issue = [&](int lpair, int c, int stage, bool valid, int slot) {
    if (valid) {
      const int pair = phys_pair(Q, lpair);
      int z, ylo;
      wpair_rows(Q, pair, z, ylo);
      const long long r0 = ((long long)z * Q.Ny + ylo) * N;
      char* st = gbase + stage * XS::BYTES;
      if (c < 3) {
        XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + (c < 0 ? 3 : c) * Q.ZS + (long long)pair * N), 8 * N, t);
        if (c >= 0) {
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.rho + c * P.RS + r0), 4 * N, t);
          XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.rho + c * P.RS + r0 + hi), 4 * N, t);
        }
        if constexpr (!HOMOG && SM) {
          if (c == 0) {
            XS::copy(mbase + slot * 8 * N, reinterpret_cast<const char*>(P.dt_rho0 + r0), 4 * N, t);
            XS::copy(mbase + slot * 8 * N + 4 * N, reinterpret_cast<const char*>(P.dt_rho0 + r0 + hi), 4 * N, t);
          }
        }
        if constexpr (!HOMOG && !SM) {
          if (c == 0) { prefetch_row<B>(P.dt_rho0 + r0, 4 * N, t); prefetch_row<B>(P.dt_rho0 + r0 + hi, 4 * N, t); }
          if (c == 2 && !ABS) { prefetch_row<B>(P.c2 + r0, 4 * N, t); prefetch_row<B>(P.c2 + r0 + hi, 4 * N, t); }
        }
      } else {
        const bool zin = (unsigned)(z + P.z0 - P.pz) < (unsigned)P.nz;
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
  float2 src[(SRC == 1 && !SRC_SM) ? A : 1];
  float2 sum[A];
  for (int it = 0; it < iters; ++it, lpair += pstep) {
    const int pair = phys_pair(Q, lpair);
    int z, ylo;
    wpair_rows(Q, pair, z, ylo);
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
      const float* mb = SM ? reinterpret_cast<const float*>(mbase + (it & 1) * 8 * N) : P.dt_rho0 + r0;   // rows lo / hi
      const long long mhi = SM ? (long long)N : hi;
      if (c < 3) {
        float2 v[B];
        if (act) {
#pragma unroll
          for (int kb = 0; kb < B; ++kb) v[kb] = zb[t + A * kb];
        }
        __syncwarp();
        wline_fft<A, B, true>(v, tw, zb, t);
        if (c < 0) {
          if constexpr (SRC == 1) {
#pragma unroll
            for (int i = 0; i < A; ++i) {
              if constexpr (SRC_SM) srl[t + B * i] = v[i]; else src[i] = v[i];
            }
          }
        } else {
          float* rho = P.rho + c * P.RS + r0;
          float2 a;
          if (c == 1) a = make_float2(P.pmly[ylo], P.pmly[ylo + Q.Ry]);
          else if (c == 2) a.x = a.y = P.pmlz[z];
#pragma unroll
          for (int i = 0; i < A; ++i) {
            const int x = t + B * i;
            if (c == 0) a.x = a.y = P.pmlx[x];
            float2 d;
            if constexpr (HOMOG) { d.x = d.y = -P.dt_rho0_s; }
            else { d.x = -mb[x]; d.y = -mb[mhi + x]; }
            float2 rn = __fmul2_rn(a, __ffma2_rn(d, v[i], __fmul2_rn(a, make_float2(rb[x], rb[N + x]))));
            if constexpr (SRC == 1) rn = cadd(rn, SRC_SM ? srl[x] : src[SRC_SM ? 0 : i]);
            rho[x] = rn.x;
            rho[hi + x] = rn.y;
            sum[i] = c == 0 ? rn : cadd(sum[i], rn);
            if constexpr (ABS) dsl[x] = c == 0 ? v[i] : cadd(dsl[x], v[i]);
          }
          if (ABS && c == 2) {
            if constexpr (ABS) {
#pragma unroll
              for (int i = 0; i < A; ++i) {
                const int x = t + B * i;
                float2 r0v;
                if constexpr (HOMOG) { r0v.x = r0v.y = P.rho0_s; } else { r0v.x = mb[x] * P.inv_dt; r0v.y = mb[mhi + x] * P.inv_dt; }
                v[i] = __fmul2_rn(r0v, dsl[x]);
                P.r1[r0 + x] = sum[i].x;
                P.r1[r0 + hi + x] = sum[i].y;
              }
              __syncwarp();
              wline_fft<A, B, false>(v, tw, zb, t);
              float2* zo = Q.Z4 + (long long)pair * N;
              if (act) {
#pragma unroll
                for (int kb = 0; kb < B; ++kb) zo[t + A * kb] = v[kb];
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < A; ++i) v[i] = sum[i];
              wline_fft<A, B, false>(v, tw, zb, t);
              zo += Q.ZS;
              if (act) {
#pragma unroll
                for (int kb = 0; kb < B; ++kb) zo[t + A * kb] = v[kb];
              }
            }
          } else if (c == 2) {
#pragma unroll
            for (int i = 0; i < A; ++i) {
              const int x = t + B * i;
              float2 c2;
              if constexpr (HOMOG) { c2.x = c2.y = P.c2_s; } else { c2.x = P.c2[r0 + x]; c2.y = P.c2[r0 + hi + x]; }
              sum[i] = __fmul2_rn(c2, sum[i]);
              if (Q.store_p) { P.p[r0 + x] = sum[i].x; P.p[r0 + hi + x] = sum[i].y; }
            }
          }
        }
      } else {
        const bool zin = (unsigned)(z + P.z0 - P.pz) < (unsigned)P.nz;
        const bool in0 = zin && (unsigned)(ylo - P.py) < (unsigned)P.ny;
        const bool in1 = zin && (unsigned)(ylo + Q.Ry - P.py) < (unsigned)P.ny;
        float2* pmg = Q.pm + r0;
        if (in0) {
#pragma unroll
          for (int i = 0; i < A; ++i) {
            const int x = t + B * i;
            sensor_update(pmg + x, zb[x], sum[i].x, Q.pm_always);
          }
        }
        if (in1) {
#pragma unroll
          for (int i = 0; i < A; ++i) {
            const int x = t + B * i;
            sensor_update(pmg + hi + x, zb[N + x], sum[i].y, Q.pm_always);
          }
        }
        __syncwarp();
        float2 f[B];
#pragma unroll
        for (int i = 0; i < A; ++i) f[i] = sum[i];
        wline_fft<A, B, false>(f, tw, zb, t);
        float2* zo = Q.ZP + (long long)pair * N;
        if (act) {
#pragma unroll
          for (int kb = 0; kb < B; ++kb) zo[t + A * kb] = f[kb];
        }
      }
      __syncwarp();
      if (ci + 2 < NI) issue(lpair, ci + 2 - C0, stage, true, it & 1);
      else issue(lpair + pstep, ci + 2 - NI - C0, stage, it + 1 < iters, (it + 1) & 1);
    }
  }
  cp_async_wait<0>();
  if (blockIdx.x == 0 && threadIdx.x == 0) *P.step = *P.step + 1;
}

// Absorbing medium, last pass of the step: p = c0^2 (sum rho + tau L1 - eta L2), running max/min, FFT_x of p -> ZP.
template <int A, int B, int TPB, bool HOMOG, bool SM>
__global__ void __launch_bounds__(TPB) kw_x_p(StepParams P, V2Params Q, int use_tau, int use_eta) {
  using XS = WStageU<A, B, TPB, HOMOG, SM>;
  constexpr int N = A * B, G = XS::GROUPS;
  constexpr int NI = 3;
  extern __shared__ __align__(16) unsigned char smraw[];
  const float4* tw = XS::load_tw(smraw, Q.tw4x);
  const int g = threadIdx.x / B, t = threadIdx.x % B;
  const bool act = (A == B) || t < A;
  char* gbase = reinterpret_cast<char*>(smraw) + g * XS::GBYTES;
  const long long hi = (long long)Q.Ry * N;
  const int nbatch = Q.Nz * (Q.Ny / 2) / G;
  const int iters = (int)blockIdx.x < nbatch ? (nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int pstep = gridDim.x * G;
  auto issue = [&](int lpair, int c, int stage, bool valid) {
    if (valid) {
      const int pair = phys_pair(Q, lpair);
      int z, ylo;
      wpair_rows(Q, pair, z, ylo);
      const long long r0 = ((long long)z * Q.Ny + ylo) * N;
      char* st = gbase + stage * XS::BYTES;
      if (c < 2) {
        XS::copy(st, reinterpret_cast<const char*>(Q.Z4 + c * Q.ZS + (long long)pair * N), 8 * N, t);
        if (c == 0) {
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.r1 + r0), 4 * N, t);
          XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.r1 + r0 + hi), 4 * N, t);
          if constexpr (!HOMOG && SM) {
            XS::copy(st + 16 * N, reinterpret_cast<const char*>(P.tau + r0), 4 * N, t);
            XS::copy(st + 20 * N, reinterpret_cast<const char*>(P.tau + r0 + hi), 4 * N, t);
          }
          if constexpr (!HOMOG && !SM) { prefetch_row<B>(P.tau + r0, 4 * N, t); prefetch_row<B>(P.tau + r0 + hi, 4 * N, t); }
        } else if constexpr (!HOMOG) {
          XS::copy(st + 8 * N, reinterpret_cast<const char*>(P.eta + r0), 4 * N, t);
          XS::copy(st + 12 * N, reinterpret_cast<const char*>(P.eta + r0 + hi), 4 * N, t);
          if constexpr (SM) {
            XS::copy(st + 16 * N, reinterpret_cast<const char*>(P.c2 + r0), 4 * N, t);
            XS::copy(st + 20 * N, reinterpret_cast<const char*>(P.c2 + r0 + hi), 4 * N, t);
          } else {
            prefetch_row<B>(P.c2 + r0, 4 * N, t);
            prefetch_row<B>(P.c2 + r0 + hi, 4 * N, t);
          }
        }
      } else {
        const bool zin = (unsigned)(z + P.z0 - P.pz) < (unsigned)P.nz;
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
  float2 acc[B];
  for (int it = 0; it < iters; ++it, lpair += pstep) {
    const int pair = phys_pair(Q, lpair);
    int z, ylo;
    wpair_rows(Q, pair, z, ylo);
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
        float2 v[B];
        if (act) {
#pragma unroll
          for (int kb = 0; kb < B; ++kb) v[kb] = zb[t + A * kb];
        }
        __syncwarp();
        wline_fft<A, B, true>(v, tw, zb, t);
#pragma unroll
        for (int i = 0; i < A; ++i) {
          const int x = t + B * i;
          if (c == 0) {
            float2 ta;
            if constexpr (HOMOG) { ta.x = ta.y = P.tau_s; }
            else if constexpr (SM) { ta.x = rb[2 * N + x]; ta.y = rb[3 * N + x]; }
            else { ta.x = P.tau[r0 + x]; ta.y = P.tau[r0 + hi + x]; }
            const float2 s0 = make_float2(rb[x], rb[N + x]);
            acc[i] = use_tau ? __ffma2_rn(ta, v[i], s0) : s0;
          } else {
            float2 et, c2;
            if constexpr (HOMOG) { et.x = et.y = -P.eta_s; c2.x = c2.y = P.c2_s; }
            else if constexpr (SM) { et.x = -rb[x]; et.y = -rb[N + x]; c2.x = rb[2 * N + x]; c2.y = rb[3 * N + x]; }
            else { et.x = -rb[x]; et.y = -rb[N + x]; c2.x = P.c2[r0 + x]; c2.y = P.c2[r0 + hi + x]; }
            if (use_eta) acc[i] = __ffma2_rn(et, v[i], acc[i]);
            acc[i] = __fmul2_rn(c2, acc[i]);
            if (Q.store_p) { P.p[r0 + x] = acc[i].x; P.p[r0 + hi + x] = acc[i].y; }
          }
        }
      } else {
        const bool zin = (unsigned)(z + P.z0 - P.pz) < (unsigned)P.nz;
        const bool in0 = zin && (unsigned)(ylo - P.py) < (unsigned)P.ny;
        const bool in1 = zin && (unsigned)(ylo + Q.Ry - P.py) < (unsigned)P.ny;
        float2* pmg = Q.pm + r0;
        if (in0) {
#pragma unroll
          for (int i = 0; i < A; ++i) {
            const int x = t + B * i;
            sensor_update(pmg + x, zb[x], acc[i].x, Q.pm_always);
          }
        }
        if (in1) {
#pragma unroll
          for (int i = 0; i < A; ++i) {
            const int x = t + B * i;
            sensor_update(pmg + hi + x, zb[N + x], acc[i].y, Q.pm_always);
          }
        }
        __syncwarp();
        wline_fft<A, B, false>(acc, tw, zb, t);
        float2* zo = Q.ZP + (long long)pair * N;
        if (act) {
#pragma unroll
          for (int kb = 0; kb < B; ++kb) zo[t + A * kb] = acc[kb];
        }
      }
      __syncwarp();
      if (c + 2 < NI) issue(lpair, c + 2, stage, true);
      else issue(lpair + pstep, c + 2 - NI, stage, it + 1 < iters);
    }
  }
  cp_async_wait<0>();
}

// x forward of the dense source slab (row pairs).  grid = ceil(nzs*(Ny/2)/G), G = 256/B
template <int A, int B>
__global__ void __launch_bounds__(256) kw_x_src(StepParams P, V2Params Q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int G = 256 / B, N = A * B;
  float4* tws = reinterpret_cast<float4*>(smraw + G * N * 8);
  for (int i = threadIdx.x; i <= N; i += 256) tws[i] = Q.tw4x[i];
  __syncthreads();
  const int g = threadIdx.x / B, t = threadIdx.x % B;
  float2* sm = reinterpret_cast<float2*>(smraw) + g * N;
  const int hy = Q.Ny / 2;
  const long long npair = (long long)Q.nzs * hy;
  long long pair = (long long)blockIdx.x * G + g;          // zr*(Ny/2) + m
  const bool ok = pair < npair;                           // the whole warp walks through the transform (warp barriers)
  if (!ok) pair = npair - 1;
  const int zr = (int)(pair / hy), m = (int)(pair - (long long)zr * hy);
  const int ylo = ((m >> Q.ry_sh) << (Q.ry_sh + 1)) | (m & (Q.Ry - 1));
  const long long r0 = ((long long)zr * Q.Ny + ylo) * N, hi = (long long)Q.Ry * N;
  float2 v[B];
#pragma unroll
  for (int i = 0; i < A; ++i) v[i] = make_float2(Q.Sslab[r0 + t + B * i], Q.Sslab[r0 + hi + t + B * i]);
  wline_fft<A, B, false>(v, tws, sm, t);
  if (ok && (A == B || t < A)) {
#pragma unroll
    for (int kb = 0; kb < B; ++kb) Q.ZSslab[pair * N + t + A * kb] = v[kb];
  }
}

// (p_max, p_min) pairs of the expanded (local) planes -> the two sensor arrays on the inner grid (this rank's planes)
static __global__ void kw_pm_crop(StepParams P, const float2* __restrict__ pm, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % P.nx);
    const long long r = i / P.nx;
    const int y = (int)(r % P.ny), jl = (int)(r / P.ny);
    const int z = jl + P.jz0 + P.pz - P.z0;                 // local expanded plane of inner plane jz0 + jl
    const float2 v = pm[((long long)z * P.Ny + (y + P.py)) * P.Nx + (x + P.px)];
    P.pmax[i] = v.x;
    P.pmin[i] = v.y;
  }
}

}  // namespace lifu
