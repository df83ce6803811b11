// fft_gen.cuh -- pipeline v3: the fused hand-written FFT passes of fft_v2.cuh for ANY 2/3/5/7-smooth axis length.
//
// Why: kspaceFirstOrder3D's pml_auto (kwave_if.py:118) picks the PML per axis that minimises the largest prime factor of
// the expanded size, so the grids the reference produces are 81 x 81 x 125 (its default SimSetup, sim_setup.py:24-36),
// 768^3 (BASELINE config C5), 512, 96, 120 ... -- not only the 64 / 256 point axes pipeline v2 is specialised for.  This
// file keeps every such grid off the library FFT: same data flow, same fusion (spectral multipliers, velocity / density /
// pressure updates, source, sensor reduction inside the FFT passes), same Z / H layouts, with the 1-D transforms done by
// a generic Stockham autosort FFT in shared memory.
//
// Line transform.  N = r_0 r_1 ... r_{s-1}, r_i in {16, 9, 8, 7, 5, 4, 3, 2}.  A CTA holds a TILE of L lines in shared
// memory as tile[n][lane] (lane fastest, row pitch L + 1 float2) and runs s radix stages over it, ping-ponging between
// two buffers: stage i reads x[j + t N/r] (t < r), multiplies the twiddles w^(t k N/(Ns r)), k = j mod Ns, does a register
// DFT of size r (packed-FP32 arithmetic of fft_v2.cuh) and writes y[(j - k) r + k + t Ns]; Ns = r_0 .. r_{i-1}.  The work
// items of a stage are (butterfly j, lane) pairs spread over the CTA's threads with the lane fastest, so shared-memory
// accesses are conflict free for every radix and the result comes out in natural order (no bit reversal, no
// constraints between forward and inverse factor orders).
//
// Layouts (float2 = complex):
//   real field  R[z][y][x]
//   Z layout    Z[z][m][kx], m < My = ceil(Ny / 2), kx < Nx : x-spectrum of the ROW PAIR row(2m) + i row(2m+1) (a missing
//               second row of an odd Ny is zero)
//   H layout    H[z][ky][kx], kx <= Nx/2, row pitch PH        half spectrum
// Strided (y, z) passes work on tiles of L consecutive kx (global accesses of L * 8 contiguous bytes); their first radix
// stage reads the column tile straight from global memory (R independent loads per work item) and their last stage writes
// straight back (GenIO), so only the stages in between and the point-wise spectral operators touch shared memory.  x
// passes work on batches of L row pairs, moved into the same tile layout with 8-byte cp.async copies (row pitch L + 1 keeps
// the transposition conflict free); what has to outlive a transform there is parked in global scratch rows (re-read from L2).
#pragma once
#include "fft_v2.cuh"

// Launch bounds of the v3 kernels.  Default: 256 threads (512 for the x passes), registers uncapped (128 in practice,
// two CTAs per SM).  -DLIFU_G3_LB512 is the experiment build: 512-thread CTAs capped at 64 registers (radices <= 8 only,
// LIFU_V3_RMAX=8), 32 warps per SM.
#ifdef LIFU_G3_LB512
#define G3_LB_S __launch_bounds__(512, 2)
#define G3_LB_X __launch_bounds__(512, 2)
#else
#define G3_LB_S __launch_bounds__(256)
#define G3_LB_X __launch_bounds__(512)
#endif

namespace lifu {

// ------------------------------------------------------------------------------------------------
// register DFTs of the odd radices (natural order, unnormalised); 2 / 4 / 8 / 16 come from fft_v2.cuh
template <bool INV> __device__ __forceinline__ float2 rot(float2 a, float c, float s) {   // a * (c - i s) (forward) / (c + i s)
  const float si = INV ? s : -s;
  return make_float2(a.x * c - a.y * si, a.x * si + a.y * c);
}
template <bool INV> __device__ __forceinline__ void dft3(float2& a, float2& b, float2& c) {
  const float S3 = 0.86602540378443864676f;
  const float2 s = cadd(b, c), d = csub(b, c);
  const float2 m = make_float2(a.x - 0.5f * s.x, a.y - 0.5f * s.y);
  const float2 e = INV ? make_float2(-S3 * d.y, S3 * d.x) : make_float2(S3 * d.y, -S3 * d.x);   // -+ i S3 d
  a = cadd(a, s);
  b = cadd(m, e);
  c = csub(m, e);
}
template <bool INV> __device__ __forceinline__ void dft5(float2 (&x)[5]) {
  const float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;
  const float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;
  const float2 s1 = cadd(x[1], x[4]), d1 = csub(x[1], x[4]);
  const float2 s2 = cadd(x[2], x[3]), d2 = csub(x[2], x[3]);
  const float2 a1 = make_float2(x[0].x + C1 * s1.x + C2 * s2.x, x[0].y + C1 * s1.y + C2 * s2.y);
  const float2 a2 = make_float2(x[0].x + C2 * s1.x + C1 * s2.x, x[0].y + C2 * s1.y + C1 * s2.y);
  float2 b1 = make_float2(S1 * d1.x + S2 * d2.x, S1 * d1.y + S2 * d2.y);
  float2 b2 = make_float2(S2 * d1.x - S1 * d2.x, S2 * d1.y - S1 * d2.y);
  // forward: X1 = a1 - i b1, X4 = a1 + i b1, X2 = a2 - i b2, X3 = a2 + i b2
  const float2 ib1 = make_float2(-b1.y, b1.x), ib2 = make_float2(-b2.y, b2.x);
  x[0] = cadd(x[0], cadd(s1, s2));
  if (!INV) { x[1] = csub(a1, ib1); x[4] = cadd(a1, ib1); x[2] = csub(a2, ib2); x[3] = cadd(a2, ib2); }
  else      { x[1] = cadd(a1, ib1); x[4] = csub(a1, ib1); x[2] = cadd(a2, ib2); x[3] = csub(a2, ib2); }
}
template <bool INV> __device__ __forceinline__ void dft7(float2 (&x)[7]) {
  const float C1 = 0.62348980185873353053f, C2 = -0.22252093395631440429f, C3 = -0.90096886790241912624f;
  const float S1 = 0.78183148246802980871f, S2 = 0.97492791218182360702f, S3 = 0.43388373911755812048f;
  const float2 s1 = cadd(x[1], x[6]), d1 = csub(x[1], x[6]);
  const float2 s2 = cadd(x[2], x[5]), d2 = csub(x[2], x[5]);
  const float2 s3 = cadd(x[3], x[4]), d3 = csub(x[3], x[4]);
  const float2 a1 = make_float2(x[0].x + C1 * s1.x + C2 * s2.x + C3 * s3.x, x[0].y + C1 * s1.y + C2 * s2.y + C3 * s3.y);
  const float2 a2 = make_float2(x[0].x + C2 * s1.x + C3 * s2.x + C1 * s3.x, x[0].y + C2 * s1.y + C3 * s2.y + C1 * s3.y);
  const float2 a3 = make_float2(x[0].x + C3 * s1.x + C1 * s2.x + C2 * s3.x, x[0].y + C3 * s1.y + C1 * s2.y + C2 * s3.y);
  const float2 b1 = make_float2(S1 * d1.x + S2 * d2.x + S3 * d3.x, S1 * d1.y + S2 * d2.y + S3 * d3.y);
  const float2 b2 = make_float2(S2 * d1.x - S3 * d2.x - S1 * d3.x, S2 * d1.y - S3 * d2.y - S1 * d3.y);
  const float2 b3 = make_float2(S3 * d1.x - S1 * d2.x + S2 * d3.x, S3 * d1.y - S1 * d2.y + S2 * d3.y);
  const float2 i1 = make_float2(-b1.y, b1.x), i2 = make_float2(-b2.y, b2.x), i3 = make_float2(-b3.y, b3.x);
  x[0] = cadd(x[0], cadd(s1, cadd(s2, s3)));
  if (!INV) { x[1] = csub(a1, i1); x[6] = cadd(a1, i1); x[2] = csub(a2, i2); x[5] = cadd(a2, i2); x[3] = csub(a3, i3); x[4] = cadd(a3, i3); }
  else      { x[1] = cadd(a1, i1); x[6] = csub(a1, i1); x[2] = cadd(a2, i2); x[5] = csub(a2, i2); x[3] = cadd(a3, i3); x[4] = csub(a3, i3); }
}
template <bool INV> __device__ __forceinline__ void dft9(float2 (&x)[9]) {
  // n = b + 3 a, k = ka + 3 kb: three DFT-3 over a, twiddles w9^(b ka), three DFT-3 over b
  const float C1 = 0.76604444311897803520f, S1 = 0.64278760968653932632f;    // w9^1
  const float C2 = 0.17364817766693034885f, S2 = 0.98480775301220805937f;    // w9^2
  const float C4 = -0.93969262078590838405f, S4 = 0.34202014332566873304f;   // w9^4
  float2 y[3][3];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    float2 p = x[b], q = x[b + 3], r = x[b + 6];
    dft3<INV>(p, q, r);
    y[b][0] = p; y[b][1] = q; y[b][2] = r;
  }
  y[1][1] = rot<INV>(y[1][1], C1, S1);
  y[1][2] = rot<INV>(y[1][2], C2, S2);
  y[2][1] = rot<INV>(y[2][1], C2, S2);
  y[2][2] = rot<INV>(y[2][2], C4, S4);
#pragma unroll
  for (int ka = 0; ka < 3; ++ka) {
    float2 p = y[0][ka], q = y[1][ka], r = y[2][ka];
    dft3<INV>(p, q, r);
    x[ka] = p; x[ka + 3] = q; x[ka + 6] = r;
  }
}
template <int R, bool INV> __device__ __forceinline__ void dft_any(float2 (&v)[R]) {
  if constexpr (R == 3) dft3<INV>(v[0], v[1], v[2]);
  else if constexpr (R == 5) dft5<INV>(v);
  else if constexpr (R == 7) dft7<INV>(v);
  else if constexpr (R == 9) dft9<INV>(v);
  else dft<R, INV>(v);
}

// ------------------------------------------------------------------------------------------------
// Optional global-memory ends of a line transform: the first stage can read its inputs straight from a strided global
// column tile (R independent loads per work item instead of a staging pass through shared memory) and the last stage can
// write its outputs straight back; element (n, lane) lives at g[n * stride + lane], lanes >= nvalid are masked, and an
// optional per-line multiplier table rides on either end.
struct GenIO {
  const float2* gin = nullptr;
  long long in_stride = 0;
  const float2* in_mul = nullptr;
  float2* gout = nullptr;
  long long out_stride = 0;
  const float2* out_mul = nullptr;
  int nvalid = 0;
};

// One radix stage over a tile: src / dst are tile[n * LP + lane].
template <int R, bool INV, bool GIN, bool GOUT>
__device__ __forceinline__ void gen_stage(int N, int Ns, int L, int lsh, int LP, const float2* __restrict__ tw,
                                          const float2* __restrict__ src, float2* __restrict__ dst, const GenIO& io) {
  const int nb = N / R;
  const int step = N / (Ns * R);
  const int items = nb << lsh;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int lane = w & (L - 1), j = w >> lsh;
    const int k = Ns == 1 ? 0 : j % Ns;
    float2 v[R];
    if (GIN) {
      const bool ok = lane < io.nvalid;
      const float2* gp = io.gin + (long long)j * io.in_stride + lane;
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = ok ? gp[(long long)t * nb * io.in_stride] : make_float2(0.f, 0.f);
      if (io.in_mul) {
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = cmul2(v[t], io.in_mul[j + t * nb]);
      }
    } else {
      const float2* sp = src + j * LP + lane;
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = sp[t * nb * LP];
    }
    if (Ns > 1) {
      const int ks = k * step;                 // t * ks < N for t < R: no wrap-around
#pragma unroll
      for (int t = 1; t < R; ++t) {
        float2 wv = __ldg(tw + t * ks);
        if (INV) wv.y = -wv.y;
        v[t] = cmul2(v[t], wv);
      }
    }
    dft_any<R, INV>(v);
    const int o0 = (j - k) * R + k;
    if (GOUT) {
      if (lane < io.nvalid) {
        float2* gp = io.gout + (long long)o0 * io.out_stride + lane;
#pragma unroll
        for (int t = 0; t < R; ++t) {
          float2 o = v[t];
          if (io.out_mul) o = cmul2(o, io.out_mul[o0 + t * Ns]);
          gp[(long long)t * Ns * io.out_stride] = o;
        }
      }
    } else {
      float2* dp = dst + o0 * LP + lane;
#pragma unroll
      for (int t = 0; t < R; ++t) dp[t * Ns * LP] = v[t];
    }
  }
}

template <bool INV, bool GIN, bool GOUT>
__device__ __forceinline__ void gen_stage_any(int r, int N, int Ns, int L, int lsh, int LP, const float2* tw, const float2* in,
                                              float2* out, const GenIO& io) {
  switch (r) {
    case 16: gen_stage<16, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    case 9: gen_stage<9, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    case 8: gen_stage<8, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    case 7: gen_stage<7, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    case 5: gen_stage<5, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    case 4: gen_stage<4, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    case 3: gen_stage<3, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
    default: gen_stage<2, INV, GIN, GOUT>(N, Ns, L, lsh, LP, tw, in, out, io); break;
  }
}

// The whole line transform of a tile.  First stage reads `src` (or io.gin when GIN), stages alternate between b0 and b1
// starting with b0; src may be b1 (it is overwritten from the second stage on) but not b0.  Returns the buffer holding
// the result (nullptr when GOUT: the last stage wrote to io.gout).  Every thread of the CTA must call it; the caller
// synchronises before (src complete) -- a barrier follows every stage that wrote shared memory.
template <bool INV, bool GIN, bool GOUT>
__device__ __noinline__ float2* gen_fft_io(const GenPlan& pl, int L, int lsh, const float2* src, float2* b0, float2* b1,
                                           const GenIO& io) {
  const int LP = L + 1;
  int Ns = 1;
  const float2* in = src;
  float2* out = b0;
  for (int s = 0; s < pl.ns; ++s) {
    const int r = pl.radix[s];
    const bool gi = GIN && s == 0, go = GOUT && s == pl.ns - 1;
    bool done = false;
    if constexpr (GIN && GOUT) { if (gi && go) { gen_stage_any<INV, true, true>(r, pl.N, Ns, L, lsh, LP, pl.tw, in, out, io); done = true; } }
    if constexpr (GIN) { if (!done && gi) { gen_stage_any<INV, true, false>(r, pl.N, Ns, L, lsh, LP, pl.tw, in, out, io); done = true; } }
    if constexpr (GOUT) { if (!done && go) { gen_stage_any<INV, false, true>(r, pl.N, Ns, L, lsh, LP, pl.tw, in, out, io); done = true; } }
    if (!done) gen_stage_any<INV, false, false>(r, pl.N, Ns, L, lsh, LP, pl.tw, in, out, io);
    if (!go) __syncthreads();
    Ns *= r;
    in = out;
    out = (out == b0) ? b1 : b0;
  }
  return GOUT ? nullptr : const_cast<float2*>(in);
}
template <bool INV>
__device__ __forceinline__ float2* gen_fft(const GenPlan& pl, int L, int lsh, const float2* src, float2* b0, float2* b1) {
  GenIO io;
  return gen_fft_io<INV, false, false>(pl, L, lsh, src, b0, b1, io);
}

__device__ __forceinline__ float kappa_rt(int poly, float a2) {
  return poly == 2 ? sinc_sqrt_poly8(a2) : (poly == 1 ? sinc_sqrt_poly(a2) : kappa_of(a2));
}
__device__ __forceinline__ float cosk_rt(int poly, float a2) {
  return poly == 2 ? cos_sqrt_poly8(a2) : (poly == 1 ? cos_sqrt_poly(a2) : cosf(sqrtf(a2)));
}

// shared-memory carve-up: tile buffers of N * (L + 1) float2
__device__ __forceinline__ float2* gen_buf(unsigned char* smraw, int N, int L, int i) {
  return reinterpret_cast<float2*>(smraw) + (size_t)i * N * (L + 1);
}

// ------------------------------------------------------------------------------------------------
// merge the rows of the pairs of a y-transformed tile (tile[y][lane], lane = kx - kx0) into the Z layout:
// Z[kx] = A + iB, Z[Nx-kx] = conj(A - iB); bins 0 and Nx/2 keep the real parts (what a C2R transform does).
__device__ __forceinline__ void g3_merge_store(const GParams& G, const float2* __restrict__ cur, float2* __restrict__ zplane,
                                               int kx0, bool mulx, const float2* __restrict__ dpx) {
  const int L = G.Ls, LP = L + 1;
  const int items = G.My << G.lsh_s;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int lane = w & (L - 1), m = w >> G.lsh_s;
    const int kx = kx0 + lane;
    if (kx >= G.Nxh) continue;
    float2 A = cur[(2 * m) * LP + lane];
    float2 B = (2 * m + 1 < G.Ny) ? cur[(2 * m + 1) * LP + lane] : make_float2(0.f, 0.f);
    if (mulx) { const float2 mx = dpx[kx]; A = cmul2(A, mx); B = cmul2(B, mx); }
    const bool selfm = (kx == 0) || (2 * kx == G.Nx);
    float2* zp = zplane + (long long)m * G.Nx;
    zp[kx] = selfm ? make_float2(A.x, B.x) : cadd_i(A, B);
    if (!selfm) zp[G.Nx - kx] = cconj(csub_i(A, B));
  }
}

// load a tile of the H layout (all ky of one plane, or all kz of one ky row) with an optional per-line multiplier
__device__ __forceinline__ void g3_load_h(const GParams& G, const float2* __restrict__ base, long long line_stride, int nlines,
                                          int kx0, float2* __restrict__ dst, const float2* __restrict__ mul) {
  const int L = G.Ls, LP = L + 1;
  const int items = nlines << G.lsh_s;
#pragma unroll 4
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int lane = w & (L - 1), n = w >> G.lsh_s;
    const int kx = kx0 + lane;
    float2 v = make_float2(0.f, 0.f);
    if (kx < G.Nxh) v = base[(long long)n * line_stride + kx];
    if (mul) v = cmul2(v, mul[n]);
    dst[n * LP + lane] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// y forward: packed row pairs -> half spectrum.  grid (tiles, planes, ncomp)
// MODE 0 pressure; 1 velocity (comp 0: x multiplier i kx e^{-i kx dx/2}, comp 1: y multiplier); 2 source slab; 3 absorption operands
template <int MODE>
__global__ void G3_LB_S g3_y_fwd(StepParams P, GParams G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = G.Ls, LP = L + 1;
  const int kx0 = blockIdx.x * L, z = blockIdx.y, comp = blockIdx.z;
  float2* b0 = gen_buf(smraw, G.Ny, L, 0);
  float2* b1 = gen_buf(smraw, G.Ny, L, 1);
  const float2* Zin = MODE == 0 ? G.ZP : ((MODE == 1 || MODE == 3) ? G.Z4 + comp * G.ZS : G.ZSslab);
  float2* Hout = MODE == 2 ? G.HSslab : G.H4 + comp * G.HS;
  const float2* zp = Zin + (long long)z * G.My * G.Nx;
  const int items = G.My << G.lsh_s;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int lane = w & (L - 1), m = w >> G.lsh_s;
    const int kx = kx0 + lane;
    float2 a2 = make_float2(0.f, 0.f), b2 = a2;
    if (kx < G.Nxh) {
      const int km = kx == 0 ? 0 : G.Nx - kx;
      const float2 d = zp[(long long)m * G.Nx + kx], mm = zp[(long long)m * G.Nx + km];
      a2 = cadd_conj(d, mm);                   // 2A = Z[k] + conj Z[-k]
      b2 = cmul_mi(csub_conj(d, mm));          // 2B = -i (Z[k] - conj Z[-k])
      if (MODE == 1 && comp == 0) { const float2 mx = P.dnx[kx]; a2 = cmul2(a2, mx); b2 = cmul2(b2, mx); }
    }
    b1[(2 * m) * LP + lane] = a2;
    if (2 * m + 1 < G.Ny) b1[(2 * m + 1) * LP + lane] = b2;
  }
  __syncthreads();
  GenIO io;
  io.gout = Hout + (long long)z * G.Ny * G.PH + kx0;
  io.out_stride = G.PH;
  io.out_mul = (MODE == 1 && comp == 1) ? P.dny : nullptr;
  io.nvalid = min(L, G.Nxh - kx0);
  gen_fft_io<false, false, true>(G.py, L, G.lsh_s, b1, b0, b1, io);
}

// z pass of the pressure gradient, out of place: H4[0] -> H4[2] = IFFT_z[kappa FFT_z p^] (blockIdx.z = 0) and
// H4[1] = IFFT_z[i kz e^{+i kz dz/2} kappa FFT_z p^] (blockIdx.z = 1).  grid (tiles, Ny, 2): the two chains run in
// different CTAs (the second read of the column tile hits L2), two tile buffers each.
__global__ void G3_LB_S g3_z_grad(StepParams P, GParams G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = G.Ls, LP = L + 1;
  const int kx0 = blockIdx.x * L, ky = blockIdx.y, pass = blockIdx.z;
  float2* b0 = gen_buf(smraw, G.Nz, L, 0);
  float2* b1 = gen_buf(smraw, G.Nz, L, 1);
  const long long zs = (long long)G.Ny * G.PH;
  const float2* col = G.H4 + (long long)ky * G.PH;
  const int items = G.Nz << G.lsh_s;
  GenIO io;
  io.gin = col + kx0;
  io.in_stride = zs;
  io.gout = G.H4 + (pass == 1 ? G.HS : 2 * G.HS) + (long long)ky * G.PH + kx0;
  io.out_stride = zs;
  io.nvalid = min(L, G.Nxh - kx0);
  float2* cur = gen_fft_io<false, true, false>(G.pz, L, G.lsh_s, b1, b0, b1, io);
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    const int lane = w & (L - 1), kz = w >> G.lsh_s;
    const int kx = min(kx0 + lane, G.Nxh - 1);
    const float kap = kappa_rt(P.poly_ok, P.ax2[kx] + P.ay2[ky] + P.az2[kz]) * G.norm;
    float2 v = cscale(cur[kz * LP + lane], kap);
    if (pass == 1) v = cmul2(v, P.dpz[kz]);
    cur[kz * LP + lane] = v;
  }
  __syncthreads();
  float2* other = cur == b0 ? b1 : b0;
  gen_fft_io<true, false, true>(G.pz, L, G.lsh_s, cur, other, cur, io);
}

// y inverse of the three gradient components + row-pair merge.  grid (tiles, Nz, 3), component = blockIdx.z
//   Z4[0] <- i kx e^{+i kx dx/2} IFFT_y[H4[2]];  Z4[1] <- IFFT_y[i ky e^{+i ky dy/2} H4[2]];  Z4[2] <- IFFT_y[H4[1]]
__global__ void G3_LB_S g3_y_inv_grad(StepParams P, GParams G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = G.Ls;
  const int kx0 = blockIdx.x * L, z = blockIdx.y, c = blockIdx.z;
  float2* b0 = gen_buf(smraw, G.Ny, L, 0);
  float2* b1 = gen_buf(smraw, G.Ny, L, 1);
  const float2* hp = G.H4 + (long long)z * G.Ny * G.PH;
  float2* zp = G.Z4 + (long long)z * G.My * G.Nx;
  GenIO io;
  io.gin = hp + (c == 2 ? G.HS : 2 * G.HS) + kx0;
  io.in_stride = G.PH;
  io.in_mul = c == 1 ? P.dpy : nullptr;
  io.nvalid = min(L, G.Nxh - kx0);
  const float2* cur = gen_fft_io<true, true, false>(G.py, L, G.lsh_s, b1, b0, b1, io);
  g3_merge_store(G, cur, zp + c * G.ZS, kx0, c == 0, P.dpx);
}

// y inverse + row-pair merge of H4[comp] -> Z4[comp].  grid (tiles, Nz, ncomp)
__global__ void G3_LB_S g3_y_inv(StepParams P, GParams G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = G.Ls;
  const int kx0 = blockIdx.x * L, z = blockIdx.y, comp = blockIdx.z;
  float2* b0 = gen_buf(smraw, G.Ny, L, 0);
  float2* b1 = gen_buf(smraw, G.Ny, L, 1);
  GenIO io;
  io.gin = G.H4 + comp * G.HS + (long long)z * G.Ny * G.PH + kx0;
  io.in_stride = G.PH;
  io.nvalid = min(L, G.Nxh - kx0);
  const float2* cur = gen_fft_io<true, true, false>(G.py, L, G.lsh_s, b1, b0, b1, io);
  g3_merge_store(G, cur, G.Z4 + comp * G.ZS + (long long)z * G.My * G.Nx, kx0, false, nullptr);
}

// z pass, in place, of (OP 0) the velocity divergence comps 0..2 [+ comp 3 = source field read from its slab, cos filter]
// and (OP 1) the two absorption operands (fractional Laplacians k^(y-2), k^(y-1)).  grid (tiles, Ny, ncomp)
template <int OP>
__global__ void G3_LB_S g3_z_pass(StepParams P, GParams G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int L = G.Ls, LP = L + 1;
  const int kx0 = blockIdx.x * L, ky = blockIdx.y;
  float2* b0 = gen_buf(smraw, G.Nz, L, 0);
  float2* b1 = gen_buf(smraw, G.Nz, L, 1);
  const long long zs = (long long)G.Ny * G.PH;
  const int items = G.Nz << G.lsh_s;
  const int comp = blockIdx.z;
  {
    float2* col = G.H4 + comp * G.HS + (long long)ky * G.PH;
    if (OP == 0 && comp == 3) {
      const float2* sp = G.HSslab + (long long)ky * G.PH;
      for (int w = threadIdx.x; w < items; w += blockDim.x) {
        const int lane = w & (L - 1), zz = w >> G.lsh_s;
        const int kx = kx0 + lane, zr = zz - G.z0s;
        float2 v = make_float2(0.f, 0.f);
        if (kx < G.Nxh && zr >= 0 && zr < G.nzs) v = sp[(long long)zr * zs + kx];
        b1[zz * LP + lane] = v;
      }
    }
    GenIO io;
    io.gin = col + kx0;
    io.in_stride = zs;
    io.gout = col + kx0;
    io.out_stride = zs;
    io.nvalid = min(L, G.Nxh - kx0);
    float2* cur;
    if (OP == 0 && comp == 3) {
      __syncthreads();
      cur = gen_fft<false>(G.pz, L, G.lsh_s, b1, b0, b1);
    } else {
      cur = gen_fft_io<false, true, false>(G.pz, L, G.lsh_s, b1, b0, b1, io);
    }
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
      const int lane = w & (L - 1), kz = w >> G.lsh_s;
      const int kx = min(kx0 + lane, G.Nxh - 1);
      float2 v = cur[kz * LP + lane];
      if (OP == 0) {
        const float a2 = P.ax2[kx] + P.ay2[ky] + P.az2[kz];
        if (comp == 3) v = cscale(v, cosk_rt(P.poly_ok, a2) * G.norm);
        else {
          v = cscale(v, kappa_rt(P.poly_ok, a2) * G.norm);
          if (comp == 2) v = cmul2(v, P.dnz[kz]);
        }
      } else {
        const float k2 = P.kx2[kx] + P.ky2[ky] + P.kz2[kz];
        const float e = comp == 0 ? P.y_minus2_half : P.y_minus1_half;
        v = cscale(v, k2 > 0.f ? __powf(k2, e) * G.norm : 0.f);
      }
      cur[kz * LP + lane] = v;
    }
    __syncthreads();
    float2* other = cur == b0 ? b1 : b0;
    gen_fft_io<true, false, true>(G.pz, L, G.lsh_s, cur, other, cur, io);
  }
}

// ------------------------------------------------------------------------------------------------
// x passes: a CTA owns a batch of L consecutive row pairs q = z * My + m (rows 2m, 2m+1 of plane z); L (a power of two)
// is a launch parameter chosen per kernel so that two tile buffers of N * (L + 1) float2 leave room for >= 2 CTAs per SM.
// Lines are moved between global memory (x contiguous) and the tile (lane fastest) by warps walking along x.
// Everything that has to outlive a transform (the source field, the du components, the partial pressure sum) is parked in
// global scratch rows owned by the same thread (P.Sf, P.r3, P.r1): they are re-read from L2 a few microseconds later,
// which is cheaper than a third, fourth and fifth tile buffer in shared memory (one CTA per SM).
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
// Lines of a pair batch -> tile (transposing).  cp.async copies of 8 bytes: every thread has its whole share of the tile
// in flight at once (no registers involved), which is what the memory system needs to stream -- the plain load / store
// loop it replaces kept 4 loads per thread in flight.  The caller's __syncthreads() follows.
__device__ __forceinline__ void g3_x_load(const GParams& G, int L, const float2* __restrict__ zfield, int q0, int nq,
                                          float2* __restrict__ dst) {
  const int LP = L + 1;
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int lane = wid; lane < L; lane += nw) {
    const float2* zp = zfield + (long long)(q0 + lane) * G.Nx;
    if (lane < nq) {
      for (int x = ln; x < G.Nx; x += 32) cp_async8(dst + x * LP + lane, zp + x);
    } else {
      for (int x = ln; x < G.Nx; x += 32) dst[x * LP + lane] = make_float2(0.f, 0.f);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
}
__device__ __forceinline__ void g3_x_store(const GParams& G, int L, float2* __restrict__ zfield, int q0, int nq,
                                           const float2* __restrict__ src) {
  const int LP = L + 1;
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int lane = wid; lane < nq; lane += nw) {
    float2* zp = zfield + (long long)(q0 + lane) * G.Nx;
#pragma unroll 8
    for (int x = ln; x < G.Nx; x += 32) zp[x] = src[x * LP + lane];
  }
}
// pair q -> plane z, lower row ylo, whether the upper row exists, offset of the lower row in a real field
__device__ __forceinline__ void g3_pair(const GParams& G, int q, int& z, int& ylo, bool& has_hi, long long& r0) {
  z = q / G.My;
  const int m = q - z * G.My;
  ylo = 2 * m;
  has_hi = ylo + 1 < G.Ny;
  r0 = ((long long)z * G.Ny + ylo) * G.Nx;
}

// IFFT_x of one gradient component (blockIdx.y), u = pml_sg (pml_sg u - dt/rho0_sg dp), FFT_x of the new u.
// grid (pair batches, 3)
template <bool HOMOG>
__global__ void G3_LB_X g3_x_u(StepParams P, GParams G, int L, int lsh) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int LP = L + 1, N = G.Nx;
  float2* b0 = gen_buf(smraw, N, L, 0);
  float2* b1 = gen_buf(smraw, N, L, 1);
  const int q0 = blockIdx.x * L;
  const int nq = min(L, G.Nz * G.My - q0);
  const int c = blockIdx.y;
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  g3_x_load(G, L, G.Z4 + c * G.ZS, q0, nq, b1);
  __syncthreads();
  float2* cur = gen_fft<true>(G.px, L, lsh, b1, b0, b1);
  for (int lane = wid; lane < nq; lane += nw) {
    int z, ylo; bool hh; long long r0;
    g3_pair(G, q0 + lane, z, ylo, hh, r0);
    float* u = P.u + c * P.RS + r0;
    const float* mrow = HOMOG ? nullptr : P.dt_rho0_sg + c * P.RS + r0;
    float2 s = make_float2(1.f, 1.f);
    if (c == 1) s = make_float2(P.sgy[ylo], hh ? P.sgy[ylo + 1] : 0.f);
    else if (c == 2) s.x = s.y = P.sgz[z];
#pragma unroll 4
    for (int x = ln; x < N; x += 32) {
      if (c == 0) s.x = s.y = P.sgx[x];
      float2 d;
      if (HOMOG) d.x = d.y = -P.dt_rho0_sg_s;
      else { d.x = -mrow[x]; d.y = hh ? -mrow[N + x] : 0.f; }
      const float2 g = cur[x * LP + lane];
      const float2 uo = make_float2(u[x], hh ? u[N + x] : 0.f);
      float2 un = __fmul2_rn(s, __ffma2_rn(d, g, __fmul2_rn(s, uo)));
      if (!hh) un.y = 0.f;
      u[x] = un.x;
      if (hh) u[N + x] = un.y;
      cur[x * LP + lane] = un;
    }
  }
  __syncthreads();
  float2* other = cur == b0 ? b1 : b0;
  const float2* res = gen_fft<false>(G.px, L, lsh, cur, other, cur);
  g3_x_store(G, L, G.Z4 + c * G.ZS, q0, nq, res);
}

// sensor rows of a pair batch: running max / min of p (kept as a pair in `acc`) on the rows inside the inner grid
__device__ __forceinline__ void g3_sensor(const StepParams& P, const GParams& G, int L, int q0, int nq,
                                          const float2* __restrict__ acc) {
  const int LP = L + 1, N = G.Nx;
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int lane = wid; lane < nq; lane += nw) {
    int z, ylo; bool hh; long long r0;
    g3_pair(G, q0 + lane, z, ylo, hh, r0);
    const bool zin = (unsigned)(z - P.pz) < (unsigned)P.nz;
    const bool in0 = zin && (unsigned)(ylo - P.py) < (unsigned)P.ny;
    const bool in1 = zin && hh && (unsigned)(ylo + 1 - P.py) < (unsigned)P.ny;
    if (!in0 && !in1) continue;
    float2* pmg = G.pm + r0;
    for (int x = P.px + ln; x < P.px + P.nx; x += 32) {
      const float2 p = acc[x * LP + lane];
      if (in0) sensor_update(pmg + x, pmg[x], p.x, G.pm_always);
      if (in1) sensor_update(pmg + N + x, pmg[N + x], p.y, G.pm_always);
    }
  }
}

// SRC: 0 none, 1 filtered source spectrum in Z4[3], 2 unfiltered dense slab.  ABS: absorbing medium (see fft_v2.cuh).
// Two tile buffers; the source rows live in P.Sf, the du rows (ABS) in P.r3, the density rows are re-read for the sum.
template <bool HOMOG, int SRC, bool ABS>
__global__ void G3_LB_X g3_x_rho_p(StepParams P, GParams G, int L, int lsh) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int LP = L + 1, N = G.Nx;
  float2* b0 = gen_buf(smraw, N, L, 0);
  float2* b1 = gen_buf(smraw, N, L, 1);
  const int q0 = blockIdx.x * L;
  const int nq = min(L, G.Nz * G.My - q0);
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (SRC == 1) {
    g3_x_load(G, L, G.Z4 + 3 * G.ZS, q0, nq, b1);
    __syncthreads();
    const float2* cur = gen_fft<true>(G.px, L, lsh, b1, b0, b1);
    for (int lane = wid; lane < nq; lane += nw) {
      int z, ylo; bool hh; long long r0;
      g3_pair(G, q0 + lane, z, ylo, hh, r0);
      for (int x = ln; x < N; x += 32) {
        const float2 v = cur[x * LP + lane];
        P.Sf[r0 + x] = v.x;
        if (hh) P.Sf[r0 + N + x] = v.y;
      }
    }
    __syncthreads();
  }
  for (int c = 0; c < 3; ++c) {
    g3_x_load(G, L, G.Z4 + c * G.ZS, q0, nq, b1);
    __syncthreads();
    const float2* cur = gen_fft<true>(G.px, L, lsh, b1, b0, b1);
    for (int lane = wid; lane < nq; lane += nw) {
      int z, ylo; bool hh; long long r0;
      g3_pair(G, q0 + lane, z, ylo, hh, r0);
      float* rho = P.rho + c * P.RS + r0;
      const float* mrow = HOMOG ? nullptr : P.dt_rho0 + r0;
      float2 a = make_float2(1.f, 1.f);
      if (c == 1) a = make_float2(P.pmly[ylo], hh ? P.pmly[ylo + 1] : 0.f);
      else if (c == 2) a.x = a.y = P.pmlz[z];
      const int zr = z - G.z0s;
      const bool sin_ = SRC == 2 && zr >= 0 && zr < G.nzs;
      const float* srow = SRC == 2 ? G.Sslab + ((long long)zr * G.Ny + ylo) * N : nullptr;
#pragma unroll 4
      for (int x = ln; x < N; x += 32) {
        if (c == 0) a.x = a.y = P.pmlx[x];
        float2 d;
        if (HOMOG) d.x = d.y = -P.dt_rho0_s;
        else { d.x = -mrow[x]; d.y = hh ? -mrow[N + x] : 0.f; }
        const float2 g = cur[x * LP + lane];
        const float2 ro = make_float2(rho[x], hh ? rho[N + x] : 0.f);
        float2 rn = __fmul2_rn(a, __ffma2_rn(d, g, __fmul2_rn(a, ro)));
        if (SRC == 1) rn = cadd(rn, make_float2(P.Sf[r0 + x], hh ? P.Sf[r0 + N + x] : 0.f));
        if (SRC == 2 && sin_) rn = cadd(rn, make_float2(srow[x], hh ? srow[N + x] : 0.f));
        rho[x] = rn.x;
        if (hh) rho[N + x] = rn.y;
        if (ABS) { P.r3[c * P.RS + r0 + x] = g.x; if (hh) P.r3[c * P.RS + r0 + N + x] = g.y; }
      }
    }
    __syncthreads();
  }
  // epilogue: every thread re-reads the rows it wrote itself
  for (int pass = ABS ? 0 : 1; pass < 2; ++pass) {
    // pass 0 (ABS only): rho0 * ((dux + duy) + duz) -> Z4[0];  pass 1: sum rho = (rho_x + rho_y) + rho_z -> p or Z4[1]
    for (int lane = wid; lane < L; lane += nw) {
      if (lane >= nq) { for (int x = ln; x < N; x += 32) b1[x * LP + lane] = make_float2(0.f, 0.f); continue; }
      int z, ylo; bool hh; long long r0;
      g3_pair(G, q0 + lane, z, ylo, hh, r0);
      const float* f0 = (pass == 0 ? P.r3 : P.rho) + r0;
      const float* mrow = HOMOG ? nullptr : P.dt_rho0 + r0;
#pragma unroll 4
      for (int x = ln; x < N; x += 32) {
        float2 sm = cadd(cadd(make_float2(f0[x], hh ? f0[N + x] : 0.f),
                              make_float2(f0[P.RS + x], hh ? f0[P.RS + N + x] : 0.f)),
                         make_float2(f0[2 * P.RS + x], hh ? f0[2 * P.RS + N + x] : 0.f));
        if (pass == 0) {
          float2 r0v;
          if (HOMOG) r0v.x = r0v.y = P.rho0_s;
          else { r0v.x = mrow[x] * P.inv_dt; r0v.y = hh ? mrow[N + x] * P.inv_dt : 0.f; }
          sm = __fmul2_rn(r0v, sm);
        } else if (ABS) {
          P.r1[r0 + x] = sm.x;
          if (hh) P.r1[r0 + N + x] = sm.y;
        } else {
          float2 c2;
          if (HOMOG) c2.x = c2.y = P.c2_s;
          else { c2.x = P.c2[r0 + x]; c2.y = hh ? P.c2[r0 + N + x] : 0.f; }
          sm = __fmul2_rn(c2, sm);
          if (G.store_p) { P.p[r0 + x] = sm.x; if (hh) P.p[r0 + N + x] = sm.y; }
        }
        b1[x * LP + lane] = sm;
      }
    }
    __syncthreads();
    if (!ABS) g3_sensor(P, G, L, q0, nq, b1);
    const float2* r = gen_fft<false>(G.px, L, lsh, b1, b0, b1);
    g3_x_store(G, L, ABS ? G.Z4 + pass * G.ZS : G.ZP, q0, nq, r);
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *P.step = *P.step + 1;
}

// Absorbing medium, last pass: p = c0^2 (sum rho + tau L1 - eta L2), sensor, FFT_x of p -> ZP.  Two tile buffers; the
// partial sum (sum rho + tau L1) waits in P.r3[0] while the second operand is transformed.
template <bool HOMOG>
__global__ void G3_LB_X g3_x_p(StepParams P, GParams G, int L, int lsh, int use_tau, int use_eta) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int LP = L + 1, N = G.Nx;
  float2* b0 = gen_buf(smraw, N, L, 0);
  float2* b1 = gen_buf(smraw, N, L, 1);
  const int q0 = blockIdx.x * L;
  const int nq = min(L, G.Nz * G.My - q0);
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  float2* cur = nullptr;
  for (int c = 0; c < 2; ++c) {
    g3_x_load(G, L, G.Z4 + c * G.ZS, q0, nq, b1);
    __syncthreads();
    cur = gen_fft<true>(G.px, L, lsh, b1, b0, b1);
    for (int lane = wid; lane < L; lane += nw) {
      if (lane >= nq) { if (c == 1) for (int x = ln; x < N; x += 32) cur[x * LP + lane] = make_float2(0.f, 0.f); continue; }
      int z, ylo; bool hh; long long r0;
      g3_pair(G, q0 + lane, z, ylo, hh, r0);
#pragma unroll 4
      for (int x = ln; x < N; x += 32) {
        const float2 v = cur[x * LP + lane];
        if (c == 0) {
          float2 ta;
          if (HOMOG) ta.x = ta.y = P.tau_s;
          else { ta.x = P.tau[r0 + x]; ta.y = hh ? P.tau[r0 + N + x] : 0.f; }
          const float2 s0 = make_float2(P.r1[r0 + x], hh ? P.r1[r0 + N + x] : 0.f);
          const float2 acc = use_tau ? __ffma2_rn(ta, v, s0) : s0;
          P.r3[r0 + x] = acc.x;
          if (hh) P.r3[r0 + N + x] = acc.y;
        } else {
          float2 et, c2;
          if (HOMOG) { et.x = et.y = -P.eta_s; c2.x = c2.y = P.c2_s; }
          else { et.x = -P.eta[r0 + x]; et.y = hh ? -P.eta[r0 + N + x] : 0.f; c2.x = P.c2[r0 + x]; c2.y = hh ? P.c2[r0 + N + x] : 0.f; }
          float2 acc = make_float2(P.r3[r0 + x], hh ? P.r3[r0 + N + x] : 0.f);
          if (use_eta) acc = __ffma2_rn(et, v, acc);
          acc = __fmul2_rn(c2, acc);
          cur[x * LP + lane] = acc;
          if (G.store_p) { P.p[r0 + x] = acc.x; if (hh) P.p[r0 + N + x] = acc.y; }
        }
      }
    }
    __syncthreads();
  }
  g3_sensor(P, G, L, q0, nq, cur);
  float2* other = cur == b0 ? b1 : b0;
  const float2* r = gen_fft<false>(G.px, L, lsh, cur, other, cur);
  g3_x_store(G, L, G.ZP, q0, nq, r);
}

// x forward of the dense source slab (row pairs of the slab planes).  grid = pair batches of the slab
__global__ void G3_LB_X g3_x_src(StepParams P, GParams G, int L, int lsh) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int LP = L + 1, N = G.Nx;
  float2* b0 = gen_buf(smraw, N, L, 0);
  float2* b1 = gen_buf(smraw, N, L, 1);
  const int q0 = blockIdx.x * L;
  const int nq = min(L, G.nzs * G.My - q0);
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int lane = wid; lane < L; lane += nw) {
    const int q = q0 + lane;
    const int zr = q / G.My, m = q - zr * G.My;
    const bool ok = lane < nq, hh = 2 * m + 1 < G.Ny;
    const float* row = G.Sslab + ((long long)zr * G.Ny + 2 * m) * N;
    for (int x = ln; x < N; x += 32)
      b1[x * LP + lane] = ok ? make_float2(row[x], hh ? row[N + x] : 0.f) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const float2* r = gen_fft<false>(G.px, L, lsh, b1, b0, b1);
  g3_x_store(G, L, G.ZSslab, q0, nq, r);
}

// Source scatter into the dense slab (same arithmetic as k_source_scatter of v1).
__global__ void __launch_bounds__(128) g3_source_scatter(StepParams P, GParams G, SourceParams S) {
  const int t = *P.step;
  const long long slab0 = (long long)G.z0s * P.Ny * P.Nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.n_src;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = S.row_ptr[i]; j < S.row_ptr[i + 1]; ++j) {
      int e = S.col[j];
      int tt = t - S.delay[e];
      if (tt >= 0 && tt < S.n_base) acc = fmaf(S.w[j] * S.gain[e], S.base[tt], acc);
    }
    G.Sslab[S.lin_exp[i] - slab0] = acc * S.scale[i];
  }
}

__global__ void g3_pm_crop(StepParams P, const float2* __restrict__ pm) {
  const long long n = (long long)P.nx * P.ny * P.nz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % P.nx);
    const long long r = i / P.nx;
    const int y = (int)(r % P.ny), z = (int)(r / P.ny);
    const float2 v = pm[((long long)(z + P.pz) * P.Ny + (y + P.py)) * P.Nx + (x + P.px)];
    P.pmax[i] = v.x;
    P.pmin[i] = v.y;
  }
}

}  // namespace lifu
