// step_kernels.cuh -- hand-written element-wise kernels of one k-space time step (v1 pipeline:
// cuFFT 3-D R2C/C2R transforms with these kernels fused between them).
//
// What each kernel replaces inside kspaceFirstOrder3D (reached from
// /root/reference/src/openlifu/sim/kwave_if.py:124-129), in k-Wave step order (ledger A8):
//   k_grad_spectral   p^ -> { i k_xi e^{+i k_xi d/2} kappa p^ }                 (3 spectra)
//   k_update_u        u_xi = pml_sg (pml_sg u_xi - dt/rho0_sg d_xi p)
//   k_div_spectral    u^_xi *= i k_xi e^{-i k_xi d/2} kappa
//   k_source_scatter  S[idx] = scale * sum_e W[i,e] gain_e s(t - n_e)
//   k_source_filter   S^ *= cos(c_ref k dt / 2)
//   k_update_rho_p    rho_xi = pml (pml rho_xi - dt rho0 d_xi u_xi) + S ; p = c0^2 sum rho ; p_max/p_min
//   k_absorb_spectral fractional Laplacians k^{y-2}, k^{y-1}
//   k_pressure_absorb p = c0^2 (sum rho + tau L1 - eta L2) ; p_max/p_min
// All real fields are x-fastest; thread index runs along x so every warp touches one or two
// contiguous 128-byte lines per field.
#pragma once
#include "common.cuh"

namespace lifu {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// sinc(sqrt(s)) and cos(sqrt(s)) as Taylor polynomials in s = (c_ref k dt/2)^2, valid (abs. error < 2e-7) for
// s <= 9.8; the host checks the bound (StepParams::poly_ok) and the kernels fall back to sinf/cosf otherwise.
__device__ __forceinline__ float sinc_sqrt_poly(float s) {
  float r = -9.183689864e-29f;
  r = fmaf(r, s, 6.446950284e-26f);
  r = fmaf(r, s, -3.868170171e-23f);
  r = fmaf(r, s, 1.957294106e-20f);
  r = fmaf(r, s, -8.220635247e-18f);
  r = fmaf(r, s, 2.811457254e-15f);
  r = fmaf(r, s, -7.647163732e-13f);
  r = fmaf(r, s, 1.605904384e-10f);
  r = fmaf(r, s, -2.505210839e-08f);
  r = fmaf(r, s, 2.755731922e-06f);
  r = fmaf(r, s, -1.984126984e-04f);
  r = fmaf(r, s, 8.333333333e-03f);
  r = fmaf(r, s, -1.666666667e-01f);
  r = fmaf(r, s, 1.000000000e+00f);
  return r;
}
__device__ __forceinline__ float cos_sqrt_poly(float s) {
  float r = -2.479596263e-27f;
  r = fmaf(r, s, 1.611737571e-24f);
  r = fmaf(r, s, -8.896791392e-22f);
  r = fmaf(r, s, 4.110317623e-19f);
  r = fmaf(r, s, -1.561920697e-16f);
  r = fmaf(r, s, 4.779477332e-14f);
  r = fmaf(r, s, -1.147074560e-11f);
  r = fmaf(r, s, 2.087675699e-09f);
  r = fmaf(r, s, -2.755731922e-07f);
  r = fmaf(r, s, 2.480158730e-05f);
  r = fmaf(r, s, -1.388888889e-03f);
  r = fmaf(r, s, 4.166666667e-02f);
  r = fmaf(r, s, -5.000000000e-01f);
  r = fmaf(r, s, 1.000000000e+00f);
  return r;
}

// Short forms for s <= 2 (the usual CFL <= 0.5 case): 8 terms, truncation error < 2e-11.
__device__ __forceinline__ float sinc_sqrt_poly8(float s) {
  float r = -7.647163732e-13f;
  r = fmaf(r, s, 1.605904384e-10f);
  r = fmaf(r, s, -2.505210839e-08f);
  r = fmaf(r, s, 2.755731922e-06f);
  r = fmaf(r, s, -1.984126984e-04f);
  r = fmaf(r, s, 8.333333333e-03f);
  r = fmaf(r, s, -1.666666667e-01f);
  r = fmaf(r, s, 1.000000000e+00f);
  return r;
}
__device__ __forceinline__ float cos_sqrt_poly8(float s) {
  float r = -1.147074560e-11f;
  r = fmaf(r, s, 2.087675699e-09f);
  r = fmaf(r, s, -2.755731922e-07f);
  r = fmaf(r, s, 2.480158730e-05f);
  r = fmaf(r, s, -1.388888889e-03f);
  r = fmaf(r, s, 4.166666667e-02f);
  r = fmaf(r, s, -5.000000000e-01f);
  r = fmaf(r, s, 1.000000000e+00f);
  return r;
}

__device__ __forceinline__ float kappa_of(float a2) {
  // kappa = sinc(c_ref k dt / 2); a2 = (c_ref k dt / 2)^2 <= (cfl*pi*sqrt(3)/2)^2, no range issues
  float a = sqrtf(a2);
  return a > 0.f ? sinf(a) / a : 1.f;
}

// POLY: 0 exact sinf/cosf, 1 long polynomials (s <= 9.8), 2 short polynomials (s <= 2)
template <int POLY> __device__ __forceinline__ float kappa_sel(float a2) {
  if constexpr (POLY == 2) return sinc_sqrt_poly8(a2);
  else if constexpr (POLY == 1) return sinc_sqrt_poly(a2);
  else return kappa_of(a2);
}
template <int POLY> __device__ __forceinline__ float cosk_sel(float a2) {
  if constexpr (POLY == 2) return cos_sqrt_poly8(a2);
  else if constexpr (POLY == 1) return cos_sqrt_poly(a2);
  else return cosf(sqrtf(a2));
}

// ------------------------------------------------------------------ K1
static __global__ void __launch_bounds__(256) k_grad_spectral(StepParams P) {
  const long long n = P.Vh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nxh);
    long long t = i / P.Nxh;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    float kap = kappa_of(P.ax2[ix] + P.ay2[iy] + P.az2[iz]) * P.invN;
    float2 v = P.c1[i];
    v.x *= kap;
    v.y *= kap;
    P.c3[i] = cmul(P.dpx[ix], v);
    P.c3[P.CS + i] = cmul(P.dpy[iy], v);
    P.c3[2 * P.CS + i] = cmul(P.dpz[iz], v);
  }
}

// ------------------------------------------------------------------ K3
static __global__ void __launch_bounds__(256) k_div_spectral(StepParams P) {
  const long long n = P.Vh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nxh);
    long long t = i / P.Nxh;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    float kap = kappa_of(P.ax2[ix] + P.ay2[iy] + P.az2[iz]) * P.invN;
    float2 a = P.c3[i], b = P.c3[P.CS + i], c = P.c3[2 * P.CS + i];
    float2 mx = P.dnx[ix], my = P.dny[iy], mz = P.dnz[iz];
    mx.x *= kap; mx.y *= kap; my.x *= kap; my.y *= kap; mz.x *= kap; mz.y *= kap;
    P.c3[i] = cmul(mx, a);
    P.c3[P.CS + i] = cmul(my, b);
    P.c3[2 * P.CS + i] = cmul(mz, c);
  }
}

// ------------------------------------------------------------------ K2
// VEC = 4 requires Nx % 4 == 0 (rows then start 16-byte aligned); VEC = 1 is the general path.
template <int VEC, bool HOMOG>
__global__ void __launch_bounds__(256) k_update_u(StepParams P) {
  const int nxv = P.Nx / VEC;
  const long long n = (long long)nxv * P.Ny * P.Nz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ixv = (int)(i % nxv);
    long long t = i / nxv;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    const float sy = P.sgy[iy], sz = P.sgz[iz];
    long long o = i * VEC;
    if constexpr (VEC == 4) {
      float4 sx = *reinterpret_cast<const float4*>(P.sgx + ixv * 4);
      float4 gx = *reinterpret_cast<const float4*>(P.r3 + o);
      float4 gy = *reinterpret_cast<const float4*>(P.r3 + P.RS + o);
      float4 gz = *reinterpret_cast<const float4*>(P.r3 + 2 * P.RS + o);
      float4 ux = *reinterpret_cast<const float4*>(P.u + o);
      float4 uy = *reinterpret_cast<const float4*>(P.u + P.RS + o);
      float4 uz = *reinterpret_cast<const float4*>(P.u + 2 * P.RS + o);
      float4 dx, dy, dz;
      if constexpr (HOMOG) {
        float s = P.dt_rho0_sg_s;
        dx = make_float4(s, s, s, s); dy = dx; dz = dx;
      } else {
        dx = *reinterpret_cast<const float4*>(P.dt_rho0_sg + o);
        dy = *reinterpret_cast<const float4*>(P.dt_rho0_sg + P.RS + o);
        dz = *reinterpret_cast<const float4*>(P.dt_rho0_sg + 2 * P.RS + o);
      }
      ux.x = sx.x * (sx.x * ux.x - dx.x * gx.x); ux.y = sx.y * (sx.y * ux.y - dx.y * gx.y);
      ux.z = sx.z * (sx.z * ux.z - dx.z * gx.z); ux.w = sx.w * (sx.w * ux.w - dx.w * gx.w);
      uy.x = sy * (sy * uy.x - dy.x * gy.x); uy.y = sy * (sy * uy.y - dy.y * gy.y);
      uy.z = sy * (sy * uy.z - dy.z * gy.z); uy.w = sy * (sy * uy.w - dy.w * gy.w);
      uz.x = sz * (sz * uz.x - dz.x * gz.x); uz.y = sz * (sz * uz.y - dz.y * gz.y);
      uz.z = sz * (sz * uz.z - dz.z * gz.z); uz.w = sz * (sz * uz.w - dz.w * gz.w);
      *reinterpret_cast<float4*>(P.u + o) = ux;
      *reinterpret_cast<float4*>(P.u + P.RS + o) = uy;
      *reinterpret_cast<float4*>(P.u + 2 * P.RS + o) = uz;
    } else {
      float sx = P.sgx[ixv];
      float dx, dy, dz;
      if constexpr (HOMOG) {
        dx = dy = dz = P.dt_rho0_sg_s;
      } else {
        dx = P.dt_rho0_sg[o]; dy = P.dt_rho0_sg[P.RS + o]; dz = P.dt_rho0_sg[2 * P.RS + o];
      }
      P.u[o] = sx * (sx * P.u[o] - dx * P.r3[o]);
      P.u[P.RS + o] = sy * (sy * P.u[P.RS + o] - dy * P.r3[P.RS + o]);
      P.u[2 * P.RS + o] = sz * (sz * P.u[2 * P.RS + o] - dz * P.r3[2 * P.RS + o]);
    }
  }
}

// ------------------------------------------------------------------ source
static __global__ void __launch_bounds__(128) k_source_scatter(StepParams P, SourceParams S) {
  const int t = *P.step;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.n_src;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = S.row_ptr[i]; j < S.row_ptr[i + 1]; ++j) {
      int e = S.col[j];
      int tt = t - S.delay[e];
      if (tt >= 0 && tt < S.n_base) acc = fmaf(S.w[j] * S.gain[e], S.base[tt], acc);
    }
    P.S[S.lin_exp[i]] = acc * S.scale[i];
  }
}

static __global__ void __launch_bounds__(256) k_source_filter(StepParams P) {
  const long long n = P.Vh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nxh);
    long long t = i / P.Nxh;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    float c = cosf(sqrtf(P.ax2[ix] + P.ay2[iy] + P.az2[iz])) * P.invN;
    float2 v = P.c1[i];
    v.x *= c; v.y *= c;
    P.c1[i] = v;
  }
}

// ------------------------------------------------------------------ K4
// SRC: 0 none, 1 dense filtered field Sf, 2 dense unfiltered field S.
// ABSORB: write the two operands of the absorption operators instead of p.
template <bool HOMOG, int SRC, bool ABSORB>
__global__ void __launch_bounds__(256) k_update_rho_p(StepParams P) {
  const long long n = P.V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nx);
    long long t = i / P.Nx;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    const float ax = P.pmlx[ix], ay = P.pmly[iy], az = P.pmlz[iz];
    float dr, c2;
    if constexpr (HOMOG) { dr = P.dt_rho0_s; c2 = P.c2_s; } else { dr = P.dt_rho0[i]; c2 = P.c2[i]; }
    const float dux = P.r3[i], duy = P.r3[P.RS + i], duz = P.r3[2 * P.RS + i];
    float rx = ax * (ax * P.rho[i] - dr * dux);
    float ry = ay * (ay * P.rho[P.RS + i] - dr * duy);
    float rz = az * (az * P.rho[2 * P.RS + i] - dr * duz);
    if constexpr (SRC != 0) {
      float s = (SRC == 1) ? P.Sf[i] : P.S[i];
      rx += s; ry += s; rz += s;
    }
    P.rho[i] = rx; P.rho[P.RS + i] = ry; P.rho[2 * P.RS + i] = rz;
    float sum = (rx + ry) + rz;
    if constexpr (ABSORB) {
      float r0;
      if constexpr (HOMOG) r0 = P.rho0_s; else r0 = P.rho0[i];
      P.r3[i] = r0 * ((dux + duy) + duz);   // operand of the tau term
      P.r3[P.RS + i] = sum;                 // operand of the eta term
      P.r1[i] = sum;                        // kept for the equation of state
    } else {
      float pv = c2 * sum;
      P.p[i] = pv;
      int jx = ix - P.px, jy = iy - P.py, jz = iz + P.z0 - P.pz;
      if ((unsigned)jx < (unsigned)P.nx && (unsigned)jy < (unsigned)P.ny && (unsigned)jz < (unsigned)P.nz) {
        long long j = ((long long)(jz - P.jz0) * P.ny + jy) * P.nx + jx;
        P.pmax[j] = fmaxf(P.pmax[j], pv);
        P.pmin[j] = fminf(P.pmin[j], pv);
      }
    }
    if (i == 0) *P.step = *P.step + 1;
  }
}

// ------------------------------------------------------------------ K5
static __global__ void __launch_bounds__(256) k_absorb_spectral(StepParams P) {
  const long long n = P.Vh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nxh);
    long long t = i / P.Nxh;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    float k2 = P.kx2[ix] + P.ky2[iy] + P.kz2[iz];
    float n1 = 0.f, n2 = 0.f;
    if (k2 > 0.f) {
      n1 = powf(k2, P.y_minus2_half) * P.invN;   // k^(y-2)
      n2 = powf(k2, P.y_minus1_half) * P.invN;   // k^(y-1)
    }
    float2 a = P.c3[i], b = P.c3[P.CS + i];
    a.x *= n1; a.y *= n1; b.x *= n2; b.y *= n2;
    P.c3[i] = a;
    P.c3[P.CS + i] = b;
  }
}

// ------------------------------------------------------------------ K6
template <bool HOMOG>
__global__ void __launch_bounds__(256) k_pressure_absorb(StepParams P, int use_tau, int use_eta) {
  const long long n = P.V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nx);
    long long t = i / P.Nx;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    float c2, tau, eta;
    if constexpr (HOMOG) { c2 = P.c2_s; tau = P.tau_s; eta = P.eta_s; }
    else { c2 = P.c2[i]; tau = P.tau[i]; eta = P.eta[i]; }
    float acc = P.r1[i];
    if (use_tau) acc = acc + tau * P.r3[i];
    if (use_eta) acc = acc - eta * P.r3[P.RS + i];
    float pv = c2 * acc;
    P.p[i] = pv;
    int jx = ix - P.px, jy = iy - P.py, jz = iz + P.z0 - P.pz;
    if ((unsigned)jx < (unsigned)P.nx && (unsigned)jy < (unsigned)P.ny && (unsigned)jz < (unsigned)P.nz) {
      long long j = ((long long)(jz - P.jz0) * P.ny + jy) * P.nx + jx;
      P.pmax[j] = fmaxf(P.pmax[j], pv);
      P.pmin[j] = fminf(P.pmin[j], pv);
    }
  }
}

// ------------------------------------------------------------------ setup kernels
static __global__ void k_fill(float* a, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    a[i] = v;
}

// Expand an inner-grid map to the PML-padded grid by edge replication (ledger A11).
// `in` holds the inner planes [plane0, ...); `out` receives n_planes expanded planes starting at the
// global plane P.z0 (a slab rank asks for one halo plane more than it owns: staggered density along z).
// T = float: x-fastest planes (sx, sy, sz) = (1, nx, nx*ny).  T = double: the caller's float64 array with its own
// element strides (e.g. C order: z fastest), rounded to float32 here exactly as numpy's astype(float32) rounds.
template <typename T>
__global__ void k_expand_edge(const T* __restrict__ in, float* __restrict__ out, StepParams P, int plane0,
                              int n_planes, long long sx, long long sy, long long sz) {
  const long long n = (long long)P.Nx * P.Ny * n_planes;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nx);
    long long t = i / P.Nx;
    int iy = (int)(t % P.Ny);
    int iz = min((int)(t / P.Ny) + P.z0, P.NzG - 1);
    int jx = min(max(ix - P.px, 0), P.nx - 1);
    int jy = min(max(iy - P.py, 0), P.ny - 1);
    int jz = min(max(iz - P.pz, 0), P.nz - 1) - plane0;
    out[i] = (float)in[jz * sz + jy * sy + jx * sx];
  }
}

// The same expansion from a LABEL volume and per-label tables (SegmentationMethod._map_params, seg_method.py:84-97, done
// on the device: one byte per voxel crosses PCIe instead of three float64 maps).  Labels outside the table give 0,
// the reference's initial value.
struct MediumLut { int n; float c0[32], rho0[32], alpha[32]; };
static __global__ void k_expand_edge_lut(const unsigned char* __restrict__ lab, float* __restrict__ c0e, float* __restrict__ rho0e,
                                  float* __restrict__ alphae, StepParams P, int n_planes, long long sx, long long sy,
                                  long long sz, MediumLut lut) {
  const long long n = (long long)P.Nx * P.Ny * n_planes;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nx);
    long long t = i / P.Nx;
    int iy = (int)(t % P.Ny);
    int iz = min((int)(t / P.Ny) + P.z0, P.NzG - 1);
    int jx = min(max(ix - P.px, 0), P.nx - 1);
    int jy = min(max(iy - P.py, 0), P.ny - 1);
    int jz = min(max(iz - P.pz, 0), P.nz - 1);
    const int l = lab[jz * sz + jy * sy + jx * sx];
    const bool ok = l < lut.n;
    c0e[i] = ok ? lut.c0[l] : 0.f;
    rho0e[i] = ok ? lut.rho0[l] : 0.f;
    alphae[i] = ok ? lut.alpha[l] : 0.f;
  }
}

// Packaging of kwave_if.py:136-141 on the device: p_min -> -p_min (float32) and
// intensity = 1e-4 * p_min^2 / (2 Z) with the float32 square and scale and the float64 divide of the numpy expression.
static __global__ void k_package(const float* __restrict__ pmin, const double* __restrict__ two_z, double two_z_s,
                          float* __restrict__ pnp, double* __restrict__ inten, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = pmin[i];
    pnp[i] = -p;
    const float sq = __fmul_rn(__fmul_rn(p, p), 1e-4f);
    inten[i] = __ddiv_rn((double)sq, two_z ? two_z[i] : two_z_s);
  }
}

// Derived medium maps on the expanded grid.  c0e/rho0e/alphae are the expanded maps;
// alpha_np_coef = 100*(1e-6/2pi)^y/(20 log10 e) converts dB/(MHz^y cm) to Np/((rad/s)^y m).
static __global__ void k_derive_medium(const float* __restrict__ c0e, const float* __restrict__ rho0e,
                                const float* __restrict__ alphae, StepParams P, float dt, float y,
                                double alpha_np_coef, double tan_term, float* dt_rho0_sg,
                                float* dt_rho0, float* c2, float* tau, float* eta) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P.V;
       i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i % P.Nx);
    long long t = i / P.Nx;
    int iy = (int)(t % P.Ny);
    int iz = (int)(t / P.Ny);
    double r = rho0e[i];
    double c = c0e[i];
    // staggered density: linear interpolation at +d/2, last plane keeps rho0 (ledger A9)
    double rx = ix + 1 < P.Nx ? 0.5 * (r + (double)rho0e[i + 1]) : r;
    double ry = iy + 1 < P.Ny ? 0.5 * (r + (double)rho0e[i + P.Nx]) : r;
    double rz = iz + P.z0 + 1 < P.NzG ? 0.5 * (r + (double)rho0e[i + (long long)P.Nx * P.Ny]) : r;
    dt_rho0_sg[i] = (float)((double)dt / rx);
    dt_rho0_sg[P.RS + i] = (float)((double)dt / ry);
    dt_rho0_sg[2 * P.RS + i] = (float)((double)dt / rz);
    dt_rho0[i] = (float)((double)dt * r);
    c2[i] = (float)(c * c);
    if (tau != nullptr) {
      double a_np = alpha_np_coef * (double)alphae[i];
      tau[i] = (float)(-2.0 * a_np * pow(c, (double)y - 1.0));
      eta[i] = (float)(2.0 * a_np * pow(c, (double)y) * tan_term);
    }
  }
}

static __global__ void k_source_points(const long long* __restrict__ idx_inner, long long n_src, StepParams P,
                                const float* __restrict__ c0e, float c0_s, double dt, double dx,
                                long long* lin_exp, float* scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_src;
       i += (long long)gridDim.x * blockDim.x) {
    long long l = idx_inner[i];
    int jx = (int)(l % P.nx);
    long long t = l / P.nx;
    int jy = (int)(t % P.ny);
    int jz = (int)(t / P.ny);
    long long e = ((long long)(jz + P.pz - P.z0) * P.Ny + (jy + P.py)) * P.Nx + (jx + P.px);   // local plane index
    lin_exp[i] = e;
    double c = c0e != nullptr ? (double)c0e[e] : (double)c0_s;
    scale[i] = (float)(2.0 * dt / (3.0 * c * dx));   // additive source scaling (ledger A6)
  }
}

}  // namespace lifu
