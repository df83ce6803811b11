// common.cuh -- shared declarations of liblifusim (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler injects itself

#include "lifusim.h"

namespace lifu {

// NVTX range for the lifetime of a scope (tracing row): nsys / ncu --nvtx show the phases of the C ABI
// calls (create, medium, source geometry, set-up / graph capture / time loop / read-back of lifu_run, plan stack, analysis).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

void set_error(const char* fmt, ...);

#define LIFU_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      lifu::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,              \
                      cudaGetErrorString(e__));                                          \
      return LIFU_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

#define LIFU_CUFFT(call)                                                                 \
  do {                                                                                   \
    cufftResult r__ = (call);                                                            \
    if (r__ != CUFFT_SUCCESS) {                                                          \
      lifu::set_error("%s:%d cuFFT error %d in %s", __FILE__, __LINE__, (int)r__, #call); \
      return LIFU_ERR_CUFFT;                                                             \
    }                                                                                    \
  } while (0)

#define LIFU_CHECK(call)          \
  do {                            \
    int s__ = (call);             \
    if (s__ != LIFU_OK) return s__; \
  } while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// Device-side view of everything the step kernels need.  Passed by value.
struct StepParams {
  int Nx, Ny, Nz, Nxh;        // expanded grid, half-spectrum length along x
  int nx, ny, nz;             // inner grid
  int px, py, pz;             // PML thickness
  int z0, NzG;                // slab decomposition: first expanded plane held here and the global Nz (0, Nz otherwise)
  int jz0;                    // first inner plane of the local sensor buffers (0 without a slab decomposition)
  long long V;                // Nx*Ny*Nz
  long long Vh;               // Nxh*Ny*Nz
  long long RS, CS;           // strides between batched real / complex fields
  float invN;                 // 1/(Nx*Ny*Nz): cuFFT transforms are unnormalised
  float inv_dt;               // 1/dt
  int poly_ok;                // (c_ref k dt/2)^2 <= 9.8 everywhere: polynomial sinc/cos are valid
  // 1-D tables
  const float2 *dpx, *dpy, *dpz;   // i k exp(+i k d/2)
  const float2 *dnx, *dny, *dnz;   // i k exp(-i k d/2)
  const float *ax2, *ay2, *az2;    // (c_ref dt k/2)^2 per axis  -> kappa = sinc(sqrt(sum))
  const float *kx2, *ky2, *kz2;    // k^2 per axis (absorption operators)
  const float *pmlx, *pmly, *pmlz, *sgx, *sgy, *sgz;
  // medium
  int homogeneous;
  float dt_rho0_sg_s, dt_rho0_s, c2_s, rho0_s, tau_s, eta_s;   // scalars (homogeneous)
  const float* dt_rho0_sg;    // [3][RS]  dt / rho0 staggered
  const float* dt_rho0;       // [RS]     dt * rho0
  const float* rho0;          // [RS]
  const float* c2;            // [RS]
  const float* tau;           // [RS]
  const float* eta;           // [RS]
  float y_minus2_half, y_minus1_half;  // (y-2)/2 and (y-1)/2 for the fractional Laplacians
  // state
  float* p;                   // [RS]
  float* u;                   // [3][RS]
  float* rho;                 // [3][RS]
  float* r3;                  // [3][RS] scratch: pressure gradients, then velocity gradients
  float* r1;                  // [RS]    scratch
  float* S;                   // [RS]    sparse source field (zeros off the mask)
  float* Sf;                  // [RS]    k-space filtered source field
  float2* c1;                 // [CS]    spectrum scratch
  float2* c3;                 // [3][CS] spectrum scratch
  float* pmax;                // inner grid, x fastest
  float* pmin;
  int* step;                  // device-side time-step counter
};

struct SourceParams {
  long long n_src;
  const long long* lin_exp;   // expanded-grid linear index per source point
  const int* row_ptr;
  const int* col;
  const float* w;
  const float* scale;         // 2 dt / (3 c0 dx) per point
  const float* base;          // base drive signal
  int n_base;
  const int* delay;           // per element
  const float* gain;          // per element
};

// Per-pipeline parameters (device pointers owned by the handle).
struct V2Params {
  int Nx, Ny, Nz, Nxh, PH;          // PH = row pitch of the H layout (multiple of 16)
  int nxt;                          // regular 16-lane kx tiles of the strided passes (Nx/32); tile nxt = Nyquist column
  int Ry;                           // radix of the y axis: rows y and y+Ry form a packed row pair
  int ry_sh, hy_sh;                 // log2(Ry), log2(Ny/2)
  long long HS;                     // stride between batched H fields  (Nz*zsH)
  long long zsH;                    // plane stride of the H layout: Ny*PH (v2) or Ny*PH + pad (wide: keeps the z stride off 2^15)
  long long ZS;                     // stride between batched Z fields  (Nz*(Ny/2)*Nx)
  const float4 *tw4x, *tw4y, *tw4z; // (w, i w), w = exp(-2 pi i m / N), m = 0..N (entry N = entry 0), per axis
  const float4 *dpy4, *dny4, *dpz4, *dnz4;   // derivative multipliers i k e^{+-i k d/2} as (m, i m)
  float2* ZP;                       // x-spectrum of the pressure (row pairs)
  float2* Z4;                       // [4][ZS] packed spectra (gradients, then velocity, then divergence + source)
  float2* H4;                       // [4][HS] half spectra
  float2* pm;                       // [Nz][Ny][Nx] running (p_max, p_min) on the expanded grid
  float norm;                       // 1 / (2 * Nx*Ny*Nz): FFT normalisation and the row-pair split factor
  // source slab (planes z0s .. z0s+nzs-1 of the expanded grid)
  int z0s, nzs;
  float* Sslab;                     // [nzs][Ny][Nx] dense source field on the slab
  float2* ZSslab;                   // [nzs][Ny/2][Nx]
  float2* HSslab;                   // [nzs][Ny][PH]
  int store_p;                      // write the real-space pressure (last step / debugging)
  int bx0;                          // added to blockIdx.x by the strided passes (nxt: a 1-wide grid does the Nyquist column only)
  int xrev;                         // persistent x kernels walk their row pairs from the last one down (L2 reuse across passes)
  int zmajor;                       // batched y passes: grid (tiles, ncomp, Nz) instead of (tiles, Nz, ncomp)
  int pm_always;                    // write the sensor pair back even when it did not change (measurement switch)
  // steady-state source (see v2_build_steady in lifusim.cu): while every element is driven, the delayed drive signals
  // span a space of rank <= 2 in time, S_t = q_1(t) F_1 + q_2(t) F_2, and the k-space source filter is applied ONCE to
  // the two spatial fields instead of to S_t on every step
  const float* FK;                  // [2][RS] filtered basis fields (real, expanded grid)
  const float* qsrc;                // [2][nws] time coefficients q_k(t0s + i)
  float* qcur;                      // [2] coefficients of the current step (written by the first kernel of the step)
  int t0s, nws;                     // first step and length of the steady window (nws = 0: off)
  int comp0;                        // component offset of k2_y_inv, first component of k2_z_div
  // z-slab decomposition of pipeline wide (wide.cu): Nz above is the LOCAL plane count, H4 the local half spectra
  // [4][Nzl][Ny][PH]; T4 holds this rank's ky rows of ALL planes, [4][NzG][Nyl][PH].  The y-forward kernels store straight
  // into the owners' T4 and the z kernels straight into the owners' H4 through `peer` (every rank's exchange buffer,
  // mapped with CUDA IPC; H4 at +0, T4 at +peerT).  G = 0: one GPU, no routing.
  int G, Nzl, NzG, Nyl, z0g, ky0;
  float2* T4;
  float2* const* peer;
  long long peerT;
  int gz0s, gnzs;                   // planes of the source mask on the whole grid (z0s / nzs above: this rank's part)
};

// Pipeline v3 (fft_gen.cuh): generic-radix fused passes.
struct GenPlan {
  int N, ns;
  int radix[8];
  const float2* tw;          // exp(-2 pi i m / N), m < N
};

struct GParams {
  int Nx, Ny, Nz, Nxh, PH, My;
  long long HS, ZS;           // strides between batched H / Z fields
  GenPlan px, py, pz;
  float2* ZP;                 // x-spectrum of the pressure (row pairs)
  float2* Z4;                 // [4][ZS]
  float2* H4;                 // [4][HS]
  float2* pm;                 // [Nz][Ny][Nx] running (p_max, p_min) on the expanded grid
  float norm;                 // 1 / (2 Nx Ny Nz)
  int z0s, nzs;               // source slab planes
  float* Sslab;               // [nzs][Ny][Nx]
  float2* ZSslab;             // [nzs][My][Nx]
  float2* HSslab;             // [nzs][Ny][PH]
  int store_p;
  int Ls, lsh_s;              // lanes per tile of the strided passes (power of two) and its log2
  int Lx, lsh_x;              // row pairs per batch of the x passes
  int pm_always;
};


}  // namespace lifu
