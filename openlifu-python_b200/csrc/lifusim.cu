// lifusim.cu -- C ABI (include/lifusim.h) and time-loop orchestration of the B200 k-space solver.
//
// Replaces, for the path openlifu.sim.run_simulation drives, what k-wave-python's
// kspaceFirstOrder3D + the kspaceFirstOrder-OMP/-CUDA binary do
// (/root/reference/src/openlifu/sim/kwave_if.py:117-129): PML sizing and grid expansion,
// k-space operators, the time loop and the p_max/p_min sensor reduction.
#include <cstdarg>
#include <functional>

#include "sim.cuh"
#include "step_kernels.cuh"
#include "fft_v2.cuh"
#include "fft_z_tma.cuh"
#include "fft_gen.cuh"
#include "slab.cuh"

namespace lifu {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int dev_alloc(lifu_sim* s, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? LIFU_ERR_NOMEM : LIFU_ERR_CUDA;
  }
  s->allocs.push_back(*p);
  return LIFU_OK;
}

// ------------------------------------------------------------------------------------------
// host-side grid logic
static int largest_prime_factor(int n) {
  int best = 1;
  for (int p = 2; (long long)p * p <= n; ++p)
    while (n % p == 0) { best = p; n /= p; }
  if (n > 1) best = n;
  return best;
}

static void pml_auto(const int32_t n[3], int32_t out[3]) {
  for (int a = 0; a < 3; ++a) {
    int best_p = 10, best_f = INT32_MAX;
    for (int p = 10; p <= 40; ++p) {
      int f = largest_prime_factor(n[a] + 2 * p);
      if (f < best_f) { best_f = f; best_p = p; }   // first minimum wins
    }
    out[a] = best_p;
  }
}

// kWaveGrid wavenumber of FFT bin i (ifftshifted; Nyquist negative for even N)
static double k_fft(int N, double d, int i) {
  int j = i <= (N - 1) / 2 ? i : i - N;   // even N: bin N/2 -> -N/2
  return (2.0 * M_PI / d) * ((double)j / (double)N);
}

static void pml_profile(int N, double d, double dt, double c, int size, double alpha, bool sg,
                        std::vector<float>& out) {
  out.assign(N, 1.0f);
  for (int i = 1; i <= size; ++i) {
    double x = (double)i;
    double l, r;
    if (sg) {
      l = alpha * (c / d) * std::pow(((x + 0.5) - size - 1.0) / (0.0 - size), 4.0);
      r = alpha * (c / d) * std::pow((x + 0.5) / size, 4.0);
    } else {
      l = alpha * (c / d) * std::pow((x - size - 1.0) / (0.0 - size), 4.0);
      r = alpha * (c / d) * std::pow(x / size, 4.0);
    }
    out[i - 1] = (float)std::exp(-l * dt / 2.0);
    out[N - size + i - 1] = (float)std::exp(-r * dt / 2.0);
  }
}

// Build every 1-D table that depends on c_ref; called once c_ref is known.
static int build_tables(lifu_sim* s) {
  const int Nx = s->N[0], Ny = s->N[1], Nz = s->N[2], Nxh = s->Nxh;
  const double* d = s->grid.d;
  const double dt = s->grid.dt, c = s->c_ref;
  const double alpha = s->grid.pml_alpha > 0 ? s->grid.pml_alpha : 2.0;
  std::vector<float> host;
  auto push = [&](const std::vector<float>& v) {
    size_t off = host.size();
    host.insert(host.end(), v.begin(), v.end());
    while (host.size() % 4) host.push_back(0.f);   // keep every table 16-byte aligned
    return off;
  };
  size_t off_dp[3], off_dn[3], off_a2[3], off_k2[3], off_pml[3], off_sg[3];
  const int Ns[3] = {Nx, Ny, Nz};
  for (int a = 0; a < 3; ++a) {
    int len = a == 0 ? Nxh : Ns[a];
    std::vector<float> dp(2 * len), dn(2 * len), a2(len), k2(len);
    for (int i = 0; i < len; ++i) {
      double k = k_fft(Ns[a], d[a], i);
      // i k exp(+-i k d/2) = k * (-+sin(k d/2) + i cos(k d/2))
      double sn = std::sin(k * d[a] / 2.0), cs = std::cos(k * d[a] / 2.0);
      dp[2 * i] = (float)(-k * sn); dp[2 * i + 1] = (float)(k * cs);
      dn[2 * i] = (float)(k * sn);  dn[2 * i + 1] = (float)(k * cs);
      double arg = c * k * dt / 2.0;
      a2[i] = (float)(arg * arg);
      k2[i] = (float)(k * k);
    }
    off_dp[a] = push(dp); off_dn[a] = push(dn); off_a2[a] = push(a2); off_k2[a] = push(k2);
    std::vector<float> pm, sg;
    pml_profile(Ns[a], d[a], dt, c, s->pml[a], alpha, false, pm);
    pml_profile(Ns[a], d[a], dt, c, s->pml[a], alpha, true, sg);
    off_pml[a] = push(pm); off_sg[a] = push(sg);
  }
  if (!s->d_tables) LIFU_CHECK(dev_alloc(s, (void**)&s->d_tables, host.size() * sizeof(float)));
  LIFU_CUDA(cudaMemcpyAsync(s->d_tables, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
  LIFU_CUDA(cudaStreamSynchronize(s->stream));
  StepParams& P = s->P;
  const float* T = s->d_tables;
  P.dpx = (const float2*)(T + off_dp[0]); P.dpy = (const float2*)(T + off_dp[1]); P.dpz = (const float2*)(T + off_dp[2]);
  P.dnx = (const float2*)(T + off_dn[0]); P.dny = (const float2*)(T + off_dn[1]); P.dnz = (const float2*)(T + off_dn[2]);
  P.ax2 = T + off_a2[0]; P.ay2 = T + off_a2[1]; P.az2 = T + off_a2[2];
  P.kx2 = T + off_k2[0]; P.ky2 = T + off_k2[1]; P.kz2 = T + off_k2[2];
  P.pmlx = T + off_pml[0]; P.pmly = T + off_pml[1]; P.pmlz = T + off_pml[2];
  P.sgx = T + off_sg[0]; P.sgy = T + off_sg[1]; P.sgz = T + off_sg[2];
  if (s->sl.on) { P.pmlz += s->sl.z0; P.sgz += s->sl.z0; }   // real-space kernels index z by the local plane
  {
    double smax = 0;
    for (int a = 0; a < 3; ++a) { double arg = c * (M_PI / d[a]) * dt / 2.0; smax += arg * arg; }
    P.poly_ok = smax <= 2.0 ? 2 : (smax <= 9.8 ? 1 : 0);
  }
  s->tables_ready = true;
  return LIFU_OK;
}

static int make_plan(lifu_sim* s, cufftHandle* h, cufftType type, int batch, size_t* ws) {
  int n[3] = {s->N[2], s->N[1], s->N[0]};
  int rembed[3] = {s->N[2], s->N[1], s->N[0]};
  int cembed[3] = {s->N[2], s->N[1], s->Nxh};
  LIFU_CUFFT(cufftCreate(h));
  LIFU_CUFFT(cufftSetAutoAllocation(*h, 0));
  size_t w = 0;
  if (type == CUFFT_R2C) {
    LIFU_CUFFT(cufftMakePlanMany(*h, 3, n, rembed, 1, (int)s->RS, cembed, 1, (int)s->CS, CUFFT_R2C, batch, &w));
  } else {
    LIFU_CUFFT(cufftMakePlanMany(*h, 3, n, cembed, 1, (int)s->CS, rembed, 1, (int)s->RS, CUFFT_C2R, batch, &w));
  }
  LIFU_CUFFT(cufftSetStream(*h, s->stream));
  if (w > *ws) *ws = w;
  return LIFU_OK;
}

static int build_plans(lifu_sim* s) {
  if (s->plans_ready) return LIFU_OK;
  if (s->RS >= (1LL << 31)) { set_error("grid too large for a single-GPU cuFFT plan"); return LIFU_ERR_INVALID; }
  size_t ws = 0;
  LIFU_CHECK(make_plan(s, &s->r2c1, CUFFT_R2C, 1, &ws));
  LIFU_CHECK(make_plan(s, &s->r2c3, CUFFT_R2C, 3, &ws));
  LIFU_CHECK(make_plan(s, &s->c2r1, CUFFT_C2R, 1, &ws));
  LIFU_CHECK(make_plan(s, &s->c2r3, CUFFT_C2R, 3, &ws));
  LIFU_CHECK(make_plan(s, &s->r2c2, CUFFT_R2C, 2, &ws));
  LIFU_CHECK(make_plan(s, &s->c2r2, CUFFT_C2R, 2, &ws));
  s->work_bytes = ws;
  LIFU_CHECK(dev_alloc(s, &s->d_work, ws));
  cufftHandle hs[6] = {s->r2c1, s->r2c3, s->c2r1, s->c2r3, s->r2c2, s->c2r2};
  for (cufftHandle h : hs) LIFU_CUFFT(cufftSetWorkArea(h, s->d_work));
  s->plans_ready = true;
  return LIFU_OK;
}

int upload_source_points(lifu_sim* s) {
  // (re)compute expanded-grid indices and the additive-source scale once both the geometry and
  // the medium are known
  if (!s->geometry_set || !s->medium_set) return LIFU_OK;
  long long i0 = 0, i1 = s->n_src;
  if (s->sl.on && s->n_src > 0) {
    // the points are sorted x fastest / z slowest: this rank's planes are one contiguous range
    std::vector<long long> h(s->n_src);
    LIFU_CUDA(cudaMemcpyAsync(h.data(), s->d_idx, sizeof(long long) * s->n_src, cudaMemcpyDeviceToHost, s->stream));
    LIFU_CUDA(cudaStreamSynchronize(s->stream));
    const long long plane = (long long)s->n[0] * s->n[1];
    i0 = std::lower_bound(h.begin(), h.end(), (long long)s->sl.jz_lo * plane) - h.begin();
    i1 = std::lower_bound(h.begin(), h.end(), (long long)(s->sl.jz_lo + s->sl.jz_n) * plane) - h.begin();
  }
  s->sl.src_i0 = i0; s->sl.src_i1 = i1;
  const long long cnt = i1 - i0;
  // the buffers are kept across calls (this runs on every medium change: cudaFree / cudaMalloc synchronise the device
  // and were seen to stall for hundreds of milliseconds now and then)
  if (cnt > s->src_pts_cap || !s->d_lin_exp) {
    if (s->d_lin_exp) { cudaFree(s->d_lin_exp); s->d_lin_exp = nullptr; }
    if (s->d_scale) { cudaFree(s->d_scale); s->d_scale = nullptr; }
    s->src_pts_cap = 0;
    LIFU_CUDA(cudaMalloc(&s->d_lin_exp, sizeof(long long) * std::max<long long>(cnt, 1)));
    LIFU_CUDA(cudaMalloc(&s->d_scale, sizeof(float) * std::max<long long>(cnt, 1)));
    s->src_pts_cap = std::max<long long>(cnt, 1);
  }
  if (cnt > 0) {
    k_source_points<<<grid_blocks(s, cnt, 128), 128, 0, s->stream>>>(
        s->d_idx + i0, cnt, s->P, s->homogeneous ? nullptr : s->d_c0e, s->c0_s, s->grid.dt, s->grid.d[0],
        s->d_lin_exp, s->d_scale);
    LIFU_CUDA(cudaGetLastError());
  }
  return LIFU_OK;
}

}  // namespace lifu

// ------------------------------------------------------------------------------------------
// small device reduction used by lifu_set_medium (min and max of a positive field)
namespace lifu {
__global__ void k_minmax(const float* __restrict__ a, long long n, float* out /*[2]*/) {
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = a[i];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[32], smx[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    mn = l < nw ? smn[l] : INFINITY;
    mx = l < nw ? smx[l] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (l == 0) {
      // values of interest are finite; order-preserving int trick is valid for non-negative floats,
      // negative minima are handled through the sign-flipped encoding below
      auto enc = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : (int)(0x80000000u - (unsigned)i); };
      atomicMin((int*)out, enc(mn));
      atomicMax((int*)out + 1, enc(mx));
    }
  }
}
}  // namespace lifu

namespace lifu {
int reduce_minmax(lifu_sim* s, const float* a, long long n, float* mn, float* mx) {
  int* red = s->P.step + 1;
  int init[2] = {INT32_MAX, INT32_MIN};
  LIFU_CUDA(cudaMemcpyAsync(red, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
  lifu::k_minmax<<<lifu::grid_blocks(s, n, 256, 4), 256, 0, s->stream>>>(a, n, (float*)red);
  LIFU_CUDA(cudaGetLastError());
  int h[2];
  LIFU_CUDA(cudaMemcpyAsync(h, red, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  LIFU_CUDA(cudaStreamSynchronize(s->stream));
  auto dec = [](int i) { int j = i >= 0 ? i : (int)(0x80000000u - (unsigned)i); float f; memcpy(&f, &j, 4); return f; };
  *mn = dec(h[0]);
  *mx = dec(h[1]);
  return LIFU_OK;
}
}  // namespace lifu

using namespace lifu;

// =========================================================================================
extern "C" {

int lifu_abi_version(void) { return LIFUSIM_ABI_VERSION; }

const char* lifu_last_error(void) { return g_err.c_str(); }

int lifu_make_time(const int32_t n[3], const double d[3], double c_ref, double cfl, int32_t* nt, double* dt) {
  if (!n || !d || !nt || !dt || c_ref <= 0 || cfl <= 0) { set_error("lifu_make_time: bad argument"); return LIFU_ERR_INVALID; }
  double s2 = 0, dmin = d[0];
  for (int a = 0; a < 3; ++a) { double L = (double)n[a] * d[a]; s2 += L * L; dmin = std::min(dmin, d[a]); }
  double t_end = std::sqrt(s2) / c_ref;
  double dt_ = cfl * dmin / c_ref;
  double q = t_end / dt_;
  int Nt = (int)std::floor(q) + 1;
  if (std::floor(q) != std::ceil(q) && std::fmod(t_end, dt_) == 0.0) Nt += 1;
  *nt = Nt;
  *dt = dt_;
  return LIFU_OK;
}

int lifu_device_count(int32_t* count) {
  if (!count) { set_error("lifu_device_count: null argument"); return LIFU_ERR_INVALID; }
  *count = 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return LIFU_OK; }
  for (int d = 0; d < ndev; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++*count;
  }
  return LIFU_OK;
}

int lifu_pml_auto(const int32_t n[3], int32_t pml_out[3]) {
  if (!n || !pml_out) { set_error("lifu_pml_auto: null argument"); return LIFU_ERR_INVALID; }
  for (int a = 0; a < 3; ++a)
    if (n[a] <= 0) { set_error("lifu_pml_auto: grid size must be positive"); return LIFU_ERR_INVALID; }
  pml_auto(n, pml_out);
  return LIFU_OK;
}

static int create_impl(const lifu_grid* g, int device, void* cuda_stream, const lifu_slab_desc* slab, lifu_sim** out) {
  NvtxRange nvtx_r("lifu_create");
  if (!g || !out) { set_error("lifu_create: null argument"); return LIFU_ERR_INVALID; }
  *out = nullptr;
  for (int a = 0; a < 3; ++a) {
    if (g->n[a] <= 0 || !(g->d[a] > 0)) { set_error("lifu_create: grid size and spacing must be positive"); return LIFU_ERR_INVALID; }
  }
  if (!(g->dt > 0) || g->nt <= 0) { set_error("lifu_create: dt and nt must be positive"); return LIFU_ERR_INVALID; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    set_error("lifu_create: no CUDA device available (%s); this library has no CPU fallback",
              ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    return LIFU_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("lifu_create: device %d out of range [0,%d)", device, ndev); return LIFU_ERR_INVALID; }
  LIFU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LIFU_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("lifu_create: device %d is sm_%d%d; liblifusim is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return LIFU_ERR_CUDA;
  }
  lifu_sim* s = new lifu_sim();
  s->grid = *g;
  s->device = device;
  s->stream = (cudaStream_t)cuda_stream;
  s->n_sm = prop.multiProcessorCount;
  int32_t pa[3];
  pml_auto(g->n, pa);
  for (int a = 0; a < 3; ++a) {
    s->n[a] = g->n[a];
    s->pml[a] = g->pml[a] >= 0 ? g->pml[a] : pa[a];
    s->N[a] = s->n[a] + 2 * s->pml[a];
    s->grid.pml[a] = s->pml[a];
  }
  const char* ng = getenv("LIFU_NO_GRAPH");
  s->use_graph = !(ng && ng[0] == '1');
  const char* pl = getenv("LIFU_PIPELINE");
  s->pipeline = (pl && !strcmp(pl, "v1")) ? 1 : ((pl && !strcmp(pl, "v2")) ? 2 : ((pl && !strcmp(pl, "v3")) ? 3 : 0));
  s->Nxh = s->N[0] / 2 + 1;
  s->V = (long long)s->N[0] * s->N[1] * s->N[2];
  s->Vin = (long long)s->n[0] * s->n[1] * s->n[2];
  int nz_loc = s->N[2];
  if (slab) {
    if (slab->nranks < 1 || slab->rank < 0 || slab->rank >= slab->nranks || slab->exchange < 0 || slab->exchange > 2) {
      set_error("lifu_create_slab: bad rank %d / nranks %d / exchange %d", slab->rank, slab->nranks, slab->exchange);
      delete s; return LIFU_ERR_INVALID;
    }
    if (s->N[2] % slab->nranks || s->N[1] % slab->nranks) {
      set_error("lifu_create_slab: expanded grid %dx%dx%d: Ny and Nz must be multiples of the %d ranks",
                s->N[0], s->N[1], s->N[2], slab->nranks);
      delete s; return LIFU_ERR_INVALID;
    }
    if (!s->stream) {   // collectives and kernels share one real stream for the handle's lifetime
      if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete s; return LIFU_ERR_CUDA; }
      s->stream = s->own_stream;
    }
    int rcs = slab_init(s, slab);
    if (rcs != LIFU_OK) { lifu_destroy(s); return rcs; }
    nz_loc = s->sl.Nzl;
    s->pipeline = 1;
    s->use_graph = false;
  }
  s->Vloc = (long long)s->N[0] * s->N[1] * nz_loc;
  s->Vsens = slab ? (long long)s->n[0] * s->n[1] * s->sl.jz_n : s->Vin;
  s->Vh = (long long)s->Nxh * s->N[1] * nz_loc;
  s->RS = round_up(s->Vloc, 128);
  s->CS = round_up(s->Vh, 64);

  StepParams& P = s->P;
  P.Nx = s->N[0]; P.Ny = s->N[1]; P.Nz = nz_loc; P.Nxh = s->Nxh;
  P.nx = s->n[0]; P.ny = s->n[1]; P.nz = s->n[2];
  P.px = s->pml[0]; P.py = s->pml[1]; P.pz = s->pml[2];
  P.z0 = slab ? s->sl.z0 : 0; P.NzG = s->N[2]; P.jz0 = slab ? s->sl.jz_lo : 0;
  P.V = s->Vloc; P.Vh = s->Vh; P.RS = s->RS; P.CS = s->CS;
  P.invN = (float)(1.0 / (double)s->V);
  P.inv_dt = (float)(1.0 / g->dt);

  int rc = LIFU_OK;
  auto A = [&](void** p, size_t bytes) { if (rc == LIFU_OK) rc = dev_alloc(s, p, bytes); };
  const size_t R = sizeof(float) * s->RS, C = sizeof(float2) * s->CS;
  A((void**)&P.p, R); A((void**)&P.u, 3 * R); A((void**)&P.rho, 3 * R); A((void**)&P.r3, 3 * R);
  A((void**)&P.r1, R); A((void**)&P.S, R); A((void**)&P.Sf, R);
  if (!slab) { A((void**)&P.c1, C); A((void**)&P.c3, 3 * C); }
  A((void**)&P.pmax, sizeof(float) * s->Vsens); A((void**)&P.pmin, sizeof(float) * s->Vsens);
  A((void**)&P.step, sizeof(int) * 4);
  if (rc == LIFU_OK) for (int i = 0; i < 3 && rc == LIFU_OK; ++i)
    if (cudaEventCreate(&s->ev[i]) != cudaSuccess) { set_error("cudaEventCreate failed"); rc = LIFU_ERR_CUDA; }
  if (rc != LIFU_OK) { lifu_destroy(s); return rc; }
  if (g->c_ref > 0) {
    s->c_ref = g->c_ref;
    rc = build_tables(s);
    if (rc != LIFU_OK) { lifu_destroy(s); return rc; }
  }
  if (slab) s->sl.ready = true;
  *out = s;
  return LIFU_OK;
}

int lifu_create(const lifu_grid* g, int device, void* cuda_stream, lifu_sim** out) {
  return create_impl(g, device, cuda_stream, nullptr, out);
}

int lifu_slab_unique_id(unsigned char id[LIFU_NCCL_ID_BYTES]) {
  if (!id) { set_error("lifu_slab_unique_id: null argument"); return LIFU_ERR_INVALID; }
  NcclApi* N = nccl_api();
  if (!N) return LIFU_ERR_STATE;
  ncclUniqueId u;
  LIFU_NCCL(N->GetUniqueId(&u));
  memcpy(id, u.internal, LIFU_NCCL_ID_BYTES);
  return LIFU_OK;
}

int lifu_create_slab(const lifu_grid* g, int device, void* cuda_stream, const lifu_slab_desc* slab, lifu_sim** out) {
  if (!slab) { set_error("lifu_create_slab: null slab descriptor"); return LIFU_ERR_INVALID; }
  return create_impl(g, device, cuda_stream, slab, out);
}

int lifu_slab_layout_of(lifu_sim* s, lifu_slab_layout* o) {
  if (!s || !o) { set_error("lifu_slab_layout_of: null argument"); return LIFU_ERR_INVALID; }
  if (!s->sl.on) { set_error("lifu_slab_layout_of: not a slab handle"); return LIFU_ERR_STATE; }
  const SlabCtx& L = s->sl;
  o->rank = L.rank; o->nranks = L.G; o->exchange = L.exchange;
  o->z0 = L.z0; o->nz = L.Nzl;
  o->sensor_z0 = L.jz_lo; o->sensor_nz = L.jz_n;
  o->medium_z0 = L.med_lo; o->medium_nz = L.med_n;
  return LIFU_OK;
}

int lifu_destroy(lifu_sim* s) {
  if (!s) return LIFU_OK;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  slab_destroy(s);
  for (int i = 0; i < 2; ++i) if (s->graph[i]) cudaGraphExecDestroy(s->graph[i]);
  cufftHandle hs[6] = {s->r2c1, s->r2c3, s->c2r1, s->c2r3, s->r2c2, s->c2r2};
  if (s->plans_ready || s->r2c1) for (cufftHandle h : hs) if (h) cufftDestroy(h);
  for (void* p : s->allocs) cudaFree(p);
  cudaFree(s->d_bli_ws);
  cudaFree(s->d_idx); cudaFree(s->d_lin_exp); cudaFree(s->d_row_ptr); cudaFree(s->d_col);
  cudaFree(s->d_w); cudaFree(s->d_scale); cudaFree(s->d_base); cudaFree(s->d_delay); cudaFree(s->d_gain);
  for (int i = 0; i < 3; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  delete s;
  return LIFU_OK;
}

// stride64 != NULL: the maps are float64 arrays of the whole inner grid with those element strides (x, y, z).
static int set_medium_impl(lifu_sim* s, const float* c0, const float* rho0, const float* alpha_db,
                           float alpha_power, int alpha_mode, int homogeneous, int plane0, int n_planes,
                           const int64_t* stride64 = nullptr, const unsigned char* labels = nullptr,
                           const MediumLut* lut = nullptr) {
  if (!s || (!labels && (!c0 || !rho0))) { set_error("lifu_set_medium: null argument"); return LIFU_ERR_INVALID; }
  if (alpha_mode < 0 || alpha_mode > 2) { set_error("lifu_set_medium: alpha_mode %d unknown", alpha_mode); return LIFU_ERR_INVALID; }
  NvtxRange nvtx_r("lifu_set_medium");
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = s->stream;
  StepParams& P = s->P;
  s->alpha_power = alpha_power;
  s->alpha_mode = alpha_mode;
  s->homogeneous = homogeneous != 0;
  for (int i = 0; i < 2; ++i) if (s->graph[i]) { cudaGraphExecDestroy(s->graph[i]); s->graph[i] = nullptr; }
  const double y = alpha_power;
  const double np_coef = 100.0 * std::pow(1e-6 / (2.0 * M_PI), y) / (20.0 * std::log10(M_E));
  const double tan_term = std::tan(M_PI * y / 2.0);
  P.y_minus2_half = (float)((y - 2.0) / 2.0);
  P.y_minus1_half = (float)((y - 1.0) / 2.0);
  double c_max;
  if (s->homogeneous) {
    float hc = 0, hr = 0, ha = 0;
    LIFU_CUDA(cudaMemcpyAsync(&hc, c0, sizeof(float), cudaMemcpyDefault, st));
    LIFU_CUDA(cudaMemcpyAsync(&hr, rho0, sizeof(float), cudaMemcpyDefault, st));
    if (alpha_db) LIFU_CUDA(cudaMemcpyAsync(&ha, alpha_db, sizeof(float), cudaMemcpyDefault, st));
    LIFU_CUDA(cudaStreamSynchronize(st));
    if (!(hc > 0) || !(hr > 0) || ha < 0) { set_error("lifu_set_medium: need c0 > 0, rho0 > 0, alpha >= 0"); return LIFU_ERR_INVALID; }
    s->c0_s = hc; s->rho0_s = hr; s->alpha_s = ha;
    s->absorbing = ha != 0.f;
    const double dt = s->grid.dt;
    P.homogeneous = 1;
    P.dt_rho0_sg_s = (float)(dt / (double)hr);
    P.dt_rho0_s = (float)(dt * (double)hr);
    P.c2_s = (float)((double)hc * (double)hc);
    P.rho0_s = hr;
    double a_np = np_coef * (double)ha;
    P.tau_s = (float)(-2.0 * a_np * std::pow((double)hc, y - 1.0));
    P.eta_s = (float)(2.0 * a_np * std::pow((double)hc, y) * tan_term);
    c_max = hc;
  } else {
    // planes of the inner grid this handle reads: all of them, or a slab's own range plus the halo
    const int need_lo = s->sl.on ? s->sl.med_lo : 0, need_n = s->sl.on ? s->sl.med_n : s->n[2];
    if (plane0 > need_lo || plane0 + n_planes < need_lo + need_n) {
      set_error("lifu_set_medium: maps cover inner planes [%d,%d) but [%d,%d) are needed", plane0, plane0 + n_planes,
                need_lo, need_lo + need_n);
      return LIFU_ERR_INVALID;
    }
    P.homogeneous = 0;
    const size_t R = sizeof(float) * s->RS;
    const int n_exp = s->sl.on ? s->sl.Nzl + 1 : s->N[2];            // a slab keeps one halo plane (staggered density)
    const long long Ve = (long long)s->N[0] * s->N[1] * n_exp;
    if (!s->d_c0e) {
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_c0e, sizeof(float) * Ve));
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_rho0e, sizeof(float) * Ve));
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_alphae, sizeof(float) * Ve));
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_med, 7 * R));
    }
    // stage the needed inner planes through the (currently idle) scratch field r3
    float* stage = P.r3;
    const long long plane = (long long)s->n[0] * s->n[1];
    const size_t inb = sizeof(float) * plane * need_n;
    const int gbe = grid_blocks(s, Ve, 256);
    const int gb = grid_blocks(s, s->Vloc, 256);
    const float* src[3] = {c0, rho0, alpha_db};
    float* dst[3] = {s->d_c0e, s->d_rho0e, s->d_alphae};
    if (labels) {
      // label volume (one byte per voxel, the caller's layout) + per-label tables: expanded maps in one pass
      LIFU_CUDA(cudaMemcpyAsync(stage, labels, (size_t)s->Vin, cudaMemcpyDefault, st));
      k_expand_edge_lut<<<gbe, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(stage), dst[0], dst[1], dst[2], P, n_exp,
                                             stride64[0], stride64[1], stride64[2], *lut);
    }
    for (int m = 0; m < 3 && !labels; ++m) {
      if (src[m] && stride64) {
        // float64 maps in the caller's layout: one block copy, conversion + re-layout on the device
        LIFU_CUDA(cudaMemcpyAsync(stage, src[m], sizeof(double) * (size_t)s->Vin, cudaMemcpyDefault, st));
        k_expand_edge<double><<<gbe, 256, 0, st>>>(reinterpret_cast<const double*>(stage), dst[m], P, 0, n_exp,
                                                   stride64[0], stride64[1], stride64[2]);
      } else if (src[m]) {
        LIFU_CUDA(cudaMemcpyAsync(stage, src[m] + (long long)(need_lo - plane0) * plane, inb, cudaMemcpyDefault, st));
        k_expand_edge<float><<<gbe, 256, 0, st>>>(stage, dst[m], P, need_lo, n_exp, 1, s->n[0], plane);
      } else {
        k_fill<<<gbe, 256, 0, st>>>(dst[m], Ve, 0.f);
      }
    }
    LIFU_CUDA(cudaGetLastError());
    float* d_dt_rho0_sg = s->d_med;
    float* d_dt_rho0 = s->d_med + 3 * s->RS;
    float* d_c2 = s->d_med + 4 * s->RS;
    float* d_tau = s->d_med + 5 * s->RS;
    float* d_eta = s->d_med + 6 * s->RS;
    k_derive_medium<<<gb, 256, 0, st>>>(s->d_c0e, s->d_rho0e, s->d_alphae, P, (float)s->grid.dt, alpha_power,
                                        np_coef, tan_term, d_dt_rho0_sg, d_dt_rho0, d_c2, d_tau, d_eta);
    LIFU_CUDA(cudaGetLastError());
    P.dt_rho0_sg = d_dt_rho0_sg; P.dt_rho0 = d_dt_rho0; P.c2 = d_c2; P.tau = d_tau; P.eta = d_eta;
    P.rho0 = s->d_rho0e;
    // c_ref = max(c0) and absorbing = any(alpha != 0) come from device reductions
    float cmax = 0, amax = 0, cmin = 0, amin = 0;
    LIFU_CHECK(reduce_minmax(s, s->d_c0e, s->Vloc, &cmin, &cmax));
    LIFU_CHECK(reduce_minmax(s, s->d_alphae, s->Vloc, &amin, &amax));
    if (s->sl.on) {
      float v[4] = {cmax, -cmin, amax, -amin};
      LIFU_CHECK(slab_allreduce_max(s, v, 4));
      cmax = v[0]; cmin = -v[1]; amax = v[2]; amin = -v[3];
    }
    if (!(cmin > 0)) { set_error("lifu_set_medium: sound speed must be positive everywhere"); return LIFU_ERR_INVALID; }
    if (amin < 0) { set_error("lifu_set_medium: attenuation must be non-negative"); return LIFU_ERR_INVALID; }
    s->absorbing = amax != 0.f;
    c_max = cmax;
    s->c0_s = cmax;
  }
  if (s->grid.c_ref <= 0) {
    if (!s->tables_ready || s->c_ref != c_max) { s->c_ref = c_max; LIFU_CHECK(build_tables(s)); }
  }
  s->medium_set = true;
  return upload_source_points(s);
}

int lifu_set_medium(lifu_sim* s, const float* c0, const float* rho0, const float* alpha_db,
                    float alpha_power, int alpha_mode, int homogeneous) {
  return set_medium_impl(s, c0, rho0, alpha_db, alpha_power, alpha_mode, homogeneous, 0, s ? s->n[2] : 0);
}

int lifu_set_medium_f64(lifu_sim* s, const double* c0, const double* rho0, const double* alpha_db,
                        const int64_t stride[3], float alpha_power, int alpha_mode) {
  if (!s || !c0 || !rho0 || !stride) { set_error("lifu_set_medium_f64: null argument"); return LIFU_ERR_INVALID; }
  if (s->sl.on) { set_error("lifu_set_medium_f64: not available on a slab handle (use lifu_set_medium_planes)"); return LIFU_ERR_STATE; }
  // the strides must be a dense permutation of (Nx, Ny, Nz): the array is copied as one block
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return stride[a] < stride[b] || (stride[a] == stride[b] && s->n[a] < s->n[b]); });
  int64_t expect = 1;
  for (int k = 0; k < 3; ++k) {
    if (s->n[order[k]] > 1 && stride[order[k]] != expect) {
      set_error("lifu_set_medium_f64: strides (%lld, %lld, %lld) do not describe a dense (%d, %d, %d) array",
                (long long)stride[0], (long long)stride[1], (long long)stride[2], s->n[0], s->n[1], s->n[2]);
      return LIFU_ERR_INVALID;
    }
    expect *= s->n[order[k]];
  }
  if (sizeof(double) * (size_t)s->Vin > sizeof(float) * 3 * (size_t)s->RS) { set_error("lifu_set_medium_f64: staging area too small"); return LIFU_ERR_NOMEM; }
  return set_medium_impl(s, reinterpret_cast<const float*>(c0), reinterpret_cast<const float*>(rho0),
                         reinterpret_cast<const float*>(alpha_db), alpha_power, alpha_mode, 0, 0, s->n[2], stride);
}

int lifu_set_medium_labels(lifu_sim* s, const uint8_t* labels, const int64_t stride[3], int32_t n_labels,
                           const double* c0, const double* rho0, const double* alpha_db, float alpha_power, int alpha_mode) {
  if (!s || !labels || !stride || !c0 || !rho0 || n_labels <= 0 || n_labels > 32) {
    set_error("lifu_set_medium_labels: bad argument (1 to 32 labels)");
    return LIFU_ERR_INVALID;
  }
  if (s->sl.on) { set_error("lifu_set_medium_labels: not available on a slab handle"); return LIFU_ERR_STATE; }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return stride[a] < stride[b] || (stride[a] == stride[b] && s->n[a] < s->n[b]); });
  int64_t expect = 1;
  for (int k = 0; k < 3; ++k) {
    if (s->n[order[k]] > 1 && stride[order[k]] != expect) {
      set_error("lifu_set_medium_labels: strides (%lld, %lld, %lld) do not describe a dense (%d, %d, %d) array",
                (long long)stride[0], (long long)stride[1], (long long)stride[2], s->n[0], s->n[1], s->n[2]);
      return LIFU_ERR_INVALID;
    }
    expect *= s->n[order[k]];
  }
  MediumLut lut;
  lut.n = n_labels;
  for (int i = 0; i < 32; ++i) {
    lut.c0[i] = i < n_labels ? (float)c0[i] : 0.f;               // data_cast='single'
    lut.rho0[i] = i < n_labels ? (float)rho0[i] : 0.f;
    lut.alpha[i] = (i < n_labels && alpha_db) ? (float)alpha_db[i] : 0.f;
  }
  return set_medium_impl(s, nullptr, nullptr, nullptr, alpha_power, alpha_mode, 0, 0, s->n[2], stride, labels, &lut);
}

int lifu_set_medium_planes(lifu_sim* s, const float* c0, const float* rho0, const float* alpha_db,
                           float alpha_power, int alpha_mode, int homogeneous, int32_t plane0, int32_t n_planes) {
  if (s && !homogeneous && (plane0 < 0 || n_planes <= 0 || plane0 + n_planes > s->n[2])) {
    set_error("lifu_set_medium_planes: planes [%d,%d) outside the inner grid (nz = %d)", plane0, plane0 + n_planes, s->n[2]);
    return LIFU_ERR_INVALID;
  }
  return set_medium_impl(s, c0, rho0, alpha_db, alpha_power, alpha_mode, homogeneous, plane0, n_planes);
}

int lifu_set_elements(lifu_sim* s, int32_t n_el, const double* pos_m, const double* size_m,
                      const double* angle_deg, double bli_tolerance, int32_t upsampling_rate, int64_t* n_src) {
  if (!s || !pos_m || !size_m || !angle_deg) { set_error("lifu_set_elements: null argument"); return LIFU_ERR_INVALID; }
  LIFU_CUDA(cudaSetDevice(s->device));
  NvtxRange nvtx_r("lifu_set_elements (BLI source geometry)");
  LIFU_CHECK(bli_build(s, n_el, pos_m, size_m, angle_deg, bli_tolerance, upsampling_rate));
  if (n_src) *n_src = s->n_src;
  return LIFU_OK;
}

int lifu_set_source_geometry(lifu_sim* s, const int64_t* idx, const int32_t* row_ptr, const int32_t* col_elem,
                             const float* w, int64_t n_src, int64_t nnz, int32_t n_el) {
  if (!s || n_src < 0 || nnz < 0 || n_el <= 0 || (n_src > 0 && (!idx || !row_ptr)) || (nnz > 0 && (!col_elem || !w))) {
    set_error("lifu_set_source_geometry: bad argument");
    return LIFU_ERR_INVALID;
  }
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = s->stream;
  s->geometry_set = false;
  cudaFree(s->d_idx); cudaFree(s->d_row_ptr); cudaFree(s->d_col); cudaFree(s->d_w);
  s->d_idx = nullptr; s->d_row_ptr = nullptr; s->d_col = nullptr; s->d_w = nullptr;
  s->idx_cap = 0; s->nnz_cap = 0;
  LIFU_CUDA(cudaMalloc(&s->d_idx, sizeof(long long) * std::max<int64_t>(n_src, 1)));
  LIFU_CUDA(cudaMalloc(&s->d_row_ptr, sizeof(int) * (n_src + 1)));
  LIFU_CUDA(cudaMalloc(&s->d_col, sizeof(int) * std::max<int64_t>(nnz, 1)));
  LIFU_CUDA(cudaMalloc(&s->d_w, sizeof(float) * std::max<int64_t>(nnz, 1)));
  s->idx_cap = std::max<int64_t>(n_src, 1); s->nnz_cap = std::max<int64_t>(nnz, 1);
  if (n_src > 0) {
    LIFU_CUDA(cudaMemcpyAsync(s->d_idx, idx, sizeof(long long) * n_src, cudaMemcpyDefault, st));
    LIFU_CUDA(cudaMemcpyAsync(s->d_row_ptr, row_ptr, sizeof(int) * (n_src + 1), cudaMemcpyDefault, st));
  } else {
    LIFU_CUDA(cudaMemsetAsync(s->d_row_ptr, 0, sizeof(int), st));
  }
  if (nnz > 0) {
    LIFU_CUDA(cudaMemcpyAsync(s->d_col, col_elem, sizeof(int) * nnz, cudaMemcpyDefault, st));
    LIFU_CUDA(cudaMemcpyAsync(s->d_w, w, sizeof(float) * nnz, cudaMemcpyDefault, st));
  }
  LIFU_CUDA(cudaStreamSynchronize(st));
  s->n_src = n_src; s->nnz = nnz; s->n_el = n_el;
  s->geometry_set = true;
  return upload_source_points(s);
}

int lifu_get_source_sizes(lifu_sim* s, int64_t* n_src, int64_t* nnz, int32_t* n_el) {
  if (!s || !s->geometry_set) { set_error("lifu_get_source_sizes: no source geometry set"); return LIFU_ERR_STATE; }
  if (n_src) *n_src = s->n_src;
  if (nnz) *nnz = s->nnz;
  if (n_el) *n_el = s->n_el;
  return LIFU_OK;
}

int lifu_get_source_geometry(lifu_sim* s, int64_t* idx, int32_t* row_ptr, int32_t* col_elem, float* w) {
  if (!s || !s->geometry_set) { set_error("lifu_get_source_geometry: no source geometry set"); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = s->stream;
  if (idx && s->n_src) LIFU_CUDA(cudaMemcpyAsync(idx, s->d_idx, sizeof(long long) * s->n_src, cudaMemcpyDefault, st));
  if (row_ptr) LIFU_CUDA(cudaMemcpyAsync(row_ptr, s->d_row_ptr, sizeof(int) * (s->n_src + 1), cudaMemcpyDefault, st));
  if (col_elem && s->nnz) LIFU_CUDA(cudaMemcpyAsync(col_elem, s->d_col, sizeof(int) * s->nnz, cudaMemcpyDefault, st));
  if (w && s->nnz) LIFU_CUDA(cudaMemcpyAsync(w, s->d_w, sizeof(float) * s->nnz, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaStreamSynchronize(st));
  return LIFU_OK;
}

int lifu_set_drive(lifu_sim* s, const float* base_signal, int32_t n_base, const int32_t* delay_samples,
                   const float* gains, int32_t n_el, int source_mode) {
  if (!s || !base_signal || !delay_samples || !gains || n_base <= 0 || n_el <= 0) {
    set_error("lifu_set_drive: bad argument");
    return LIFU_ERR_INVALID;
  }
  if (source_mode != LIFU_SOURCE_ADDITIVE && source_mode != LIFU_SOURCE_ADDITIVE_NO_CORRECTION) {
    set_error("lifu_set_drive: source_mode %d unknown", source_mode);
    return LIFU_ERR_INVALID;
  }
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = s->stream;
  if (n_base != s->n_base || !s->d_base) {
    cudaFree(s->d_base); s->d_base = nullptr;
    LIFU_CUDA(cudaMalloc(&s->d_base, sizeof(float) * n_base));
  }
  if (n_el != s->drive_n_el || !s->d_delay) {
    cudaFree(s->d_delay); cudaFree(s->d_gain); s->d_delay = nullptr; s->d_gain = nullptr;
    LIFU_CUDA(cudaMalloc(&s->d_delay, sizeof(int) * n_el));
    LIFU_CUDA(cudaMalloc(&s->d_gain, sizeof(float) * n_el));
  }
  std::vector<int> hd(n_el);
  LIFU_CUDA(cudaMemcpyAsync(s->d_base, base_signal, sizeof(float) * n_base, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaMemcpyAsync(s->d_delay, delay_samples, sizeof(int) * n_el, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaMemcpyAsync(s->d_gain, gains, sizeof(float) * n_el, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaMemcpyAsync(hd.data(), s->d_delay, sizeof(int) * n_el, cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaStreamSynchronize(st));
  int md = 0;
  for (int v : hd) {
    if (v < 0) { set_error("lifu_set_drive: negative delay sample count"); return LIFU_ERR_INVALID; }
    md = std::max(md, v);
  }
  if (source_mode != s->source_mode)
    for (int i = 0; i < 2; ++i) if (s->graph[i]) { cudaGraphExecDestroy(s->graph[i]); s->graph[i] = nullptr; }
  s->n_base = n_base; s->drive_n_el = n_el; s->max_delay = md; s->source_mode = source_mode;
  s->drive_set = true;
  return LIFU_OK;
}

}  // extern "C"


// =========================================================================================
// time loop
namespace lifu {

template <int VEC>
static void launch_update_u(lifu_sim* s, int gb) {
  if (s->homogeneous) k_update_u<VEC, true><<<gb, 256, 0, s->stream>>>(s->P);
  else k_update_u<VEC, false><<<gb, 256, 0, s->stream>>>(s->P);
}

template <bool HOMOG, int SRC>
static void launch_rho_p(lifu_sim* s, int gb) {
  if (s->absorbing) k_update_rho_p<HOMOG, SRC, true><<<gb, 256, 0, s->stream>>>(s->P);
  else k_update_rho_p<HOMOG, SRC, false><<<gb, 256, 0, s->stream>>>(s->P);
}


// ------------------------------------------------------------------------------------------
// pipeline v2: fused hand-written FFT passes (fft_v2.cuh)
static int radix_of(int n) { return n == 64 ? 8 : (n == 256 ? 16 : 0); }

static bool v2_square(const lifu_sim* s) {
  for (int a = 0; a < 3; ++a) if (radix_of(s->N[a]) == 0) return false;
  return true;
}
// pipeline "wide" (fft_wide.cuh, wide.cu): the same fused passes for 64 / 128 / 256 / 512 / 768 / 1024-point axes.  Taken for
// non-square factorisations when asked for (LIFU_PIPELINE=v2) or when LIFU_WIDE_AUTO is not "0".
static bool wide_ok(const lifu_sim* s) {
  for (int a = 0; a < 3; ++a) if (!wide_ab(s->N[a], nullptr, nullptr)) return false;
  if (s->drive_set && s->source_mode != LIFU_SOURCE_ADDITIVE) return false;
  return true;
}
static bool wide_auto() {
  const char* e = getenv("LIFU_WIDE_AUTO");
  return !(e && e[0] == '0');
}
// Slab decomposition on the fused passes (wide.cu): peer-store exchange only; the block arithmetic of the routed stores
// needs G | B of the y factorisation (Nyl = A B / G rows per rank) and G | A of the z factorisation (Nzl planes per rank).
static bool slab_wide_ok(const lifu_sim* s) {
  if (!s->sl.on || s->sl.exchange != 2 || !wide_ok(s)) return false;
  const char* e = getenv("LIFU_SLAB_WIDE");
  if (e && e[0] == '0') return false;
  int A = 0, B = 0;
  wide_ab(s->N[1], &A, &B);
  if (B % s->sl.G) return false;
  wide_ab(s->N[2], &A, &B);
  if (A % s->sl.G) return false;
  return true;
}
static bool v2_eligible(const lifu_sim* s) {
  if (s->pipeline == 1 || s->pipeline == 3) return false;
  if (v2_square(s)) return true;
  return wide_ok(s) && (s->pipeline == 2 || wide_auto());
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (liblifusim does not link libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (tmap_encode_fn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// Tile = 16 kx of one ky row over all z planes of a complex (8-byte) field [comp][z][ky][kx(PH)]
static bool encode_z_tile_map(void* out, void* base, int PH, int Ny, int nz, int ncomp) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc || nz > 256) return false;
  cuuint64_t dims[4] = {(cuuint64_t)PH, (cuuint64_t)Ny, (cuuint64_t)nz, (cuuint64_t)ncomp};
  cuuint64_t strides[3] = {(cuuint64_t)PH * 8, (cuuint64_t)PH * 8 * Ny, (cuuint64_t)PH * 8 * Ny * nz};
  cuuint32_t box[4] = {16, 1, (cuuint32_t)nz, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const int rank = ncomp > 0 ? 4 : 3;
  CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int v2_setup(lifu_sim* s) {
  V2Params& Q = s->Q;
  const bool sl = s->sl.on;                  // one slab of a decomposed grid: local planes, exchange buffers of slab_init
  if (!s->v2_ready) {
    Q.Nx = s->N[0]; Q.Ny = s->N[1]; Q.Nz = sl ? s->sl.Nzl : s->N[2]; Q.Nxh = s->Nxh;
    Q.PH = (int)round_up(s->Nxh, 16);
    Q.nxt = s->N[0] / 32;
    {
      // wide: pad every H plane so that the z stride is not a multiple of 32 KB (LIFU_WIDE_HPAD, in complex elements)
      const char* hp = getenv("LIFU_WIDE_HPAD");
      const int pad = (hp && hp[0]) ? atoi(hp) : 0;        // measured: no effect on 256^3 / 512^3 (profiles/r2_wide_summary.md)
      Q.zsH = (long long)Q.Ny * Q.PH + ((!sl && (!v2_square(s) || (getenv("LIFU_WIDE_SQUARE") && getenv("LIFU_WIDE_SQUARE")[0] == '1'))) ? pad : 0);
    }
    Q.HS = sl ? s->sl.Hl : (long long)Q.Nz * Q.zsH;
    if (sl && s->sl.Hl != (long long)Q.Nz * Q.zsH) { set_error("slab exchange buffers do not match the padded half-spectrum layout"); return LIFU_ERR_STATE; }
    Q.ZS = (long long)Q.Nz * (Q.Ny / 2) * Q.Nx;
    Q.norm = (float)(1.0 / (2.0 * (double)s->V));
    const char* wsq = getenv("LIFU_WIDE_SQUARE");           // measurement switch: square grids through the wide kernels too
    s->v2_wide = sl || !v2_square(s) || (wsq && wsq[0] == '1' && wide_ok(s));
    for (int a = 0; a < 3; ++a) {
      s->R[a] = radix_of(s->N[a]);
      if (s->v2_wide) wide_ab(s->N[a], nullptr, &s->R[a]);     // threads per line (B of N = A x B)
      const int n = s->N[a];
      std::vector<float4> tw(n + 1);
      for (int m = 0; m <= n; ++m) {
        double ang = -2.0 * M_PI * (double)(m % n) / (double)n;
        float c = (float)std::cos(ang), sn = (float)std::sin(ang);
        tw[m] = make_float4(c, sn, -sn, c);                          // (w, i w)
      }
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_tw[a], sizeof(float4) * (n + 1)));
      LIFU_CUDA(cudaMemcpyAsync(s->d_tw[a], tw.data(), sizeof(float4) * (n + 1), cudaMemcpyHostToDevice, s->stream));
      LIFU_CUDA(cudaStreamSynchronize(s->stream));
    }
    Q.Ry = s->R[1];
    Q.ry_sh = Q.Ry == 8 ? 3 : (Q.Ry == 16 ? 4 : 5);
    Q.hy_sh = 0;
    while ((1 << Q.hy_sh) < Q.Ny / 2) ++Q.hy_sh;
    Q.tw4x = s->d_tw[0]; Q.tw4y = s->d_tw[1]; Q.tw4z = s->d_tw[2];
    {
      // derivative multipliers i k e^{+-i k d/2} of the y and z axes as (m, i m) pairs
      const int Ny = s->N[1], Nz = s->N[2];
      std::vector<float4> mul(2 * Ny + 2 * Nz);
      auto fill = [&](float4* dst, int n, double d, double sign) {
        for (int i = 0; i < n; ++i) {
          double k = k_fft(n, d, i);
          float re = (float)(-sign * k * std::sin(k * d / 2.0)), im = (float)(k * std::cos(k * d / 2.0));
          dst[i] = make_float4(re, im, -im, re);
        }
      };
      fill(mul.data(), Ny, s->grid.d[1], +1.0);
      fill(mul.data() + Ny, Ny, s->grid.d[1], -1.0);
      fill(mul.data() + 2 * Ny, Nz, s->grid.d[2], +1.0);
      fill(mul.data() + 2 * Ny + Nz, Nz, s->grid.d[2], -1.0);
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_mul4, sizeof(float4) * mul.size()));
      LIFU_CUDA(cudaMemcpyAsync(s->d_mul4, mul.data(), sizeof(float4) * mul.size(), cudaMemcpyHostToDevice, s->stream));
      LIFU_CUDA(cudaStreamSynchronize(s->stream));
      Q.dpy4 = s->d_mul4; Q.dny4 = s->d_mul4 + Ny; Q.dpz4 = s->d_mul4 + 2 * Ny; Q.dnz4 = s->d_mul4 + 2 * Ny + Nz;
    }
    LIFU_CHECK(dev_alloc(s, (void**)&Q.ZP, sizeof(float2) * Q.ZS));
    LIFU_CHECK(dev_alloc(s, (void**)&Q.Z4, sizeof(float2) * 4 * Q.ZS));
    if (sl) {
      Q.H4 = s->sl.xbuf; Q.T4 = s->sl.xbuf + 4 * s->sl.Hl;
      Q.G = s->sl.G; Q.Nzl = s->sl.Nzl; Q.NzG = s->N[2]; Q.Nyl = s->sl.Nyl; Q.z0g = s->sl.z0; Q.ky0 = s->sl.ky0;
      Q.peer = s->sl.d_peer; Q.peerT = 4 * s->sl.Hl;
    } else {
      LIFU_CHECK(dev_alloc(s, (void**)&Q.H4, sizeof(float2) * 4 * Q.HS));
      Q.G = 0; Q.T4 = nullptr; Q.peer = nullptr;
    }
    LIFU_CHECK(dev_alloc(s, (void**)&Q.pm, sizeof(float2) * s->Vloc));
    s->v2_ready = true;
  }
  // source slab: z range of the mask (indices are sorted, x fastest, so first/last give min/max z)
  int z0 = 0, nz = 1;
  const long long n_pts = sl ? s->sl.src_i1 - s->sl.src_i0 : s->n_src;        // a slab handle holds its own planes' points
  if (n_pts > 0) {
    long long first = 0, last = 0;
    LIFU_CUDA(cudaMemcpyAsync(&first, s->d_lin_exp, sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    LIFU_CUDA(cudaMemcpyAsync(&last, s->d_lin_exp + (n_pts - 1), sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    LIFU_CUDA(cudaStreamSynchronize(s->stream));
    const long long plane = (long long)s->N[0] * s->N[1];
    z0 = (int)(first / plane);
    nz = (int)(last / plane) - z0 + 1;
  }
  if (sl) {
    // planes of the mask on the WHOLE grid (every rank filters the same source field), then this rank's share of them:
    // the owners store all their planes of the range, zero where they hold no points
    float v[2] = {n_pts > 0 ? -(float)(s->sl.z0 + z0) : -1e9f, n_pts > 0 ? (float)(s->sl.z0 + z0 + nz - 1) : -1e9f};
    LIFU_CHECK(slab_allreduce_max(s, v, 2));
    if (v[1] < 0.f) { Q.gz0s = 0; Q.gnzs = 0; }
    else { Q.gz0s = (int)(-v[0]); Q.gnzs = (int)v[1] - Q.gz0s + 1; }
    const int lo = std::max(Q.gz0s, s->sl.z0), hi = std::min(Q.gz0s + Q.gnzs, s->sl.z0 + s->sl.Nzl);
    z0 = hi > lo ? lo - s->sl.z0 : 0;
    nz = hi > lo ? hi - lo : 0;
  }
  if (std::max(nz, 1) > s->slab_planes_alloc) {
    const int nzs_keep = nz;
    nz = std::max(nz, 1);
    // (re)allocate the slab buffers; old ones stay in the handle's allocation list until destroy
    LIFU_CHECK(dev_alloc(s, (void**)&Q.Sslab, sizeof(float) * (size_t)nz * s->N[1] * s->N[0]));
    LIFU_CHECK(dev_alloc(s, (void**)&Q.ZSslab, sizeof(float2) * (size_t)nz * (s->N[1] / 2) * s->N[0]));
    LIFU_CHECK(dev_alloc(s, (void**)&Q.HSslab, sizeof(float2) * (size_t)nz * Q.zsH));
    s->slab_planes_alloc = nz;
    nz = nzs_keep;
  }
  Q.z0s = z0; Q.nzs = nz;
  Q.store_p = 0;
  Q.bx0 = 0;
  Q.comp0 = 0; Q.nws = 0; Q.t0s = 0;
  // traversal-order switches (LIFU_V2_ORDER bit 0: x kernels from the last row pair down -- on by default, bit 1:
  // plane-major batched y passes -- off, measured slower; LIFU_PM_ALWAYS=1: unconditional sensor write-back);
  // measurements in profiles/r2_order_experiments.md
  {
    const char* ord = getenv("LIFU_V2_ORDER");
    const int o = ord ? atoi(ord) : 1;
    Q.xrev = o & 1; Q.zmajor = (o >> 1) & 1;
    const char* pa = getenv("LIFU_PM_ALWAYS");
    Q.pm_always = (pa && pa[0] == '1') ? 1 : 0;
    const char* yg = getenv("LIFU_V2_YGRAD");
    s->v2_ygrad_split = yg && !strcmp(yg, "split");
  }
  // TMA descriptors of the z passes.  Opt-in (LIFU_Z_TMA=1): on C2 the persistent TMA-fed kernels measure 3 % slower
  // than the per-thread-load kernels (profiles/r1_z_tma.md) -- the z passes are bound by their two 256-point
  // transforms per element, not by load latency.
  const char* zt = getenv("LIFU_Z_TMA");
  s->z_tma = !s->v2_wide && (zt && zt[0] == '1') && encode_z_tile_map(s->tmH, Q.H4, Q.PH, Q.Ny, Q.Nz, 4) &&
             encode_z_tile_map(s->tmS, Q.HSslab, Q.PH, Q.Ny, Q.nzs, 0);
  return LIFU_OK;
}

// launch helpers: every v2 kernel uses dynamic shared memory above the 48 KB default
template <typename K, typename... Args>
static void v2_launch(K kernel, dim3 grid, int threads, size_t sm, cudaStream_t st, Args... args) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  kernel<<<grid, threads, sm, st>>>(args...);
}
#define V2_R(RVAL, EXPR) do { if ((RVAL) == 8) { constexpr int RR = 8; EXPR; } else { constexpr int RR = 16; EXPR; } } while (0)

template <int R> static void v2_launch_x_u(lifu_sim* s, int nbatch) {
  dim3 grid(std::min(nbatch, s->n_sm * (s->homogeneous ? 3 : 2)));
  if (s->homogeneous) v2_launch(k2_x_u<R, true>, grid, 128, XStageU<R, true>::SMEM, s->stream, s->P, s->Q);
  else v2_launch(k2_x_u<R, false>, grid, 128, XStageU<R, false>::SMEM, s->stream, s->P, s->Q);
}
template <int R, int SRC> static void v2_launch_x_rho_p(lifu_sim* s, int nbatch) {
  dim3 grid(std::min(nbatch, s->n_sm * (s->homogeneous ? 3 : 2)));
  if (s->absorbing) {
    if (s->homogeneous) v2_launch(k2_x_rho_p<R, true, SRC, true>, grid, 128, XStageRho<R, true>::SMEM, s->stream, s->P, s->Q);
    else v2_launch(k2_x_rho_p<R, false, SRC, true>, grid, 128, XStageRho<R, false>::SMEM, s->stream, s->P, s->Q);
  } else {
    if (s->homogeneous) v2_launch(k2_x_rho_p<R, true, SRC>, grid, 128, XStageRho<R, true>::SMEM, s->stream, s->P, s->Q);
    else v2_launch(k2_x_rho_p<R, false, SRC>, grid, 128, XStageRho<R, false>::SMEM, s->stream, s->P, s->Q);
  }
}
template <int R> static void v2_launch_x_p(lifu_sim* s, int nbatch) {
  dim3 grid(std::min(nbatch, s->n_sm * (s->homogeneous ? 3 : 2)));
  const int use_tau = s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION, use_eta = s->alpha_mode != LIFU_ALPHA_NO_DISPERSION;
  if (s->homogeneous) v2_launch(k2_x_p<R, true>, grid, 128, XStageU<R, true>::SMEM, s->stream, s->P, s->Q, use_tau, use_eta);
  else v2_launch(k2_x_p<R, false>, grid, 128, XStageU<R, false>::SMEM, s->stream, s->P, s->Q, use_tau, use_eta);
}

// ------------------------------------------------------------------------------------------
// Steady-state source.  S_t = sum_e W_e gain_e s(t - n_e) (A.7); while every element is inside its burst,
// t in [max n_e, min n_e + n_base), the delayed copies s(t - n) of the base signal span a space of rank <= 2 for the
// sinusoidal bursts OpenLIFU drives (kwave_if.py:101-102) -- checked numerically here, not assumed: pivoted Gram-Schmidt
// over the distinct delays, accepted only when the residual is below 5e-7 of the signal norm.  Then
// S_t = q_1(t) F_1 + q_2(t) F_2 with F_k = sum_e W_e gain_e c_{k,e}, and the k-space source filter cos(c_ref k dt/2),
// being linear and time-invariant, is applied once to F_1 and F_2 (set-up: four small launches each) instead of to S_t on
// every step: on those steps the source costs 8 B per voxel of reads in k2_x_rho_p and nothing else.
static int v2_build_steady(lifu_sim* s, int nt) {
  V2Params& Q = s->Q;
  Q.nws = 0; Q.t0s = 0;
  const char* env = getenv("LIFU_SOURCE_STEADY");
  if ((env && env[0] == '0') || s->source_mode != LIFU_SOURCE_ADDITIVE || s->n_src <= 0) return LIFU_OK;
  const int n_el = s->n_el, nb = s->n_base;
  std::vector<float> base(nb), gain(n_el);
  std::vector<int> delay(n_el);
  cudaStream_t st = s->stream;
  LIFU_CUDA(cudaMemcpyAsync(base.data(), s->d_base, sizeof(float) * nb, cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaMemcpyAsync(gain.data(), s->d_gain, sizeof(float) * n_el, cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaMemcpyAsync(delay.data(), s->d_delay, sizeof(int) * n_el, cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaStreamSynchronize(st));
  int dmin = INT32_MAX, dmax = 0;
  for (int d : delay) { dmin = std::min(dmin, d); dmax = std::max(dmax, d); }
  const int t0 = dmax, t1 = std::min(dmin + nb, nt);
  const int nw = t1 - t0;
  if (nw < 24) return LIFU_OK;
  std::vector<int> dd(delay);
  std::sort(dd.begin(), dd.end());
  dd.erase(std::unique(dd.begin(), dd.end()), dd.end());
  const int m = (int)dd.size();
  // rows A_j[i] = s(t0 + i - dd[j]); pivoted Gram-Schmidt in float64
  auto A = [&](int j, int i) { return (double)base[t0 + i - dd[j]]; };
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) { double r = 0; for (int i = 0; i < nw; ++i) r += a[i] * b[i]; return r; };
  std::vector<std::vector<double>> R(m, std::vector<double>(nw));
  double top = 0; int p1 = 0;
  for (int j = 0; j < m; ++j) {
    for (int i = 0; i < nw; ++i) R[j][i] = A(j, i);
    const double n2 = dot(R[j], R[j]);
    if (n2 > top) { top = n2; p1 = j; }
  }
  if (!(top > 0)) return LIFU_OK;
  std::vector<double> q1(R[p1]), q2(nw, 0.0);
  { const double inv = 1.0 / std::sqrt(top); for (double& v : q1) v *= inv; }
  std::vector<double> c1(m), c2(m, 0.0);
  double top2 = 0; int p2 = -1;
  for (int j = 0; j < m; ++j) {
    c1[j] = dot(R[j], q1);
    for (int i = 0; i < nw; ++i) R[j][i] -= c1[j] * q1[i];
    const double n2 = dot(R[j], R[j]);
    if (n2 > top2) { top2 = n2; p2 = j; }
  }
  const double tol2 = 5e-7 * 5e-7 * top;
  if (p2 >= 0 && top2 > tol2) {
    q2 = R[p2];
    const double inv = 1.0 / std::sqrt(top2);
    for (double& v : q2) v *= inv;
    for (int j = 0; j < m; ++j) {
      c2[j] = dot(R[j], q2);
      for (int i = 0; i < nw; ++i) R[j][i] -= c2[j] * q2[i];
      if (dot(R[j], R[j]) > tol2) return LIFU_OK;          // rank > 2: keep the generic path on every step
    }
  }
  // device side: coefficients per element, time coefficients, the two filtered basis fields
  std::vector<float> coef(2 * (size_t)n_el), qs(2 * (size_t)nw);
  for (int e = 0; e < n_el; ++e) {
    const int j = (int)(std::lower_bound(dd.begin(), dd.end(), delay[e]) - dd.begin());
    coef[e] = (float)((double)gain[e] * c1[j]);
    coef[n_el + e] = (float)((double)gain[e] * c2[j]);
  }
  for (int i = 0; i < nw; ++i) { qs[i] = (float)q1[i]; qs[nw + i] = (float)q2[i]; }
  if (!s->d_fk) LIFU_CHECK(dev_alloc(s, (void**)&s->d_fk, sizeof(float) * 2 * (size_t)s->RS));
  if (!s->d_qcur) LIFU_CHECK(dev_alloc(s, (void**)&s->d_qcur, sizeof(float) * 4));
  if ((size_t)nw > s->qsrc_cap || !s->d_qsrc) {
    LIFU_CHECK(dev_alloc(s, (void**)&s->d_qsrc, sizeof(float) * 2 * (size_t)nw));
    s->qsrc_cap = nw;
  }
  if ((size_t)n_el > s->coef_cap || !s->d_coef) {
    LIFU_CHECK(dev_alloc(s, (void**)&s->d_coef, sizeof(float) * 2 * (size_t)n_el));
    s->coef_cap = n_el;
  }
  LIFU_CUDA(cudaMemcpyAsync(s->d_qsrc, qs.data(), sizeof(float) * qs.size(), cudaMemcpyHostToDevice, st));
  LIFU_CUDA(cudaMemcpyAsync(s->d_coef, coef.data(), sizeof(float) * coef.size(), cudaMemcpyHostToDevice, st));
  LIFU_CUDA(cudaStreamSynchronize(st));                      // the host vectors go out of scope
  const int Rx = s->R[0], Ry = s->R[1], Rz = s->R[2];
  const unsigned tx = Q.nxt + 1;
  V2Params Qs = Q;
  Qs.comp0 = 3; Qs.zmajor = 0;
  const int poly = s->P.poly_ok;
  const int gs = (int)(((long long)Q.nzs * (Q.Ny / 2) + (256 / Rx) - 1) / (256 / Rx));
  const int gall = (int)((long long)Q.Nz * (Q.Ny / 2) / (256 / Rx));
  const size_t smx = (size_t)(256 / Rx) * Rx * Rx * 8 + 16 * (Rx * Rx + 1);
  for (int k = 0; k < 2; ++k) {
    LIFU_CUDA(cudaMemsetAsync(Q.Sslab, 0, sizeof(float) * (size_t)Q.nzs * s->N[1] * s->N[0], st));
    k2_source_basis<<<grid_blocks(s, s->n_src, 128), 128, 0, st>>>(s->P, Qs, s->S, s->d_coef + (size_t)k * n_el);
    V2_R(Rx, (v2_launch(k2_x_src<RR>, dim3(gs), 256, smx, st, s->P, Qs)));
    V2_R(Ry, (v2_launch(k2_y_fwd<RR, 2>, dim3(tx, Q.nzs, 1), 16 * RR, Strided<RR>::smem(1), st, s->P, Qs)));
    if (poly == 2) V2_R(Rz, (v2_launch(k2_z_div<RR, 2>, dim3(tx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qs, 4)));
    else if (poly == 1) V2_R(Rz, (v2_launch(k2_z_div<RR, 1>, dim3(tx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qs, 4)));
    else V2_R(Rz, (v2_launch(k2_z_div<RR, 0>, dim3(tx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qs, 4)));
    V2_R(Ry, (v2_launch(k2_y_inv<RR>, dim3(tx, Q.Nz, 1), 16 * RR, Strided<RR>::smem(1), st, s->P, Qs)));
    V2_R(Rx, (v2_launch(k2_x_inv_real<RR>, dim3(gall), 256, smx, st, s->P, Qs, (const float2*)(Q.Z4 + 3 * Q.ZS), s->d_fk + (size_t)k * s->RS)));
  }
  LIFU_CUDA(cudaMemsetAsync(Q.Sslab, 0, sizeof(float) * (size_t)Q.nzs * s->N[1] * s->N[0], st));
  LIFU_CUDA(cudaGetLastError());
  Q.FK = s->d_fk; Q.qsrc = s->d_qsrc; Q.qcur = s->d_qcur;
  Q.t0s = t0; Q.nws = nw;
  return LIFU_OK;
}

// kind: 0 no source, 1 source active (generic path), 2 steady window (rank-2 source, k2_x_rho_p<SRC = 3>)
static int enqueue_step_v2(lifu_sim* s, int kind, int* n_kernels, const std::function<void(const char*, double)>& mark) {
  const V2Params& Q = s->Q;
  cudaStream_t st = s->stream;
  const int Rx = s->R[0], Ry = s->R[1], Rz = s->R[2];
  const unsigned tx = Q.nxt + 1;                                   // regular kx tiles + the Nyquist slot
  const int gx = (int)((long long)Q.Nz * (Q.Ny / 2) / (128 / Rx));   // row-pair batches of the persistent x kernels
  int nk = 0;
  auto ygrid = [&](unsigned planes, unsigned ncomp_) { return Q.zmajor ? dim3(tx, ncomp_, planes) : dim3(tx, planes, ncomp_); };
  const int src = kind == 0 ? 0 : (kind == 2 ? 3 : (s->source_mode == LIFU_SOURCE_ADDITIVE ? 1 : 2));
  const double srcf = (double)Q.nzs / Q.Nz;   // slab share of a full pass
  const int poly = s->P.poly_ok;
  // (1) pressure gradient
  V2_R(Ry, (v2_launch(k2_y_fwd<RR, 0>, ygrid(Q.Nz, 1), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
  ++nk; mark("k2_y_fwd_p", 8);
  // z passes: regular kx tiles through the persistent TMA-fed kernel, the Nyquist column through the per-thread-load
  // kernel on a 1-wide grid (zx = grid.x of that launch, Qn.bx0 selects the column)
  const bool ztma = s->z_tma;
  V2Params Qn = Q;
  if (ztma) Qn.bx0 = Q.nxt;
  const unsigned zx = ztma ? 1u : tx;
  const CUtensorMap& tmH = *reinterpret_cast<const CUtensorMap*>(s->tmH);
  const CUtensorMap& tmS = *reinterpret_cast<const CUtensorMap*>(s->tmS);
  const dim3 zgrid((unsigned)std::min(Q.nxt * Q.Ny, 2 * s->n_sm));
#define V2_ZTMA(OP, NCOMP)                                                                                                   \
  do {                                                                                                                        \
    if (poly == 2) V2_R(Rz, (v2_launch(k2_z_tma<RR, OP, 2>, zgrid, 16 * RR, ZTma<RR>::SMEM, st, s->P, Q, tmH, tmS, NCOMP)));   \
    else if (poly == 1) V2_R(Rz, (v2_launch(k2_z_tma<RR, OP, 1>, zgrid, 16 * RR, ZTma<RR>::SMEM, st, s->P, Q, tmH, tmS, NCOMP))); \
    else V2_R(Rz, (v2_launch(k2_z_tma<RR, OP, 0>, zgrid, 16 * RR, ZTma<RR>::SMEM, st, s->P, Q, tmH, tmS, NCOMP)));             \
    ++nk;                                                                                                                     \
  } while (0)
  if (ztma) V2_ZTMA(ZOP_GRAD, 1);
  if (poly == 2) V2_R(Rz, (v2_launch(k2_z_grad<RR, 2>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn)));
  else if (poly == 1) V2_R(Rz, (v2_launch(k2_z_grad<RR, 1>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn)));
  else V2_R(Rz, (v2_launch(k2_z_grad<RR, 0>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn)));
  ++nk; mark("k2_z_grad", 12);
  if (s->v2_ygrad_split) V2_R(Ry, (v2_launch(k2_y_inv_grad_split<RR>, dim3(tx, Q.Nz, 3), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
  else V2_R(Ry, (v2_launch(k2_y_inv_grad<RR>, dim3(tx, Q.Nz), 16 * RR, Strided<RR>::smem(2), st, s->P, Q)));
  ++nk; mark("k2_y_inv_grad", 20);
  // (2) velocity update + forward x transform of the new velocity
  V2_R(Rx, (v2_launch_x_u<RR>(s, gx)));
  ++nk; mark("k2_x_u", s->homogeneous ? 48 : 60);
  V2_R(Ry, (v2_launch(k2_y_fwd<RR, 1>, ygrid(Q.Nz, 3), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
  ++nk; mark("k2_y_fwd_u", 24);
  // (3) source field on its slab
  if (src == 1 || src == 2) {
    k2_source_scatter<<<grid_blocks(s, s->n_src, 128), 128, 0, st>>>(s->P, Q, s->S);
    ++nk; mark("k2_source_scatter", 0);
    if (src == 1) {
      const int gs = (int)(((long long)Q.nzs * (Q.Ny / 2) + (256 / Rx) - 1) / (256 / Rx));
      V2_R(Rx, (v2_launch(k2_x_src<RR>, dim3(gs), 256, (size_t)(256 / RR) * RR * RR * 8 + 16 * (RR * RR + 1), st, s->P, Q)));
      ++nk; mark("k2_x_src", 8 * srcf);
      V2_R(Ry, (v2_launch(k2_y_fwd<RR, 2>, ygrid(Q.nzs, 1), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
      ++nk; mark("k2_y_fwd_src", 8 * srcf);
    }
  }
  // (4) divergence (+ filtered source) through z and back through y
  const int ncomp = src == 1 ? 4 : 3;
  if (ztma) V2_ZTMA(ZOP_DIV, ncomp);
  if (poly == 2) V2_R(Rz, (v2_launch(k2_z_div<RR, 2>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn, ncomp)));
  else if (poly == 1) V2_R(Rz, (v2_launch(k2_z_div<RR, 1>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn, ncomp)));
  else V2_R(Rz, (v2_launch(k2_z_div<RR, 0>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn, ncomp)));
  ++nk; mark("k2_z_div", 24 + (src == 1 ? 4 + 4 * srcf : 0));
  V2_R(Ry, (v2_launch(k2_y_inv<RR>, ygrid(Q.Nz, ncomp), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
  ++nk; mark("k2_y_inv", 8 * ncomp);
  // (5) density update, source, equation of state, sensor, forward x transform of p
  if (src == 0) V2_R(Rx, (v2_launch_x_rho_p<RR, 0>(s, gx)));
  else if (src == 1) V2_R(Rx, (v2_launch_x_rho_p<RR, 1>(s, gx)));
  else if (src == 2) V2_R(Rx, (v2_launch_x_rho_p<RR, 2>(s, gx)));
  else V2_R(Rx, (v2_launch_x_rho_p<RR, 3>(s, gx)));
  ++nk;
  const double sens = (double)s->n[1] * s->n[2] / ((double)s->N[1] * s->N[2]);   // sensor rows are full x lines
  if (!s->absorbing) {
    mark("k2_x_rho_p", 12 + 24 + 16 * sens + 4 + (s->homogeneous ? 0 : 8) + (src == 1 ? 4 : (src == 3 ? 8 : 0)));
  } else {
    // (6) absorbing medium: the two fractional Laplacians, then the equation of state
    mark("k2_x_rho_abs", 12 + 24 + 4 + 8 + (s->homogeneous ? 0 : 8) + (src == 1 ? 4 : (src == 3 ? 8 : 0)));
    V2_R(Ry, (v2_launch(k2_y_fwd<RR, 3>, ygrid(Q.Nz, 2), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
    ++nk; mark("k2_y_fwd_abs", 16);
    if (ztma) V2_ZTMA(ZOP_ABS, 2);
    V2_R(Rz, (v2_launch(k2_z_absorb<RR>, dim3(zx, Q.Ny), 16 * RR, Strided<RR>::smem(2), st, s->P, Qn)));
    ++nk; mark("k2_z_absorb", 16);
    V2_R(Ry, (v2_launch(k2_y_inv<RR>, ygrid(Q.Nz, 2), 16 * RR, Strided<RR>::smem(1), st, s->P, Q)));
    ++nk; mark("k2_y_inv_abs", 16);
    V2_R(Rx, (v2_launch_x_p<RR>(s, gx)));
    ++nk; mark("k2_x_p", 8 + 4 + 16 * sens + 4 + (s->homogeneous ? 0 : 12));
  }
  LIFU_CUDA(cudaGetLastError());
  if (n_kernels) *n_kernels = nk;
  return LIFU_OK;
}

// ------------------------------------------------------------------------------------------
// pipeline v3: the same fused passes for any 2/3/5/7-smooth axis length (fft_gen.cuh)
static bool gen_factor(int n, int* radix, int* ns) {
  static const int cand[8] = {16, 9, 8, 7, 5, 4, 3, 2};
  const char* rm = getenv("LIFU_V3_RMAX");                  // tuning: largest radix allowed (e.g. 8)
  const int rmax = (rm && rm[0]) ? atoi(rm) : 16;
  int k = 0;
  for (int c = 0; c < 8 && n > 1; ++c) {
    const int r = cand[c];
    if (r > rmax) continue;
    while (n % r == 0 && n > 1) {
      if (r == 16 && n / 16 == 2) break;          // 32 -> 8 x 4 rather than 16 x 2
      if (k >= 8) return false;
      radix[k++] = r;
      n /= r;
    }
  }
  if (ns) *ns = k;
  return n == 1 && k >= 1;
}

static const size_t kV3SmemCap = 200 * 1024;

static bool v3_eligible(const lifu_sim* s) {
  if (s->pipeline == 1) return false;
  int rad[8], ns;
  for (int a = 0; a < 3; ++a) {
    if (s->N[a] < 4 || !gen_factor(s->N[a], rad, &ns)) return false;
    if ((size_t)2 * s->N[a] * 3 * sizeof(float2) > kV3SmemCap) return false;     // two tile buffers of two lanes must fit
  }
  return true;
}

static bool v3_auto(const lifu_sim* s) {
  const char* e = getenv("LIFU_V3_AUTO");
  if (e) return e[0] == '1';
  return false;
}

static int v3_lanes(int n, int nbuf, size_t budget, int lmax) {
  int L = lmax;
  while (L > 1 && (size_t)nbuf * n * (L + 1) * sizeof(float2) > budget) L >>= 1;
  return L;
}

static int v3_setup(lifu_sim* s) {
  GParams& G = s->G;
  const int Nx = s->N[0], Ny = s->N[1], Nz = s->N[2];
  if (!s->v3_ready) {
    G.Nx = Nx; G.Ny = Ny; G.Nz = Nz; G.Nxh = s->Nxh;
    G.PH = (int)round_up(s->Nxh, 4);
    G.My = (Ny + 1) / 2;
    G.HS = (long long)Nz * Ny * G.PH;
    G.ZS = (long long)Nz * G.My * Nx;
    G.norm = (float)(1.0 / (2.0 * (double)s->V));
    GenPlan* pl[3] = {&G.px, &G.py, &G.pz};
    for (int a = 0; a < 3; ++a) {
      const int n = s->N[a];
      pl[a]->N = n;
      if (!gen_factor(n, pl[a]->radix, &pl[a]->ns)) { set_error("v3: axis length %d is not 2/3/5/7-smooth", n); return LIFU_ERR_STATE; }
      std::vector<float2> tw(n);
      for (int m = 0; m < n; ++m) {
        const double ang = -2.0 * M_PI * (double)m / (double)n;
        tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
      }
      LIFU_CHECK(dev_alloc(s, (void**)&s->d_gtw[a], sizeof(float2) * n));
      LIFU_CUDA(cudaMemcpyAsync(s->d_gtw[a], tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, s->stream));
      LIFU_CUDA(cudaStreamSynchronize(s->stream));
      pl[a]->tw = s->d_gtw[a];
    }
    LIFU_CHECK(dev_alloc(s, (void**)&G.ZP, sizeof(float2) * G.ZS));
    LIFU_CHECK(dev_alloc(s, (void**)&G.Z4, sizeof(float2) * 4 * G.ZS));
    LIFU_CHECK(dev_alloc(s, (void**)&G.H4, sizeof(float2) * 4 * G.HS));
    LIFU_CHECK(dev_alloc(s, (void**)&G.pm, sizeof(float2) * s->V));
    // tile widths: as wide as two CTAs per SM allow (strided passes: two buffers; x passes: up to five)
    s->v3_ready = true;
  }
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return (e && e[0]) ? atoi(e) : dflt; };
  G.Ls = v3_lanes(std::max(Ny, Nz), 2, 108 * 1024, 16);
  {
    const int o = env_int("LIFU_V3_LS", 0);                     // tuning override: lanes per strided tile (power of two)
    if (o == 1 || o == 2 || o == 4 || o == 8 || o == 16) G.Ls = v3_lanes(std::max(Ny, Nz), 2, kV3SmemCap, o);
  }
  G.lsh_s = 0; while ((1 << G.lsh_s) < G.Ls) ++G.lsh_s;
  // the x passes hold two tile buffers as well (what outlives a transform is parked in global scratch rows)
  G.Lx = v3_lanes(Nx, 2, 108 * 1024, 16);
  {
    const int o = env_int("LIFU_V3_LX", 0);
    if (o == 1 || o == 2 || o == 4 || o == 8 || o == 16) G.Lx = v3_lanes(Nx, 2, kV3SmemCap, o);
  }
  G.lsh_x = 0; while ((1 << G.lsh_x) < G.Lx) ++G.lsh_x;
  s->v3_ts = std::min(512, std::max(64, env_int("LIFU_V3_TS", 256) / 32 * 32));
  s->v3_tx = std::min(512, std::max(64, env_int("LIFU_V3_TX", 256) / 32 * 32));
  int z0 = 0, nz = 1;
  if (s->n_src > 0) {
    long long first = 0, last = 0;
    LIFU_CUDA(cudaMemcpyAsync(&first, s->d_lin_exp, sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    LIFU_CUDA(cudaMemcpyAsync(&last, s->d_lin_exp + (s->n_src - 1), sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    LIFU_CUDA(cudaStreamSynchronize(s->stream));
    const long long plane = (long long)Nx * Ny;
    z0 = (int)(first / plane);
    nz = (int)(last / plane) - z0 + 1;
  }
  if (nz > s->v3_slab_planes) {
    LIFU_CHECK(dev_alloc(s, (void**)&G.Sslab, sizeof(float) * (size_t)nz * Ny * Nx));
    LIFU_CHECK(dev_alloc(s, (void**)&G.ZSslab, sizeof(float2) * (size_t)nz * G.My * Nx));
    LIFU_CHECK(dev_alloc(s, (void**)&G.HSslab, sizeof(float2) * (size_t)nz * Ny * G.PH));
    s->v3_slab_planes = nz;
  }
  G.z0s = z0; G.nzs = nz;
  G.store_p = 0;
  const char* pa = getenv("LIFU_PM_ALWAYS");
  G.pm_always = (pa && pa[0] == '1') ? 1 : 0;
  return LIFU_OK;
}

static int enqueue_step_v3(lifu_sim* s, bool src_active, int* n_kernels, const std::function<void(const char*, double)>& mark) {
  const GParams& G = s->G;
  cudaStream_t st = s->stream;
  const unsigned tx = (unsigned)((G.Nxh + G.Ls - 1) / G.Ls);
  const size_t smy = (size_t)2 * G.Ny * (G.Ls + 1) * sizeof(float2);
  const size_t smz = (size_t)2 * G.Nz * (G.Ls + 1) * sizeof(float2);
  const size_t smx = (size_t)G.Nx * (G.Lx + 1) * sizeof(float2);      // one x tile buffer
  const unsigned gx = (unsigned)(((long long)G.Nz * G.My + G.Lx - 1) / G.Lx);
  const int src = !src_active ? 0 : (s->source_mode == LIFU_SOURCE_ADDITIVE ? 1 : 2);
  const double srcf = (double)G.nzs / G.Nz;
  const int TS = s->v3_ts, TX = s->v3_tx;                              // threads per CTA of the strided / x passes
  int nk = 0;
  // (1) pressure gradient
  v2_launch(g3_y_fwd<0>, dim3(tx, G.Nz, 1), TS, smy, st, s->P, G); ++nk; mark("g3_y_fwd_p", 8);
  v2_launch(g3_z_grad, dim3(tx, G.Ny, 2), TS, smz, st, s->P, G); ++nk; mark("g3_z_grad", 12);
  v2_launch(g3_y_inv_grad, dim3(tx, G.Nz, 3), TS, smy, st, s->P, G); ++nk; mark("g3_y_inv_grad", 20);
  // (2) velocity update + forward x transform of the new velocity
  if (s->homogeneous) v2_launch(g3_x_u<true>, dim3(gx, 3), TX, 2 * smx, st, s->P, G, G.Lx, G.lsh_x);
  else v2_launch(g3_x_u<false>, dim3(gx, 3), TX, 2 * smx, st, s->P, G, G.Lx, G.lsh_x);
  ++nk; mark("g3_x_u", s->homogeneous ? 48 : 60);
  v2_launch(g3_y_fwd<1>, dim3(tx, G.Nz, 3), TS, smy, st, s->P, G); ++nk; mark("g3_y_fwd_u", 24);
  // (3) source field on its slab
  if (src != 0) {
    g3_source_scatter<<<grid_blocks(s, s->n_src, 128), 128, 0, st>>>(s->P, G, s->S);
    ++nk; mark("g3_source_scatter", 0);
    if (src == 1) {
      const unsigned gs = (unsigned)(((long long)G.nzs * G.My + G.Lx - 1) / G.Lx);
      v2_launch(g3_x_src, dim3(gs), TX, 2 * smx, st, s->P, G, G.Lx, G.lsh_x); ++nk; mark("g3_x_src", 8 * srcf);
      v2_launch(g3_y_fwd<2>, dim3(tx, G.nzs, 1), TS, smy, st, s->P, G); ++nk; mark("g3_y_fwd_src", 8 * srcf);
    }
  }
  // (4) divergence (+ filtered source) through z and back through y
  const int ncomp = src == 1 ? 4 : 3;
  v2_launch(g3_z_pass<0>, dim3(tx, G.Ny, ncomp), TS, smz, st, s->P, G); ++nk;
  mark("g3_z_div", 24 + (src == 1 ? 4 + 4 * srcf : 0));
  v2_launch(g3_y_inv, dim3(tx, G.Nz, ncomp), TS, smy, st, s->P, G); ++nk; mark("g3_y_inv", 8 * ncomp);
  // (5) density update, source, equation of state, sensor, forward x transform of p
#define V3_RHO(H, SRCV, A) v2_launch(g3_x_rho_p<H, SRCV, A>, dim3(gx), TX, 2 * smx, st, s->P, G, G.Lx, G.lsh_x)
#define V3_RHO_SRC(H, A) do { if (src == 0) V3_RHO(H, 0, A); else if (src == 1) V3_RHO(H, 1, A); else V3_RHO(H, 2, A); } while (0)
  if (s->homogeneous) { if (s->absorbing) V3_RHO_SRC(true, true); else V3_RHO_SRC(true, false); }
  else { if (s->absorbing) V3_RHO_SRC(false, true); else V3_RHO_SRC(false, false); }
#undef V3_RHO_SRC
#undef V3_RHO
  ++nk;
  const double sens = (double)s->Vin / (double)s->V;
  if (!s->absorbing) {
    mark("g3_x_rho_p", 12 + 24 + 16 * sens + 4 + (s->homogeneous ? 0 : 8) + (src == 1 ? 4 : 0));
  } else {
    mark("g3_x_rho_abs", 12 + 24 + 4 + 8 + (s->homogeneous ? 0 : 8) + (src == 1 ? 4 : 0));
    v2_launch(g3_y_fwd<3>, dim3(tx, G.Nz, 2), TS, smy, st, s->P, G); ++nk; mark("g3_y_fwd_abs", 16);
    v2_launch(g3_z_pass<1>, dim3(tx, G.Ny, 2), TS, smz, st, s->P, G); ++nk; mark("g3_z_absorb", 16);
    v2_launch(g3_y_inv, dim3(tx, G.Nz, 2), TS, smy, st, s->P, G); ++nk; mark("g3_y_inv_abs", 16);
    const int use_tau = s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION, use_eta = s->alpha_mode != LIFU_ALPHA_NO_DISPERSION;
    if (s->homogeneous) v2_launch(g3_x_p<true>, dim3(gx), TX, 2 * smx, st, s->P, G, G.Lx, G.lsh_x, use_tau, use_eta);
    else v2_launch(g3_x_p<false>, dim3(gx), TX, 2 * smx, st, s->P, G, G.Lx, G.lsh_x, use_tau, use_eta);
    ++nk; mark("g3_x_p", 8 + 4 + 16 * sens + 4 + (s->homogeneous ? 0 : 12));
  }
  LIFU_CUDA(cudaGetLastError());
  if (n_kernels) *n_kernels = nk;
  return LIFU_OK;
}

// ------------------------------------------------------------------------------------------
// one time step of a z-slab decomposed grid (slab.cuh): same k-Wave step order (ledger A8), every 3-D
// transform split into local 2-D transforms, an exchange over NVLink and local 1-D transforms.
static int enqueue_step_slab(lifu_sim* s, bool src_active, int* n_kernels, int* n_ffts,
                             const std::function<void(const char*, double)>& mark) {
  StepParams& P = s->P;
  SlabCtx& L = s->sl;
  cudaStream_t st = s->stream, xs = L.xs;
  LIFU_CHECK(slab_plans(s));
  LIFU_CUFFT(cufftSetStream(L.r2c2d, st)); LIFU_CUFFT(cufftSetStream(L.c2r2d, st)); LIFU_CUFFT(cufftSetStream(L.c2c1d, st));
  const SlabParams S = SlabHost::params(s, false);
  float2* H = S.H; float2* T = S.T;
  const long long Hl = L.Hl;
  const int gbh = grid_blocks(s, Hl, 256);
  const int gbr = grid_blocks(s, s->Vloc, 256);
  int nk = 0, nf = 0, ne = 0, rc = LIFU_OK;
  const int xk = L.exchange == 2 ? 1 : 2;                       // kernels per exchange
  auto fft2_r2c = [&](float* r, float2* c) { ++nf; return cufftExecR2C(L.r2c2d, r, (cufftComplex*)c); };
  auto fft2_c2r = [&](float2* c, float* r) { ++nf; return cufftExecC2R(L.c2r2d, (cufftComplex*)c, r); };
  auto fftz = [&](float2* c, int dir) { ++nf; return cufftExecC2C(L.c2c1d, (cufftComplex*)c, (cufftComplex*)c, dir); };
  // Hand-offs between the compute stream and the exchange stream.  An exchange of field f is: wait for its producer
  // on `st`, push + barrier on `xs`, and an event the consumer on `st` waits for right before it needs the field --
  // so the transforms of field f+1 run while field f is on the wire.
  auto hand = [&](cudaStream_t from, cudaStream_t to) {
    cudaEvent_t e = L.ev[ne++ & 31];
    if (cudaEventRecord(e, from) != cudaSuccess || cudaStreamWaitEvent(to, e, 0) != cudaSuccess) rc = LIFU_ERR_CUDA;
  };
  cudaEvent_t done[4];                                          // "exchange of field f finished", consumed on st
  const bool ov = L.overlap;
  if (!ov) xs = st;                                             // batched mode: one stream, one barrier per batch of fields
  auto xfwd = [&](int f, int kind, bool last) {
    if (ov) hand(st, xs);
    if (rc == LIFU_OK) rc = slab_exchange_fwd(s, xs, f, kind, ov || last);
    if (ov) { done[f] = L.ev[ne++ & 31]; if (cudaEventRecord(done[f], xs) != cudaSuccess) rc = LIFU_ERR_CUDA; }
    nk += xk;
  };
  auto xback = [&](int f, int fd, bool last) {
    if (ov) hand(st, xs);
    if (rc == LIFU_OK) rc = slab_exchange_back(s, xs, f, fd, ov || last);
    if (ov) { done[f] = L.ev[ne++ & 31]; if (cudaEventRecord(done[f], xs) != cudaSuccess) rc = LIFU_ERR_CUDA; }
    nk += xk;
  };
  auto need = [&](int f) { if (ov && cudaStreamWaitEvent(st, done[f], 0) != cudaSuccess) rc = LIFU_ERR_CUDA; };

  // (1) pressure gradient: 1 field out, 2 fields back
  LIFU_CUFFT(fft2_r2c(P.p, H));
  xfwd(0, 0, true);
  need(0);
  mark("r2c_p+xchg", 16);
  LIFU_CUFFT(fftz(T, CUFFT_FORWARD));
  k_slab_grad_z<<<gbh, 256, 0, st>>>(P, S); ++nk;
  LIFU_CUFFT(fftz(T, CUFFT_INVERSE));
  xback(0, 3, false);                                              // T0 -> H3 (kappa p^: all of the x / y content)
  LIFU_CUFFT(fftz(T + Hl, CUFFT_INVERSE));
  xback(1, 2, true);                                              // T1 -> H2 (d/dz)
  mark("z_grad", 4 + 12 + 16);
  need(0);
  k_slab_grad_xy<<<gbh, 256, 0, st>>>(P, S); ++nk;
  LIFU_CUFFT(fft2_c2r(H, P.r3));
  LIFU_CUFFT(fft2_c2r(H + Hl, P.r3 + P.RS));
  need(1);
  LIFU_CUFFT(fft2_c2r(H + 2 * Hl, P.r3 + 2 * P.RS));
  mark("xchg_back_grad+xy+c2r", 16 + 12 + 24);
  if (s->N[0] % 4 == 0) launch_update_u<4>(s, grid_blocks(s, s->Vloc / 4, 256));
  else launch_update_u<1>(s, gbr);
  ++nk;
  mark("k_update_u", s->homogeneous ? 36 : 48);
  // (2) source field (built before the divergence so that it rides the same pipeline)
  int src = 0;
  if (src_active) {
    if (L.src_i1 > L.src_i0) {
      k_source_scatter<<<grid_blocks(s, L.src_i1 - L.src_i0, 128), 128, 0, st>>>(P, s->S); ++nk;
    }
    mark("k_source_scatter", 0);
    src = s->source_mode == LIFU_SOURCE_ADDITIVE ? 1 : 2;
  }
  // (3) velocity divergence (+ k-space filtered source): 3 (4) fields out and back, pipelined field by field
  const int nfld = src == 1 ? 4 : 3;
  for (int c = 0; c < nfld; ++c) {
    LIFU_CUFFT(fft2_r2c(c < 3 ? P.u + c * P.RS : P.S, H + c * Hl));
    xfwd(c, c == 0 ? 1 : (c == 1 ? 2 : 0), c == nfld - 1);
  }
  mark("r2c_u", 8 * nfld);
  for (int c = 0; c < nfld; ++c) {
    need(c);
    LIFU_CUFFT(fftz(T + c * Hl, CUFFT_FORWARD));
    if (c < 2) k_slab_div_z<0><<<gbh, 256, 0, st>>>(P, S, c);
    else if (c == 2) k_slab_div_z<1><<<gbh, 256, 0, st>>>(P, S, c);
    else k_slab_div_z<2><<<gbh, 256, 0, st>>>(P, S, c);
    ++nk;
    LIFU_CUFFT(fftz(T + c * Hl, CUFFT_INVERSE));
    xback(c, c, c == nfld - 1);
  }
  mark("xchg_fwd_u+z_div", 8 * nfld + 16 * nfld);
  for (int c = 0; c < nfld; ++c) {
    need(c);
    LIFU_CUFFT(fft2_c2r(H + c * Hl, c < 3 ? P.r3 + c * P.RS : P.Sf));
  }
  mark("xchg_back_div+c2r", 16 * nfld);
  // (4) density update, source, equation of state, sensor
  if (s->homogeneous) {
    if (src == 0) launch_rho_p<true, 0>(s, gbr); else if (src == 1) launch_rho_p<true, 1>(s, gbr); else launch_rho_p<true, 2>(s, gbr);
  } else {
    if (src == 0) launch_rho_p<false, 0>(s, gbr); else if (src == 1) launch_rho_p<false, 1>(s, gbr); else launch_rho_p<false, 2>(s, gbr);
  }
  ++nk;
  mark("k_update_rho_p", (s->homogeneous ? 56 : 64) + (src ? 4 : 0));
  if (s->absorbing) {
    for (int c = 0; c < 2; ++c) {
      LIFU_CUFFT(fft2_r2c(P.r3 + c * P.RS, H + c * Hl));
      xfwd(c, 0, c == 1);
    }
    for (int c = 0; c < 2; ++c) {
      need(c);
      LIFU_CUFFT(fftz(T + c * Hl, CUFFT_FORWARD));
      k_slab_absorb_z<<<gbh, 256, 0, st>>>(P, S, c, c == 0 ? P.y_minus2_half : P.y_minus1_half); ++nk;
      LIFU_CUFFT(fftz(T + c * Hl, CUFFT_INVERSE));
      xback(c, c, c == 1);
    }
    mark("absorb_fwd+z", 16 + 16 + 32);
    for (int c = 0; c < 2; ++c) {
      need(c);
      LIFU_CUFFT(fft2_c2r(H + c * Hl, P.r3 + c * P.RS));
    }
    mark("xchg_back_absorb+c2r", 16 + 16);
    const int use_tau = s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION, use_eta = s->alpha_mode != LIFU_ALPHA_NO_DISPERSION;
    if (s->homogeneous) k_pressure_absorb<true><<<gbr, 256, 0, st>>>(P, use_tau, use_eta);
    else k_pressure_absorb<false><<<gbr, 256, 0, st>>>(P, use_tau, use_eta);
    ++nk;
    mark("k_pressure_absorb", s->homogeneous ? 32 : 44);
  }
  if (rc != LIFU_OK) { if (g_err.empty()) set_error("slab step: stream/event hand-off failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
  LIFU_CUDA(cudaGetLastError());
  if (n_kernels) *n_kernels = nk;
  if (n_ffts) *n_ffts = nf;
  return LIFU_OK;
}

// Enqueue one time step on the handle's stream.  Counts hand-written kernels / FFT executions.
// kind: 0 no source, 1 source active, 2 steady window of the v2 pipeline (v2_build_steady)
static int enqueue_step(lifu_sim* s, int kind, int* n_kernels, int* n_ffts) {
  const bool src_active = kind != 0;
  StepParams& P = s->P;
  cudaStream_t st = s->stream;
  // optional per-stage timing (lifu_profile_stages): an event after every stage
  auto mark = [&](const char* name, double bytes_per_voxel) {
    if (!s->prof_on) return;
    cudaEvent_t e = nullptr;
    if (s->prof_used < (int)s->prof_ev.size()) e = s->prof_ev[s->prof_used];
    else { cudaEventCreate(&e); s->prof_ev.push_back(e); }
    cudaEventRecord(e, st);
    if (s->prof_used >= (int)s->prof_names.size()) { s->prof_names.push_back(name); s->prof_bytes.push_back(bytes_per_voxel); }
    ++s->prof_used;
  };
  mark("begin", 0);
  if (s->sl.on && !s->last_used_v2) return enqueue_step_slab(s, src_active, n_kernels, n_ffts, mark);
  if (s->last_used_v2) {
    std::function<int()> bar;
    if (s->sl.on) bar = [s]() { return slab_barrier(s, s->stream); };
    int rc2 = s->v2_wide ? wide_enqueue_step(s, kind, n_kernels, mark, bar) : enqueue_step_v2(s, kind, n_kernels, mark);
    if (n_ffts) *n_ffts = 0;
    return rc2;
  }
  if (s->last_used_v3) {
    int rc3 = enqueue_step_v3(s, src_active, n_kernels, mark);
    if (n_ffts) *n_ffts = 0;
    return rc3;
  }
  const int gbh = grid_blocks(s, s->Vh, 256);
  const int gbr = grid_blocks(s, s->V, 256);
  int nk = 0, nf = 0;
  // (1) grad p -> u
  LIFU_CUFFT(cufftExecR2C(s->r2c1, P.p, (cufftComplex*)P.c1)); nf += 1;
  mark("cufft_r2c_p", 8);
  k_grad_spectral<<<gbh, 256, 0, st>>>(P); ++nk;
  mark("k_grad_spectral", 16);
  LIFU_CUFFT(cufftExecC2R(s->c2r3, (cufftComplex*)P.c3, P.r3)); nf += 3;
  mark("cufft_c2r_grad_x3", 24);
  if (s->N[0] % 4 == 0) launch_update_u<4>(s, grid_blocks(s, s->V / 4, 256));
  else launch_update_u<1>(s, gbr);
  ++nk;
  mark("k_update_u", s->homogeneous ? 36 : 48);
  // (2) div u
  LIFU_CUFFT(cufftExecR2C(s->r2c3, P.u, (cufftComplex*)P.c3)); nf += 3;
  mark("cufft_r2c_u_x3", 24);
  k_div_spectral<<<gbh, 256, 0, st>>>(P); ++nk;
  mark("k_div_spectral", 24);
  LIFU_CUFFT(cufftExecC2R(s->c2r3, (cufftComplex*)P.c3, P.r3)); nf += 3;
  mark("cufft_c2r_div_x3", 24);
  // (4) source field
  int src = 0;
  if (src_active) {
    k_source_scatter<<<grid_blocks(s, s->n_src, 128), 128, 0, st>>>(P, s->S); ++nk;
    mark("k_source_scatter", 0);
    if (s->source_mode == LIFU_SOURCE_ADDITIVE) {
      LIFU_CUFFT(cufftExecR2C(s->r2c1, P.S, (cufftComplex*)P.c1)); nf += 1;
      mark("cufft_r2c_src", 8);
      k_source_filter<<<gbh, 256, 0, st>>>(P); ++nk;
      mark("k_source_filter", 8);
      LIFU_CUFFT(cufftExecC2R(s->c2r1, (cufftComplex*)P.c1, P.Sf)); nf += 1;
      mark("cufft_c2r_src", 8);
      src = 1;
    } else {
      src = 2;
    }
  }
  // (3)+(5)+(6) rho update, equation of state, sensor
  if (s->homogeneous) {
    if (src == 0) launch_rho_p<true, 0>(s, gbr); else if (src == 1) launch_rho_p<true, 1>(s, gbr); else launch_rho_p<true, 2>(s, gbr);
  } else {
    if (src == 0) launch_rho_p<false, 0>(s, gbr); else if (src == 1) launch_rho_p<false, 1>(s, gbr); else launch_rho_p<false, 2>(s, gbr);
  }
  ++nk;
  mark("k_update_rho_p", (s->homogeneous ? 56 : 64) + (src ? 4 : 0));
  if (s->absorbing) {
    LIFU_CUFFT(cufftExecR2C(s->r2c2, P.r3, (cufftComplex*)P.c3)); nf += 2;
    mark("cufft_r2c_absorb_x2", 16);
    k_absorb_spectral<<<gbh, 256, 0, st>>>(P); ++nk;
    mark("k_absorb_spectral", 16);
    LIFU_CUFFT(cufftExecC2R(s->c2r2, (cufftComplex*)P.c3, P.r3)); nf += 2;
    mark("cufft_c2r_absorb_x2", 16);
    const int use_tau = s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION, use_eta = s->alpha_mode != LIFU_ALPHA_NO_DISPERSION;
    if (s->homogeneous) k_pressure_absorb<true><<<gbr, 256, 0, st>>>(P, use_tau, use_eta);
    else k_pressure_absorb<false><<<gbr, 256, 0, st>>>(P, use_tau, use_eta);
    ++nk;
    mark("k_pressure_absorb", s->homogeneous ? 32 : 44);
  }
  LIFU_CUDA(cudaGetLastError());
  if (n_kernels) *n_kernels = nk;
  if (n_ffts) *n_ffts = nf;
  return LIFU_OK;
}

static double bytes_model(const lifu_sim* s, bool src_active) {
  // (steps of the steady window are counted like any other source-active step: the model is the reference algorithm's)
  // algorithmic bytes per voxel-step (DESIGN.md section 4 / SURVEY.md 8d)
  double ffts = 10, k1 = 16, k3 = 24;
  double k2 = s->homogeneous ? 36 : 48;
  double k4 = s->homogeneous ? 56 : 64;
  double b = k1 + k2 + k3 + k4;
  if (s->absorbing) {
    if (s->alpha_mode != LIFU_ALPHA_NO_ABSORPTION) { ffts += 2; b += s->homogeneous ? 16 : 20; }
    if (s->alpha_mode != LIFU_ALPHA_NO_DISPERSION) { ffts += 2; b += 20; }
  }
  if (src_active && s->source_mode == LIFU_SOURCE_ADDITIVE) { ffts += 2; b += 8; }
  return b + 8.0 * ffts;
}

}  // namespace lifu

extern "C" {

int lifu_run(lifu_sim* s, float* p_max, float* p_min, lifu_stats* stats) {
  if (!s) { set_error("lifu_run: null handle"); return LIFU_ERR_INVALID; }
  if (!s->medium_set) { set_error("lifu_run: call lifu_set_medium first"); return LIFU_ERR_STATE; }
  if (!s->geometry_set) { set_error("lifu_run: call lifu_set_elements or lifu_set_source_geometry first"); return LIFU_ERR_STATE; }
  if (!s->drive_set) { set_error("lifu_run: call lifu_set_drive first"); return LIFU_ERR_STATE; }
  if (s->drive_n_el != s->n_el) {
    set_error("lifu_run: drive has %d elements but the source geometry has %d", s->drive_n_el, s->n_el);
    return LIFU_ERR_STATE;
  }
  if (!s->tables_ready) { set_error("lifu_run: internal error, tables not built"); return LIFU_ERR_STATE; }
  NvtxRange nvtx_run("lifu_run");
  nvtxRangePushA("lifu_run: set-up");
  struct PopOnExit { int n = 1; ~PopOnExit() { for (int i = 0; i < n; ++i) nvtxRangePop(); } } nvtx_phase;
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t user_stream = s->stream;
  cudaStream_t own = nullptr;
  if (s->stream == nullptr) {   // stream capture needs a non-default stream
    LIFU_CUDA(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
    s->stream = own;
  }
  struct Restore {
    lifu_sim* s; cudaStream_t user, own;
    ~Restore() {
      if (own) {
        cudaStreamSynchronize(own);
        cufftHandle hs[6] = {s->r2c1, s->r2c3, s->c2r1, s->c2r3, s->r2c2, s->c2r2};
        if (s->plans_ready) for (cufftHandle h : hs) cufftSetStream(h, user);
        cudaStreamDestroy(own);
      }
      s->stream = user;
    }
  } restore{s, user_stream, own};
  cudaStream_t st = s->stream;
  if (s->pipeline == 2 && !v2_eligible(s)) {
    set_error("lifu_run: LIFU_PIPELINE=v2 needs 64 / 128 / 256 / 512 / 768 / 1024-point axes (grid is %dx%dx%d)",
              s->N[0], s->N[1], s->N[2]);
    return LIFU_ERR_STATE;
  }
  if (s->pipeline == 3 && !v3_eligible(s)) {
    set_error("lifu_run: LIFU_PIPELINE=v3 needs 2/3/5/7-smooth axes (grid is %dx%dx%d)", s->N[0], s->N[1], s->N[2]);
    return LIFU_ERR_STATE;
  }
  s->last_used_v2 = s->sl.on ? slab_wide_ok(s) : v2_eligible(s);
  // v3 (generic radices) is taken when asked for (LIFU_PIPELINE=v3) or, automatically, on the grids where it has been
  // measured faster than the library-FFT pipeline (profiles/r2_v3_summary.md): LIFU_V3_AUTO=0 / 1 overrides
  s->last_used_v3 = !s->sl.on && !s->last_used_v2 && v3_eligible(s) && (s->pipeline == 3 || v3_auto(s));
  if (s->sl.on && !s->last_used_v2) {
    LIFU_CHECK(slab_plans(s));
  } else if (s->last_used_v2) {
    LIFU_CHECK(v2_setup(s));
  } else if (s->last_used_v3) {
    LIFU_CHECK(v3_setup(s));
  } else {
    LIFU_CHECK(build_plans(s));
    cufftHandle hs[6] = {s->r2c1, s->r2c3, s->c2r1, s->c2r3, s->r2c2, s->c2r2};
    for (cufftHandle h : hs) LIFU_CUFFT(cufftSetStream(h, st));
  }
  StepParams& P = s->P;
  SourceParams& S = s->S;
  S.n_src = s->sl.src_i1 - s->sl.src_i0; S.lin_exp = s->d_lin_exp; S.row_ptr = s->d_row_ptr + s->sl.src_i0; S.col = s->d_col; S.w = s->d_w;
  S.scale = s->d_scale; S.base = s->d_base; S.n_base = s->n_base; S.delay = s->d_delay; S.gain = s->d_gain;

  LIFU_CUDA(cudaEventRecord(s->ev[0], st));
  const size_t R = sizeof(float) * s->RS;
  LIFU_CUDA(cudaMemsetAsync(P.p, 0, R, st));
  LIFU_CUDA(cudaMemsetAsync(P.u, 0, 3 * R, st));
  LIFU_CUDA(cudaMemsetAsync(P.rho, 0, 3 * R, st));
  LIFU_CUDA(cudaMemsetAsync(P.S, 0, R, st));
  LIFU_CUDA(cudaMemsetAsync(P.step, 0, sizeof(int), st));
  if (s->last_used_v2) {
    LIFU_CUDA(cudaMemsetAsync(s->Q.ZP, 0, sizeof(float2) * s->Q.ZS, st));
    LIFU_CUDA(cudaMemsetAsync(s->Q.Sslab, 0, sizeof(float) * (size_t)s->Q.nzs * s->N[1] * s->N[0], st));
    k2_pm_init<<<grid_blocks(s, s->Vloc, 256), 256, 0, st>>>(s->Q.pm, s->Vloc);
  }
  if (s->last_used_v3) {
    LIFU_CUDA(cudaMemsetAsync(s->G.ZP, 0, sizeof(float2) * s->G.ZS, st));
    LIFU_CUDA(cudaMemsetAsync(s->G.Sslab, 0, sizeof(float) * (size_t)s->G.nzs * s->N[1] * s->N[0], st));
    k2_pm_init<<<grid_blocks(s, s->V, 256), 256, 0, st>>>(s->G.pm, s->V);
  }
  k_fill<<<grid_blocks(s, s->Vsens, 256), 256, 0, st>>>(P.pmax, s->Vsens, -INFINITY);
  k_fill<<<grid_blocks(s, s->Vsens, 256), 256, 0, st>>>(P.pmin, s->Vsens, INFINITY);
  LIFU_CUDA(cudaGetLastError());

  const int nt = s->grid.nt;
  const int L = s->n_src > 0 ? std::min(nt, s->max_delay + s->n_base) : 0;
  nvtxRangePop(); nvtxRangePushA("lifu_run: steady-source basis + graph capture");
  // steady window of the source (v2 pipeline): steps [w0, w1) run the rank-2 source path
  int w0 = 0, w1 = 0;
  if (s->last_used_v2 && !s->v2_wide && L > 0) {
    LIFU_CHECK(v2_build_steady(s, nt));
    if (s->Q.nws > 0) { w0 = s->Q.t0s; w1 = s->Q.t0s + s->Q.nws; }
  }
  auto kind_of = [&](int t) { return t >= L ? 0 : ((t >= w0 && t < w1) ? 2 : 1); };
  int nk[3] = {0, 0, 0}, nf[3] = {0, 0, 0};
  long long cnt[3] = {0, 0, 0};
  for (int t = 0; t < nt; ++t) ++cnt[kind_of(t)];
  // capture one step of every kind that occurs as a CUDA graph (launch-bound at C1 sizes)
  cudaGraphExec_t gexec[3] = {nullptr, nullptr, nullptr};
  bool graphs = s->use_graph;
  if (graphs) {
    for (int v = 0; v < 3 && graphs; ++v) {
      if (cnt[v] == 0) continue;
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { graphs = false; break; }
      int rc = enqueue_step(s, v, &nk[v], &nf[v]);
      cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc != LIFU_OK || ce != cudaSuccess || g == nullptr) { graphs = false; if (g) cudaGraphDestroy(g); break; }
      ce = cudaGraphInstantiate(&gexec[v], g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { graphs = false; gexec[v] = nullptr; break; }
    }
    if (!graphs) {
      cudaGetLastError();   // clear the sticky-free capture error and run without graphs
      for (int v = 0; v < 3; ++v) if (gexec[v]) { cudaGraphExecDestroy(gexec[v]); gexec[v] = nullptr; }
    }
  }
  LIFU_CUDA(cudaEventRecord(s->ev[1], st));
  nvtxRangePop(); nvtxRangePushA("lifu_run: time loop");
  int rc = LIFU_OK;
  for (int t = 0; t < nt && rc == LIFU_OK; ++t) {
    const int v = kind_of(t);
    if ((s->last_used_v2 || s->last_used_v3) && t == nt - 1) {   // last step also materialises the real-space pressure
      s->Q.store_p = 1; s->G.store_p = 1;
      int k1 = 0, f1 = 0;
      rc = enqueue_step(s, v, &k1, &f1);
      s->Q.store_p = 0; s->G.store_p = 0;
      if (nk[v] == 0) { nk[v] = k1; nf[v] = f1; }
    } else if (graphs) {
      if (cudaGraphLaunch(gexec[v], st) != cudaSuccess) { set_error("cudaGraphLaunch failed at step %d: %s", t, cudaGetErrorString(cudaGetLastError())); rc = LIFU_ERR_CUDA; }
    } else {
      rc = enqueue_step(s, v, &nk[v], &nf[v]);
    }
  }
  if (rc == LIFU_OK && cudaEventRecord(s->ev[2], st) != cudaSuccess) rc = LIFU_ERR_CUDA;
  nvtxRangePop(); nvtxRangePushA("lifu_run: crop + read-back + wait");
  if (rc == LIFU_OK && s->last_used_v2) {
    if (s->v2_wide) { if (s->Vsens > 0) wide_pm_crop(s); }
    else k2_pm_crop<<<grid_blocks(s, s->Vin, 256), 256, 0, st>>>(s->P, s->Q);
  }
  if (rc == LIFU_OK && s->last_used_v3) g3_pm_crop<<<grid_blocks(s, s->Vin, 256), 256, 0, st>>>(s->P, s->G.pm);
  if (rc == LIFU_OK && p_max && s->Vsens) if (cudaMemcpyAsync(p_max, P.pmax, sizeof(float) * s->Vsens, cudaMemcpyDefault, st) != cudaSuccess) rc = LIFU_ERR_CUDA;
  if (rc == LIFU_OK && p_min && s->Vsens) if (cudaMemcpyAsync(p_min, P.pmin, sizeof(float) * s->Vsens, cudaMemcpyDefault, st) != cudaSuccess) rc = LIFU_ERR_CUDA;
  cudaError_t se = cudaStreamSynchronize(st);
  for (int v = 0; v < 3; ++v) if (gexec[v]) cudaGraphExecDestroy(gexec[v]);
  if (rc == LIFU_ERR_CUDA && g_err.empty()) set_error("lifu_run: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc != LIFU_OK) return rc;
  if (se != cudaSuccess) { set_error("lifu_run: time loop failed: %s", cudaGetErrorString(se)); return LIFU_ERR_CUDA; }

  lifu_stats& r = s->last;
  memset(&r, 0, sizeof(r));
  r.voxels = s->V;
  for (int a = 0; a < 3; ++a) { r.n_exp[a] = s->N[a]; r.pml[a] = s->pml[a]; }
  r.steps = nt;
  r.source_steps = L;
  r.kernel_launches = cnt[0] * nk[0] + cnt[1] * nk[1] + cnt[2] * nk[2];
  r.fft_launches = cnt[0] * nf[0] + cnt[1] * nf[1] + cnt[2] * nf[2];
  r.steady_source_steps = (int32_t)cnt[2];
  float ms = 0;
  cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]); r.loop_ms = ms;
  cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); r.setup_ms = ms;
  r.bytes_per_voxel_step = (L * bytes_model(s, true) + (double)(nt - L) * bytes_model(s, false)) / (double)nt;
  r.homogeneous = s->homogeneous; r.absorbing = s->absorbing;
  if (stats) *stats = r;
  return LIFU_OK;
}

int lifu_set_two_z(lifu_sim* s, const double* two_z, int64_t n) {
  if (!s || !two_z || (n != 1 && n != s->Vin)) { set_error("lifu_set_two_z: need 1 value or one per inner-grid voxel"); return LIFU_ERR_INVALID; }
  if (s->sl.on) { set_error("lifu_set_two_z: not available on a slab handle"); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(s->device));
  if (n == 1) {
    LIFU_CUDA(cudaMemcpyAsync(&s->two_z_s, two_z, sizeof(double), cudaMemcpyDefault, s->stream));
    LIFU_CUDA(cudaStreamSynchronize(s->stream));
    s->two_z_mode = 1;
    return LIFU_OK;
  }
  if (!s->d_two_z) LIFU_CHECK(dev_alloc(s, (void**)&s->d_two_z, sizeof(double) * (size_t)s->Vin));
  LIFU_CUDA(cudaMemcpyAsync(s->d_two_z, two_z, sizeof(double) * (size_t)s->Vin, cudaMemcpyDefault, s->stream));
  LIFU_CUDA(cudaStreamSynchronize(s->stream));
  s->two_z_mode = 2;
  return LIFU_OK;
}

int lifu_get_packaged(lifu_sim* s, float* p_max, float* pnp, double* intensity) {
  if (!s || !pnp || !intensity) { set_error("lifu_get_packaged: null argument"); return LIFU_ERR_INVALID; }
  if (s->sl.on) { set_error("lifu_get_packaged: not available on a slab handle"); return LIFU_ERR_STATE; }
  if (s->last.steps <= 0) { set_error("lifu_get_packaged: call lifu_run first"); return LIFU_ERR_STATE; }
  if (s->two_z_mode == 0) { set_error("lifu_get_packaged: call lifu_set_two_z first"); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = s->stream;
  // results are staged in the scratch field r3 (idle between runs): [Vin] float, then [Vin] double
  float* d_pnp = s->P.r3;
  double* d_int = reinterpret_cast<double*>(s->P.r3 + round_up(s->Vin, 4));
  if (sizeof(float) * (size_t)round_up(s->Vin, 4) + sizeof(double) * (size_t)s->Vin > sizeof(float) * 3 * (size_t)s->RS) {
    set_error("lifu_get_packaged: staging area too small");
    return LIFU_ERR_NOMEM;
  }
  k_package<<<grid_blocks(s, s->Vin, 256), 256, 0, st>>>(s->P.pmin, s->two_z_mode == 2 ? s->d_two_z : nullptr, s->two_z_s,
                                                         d_pnp, d_int, s->Vin);
  LIFU_CUDA(cudaGetLastError());
  if (p_max) LIFU_CUDA(cudaMemcpyAsync(p_max, s->P.pmax, sizeof(float) * (size_t)s->Vin, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaMemcpyAsync(pnp, d_pnp, sizeof(float) * (size_t)s->Vin, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaMemcpyAsync(intensity, d_int, sizeof(double) * (size_t)s->Vin, cudaMemcpyDefault, st));
  LIFU_CUDA(cudaStreamSynchronize(st));
  return LIFU_OK;
}

int lifu_profile_stages(lifu_sim* s, int reps, int with_source, int max_stages, char* names, int name_stride,
                        double* ms, double* bytes_per_voxel, int* n_stages) {
  if (!s || reps <= 0 || !ms || !n_stages) { set_error("lifu_profile_stages: bad argument"); return LIFU_ERR_INVALID; }
  if (!(s->plans_ready || s->v2_ready || s->v3_ready || s->sl.plans) || !s->medium_set || !s->geometry_set || !s->drive_set) {
    set_error("lifu_profile_stages: call lifu_run once first");
    return LIFU_ERR_STATE;
  }
  LIFU_CUDA(cudaSetDevice(s->device));
  cudaStream_t user = s->stream, own = nullptr;
  if (!user) { LIFU_CUDA(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking)); s->stream = own; }
  cufftHandle hs[6] = {s->r2c1, s->r2c3, s->c2r1, s->c2r3, s->r2c2, s->c2r2};
  if (s->plans_ready) for (cufftHandle h : hs) cufftSetStream(h, s->stream);
  std::vector<double> acc;
  int rc = LIFU_OK;
  s->prof_names.clear(); s->prof_bytes.clear();
  for (int r = 0; r < reps && rc == LIFU_OK; ++r) {
    if (with_source) cudaMemsetAsync(s->P.step, 0, sizeof(int), s->stream);
    s->prof_on = true; s->prof_used = 0;
    // with_source: 0 no source, 1 generic source step, 2 step of the steady window (v2 pipeline after a run that had one)
    int kind = (with_source != 0 && s->n_src > 0) ? 1 : 0;
    if (with_source == 2 && s->last_used_v2 && s->Q.nws > 0) {
      kind = 2;
      cudaMemcpyAsync(s->P.step, &s->Q.t0s, sizeof(int), cudaMemcpyHostToDevice, s->stream);
    }
    rc = enqueue_step(s, kind, nullptr, nullptr);
    s->prof_on = false;
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) { set_error("lifu_profile_stages: step failed"); rc = LIFU_ERR_CUDA; }
    if (rc != LIFU_OK) break;
    if (acc.empty()) acc.assign(s->prof_used, 0.0);
    for (int i = 1; i < s->prof_used; ++i) {
      float t = 0; cudaEventElapsedTime(&t, s->prof_ev[i - 1], s->prof_ev[i]);
      acc[i] += t;
    }
  }
  if (s->plans_ready) for (cufftHandle h : hs) cufftSetStream(h, user);
  if (own) { cudaStreamDestroy(own); }
  s->stream = user;
  if (rc != LIFU_OK) return rc;
  int n = (int)acc.size() - 1;
  if (n > max_stages) n = max_stages;
  for (int i = 0; i < n; ++i) {
    ms[i] = acc[i + 1] / reps;
    if (bytes_per_voxel) bytes_per_voxel[i] = s->prof_bytes[i + 1];
    if (names && name_stride > 0) { strncpy(names + (size_t)i * name_stride, s->prof_names[i + 1], name_stride - 1); names[(size_t)i * name_stride + name_stride - 1] = 0; }
  }
  *n_stages = n;
  return LIFU_OK;
}

int lifu_get_field(lifu_sim* s, int which, float* out) {
  if (!s || !out || which < 0 || which > 6) { set_error("lifu_get_field: bad argument"); return LIFU_ERR_INVALID; }
  LIFU_CUDA(cudaSetDevice(s->device));
  const float* src = which == 0 ? s->P.p : (which <= 3 ? s->P.u + (which - 1) * s->RS : s->P.rho + (which - 4) * s->RS);
  LIFU_CUDA(cudaMemcpyAsync(out, src, sizeof(float) * s->Vloc, cudaMemcpyDefault, s->stream));
  LIFU_CUDA(cudaStreamSynchronize(s->stream));
  return LIFU_OK;
}

int lifu_get_info(lifu_sim* s, lifu_stats* stats) {
  if (!s || !stats) { set_error("lifu_get_info: null argument"); return LIFU_ERR_INVALID; }
  *stats = s->last;
  stats->voxels = s->V;
  for (int a = 0; a < 3; ++a) { stats->n_exp[a] = s->N[a]; stats->pml[a] = s->pml[a]; }
  stats->homogeneous = s->homogeneous; stats->absorbing = s->absorbing;
  return LIFU_OK;
}

}  // extern "C"
