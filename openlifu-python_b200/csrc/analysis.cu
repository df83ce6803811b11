// analysis.cu -- beam analysis of the simulated fields on the device.
//
// Device side of Solution.analyze (/root/reference/src/openlifu/plan/solution.py:181-270): per focus
//   k_focus_reduce    one pass over (pnp, ipa, max-over-foci ipa): focus-frame ellipsoid distance, main-lobe /
//                     side-lobe / z_ok selections, the eight masked maxima and the selection sizes
//   k_focus_centroid  value-weighted centroid sums over the -3 dB part of the main lobe
//   k_line_samples    trilinear samples along the three focus-frame axes (beam widths)
// and, while the fields are staged, k_max_into keeps the running maximum of the intensity over the foci (the
// reference masks the whole (focus, x, y, z) stack for I_SPTA, solution.py:245,270).
//
// Geometry is float64 with explicit round-to-nearest intrinsics (no FMA contraction) in the order numpy
// evaluates the reference expressions, so selections and samples are bit-identical to the host evaluation.
// HBM-bound streaming reads: 12 B/voxel (20 B with more than one focus) for k_focus_reduce; the other two
// kernels touch only the main lobe / 6*N points.
#include <algorithm>

#include "common.cuh"

namespace lifu {

struct AnaGeom {
  int n[3];               // Nx, Ny, Nz
  int md[3];              // sizes of the memory dims, slowest first
  int ax_of[3];           // logical axis (0 x, 1 y, 2 z) of memory dim 0, 1, 2
  unsigned V;             // voxels
  const double* t[3][3];  // t[i][a][k] = w[i][a] * axis_a[k]
  double w3[3], ar[3];
  const double* axis[3];  // coordinate vectors
  const unsigned char* z_ok;
};

__device__ __forceinline__ void ana_index(const AnaGeom& g, unsigned m, int& ix, int& iy, int& iz) {
  unsigned i2 = m % (unsigned)g.md[2];
  unsigned t = m / (unsigned)g.md[2];
  unsigned i1 = t % (unsigned)g.md[1];
  unsigned i0 = t / (unsigned)g.md[1];
  int li[3];
  li[0] = li[1] = li[2] = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (g.ax_of[0] == a) li[a] = (int)i0;
    else if (g.ax_of[1] == a) li[a] = (int)i1;
    else li[a] = (int)i2;
  }
  ix = li[0]; iy = li[1]; iz = li[2];
}

// sqrt(sum_i (((t_i0[x] + t_i1[y]) + t_i2[z]) + w_i3) / ar_i)^2), the order of FocusFrame.distance
// (host mirror of calc_dist_from_focus, solution_analysis.py:384-403)
__device__ __forceinline__ double ana_dist(const AnaGeom& g, int ix, int iy, int iz) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double c = __dadd_rn(__dadd_rn(__dadd_rn(__ldg(g.t[i][0] + ix), __ldg(g.t[i][1] + iy)), __ldg(g.t[i][2] + iz)), g.w3[i]);
    c = __ddiv_rn(c, g.ar[i]);
    acc = __dadd_rn(acc, __dmul_rn(c, c));
  }
  return __dsqrt_rn(acc);
}

constexpr int ANA_THREADS = 256;
constexpr int N_MAX = 8;   // main/side/global pnp, main/side/global ipa, main/global ipa over all foci
constexpr int N_CNT = 3;

__device__ __forceinline__ double warp_fmax(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fmax() returns the non-NaN operand, so starting from NaN gives "maximum ignoring NaN, NaN when nothing
// was selected" -- the semantics of DataArray.where(mask).max().
__global__ void __launch_bounds__(ANA_THREADS)
k_focus_reduce(AnaGeom g, const float* __restrict__ pnp, const double* __restrict__ ipa,
               const double* __restrict__ ipa_all, float scale, double r_main, double r_side,
               double* __restrict__ part_max, long long* __restrict__ part_cnt) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  double mx[N_MAX];
#pragma unroll
  for (int k = 0; k < N_MAX; ++k) mx[k] = qnan;
  long long cnt[N_CNT] = {0, 0, 0};
  for (unsigned m = blockIdx.x * ANA_THREADS + threadIdx.x; m < g.V; m += gridDim.x * ANA_THREADS) {
    int ix, iy, iz;
    ana_index(g, m, ix, iy, iz);
    double d = ana_dist(g, ix, iy, iz);
    bool zok = g.z_ok[iz] != 0;
    bool in_main = d < r_main;
    bool in_side = (d > r_side) && zok;
    double p = (double)__fmul_rn(pnp[m], scale);
    double I = ipa[m];
    double Ia = ipa_all[m];
    if (in_main) { mx[0] = fmax(mx[0], p); mx[3] = fmax(mx[3], I); mx[6] = fmax(mx[6], Ia); ++cnt[0]; }
    if (in_side) { mx[1] = fmax(mx[1], p); mx[4] = fmax(mx[4], I); ++cnt[1]; }
    if (zok)     { mx[2] = fmax(mx[2], p); mx[5] = fmax(mx[5], I); mx[7] = fmax(mx[7], Ia); ++cnt[2]; }
  }
  __shared__ double s_mx[ANA_THREADS / 32][N_MAX];
  __shared__ long long s_cnt[ANA_THREADS / 32][N_CNT];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N_MAX; ++k) { double v = warp_fmax(mx[k]); if (lane == 0) s_mx[wid][k] = v; }
#pragma unroll
  for (int k = 0; k < N_CNT; ++k) { long long v = warp_sum_ll(cnt[k]); if (lane == 0) s_cnt[wid][k] = v; }
  __syncthreads();
  if (threadIdx.x < N_MAX) {
    double v = qnan;
    for (int w = 0; w < ANA_THREADS / 32; ++w) v = fmax(v, s_mx[w][threadIdx.x]);
    part_max[blockIdx.x * N_MAX + threadIdx.x] = v;
  } else if (threadIdx.x < N_MAX + N_CNT) {
    int k = threadIdx.x - N_MAX;
    long long v = 0;
    for (int w = 0; w < ANA_THREADS / 32; ++w) v += s_cnt[w][k];
    part_cnt[blockIdx.x * N_CNT + k] = v;
  }
}

// w = pnp where (main lobe and pnp > cutoff) else 0 (float32, as the reference's where()); sums of w and w*axis.
__global__ void __launch_bounds__(ANA_THREADS)
k_focus_centroid(AnaGeom g, const float* __restrict__ pnp, float scale, double r_main, float cutoff,
                 double* __restrict__ part_sum, long long* __restrict__ part_cnt) {
  double s[4] = {0, 0, 0, 0};
  long long cnt = 0;
  for (unsigned m = blockIdx.x * ANA_THREADS + threadIdx.x; m < g.V; m += gridDim.x * ANA_THREADS) {
    int ix, iy, iz;
    ana_index(g, m, ix, iy, iz);
    if (!(ana_dist(g, ix, iy, iz) < r_main)) continue;
    float p = __fmul_rn(pnp[m], scale);
    if (!(p > cutoff)) continue;
    double w = (double)p;
    s[0] += w;
    s[1] += w * __ldg(g.axis[0] + ix);
    s[2] += w * __ldg(g.axis[1] + iy);
    s[3] += w * __ldg(g.axis[2] + iz);
    ++cnt;
  }
  __shared__ double s_s[ANA_THREADS / 32][4];
  __shared__ long long s_c[ANA_THREADS / 32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) { double v = warp_sum(s[k]); if (lane == 0) s_s[wid][k] = v; }
  { long long v = warp_sum_ll(cnt); if (lane == 0) s_c[wid] = v; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0;
    for (int w = 0; w < ANA_THREADS / 32; ++w) v += s_s[w][threadIdx.x];
    part_sum[blockIdx.x * 4 + threadIdx.x] = v;
  } else if (threadIdx.x == 4) {
    long long v = 0;
    for (int w = 0; w < ANA_THREADS / 32; ++w) v += s_c[w];
    part_cnt[blockIdx.x] = v;
  }
}

// lower node and fraction of q on a monotonically increasing axis (np.searchsorted(side='right') - 1, clipped)
__device__ __forceinline__ bool ana_bracket(const double* __restrict__ ax, int n, double q, int& i, double& t) {
  if (n == 1) { i = 0; t = 0.0; return q == ax[0]; }
  int lo = 0, hi = n;               // first index with ax[idx] > q
  while (lo < hi) { int mid = (lo + hi) >> 1; if (ax[mid] <= q) lo = mid + 1; else hi = mid; }
  i = min(max(lo - 1, 0), n - 2);
  t = __ddiv_rn(__dsub_rn(q, ax[i]), __dsub_rn(ax[i + 1], ax[i]));
  return q >= ax[0] && q <= ax[n - 1];
}

__device__ __forceinline__ double ana_lerp(double a, double b, double t) { return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), t)); }

// Trilinear samples, x then y then z (the order of a dimension-by-dimension linear interpolant over dims (x,y,z)).
__global__ void k_line_samples(AnaGeom g, const float* __restrict__ pnp, float scale, long long sx, long long sy,
                               long long sz, const double* __restrict__ pts, int n_pts, double* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pts) return;
  int ix, iy, iz;
  double tx, ty, tz;
  bool ok = ana_bracket(g.axis[0], g.n[0], pts[3 * k + 0], ix, tx);
  ok = ana_bracket(g.axis[1], g.n[1], pts[3 * k + 1], iy, ty) && ok;
  ok = ana_bracket(g.axis[2], g.n[2], pts[3 * k + 2], iz, tz) && ok;
  if (!ok) { out[k] = __longlong_as_double(0x7ff8000000000000LL); return; }
  int jx = min(ix + 1, g.n[0] - 1), jy = min(iy + 1, g.n[1] - 1), jz = min(iz + 1, g.n[2] - 1);
  auto v = [&](int a, int b, int c) { return (double)__fmul_rn(pnp[a * sx + b * sy + c * sz], scale); };
  double c00 = ana_lerp(v(ix, iy, iz), v(jx, iy, iz), tx);
  double c10 = ana_lerp(v(ix, jy, iz), v(jx, jy, iz), tx);
  double c01 = ana_lerp(v(ix, iy, jz), v(jx, iy, jz), tx);
  double c11 = ana_lerp(v(ix, jy, jz), v(jx, jy, jz), tx);
  out[k] = ana_lerp(ana_lerp(c00, c10, ty), ana_lerp(c01, c11, ty), tz);
}

__global__ void k_max_into(double* __restrict__ acc, const double* __restrict__ v, unsigned n) {
  for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) acc[m] = fmax(acc[m], v[m]);
}

}  // namespace lifu

using namespace lifu;

struct lifu_analysis {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n[3] = {0, 0, 0};
  int n_foci = 0;
  size_t V = 0;
  int sms = 148;
  int blocks = 0;
  std::vector<double> axis[3];
  double* d_axis = nullptr;        // x | y | z
  double* d_terms = nullptr;       // [3][x | y | z]
  unsigned char* d_zok = nullptr;
  float* d_pnp = nullptr;          // [n_foci][V]
  double* d_ipa = nullptr;         // [n_foci][V]
  double* d_all = nullptr;         // [V] maximum over the foci staged so far (aliases d_ipa when n_foci == 1)
  std::vector<char> staged;
  int n_staged = 0;
  int64_t stride[3] = {0, 0, 0};
  int md[3], ax_of[3];
  double* d_part = nullptr;        // per-block partial results
  long long* d_cnt = nullptr;
  double* d_pts = nullptr;
  double* d_line = nullptr;
  int line_cap = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

static void ana_free(lifu_analysis* a) {
  if (!a) return;
  cudaSetDevice(a->device);
  for (void* p : {(void*)a->d_axis, (void*)a->d_terms, (void*)a->d_zok, (void*)a->d_pnp, (void*)a->d_ipa,
                  (void*)(a->n_foci > 1 ? a->d_all : nullptr), (void*)a->d_part, (void*)a->d_cnt, (void*)a->d_pts,
                  (void*)a->d_line})
    if (p) cudaFree(p);
  if (a->e0) cudaEventDestroy(a->e0);
  if (a->e1) cudaEventDestroy(a->e1);
  delete a;
}

#define ANA_ALLOC(ptr, bytes)                                                                          \
  do {                                                                                                 \
    cudaError_t e__ = cudaMalloc((void**)&(ptr), (bytes));                                             \
    if (e__ != cudaSuccess) {                                                                          \
      set_error("lifu_analysis: cudaMalloc(%zu bytes) failed: %s", (size_t)(bytes), cudaGetErrorString(e__)); \
      cudaGetLastError();                                                                              \
      ana_free(a);                                                                                     \
      return LIFU_ERR_NOMEM;                                                                           \
    }                                                                                                  \
  } while (0)

extern "C" {

int lifu_analysis_create(int device, void* cuda_stream, const int32_t n[3], int32_t n_foci, const double* x,
                         const double* y, const double* z, const uint8_t* z_ok, lifu_analysis** out) {
  if (!n || !x || !y || !z || !out || n_foci < 1 || n[0] < 1 || n[1] < 1 || n[2] < 1) {
    set_error("lifu_analysis_create: bad argument");
    return LIFU_ERR_INVALID;
  }
  size_t V = (size_t)n[0] * n[1] * n[2];
  if (V >= (1ull << 32)) { set_error("lifu_analysis_create: more than 2^32 voxels per field"); return LIFU_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("lifu_analysis_create: no CUDA device visible; the device analysis has no CPU fallback");
    return LIFU_ERR_CUDA;
  }
  LIFU_CUDA(cudaSetDevice(device));
  auto* a = new lifu_analysis();
  a->device = device;
  a->stream = (cudaStream_t)cuda_stream;
  a->n_foci = n_foci;
  a->V = V;
  const double* src[3] = {x, y, z};
  for (int k = 0; k < 3; ++k) { a->n[k] = n[k]; a->axis[k].assign(src[k], src[k] + n[k]); }
  cudaDeviceGetAttribute(&a->sms, cudaDevAttrMultiProcessorCount, device);
  a->blocks = (int)std::min<size_t>((V + ANA_THREADS - 1) / ANA_THREADS, (size_t)a->sms * 8);
  a->staged.assign(n_foci, 0);
  size_t nsum = (size_t)n[0] + n[1] + n[2];
  ANA_ALLOC(a->d_axis, nsum * sizeof(double));
  ANA_ALLOC(a->d_terms, 3 * nsum * sizeof(double));
  ANA_ALLOC(a->d_zok, (size_t)n[2]);
  ANA_ALLOC(a->d_pnp, (size_t)n_foci * V * sizeof(float));
  ANA_ALLOC(a->d_ipa, (size_t)n_foci * V * sizeof(double));
  if (n_foci > 1) ANA_ALLOC(a->d_all, V * sizeof(double));
  else a->d_all = a->d_ipa;
  ANA_ALLOC(a->d_part, (size_t)a->blocks * N_MAX * sizeof(double));
  ANA_ALLOC(a->d_cnt, (size_t)a->blocks * N_CNT * sizeof(long long));
  std::vector<double> flat;
  for (int k = 0; k < 3; ++k) flat.insert(flat.end(), a->axis[k].begin(), a->axis[k].end());
  std::vector<unsigned char> zk(n[2], 1);
  if (z_ok) zk.assign(z_ok, z_ok + n[2]);
  cudaError_t e = cudaMemcpyAsync(a->d_axis, flat.data(), nsum * sizeof(double), cudaMemcpyHostToDevice, a->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(a->d_zok, zk.data(), (size_t)n[2], cudaMemcpyHostToDevice, a->stream);
  if (e == cudaSuccess) e = cudaEventCreate(&a->e0);
  if (e == cudaSuccess) e = cudaEventCreate(&a->e1);
  if (e == cudaSuccess) e = cudaStreamSynchronize(a->stream);
  if (e != cudaSuccess) { set_error("lifu_analysis_create: %s", cudaGetErrorString(e)); ana_free(a); return LIFU_ERR_CUDA; }
  *out = a;
  return LIFU_OK;
}

int lifu_analysis_set_focus(lifu_analysis* a, int32_t focus, const float* pnp, const double* ipa, const int64_t stride[3]) {
  if (!a || !pnp || !ipa || !stride || focus < 0 || focus >= a->n_foci) { set_error("lifu_analysis_set_focus: bad argument"); return LIFU_ERR_INVALID; }
  // the strides must be a dense permutation of (Nx, Ny, Nz)
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int p, int q) { return stride[p] > stride[q] || (stride[p] == stride[q] && a->n[p] > a->n[q]); });
  int64_t expect = 1;
  for (int k = 2; k >= 0; --k) {
    if (a->n[order[k]] > 1 && stride[order[k]] != expect) {
      set_error("lifu_analysis_set_focus: strides (%lld, %lld, %lld) do not describe a dense (%d, %d, %d) array",
                (long long)stride[0], (long long)stride[1], (long long)stride[2], a->n[0], a->n[1], a->n[2]);
      return LIFU_ERR_INVALID;
    }
    expect *= a->n[order[k]];
  }
  int64_t dense[3];
  expect = 1;
  for (int k = 2; k >= 0; --k) { dense[order[k]] = expect; expect *= a->n[order[k]]; }
  if (a->n_staged > 0 && (dense[0] != a->stride[0] || dense[1] != a->stride[1] || dense[2] != a->stride[2])) {
    set_error("lifu_analysis_set_focus: every focus must use the same memory layout");
    return LIFU_ERR_INVALID;
  }
  for (int k = 0; k < 3; ++k) { a->stride[k] = dense[k]; a->md[k] = a->n[order[k]]; a->ax_of[k] = order[k]; }
  LIFU_CUDA(cudaSetDevice(a->device));
  float* dp = a->d_pnp + (size_t)focus * a->V;
  double* di = a->d_ipa + (size_t)focus * a->V;
  LIFU_CUDA(cudaMemcpyAsync(dp, pnp, a->V * sizeof(float), cudaMemcpyDefault, a->stream));
  LIFU_CUDA(cudaMemcpyAsync(di, ipa, a->V * sizeof(double), cudaMemcpyDefault, a->stream));
  if (a->n_foci > 1) {
    if (a->n_staged == 0) LIFU_CUDA(cudaMemcpyAsync(a->d_all, di, a->V * sizeof(double), cudaMemcpyDeviceToDevice, a->stream));
    else if (!a->staged[focus]) {
      k_max_into<<<a->blocks, ANA_THREADS, 0, a->stream>>>(a->d_all, di, (unsigned)a->V);
      LIFU_CUDA(cudaGetLastError());
    } else {
      // a focus is being replaced: rebuild the running maximum from every staged field
      a->staged[focus] = 1;
      bool first = true;
      for (int f = 0; f < a->n_foci; ++f) {
        if (!a->staged[f]) continue;
        const double* src = a->d_ipa + (size_t)f * a->V;
        if (first) { LIFU_CUDA(cudaMemcpyAsync(a->d_all, src, a->V * sizeof(double), cudaMemcpyDeviceToDevice, a->stream)); first = false; }
        else k_max_into<<<a->blocks, ANA_THREADS, 0, a->stream>>>(a->d_all, src, (unsigned)a->V);
      }
      LIFU_CUDA(cudaGetLastError());
    }
  }
  if (!a->staged[focus]) { a->staged[focus] = 1; ++a->n_staged; }
  LIFU_CUDA(cudaStreamSynchronize(a->stream));   // the caller's buffers are free again when this returns
  return LIFU_OK;
}

int lifu_analysis_run_focus(lifu_analysis* a, int32_t focus, const lifu_focus_query* q, const double* line_pts,
                            lifu_focus_metrics* out, double* line_vals) {
  if (!a || !q || !out || focus < 0 || focus >= a->n_foci) { set_error("lifu_analysis_run_focus: bad argument"); return LIFU_ERR_INVALID; }
  if (a->n_staged != a->n_foci) {
    set_error("lifu_analysis_run_focus: %d of %d foci staged (the I_SPTA maxima run over every focus' field)", a->n_staged, a->n_foci);
    return LIFU_ERR_STATE;
  }
  int n_pts = 0;
  for (int k = 0; k < 3; ++k) {
    if (q->n_line[k] < 0) { set_error("lifu_analysis_run_focus: negative n_line"); return LIFU_ERR_INVALID; }
    n_pts += q->n_line[k];
  }
  if (n_pts > 0 && (!line_pts || !line_vals)) { set_error("lifu_analysis_run_focus: line buffers missing"); return LIFU_ERR_INVALID; }
  for (int i = 0; i < 3; ++i)
    if (!(q->aspect[i] != 0.0)) { set_error("lifu_analysis_run_focus: zero aspect ratio"); return LIFU_ERR_INVALID; }
  LIFU_CUDA(cudaSetDevice(a->device));
  cudaStream_t st = a->stream;

  // per-axis products w[i][a] * axis_a (plain float64 multiplies, as numpy forms them)
  size_t nsum = (size_t)a->n[0] + a->n[1] + a->n[2];
  std::vector<double> terms(3 * nsum);
  AnaGeom g;
  g.V = (unsigned)a->V;
  for (int k = 0; k < 3; ++k) { g.n[k] = a->n[k]; g.md[k] = a->md[k]; g.ax_of[k] = a->ax_of[k]; }
  size_t off_axis[3] = {0, (size_t)a->n[0], (size_t)a->n[0] + a->n[1]};
  for (int i = 0; i < 3; ++i) {
    for (int ax = 0; ax < 3; ++ax) {
      double* dst = terms.data() + i * nsum + off_axis[ax];
      const volatile double wv = q->w[i][ax];
      for (int k = 0; k < a->n[ax]; ++k) dst[k] = wv * a->axis[ax][k];
      g.t[i][ax] = a->d_terms + i * nsum + off_axis[ax];
    }
    g.w3[i] = q->w[i][3];
    g.ar[i] = q->aspect[i];
  }
  for (int ax = 0; ax < 3; ++ax) g.axis[ax] = a->d_axis + off_axis[ax];
  g.z_ok = a->d_zok;
  LIFU_CUDA(cudaMemcpyAsync(a->d_terms, terms.data(), terms.size() * sizeof(double), cudaMemcpyHostToDevice, st));

  const float* dp = a->d_pnp + (size_t)focus * a->V;
  const double* di = a->d_ipa + (size_t)focus * a->V;
  LIFU_CUDA(cudaEventRecord(a->e0, st));
  k_focus_reduce<<<a->blocks, ANA_THREADS, 0, st>>>(g, dp, di, a->d_all, q->pnp_scale, q->mainlobe_radius, q->sidelobe_radius,
                                                    a->d_part, a->d_cnt);
  LIFU_CUDA(cudaGetLastError());
  std::vector<double> hmax((size_t)a->blocks * N_MAX);
  std::vector<long long> hcnt((size_t)a->blocks * N_CNT);
  LIFU_CUDA(cudaMemcpyAsync(hmax.data(), a->d_part, hmax.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaMemcpyAsync(hcnt.data(), a->d_cnt, hcnt.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaStreamSynchronize(st));
  double mx[N_MAX];
  long long cnt[N_CNT] = {0, 0, 0};
  for (int k = 0; k < N_MAX; ++k) mx[k] = std::nan("");
  for (int b = 0; b < a->blocks; ++b) {
    for (int k = 0; k < N_MAX; ++k) mx[k] = std::fmax(mx[k], hmax[(size_t)b * N_MAX + k]);
    for (int k = 0; k < N_CNT; ++k) cnt[k] += hcnt[(size_t)b * N_CNT + k];
  }
  out->main_pnp = mx[0]; out->side_pnp = mx[1]; out->global_pnp = mx[2];
  out->main_ipa = mx[3]; out->side_ipa = mx[4]; out->global_ipa = mx[5];
  out->main_ipa_all = mx[6]; out->global_ipa_all = mx[7];
  out->n_main = cnt[0]; out->n_side = cnt[1]; out->n_global = cnt[2];

  // centroid of the -3 dB part of the main lobe; the cutoff is compared in float32 like the float32 field is
  const volatile double cut64 = out->main_pnp * q->centroid_factor;
  float cutoff = (float)cut64;
  k_focus_centroid<<<a->blocks, ANA_THREADS, 0, st>>>(g, dp, q->pnp_scale, q->mainlobe_radius, cutoff, a->d_part, a->d_cnt);
  LIFU_CUDA(cudaGetLastError());
  if (n_pts > 0) {
    if (n_pts > a->line_cap) {
      if (a->d_pts) cudaFree(a->d_pts);
      if (a->d_line) cudaFree(a->d_line);
      a->d_pts = a->d_line = nullptr;
      a->line_cap = 0;
      LIFU_CUDA(cudaMalloc((void**)&a->d_pts, (size_t)n_pts * 3 * sizeof(double)));
      LIFU_CUDA(cudaMalloc((void**)&a->d_line, (size_t)n_pts * sizeof(double)));
      a->line_cap = n_pts;
    }
    LIFU_CUDA(cudaMemcpyAsync(a->d_pts, line_pts, (size_t)n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    k_line_samples<<<(n_pts + 127) / 128, 128, 0, st>>>(g, dp, q->pnp_scale, a->stride[0], a->stride[1], a->stride[2], a->d_pts,
                                                        n_pts, a->d_line);
    LIFU_CUDA(cudaGetLastError());
  }
  LIFU_CUDA(cudaEventRecord(a->e1, st));
  std::vector<double> hs((size_t)a->blocks * 4);
  std::vector<long long> hc((size_t)a->blocks);
  LIFU_CUDA(cudaMemcpyAsync(hs.data(), a->d_part, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaMemcpyAsync(hc.data(), a->d_cnt, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (n_pts > 0) LIFU_CUDA(cudaMemcpyAsync(line_vals, a->d_line, (size_t)n_pts * sizeof(double), cudaMemcpyDeviceToHost, st));
  LIFU_CUDA(cudaStreamSynchronize(st));
  double s[4] = {0, 0, 0, 0};
  long long nc = 0;
  for (int b = 0; b < a->blocks; ++b) {
    for (int k = 0; k < 4; ++k) s[k] += hs[(size_t)b * 4 + k];
    nc += hc[b];
  }
  out->cen_w = s[0]; out->cen_wx = s[1]; out->cen_wy = s[2]; out->cen_wz = s[3];
  out->n_centroid = nc;
  float ms = 0.f;
  LIFU_CUDA(cudaEventElapsedTime(&ms, a->e0, a->e1));
  out->kernel_ms = ms;
  return LIFU_OK;
}

int lifu_analysis_destroy(lifu_analysis* a) {
  ana_free(a);
  return LIFU_OK;
}

}  // extern "C"
