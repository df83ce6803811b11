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

// Selection of one voxel: main lobe (dist < r_main) and "outside the side-lobe radius" (dist > r_side), decided
// exactly as the reference does, i.e. on sqrt_rn(sum_i ((c_i / ar_i)^2)).  The divisions and the square root are only
// evaluated for voxels whose squared distance (formed with reciprocals) lies within 1e-12 relative of a threshold:
// the two evaluations differ by a few ulp (< 2e-15 relative), so outside that band the comparison cannot flip.
struct AnaBand {
  double inv_ar[3];
  double lo_m, hi_m, lo_s, hi_s;   // r^2 (1 -+ 1e-12) for the main-lobe and the side-lobe radius
  int exact_only;                  // degenerate radii: always take the exact path
};

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

// The reference's distance, operation by operation (rare path).
__device__ __forceinline__ double ana_exact_dist(double c0, double c1, double c2, double a0, double a1, double a2) {
  const double q0 = __ddiv_rn(c0, a0), q1 = __ddiv_rn(c1, a1), q2 = __ddiv_rn(c2, a2);
  double e = __dmul_rn(q0, q0);                       // 0 + q0^2
  e = __dadd_rn(e, __dmul_rn(q1, q1));
  e = __dadd_rn(e, __dmul_rn(q2, q2));
  return __dsqrt_rn(e);
}

// Row-wise traversal shared by the two field passes: a warp owns rows of the fastest memory axis (lanes stride over
// it, so loads are coalesced and no per-voxel division is needed); the focus-frame terms of the two slower axes are
// row constants.  FAST = logical axis (0 x, 1 y, 2 z) that is fastest in memory.
template <int FAST>
struct AnaRow {
  double ra[3], rb[3];      // row-constant terms of the two non-fast axes, in logical axis order
  const double* tf[3];      // per-voxel term tables of the fast axis
  static constexpr int A = FAST == 0 ? 1 : 0, B = FAST == 2 ? 1 : 2;   // the two slow logical axes, ascending
  int ia, ib;               // the row's indices along A and B
  __device__ __forceinline__ void begin(const AnaGeom& g, unsigned row) {
    const unsigned i1 = row % (unsigned)g.md[1], i0 = row / (unsigned)g.md[1];
    const bool swapped = g.ax_of[0] != A;                               // memory dim 0 holds logical axis B
    ia = (int)(swapped ? i1 : i0);
    ib = (int)(swapped ? i0 : i1);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      ra[i] = __ldg(g.t[i][A] + ia);
      rb[i] = __ldg(g.t[i][B] + ib);
      tf[i] = g.t[i][FAST];
    }
  }
  __device__ __forceinline__ int index_of(int axis, int k) const { return axis == FAST ? k : (axis == A ? ia : ib); }
  // ((t_x + t_y) + t_z) + w3 in the reference's order, whichever axis is the per-voxel one
  __device__ __forceinline__ double coord(const AnaGeom& g, int i, int k) const {
    const double v = __ldg(tf[i] + k);
    double c;
    if (FAST == 0) c = __dadd_rn(__dadd_rn(v, ra[i]), rb[i]);
    else if (FAST == 1) c = __dadd_rn(__dadd_rn(ra[i], v), rb[i]);
    else c = __dadd_rn(__dadd_rn(ra[i], rb[i]), v);
    return __dadd_rn(c, g.w3[i]);
  }
  __device__ __forceinline__ void select(const AnaGeom& g, const AnaBand& b, double r_main, double r_side, int k,
                                         bool& in_main, bool& out_side) const {
    double c[3], acc = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c[i] = coord(g, i, k);
      const double s = c[i] * b.inv_ar[i];
      acc = fma(s, s, acc);
    }
    const bool near_m = acc > b.lo_m && acc < b.hi_m, near_s = acc > b.lo_s && acc < b.hi_s;
    if (b.exact_only || near_m || near_s) {
      const double d = ana_exact_dist(c[0], c[1], c[2], g.ar[0], g.ar[1], g.ar[2]);
      in_main = d < r_main;
      out_side = d > r_side;
    } else {
      in_main = acc <= b.lo_m;
      out_side = acc >= b.hi_s;
    }
  }
};

// fmax() / fmaxf() return the non-NaN operand, so starting from NaN gives "maximum ignoring NaN, NaN when nothing
// was selected" -- the semantics of DataArray.where(mask).max().  The pressure maxima are taken in float32 (the
// field's own type; the conversion to float64 is monotone and exact).
template <int FAST>
__global__ void __launch_bounds__(ANA_THREADS, 3)
k_focus_reduce(AnaGeom g, AnaBand band, const float* __restrict__ pnp, const double* __restrict__ ipa,
               const double* __restrict__ ipa_all, float scale, double r_main, double r_side,
               double* __restrict__ part_max, long long* __restrict__ part_cnt) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const float qnanf = __int_as_float(0x7fc00000);
  float mp[3] = {qnanf, qnanf, qnanf};                 // main / side / global pressure
  double mi[5] = {qnan, qnan, qnan, qnan, qnan};       // main / side / global intensity, main / global over all foci
  int cnt[N_CNT] = {0, 0, 0};
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned rows = (unsigned)g.md[0] * (unsigned)g.md[1], nfast = (unsigned)g.md[2];
  const unsigned warps = gridDim.x * (ANA_THREADS / 32);
  AnaRow<FAST> R;
  for (unsigned row = blockIdx.x * (ANA_THREADS / 32) + wid; row < rows; row += warps) {
    R.begin(g, row);
    const size_t base = (size_t)row * nfast;
    const bool row_zok = FAST == 2 ? true : g.z_ok[R.index_of(2, 0)] != 0;
    // register double buffer over the row: the streamed operands of element k + 32 are in flight while element k
    // is classified
    float praw_n = 0.f;
    double I_n = 0.0, Ia_n = 0.0;
    if ((unsigned)lane < nfast) { praw_n = __ldcs(pnp + base + lane); I_n = __ldcs(ipa + base + lane); Ia_n = __ldcs(ipa_all + base + lane); }
    for (unsigned k = lane; k < nfast; k += 32) {
      const float praw = praw_n;
      const double I = I_n, Ia = Ia_n;
      if (k + 32 < nfast) { praw_n = __ldcs(pnp + base + k + 32); I_n = __ldcs(ipa + base + k + 32); Ia_n = __ldcs(ipa_all + base + k + 32); }
      bool in_main, out_side;
      R.select(g, band, r_main, r_side, (int)k, in_main, out_side);
      const bool zok = FAST == 2 ? g.z_ok[k] != 0 : row_zok;
      const bool in_side = out_side && zok;
      const float p = __fmul_rn(praw, scale);
      if (in_main) { mp[0] = fmaxf(mp[0], p); mi[0] = fmax(mi[0], I); mi[3] = fmax(mi[3], Ia); ++cnt[0]; }
      if (in_side) { mp[1] = fmaxf(mp[1], p); mi[1] = fmax(mi[1], I); ++cnt[1]; }
      if (zok)     { mp[2] = fmaxf(mp[2], p); mi[2] = fmax(mi[2], I); mi[4] = fmax(mi[4], Ia); ++cnt[2]; }
    }
  }
  double mx[N_MAX] = {(double)mp[0], (double)mp[1], (double)mp[2], mi[0], mi[1], mi[2], mi[3], mi[4]};
  __shared__ double s_mx[ANA_THREADS / 32][N_MAX];
  __shared__ long long s_cnt[ANA_THREADS / 32][N_CNT];
#pragma unroll
  for (int k = 0; k < N_MAX; ++k) { double v = warp_fmax(mx[k]); if (lane == 0) s_mx[wid][k] = v; }
#pragma unroll
  for (int k = 0; k < N_CNT; ++k) { long long v = warp_sum_ll((long long)cnt[k]); if (lane == 0) s_cnt[wid][k] = v; }
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
template <int FAST>
__global__ void __launch_bounds__(ANA_THREADS)
k_focus_centroid(AnaGeom g, AnaBand band, const float* __restrict__ pnp, float scale, double r_main, float cutoff,
                 double* __restrict__ part_sum, long long* __restrict__ part_cnt) {
  double s[4] = {0, 0, 0, 0};
  long long cnt = 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned rows = (unsigned)g.md[0] * (unsigned)g.md[1], nfast = (unsigned)g.md[2];
  const unsigned warps = gridDim.x * (ANA_THREADS / 32);
  AnaRow<FAST> R;
  for (unsigned row = blockIdx.x * (ANA_THREADS / 32) + wid; row < rows; row += warps) {
    R.begin(g, row);
    const size_t base = (size_t)row * nfast;
    for (unsigned k = lane; k < nfast; k += 32) {
      bool in_main, out_side;
      R.select(g, band, r_main, r_main, (int)k, in_main, out_side);
      if (!in_main) continue;
      const float p = __fmul_rn(pnp[base + k], scale);
      if (!(p > cutoff)) continue;
      const double w = (double)p;
      s[0] += w;
      s[1] += w * __ldg(g.axis[0] + R.index_of(0, (int)k));
      s[2] += w * __ldg(g.axis[1] + R.index_of(1, (int)k));
      s[3] += w * __ldg(g.axis[2] + R.index_of(2, (int)k));
      ++cnt;
    }
  }
  __shared__ double s_s[ANA_THREADS / 32][4];
  __shared__ long long s_c[ANA_THREADS / 32];
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
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
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
  if (a->e2) cudaEventDestroy(a->e2);
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
  if (e == cudaSuccess) e = cudaEventCreate(&a->e2);
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
  NvtxRange nvtx_r("lifu_analysis_run_focus");
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

  AnaBand band, band_c;
  {
    const double rm = q->mainlobe_radius, rs = q->sidelobe_radius, eps = 1e-12;
    for (int i = 0; i < 3; ++i) band.inv_ar[i] = 1.0 / q->aspect[i];
    band.lo_m = rm * rm * (1.0 - eps); band.hi_m = rm * rm * (1.0 + eps);
    band.lo_s = rs * rs * (1.0 - eps); band.hi_s = rs * rs * (1.0 + eps);
    band.exact_only = !(rm > 0.0 && rs > 0.0 && std::isfinite(rm) && std::isfinite(rs)) ? 1 : 0;
    band_c = band;                                   // centroid pass: only the main-lobe radius matters
    band_c.lo_s = band.lo_m; band_c.hi_s = band.hi_m;
  }
  const float* dp = a->d_pnp + (size_t)focus * a->V;
  const double* di = a->d_ipa + (size_t)focus * a->V;
  LIFU_CUDA(cudaEventRecord(a->e0, st));
  const int fast = a->ax_of[2];
#define ANA_FAST(KERNEL, ...)                                                          \
  do {                                                                                 \
    if (fast == 0) KERNEL<0><<<a->blocks, ANA_THREADS, 0, st>>>(__VA_ARGS__);          \
    else if (fast == 1) KERNEL<1><<<a->blocks, ANA_THREADS, 0, st>>>(__VA_ARGS__);     \
    else KERNEL<2><<<a->blocks, ANA_THREADS, 0, st>>>(__VA_ARGS__);                    \
  } while (0)
  ANA_FAST(k_focus_reduce, g, band, dp, di, a->d_all, q->pnp_scale, q->mainlobe_radius, q->sidelobe_radius, a->d_part, a->d_cnt);
  LIFU_CUDA(cudaGetLastError());
  LIFU_CUDA(cudaEventRecord(a->e2, st));
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
  ANA_FAST(k_focus_centroid, g, band_c, dp, q->pnp_scale, q->mainlobe_radius, cutoff, a->d_part, a->d_cnt);
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
  LIFU_CUDA(cudaEventElapsedTime(&ms, a->e0, a->e2));
  out->reduce_ms = ms;
  return LIFU_OK;
}

int lifu_analysis_destroy(lifu_analysis* a) {
  ana_free(a);
  return LIFU_OK;
}

}  // extern "C"
