// bli.cu -- K0: off-grid rectangular transducer elements -> source mask + sparse weights.
//
// Replaces kWaveArray.get_array_binary_mask / get_distributed_source_signal (weights part) as
// driven by /root/reference/src/openlifu/sim/kwave_if.py:29-47 (get_karray), :75-77 (get_source),
// which the reference evaluates in pure Python twice per focus.  Method:
// band-limited interpolation with a truncated sinc (Wise et al., JASA 2019): every integration
// point of an element spreads onto the "star" |i*j*k| <= h, h = ceil(1/(pi*tol)), around its
// nearest grid node; an axis on which the point sits on a node (within dx*1e-3) collapses.
//
// Integer decisions (nearest node, on-grid flags, point counts) are taken on the host in
// float64 so that the mask is bit-exact; the O(points x star) arithmetic runs on the GPU as a
// *gather*: one thread per voxel of the element's bounding box walks the element's points in
// their reference order, which makes the single-precision accumulation deterministic.
#include <cub/cub.cuh>

#include "sim.cuh"

namespace lifu {

struct BliElem {
  int pt_off, n_pts;
  int lo[3], nb[3];
  long long box_off;
  double scale;
};

struct BliPoint {
  int c[3];      // nearest grid node per axis
  int ongrid;    // bit a set: collapse axis a
};

// tab[(p*3+a)*TW + (o+h)] = sinc(pi/d_a * (x_vec_a[c_a+o] - point_a)); TW = 2h+1
__global__ void k_bli_tables(const double* __restrict__ pts, const BliPoint* __restrict__ bp,
                             long long n_pts, int h, int n0, int n1, int n2, double d0, double d1,
                             double d2, double* __restrict__ tab) {
  const int TW = 2 * h + 1;
  long long total = n_pts * 3 * TW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int o = (int)(i % TW) - h;
    long long t = i / TW;
    int a = (int)(t % 3);
    long long p = t / 3;
    int N = a == 0 ? n0 : (a == 1 ? n1 : n2);
    double d = a == 0 ? d0 : (a == 1 ? d1 : d2);
    int g = bp[p].c[a] + o;
    double v = 0.0;
    if (g >= 0 && g < N) {
      // kWaveGrid position of node g: (N*d) * ((g - floor(N/2)) / N)
      double xg = ((double)N * d) * ((double)(g - N / 2) / (double)N);
      double x = (3.14159265358979323846 / d) * (xg - pts[p * 3 + a]);
      v = x != 0.0 ? sin(x) / x : 1.0;
    }
    tab[i] = v;
  }
}

__global__ void __launch_bounds__(128) k_bli_gather(const BliElem* __restrict__ el, int n_el,
                                                    const BliPoint* __restrict__ bp,
                                                    const double* __restrict__ tab, int h,
                                                    float* __restrict__ wbox,
                                                    unsigned char* __restrict__ tbox) {
  const int e = blockIdx.y;
  const BliElem E = el[e];
  const int TW = 2 * h + 1;
  const long long nvox = (long long)E.nb[0] * E.nb[1] * E.nb[2];
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox;
       v += (long long)gridDim.x * blockDim.x) {
    int bx = (int)(v % E.nb[0]);
    long long t = v / E.nb[0];
    int by = (int)(t % E.nb[1]);
    int bz = (int)(t / E.nb[1]);
    int gx = bx + E.lo[0], gy = by + E.lo[1], gz = bz + E.lo[2];
    float acc = 0.f;
    unsigned char touched = 0;
    for (int q = 0; q < E.n_pts; ++q) {
      const long long p = E.pt_off + q;
      const BliPoint B = bp[p];
      int ox = gx - B.c[0], oy = gy - B.c[1], oz = gz - B.c[2];
      if (abs(ox) > h || abs(oy) > h || abs(oz) > h) continue;
      if (((B.ongrid & 1) && ox != 0) || ((B.ongrid & 2) && oy != 0) || ((B.ongrid & 4) && oz != 0)) continue;
      if (abs(ox * oy * oz) > h) continue;
      const double* tp = tab + p * 3 * TW;
      double w = tp[ox + h] * tp[TW + oy + h] * tp[2 * TW + oz + h];
      // reference accumulates float64 terms into a float32 grid: round after every add
      acc = (float)((double)acc + E.scale * w);
      touched = 1;
    }
    wbox[E.box_off + v] = acc;
    tbox[E.box_off + v] = touched;
  }
}

__global__ void k_bli_mark(const BliElem* __restrict__ el, const unsigned char* __restrict__ tbox,
                           int nx, int ny, unsigned char* __restrict__ mask) {
  const BliElem E = el[blockIdx.y];
  const long long nvox = (long long)E.nb[0] * E.nb[1] * E.nb[2];
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox;
       v += (long long)gridDim.x * blockDim.x) {
    if (!tbox[E.box_off + v]) continue;
    int bx = (int)(v % E.nb[0]);
    long long t = v / E.nb[0];
    int by = (int)(t % E.nb[1]);
    int bz = (int)(t / E.nb[1]);
    long long lin = ((long long)(bz + E.lo[2]) * ny + (by + E.lo[1])) * nx + (bx + E.lo[0]);
    mask[lin] = 1;
  }
}

// FILL = false: cnt[i] = number of elements with a non-zero weight at source point i.
// FILL = true : write (element, weight) pairs at row_ptr[i]...
template <bool FILL>
__global__ void k_bli_rows(const long long* __restrict__ idx, long long n_src, const BliElem* __restrict__ el,
                           int n_el, const float* __restrict__ wbox, const unsigned char* __restrict__ tbox,
                           int nx, int ny, int* __restrict__ cnt, const int* __restrict__ row_ptr,
                           int* __restrict__ col, float* __restrict__ w) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_src;
       i += (long long)gridDim.x * blockDim.x) {
    long long l = idx[i];
    int gx = (int)(l % nx);
    long long t = l / nx;
    int gy = (int)(t % ny);
    int gz = (int)(t / ny);
    int c = 0;
    int base = FILL ? row_ptr[i] : 0;
    for (int e = 0; e < n_el; ++e) {
      const BliElem& E = el[e];
      int bx = gx - E.lo[0], by = gy - E.lo[1], bz = gz - E.lo[2];
      if ((unsigned)bx >= (unsigned)E.nb[0] || (unsigned)by >= (unsigned)E.nb[1] || (unsigned)bz >= (unsigned)E.nb[2]) continue;
      long long v = E.box_off + ((long long)bz * E.nb[1] + by) * E.nb[0] + bx;
      if (!tbox[v]) continue;
      float wv = wbox[v];
      if (wv == 0.f) continue;   // get_distributed_source_signal keys on weights != 0
      if (FILL) { col[base + c] = e; w[base + c] = wv; }
      ++c;
    }
    if (!FILL) cnt[i] = c;
  }
}

static inline double grid_pos(int N, double d, int i) {
  return ((double)N * d) * ((double)(i - N / 2) / (double)N);
}

// numpy argmin(|x_vec - p|): first index attaining the minimum.
static int closest_node(int N, double d, double p) {
  double g = (p - grid_pos(N, d, 0)) / d;
  long long c0 = (long long)std::floor(g + 0.5);
  if (c0 < 0) c0 = 0;
  if (c0 > N - 1) c0 = N - 1;
  int lo = (int)std::max<long long>(c0 - 2, 0), hi = (int)std::min<long long>(c0 + 2, N - 1);
  int best = lo;
  double bd = std::fabs(grid_pos(N, d, lo) - p);
  for (int i = lo + 1; i <= hi; ++i) {
    double di = std::fabs(grid_pos(N, d, i) - p);
    if (di < bd) { bd = di; best = i; }
  }
  return best;
}

static inline double round_half_even(double v) { return std::nearbyint(v); }  // np.round

int bli_build(lifu_sim* s, int n_el, const double* pos, const double* size, const double* ang,
              double tol, int ups) {
  if (n_el <= 0 || tol <= 0.0 || ups <= 0) {
    set_error("lifu_set_elements: need n_el > 0, bli_tolerance > 0 (got %g), upsampling_rate > 0", tol);
    return LIFU_ERR_INVALID;
  }
  const int h = (int)std::ceil(1.0 / (M_PI * tol));
  const int* n = s->n;
  const double* d = s->grid.d;
  const double thr = d[0] * 1e-3;
  std::vector<BliElem> elems(n_el);
  std::vector<double> pts;
  std::vector<BliPoint> bps;
  long long box_total = 0;
  const double deg = M_PI / 180.0;
  for (int e = 0; e < n_el; ++e) {
    const double Lx = size[2 * e], Ly = size[2 * e + 1];
    if (!(Lx > 0.0) || !(Ly > 0.0)) { set_error("element %d has non-positive size", e); return LIFU_ERR_INVALID; }
    const double m_grid = (Lx * Ly) / (d[0] * d[0]);
    const long long m_int = (long long)std::ceil(m_grid * (double)ups);
    int npx = (int)round_half_even(std::sqrt((double)m_int * Lx / Ly));
    if (npx < 1) npx = 1;
    int npy = (int)round_half_even((double)m_int / (double)npx);
    if (npy < 1) npy = 1;
    // R = Rz(roll) Ry(az) Rx(el); theta = (el, az, roll) degrees
    const double tx = ang[3 * e] * deg, ty = ang[3 * e + 1] * deg, tz = ang[3 * e + 2] * deg;
    const double Rx[9] = {1, 0, 0, 0, cos(tx), -sin(tx), 0, sin(tx), cos(tx)};
    const double Ry[9] = {cos(ty), 0, sin(ty), 0, 1, 0, -sin(ty), 0, cos(ty)};
    const double Rz[9] = {cos(tz), -sin(tz), 0, sin(tz), cos(tz), 0, 0, 0, 1};
    double T[9], R[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) { T[3 * r + c] = 0; for (int k = 0; k < 3; ++k) T[3 * r + c] += Ry[3 * r + k] * Rx[3 * k + c]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) { R[3 * r + c] = 0; for (int k = 0; k < 3; ++k) R[3 * r + c] += Rz[3 * r + k] * T[3 * k + c]; }
    BliElem& E = elems[e];
    E.pt_off = (int)bps.size();
    E.n_pts = npx * npy;
    E.scale = m_grid / (double)E.n_pts;
    int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    const double dxp = 2.0 / npx, dyp = 2.0 / npy;
    for (int ix = 0; ix < npx; ++ix) {
      // np.linspace(-1+dx/2, 1-dx/2, npx)
      double a0 = -1 + dxp / 2, a1 = 1 - dxp / 2;
      double u = npx > 1 ? a0 + (a1 - a0) / (double)(npx - 1) * (double)ix : a0;
      if (npx > 1 && ix == npx - 1) u = a1;
      for (int iy = 0; iy < npy; ++iy) {
        double b0 = -1 + dyp / 2, b1 = 1 - dyp / 2;
        double v = npy > 1 ? b0 + (b1 - b0) / (double)(npy - 1) * (double)iy : b0;
        if (npy > 1 && iy == npy - 1) v = b1;
        // A = R * diag(Lx, Ly, 1)/2 applied to the canonical point (u, v, 0)
        BliPoint B;
        B.ongrid = 0;
        for (int a = 0; a < 3; ++a) {
          double q = ((R[3 * a] * (Lx / 2.0)) * u + (R[3 * a + 1] * (Ly / 2.0)) * v) + pos[3 * e + a];
          pts.push_back(q);
          B.c[a] = closest_node(n[a], d[a], q);
          if (std::fabs(grid_pos(n[a], d[a], B.c[a]) - q) < thr) B.ongrid |= (1 << a);
          int l = (B.ongrid >> a) & 1 ? B.c[a] : B.c[a] - h;
          int u2 = (B.ongrid >> a) & 1 ? B.c[a] : B.c[a] + h;
          lo[a] = std::min(lo[a], l);
          hi[a] = std::max(hi[a], u2);
        }
        bps.push_back(B);
      }
    }
    for (int a = 0; a < 3; ++a) {
      E.lo[a] = std::max(lo[a], 0);
      int up = std::min(hi[a], n[a] - 1);
      E.nb[a] = std::max(up - E.lo[a] + 1, 0);
    }
    E.box_off = box_total;
    box_total += (long long)E.nb[0] * E.nb[1] * E.nb[2];
  }
  const long long n_pts = (long long)bps.size();
  if (s->Vin >= (1LL << 31)) { set_error("inner grid too large for the source mask scan"); return LIFU_ERR_INVALID; }

  cudaStream_t st = s->stream;
  // Memory.  Rebuilding the geometry is a per-candidate-pose operation (Protocol.simulate_candidates): every cudaMalloc /
  // cudaFree on a device that holds gigabytes of solver state costs milliseconds to tens of milliseconds and synchronises,
  // so the temporaries come out of ONE workspace kept on the handle (grown when needed) and the result buffers are
  // reused while their capacity suffices.
#define BLI_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return LIFU_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)
  const int TW = 2 * h + 1;
  const long long cap = std::max<long long>(std::min<long long>(box_total, s->Vin), 1);   // upper bound on selected points
  size_t sel_bytes = 0, scan_bytes = 0;
  cub::CountingInputIterator<long long> counting(0);
  BLI_CUDA(cub::DeviceSelect::Flagged(nullptr, sel_bytes, counting, (unsigned char*)nullptr, (long long*)nullptr, (long long*)nullptr,
                                      (int)s->Vin, st));
  BLI_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, (int)(cap + 1), st));
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_el = 0;
  const size_t o_bp = o_el + up(sizeof(BliElem) * n_el);
  const size_t o_pts = o_bp + up(sizeof(BliPoint) * n_pts);
  const size_t o_tab = o_pts + up(sizeof(double) * 3 * n_pts);
  const size_t o_wbox = o_tab + up(sizeof(double) * 3 * TW * n_pts);
  const size_t o_tbox = o_wbox + up(sizeof(float) * std::max<long long>(box_total, 1));
  const size_t o_mask = o_tbox + up(std::max<long long>(box_total, 1));
  const size_t o_nsel = o_mask + up(s->Vin);
  const size_t o_cnt = o_nsel + up(sizeof(long long));
  const size_t o_tmp = o_cnt + up(sizeof(int) * (cap + 1));
  const size_t ws_bytes = o_tmp + up(std::max(sel_bytes, scan_bytes));
  if (ws_bytes > s->bli_ws_cap) {
    if (s->d_bli_ws) { cudaFree(s->d_bli_ws); s->d_bli_ws = nullptr; s->bli_ws_cap = 0; }
    BLI_CUDA(cudaMalloc(&s->d_bli_ws, ws_bytes));
    s->bli_ws_cap = ws_bytes;
  }
  char* ws = reinterpret_cast<char*>(s->d_bli_ws);
  BliElem* d_el = reinterpret_cast<BliElem*>(ws + o_el);
  BliPoint* d_bp = reinterpret_cast<BliPoint*>(ws + o_bp);
  double* d_pts = reinterpret_cast<double*>(ws + o_pts);
  double* d_tab = reinterpret_cast<double*>(ws + o_tab);
  float* d_wbox = reinterpret_cast<float*>(ws + o_wbox);
  unsigned char* d_tbox = reinterpret_cast<unsigned char*>(ws + o_tbox);
  unsigned char* d_mask = reinterpret_cast<unsigned char*>(ws + o_mask);
  long long* d_nsel = reinterpret_cast<long long*>(ws + o_nsel);
  int* d_cnt = reinterpret_cast<int*>(ws + o_cnt);
  void* d_tmp = ws + o_tmp;
  BLI_CUDA(cudaMemcpyAsync(d_el, elems.data(), sizeof(BliElem) * n_el, cudaMemcpyHostToDevice, st));
  BLI_CUDA(cudaMemcpyAsync(d_bp, bps.data(), sizeof(BliPoint) * n_pts, cudaMemcpyHostToDevice, st));
  BLI_CUDA(cudaMemcpyAsync(d_pts, pts.data(), sizeof(double) * 3 * n_pts, cudaMemcpyHostToDevice, st));
  BLI_CUDA(cudaMemsetAsync(d_mask, 0, s->Vin, st));

  k_bli_tables<<<grid_blocks(s, n_pts * 3 * TW, 256), 256, 0, st>>>(d_pts, d_bp, n_pts, h, n[0], n[1], n[2],
                                                                    d[0], d[1], d[2], d_tab);
  long long max_box = 1;
  for (auto& E : elems) max_box = std::max(max_box, (long long)E.nb[0] * E.nb[1] * E.nb[2]);
  dim3 gg((unsigned)std::min<long long>((max_box + 127) / 128, 4096), n_el);
  k_bli_gather<<<gg, 128, 0, st>>>(d_el, n_el, d_bp, d_tab, h, d_wbox, d_tbox);
  k_bli_mark<<<gg, 128, 0, st>>>(d_el, d_tbox, n[0], n[1], d_mask);
  BLI_CUDA(cudaGetLastError());

  // compact the mask into sorted linear indices (x fastest == matlab_find order)
  s->geometry_set = false;
  if (cap > s->idx_cap || !s->d_idx) {
    if (s->d_idx) { cudaFree(s->d_idx); s->d_idx = nullptr; }
    if (s->d_row_ptr) { cudaFree(s->d_row_ptr); s->d_row_ptr = nullptr; }
    s->idx_cap = 0;
    BLI_CUDA(cudaMalloc(&s->d_idx, sizeof(long long) * cap));
    BLI_CUDA(cudaMalloc(&s->d_row_ptr, sizeof(int) * (cap + 1)));
    s->idx_cap = cap;
  }
  size_t tmp_bytes = sel_bytes;
  BLI_CUDA(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, counting, d_mask, s->d_idx, d_nsel, (int)s->Vin, st));
  long long n_src = 0;
  BLI_CUDA(cudaMemcpyAsync(&n_src, d_nsel, sizeof(long long), cudaMemcpyDeviceToHost, st));
  BLI_CUDA(cudaStreamSynchronize(st));
  s->n_src = n_src;
  s->n_el = n_el;

  BLI_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (n_src + 1), st));
  if (n_src > 0) {
    k_bli_rows<false><<<grid_blocks(s, n_src, 128), 128, 0, st>>>(s->d_idx, n_src, d_el, n_el, d_wbox, d_tbox,
                                                                  n[0], n[1], d_cnt, nullptr, nullptr, nullptr);
  }
  size_t tmp2 = scan_bytes;
  BLI_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp2, d_cnt, s->d_row_ptr, (int)(n_src + 1), st));
  int nnz = 0;
  BLI_CUDA(cudaMemcpyAsync(&nnz, s->d_row_ptr + n_src, sizeof(int), cudaMemcpyDeviceToHost, st));
  BLI_CUDA(cudaStreamSynchronize(st));
  s->nnz = nnz;
  if ((long long)nnz > s->nnz_cap || !s->d_col) {
    if (s->d_col) { cudaFree(s->d_col); s->d_col = nullptr; }
    if (s->d_w) { cudaFree(s->d_w); s->d_w = nullptr; }
    s->nnz_cap = 0;
    const long long want = std::max<long long>((long long)nnz + nnz / 8, 1);      // some slack: poses differ by a few points
    BLI_CUDA(cudaMalloc(&s->d_col, sizeof(int) * want));
    BLI_CUDA(cudaMalloc(&s->d_w, sizeof(float) * want));
    s->nnz_cap = want;
  }
  if (n_src > 0) {
    k_bli_rows<true><<<grid_blocks(s, n_src, 128), 128, 0, st>>>(s->d_idx, n_src, d_el, n_el, d_wbox, d_tbox,
                                                                 n[0], n[1], nullptr, s->d_row_ptr, s->d_col, s->d_w);
  }
  BLI_CUDA(cudaGetLastError());
  BLI_CUDA(cudaStreamSynchronize(st));
#undef BLI_CUDA
  s->geometry_set = true;
  return upload_source_points(s);
}

}  // namespace lifu
