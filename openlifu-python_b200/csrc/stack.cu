// stack.cu -- the fields of every focus of one plan kept on the device (C ABI: lifu_stack_*).
//
// What it replaces on the host: the per-focus packaging of run_simulation
// (/root/reference/src/openlifu/sim/kwave_if.py:131-146), the xa.concat over foci of Protocol.calc_solution
// (plan/protocol.py:340-347), the in-place rescaling of Solution.scale (plan/solution.py:334-336) and the
// aggregation over foci (plan/protocol.py:382-392: p_min / p_max -> max over foci, intensity -> mean over foci).
// The solver's p_max / p_min never leave HBM between lifu_run and the beam analysis (lifu_analysis_set_focus takes
// the device pointers of lifu_stack_pointers); one device -> host copy hands the finished stack to the caller.
//
// Layout: [focus][z][y][x] (x fastest) per variable: p_max float32, pnp = -p_min float32, intensity float64 --
// the arrays of the returned Dataset, bit for bit (same IEEE operations as the numpy expressions they replace).
#include <algorithm>

#include "sim.cuh"

namespace lifu {

// Packaging of kwave_if.py:136-141: pnp = -p_min (float32), intensity = 1e-4 * p_min^2 / (2 Z) with the float32 square and
// scale and the float64 divide of the numpy expression (same arithmetic as k_package of step_kernels.cuh).
__global__ void k_stack_package(const float* __restrict__ pmin, const double* __restrict__ two_z, double two_z_s,
                                float* __restrict__ pnp, double* __restrict__ inten, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = pmin[i];
    pnp[i] = -p;
    const float sq = __fmul_rn(__fmul_rn(p, p), 1e-4f);
    inten[i] = __ddiv_rn((double)sq, two_z ? two_z[i] : two_z_s);
  }
}

// numpy >= 2 semantics of `float32_array *= np.float64(s)`: product formed in float64, rounded to float32
__global__ void k_stack_scale(float* __restrict__ pmax, float* __restrict__ pnp, double* __restrict__ inten, long long n,
                              double s, double s2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    pmax[i] = __double2float_rn(__dmul_rn((double)pmax[i], s));
    pnp[i] = __double2float_rn(__dmul_rn((double)pnp[i], s));
    inten[i] = __dmul_rn(inten[i], s2);
  }
}

// max over foci of the two pressures (NaN-skipping, NaN when every focus is NaN = np.nanmax), mean over foci of the
// intensity: sum in focus order, one division (= np.mean along the leading axis); NaNs, if any, are skipped (np.nanmean)
__global__ void k_stack_aggregate(const float* __restrict__ pmax, const float* __restrict__ pnp, const double* __restrict__ inten,
                                  long long n, int nf, float* __restrict__ o_pmax, float* __restrict__ o_pnp,
                                  double* __restrict__ o_int) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = pmax[i], b = pnp[i];
    double sum = inten[i];
    bool any_nan = isnan(sum);
    for (int f = 1; f < nf; ++f) {
      a = fmaxf(a, pmax[(long long)f * n + i]);
      b = fmaxf(b, pnp[(long long)f * n + i]);
      const double v = inten[(long long)f * n + i];
      any_nan |= isnan(v);
      sum = __dadd_rn(sum, v);
    }
    double mean = __ddiv_rn(sum, (double)nf);
    if (any_nan) {
      double s2 = 0.0; int c = 0; bool first = true;
      for (int f = 0; f < nf; ++f) {
        const double v = inten[(long long)f * n + i];
        if (isnan(v)) continue;
        s2 = first ? v : __dadd_rn(s2, v);
        first = false; ++c;
      }
      mean = c ? __ddiv_rn(s2, (double)c) : nan("");
    }
    o_pmax[i] = a;
    o_pnp[i] = b;
    o_int[i] = mean;
  }
}

}  // namespace lifu

struct lifu_stack {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n[3] = {0, 0, 0};
  int n_foci = 0;
  long long V = 0;
  float* d_pmax = nullptr;     // [F][V]
  float* d_pnp = nullptr;      // [F][V]
  double* d_int = nullptr;     // [F][V]
  float* d_agg_f = nullptr;    // [2][V] aggregated pressures
  double* d_agg_d = nullptr;   // [V] aggregated intensity
  std::vector<char> filled;
  int blocks = 148 * 8;
};

using namespace lifu;

static void stack_free(lifu_stack* k) {
  if (!k) return;
  cudaSetDevice(k->device);
  cudaFree(k->d_pmax); cudaFree(k->d_pnp); cudaFree(k->d_int); cudaFree(k->d_agg_f); cudaFree(k->d_agg_d);
  delete k;
}

extern "C" {

int lifu_stack_create(int device, void* cuda_stream, const int32_t n[3], int32_t n_foci, lifu_stack** out) {
  if (!n || !out || n_foci <= 0 || n[0] <= 0 || n[1] <= 0 || n[2] <= 0) { set_error("lifu_stack_create: bad argument"); return LIFU_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("lifu_stack_create: no CUDA device %d (this library has no CPU fallback)", device);
    return LIFU_ERR_CUDA;
  }
  LIFU_CUDA(cudaSetDevice(device));
  lifu_stack* k = new lifu_stack();
  k->device = device;
  k->stream = (cudaStream_t)cuda_stream;
  for (int a = 0; a < 3; ++a) k->n[a] = n[a];
  k->n_foci = n_foci;
  k->V = (long long)n[0] * n[1] * n[2];
  k->filled.assign(n_foci, 0);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  k->blocks = sms * 8;
  const size_t fv = (size_t)n_foci * (size_t)k->V;
  cudaError_t e = cudaMalloc(&k->d_pmax, fv * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&k->d_pnp, fv * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&k->d_int, fv * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&k->d_agg_f, 2 * (size_t)k->V * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&k->d_agg_d, (size_t)k->V * sizeof(double));
  if (e != cudaSuccess) {
    set_error("lifu_stack_create: cudaMalloc for %d foci of %lld voxels failed: %s", n_foci, k->V, cudaGetErrorString(e));
    stack_free(k);
    return e == cudaErrorMemoryAllocation ? LIFU_ERR_NOMEM : LIFU_ERR_CUDA;
  }
  *out = k;
  return LIFU_OK;
}

int lifu_stack_destroy(lifu_stack* k) {
  if (!k) return LIFU_OK;
  cudaSetDevice(k->device);
  cudaStreamSynchronize(k->stream);
  stack_free(k);
  return LIFU_OK;
}

int lifu_stack_put(lifu_stack* k, int32_t focus, lifu_sim* s) {
  if (!k || !s || focus < 0 || focus >= k->n_foci) { set_error("lifu_stack_put: bad argument"); return LIFU_ERR_INVALID; }
  if (s->sl.on) { set_error("lifu_stack_put: not available on a slab handle"); return LIFU_ERR_STATE; }
  if (s->device != k->device) { set_error("lifu_stack_put: solver on device %d, stack on device %d", s->device, k->device); return LIFU_ERR_INVALID; }
  if (s->n[0] != k->n[0] || s->n[1] != k->n[1] || s->n[2] != k->n[2]) { set_error("lifu_stack_put: grid sizes differ"); return LIFU_ERR_INVALID; }
  if (s->last.steps <= 0) { set_error("lifu_stack_put: call lifu_run first"); return LIFU_ERR_STATE; }
  if (s->two_z_mode == 0) { set_error("lifu_stack_put: call lifu_set_two_z first"); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(k->device));
  NvtxRange nvtx_r("lifu_stack_put");
  // the solver's stream orders this after the time loop; the stack's stream is joined through an event
  cudaStream_t st = s->stream;
  const size_t off = (size_t)focus * (size_t)k->V;
  LIFU_CUDA(cudaMemcpyAsync(k->d_pmax + off, s->P.pmax, sizeof(float) * (size_t)k->V, cudaMemcpyDeviceToDevice, st));
  k_stack_package<<<k->blocks, 256, 0, st>>>(s->P.pmin, s->two_z_mode == 2 ? s->d_two_z : nullptr, s->two_z_s, k->d_pnp + off,
                                       k->d_int + off, k->V);
  LIFU_CUDA(cudaGetLastError());
  LIFU_CUDA(cudaStreamSynchronize(st));
  k->filled[focus] = 1;
  return LIFU_OK;
}

int lifu_stack_scale(lifu_stack* k, int32_t focus, double s, double s2) {
  if (!k || focus < 0 || focus >= k->n_foci) { set_error("lifu_stack_scale: bad argument"); return LIFU_ERR_INVALID; }
  if (!k->filled[focus]) { set_error("lifu_stack_scale: focus %d has no fields yet", focus); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(k->device));
  const size_t off = (size_t)focus * (size_t)k->V;
  k_stack_scale<<<k->blocks, 256, 0, k->stream>>>(k->d_pmax + off, k->d_pnp + off, k->d_int + off, k->V, s, s2);
  LIFU_CUDA(cudaGetLastError());
  LIFU_CUDA(cudaStreamSynchronize(k->stream));
  return LIFU_OK;
}

int lifu_stack_pointers(lifu_stack* k, int32_t focus, float** p_max, float** pnp, double** intensity) {
  if (!k || focus < 0 || focus >= k->n_foci) { set_error("lifu_stack_pointers: bad argument"); return LIFU_ERR_INVALID; }
  const size_t off = (size_t)focus * (size_t)k->V;
  if (p_max) *p_max = k->d_pmax + off;
  if (pnp) *pnp = k->d_pnp + off;
  if (intensity) *intensity = k->d_int + off;
  return LIFU_OK;
}

int lifu_stack_get(lifu_stack* k, int32_t focus, float* p_max, float* pnp, double* intensity) {
  if (!k || focus < -1 || focus >= k->n_foci) { set_error("lifu_stack_get: bad argument"); return LIFU_ERR_INVALID; }
  for (int f = 0; f < k->n_foci; ++f)
    if ((focus < 0 || f == focus) && !k->filled[f]) { set_error("lifu_stack_get: focus %d has no fields yet", f); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(k->device));
  const size_t off = focus < 0 ? 0 : (size_t)focus * (size_t)k->V;
  const size_t cnt = (focus < 0 ? (size_t)k->n_foci : 1) * (size_t)k->V;
  if (p_max) LIFU_CUDA(cudaMemcpyAsync(p_max, k->d_pmax + off, cnt * sizeof(float), cudaMemcpyDefault, k->stream));
  if (pnp) LIFU_CUDA(cudaMemcpyAsync(pnp, k->d_pnp + off, cnt * sizeof(float), cudaMemcpyDefault, k->stream));
  if (intensity) LIFU_CUDA(cudaMemcpyAsync(intensity, k->d_int + off, cnt * sizeof(double), cudaMemcpyDefault, k->stream));
  LIFU_CUDA(cudaStreamSynchronize(k->stream));
  return LIFU_OK;
}

int lifu_stack_aggregate(lifu_stack* k, float* p_max_max, float* pnp_max, double* intensity_mean) {
  if (!k || !p_max_max || !pnp_max || !intensity_mean) { set_error("lifu_stack_aggregate: null argument"); return LIFU_ERR_INVALID; }
  for (int f = 0; f < k->n_foci; ++f)
    if (!k->filled[f]) { set_error("lifu_stack_aggregate: focus %d has no fields yet", f); return LIFU_ERR_STATE; }
  LIFU_CUDA(cudaSetDevice(k->device));
  NvtxRange nvtx_r("lifu_stack_aggregate");
  k_stack_aggregate<<<k->blocks, 256, 0, k->stream>>>(k->d_pmax, k->d_pnp, k->d_int, k->V, k->n_foci, k->d_agg_f,
                                                      k->d_agg_f + k->V, k->d_agg_d);
  LIFU_CUDA(cudaGetLastError());
  LIFU_CUDA(cudaMemcpyAsync(p_max_max, k->d_agg_f, (size_t)k->V * sizeof(float), cudaMemcpyDefault, k->stream));
  LIFU_CUDA(cudaMemcpyAsync(pnp_max, k->d_agg_f + k->V, (size_t)k->V * sizeof(float), cudaMemcpyDefault, k->stream));
  LIFU_CUDA(cudaMemcpyAsync(intensity_mean, k->d_agg_d, (size_t)k->V * sizeof(double), cudaMemcpyDefault, k->stream));
  LIFU_CUDA(cudaStreamSynchronize(k->stream));
  return LIFU_OK;
}

}  // extern "C"
