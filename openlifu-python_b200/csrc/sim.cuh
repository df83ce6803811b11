// sim.cuh -- the opaque handle behind lifu_sim.
#pragma once
#include <functional>

#include "common.cuh"

namespace lifu {
// One rank's share of a z-slab decomposed grid (BASELINE.json config C5).
struct SlabCtx {
  bool on = false;
  bool ready = false;               // creation completed on this rank (destroy then synchronises with the peers)
  int rank = 0, G = 1;
  int z0 = 0, Nzl = 0;              // expanded planes [z0, z0 + Nzl) live here
  int Nyl = 0, ky0 = 0;             // ky rows [ky0, ky0 + Nyl) of the transposed spectra live here
  int jz_lo = 0, jz_n = 0;          // inner (sensor) planes owned
  int med_lo = 0, med_n = 0;        // inner planes of the medium maps this rank reads (incl. the +1 halo)
  long long Vl = 0, Hl = 0;         // local real voxels, local half-spectrum elements (= transposed elements)
  long long src_i0 = 0, src_i1 = 0; // this rank's range of the sorted source points
  int exchange = 0;                 // 1 NCCL send/recv, 2 peer stores over NVLink (CUDA IPC)
  void* comm = nullptr;             // ncclComm_t
  float2* xbuf = nullptr;           // [ H x4 | T x4 ] one allocation so one IPC handle covers both
  float2* pack = nullptr;           // [4][Hl] staging of the NCCL exchange
  float2** d_peer = nullptr;        // device table [G] of every rank's xbuf (own entry = local pointer)
  std::vector<void*> opened;        // IPC mappings to close
  float* d_bar = nullptr;           // barrier / small reductions
  cudaStream_t xs = nullptr;        // exchange stream (push kernels, barriers, NCCL calls)
  cudaEvent_t ev[32] = {};          // producer -> exchange and exchange -> consumer hand-offs of one time step
  int push_ctas_per_sm = 4;         // SM share of the NVLink-bound push kernels (transforms run beside them)
  bool overlap = false;             // per-field exchanges on `xs` beside the transforms (LIFU_SLAB_OVERLAP overrides)
  cufftHandle r2c2d = 0, c2r2d = 0, c2c1d = 0;
  bool plans = false;
  void* d_fftwork = nullptr;
};
}  // namespace lifu

struct lifu_sim {
  lifu_grid grid{};
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_sm = 148;

  // geometry
  int N[3]{}, n[3]{}, pml[3]{}, Nxh = 0;
  long long V = 0, Vin = 0;            // expanded / inner voxels of the WHOLE grid
  long long Vloc = 0, Vsens = 0;       // voxels held here (== V, Vin unless this is one slab of a decomposed grid)
  long long Vh = 0, RS = 0, CS = 0;    // local half-spectrum elements; strides between batched real / complex fields
  cudaStream_t own_stream = nullptr;   // created when a slab handle was given the default stream
  double c_ref = 0.0;
  bool tables_ready = false;

  // device buffers (all owned)
  std::vector<void*> allocs;
  float* d_tables = nullptr;   // all 1-D tables packed
  lifu::StepParams P{};
  lifu::SourceParams S{};

  // medium
  bool medium_set = false;
  bool homogeneous = true, absorbing = false;
  int alpha_mode = 0;
  float alpha_power = 0.9f;
  float *d_c0e = nullptr, *d_rho0e = nullptr, *d_alphae = nullptr;
  float c0_s = 0, rho0_s = 0, alpha_s = 0;
  float* d_med = nullptr;      // packed derived maps
  long long src_pts_cap = 0;   // capacity (points) of d_lin_exp / d_scale
  double* d_two_z = nullptr;   // 2 * density * sound_speed on the inner grid (float64, x fastest) for the intensity
  double two_z_s = 0.0;        // ... or one value
  int two_z_mode = 0;          // 0 not set, 1 scalar, 2 map

  // source geometry (device)
  long long n_src = 0, nnz = 0;
  int n_el = 0;
  long long* d_idx = nullptr;        // inner-grid linear indices
  long long* d_lin_exp = nullptr;
  int* d_row_ptr = nullptr;
  int* d_col = nullptr;
  float* d_w = nullptr;
  float* d_scale = nullptr;
  bool geometry_set = false;
  long long idx_cap = 0, nnz_cap = 0;  // capacities of d_idx / d_row_ptr (points) and d_col / d_w (non-zeros)
  void* d_bli_ws = nullptr;            // workspace of the GPU source-geometry build (bli.cu), kept across rebuilds
  size_t bli_ws_cap = 0;

  // drive
  float* d_base = nullptr;
  int n_base = 0;
  int* d_delay = nullptr;
  float* d_gain = nullptr;
  int drive_n_el = 0;
  int max_delay = 0;
  int source_mode = 0;
  bool drive_set = false;

  // FFT plans (cuFFT, shared work area)
  cufftHandle r2c1 = 0, r2c3 = 0, c2r1 = 0, c2r3 = 0, r2c2 = 0, c2r2 = 0;
  bool plans_ready = false;
  void* d_work = nullptr;
  size_t work_bytes = 0;

  // CUDA graphs of one time step: [0] source active, [1] source off
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  bool use_graph = true;

  // pipeline v2 (fused hand-written FFT passes); used when every axis is 64 or 256 and the medium is lossless
  int pipeline = 0;            // 0 auto, 1 force v1 (cuFFT), 2 force v2
  bool v2_ready = false;
  int R[3] = {0, 0, 0};        // radix per axis (N = R*R)
  lifu::V2Params Q{};
  float4* d_tw[3] = {nullptr, nullptr, nullptr};
  float4* d_mul4 = nullptr;    // dpy4, dny4, dpz4, dnz4 packed
  long long slab_planes_alloc = 0;
  bool z_tma = false;          // z passes through the persistent TMA-fed kernels (fft_z_tma.cuh)
  unsigned char tmH[128] __attribute__((aligned(64))) = {};   // CUtensorMap of H4[comp][z][ky][kx]
  unsigned char tmS[128] __attribute__((aligned(64))) = {};   // CUtensorMap of the source slab spectrum
  bool last_used_v2 = false;
  bool v2_wide = false;        // the grid has a non-square axis (128, 512, 768, 1024): passes of fft_wide.cuh / wide.cu
  bool v2_ygrad_split = false; // gradient y-inverse with one component per CTA (LIFU_V2_YGRAD=split)
  float* d_fk = nullptr;       // steady-state source: [2][RS] filtered basis fields
  float* d_qsrc = nullptr;     // [2][nws] time coefficients
  float* d_qcur = nullptr;     // coefficients of the current step
  float* d_coef = nullptr;     // [2][n_el] per-element spatial coefficients (set-up)
  size_t qsrc_cap = 0, coef_cap = 0;

  // pipeline v3 (fft_gen.cuh): the fused passes for any 2/3/5/7-smooth grid
  bool v3_ready = false, last_used_v3 = false;
  lifu::GParams G{};
  float2* d_gtw[3] = {nullptr, nullptr, nullptr};
  long long v3_slab_planes = 0;
  int v3_ts = 256, v3_tx = 256;   // threads per CTA of the strided / x passes (LIFU_V3_TS / LIFU_V3_TX)

  // per-stage profiling (lifu_profile_stages)
  bool prof_on = false;
  int prof_used = 0;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<const char*> prof_names;
  std::vector<double> prof_bytes;

  // slab decomposition over several GPUs (slab.cuh); sl.on == false for the single-GPU paths
  lifu::SlabCtx sl{};

  lifu_stats last{};
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

namespace lifu {
int dev_alloc(lifu_sim* s, void** p, size_t bytes);
int bli_build(lifu_sim* s, int n_el, const double* pos, const double* size, const double* ang,
              double tol, int ups);
int upload_source_points(lifu_sim* s);
// pipeline "wide" (wide.cu): axis lengths N = A x B it covers, one time step on the handle's stream
bool wide_ab(int n, int* A, int* B);
int wide_enqueue_step(lifu_sim* s, int kind, int* n_kernels, const std::function<void(const char*, double)>& mark,
                      const std::function<int()>& barrier);
void wide_pm_crop(lifu_sim* s);
inline int grid_blocks(const lifu_sim* s, long long n, int threads, int per_sm = 8) {
  long long need = (n + threads - 1) / threads;
  long long cap = (long long)s->n_sm * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}
}  // namespace lifu
