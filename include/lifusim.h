/* lifusim.h -- C ABI of the B200-native k-space acoustic solver behind
 * openlifu.sim.run_simulation.
 *
 * This is the drop-in boundary.  Every entry point replaces one piece of what the
 * reference adapter /root/reference/src/openlifu/sim/kwave_if.py does through
 * k-wave-python 0.4.0 (kWaveGrid / kWaveArray / kWaveMedium / kSensor / kSource /
 * kspaceFirstOrder3D -> HDF5 -> kspaceFirstOrder-OMP|CUDA subprocess).  The reference
 * has no FFI of its own (it is pure Python); the binding a
 * maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary.
 *   - every data pointer may be a HOST or a DEVICE pointer (unified virtual addressing
 *     resolves it; copies use cudaMemcpyDefault on the handle's stream).
 *   - 3-D arrays are "x fastest" = Fortran order of the reference's (Nx,Ny,Nz) arrays
 *     = what kwave_if.py:132-141 reshapes with order='F'.
 *   - every function returns 0 on success or a negative lifu_status; the message of the
 *     last failure on the calling thread is lifu_last_error().  No exception crosses.
 *   - a handle is bound to one (device, stream); it is not thread-safe; distinct handles
 *     may be used concurrently (one per GPU when foci are sharded).
 *   - the handle never frees or retains caller buffers after the call returns.
 */
#ifndef LIFUSIM_H
#define LIFUSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIFUSIM_ABI_VERSION 1

typedef enum lifu_status {
  LIFU_OK = 0,
  LIFU_ERR_INVALID = -1,   /* bad argument */
  LIFU_ERR_CUDA = -2,      /* CUDA runtime failure */
  LIFU_ERR_CUFFT = -3,     /* cuFFT failure */
  LIFU_ERR_STATE = -4,     /* call order violated (e.g. run before set_medium) */
  LIFU_ERR_NOMEM = -5
} lifu_status;

/* alpha_mode of kWaveMedium (kwave_if.py:56,62).  The k-Wave binary receives only
 * alpha_coeff/alpha_power, so LIFU_ALPHA_BINARY (both terms) is what the reference's
 * OMP path computes (SURVEY.md ledger A7). */
typedef enum lifu_alpha_mode {
  LIFU_ALPHA_BINARY = 0,        /* absorption (tau) and dispersion (eta) terms */
  LIFU_ALPHA_NO_DISPERSION = 1, /* eta = 0 */
  LIFU_ALPHA_NO_ABSORPTION = 2  /* tau = 0 */
} lifu_alpha_mode;

typedef enum lifu_source_mode {
  LIFU_SOURCE_ADDITIVE = 0,               /* k-space corrected additive source (k-Wave default) */
  LIFU_SOURCE_ADDITIVE_NO_CORRECTION = 1
} lifu_source_mode;

/* Grid + time axis.  Replaces get_kgrid (kwave_if.py:13-27) and the pml_auto /
 * pml_inside=False expansion done inside kspaceFirstOrder3D (kwave_if.py:117-122). */
typedef struct lifu_grid {
  int32_t n[3];     /* inner grid size Nx,Ny,Nz = len(coords) (kwave_if.py:19) */
  int32_t pml[3];   /* PML thickness per axis; any entry < 0 -> pml_auto for that axis */
  double d[3];      /* spacing in metres (kwave_if.py:20) */
  double dt;        /* time step [s] */
  int32_t nt;       /* number of time steps */
  double pml_alpha; /* <= 0 -> 2.0 (k-Wave default) */
  double c_ref;     /* <= 0 -> max(c0) once the medium is set (k-Wave default) */
} lifu_grid;

typedef struct lifu_stats {
  int64_t voxels;          /* expanded grid voxels advanced per step (incl. PML) */
  int32_t n_exp[3];        /* expanded grid size */
  int32_t pml[3];
  int32_t steps;           /* time steps executed */
  int32_t source_steps;    /* steps on which the source was active */
  int64_t kernel_launches; /* hand-written kernels launched inside the time loop */
  int64_t fft_launches;    /* FFT transforms executed inside the time loop */
  double loop_ms;          /* CUDA-event time of the time loop on the handle's stream */
  double setup_ms;         /* CUDA-event time of uploads/initialisation in lifu_run */
  double bytes_per_voxel_step; /* algorithmic bytes model, averaged over steps */
  int32_t homogeneous;
  int32_t absorbing;
  int32_t steady_source_steps; /* steps on which the source ran as q1(t) F1 + q2(t) F2 (filtered once) */
  int32_t reserved;
} lifu_stats;

typedef struct lifu_sim lifu_sim; /* opaque */

/* ---- pure host helpers (no GPU needed) -------------------------------------------- */
int lifu_abi_version(void);
const char* lifu_last_error(void);
/* kWaveGrid.makeTime(c_ref, cfl) as called at kwave_if.py:22-23; writes nt, dt. */
int lifu_make_time(const int32_t n[3], const double d[3], double c_ref, double cfl,
                   int32_t* nt, double* dt);
/* get_optimal_pml_size for pml_auto=True (kwave_if.py:118): range [10,40]. */
int lifu_pml_auto(const int32_t n[3], int32_t pml_out[3]);

/* Number of CUDA devices this library can run on (sm_100 only); 0 when there is none or the driver is missing.
 * Replaces the NVML probe of util/checkgpu.py:6-14 where pynvml is not installed: Protocol.calc_solution(use_gpu=None)
 * (plan/protocol.py:294-295) asks the library that will do the work.  Never fails. */
int lifu_device_count(int32_t* count);

/* ---- lifecycle --------------------------------------------------------------------- */
int lifu_create(const lifu_grid* grid, int device, void* cuda_stream, lifu_sim** out);
int lifu_destroy(lifu_sim* sim);

/* Replaces get_medium (kwave_if.py:49-63) + the medium expansion / staggered density /
 * absorption coefficients derived inside kspaceFirstOrder3D.  Maps are float32 on the
 * INNER grid, x fastest.  homogeneous != 0: each pointer addresses ONE float. */
int lifu_set_medium(lifu_sim* sim, const float* c0, const float* rho0, const float* alpha_db,
                    float alpha_power, int alpha_mode, int homogeneous);

/* Same as lifu_set_medium for heterogeneous media, taking the arrays the reference holds --
 * params['sound_speed'|'density'|'attenuation'].data: float64, (Nx,Ny,Nz), C order (sim_setup.py:107-116,
 * seg_method.py:84-97) -- without a host-side conversion: each map is copied as one block and rounded to
 * float32 (round-to-nearest, = data_cast='single') and re-laid out on the device.  stride = element
 * strides of (x, y, z); must describe a dense array.  alpha_db may be NULL.  Not for slab handles. */
int lifu_set_medium_f64(lifu_sim* sim, const double* c0, const double* rho0, const double* alpha_db,
                        const int64_t stride[3], float alpha_power, int alpha_mode);

/* Heterogeneous medium from a LABEL volume and per-label tables: what SegmentationMethod._map_params
 * (seg/seg_method.py:84-97) expands on the host into three float64 maps, expanded on the device instead -- one byte
 * per voxel is uploaded.  labels: uint8, inner grid, element strides of (x, y, z) (dense); tables: n_labels <= 32
 * float64 entries, rounded to float32 (data_cast='single'); labels >= n_labels give 0 like the reference's initial
 * value (and are then rejected: the sound speed must be positive).  alpha_db may be NULL.  Not for slab handles. */
int lifu_set_medium_labels(lifu_sim* sim, const uint8_t* labels, const int64_t stride[3], int32_t n_labels,
                           const double* c0, const double* rho0, const double* alpha_db, float alpha_power,
                           int alpha_mode);

/* Replaces get_karray + get_array_binary_mask + the weight half of
 * get_distributed_source_signal (kwave_if.py:29-47,75-77): off-grid rectangular elements
 * spread with the truncated-sinc band-limited interpolant, computed on the GPU.
 *   pos_m     [n_el*3] element centres in metres in the k-Wave grid frame, i.e. already
 *             shifted by array_offset = -mean(coords) (kwave_if.py:108)
 *   size_m    [n_el*2] (width, length)
 *   angle_deg [n_el*3] (el, az, roll) = rotations about x, y, z (element.py:216-226)
 * Host pointers (float64).  Returns the number of source points through n_src. */
int lifu_set_elements(lifu_sim* sim, int32_t n_el, const double* pos_m, const double* size_m,
                      const double* angle_deg, double bli_tolerance, int32_t upsampling_rate,
                      int64_t* n_src);

/* Explicit alternative to lifu_set_elements: caller supplies the mask indices (sorted,
 * x-fastest linear, INNER grid) and the weights as CSR over source points. */
int lifu_set_source_geometry(lifu_sim* sim, const int64_t* idx, const int32_t* row_ptr,
                             const int32_t* col_elem, const float* w, int64_t n_src,
                             int64_t nnz, int32_t n_el);

/* Read back the geometry built by either call above (for parity checks / caching --
 * the reference's dead grid-weights cache API, db/database.py:41-55).
 * Any output pointer may be NULL.  Sizes: idx[n_src], row_ptr[n_src+1], col/w[nnz]. */
int lifu_get_source_sizes(lifu_sim* sim, int64_t* n_src, int64_t* nnz, int32_t* n_el);
int lifu_get_source_geometry(lifu_sim* sim, int64_t* idx, int32_t* row_ptr, int32_t* col_elem,
                             float* w);

/* Replaces the drive-signal half: kwave_if.py:101-103 + Transducer.calc_output
 * (xdc/transducer.py:95-112).  Element e emits gains[e]*base_signal[t - delay_samples[e]].
 * delay_samples[e] = int(delay_e/dt) is computed by the host (integer, bit-exact). */
int lifu_set_drive(lifu_sim* sim, const float* base_signal, int32_t n_base,
                   const int32_t* delay_samples, const float* gains, int32_t n_el,
                   int source_mode);

/* Replaces kspaceFirstOrder3D(...) with sensor.record=['p_max','p_min'] over the whole
 * inner grid (kwave_if.py:65-69,124-129).  p_max/p_min: float32[Nx*Ny*Nz] x fastest; p_min is
 * the raw minimum (the adapter negates it, kwave_if.py:136).  stats may be NULL. */
int lifu_run(lifu_sim* sim, float* p_max, float* p_min, lifu_stats* stats);

/* Replaces the packaging arithmetic of kwave_if.py:136-141 on the device.  lifu_set_two_z stores
 * 2 * density * sound_speed of the params maps (float64: ONE value, or one per inner-grid voxel, x fastest;
 * host or device pointer).  After lifu_run (whose p_max / p_min may then be NULL), lifu_get_packaged writes
 * p_max, pnp = -p_min (float32) and intensity = 1e-4 * p_min^2 / two_z (float32 square and scale, float64
 * divide -- the numpy expression, bit for bit) to the caller's buffers; p_max may be NULL. */
int lifu_set_two_z(lifu_sim* sim, const double* two_z, int64_t n);
int lifu_get_packaged(lifu_sim* sim, float* p_max, float* pnp, double* intensity);

/* ---- one oversized grid over several GPUs (BASELINE.json config C5) ----
 * k-Wave's binaries are single-device, so nothing in the reference binds these; they extend the
 * kspaceFirstOrder3D replacement (kwave_if.py:117-129) to grids that do not fit one GPU.  One process
 * per GPU of one NVSwitch node; rank r holds the expanded planes [r*Nz/G, (r+1)*Nz/G) of every field.
 * The expanded Ny and Nz must be multiples of G.  3-D transforms = local 2-D (x,y) transforms + one
 * exchange over NVLink + local 1-D z transforms.
 *   lifu_slab_unique_id   rank 0 calls it and broadcasts the 128 bytes (ncclUniqueId) to every rank
 *   lifu_create_slab      collective: builds the NCCL communicator and maps the peers' exchange buffers
 *   lifu_slab_layout      which planes this rank owns / must be given
 * On a slab handle:
 *   lifu_set_medium_planes  maps cover the inner planes [plane0, plane0+n_planes) of the INNER grid and
 *                           must include [medium_z0, medium_z0+medium_nz); lifu_set_medium = all planes.
 *                           Collective (c_ref = global max c0).
 *   lifu_set_elements / lifu_set_source_geometry  take the FULL geometry on every rank (each keeps its part)
 *   lifu_run              collective; p_max/p_min receive this rank's inner planes
 *                         [sensor_z0, sensor_z0+sensor_nz), x fastest.
 *   lifu_destroy          collective. */
#define LIFU_NCCL_ID_BYTES 128
typedef enum lifu_exchange {
  LIFU_EXCHANGE_AUTO = 0,  /* peer stores when CUDA IPC mapping works on every rank, else NCCL */
  LIFU_EXCHANGE_NCCL = 1,  /* pack + grouped ncclSend/ncclRecv + unpack */
  LIFU_EXCHANGE_PEER = 2   /* kernels store straight into the destination rank's buffer over NVLink */
} lifu_exchange;
typedef struct lifu_slab_desc {
  int32_t rank, nranks;
  int32_t exchange;                      /* lifu_exchange */
  unsigned char nccl_id[LIFU_NCCL_ID_BYTES];
} lifu_slab_desc;
typedef struct lifu_slab_layout {
  int32_t rank, nranks, exchange;        /* exchange: the mode in use (1 or 2) */
  int32_t z0, nz;                        /* expanded planes held by this rank */
  int32_t sensor_z0, sensor_nz;          /* inner planes of p_max/p_min written by lifu_run (nz may be 0) */
  int32_t medium_z0, medium_nz;          /* inner planes of the medium maps this rank reads */
} lifu_slab_layout;
int lifu_slab_unique_id(unsigned char id[LIFU_NCCL_ID_BYTES]);
int lifu_create_slab(const lifu_grid* grid, int device, void* cuda_stream, const lifu_slab_desc* slab,
                     lifu_sim** out);
int lifu_slab_layout_of(lifu_sim* sim, lifu_slab_layout* out);
int lifu_set_medium_planes(lifu_sim* sim, const float* c0, const float* rho0, const float* alpha_db,
                           float alpha_power, int alpha_mode, int homogeneous, int32_t plane0,
                           int32_t n_planes);

/* Debug / test access to the state after lifu_run: which = 0 p, 1..3 u_x,u_y,u_z,
 * 4..6 rho_x,rho_y,rho_z; out: float32 on the EXPANDED grid (a slab handle: its own planes), x fastest. */
int lifu_get_field(lifu_sim* sim, int which, float* out);
int lifu_get_info(lifu_sim* sim, lifu_stats* stats);

/* Measurement aid (bench.py roofline): after a lifu_run, execute `reps` further time steps with a
 * CUDA event after every stage on the handle's stream and return the mean duration per stage
 * [ms], its name and its algorithmic bytes per voxel.  with_source selects the
 * source-active step variant.  names: max_stages * name_stride chars (may be NULL). */
int lifu_profile_stages(lifu_sim* sim, int reps, int with_source, int max_stages, char* names,
                        int name_stride, double* ms, double* bytes_per_voxel, int* n_stages);

/* ---- beam analysis of the simulated fields ------------------------------------
 * The O(V) passes of Solution.analyze (/root/reference/src/openlifu/plan/solution.py:135-281), which the
 * reference runs twice per calc_solution (inside Solution.scale, plan/protocol.py:372, and at :396): the
 * focus-frame distance map and the ellipsoid masks (solution_analysis.py:319-442 get_focus_matrix,
 * get_offset_grid, calc_dist_from_focus, get_mask), the masked maxima (solution.py:196-199,243-258), the
 * -3 dB value-weighted centroid of the main lobe (solution.py:200-208, find_centroid :306-317) and the
 * trilinear samples along the three focus-frame axes that the beam widths are read from
 * (interp_transformed_axis :444-487, get_beam_bounds :489-535).  Scalar post-processing (unit factors,
 * duty cycles, beam-width edges of the sampled lines, MI/TIC/power) stays on the host.
 *
 * One handle holds the fields of every focus of a solution on the device.  All geometry is float64 with
 * round-to-nearest operations in the reference's evaluation order, so masks, maxima and line samples are
 * bit-identical to the numpy evaluation; the centroid sums are accumulated in float64.
 *   lifu_analysis_create      n = (Nx,Ny,Nz) of the fields; x,y,z = float64 coordinate vectors (analysis units);
 *                             z_ok[Nz] = 1 where z > sidelobe_zmin (solution.py:194), NULL = all ones
 *   lifu_analysis_set_focus   pnp = float32 peak-negative-pressure field of this focus in its stored unit,
 *                             ipa = float64 intensity field; HOST or DEVICE pointers; stride = element strides
 *                             of (x,y,z) and must describe a dense array (C or Fortran order or any permutation)
 *   lifu_analysis_run_focus   line_pts: host float64 [(n_line[0]+n_line[1]+n_line[2])][3] sample points in grid
 *                             coordinates; line_vals: host float64, same count (NaN outside the grid) */
typedef struct lifu_analysis lifu_analysis; /* opaque */

typedef struct lifu_focus_query {
  double w[3][4];            /* rows 0..2 of inverse(get_focus_matrix(focus, origin)):
                                frame coordinate i = ((w[i][0]*x + w[i][1]*y) + w[i][2]*z) + w[i][3] */
  double aspect[3];          /* mainlobe_aspect_ratio: coordinate i is divided by aspect[i] */
  double mainlobe_radius;    /* main lobe: dist <  mainlobe_radius */
  double sidelobe_radius;    /* side lobe: dist >  sidelobe_radius and z_ok */
  double centroid_factor;    /* centroid keeps main-lobe voxels with pnp > float32(main_pnp * centroid_factor) */
  float pnp_scale;           /* float32 factor from the stored pressure unit to the analysis unit (Pa -> MPa: 1e-6f) */
  int32_t n_line[3];         /* samples on the lateral / elevation / axial line (0 = none) */
} lifu_focus_query;

typedef struct lifu_focus_metrics {
  double main_pnp, side_pnp, global_pnp;   /* maxima of float32(pnp*pnp_scale) over main lobe / side lobe / z_ok;
                                              NaN when the selection holds no non-NaN value */
  double main_ipa, side_ipa, global_ipa;   /* the same selections of this focus' intensity */
  double main_ipa_all, global_ipa_all;     /* main-lobe / z_ok maximum over EVERY focus' intensity field
                                              (the reference masks the whole stack, solution.py:245,266) */
  int64_t n_main, n_side, n_global;        /* voxels in each selection */
  double cen_w, cen_wx, cen_wy, cen_wz;    /* sum of w, w*x, w*y, w*z over the -3 dB part of the main lobe */
  int64_t n_centroid;
  double kernel_ms;                        /* CUDA-event time from the first to the last kernel of this call
                                              (includes the host round trip between the two passes) */
  double reduce_ms;                        /* CUDA-event time of k_focus_reduce alone (the pass that streams the fields) */
} lifu_focus_metrics;

int lifu_analysis_create(int device, void* cuda_stream, const int32_t n[3], int32_t n_foci, const double* x,
                         const double* y, const double* z, const uint8_t* z_ok, lifu_analysis** out);
int lifu_analysis_set_focus(lifu_analysis* a, int32_t focus, const float* pnp, const double* ipa,
                            const int64_t stride[3]);
int lifu_analysis_run_focus(lifu_analysis* a, int32_t focus, const lifu_focus_query* q, const double* line_pts,
                            lifu_focus_metrics* out, double* line_vals);
int lifu_analysis_destroy(lifu_analysis* a);

/* ---- plan stack: the fields of every focus of one plan kept on the device ---------------
 * Replaces, on the device, the per-focus packaging of run_simulation (kwave_if.py:131-146), the xa.concat over foci
 * (plan/protocol.py:340-347), the in-place rescaling of Solution.scale (plan/solution.py:334-336) and the
 * aggregation over foci (plan/protocol.py:382-392).  Layout per variable: [focus][z][y][x] (x fastest); p_max and
 * pnp = -p_min float32, intensity float64 -- the arrays of the returned Dataset, bit for bit. */
typedef struct lifu_stack lifu_stack; /* opaque */
int lifu_stack_create(int device, void* cuda_stream, const int32_t n[3], int32_t n_foci, lifu_stack** out);
int lifu_stack_destroy(lifu_stack* stack);
/* Package the result of the solver's last lifu_run into slot `focus` (needs lifu_set_two_z): no host round trip. */
int lifu_stack_put(lifu_stack* stack, int32_t focus, lifu_sim* sim);
/* Solution.scale on one focus: pressures *= s (product in float64, rounded to float32: numpy >= 2 semantics of
 * `float32_array *= np.float64(s)`), intensity *= s2.  The caller passes s2 = `s ** 2` as ITS arithmetic forms it
 * (numpy's scalar power is not always the correctly rounded s * s). */
int lifu_stack_scale(lifu_stack* stack, int32_t focus, double s, double s2);
/* Device pointers of one focus (for lifu_analysis_set_focus with strides (1, Nx, Nx*Ny)). */
int lifu_stack_pointers(lifu_stack* stack, int32_t focus, float** p_max, float** pnp, double** intensity);
/* Copy out one focus (focus >= 0) or the whole stack (focus = -1); host or device destinations, any may be NULL. */
int lifu_stack_get(lifu_stack* stack, int32_t focus, float* p_max, float* pnp, double* intensity);
/* max over foci of p_max and pnp (NaN-skipping), mean over foci of the intensity (sum in focus order, one division). */
int lifu_stack_aggregate(lifu_stack* stack, float* p_max_max, float* pnp_max, double* intensity_mean);

#ifdef __cplusplus
}
#endif
#endif /* LIFUSIM_H */
