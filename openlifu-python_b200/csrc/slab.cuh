// slab.cuh -- one oversized grid decomposed into z slabs over the GPUs of one NVSwitch node
// (BASELINE.json config C5).  One process per GPU; rank r holds the expanded
// planes [r Nz/G, (r+1) Nz/G) of every real field.  k-Wave's binaries, which the reference drives from
// /root/reference/src/openlifu/sim/kwave_if.py:117-129, are single-device: there is no reference
// counterpart of this file, only of the arithmetic it distributes.
//
// A 3-D transform becomes: local 2-D (x, y) transforms (cuFFT, batch = local planes), ONE exchange that
// trades the z split for a ky split, local 1-D transforms along z (cuFFT, strided).  Everything between
// the transforms is ours:
//   * the exchange itself.  Peer mode: the kernel that applies the (x, y) derivative multipliers stores
//     its result straight into the destination rank's transposed buffer over NVLink (CUDA IPC mappings
//     of every rank's exchange buffer), so the transpose costs no extra pass over HBM and no staging
//     copy; a one-float ncclAllReduce on the same stream is the barrier.  NCCL mode (fallback and
//     comparison): the same kernel packs per-destination blocks, grouped ncclSend/ncclRecv move them.
//   * the spectral operators in the transposed layout T[z][kyl][kx] (kappa, i kz e^{+-i kz dz/2},
//     cos(c_ref k dt/2) source filter, fractional Laplacians).
// NVLink traffic is kept minimal: the pressure gradient sends 1 field forward and 2 back (the x and y
// derivative multipliers commute with the z transforms and are applied after the return trip).
//
// Layouts (float2 complex, x fastest):
//   H[f][zl][ky][kx]   2-D spectra of the local planes, kx = 0..Nx/2      (Hl = Nzl*Ny*Nxh per field)
//   T[f][z][kyl][kx]   all planes, this rank's Ny/G ky rows                (same element count)
// Both live in ONE allocation per rank, xbuf = [H x4 | T x4], so one IPC handle exposes them.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "sim.cuh"
#include "step_kernels.cuh"

namespace lifu {

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: liblifusim.so must load (and the single-GPU paths must run) without libnccl.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  // One libnccl.so.2 per process (the loader dedupes by soname): reuse the copy the host already loaded
  // (PyTorch bundles a newer NCCL than the system one and fails to import on top of an older copy), then
  // the path the host names in LIFU_NCCL_LIB, then the system library.
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  const char* envp = getenv("LIFU_NCCL_LIB");
  if (!h && envp && envp[0]) h = dlopen(envp, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("slab decomposition needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror()); return nullptr; }
  bool ok = true;
  auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) ok = false; return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (!ok) { set_error("libnccl.so.2 lacks a required symbol"); return nullptr; }
  api.lib = h;
  return &api;
}

#define LIFU_NCCL(call)                                                                    \
  do {                                                                                     \
    ncclResult_t r__ = (call);                                                             \
    if (r__ != ncclSuccess) {                                                              \
      lifu::set_error("%s:%d NCCL error in %s: %s", __FILE__, __LINE__, #call,             \
                      lifu::nccl_api()->GetErrorString(r__));                              \
      return LIFU_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------------
// device-side view
struct SlabParams {
  int Nx, Ny, Nxh, Nzl, NzG, Nyl, G;
  int z0, ky0;
  long long Hl;                 // stride between fields in either layout (>= Hn: sized for the padded rows of wide.cu)
  long long Hn;                 // elements per field = Nzl*Ny*Nxh = NzG*Nyl*Nxh
  float2* H;                    // local [4][Hl]
  float2* T;                    // local [4][Hl]
  float2* const* peer;          // [G] xbuf of every rank (H at +0, T at +4*Hl), or pack staging in NCCL mode
  long long peer_T_off;         // element offset of T inside a peer entry
  int zoff_T;                   // plane offset of this rank's block inside the destination T (z0 peer mode, 0 staging)
  int kyoff_H;                  // ky offset of this rank's block inside the destination H (ky0 peer mode, 0 staging)
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Forward exchange, sender side: H[f] (local planes, all ky) -> block (z, kyl, kx) of rank q = ky / Nyl.
// KIND 0: plain; 1: x i kx e^{-i kx dx/2} (velocity x); 2: x i ky e^{-i ky dy/2} (velocity y) -- the in-plane
// divergence multipliers ride on the exchange.  Peer mode writes into rank q's T at plane z0 + zl; staging
// mode writes block q of the pack buffer (a block is [planes][Nyl][Nxh] in both).
template <int KIND>
__global__ void __launch_bounds__(256) k_slab_push_fwd(StepParams P, SlabParams S, int f) {
  const int rows = S.Nzl * S.Ny;                       // (zl, ky) rows of Nxh elements
  const int kx = threadIdx.x & 31;
  const int row_in_blk = threadIdx.x >> 5;             // 8 rows per CTA pass
  const float2* src = S.H + f * S.Hl;
  for (int row = blockIdx.x * 8 + row_in_blk; row < rows; row += gridDim.x * 8) {
    const int zl = row / S.Ny, ky = row - zl * S.Ny;
    const int q = ky / S.Nyl, kyl = ky - q * S.Nyl;
    float2* dst = S.peer[q] + S.peer_T_off + f * S.Hl + ((long long)(S.zoff_T + zl) * S.Nyl + kyl) * S.Nxh;
    const float2* sp = src + (long long)row * S.Nxh;
    float2 my = make_float2(1.f, 0.f);
    if (KIND == 2) my = P.dny[ky];
    for (int x = kx; x < S.Nxh; x += 32) {
      float2 v = sp[x];
      if (KIND == 1) v = cmulf(P.dnx[x], v);
      if (KIND == 2) v = cmulf(my, v);
      dst[x] = v;
    }
  }
}

// Backward exchange, sender side: T[f] (all planes, local ky rows) -> rank q = z / Nzl, H[fd] at
// (zl, ky0 + kyl, kx).  Staging mode is not needed for this direction (blocks of T are contiguous), so
// this kernel is peer-mode only; NCCL mode receives into the pack buffer and runs k_slab_unpack_back.
__global__ void __launch_bounds__(256) k_slab_push_back(SlabParams S, int f0, int nf, int fd0, int fdstep) {
  const int rows = S.NzG * S.Nyl;
  const int kx = threadIdx.x & 31;
  const int row_in_blk = threadIdx.x >> 5;
  for (int f = 0; f < nf; ++f) {
    const float2* src = S.T + (f0 + f) * S.Hl;
    for (int row = blockIdx.x * 8 + row_in_blk; row < rows; row += gridDim.x * 8) {
      const int z = row / S.Nyl, kyl = row - z * S.Nyl;
      const int q = z / S.Nzl, zl = z - q * S.Nzl;
      float2* dst = S.peer[q] + (fd0 + f * fdstep) * S.Hl + ((long long)zl * S.Ny + S.ky0 + kyl) * S.Nxh;
      const float2* sp = src + (long long)row * S.Nxh;
      for (int x = kx; x < S.Nxh; x += 32) dst[x] = sp[x];
    }
  }
}

// NCCL mode, receiver side of the backward exchange: pack[f][q][zl][kyl][kx] -> H[fd][zl][q*Nyl + kyl][kx]
__global__ void __launch_bounds__(256) k_slab_unpack_back(SlabParams S, const float2* __restrict__ pack, int nf, int fd0, int fdstep) {
  const int rows = S.G * S.Nzl * S.Nyl;
  const int kx = threadIdx.x & 31;
  const int row_in_blk = threadIdx.x >> 5;
  for (int f = 0; f < nf; ++f) {
    for (int row = blockIdx.x * 8 + row_in_blk; row < rows; row += gridDim.x * 8) {
      const int kyl = row % S.Nyl;
      const int t = row / S.Nyl;
      const int zl = t % S.Nzl, q = t / S.Nzl;
      float2* dst = S.H + (fd0 + f * fdstep) * S.Hl + ((long long)zl * S.Ny + q * S.Nyl + kyl) * S.Nxh;
      const float2* sp = pack + f * S.Hl + (long long)row * S.Nxh;
      for (int x = kx; x < S.Nxh; x += 32) dst[x] = sp[x];
    }
  }
}

// ------------------------------------------------------------------ spectral operators, T layout
// index -> (kx, global ky, kz)
__device__ __forceinline__ void t_index(const SlabParams& S, long long i, int& kx, int& ky, int& kz) {
  kx = (int)(i % S.Nxh);
  long long t = i / S.Nxh;
  ky = S.ky0 + (int)(t % S.Nyl);
  kz = (int)(t / S.Nyl);
}

// pressure gradient: T0 <- kappa p^ / N ; T1 <- i kz e^{+i kz dz/2} kappa p^ / N
__global__ void __launch_bounds__(256) k_slab_grad_z(StepParams P, SlabParams S) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.Hn; i += (long long)gridDim.x * blockDim.x) {
    int kx, ky, kz;
    t_index(S, i, kx, ky, kz);
    const float kap = kappa_of(P.ax2[kx] + P.ay2[ky] + P.az2[kz]) * P.invN;
    float2 v = S.T[i];
    v.x *= kap; v.y *= kap;
    S.T[i] = v;
    S.T[S.Hl + i] = cmulf(P.dpz[kz], v);
  }
}

// one field of the velocity divergence / source, in place on T[f]:
// KIND 0: x kappa/N (x and y components: their i k multipliers rode on the exchange);
// KIND 1: x i kz e^{-i kz dz/2} kappa/N (z component);  KIND 2: x cos(c_ref k dt/2)/N (source field)
template <int KIND>
__global__ void __launch_bounds__(256) k_slab_div_z(StepParams P, SlabParams S, int f) {
  float2* T = S.T + f * S.Hl;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.Hn; i += (long long)gridDim.x * blockDim.x) {
    int kx, ky, kz;
    t_index(S, i, kx, ky, kz);
    const float a2 = P.ax2[kx] + P.ay2[ky] + P.az2[kz];
    float2 a = T[i];
    if (KIND == 2) {
      const float cs = cosf(sqrtf(a2)) * P.invN;
      a.x *= cs; a.y *= cs;
    } else {
      const float kap = kappa_of(a2) * P.invN;
      a.x *= kap; a.y *= kap;
      if (KIND == 1) a = cmulf(P.dnz[kz], a);
    }
    T[i] = a;
  }
}

// absorption operators, one field in place: T[f] *= k^(2e)/N with e = (y-2)/2 (tau operand) or (y-1)/2 (eta operand)
__global__ void __launch_bounds__(256) k_slab_absorb_z(StepParams P, SlabParams S, int f, float e) {
  float2* T = S.T + f * S.Hl;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.Hn; i += (long long)gridDim.x * blockDim.x) {
    int kx, ky, kz;
    t_index(S, i, kx, ky, kz);
    const float k2 = P.kx2[kx] + P.ky2[ky] + P.kz2[kz];
    const float n1 = k2 > 0.f ? powf(k2, e) * P.invN : 0.f;
    float2 a = T[i];
    a.x *= n1; a.y *= n1;
    T[i] = a;
  }
}

// after the return trip of the gradient: H3 = IFFT_z[kappa p^]  ->  H0 = i kx e^{+..} H3, H1 = i ky e^{+..} H3
__global__ void __launch_bounds__(256) k_slab_grad_xy(StepParams P, SlabParams S) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < S.Hn; i += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(i % S.Nxh);
    const int ky = (int)((i / S.Nxh) % S.Ny);
    const float2 a = S.H[3 * S.Hl + i];
    S.H[i] = cmulf(P.dpx[kx], a);
    S.H[S.Hl + i] = cmulf(P.dpy[ky], a);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
struct SlabHost {
  static SlabParams params(const lifu_sim* s, bool staging) {
    const SlabCtx& L = s->sl;
    SlabParams S{};
    S.Nx = s->N[0]; S.Ny = s->N[1]; S.Nxh = s->Nxh; S.Nzl = L.Nzl; S.NzG = s->N[2]; S.Nyl = L.Nyl; S.G = L.G;
    S.z0 = L.z0; S.ky0 = L.ky0; S.Hl = L.Hl; S.Hn = (long long)s->Nxh * s->N[1] * L.Nzl;
    S.H = L.xbuf; S.T = L.xbuf + 4 * L.Hl;
    S.peer = L.d_peer;
    S.peer_T_off = staging ? 0 : 4 * L.Hl;
    S.zoff_T = staging ? 0 : L.z0;
    S.kyoff_H = staging ? 0 : L.ky0;
    return S;
  }
};

inline int slab_barrier(lifu_sim* s, cudaStream_t st) {
  NcclApi* N = nccl_api();
  LIFU_NCCL(N->AllReduce(s->sl.d_bar, s->sl.d_bar + 1, 1, ncclFloat, ncclSum, (ncclComm_t)s->sl.comm, st));
  return LIFU_OK;
}

// all-reduce (max) of a few host floats
inline int slab_allreduce_max(lifu_sim* s, float* v, int n) {
  NcclApi* N = nccl_api();
  LIFU_CUDA(cudaMemcpyAsync(s->sl.d_bar + 8, v, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
  LIFU_NCCL(N->AllReduce(s->sl.d_bar + 8, s->sl.d_bar + 8, n, ncclFloat, ncclMax, (ncclComm_t)s->sl.comm, s->stream));
  LIFU_CUDA(cudaMemcpyAsync(v, s->sl.d_bar + 8, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream));
  LIFU_CUDA(cudaStreamSynchronize(s->stream));
  return LIFU_OK;
}

inline int slab_init(lifu_sim* s, const lifu_slab_desc* d) {
  SlabCtx& L = s->sl;
  NcclApi* N = nccl_api();
  if (!N) return LIFU_ERR_STATE;
  L.on = true;
  L.rank = d->rank; L.G = d->nranks;
  L.Nzl = s->N[2] / L.G; L.z0 = L.rank * L.Nzl;
  L.Nyl = s->N[1] / L.G; L.ky0 = L.rank * L.Nyl;
  const int lo = std::max(L.z0 - s->pml[2], 0), hi = std::min(L.z0 + L.Nzl - s->pml[2], s->n[2]);
  L.jz_lo = std::min(lo, s->n[2]); L.jz_n = std::max(hi - lo, 0);
  const int mlo = std::min(std::max(L.z0 - s->pml[2], 0), s->n[2] - 1);
  const int mhi = std::min(std::max(L.z0 + L.Nzl - s->pml[2], 0), s->n[2] - 1);   // + halo plane
  L.med_lo = mlo; L.med_n = mhi - mlo + 1;
  L.Vl = (long long)s->N[0] * s->N[1] * L.Nzl;
  L.Hl = (long long)s->Nxh * s->N[1] * L.Nzl;
  // grids the fused passes cover (wide.cu) keep their half spectra with a row pitch of round_up(Nxh, 16): size the exchange
  // buffers for that layout (the library-FFT path uses Hl only as the stride between fields)
  if (wide_ab(s->N[0], nullptr, nullptr) && wide_ab(s->N[1], nullptr, nullptr) && wide_ab(s->N[2], nullptr, nullptr))
    L.Hl = round_up(s->Nxh, 16) * s->N[1] * L.Nzl;
  ncclUniqueId id;
  static_assert(sizeof(id.internal) == LIFU_NCCL_ID_BYTES, "ncclUniqueId size");
  memcpy(id.internal, d->nccl_id, LIFU_NCCL_ID_BYTES);
  ncclComm_t comm = nullptr;
  LIFU_NCCL(N->CommInitRank(&comm, L.G, id, L.rank));
  L.comm = comm;
  LIFU_CHECK(dev_alloc(s, (void**)&L.d_bar, sizeof(float) * 64));
  LIFU_CUDA(cudaMemsetAsync(L.d_bar, 0, sizeof(float) * 64, s->stream));
  // exchange buffers: a plain cudaMalloc (IPC needs one; never from a pool)
  LIFU_CUDA(cudaMalloc((void**)&L.xbuf, sizeof(float2) * 8 * L.Hl));
  LIFU_CUDA(cudaMemsetAsync(L.xbuf, 0, sizeof(float2) * 8 * L.Hl, s->stream));
  LIFU_CHECK(dev_alloc(s, (void**)&L.d_peer, sizeof(float2*) * L.G));
  std::vector<float2*> peers(L.G, nullptr);
  int want = d->exchange;
  if (want != 1) {
    // exchange the IPC handles of xbuf through NCCL and map every peer's buffer
    cudaIpcMemHandle_t mine;
    float ok = cudaIpcGetMemHandle(&mine, L.xbuf) == cudaSuccess ? 1.f : 0.f;
    if (ok == 0.f) cudaGetLastError();
    unsigned char* d_h = nullptr;
    LIFU_CUDA(cudaMalloc((void**)&d_h, sizeof(mine) * (L.G + 1)));
    LIFU_CUDA(cudaMemcpyAsync(d_h + sizeof(mine) * L.G, &mine, sizeof(mine), cudaMemcpyHostToDevice, s->stream));
    LIFU_NCCL(N->AllGather(d_h + sizeof(mine) * L.G, d_h, sizeof(mine), ncclChar, comm, s->stream));
    std::vector<cudaIpcMemHandle_t> all(L.G);
    LIFU_CUDA(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * L.G, cudaMemcpyDeviceToHost, s->stream));
    LIFU_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d_h);
    for (int q = 0; q < L.G && ok == 1.f; ++q) {
      if (q == L.rank) { peers[q] = L.xbuf; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0.f; break; }
      L.opened.push_back(p);
      peers[q] = (float2*)p;
    }
    float agree = -ok;                                       // max(-ok) == 0  <=>  some rank failed
    LIFU_CHECK(slab_allreduce_max(s, &agree, 1));
    if (agree == 0.f) {
      for (void* p : L.opened) cudaIpcCloseMemHandle(p);
      L.opened.clear();
      if (want == 2) { set_error("lifu_create_slab: CUDA IPC peer mapping failed on at least one rank"); return LIFU_ERR_CUDA; }
      want = 1;
    } else {
      want = 2;
    }
  }
  L.exchange = want;
  if (want == 1) {
    LIFU_CHECK(dev_alloc(s, (void**)&L.pack, sizeof(float2) * 4 * L.Hl));
    // staging "peers": block q of the pack buffer; a block is [Nzl][Nyl][Nxh] = Hl / G elements per field
    for (int q = 0; q < L.G; ++q) peers[q] = L.pack + (long long)q * ((long long)L.Nzl * L.Nyl * s->Nxh);
  }
  LIFU_CUDA(cudaMemcpyAsync(L.d_peer, peers.data(), sizeof(float2*) * L.G, cudaMemcpyHostToDevice, s->stream));
  LIFU_CUDA(cudaStreamSynchronize(s->stream));
  // exchange stream: every push kernel, barrier and NCCL call of the time loop runs on it, in the same order on
  // every rank; the handle's stream does the transforms and the real-space kernels and meets it through events
  LIFU_CUDA(cudaStreamCreateWithFlags(&L.xs, cudaStreamNonBlocking));
  for (auto& e : L.ev) LIFU_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (const char* e = getenv("LIFU_SLAB_PUSH_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 8) L.push_ctas_per_sm = v; }
  // overlap: per-field exchanges on the exchange stream beside the transforms of the next field (measured +3 % at
  // G = 2, where half of every exchange is a local copy); otherwise one barrier per batch of fields on one stream
  L.overlap = L.G >= 2;
  if (const char* e = getenv("LIFU_SLAB_OVERLAP")) L.overlap = e[0] == '1';
  return LIFU_OK;
}

inline void slab_destroy(lifu_sim* s) {
  SlabCtx& L = s->sl;
  if (!L.on) return;
  if (L.plans) { cufftDestroy(L.r2c2d); cufftDestroy(L.c2r2d); cufftDestroy(L.c2c1d); }
  if (L.xs) { cudaStreamSynchronize(L.xs); }
  cudaFree(L.d_fftwork);
  for (void* p : L.opened) cudaIpcCloseMemHandle(p);
  L.opened.clear();
  // every rank must have stopped touching peers' buffers before any of them is freed
  if (L.comm) {
    NcclApi* N = nccl_api();
    if (N && L.d_bar && L.ready) { N->AllReduce(L.d_bar, L.d_bar + 1, 1, ncclFloat, ncclSum, (ncclComm_t)L.comm, s->stream); cudaStreamSynchronize(s->stream); }
    cudaFree(L.xbuf);
    if (N) N->CommDestroy((ncclComm_t)L.comm);
    L.comm = nullptr;
  } else {
    cudaFree(L.xbuf);
  }
  L.xbuf = nullptr;
  if (L.xs) { cudaStreamDestroy(L.xs); L.xs = nullptr; }
  for (auto& e : L.ev) if (e) { cudaEventDestroy(e); e = nullptr; }
}

inline int slab_plans(lifu_sim* s) {
  SlabCtx& L = s->sl;
  if (L.plans) return LIFU_OK;
  const int Nx = s->N[0], Ny = s->N[1], Nxh = s->Nxh;
  int n2[2] = {Ny, Nx};
  int re[2] = {Ny, Nx}, ce[2] = {Ny, Nxh};
  size_t w = 0, ws = 0;
  LIFU_CUFFT(cufftCreate(&L.r2c2d)); LIFU_CUFFT(cufftSetAutoAllocation(L.r2c2d, 0));
  LIFU_CUFFT(cufftMakePlanMany(L.r2c2d, 2, n2, re, 1, Nx * Ny, ce, 1, Ny * Nxh, CUFFT_R2C, L.Nzl, &w)); ws = std::max(ws, w);
  LIFU_CUFFT(cufftCreate(&L.c2r2d)); LIFU_CUFFT(cufftSetAutoAllocation(L.c2r2d, 0));
  LIFU_CUFFT(cufftMakePlanMany(L.c2r2d, 2, n2, ce, 1, Ny * Nxh, re, 1, Nx * Ny, CUFFT_C2R, L.Nzl, &w)); ws = std::max(ws, w);
  const int PL = L.Nyl * Nxh;                               // elements per transposed plane
  int n1[1] = {s->N[2]};
  int e1[1] = {s->N[2]};
  LIFU_CUFFT(cufftCreate(&L.c2c1d)); LIFU_CUFFT(cufftSetAutoAllocation(L.c2c1d, 0));
  LIFU_CUFFT(cufftMakePlanMany(L.c2c1d, 1, n1, e1, PL, 1, e1, PL, 1, CUFFT_C2C, PL, &w)); ws = std::max(ws, w);
  LIFU_CUDA(cudaMalloc(&L.d_fftwork, std::max<size_t>(ws, 16)));
  LIFU_CUFFT(cufftSetWorkArea(L.r2c2d, L.d_fftwork));
  LIFU_CUFFT(cufftSetWorkArea(L.c2r2d, L.d_fftwork));
  LIFU_CUFFT(cufftSetWorkArea(L.c2c1d, L.d_fftwork));
  L.plans = true;
  return LIFU_OK;
}

// Forward exchange of one field, H[f] -> T[f] on every rank, enqueued on `st` (the exchange stream): the sender-side
// kernel (with the in-plane multiplier `kind`), then the barrier (peer mode) or the grouped send/recv (NCCL mode).
// The push kernel is NVLink-bound, so it is given a fraction of the SMs: transforms of the next field run beside it.
inline int slab_exchange_fwd(lifu_sim* s, cudaStream_t st, int f, int kind, bool barrier = true) {
  SlabCtx& L = s->sl;
  const bool staging = L.exchange == 1;
  SlabParams S = SlabHost::params(s, staging);
  const int rows = L.Nzl * s->N[1];
  const int gb = std::min((rows + 7) / 8, s->n_sm * (L.overlap ? L.push_ctas_per_sm : 8));
  if (kind == 1) k_slab_push_fwd<1><<<gb, 256, 0, st>>>(s->P, S, f);
  else if (kind == 2) k_slab_push_fwd<2><<<gb, 256, 0, st>>>(s->P, S, f);
  else k_slab_push_fwd<0><<<gb, 256, 0, st>>>(s->P, S, f);
  LIFU_CUDA(cudaGetLastError());
  if (!staging) return barrier ? slab_barrier(s, st) : LIFU_OK;
  NcclApi* N = nccl_api();
  const long long blk = (long long)L.Nzl * L.Nyl * s->Nxh;   // complex elements per (field, destination); Hl >= G * blk
  LIFU_NCCL(N->GroupStart());
  for (int q = 0; q < L.G; ++q) {
    LIFU_NCCL(N->Send(L.pack + f * L.Hl + q * blk, 2 * blk, ncclFloat, q, (ncclComm_t)L.comm, st));
    LIFU_NCCL(N->Recv(S.T + f * L.Hl + q * blk, 2 * blk, ncclFloat, q, (ncclComm_t)L.comm, st));
  }
  LIFU_NCCL(N->GroupEnd());
  return LIFU_OK;
}

// Backward exchange of one field, T[f] -> H[fd] on every rank, enqueued on `st`.
inline int slab_exchange_back(lifu_sim* s, cudaStream_t st, int f, int fd, bool barrier = true) {
  SlabCtx& L = s->sl;
  SlabParams S = SlabHost::params(s, false);
  const int rows = s->N[2] * L.Nyl;
  const int gb = std::min((rows + 7) / 8, s->n_sm * (L.overlap ? L.push_ctas_per_sm : 8));
  if (L.exchange == 2) {
    k_slab_push_back<<<gb, 256, 0, st>>>(S, f, 1, fd, 1);
    LIFU_CUDA(cudaGetLastError());
    return barrier ? slab_barrier(s, st) : LIFU_OK;
  }
  NcclApi* N = nccl_api();
  const long long blk = (long long)L.Nzl * L.Nyl * s->Nxh;
  LIFU_NCCL(N->GroupStart());
  for (int q = 0; q < L.G; ++q) {
    LIFU_NCCL(N->Send(S.T + f * L.Hl + q * blk, 2 * blk, ncclFloat, q, (ncclComm_t)L.comm, st));
    LIFU_NCCL(N->Recv(L.pack + f * L.Hl + q * blk, 2 * blk, ncclFloat, q, (ncclComm_t)L.comm, st));
  }
  LIFU_NCCL(N->GroupEnd());
  k_slab_unpack_back<<<std::min((rows + 7) / 8, s->n_sm * 8), 256, 0, st>>>(S, L.pack + f * L.Hl, 1, fd, 1);
  LIFU_CUDA(cudaGetLastError());
  return LIFU_OK;
}

}  // namespace lifu
