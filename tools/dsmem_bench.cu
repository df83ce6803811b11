// dsmem_bench.cu -- feasibility probe for the fused y-z pass of DESIGN.md section 6: how fast can the CTAs of one
// thread-block cluster transpose a tile through distributed shared memory on B200?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench tools/dsmem_bench.cu && ./dsmem_bench
//
// Model of the fused pass: a cluster of C CTAs holds one 256 x 256 complex (float2) y-z plane, CTA r owning 256 / C
// z rows.  After the y transforms every CTA scatters its rows so that CTA q receives the 256 / C y columns it will
// transform along z: an all-to-all in which (C - 1) / C of the bytes cross SM boundaries.  The kernel below performs
// exactly that scatter (8-byte st.shared::cluster stores, coalesced along the destination row) ITER times per launch and
// reports bytes moved per SM per clock and the aggregate rate.  A local-only variant (every store into the CTA's own
// shared memory) gives the reference point.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

constexpr int NY = 256, NZ = 256;
constexpr int U = 4;          // stores per thread per pass
constexpr int W = 2;          // float2 elements per store (2 = 16-byte st.shared::cluster)
constexpr int THREADS = 512;

template <int C, bool REMOTE>
__global__ void __launch_bounds__(THREADS) k_transpose(float2* out, int iters, int rep) {
  extern __shared__ __align__(16) unsigned char smraw[];
  // two buffers of (NZ / C) x NY float2 each: src = my z rows (all y), dst = my y columns (all z), as [y_local][z]
  constexpr int ROWS = NZ / C;                 // z rows owned (and y columns received)
  constexpr int SP = NY + 1;                   // padded source pitch: the strided reads below stay bank-conflict free
  float2* src = reinterpret_cast<float2*>(smraw);
  float2* dst = src + ROWS * SP;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  for (int i = threadIdx.x; i < ROWS * SP; i += blockDim.x) src[i] = make_float2((float)(rank * 1000 + i), 1.f);
  cluster.sync();
  for (int it = 0; it < iters; ++it) {
    // element (zl, y) of my rows goes to CTA q = y / ROWS, position [y % ROWS][rank * ROWS + zl]
    // thread mapping: consecutive threads take consecutive zl (destination-contiguous 8-byte stores)
    for (int r = 0; r < rep; ++r)                       // rep scatters per cluster barrier (rep = 0: barrier cost alone)
      for (int e0 = threadIdx.x; e0 < ROWS * NY / W; e0 += blockDim.x * U) {
        float2 v[U][W];
#pragma unroll
        for (int k = 0; k < U; ++k) {                       // U x W independent loads in flight per thread
          const int e = (e0 + k * blockDim.x) * W;          // W consecutive z rows of one y column
#pragma unroll
          for (int w = 0; w < W; ++w) v[k][w] = src[(e % ROWS + w) * SP + e / ROWS];
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int e = (e0 + k * blockDim.x) * W;
          const int zl = e % ROWS, y = e / ROWS;
          const int q = y / ROWS, yl = y % ROWS;
          float2* pq = REMOTE ? cluster.map_shared_rank(dst, q) : dst;      // one mapa per store
          float2* d = pq + yl * NZ + rank * ROWS + zl;
          if (W == 2) *reinterpret_cast<float4*>(d) = make_float4(v[k][0].x, v[k][0].y, v[k][W - 1].x, v[k][W - 1].y);
          else d[0] = v[k][0];
        }
      }
    cluster.sync();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = dst[rank];
}

template <int C, bool REMOTE>
static void run(int n_sm, int iters, int rep) {
  constexpr int ROWS = NZ / C;
  const size_t smem = ((size_t)ROWS * (NY + 1) + (size_t)ROWS * NY) * sizeof(float2);
  cudaFuncSetAttribute(k_transpose<C, REMOTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_transpose<C, REMOTE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int clusters = n_sm / C;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * C);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  float2* out;
  cudaMalloc(&out, sizeof(float2) * clusters * C);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaError_t err = cudaLaunchKernelEx(&cfg, k_transpose<C, REMOTE>, out, 2, rep);   // warm-up
  cudaDeviceSynchronize();
  if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    printf("cluster %2d %s: launch failed (%s)\n", C, REMOTE ? "dsmem" : "local", cudaGetErrorString(err));
    return;
  }
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, k_transpose<C, REMOTE>, out, iters, rep);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double bytes_per_cta = (double)iters * rep * ROWS * NY * sizeof(float2);
  const double total = bytes_per_cta * clusters * C;
  if (rep == 0) {
    printf("cluster %2d: cluster.sync() alone: %.2f us each\n", C, ms * 1e3 / iters);
    cudaFree(out);
    return;
  }
  printf("cluster %2d %s: %3d CTAs x %3zu KB smem, %d iters: %.3f ms  -> %.1f GB/s per SM, %.2f TB/s aggregate, %.1f B/clk/SM at %d MHz (%.0f %% of bytes remote)\n",
         C, REMOTE ? "dsmem" : "local", clusters * C, smem >> 10, iters, ms, bytes_per_cta / (ms * 1e-3) / 1e9,
         total / (ms * 1e-3) / 1e12, bytes_per_cta / (ms * 1e-3) / (khz * 1e3), khz / 1000, REMOTE ? 100.0 * (C - 1) / C : 0.0);
  cudaFree(out);
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 200;
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs: %d; tile: %d x %d float2 plane (512 KB) per cluster\n", n_sm, NY, NZ);
  const int rep = argc > 2 ? atoi(argv[2]) : 8;
  run<8, true>(n_sm, iters, 0);
  run<8, false>(n_sm, iters, rep);
  run<8, true>(n_sm, iters, rep);
  run<16, true>(n_sm, iters, 0);
  run<16, false>(n_sm, iters, rep);
  run<16, true>(n_sm, iters, rep);
  return 0;
}
