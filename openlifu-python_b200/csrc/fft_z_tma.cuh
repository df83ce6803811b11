// fft_z_tma.cuh -- z passes of pipeline v2 as persistent, TMA-fed kernels.
//
// Same arithmetic as k2_z_grad / k2_z_div / k2_z_absorb of fft_v2.cuh (which they replace for the regular kx
// tiles; the single Nyquist column still goes through those kernels on a 1-wide grid).  What changes is how
// the data reaches the SM.  A tile -- 16 consecutive kx of one ky row over ALL z, 16 x Nz x 8 bytes -- is one
// `cp.async.bulk.tensor` (TMA) box of the 4-D tensor H4[comp][z][ky][kx]: one elected thread issues it, the
// copy engine walks the Nz strided 128-byte rows, and an mbarrier flips when the bytes have landed.  CTAs are
// persistent (2 per SM) and keep the loads of the next two items in flight while they transform the current
// one, so the first-load latency that the short-lived CTAs of the old kernels paid per tile (about a fifth of
// their lifetime, profiles/r1_zdiv_source.md) is gone and no registers are spent on prefetch.
//
// Shared memory per CTA: 2 stages + 1 exchange buffer of Nz*16 complex (32 KB each at Nz = 256) + twiddles.
// A consumed stage doubles as the exchange buffer of the forward transform; the inverse transforms exchange
// through X.  Item n+2 is issued into item n's stage as soon as every thread is past the forward exchange.
#pragma once
#include <cuda.h>

#include "fft_v2.cuh"

namespace lifu {

enum { ZOP_GRAD = 0, ZOP_DIV = 1, ZOP_ABS = 2 };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(b)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int R>
struct ZTma {
  static constexpr int N = R * R;
  static constexpr int XCH = Strided<R>::XCH;                 // bytes of one stage / exchange buffer
  static constexpr int SMEM = 3 * XCH + Strided<R>::TW + 32 + 128;  // stages, X, twiddles, 2 mbarriers, alignment slack
};

// OP = ZOP_GRAD: items = tiles of H4[0];        out: H4[0] <- IFFT_z[kappa p^], H4[1] <- IFFT_z[i kz e^{+i kz dz/2} kappa p^]
// OP = ZOP_DIV : items = (tile, comp < ncomp);  in place: kappa (comp 0, 1), i kz e^{-i kz dz/2} kappa (comp 2),
//                                               cos(c_ref k dt/2) (comp 3, read from the source slab through tmS)
// OP = ZOP_ABS : items = (tile, comp < 2);      in place: k^(y-2) (comp 0), k^(y-1) (comp 1)
template <int R, int OP, int POLY>
__global__ void __launch_bounds__(16 * R, 2) k2_z_tma(StepParams P, V2Params Q, const __grid_constant__ CUtensorMap tmH,
                                                      const __grid_constant__ CUtensorMap tmS, int ncomp) {
  extern __shared__ __align__(16) unsigned char smraw_[];
  unsigned char* const smraw = smraw_ + ((128u - (smem_u32(smraw_) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  constexpr int N = R * R, XCH = ZTma<R>::XCH, THREADS = 16 * R;
  float2* const stg0 = reinterpret_cast<float2*>(smraw);
  float2* const X = reinterpret_cast<float2*>(smraw + 2 * XCH);
  float4* const tws = reinterpret_cast<float4*>(smraw + 3 * XCH);
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(smraw + 3 * XCH + Strided<R>::TW);
  const int tid = threadIdx.x, l = tid & 15, t = tid >> 4;
  const int ntile = Q.nxt * Q.Ny;
  const int nin = OP == ZOP_GRAD ? 1 : ncomp;
  const int mytiles = (int)blockIdx.x < ntile ? (ntile - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nitems = mytiles * nin;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int n) {   // elected thread: start the TMA load of item n into stage n & 1
    const int tile = (int)blockIdx.x + (n / nin) * (int)gridDim.x, comp = n % nin, s = n & 1;
    const int kx0 = (tile % Q.nxt) * 16, ky = tile / Q.nxt;
    unsigned char* dst = smraw + s * XCH;
    if (comp < 3) {
      mbar_expect_tx(&mbar[s], XCH);
      tma_load_4d(dst, &tmH, &mbar[s], kx0, ky, 0, comp);
    } else {
      mbar_expect_tx(&mbar[s], 16 * 8 * Q.nzs);
      tma_load_3d(dst, &tmS, &mbar[s], kx0, ky, 0);
    }
  };
  if (tid == 0) {
    if (nitems > 0) issue(0);
    if (nitems > 1) issue(1);
  }
  for (int i = tid; i <= N; i += THREADS) tws[i] = Q.tw4z[i];
  __syncthreads();
  const int zs = Q.Ny * Q.PH, jstep = R * zs;
  float kap[OP == ZOP_ABS ? 1 : R];
  float axy = 0.f;
  float2 v[R];
  for (int n = 0; n < nitems; ++n) {
    const int s = n & 1;
    const int tile = (int)blockIdx.x + (n / nin) * (int)gridDim.x, comp = n % nin;
    const int kx = (tile % Q.nxt) * 16 + l, ky = tile / Q.nxt;
    float2* const stg = stg0 + s * (XCH / 8);
    mbar_wait(&mbar[s], (n >> 1) & 1);
    if (OP != ZOP_DIV || comp < 3) {
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = stg[(t + R * j) * 16 + l];
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int zr = t + R * j - Q.z0s;
        v[j] = (zr >= 0 && zr < Q.nzs) ? stg[zr * 16 + l] : make_float2(0.f, 0.f);
      }
    }
    if (comp == 0) {
      if (OP == ZOP_ABS) {
        axy = P.kx2[kx] + P.ky2[ky];
      } else {
        axy = P.ax2[kx] + P.ay2[ky];
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) kap[OP == ZOP_ABS ? 0 : k1] = kappa_sel<POLY>(axy + P.az2[t + R * k1]) * Q.norm;
      }
    }
    __syncthreads();                                   // every input of this stage is in registers: it becomes the exchange buffer
    strided_fft<R, false>(v, tws, stg, l, t);
    float2* hp = Q.H4 + (long long)t * zs + ky * Q.PH + kx;
    if (OP == ZOP_GRAD) {
      float2 w[R];
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        v[k1] = cscale(v[k1], kap[OP == ZOP_ABS ? 0 : k1]);
        w[k1] = cmul4(v[k1], Q.dpz4[t + R * k1]);
      }
      strided_fft<R, true>(v, tws, X, l, t);           // past its barrier nobody reads the stage any more
      if (tid == 0 && n + 2 < nitems) { fence_proxy_async(); issue(n + 2); }
#pragma unroll
      for (int j = 0; j < R; ++j) hp[j * jstep] = v[j];
      __syncthreads();                                 // X is read out before the second inverse transform reuses it
      strided_fft<R, true>(w, tws, X, l, t);
      float2* hq = hp + Q.HS;
#pragma unroll
      for (int j = 0; j < R; ++j) hq[j * jstep] = w[j];
    } else {
      if (OP == ZOP_DIV) {
        if (comp < 2) {
#pragma unroll
          for (int k1 = 0; k1 < R; ++k1) v[k1] = cscale(v[k1], kap[OP == ZOP_ABS ? 0 : k1]);
        } else if (comp == 2) {
#pragma unroll
          for (int k1 = 0; k1 < R; ++k1) v[k1] = cmul4(cscale(v[k1], kap[OP == ZOP_ABS ? 0 : k1]), Q.dnz4[t + R * k1]);
        } else {
#pragma unroll
          for (int k1 = 0; k1 < R; ++k1) v[k1] = cscale(v[k1], cosk_sel<POLY>(axy + P.az2[t + R * k1]) * Q.norm);
        }
      } else {
        const float e = comp == 0 ? P.y_minus2_half : P.y_minus1_half;
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
          const float k2 = axy + P.kz2[t + R * k1];
          v[k1] = cscale(v[k1], k2 > 0.f ? __powf(k2, e) * Q.norm : 0.f);
        }
      }
      strided_fft<R, true>(v, tws, X, l, t);
      if (tid == 0 && n + 2 < nitems) { fence_proxy_async(); issue(n + 2); }
      float2* op = hp + comp * Q.HS;
#pragma unroll
      for (int j = 0; j < R; ++j) op[j * jstep] = v[j];
    }
  }
}

}  // namespace lifu
