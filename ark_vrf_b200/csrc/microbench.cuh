// Integer-multiply roofline probes and field / group operation microbenchmarks (avrf_microbench).
#pragma once
#include "msm.cuh"

namespace avrf {
// ---- microbenchmarks (integer-multiply roofline probe) -----------------------------------
__global__ void __launch_bounds__(256) k_mb_imad(uint64_t* out, uint32_t iters, uint32_t seed) {
  // 8 independent IMAD.WIDE.U32 accumulation chains per thread
  uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 12345u;
  uint64_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = i + seed;
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + i), "r"(b + u));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// carry-chained form: mad.lo.cc / madc.hi.cc rows as used by mont_mul (IMAD.WIDE.U32[.X] with
// carry predicates); 4 independent rows of 4 chained wide MACs per thread.
__global__ void __launch_bounds__(256) k_mb_imadx(uint32_t* out, uint32_t iters, uint32_t seed) {
  uint32_t a[8], acc[4][8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 2654435761u + seed + i * 977u;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int i = 0; i < 8; i++) acc[r][i] = seed + r * 8 + i;
  uint32_t b = blockIdx.x * 40503u + 12345u;
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
#pragma unroll
      for (int r = 0; r < 4; r++) mad_row(acc[r], a, b + r + u);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[r][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// wide MAC with carry-OUT only (consumed by an ALU addc) / carry-IN only (produced by an ALU add.cc):
// which half of the carry plumbing makes IMAD.WIDE.U32.X issue at 4 cycles?
template <int VARIANT>
__global__ void __launch_bounds__(256) k_mb_imadc(uint32_t* out, uint32_t iters, uint32_t seed) {
  uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 12345u;
  uint32_t lo[8], hi[8], sink = seed, t = seed * 3u;
#pragma unroll
  for (int i = 0; i < 8; i++) { lo[i] = i + seed; hi[i] = i * 7u + seed; }
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (VARIANT == 0) {          // carry-out only
          lo[i] = mad_lo_cc(a + i, b + u, lo[i]);
          hi[i] = madc_hi_cc(a + i, b + u, hi[i]);
          sink = addc(sink, 0);
        } else {                     // carry-in only
          t = add_cc(t, a);
          lo[i] = madc_lo_cc(a + i, b + u, lo[i]);
          hi[i] = madc_hi(a + i, b + u, hi[i]);
        }
      }
    }
  }
  uint32_t s = sink ^ t;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= lo[i] ^ hi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain 32-bit IMAD (lo) streams
__global__ void __launch_bounds__(256) k_mb_imad32(uint32_t* out, uint32_t iters, uint32_t seed) {
  uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 12345u;
  uint32_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = i + seed;
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a + i), "r"(b + u));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(128, 4) k_mb_mul(Fe* out, uint32_t iters) {
  Fe a, b;
  for (int i = 0; i < 8; i++) { a.v[i] = threadIdx.x * 77u + i; b.v[i] = blockIdx.x * 13u + 5u * i + 1u; }
  a.v[7] &= 0x0fffffffu;
  b.v[7] &= 0x0fffffffu;
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
    mont_mul<FQ_BAND>(a, a, b);
    mont_mul<FQ_BAND>(b, b, a);
  }
  fe_add<FQ_BAND>(a, a, b);
  store_fe(out + blockIdx.x * blockDim.x + threadIdx.x, a);
}

__global__ void __launch_bounds__(128, 4) k_mb_madd(Ext* out, const BaseRec* pts, uint32_t npts, uint32_t iters) {
  Ext acc;
  ext_identity<SUITE_BAND>(acc);
  uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
    AffineK q;
    load_affinek(q, pts + (idx % npts));
    idx = idx * 1664525u + 1013904223u;
    ext_madd<SUITE_BAND>(acc, q.x, q.y, q.k);
  }
  store_ext(out + blockIdx.x * blockDim.x + threadIdx.x, acc);
}

}  // namespace avrf
