// GPU Pippenger for the batch equation of thin::BatchVerifier::verify (src/thin.rs:282-319).
//
// Replaces `ark_ec::VariableBaseMSM::msm_unchecked` (ark-ec 0.6; call site src/thin.rs:319)
// and the scalar assembly loop at src/thin.rs:287-313.  Any correct MSM returns the same
// group element, and only `is_zero()` of it is observable (thin.rs:320-324).
//
// Shape: signed 16-bit windows (16 windows, 2^15 buckets each => 2^19 bins).
//   k_scalars      per proof : w_j, the 2+2M scalars, signed digits, bin histogram, sum w_j s_j
//   k_gscalar      1 thread  : g = -sum w_j s_j, digits of the shared G term
//   k_scan_*       bins      : exclusive scans -> entry offsets, rank among non-empty bins
//   k_scatter      per point : counting-sort scatter of (point, sign) into bin order   [HBM]
//   k_sort_transpose + k_scatter_window : the same for large batches, window by window, bin offsets in shared memory
//   k_accumulate   per L-entry segment of the sorted array: mixed additions, partial sums
//                  flushed per bin (perfect load balance under bucket skew, H3)         [IMAD]
//   k_combine(_big) per bin  : fold a bin's partial sums to one
//   k_bucket_reduce, k_window_sum, k_fold : sum_b b*B_b per window, Horner over windows
#pragma once
#include "thin.cuh"

namespace avrf {

constexpr int MSM_WBITS = 16;
constexpr int MSM_NWIN = 16;
constexpr int MSM_NBUCKET = 1 << (MSM_WBITS - 1);          // 32768 per window
constexpr int MSM_NBINS = MSM_NWIN * MSM_NBUCKET;          // 524288
constexpr int MSM_CHUNK = 16;                              // buckets per reduce thread
constexpr int MSM_NCHUNK = MSM_NBUCKET / MSM_CHUNK;        // 2048 per window

struct Seed64 { uint64_t w[8]; };                          // SHA-512 digest as 8 big-endian words

#ifdef __CUDACC__

__device__ __forceinline__ void load_affinek(AffineK& q, const BaseRec* p) {
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4 a0 = __ldg(s + 0), a1 = __ldg(s + 1), a2 = __ldg(s + 2), a3 = __ldg(s + 3), a4 = __ldg(s + 4), a5 = __ldg(s + 5);
  q.x.v[0] = a0.x; q.x.v[1] = a0.y; q.x.v[2] = a0.z; q.x.v[3] = a0.w;
  q.x.v[4] = a1.x; q.x.v[5] = a1.y; q.x.v[6] = a1.z; q.x.v[7] = a1.w;
  q.y.v[0] = a2.x; q.y.v[1] = a2.y; q.y.v[2] = a2.z; q.y.v[3] = a2.w;
  q.y.v[4] = a3.x; q.y.v[5] = a3.y; q.y.v[6] = a3.z; q.y.v[7] = a3.w;
  q.k.v[0] = a4.x; q.k.v[1] = a4.y; q.k.v[2] = a4.z; q.k.v[3] = a4.w;
  q.k.v[4] = a5.x; q.k.v[5] = a5.y; q.k.v[6] = a5.z; q.k.v[7] = a5.w;
}

__device__ __forceinline__ void store_fe(Fe* dst, const Fe& a) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  d[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

__device__ __forceinline__ void load_fe(Fe& a, const Fe* src) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4 lo = s[0], hi = s[1];
  a.v[0] = lo.x; a.v[1] = lo.y; a.v[2] = lo.z; a.v[3] = lo.w;
  a.v[4] = hi.x; a.v[5] = hi.y; a.v[6] = hi.z; a.v[7] = hi.w;
}

__device__ __forceinline__ void store_ext(Ext* dst, const Ext& p) {
  store_fe(&dst->x, p.x); store_fe(&dst->y, p.y); store_fe(&dst->z, p.z); store_fe(&dst->t, p.t);
}

__device__ __forceinline__ void load_ext(Ext& p, const Ext* src) {
  load_fe(p.x, &src->x); load_fe(p.y, &src->y); load_fe(p.z, &src->z); load_fe(p.t, &src->t);
}

// ---------------------------------------------------------------------------------------
// Digits + histogram
// ---------------------------------------------------------------------------------------
// Digits of one scalar, its histogram contribution, and - from the same (returning) atomic - the rank
// of each entry inside its bin, so that the scatter pass needs no second round of atomics.
__device__ __forceinline__ void emit_scalar(uint4* digits, uint32_t* hist, uint4* ranks, Fe* scalars_tap, size_t point,
                                            const Fe& k) {
  int32_t dg[16];
  recode_signed16(dg, k);
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 8; i++) pk[i] = ((uint32_t)dg[2 * i] & 0xffffu) | ((uint32_t)dg[2 * i + 1] << 16);
  digits[2 * point] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  digits[2 * point + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  uint32_t rk[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int32_t d = dg[i];
    rk[i] = 0;
    if (d != 0) {
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      rk[i] = atomicAdd(&hist[i * MSM_NBUCKET + mag - 1], 1u);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) ranks[4 * point + i] = make_uint4(rk[4 * i], rk[4 * i + 1], rk[4 * i + 2], rk[4 * i + 3]);
  if (scalars_tap) store_fe(&scalars_tap[point], k);
}

struct ScalArgs {
  const uint32_t* cs;       // 16 words per proof: c (4), 0 (4), s canonical (8)
  const uint32_t* z;        // 4 words per I/O pair
  const uint32_t* io_off;   // n+1
  uint4* digits;            // 2 x uint4 per point
  uint4* ranks;             // 4 x uint4 per point: rank of each entry in its bin
  uint32_t* hist;           // MSM_NBINS
  uint32_t* gpart;          // 10 words per block: sum of w_j s_j as a plain integer
  uint32_t* w_tap;          // optional, 4 words per proof
  Fe* scalars_tap;          // optional
  Seed64 seed;
  uint64_t first_index;
  const uint64_t* segs;     // optional (nseg > 1): pairs (local start, global first index), ascending local start
  uint32_t nseg;
  uint32_t n;
};

// Global index (position in the batch transcript, thin.rs:273-289) of this handle's proof j.  A handle normally
// holds one contiguous run [first_index, first_index + n); the shards of a multi-GPU batch that was pushed in
// several calls hold several runs.
__device__ __forceinline__ uint64_t global_index(uint64_t first_index, const uint64_t* segs, uint32_t nseg, uint32_t j) {
  if (nseg <= 1) return first_index + j;
  uint32_t lo = 0, hi = nseg;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (segs[2 * mid] <= j) lo = mid;
    else hi = mid;
  }
  return segs[2 * lo + 1] + (j - segs[2 * lo]);
}

// 10-limb integer add with carry chain
__device__ __forceinline__ void add10(uint32_t* a, const uint32_t* b) {
  a[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 9; i++) a[i] = addc_cc(a[i], b[i]);
  a[9] = addc(a[9], b[9]);
}

// One thread per proof.  (Four consecutive proofs share a 64-byte squeeze block, thin.rs:289 / transcript.rs:255-273, and
// one thread per FOUR proofs saves three of four compressions - measured: k_scalars 0.58 -> 0.56 ms, it is bound by its
// 59 M returning atomics, but k_scatter 0.93 -> 1.35 ms, because the ranks handed out by the atomics no longer follow the
// point order and the scattered stores lose their locality.  Kept at one.)
constexpr uint32_t SCAL_PER_THREAD = 1;
template <int S>
__global__ void __launch_bounds__(128) k_scalars(ScalArgs a) {
  constexpr int FR = SuiteT<S>::FR;
  uint32_t j0 = (blockIdx.x * blockDim.x + threadIdx.x) * SCAL_PER_THREAD;
  uint32_t acc[10];
#pragma unroll
  for (int i = 0; i < 10; i++) acc[i] = 0;
  uint64_t blk[8], have = ~(uint64_t)0;
#pragma unroll 1
  for (uint32_t j = j0; j < j0 + SCAL_PER_THREAD && j < a.n; j++) {
    uint64_t jg = global_index(a.first_index, a.segs, a.nseg, j);
    if ((jg >> 2) != have) {
      have = jg >> 2;
      sha512_xof_block(blk, a.seed.w, have);           // thin.rs:289 via transcript.rs:255-273
    }
    Fe w, c, s, wM, wc, ws;
    fe_zero(w);
    fe_zero(c);
    digest_le128(w.v, blk, 16 * (uint32_t)(jg & 3));
    const uint32_t* csj = a.cs + 16 * (size_t)j;
#pragma unroll
    for (int i = 0; i < 4; i++) c.v[i] = csj[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s.v[i] = csj[8 + i];
    if (a.w_tap) {
#pragma unroll
      for (int i = 0; i < 4; i++) a.w_tap[4 * (size_t)j + i] = w.v[i];
    }
    to_mont<FR>(wM, w);
    mont_mul_c<FR>(wc, wM, c);                           // w*c      (canonical)
    mont_mul_c<FR>(ws, wM, s);                           // w*s      (canonical)
    uint32_t io0 = a.io_off[j], io1 = a.io_off[j + 1];
    size_t pbase = 2 * (size_t)j + 2 * (size_t)io0;
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pbase + 0, w);    // R_j  : w           thin.rs:295-296
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pbase + 1, wc);   // pk_j : w c z0      thin.rs:299-300
    for (uint32_t i = io0; i < io1; i++) {
      Fe z, zM, t;
      fe_zero(z);
#pragma unroll
      for (int q = 0; q < 4; q++) z.v[q] = a.z[4 * (size_t)i + q];
      to_mont<FR>(zM, z);
      mont_mul_c<FR>(t, wc, zM);                         // O_i : w c z_i         thin.rs:307-308
      emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pbase + 2 + 2 * (size_t)(i - io0), t);
      mont_mul_c<FR>(t, ws, zM);
      fe_neg<FR>(t, t);                                // I_i : -(w s z_i)      thin.rs:310-311
      emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pbase + 3 + 2 * (size_t)(i - io0), t);
    }
    uint32_t wsv[10];
#pragma unroll
    for (int i = 0; i < 8; i++) wsv[i] = ws.v[i];      // g -= w s z0           thin.rs:303
    wsv[8] = wsv[9] = 0;
    add10(acc, wsv);
  }
  // block sum of w_j s_j (plain 320-bit integers)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    uint32_t o[10];
#pragma unroll
    for (int i = 0; i < 10; i++) o[i] = __shfl_down_sync(0xffffffffu, acc[i], off);
    add10(acc, o);
  }
  __shared__ uint32_t sm[4][10];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 10; i++) sm[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < 4; wv++) add10(acc, sm[wv]);
#pragma unroll
    for (int i = 0; i < 10; i++) a.gpart[10 * (size_t)blockIdx.x + i] = acc[i];
  }
}

// g = -(sum of block partials) mod r; emits the digits of the shared generator term
// (thin.rs:315-317) as the last MSM point.  One block of 256 threads.
template <int S>
__global__ void __launch_bounds__(256) k_gscalar(const uint32_t* gpart, uint32_t nblocks, uint4* digits, uint4* ranks,
                                                 uint32_t* hist, Fe* scalars_tap, BaseRec* pts, size_t gpoint) {
  constexpr int FR = SuiteT<S>::FR;
  __shared__ uint32_t sm[8][10];
  uint32_t acc[10];
#pragma unroll
  for (int i = 0; i < 10; i++) acc[i] = 0;
  for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) add10(acc, gpart + 10 * (size_t)b);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    uint32_t o[10];
#pragma unroll
    for (int i = 0; i < 10; i++) o[i] = __shfl_down_sync(0xffffffffu, acc[i], off);
    add10(acc, o);
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 10; i++) sm[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int wv = 1; wv < 8; wv++) add10(acc, sm[wv]);
  Fe lo, hi, t, g;
  fe_zero(hi);
  for (int i = 0; i < 8; i++) lo.v[i] = acc[i];
  hi.v[0] = acc[8];
  hi.v[1] = acc[9];
  reduce_once<FR>(lo, lo);
  to_mont<FR>(t, hi);                                  // hi * 2^256 mod r (canonical)
  fe_add<FR>(g, lo, t);
  fe_neg<FR>(g, g);
  emit_scalar(digits, hist, ranks, scalars_tap, gpoint, g);
  Affine Ga;
  AffineK G;
  fe_set(Ga.x, AVRF_CC(S).gx);
  fe_set(Ga.y, AVRF_CC(S).gy);
  affine_to_k<S>(G, Ga);
  store_fe(&pts[gpoint].x, G.x);
  store_fe(&pts[gpoint].y, G.y);
  store_fe(&pts[gpoint].k, G.k);
}

// ---------------------------------------------------------------------------------------
// Pedersen VRF batch equation (reference src/pedersen.rs:341-426): 5 bases per proof
// (O_m, Ok, I_m, Yb, R) with scalars (t c, t, -t s, u c, u) and two shared bases G, B with
// -sum u s, -sum u sb.  t_i, u_i = the two 16-byte halves of proof i's 32-byte squeeze.
// ---------------------------------------------------------------------------------------
struct PedScalArgs {
  const uint32_t* cs;       // 24 words per proof: c (4) 0 (4) s (8) sb (8)
  uint4* digits;
  uint4* ranks;
  uint32_t* hist;
  uint32_t* gpart;          // 20 words per block: sum u s, sum u sb
  uint32_t* w_tap;          // optional: 8 words per proof (t, u)
  Fe* scalars_tap;
  Seed64 seed;
  uint64_t first_index;
  const uint64_t* segs;
  uint32_t nseg;
  uint32_t n;
};

template <int S>
__global__ void __launch_bounds__(128) k_scalars_ped(PedScalArgs a) {
  constexpr int FR = SuiteT<S>::FR;
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t accg[10], accb[10];
#pragma unroll
  for (int i = 0; i < 10; i++) accg[i] = accb[i] = 0;
  if (j < a.n) {
    uint64_t jg = global_index(a.first_index, a.segs, a.nseg, j);
    uint64_t blk[8];
    sha512_xof_block(blk, a.seed.w, jg >> 1);          // pedersen.rs:373-381: 32 bytes per proof
    Fe t, u, c, s, sb, tM, uM, x;
    fe_zero(t); fe_zero(u); fe_zero(c);
    digest_le128(t.v, blk, 32 * (uint32_t)(jg & 1));
    digest_le128(u.v, blk, 32 * (uint32_t)(jg & 1) + 16);
    const uint32_t* csj = a.cs + 24 * (size_t)j;
#pragma unroll
    for (int i = 0; i < 4; i++) c.v[i] = csj[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { s.v[i] = csj[8 + i]; sb.v[i] = csj[16 + i]; }
    if (a.w_tap) {
#pragma unroll
      for (int i = 0; i < 4; i++) { a.w_tap[8 * (size_t)j + i] = t.v[i]; a.w_tap[8 * (size_t)j + 4 + i] = u.v[i]; }
    }
    to_mont<FR>(tM, t);
    to_mont<FR>(uM, u);
    size_t pb = 5 * (size_t)j;
    mont_mul_c<FR>(x, tM, c);
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pb + 0, x);       // O_m : t c        pedersen.rs:391-392
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pb + 1, t);       // Ok  : t          :394-395
    mont_mul_c<FR>(x, tM, s);
    fe_neg<FR>(x, x);
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pb + 2, x);       // I_m : -t s       :397-398
    mont_mul_c<FR>(x, uM, c);
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pb + 3, x);       // Yb  : u c        :401-402
    emit_scalar(a.digits, a.hist, a.ranks, a.scalars_tap, pb + 4, u);       // R   : u          :404-405
    mont_mul_c<FR>(x, uM, s);
#pragma unroll
    for (int i = 0; i < 8; i++) accg[i] = x.v[i];                  // g += u s         :408
    mont_mul_c<FR>(x, uM, sb);
#pragma unroll
    for (int i = 0; i < 8; i++) accb[i] = x.v[i];                  // b += u sb        :409
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    uint32_t o[10];
#pragma unroll
    for (int i = 0; i < 10; i++) o[i] = __shfl_down_sync(0xffffffffu, accg[i], off);
    add10(accg, o);
#pragma unroll
    for (int i = 0; i < 10; i++) o[i] = __shfl_down_sync(0xffffffffu, accb[i], off);
    add10(accb, o);
  }
  __shared__ uint32_t sm[4][20];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 10; i++) { sm[warp][i] = accg[i]; sm[warp][10 + i] = accb[i]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < 4; wv++) { add10(accg, sm[wv]); add10(accb, sm[wv] + 10); }
#pragma unroll
    for (int i = 0; i < 10; i++) {
      a.gpart[20 * (size_t)blockIdx.x + i] = accg[i];
      a.gpart[20 * (size_t)blockIdx.x + 10 + i] = accb[i];
    }
  }
}

// -(sum of block partials) mod r for the two shared bases G and B (pedersen.rs:412-417); one warp.
template <int S>
__global__ void __launch_bounds__(32) k_gscalar_ped(const uint32_t* gpart, uint32_t nblocks, uint4* digits, uint4* ranks,
                                                    uint32_t* hist, Fe* scalars_tap, BaseRec* pts, size_t gpoint) {
  constexpr int FR = SuiteT<S>::FR;
  uint32_t acc[2][10];
  for (int w = 0; w < 2; w++)
    for (int i = 0; i < 10; i++) acc[w][i] = 0;
  for (uint32_t b = threadIdx.x; b < nblocks; b += 32) {
    add10(acc[0], gpart + 20 * (size_t)b);
    add10(acc[1], gpart + 20 * (size_t)b + 10);
  }
  for (int w = 0; w < 2; w++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      uint32_t o[10];
#pragma unroll
      for (int i = 0; i < 10; i++) o[i] = __shfl_down_sync(0xffffffffu, acc[w][i], off);
      add10(acc[w], o);
    }
  }
  if (threadIdx.x != 0) return;
  for (int w = 0; w < 2; w++) {
    Fe lo, hi, t, g;
    fe_zero(hi);
    for (int i = 0; i < 8; i++) lo.v[i] = acc[w][i];
    hi.v[0] = acc[w][8];
    hi.v[1] = acc[w][9];
    reduce_once<FR>(lo, lo);
    to_mont<FR>(t, hi);
    fe_add<FR>(g, lo, t);
    fe_neg<FR>(g, g);
    emit_scalar(digits, hist, ranks, scalars_tap, gpoint + w, g);
    Affine Pa;
    AffineK P;
    fe_set(Pa.x, w == 0 ? AVRF_CC(S).gx : AVRF_CC(S).bx);
    fe_set(Pa.y, w == 0 ? AVRF_CC(S).gy : AVRF_CC(S).by);
    affine_to_k<S>(P, Pa);
    store_fe(&pts[gpoint + w].x, P.x);
    store_fe(&pts[gpoint + w].y, P.y);
    store_fe(&pts[gpoint + w].k, P.k);
  }
}

// ---------------------------------------------------------------------------------------
// Scans over the 2^19 bins: entry offsets (offs, NBINS+1 entries) and the rank of every bin among
// the non-empty ones (nzr).
// ---------------------------------------------------------------------------------------

// 512 blocks x 1024 threads: per-block exclusive scan, block totals out.
__global__ void __launch_bounds__(1024) k_scan_local(const uint32_t* hist, uint32_t* offs, uint32_t* nzr,
                                                     uint32_t* btot) {
  __shared__ uint32_t se[1024], st[1024];
  uint32_t tid = threadIdx.x, b = blockIdx.x * 1024 + tid;
  uint32_t c = hist[b], t = c != 0u ? 1u : 0u;
  se[tid] = c;
  st[tid] = t;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    uint32_t ve = 0, vt = 0;
    if (tid >= d) { ve = se[tid - d]; vt = st[tid - d]; }
    __syncthreads();
    se[tid] += ve;
    st[tid] += vt;
    __syncthreads();
  }
  offs[b] = se[tid] - c;
  nzr[b] = st[tid] - t;
  if (tid == 1023) { btot[2 * blockIdx.x] = se[tid]; btot[2 * blockIdx.x + 1] = st[tid]; }
}

// one block of 512 threads: exclusive scan of block totals; totals[0]=entries, totals[1]=non-empty bins
__global__ void __launch_bounds__(512) k_scan_totals(uint32_t* btot, uint32_t* totals, uint32_t* offs) {
  __shared__ uint32_t se[512], st[512];
  uint32_t tid = threadIdx.x;
  uint32_t c = btot[2 * tid], t = btot[2 * tid + 1];
  se[tid] = c;
  st[tid] = t;
  __syncthreads();
  for (uint32_t d = 1; d < 512; d <<= 1) {
    uint32_t ve = 0, vt = 0;
    if (tid >= d) { ve = se[tid - d]; vt = st[tid - d]; }
    __syncthreads();
    se[tid] += ve;
    st[tid] += vt;
    __syncthreads();
  }
  btot[2 * tid] = se[tid] - c;
  btot[2 * tid + 1] = st[tid] - t;
  if (tid == 511) { totals[0] = se[tid]; totals[1] = st[tid]; totals[2] = 0; offs[MSM_NBINS] = se[tid]; }
}

__global__ void __launch_bounds__(1024) k_scan_add(uint32_t* offs, uint32_t* nzr, const uint32_t* btot) {
  uint32_t b = blockIdx.x * 1024 + threadIdx.x;
  offs[b] += btot[2 * blockIdx.x];
  nzr[b] += btot[2 * blockIdx.x + 1];
}

// ---------------------------------------------------------------------------------------
// Scatter (counting sort, second pass)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(const uint4* __restrict__ digits, const uint4* __restrict__ ranks,
                                                 const uint32_t* __restrict__ offs, uint32_t* __restrict__ entries,
                                                 size_t npoints) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npoints) return;
  uint4 d0 = digits[2 * p], d1 = digits[2 * p + 1];
  uint32_t pk[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
  uint32_t rk[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 r = ranks[4 * p + i];
    rk[4 * i] = r.x; rk[4 * i + 1] = r.y; rk[4 * i + 2] = r.z; rk[4 * i + 3] = r.w;
  }
  // position = bin offset + the rank the histogram atomic handed out: all 16 offset loads are issued
  // before the scattered 4-byte stores
  uint32_t base[16], val[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int32_t d = (int32_t)(int16_t)((pk[i >> 1] >> (16 * (i & 1))) & 0xffffu);
    uint32_t neg = d < 0;
    uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
    val[i] = ((uint32_t)p << 1) | neg;
    base[i] = d != 0 ? __ldg(offs + i * MSM_NBUCKET + mag - 1) : 0xffffffffu;
  }
#pragma unroll
  for (int i = 0; i < 16; i++)
    if (base[i] != 0xffffffffu) entries[base[i] + rk[i]] = val[i];
}

// ---------------------------------------------------------------------------------------
// Scatter, window by window.  k_scatter above issues, per point, 16 scattered loads of bin offsets and 16 scattered
// 4-byte stores: it is bound by L1 wavefronts (one per distinct line per warp instruction), half of them the offset
// loads.  Here the digits and ranks are first transposed to window-major arrays (coalesced both ways through shared
// memory), then every block works on ONE window with that window's 32 768 bin offsets in shared memory (128 KiB):
// its reads are coalesced, the offset look-ups are shared-memory reads, and only the entry stores stay scattered -
// into a 15 MB slice of the entry array that stays resident in L2 while the window is being written.
// ---------------------------------------------------------------------------------------
constexpr int SORT_TP = 256;        // points per transpose block

__global__ void __launch_bounds__(SORT_TP) k_sort_transpose(const uint4* __restrict__ digits, const uint4* __restrict__ ranks,
                                                            size_t npoints, size_t stride, uint16_t* __restrict__ dig_w,
                                                            uint32_t* __restrict__ rank_w) {
  __shared__ uint16_t sd[MSM_NWIN][SORT_TP + 2];
  __shared__ uint32_t sr[MSM_NWIN][SORT_TP + 1];
  size_t p0 = (size_t)blockIdx.x * SORT_TP, p = p0 + threadIdx.x;
  uint32_t t = threadIdx.x;
  if (p < npoints) {
    uint4 d0 = digits[2 * p], d1 = digits[2 * p + 1];
    uint32_t pk[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
    for (int i = 0; i < 8; i++) { sd[2 * i][t] = (uint16_t)(pk[i] & 0xffffu); sd[2 * i + 1][t] = (uint16_t)(pk[i] >> 16); }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint4 r = ranks[4 * p + i];
      sr[4 * i][t] = r.x; sr[4 * i + 1][t] = r.y; sr[4 * i + 2][t] = r.z; sr[4 * i + 3][t] = r.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < MSM_NWIN; i++) { sd[i][t] = 0; sr[i][t] = 0; }
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < MSM_NWIN; w++) {
    dig_w[(size_t)w * stride + p] = sd[w][t];            // stride is a multiple of SORT_TP: always in range
    rank_w[(size_t)w * stride + p] = sr[w][t];
  }
}

__global__ void __launch_bounds__(1024) k_scatter_window(const uint16_t* __restrict__ dig_w, const uint32_t* __restrict__ rank_w,
                                                         size_t npoints, size_t stride, const uint32_t* __restrict__ offs,
                                                         uint32_t* __restrict__ entries) {
  extern __shared__ uint32_t s_offs[];                   // MSM_NBUCKET offsets of this block's window
  const uint32_t w = blockIdx.y;
  for (uint32_t i = threadIdx.x; i < MSM_NBUCKET; i += blockDim.x) s_offs[i] = offs[w * MSM_NBUCKET + i];
  __syncthreads();
  // four consecutive points per thread and iteration: one 8-byte and one 16-byte coalesced load, four scattered stores;
  // the loads of two iterations are in flight together
  const size_t per = ((npoints + gridDim.x - 1) / gridDim.x + 4095) & ~(size_t)4095;
  const size_t p0 = (size_t)blockIdx.x * per, p1 = p0 + per < npoints ? p0 + per : npoints;
  const uint2* dg = reinterpret_cast<const uint2*>(dig_w + (size_t)w * stride);
  const uint4* rk = reinterpret_cast<const uint4*>(rank_w + (size_t)w * stride);
  auto put = [&](size_t p, uint32_t d16, uint32_t rank) {
    int32_t d = (int32_t)(int16_t)d16;
    if (d == 0 || p >= p1) return;
    uint32_t neg = d < 0;
    uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
    entries[s_offs[mag - 1] + rank] = ((uint32_t)p << 1) | neg;
  };
#pragma unroll 2
  for (size_t p = p0 + 4 * (size_t)threadIdx.x; p < p1; p += 4 * (size_t)blockDim.x) {
    uint2 d = __ldg(dg + (p >> 2));
    uint4 r = __ldg(rk + (p >> 2));
    put(p, d.x & 0xffffu, r.x);
    put(p + 1, d.x >> 16, r.y);
    put(p + 2, d.y & 0xffffu, r.z);
    put(p + 3, d.y >> 16, r.w);
  }
}

// ---------------------------------------------------------------------------------------
// Bucket accumulation: the IMAD-bound kernel.  The sorted entry array is cut into equal segments, one per thread of a
// grid that is an exact multiple of the resident blocks (4 waves of SMs x blocks per SM: q = ceil(E / threads)
// consecutive entries each), so every lane of every warp performs the same number of unified mixed additions - no
// bucket-size imbalance, no partly filled last wave, and still enough blocks for the hardware scheduler to
// even out the SMs (one single wave is slower: the slowest SM sets the time).  When a segment crosses into the next bin the running
// sum is flushed to a slot; the partial sums of bin b lie in the contiguous slots
// offs[b] / q + nzr[b] ... (offs[b] + cnt - 1) / q + nzr[b].
// ---------------------------------------------------------------------------------------
struct AccArgs {
  const uint32_t* entries;
  const uint32_t* offs;     // NBINS + 1
  const uint32_t* hist;
  const uint32_t* nzr;
  const uint32_t* totals;   // [0] = number of entries
  const BaseRec* pts;
  Ext* slots;
  uint32_t nthr;            // threads of the accumulation grid
};

// Entries per accumulation thread for E sorted entries (every kernel of the tail derives it the same way).
__device__ __forceinline__ uint32_t seg_len(uint32_t E, uint32_t nthr) {
  uint32_t q = (E + nthr - 1) / nthr;
  return q < 8u ? 8u : q;
}

__device__ __forceinline__ uint32_t first_slot(const uint32_t* offs, const uint32_t* nzr, uint32_t bin, uint32_t q) {
  return offs[bin] / q + nzr[bin];
}

// The entry indices are read two iterations ahead and the next base is prefetched into L2 while the
// current addition runs.  (Staging the next base in shared memory with cp.async, double-buffered, was
// measured too: 7.66 ms vs 7.67 ms - the gathers are already hidden, the kernel is IMAD-pipe bound.)
template <int S, int LB>
__global__ void __launch_bounds__(128, LB) k_accumulate(AccArgs a) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t E = a.totals[0];
  uint32_t q = seg_len(E, a.nthr);
  uint64_t e64 = (uint64_t)t * q;
  if (e64 >= E) return;
  uint32_t e = (uint32_t)e64;
  uint32_t end = (E - e) > q ? e + q : E;
  // bin of the first entry: the last bin with offs[bin] <= e (it is non-empty)
  uint32_t lo = 0, hi = MSM_NBINS;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(a.offs + mid) <= e) lo = mid;
    else hi = mid;
  }
  uint32_t bin = lo;
  uint32_t bend = __ldg(a.offs + bin) + __ldg(a.hist + bin);
  Ext acc;
  ext_identity<S>(acc);
  uint32_t v0 = __ldg(a.entries + e);
  uint32_t v1 = e + 1 < end ? __ldg(a.entries + e + 1) : 0;
#pragma unroll 1
  for (; e < end; e++) {
    if (e == bend) {                                   // segment crosses into the next non-empty bin
      store_ext(a.slots + t + __ldg(a.nzr + bin), acc);
      ext_identity<S>(acc);
      do { bin++; } while (__ldg(a.hist + bin) == 0);
      bend = __ldg(a.offs + bin) + __ldg(a.hist + bin);
    }
    uint32_t v = v0;
    AffineK q;
    if (e + 1 < end) {
      const char* nb = reinterpret_cast<const char*>(a.pts + (v1 >> 1));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nb));          // one 128-byte line per record
    }
    uint32_t v2 = e + 2 < end ? __ldg(a.entries + e + 2) : 0;
    load_affinek(q, a.pts + (v >> 1));
    v0 = v1;
    v1 = v2;
    base_cneg<S>(q, v & 1);
    ext_madd<S>(acc, q.x, q.y, q.k);
  }
  store_ext(a.slots + t + __ldg(a.nzr + bin), acc);
}

__device__ __forceinline__ void shfl_down_ext(Ext& o, const Ext& p, int off) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    o.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], off);
    o.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], off);
    o.z.v[i] = __shfl_down_sync(0xffffffffu, p.z.v[i], off);
    o.t.v[i] = __shfl_down_sync(0xffffffffu, p.t.v[i], off);
  }
}

constexpr uint32_t COMBINE_SERIAL_MAX = 8;

// One thread per bin: fold the (few) partial sums of a bin into its first slot.  Bins with more
// than COMBINE_SERIAL_MAX partials (the weight-carry bucket, tiny-batch skew) go to k_combine_big.
template <int S>
__global__ void __launch_bounds__(128) k_combine(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ offs,
                                                 const uint32_t* __restrict__ nzr, Ext* __restrict__ slots,
                                                 uint32_t nthr, uint32_t* __restrict__ totals, uint32_t* __restrict__ biglist) {
  uint32_t bin = blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= MSM_NBINS) return;
  uint32_t c = hist[bin];
  if (c == 0) return;
  uint32_t o = offs[bin], r = nzr[bin], q = seg_len(totals[0], nthr);
  uint32_t s0 = o / q + r, s1 = (o + c - 1) / q + r;
  uint32_t n = s1 - s0 + 1;
  if (n == 1) return;
  if (n > COMBINE_SERIAL_MAX) {
    biglist[atomicAdd(&totals[2], 1u)] = bin;
    return;
  }
  Ext acc;
  load_ext(acc, slots + s0);
#pragma unroll 1
  for (uint32_t i = 1; i < n; i++) {
    Ext q;
    load_ext(q, slots + s0 + i);
    ext_add_c<S>(acc, acc, q);
  }
  store_ext(slots + s0, acc);
}

// One block (256 threads) per big bin, grid-stride over the list.
template <int S>
__global__ void __launch_bounds__(256) k_combine_big(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ offs,
                                                     const uint32_t* __restrict__ nzr, Ext* __restrict__ slots,
                                                     uint32_t nthr, const uint32_t* __restrict__ totals,
                                                     const uint32_t* __restrict__ biglist) {
  __shared__ Ext sm[256];
  uint32_t nbig = totals[2], tid = threadIdx.x, q = seg_len(totals[0], nthr);
  for (uint32_t i = blockIdx.x; i < nbig; i += gridDim.x) {
    uint32_t bin = biglist[i];
    uint32_t c = hist[bin], o = offs[bin], r = nzr[bin];
    uint32_t s0 = o / q + r, s1 = (o + c - 1) / q + r;
    uint32_t n = s1 - s0 + 1;
    Ext acc;
    ext_identity<S>(acc);
#pragma unroll 1
    for (uint32_t k = tid; k < n; k += 256) {
      Ext pq;
      load_ext(pq, slots + s0 + k);
      ext_add_c<S>(acc, acc, pq);
    }
    sm[tid] = acc;
    __syncthreads();
#pragma unroll 1
    for (uint32_t d = 128; d > 0; d >>= 1) {
      if (tid < d) {
        Ext q = sm[tid + d];
        ext_add_c<S>(acc, acc, q);
        sm[tid] = acc;
      }
      __syncthreads();
    }
    if (tid == 0) store_ext(slots + s0, acc);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Quad-cooperative point arithmetic for the latency-bound tail: the four lanes of a quad
// (lane & 3) each compute ONE of the independent field multiplications of a point operation and
// exchange the products by shuffle, so a doubling costs 2 multiplication latencies instead of 8
// and an addition 3 instead of 10.  Every lane of the warp holds the whole point.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void fe_sel4(Fe& r, int q, const Fe& a0, const Fe& a1, const Fe& a2, const Fe& a3) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = q == 0 ? a0.v[i] : q == 1 ? a1.v[i] : q == 2 ? a2.v[i] : a3.v[i];
}

__device__ __forceinline__ void quad_gather(Fe& a, Fe& b, Fe& c, Fe& d, const Fe& m) {
  int base = (threadIdx.x & 31) & ~3;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.v[i] = __shfl_sync(0xffffffffu, m.v[i], base + 0);
    b.v[i] = __shfl_sync(0xffffffffu, m.v[i], base + 1);
    c.v[i] = __shfl_sync(0xffffffffu, m.v[i], base + 2);
    d.v[i] = __shfl_sync(0xffffffffu, m.v[i], base + 3);
  }
}

template <int S>
__device__ __forceinline__ void quad_dbl(Ext& p) {
  constexpr int FQ = SuiteT<S>::FQ;
  int q = threadIdx.x & 3;
  Fe t0, u, v, m, A, B, C, D, E, F, G, H;
  fe_add<FQ>(t0, p.x, p.y);
  fe_sel4(u, q, p.x, p.y, p.z, t0);
  m = mont_mul_v<FQ>(u, u);
  quad_gather(A, B, C, E, m);                      // X^2, Y^2, Z^2, (X+Y)^2
  fe_dbl<FQ>(C, C);
  a_times<S>(D, A);
  fe_sub<FQ>(E, E, A);
  fe_sub<FQ>(E, E, B);
  fe_add<FQ>(G, D, B);
  fe_sub<FQ>(F, G, C);
  fe_sub<FQ>(H, D, B);
  fe_sel4(u, q, E, G, E, F);
  fe_sel4(v, q, F, H, H, G);
  m = mont_mul_v<FQ>(u, v);
  quad_gather(p.x, p.y, p.t, p.z, m);              // E*F, G*H, E*H, F*G
}

template <int S>
__device__ __forceinline__ void quad_add(Ext& p, const Ext& o) {
  constexpr int FQ = SuiteT<S>::FQ;
  int q = threadIdx.x & 3;
  Fe u, v, m, A, B, C, D, E, F, G, H, t0, t1, dd, x0, x1;
  fe_sel4(u, q, p.x, p.y, p.t, p.z);
  fe_sel4(v, q, o.x, o.y, o.t, o.z);
  m = mont_mul_v<FQ>(u, v);
  quad_gather(A, B, C, D, m);                      // X1X2, Y1Y2, T1T2, Z1Z2
  fe_add<FQ>(t0, p.x, p.y);
  fe_add<FQ>(t1, o.x, o.y);
  fe_set(dd, AVRF_CC(S).d);
  fe_sel4(u, q, C, t0, C, t0);
  fe_sel4(v, q, dd, t1, dd, t1);
  m = mont_mul_v<FQ>(u, v);
  quad_gather(C, E, x0, x1, m);                    // d*T1T2, (X1+Y1)(X2+Y2)
  fe_sub<FQ>(E, E, A);
  fe_sub<FQ>(E, E, B);
  fe_sub<FQ>(F, D, C);
  fe_add<FQ>(G, D, C);
  sub_a_times<S>(H, B, A);
  fe_sel4(u, q, E, G, E, F);
  fe_sel4(v, q, F, H, H, G);
  m = mont_mul_v<FQ>(u, v);
  quad_gather(p.x, p.y, p.t, p.z, m);
}

// One thread per chunk of MSM_CHUNK buckets of one window: sum_b b * B_b over the chunk.
// (A quad-cooperative variant - 4 lanes per chunk, uniform control flow - was measured at 0.86 ms against
// 0.43 ms: with ~1000 warps this kernel is throughput-bound and the select/shuffle overhead of the quad
// operations costs more than the shorter dependency chains save.  Quads pay off only in k_fold.)
template <int S>
__global__ void __launch_bounds__(128) k_bucket_reduce(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ offs,
                                                       const uint32_t* __restrict__ nzr, uint32_t nthr,
                                                       const uint32_t* __restrict__ totals,
                                                       const Ext* __restrict__ sums, Ext* __restrict__ chunk_out) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. NWIN*NCHUNK
  if (g >= MSM_NWIN * MSM_NCHUNK) return;
  uint32_t win = g / MSM_NCHUNK, chunk = g % MSM_NCHUNK;
  uint32_t bin0 = win * MSM_NBUCKET + chunk * MSM_CHUNK;
  const uint32_t seglen = seg_len(totals[0], nthr);
  Ext run, acc;
  ext_identity<S>(run);
  ext_identity<S>(acc);
  bool any = false;
#pragma unroll 1
  for (int i = MSM_CHUNK - 1; i >= 0; i--) {
    uint32_t bin = bin0 + i;
    if (hist[bin] != 0) {
      Ext q;
      load_ext(q, sums + first_slot(offs, nzr, bin, seglen));
      if (any) ext_add_c<S>(run, run, q);
      else run = q;
      any = true;
    }
    if (any) ext_add_c<S>(acc, acc, run);
  }
  if (any && chunk > 0) {
    uint32_t k[8] = {chunk * MSM_CHUNK, 0, 0, 0, 0, 0, 0, 0};
    Ext m;
    ext_scalar_mul<S>(m, run, k, 16);
    ext_add_c<S>(acc, acc, m);
  }
  store_ext(chunk_out + g, acc);
}

// One block (256 threads) per window: W_k = sum of its 2048 chunk sums.
template <int S>
__global__ void __launch_bounds__(256) k_window_sum(const Ext* __restrict__ chunk_out, Ext* __restrict__ wsum) {
  __shared__ Ext sm[256];
  uint32_t win = blockIdx.x, tid = threadIdx.x;
  Ext acc;
  load_ext(acc, chunk_out + win * MSM_NCHUNK + tid * 8);
#pragma unroll 1
  for (int i = 1; i < 8; i++) {
    Ext q;
    load_ext(q, chunk_out + win * MSM_NCHUNK + tid * 8 + i);
    ext_add_c<S>(acc, acc, q);
  }
  sm[tid] = acc;
  __syncthreads();
#pragma unroll 1
  for (uint32_t d = 128; d > 0; d >>= 1) {
    if (tid < d) {
      Ext q = sm[tid + d];
      ext_add_c<S>(acc, acc, q);
      sm[tid] = acc;
    }
    __syncthreads();
  }
  if (tid == 0) store_ext(wsum + win, acc);
}

// What a shard of a multi-GPU batch hands to the combining device: its partial sum, its identity-gate flags
// (thin.rs:266-271) and whether the partial is the identity.  k_fold stores it straight into the peer-mapped
// slot of the combining device when one is given (the exchange of SURVEY.md 8e, fused into the kernel's tail).
struct ShardSlot { Ext partial; int gate; int is_identity; int pad[2]; };

// Horner over windows: partial = sum_k 2^(16k) W_k.  flags[1] = partial is the identity.
// One warp, quad-cooperative (240 serial doublings: the critical path of the tail).
template <int S>
__global__ void __launch_bounds__(32) k_fold(const Ext* __restrict__ wsum, Ext* __restrict__ partial, int* flags,
                                             ShardSlot* remote) {
  Ext acc;
  load_ext(acc, wsum + MSM_NWIN - 1);
#pragma unroll 1
  for (int k = MSM_NWIN - 2; k >= 0; k--) {
#pragma unroll 1
    for (int i = 0; i < MSM_WBITS; i++) quad_dbl<S>(acc);
    Ext q;
    load_ext(q, wsum + k);
    quad_add<S>(acc, q);
  }
  if (threadIdx.x == 0) {
    int id = ext_is_identity<S>(acc) ? 1 : 0;
    store_ext(partial, acc);
    flags[1] = id;
    if (remote) {
      store_ext(&remote->partial, acc);
      remote->gate = flags[0];
      remote->is_identity = id;
      __threadfence_system();
    }
  }
}

// Sum of n partial points (multi-GPU tail, thin.rs:320-324).
template <int S>
__global__ void k_combine_partials(const Ext* __restrict__ parts, uint32_t n, Ext* __restrict__ out, int* flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Ext acc;
  ext_identity<S>(acc);
  for (uint32_t i = 0; i < n; i++) {
    Ext q;
    load_ext(q, parts + i);
    ext_add_c<S>(acc, acc, q);
  }
  store_ext(out, acc);
  flags[1] = ext_is_identity<S>(acc) ? 1 : 0;
}

// The same over the shard slots of a multi-GPU batch: flags[0] = OR of the shards' gate flags, flags[1] = the
// total is the identity.  Empty shards leave the identity in their slot.
template <int S>
__global__ void k_combine_shards(const ShardSlot* __restrict__ slots, uint32_t n, Ext* __restrict__ out, int* flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Ext acc;
  ext_identity<S>(acc);
  int gate = 0;
  for (uint32_t i = 0; i < n; i++) {
    Ext q;
    load_ext(q, &slots[i].partial);
    ext_add_c<S>(acc, acc, q);
    gate |= slots[i].gate;
  }
  store_ext(out, acc);
  flags[0] = gate;
  flags[1] = ext_is_identity<S>(acc) ? 1 : 0;
}

#endif  // __CUDACC__

}  // namespace avrf
