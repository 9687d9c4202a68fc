// tcgen05 / TMEM / TMA / mbarrier primitives shared by the tensor-core kernels (sm_100a inline PTX).
//
// Numerics: every GEMM runs as "bf16x6": both operands are split into three bf16 terms
// (v ~ v0 + v1 + v2, round-to-nearest at each level, 24 mantissa bits in total) and the six products
// a0w0, a0w1, a1w0, a0w2, a1w1, a2w0 are accumulated in fp32 in TMEM, smallest first.  Measured on B200
// (scripts/tc_probe.cu): max error 6.9e-7 on outputs of magnitude 5.7, below a plain fp32 FMA loop.
//
// Operand conventions (M = 128 rows = TMEM lanes, thread r <-> row r):
//   A : TMEM, K-major, 2 bf16 per 32-bit column (element k in the low half of column k/2 when k even),
//       split s of a 64-wide operand at columns a + 32*s.  Written with tcgen05.st by the row's threads.
//   B : shared memory, UMMA canonical no-swizzle layout made of 8x8 "core matrices" (128 contiguous
//       bytes); descriptor = start address, LBO (stride between core matrices along K), SBO (along N).
//   D : TMEM fp32, lane = row, column = n.  Read back with tcgen05.ld (32x32b: one row per thread).
#pragma once
#include "common.cuh"

namespace eqd {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned addr = smem_u32(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src));
}
// barrier among the 256 threads of one tile group (named barriers 1, 2)
__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); }
// barrier between the two warps that own the two column halves of the same 32 rows (named barriers 3..10)
// group barrier + AND-reduction of a predicate over its 256 threads
__device__ __forceinline__ bool wg_barrier_and(int wg, bool pred) {
  unsigned r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, %2, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"((unsigned)pred), "r"(wg + 1)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// No-swizzle canonical B descriptor (version bits for sm_100): lbo / sbo in bytes
__device__ __forceinline__ unsigned long long b_desc_ex(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (unsigned long long)((saddr >> 4) & 0x3FFF) | ((unsigned long long)(lbo >> 4) << 16) |
         ((unsigned long long)(sbo >> 4) << 32) | (1ull << 46);
}
// K-major weight panels [N=64][K]: LBO (K direction) = 1024 B, SBO (N direction) = 128 B
__device__ __forceinline__ unsigned long long b_desc(unsigned saddr) { return b_desc_ex(saddr, 1024, 128); }
// D[128x64] (+)= A[tmem, 128x16 bf16] * B[smem desc, 64x16 bf16]^T
// instruction descriptor: D = f32, A = B = bf16, A K-major, M = 128; b_mn_major selects a [K][N] B operand
__host__ __device__ constexpr unsigned umma_idesc(unsigned n, unsigned b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_ts_i(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned idesc,
                                          unsigned accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc),
      "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned accum) {
  const unsigned idesc = umma_idesc(64, 0);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc),
      "r"(accum) : "memory");
}
// All 6 cross products of the bf16x3 splits, smallest terms first; A split s at a_base + s*a_split_cols,
// weight split s at w_saddr + s*w_split_bytes (K-major panel, 2048 B per k-block of 16).
// accum0 = 0 starts a fresh accumulator, 1 adds to what D already holds.
__device__ __forceinline__ void issue_gemm(unsigned d_tmem, unsigned a_base, unsigned a_split_cols, unsigned w_saddr,
                                           unsigned w_split_bytes, int kblocks, unsigned accum0 = 0) {
  const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
  unsigned accum = accum0;
#pragma unroll
  for (int pr = 0; pr < 6; ++pr)
    for (int kb = 0; kb < kblocks; ++kb) {
      umma_ts(d_tmem, a_base + pa[pr] * a_split_cols + kb * 8, b_desc(w_saddr + pb[pr] * w_split_bytes + kb * 2048), accum);
      accum = 1;
    }
}
// Same for a K-major [N][K] panel of any N (multiple of 16): a k-block of 16 is 32 N bytes, LBO = 16 N, SBO = 128.
template <int N>
__device__ __forceinline__ void issue_gemm_n(unsigned d_tmem, unsigned a_base, unsigned a_split_cols, unsigned w_saddr,
                                             unsigned w_split_bytes, int kblocks, unsigned accum0 = 0) {
  const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
  const unsigned idesc = umma_idesc(N, 0);
  unsigned accum = accum0;
#pragma unroll
  for (int pr = 0; pr < 6; ++pr)
    for (int kb = 0; kb < kblocks; ++kb) {
      umma_ts_i(d_tmem, a_base + pa[pr] * a_split_cols + kb * 8,
                b_desc_ex(w_saddr + pb[pr] * w_split_bytes + kb * (N * 32), N * 16, 128), idesc, accum);
      accum = 1;
    }
}
// End of a kernel that owns tensor memory: every thread has passed its last tcgen05 operation (callers fence + sync
// first); warp 0, which allocated the 512 columns, releases them.
__device__ __forceinline__ void tmem_release(unsigned tmem_base, int warp) {
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bf16x3 split of register tiles and TMEM stores / loads -------------------------------------
__device__ __forceinline__ unsigned cvt_bf16x2(float hi, float lo) {
  unsigned d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// (v0, v1) -> three packed bf16x2 words (v0 in the low half = even k)
__device__ __forceinline__ void split3_pair(float v0, float v1, unsigned& p0, unsigned& p1, unsigned& p2) {
  p0 = cvt_bf16x2(v1, v0);
  // v - hi as fma(hi, -1, v) on both lanes at once (fma.f32x2; exact product, same rounding as the subtraction)
  const float2 m1 = make_float2(-1.f, -1.f);
  float2 rr = __ffma2_rn(make_float2(__uint_as_float(p0 << 16), __uint_as_float(p0 & 0xFFFF0000u)), m1, make_float2(v0, v1));
  p1 = cvt_bf16x2(rr.y, rr.x);
  rr = __ffma2_rn(make_float2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xFFFF0000u)), m1, rr);
  p2 = cvt_bf16x2(rr.y, rr.x);
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
               "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(unsigned taddr, unsigned (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// 32 fp32 values (this thread's half of its row) -> bf16x3 -> TMEM: split s lands at a_taddr + 32*s, 16 columns
// (split_cols: columns between consecutive splits of the A operand, 32 for a 64-wide K, 40 for K = 80)
__device__ __forceinline__ void store_half_split3(unsigned a_taddr, const float (&v)[32], unsigned split_cols = 32) {
  unsigned p0[16], p1[16], p2[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) split3_pair(v[2 * c], v[2 * c + 1], p0[c], p1[c], p2[c]);
  tmem_st16(a_taddr, p0);
  tmem_st16(a_taddr + split_cols, p1);
  tmem_st16(a_taddr + 2 * split_cols, p2);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 8 fp32 values + 8 zeros (one K block of 16) -> bf16x3 -> 8 TMEM columns per split
__device__ __forceinline__ void store_extra8_split3(unsigned a_taddr, const float (&t)[8], unsigned split_cols) {
  unsigned p0[8], p1[8], p2[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) split3_pair(t[2 * c], t[2 * c + 1], p0[c], p1[c], p2[c]);
#pragma unroll
  for (int c = 4; c < 8; ++c) p0[c] = p1[c] = p2[c] = 0u;
  tmem_st8(a_taddr, p0);
  tmem_st8(a_taddr + split_cols, p1);
  tmem_st8(a_taddr + 2 * split_cols, p2);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16f(unsigned taddr, float (&v)[16]) {
  unsigned a[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
        "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]);
}
__device__ __forceinline__ void tmem_st4(unsigned taddr, const unsigned* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld32f(unsigned taddr, float (&v)[32]) {
  unsigned a[32];
  tmem_ld32_nowait(taddr, a);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]);
}


}  // namespace eqd
