// Segmented cross attention of a 64-wide IEGMN layer (rigid_docking_model.py:46-64, 247-256) on the tensor
// cores (tcgen05, bf16x6):   mu_i = sum_j softmax_j(q_i . k_j) v_j   over the partner protein's nodes j
// (the per-pair block of the reference's dense masked softmax; no 1/sqrt(d)).
//
// A tile = 128 query nodes of one protein; two tile groups of 256 threads per CTA (2 threads per query row).
// K and V of every node arrive as bf16x3 8-node blocks (written by the projection kernel), so a run of
// 8 blocks (64 keys) is TMA-bulk-copied straight into shared memory as a UMMA B operand:
//   S = Q K^T   : A = Q (TMEM, bf16x3), B = K blocks, K-major  (n = key, k = d)
//   O += P V    : A = P (TMEM, bf16x3), B = V blocks, MN-major (k = key, n = d)
// Two passes over the keys (row maxima first, then exp / P.V) instead of an online softmax: the extra S GEMMs
// are cheap on the tensor pipe and O never has to be rescaled in TMEM.
#include "tc_common.cuh"

namespace eqd {

// -DATTN_PROF: thread 0 of CTA 0 accumulates the SM cycles it spends in each phase of the tile loop (scripts/attn_variants.py)
#ifdef ATTN_PROF
__device__ long long g_attn_prof[16];
#define PROF_MARK(k) do { if (tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); prof_acc[k] += t_ - prof_t; prof_t = t_; } } while (0)
#else
#define PROF_MARK(k) do { } while (0)
#endif

#define AT_THREADS 512
#define AT_KEYS 64            // keys per chunk = 8 blocks
#define AT_CHUNK_BYTES 8192   // per split

// X5 = the 69-wide layer 0: the tensor cores handle channels 0..63 exactly as in a 64-wide layer; channels 64..68 of
// Q, K, V (fp32 in x5[n][16] = [K64..67 | V64..67 | K68 V68 | Q64..68 | 0], written by the layer-0 projection) are a
// rank-5 update of the scores and five extra output columns, done with plain FMAs next to the exp().
template <bool X5>
struct AtGroupSmem {
  unsigned char k[2][3][AT_CHUNK_BYTES];  // double-buffered K chunks (3 splits)
  unsigned char v[2][3][AT_CHUNK_BYTES];
  float red[EQD_TM * 2];
  float x5c[X5 ? 2 : 1][X5 ? AT_KEYS * 16 : 4];   // x5 rows of the K chunk in flight (same double buffering as k)
  float red5[X5 ? EQD_TM * 2 * 5 : 4];
};
template <bool X5>
struct AtSmem {
  AtGroupSmem<X5> grp[2];
  unsigned long long k_bar[2][2], v_bar[2][2], mma_bar[2];
  unsigned int tmem_base;
};

template <bool X5>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(eqd_graph g, const float* __restrict__ proj, int pw, const unsigned char* __restrict__ kv,
                    long kv_split_stride, const float* __restrict__ x5, float* __restrict__ mu, int ldmu) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AtSmem<X5>& S = *reinterpret_cast<AtSmem<X5>*>(smem_raw);
  const int tid = threadIdx.x, wg = tid >> 8, q = tid & 255, half = q >> 7, r = q & 127, warp = tid >> 5;
  AtGroupSmem<X5>& G = S.grp[wg];
  TRACE_START(1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int a = 0; a < 2; ++a) {
      mbar_init(&S.mma_bar[a], 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&S.k_bar[a][b], 1);
        mbar_init(&S.v_bar[a][b], 1);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, 0, 14);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, 0, 13);
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wg_u = warp_u >> 3;
  const bool issuer_warp = (warp_u & 7) == 0;
  const unsigned tmem_wg = __shfl_sync(0xffffffffu, S.tmem_base, 0) + (unsigned)wg_u * 256;
  const unsigned tmem = tmem_wg + ((unsigned)((warp & 3) * 32) << 16);
  // columns: Q (A) 0..95 | S (fp32, 64) / P (A, 3x32) 96..191 | O 192..255
  const unsigned k_saddr = smem_u32(S.grp[wg_u].k), v_saddr = smem_u32(S.grp[wg_u].v);
  unsigned kph[2] = {0, 0}, vph[2] = {0, 0}, mph = 0;
  const int B = g.n_pairs;
  const unsigned char* k_g = kv;                              // which = 0
  const unsigned char* v_g = kv + 3 * kv_split_stride;        // which = 1

  // a chunk of K or V blocks; x5buf >= 0: also the fp32 x5 rows of those 64 keys (they travel with the K chunks)
  auto load_chunk = [&](const unsigned char* src, unsigned char (*dst)[AT_CHUNK_BYTES], unsigned long long* bar, int blk0,
                        int x5buf) {
    if (q == 0) {
      const bool with5 = X5 && x5buf >= 0;
      mbar_expect_tx(bar, 3 * AT_CHUNK_BYTES + (with5 ? AT_KEYS * 64 : 0));
#pragma unroll
      for (int s = 0; s < 3; ++s) bulk_g2s(dst[s], src + s * kv_split_stride + (long)blk0 * 1024, AT_CHUNK_BYTES, bar);
      if (with5) bulk_g2s(G.x5c[x5buf], x5 + (long)blk0 * 8 * 16, AT_KEYS * 64, bar);
    }
  };
  // s[i] += <Q5, K5[key i of my half]>   (keys are uniform across the warp: broadcast shared loads)
  auto add_s5 = [&](float (&s)[32], const float* xc, const float (&q5)[5]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float* row = xc + (half * 32 + i) * 16;
      const float4 k4 = *reinterpret_cast<const float4*>(row);
      const float k8 = row[8];
      s[i] += q5[0] * k4.x + q5[1] * k4.y + q5[2] * k4.z + q5[3] * k4.w + q5[4] * k8;
    }
  };
  // S = Q K^T for one 64-key chunk in K buffer `kb_`.  hi_only: just the leading bf16 x bf16 product (4 MMAs instead
  // of 24) -- enough for pass 1, which only needs each row's maximum to within a few units to keep exp() in range; the
  // softmax result does not depend on which shift is subtracted.
  auto issue_s = [&](int kb_, bool hi_only) {
    if (issuer_warp) {
      tc_fence_after();
      if (elect_one()) {
        const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
        unsigned accum = 0;
#pragma unroll
        for (int pr = 0; pr < 6; ++pr) {
          if (hi_only && pr < 5) continue;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_ts(tmem_wg + 96, tmem_wg + pa[pr] * 32 + kk * 8,
                    b_desc_ex(k_saddr + (kb_ * 3 + pb[pr]) * AT_CHUNK_BYTES + kk * 256, 128, 1024), accum);
            accum = 1;
          }
        }
        umma_commit(&S.mma_bar[wg_u]);
      }
      __syncwarp();
    }
  };
  // O (+)= P V for one chunk in V buffer `vb_`
  auto issue_pv = [&](int vb_, unsigned accum0) {
    if (issuer_warp) {
      tc_fence_after();
      if (elect_one()) {
        const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
        const unsigned idesc = umma_idesc(64, 1);  // B is MN-major: [key][d]
        unsigned accum = accum0;
#pragma unroll
        for (int pr = 0; pr < 6; ++pr)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_ts_i(tmem_wg + 192, tmem_wg + 96 + pa[pr] * 32 + kk * 8,
                      b_desc_ex(v_saddr + (vb_ * 3 + pb[pr]) * AT_CHUNK_BYTES + kk * 2048, 1024, 128), idesc, accum);
            accum = 1;
          }
        umma_commit(&S.mma_bar[wg_u]);
      }
      __syncwarp();
    }
  };
  auto wait_mma = [&]() {
    mbar_wait(&S.mma_bar[wg], mph);
    mph ^= 1;
    tc_fence_after();
  };

#ifdef ATTN_PROF
  __shared__ long long prof_acc[16];
  long long prof_t = clock64();
  if (tid == 0)
    for (int k = 0; k < 16; ++k) prof_acc[k] = 0;
#endif
  for (int tile = blockIdx.x * 2 + wg; tile < g.n_node_tiles; tile += gridDim.x * 2) {
    PROF_MARK(15);
    if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, 1);
    const int seg = g.node_tiles[2 * tile], node0 = g.node_tiles[2 * tile + 1];
    const int nvalid = min(EQD_TM, g.seg_ptr[seg + 1] - node0);
    const int pseg = seg < B ? seg + B : seg - B;
    const int j0 = g.seg_ptr[pseg], j1 = g.seg_ptr[pseg + 1];
    const int blk_lo = j0 >> 3, blk_hi = (j1 + 7) >> 3;
    const int nchunks = (blk_hi - blk_lo + 7) >> 3;
    const int node = node0 + r;
    const bool valid = r < nvalid;
    PROF_MARK(0);   // tile metadata (dependent global loads)
    if (nchunks > 0) load_chunk(k_g, G.k[0], &S.k_bar[wg][0], blk_lo, 0);
    float q5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (X5 && valid) {
#pragma unroll
      for (int e = 0; e < 5; ++e) q5[e] = x5[(long)node * 16 + 10 + e];
    }
    {  // Q row -> bf16x3 -> TMEM
      float v[32];
      const float4* sp = reinterpret_cast<const float4*>(proj + (long)node * pw + 128 + half * 32);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 t = valid ? sp[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        v[c4 * 4] = t.x; v[c4 * 4 + 1] = t.y; v[c4 * 4 + 2] = t.z; v[c4 * 4 + 3] = t.w;
      }
      store_half_split3(tmem + half * 16, v);
    }
    tc_fence_before();
    wg_barrier(wg);
    PROF_MARK(1);   // Q row -> TMEM + barrier
    // ---------------- pass 1: row maxima ----------------------------------------------------------------
    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      const int kb_ = c & 1;
      if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, (c << 4) | 2);
      mbar_wait(&S.k_bar[wg][kb_], kph[kb_]);
      kph[kb_] ^= 1;
      PROF_MARK(2);   // pass 1: K chunk wait
      issue_s(kb_, true);
      // the other K buffer was last read by the S GEMM of chunk c-1, already waited for: prefetch into it
      if (c + 1 < nchunks) load_chunk(k_g, G.k[kb_ ^ 1], &S.k_bar[wg][kb_ ^ 1], blk_lo + 8 * (c + 1), kb_ ^ 1);
      else load_chunk(k_g, G.k[kb_ ^ 1], &S.k_bar[wg][kb_ ^ 1], blk_lo, kb_ ^ 1);   // first chunk of pass 2
      if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, (c << 4) | 3);
      wait_mma();
      PROF_MARK(3);   // pass 1: issue + S(hi) MMA wait
      float s[32];
      tmem_ld32f(tmem + 96 + half * 32, s);
      if (X5) add_s5(s, G.x5c[kb_], q5);
      const int key0 = (blk_lo + 8 * c) * 8 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        int kn = key0 + i;
        mx = fmaxf(mx, (kn >= j0 && kn < j1) ? s[i] : -INFINITY);
      }
      tc_fence_before();
      wg_barrier(wg);  // S drained before the next S GEMM overwrites it
      PROF_MARK(4);   // pass 1: ld + max + barrier
    }
    G.red[r * 2 + half] = mx;
    wg_barrier(wg);
    mx = fmaxf(G.red[r * 2], G.red[r * 2 + 1]);
    // ---------------- pass 2: P = exp(S - max), O += P V ----------------------------------------------------
    float l = 0.f;
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
    if (nchunks > 0) load_chunk(v_g, G.v[0], &S.v_bar[wg][0], blk_lo, -1);
    float o5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < nchunks; ++c) {
      const int kb_ = (nchunks + c) & 1, vb_ = c & 1;   // K buffers keep alternating after pass 1
      if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, (c << 4) | 4);
      mbar_wait(&S.k_bar[wg][kb_], kph[kb_]);
      kph[kb_] ^= 1;
      PROF_MARK(5);   // pass 2: K chunk wait (+ row-max exchange on the first chunk)
      issue_s(kb_, false);
      if (c + 1 < nchunks) {
        load_chunk(k_g, G.k[kb_ ^ 1], &S.k_bar[wg][kb_ ^ 1], blk_lo + 8 * (c + 1), kb_ ^ 1);
        load_chunk(v_g, G.v[vb_ ^ 1], &S.v_bar[wg][vb_ ^ 1], blk_lo + 8 * (c + 1), -1);   // its last reader (P V of c-1) is done
      }
      if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, (c << 4) | 5);
      wait_mma();
      PROF_MARK(6);   // pass 2: issue + S MMA wait
      float s[32];
      tmem_ld32f(tmem + 96 + half * 32, s);
      if (X5) add_s5(s, G.x5c[kb_], q5);
      const int key0 = (blk_lo + 8 * c) * 8 + half * 32;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        int kn = key0 + i;
        float pj = (kn >= j0 && kn < j1) ? expf(s[i] - mx) : 0.f;
        s[i] = pj;
        l4[i & 3] += pj;
        if (X5) {   // the five extra output columns: o5 += p V5[key]
          const float* row = G.x5c[kb_] + (half * 32 + i) * 16;
          const float4 v4 = *reinterpret_cast<const float4*>(row + 4);
          o5[0] = fmaf(pj, v4.x, o5[0]);
          o5[1] = fmaf(pj, v4.y, o5[1]);
          o5[2] = fmaf(pj, v4.z, o5[2]);
          o5[3] = fmaf(pj, v4.w, o5[3]);
          o5[4] = fmaf(pj, row[9], o5[4]);
        }
      }
      l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tc_fence_before();
      PROF_MARK(7);   // pass 2: ld + exp
      wg_barrier(wg);  // every S value is in registers: the P splits may overwrite the S columns
      PROF_MARK(8);   // pass 2: barrier 1
      store_half_split3(tmem + 96 + half * 16, s);
      tc_fence_before();
      wg_barrier(wg);
      PROF_MARK(9);   // pass 2: P split/store + barrier 2
      if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, (c << 4) | 6);
      mbar_wait(&S.v_bar[wg][vb_], vph[vb_]);
      vph[vb_] ^= 1;
      issue_pv(vb_, 0u);
      if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, tile, (c << 4) | 7);
      wait_mma();      // P (= the S region) and this V buffer are free again
      PROF_MARK(10);  // pass 2: V wait + issue + P.V MMA wait
      // The tensor core truncates (round-toward-zero) every time it adds into an fp32 accumulator, a systematic
      // bias that grows with the number of accumulation steps; each 64-key chunk is therefore accumulated on its
      // own (4 full-magnitude steps) and the chunks are summed here with round-to-nearest FADDs.
      {
        float oc[32];
        tmem_ld32f(tmem + 192 + half * 32, oc);
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] += oc[i];
      }
      // Every thread must have OBSERVED this chunk's v_bar and P.V mma_bar phases before the issuing warp may start the
      // next ones on the same mbarriers (next chunk's S GEMM commit, the V refill two chunks ahead): a warp that is
      // held up for a microsecond between the barrier above and its try_wait would otherwise be lapped -- two phase
      // flips look like none -- and spin forever.  (This was a real, rare hang: ~1 in 10^4 launches back to back.)
      tc_fence_before();
      wg_barrier(wg);
      PROF_MARK(11);  // pass 2: O ld + accumulate + barrier 3
    }
    // ---------------- mu = O / l -----------------------------------------------------------------------------
    G.red[r * 2 + half] = l;
    if (X5) {
#pragma unroll
      for (int e = 0; e < 5; ++e) G.red5[(r * 2 + half) * 5 + e] = o5[e];
    }
    wg_barrier(wg);
    l = G.red[r * 2] + G.red[r * 2 + 1];
    {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      if (X5 && valid && half == 0) {   // mu[64..68], then zeros up to the row stride
        float e8[8];
#pragma unroll
        for (int e = 0; e < 5; ++e) e8[e] = (G.red5[r * 10 + e] + G.red5[r * 10 + 5 + e]) * inv;
        e8[5] = e8[6] = e8[7] = 0.f;
        float4* de = reinterpret_cast<float4*>(mu + (long)node * ldmu + 64);
        de[0] = make_float4(e8[0], e8[1], e8[2], e8[3]);
        de[1] = make_float4(e8[4], e8[5], e8[6], e8[7]);
      }
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(mu + (long)node * ldmu + half * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          dst[c4] = make_float4(o_acc[c4 * 4] * inv, o_acc[c4 * 4 + 1] * inv, o_acc[c4 * 4 + 2] * inv, o_acc[c4 * 4 + 3] * inv);
      }
    }
    tc_fence_before();
    wg_barrier(wg);
    PROF_MARK(12);  // mu = O / l, stores, barrier
  }
#ifdef ATTN_PROF
  if (tid == 0 && blockIdx.x == 0)
    for (int k = 0; k < 16; ++k) g_attn_prof[k] = prof_acc[k];
#endif
  if (q == 0) TRACE_PHASE(1, blockIdx.x * 2 + wg, 0xffff, 15);
  tc_fence_before();
  __syncthreads();
  TRACE_END(1);
  tmem_release(S.tmem_base, warp);
}

}  // namespace eqd

EQD_TRACE_SETTER(eqd_trace_set_attn)

#ifdef ATTN_PROF
extern "C" int eqd_attn_prof_read(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, eqd::g_attn_prof, sizeof(long long) * 16);
}
#endif

template <bool X5>
static int launch_attention_tc(const eqd_graph* g, const float* proj, int pw, const void* kv, const float* x5, float* mu,
                               int ldmu, void* stream) {
  if (reinterpret_cast<uintptr_t>(kv) & 15) return EQD_ERR_BAD_ARG;
  if (g->n_node_tiles <= 0) return EQD_OK;
  size_t smem = sizeof(eqd::AtSmem<X5>) + 128;
  EQD_SET_SMEM((eqd::attention_tc_kernel<X5>), smem);
  int grid = (g->n_node_tiles + 1) / 2;
  if (grid > 148) grid = 148;
  long split_stride = (long)((g->n_nodes + 7) / 8 + 8) * 1024;
  eqd::attention_tc_kernel<X5><<<grid, AT_THREADS, smem, (cudaStream_t)stream>>>(
      *g, proj, pw, reinterpret_cast<const unsigned char*>(kv), split_stride, x5, mu, ldmu);
  EQD_CUDA_LAUNCH_CHECK();
  return EQD_OK;
}

extern "C" int eqd_attention_tc(const eqd_graph* g, const float* proj, const void* kv, float* mu, void* stream) {
  if (!g || !proj || !kv || !mu) return EQD_ERR_BAD_ARG;
  return launch_attention_tc<false>(g, proj, 320, kv, nullptr, mu, EQD_HID, stream);
}

extern "C" int eqd_attention_tc0(const eqd_graph* g, const float* proj, const void* kv, const float* x5, float* mu,
                                 void* stream) {
  if (!g || !proj || !kv || !x5 || !mu) return EQD_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(x5) & 15) return EQD_ERR_BAD_ARG;
  return launch_attention_tc<true>(g, proj, 128 + 3 * 72, kv, x5, mu, EQD_H0_PAD, stream);
}
