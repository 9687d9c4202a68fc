// Stand-alone probe of the tcgen05 mechanics used by the tensor-core edge stage (sm_100a):
//   D[128x64] (fp32, TMEM) = A[128xK] (bf16, TMEM, written by tcgen05.st) * B[64xK]^T (bf16, smem, K-major,
//   no-swizzle canonical layout), with the bf16x3 split ("bf16x6": 6 products) that emulates fp32.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe scripts/tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>

#define K_DIM 64
#define N_DIM 64

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no-swizzle canonical layout of a [N][K] bf16 operand: element (n,k) at byte
//   (k/8)*LBO + (n/8)*SBO + (n%8)*16 + (k%8)*2,  LBO = N*16, SBO = 128
__host__ __device__ inline int b_off_bytes(int n, int k) { return (k / 8) * (N_DIM * 16) + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; }

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);               // start address
  d |= (uint64_t)(((N_DIM * 16) >> 4) & 0x3FFF) << 16;  // leading byte offset (K direction)
  d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;           // stride byte offset (N direction)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  return d;                                             // layout_type = 0 (no swizzle)
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// split v into 3 bf16 (round to nearest): v ~ b0 + b1 + b2
__device__ __forceinline__ void split3(float v, __nv_bfloat16& b0, __nv_bfloat16& b1, __nv_bfloat16& b2) {
  b0 = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(b0);
  b1 = __float2bfloat16_rn(r);
  r -= __bfloat162float(b1);
  b2 = __float2bfloat16_rn(r);
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       float* __restrict__ D) {
  __shared__ __align__(128) unsigned char bsm[3][N_DIM * K_DIM * 2];  // 3 splits of B, canonical layout
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // B (weights) -> 3 bf16 splits in canonical layout
  for (int idx = tid; idx < N_DIM * K_DIM; idx += 128) {
    int n = idx / K_DIM, k = idx % K_DIM;
    __nv_bfloat16 b0, b1, b2;
    split3(W[n * K_DIM + k], b0, b1, b2);
    int off = b_off_bytes(n, k);
    *reinterpret_cast<__nv_bfloat16*>(&bsm[0][off]) = b0;
    *reinterpret_cast<__nv_bfloat16*>(&bsm[1][off]) = b1;
    *reinterpret_cast<__nv_bfloat16*>(&bsm[2][off]) = b2;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // make the generic-proxy smem writes of B visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t d_col = 0, a_col = 64;  // D: 64 columns; A splits: 3 x 32 columns

  // A row of this thread -> 3 bf16 splits -> TMEM (lane = row, 2 bf16 per 32-bit column)
  {
    uint32_t p0[32], p1[32], p2[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      __nv_bfloat16 x0, x1, x2, y0, y1, y2;
      split3(A[tid * K_DIM + 2 * c], x0, x1, x2);
      split3(A[tid * K_DIM + 2 * c + 1], y0, y1, y2);
      p0[c] = pack2(x0, y0);
      p1[c] = pack2(x1, y1);
      p2[c] = pack2(x2, y2);
    }
#define ST32(arr, col)                                                                                                  \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tmem_base + lane_base + (col)), \
               "r"(arr[0]), "r"(arr[1]), "r"(arr[2]), "r"(arr[3]), "r"(arr[4]), "r"(arr[5]), "r"(arr[6]), "r"(arr[7]),  \
               "r"(arr[8]), "r"(arr[9]), "r"(arr[10]), "r"(arr[11]), "r"(arr[12]), "r"(arr[13]), "r"(arr[14]),          \
               "r"(arr[15]), "r"(arr[16]), "r"(arr[17]), "r"(arr[18]), "r"(arr[19]), "r"(arr[20]), "r"(arr[21]),        \
               "r"(arr[22]), "r"(arr[23]), "r"(arr[24]), "r"(arr[25]), "r"(arr[26]), "r"(arr[27]), "r"(arr[28]),        \
               "r"(arr[29]), "r"(arr[30]), "r"(arr[31]) : "memory")
    ST32(p0, a_col);
    ST32(p1, a_col + 32);
    ST32(p2, a_col + 64);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // idesc: D=f32, A=B=bf16, K-major both, N=64, M=128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N_DIM >> 3) << 17) | ((128 >> 4) << 24);
    // products in increasing magnitude order: (a2,b0) (a0,b2) (a1,b1) (a1,b0) (a0,b1) (a0,b0)
    const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
    uint32_t accum = 0;
    for (int pr = 0; pr < 6; ++pr) {
      for (int kb = 0; kb < K_DIM / 16; ++kb) {
        uint32_t a_addr = tmem_base + a_col + pa[pr] * 32 + kb * 8;  // 16 bf16 = 8 columns per k-block
        uint64_t b_desc = make_b_desc(smem_u32(&bsm[pb[pr]][0]) + kb * 2 * (N_DIM * 16));
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base + d_col),
            "r"(a_addr), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,"
        "%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(tmem_base + lane_base + d_col));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 64; ++c) D[tid * N_DIM + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

int main() {
  std::vector<float> A(128 * K_DIM), W(N_DIM * K_DIM), D(128 * N_DIM, -1.f);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f + 0.1f;
  for (auto& v : W) v = (rand() / (float)RAND_MAX) * 0.5f + 0.01f;
  float *dA, *dW, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  probe_kernel<<<1, 128>>>(dA, dW, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_err32 = 0, max_ref = 0, bias = 0, bias32 = 0, rms = 0, rms32 = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N_DIM; ++n) {
      double ref = 0; float ref32 = 0.f;
      for (int k = 0; k < K_DIM; ++k) { ref += (double)A[m * K_DIM + k] * W[n * K_DIM + k]; ref32 = fmaf(A[m * K_DIM + k], W[n * K_DIM + k], ref32); }
      double e = D[m * N_DIM + n] - ref, e32 = (double)ref32 - ref, sg = ref >= 0 ? 1.0 : -1.0;
      max_err = fmax(max_err, fabs(e));
      max_err32 = fmax(max_err32, fabs(e32));
      max_ref = fmax(max_ref, fabs(ref));
      bias += e * sg; bias32 += e32 * sg; rms += e * e; rms32 += e32 * e32;
    }
  int cnt = 128 * N_DIM;
  printf("max|D - fp64 ref| = %.3e  (plain fp32 FMA loop: %.3e, max|ref| = %.3f)  D[0][0..3] = %f %f %f %f\n", max_err, max_err32,
         max_ref, D[0], D[1], D[2], D[3]);
  printf("signed bias toward larger magnitude: tc %.3e  fma %.3e ; rms: tc %.3e  fma %.3e\n", bias / cnt, bias32 / cnt, sqrt(rms / cnt), sqrt(rms32 / cnt));
  printf(max_err < 1e-5 ? "PROBE PASS\n" : "PROBE FAIL\n");
  return 0;
}
