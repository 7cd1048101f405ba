// Unit-test kernel for the tcgen05 building blocks (see include/vlg_selftest.h).
#include "../../include/vlg.h"
#include "../../include/vlg_selftest.h"
#include <cuda_fp16.h>

#include "vlg_common.cuh"
#include "vlg_tcgen05.cuh"

namespace vlg {
namespace {

using namespace tc;

// one CTA, 128 threads.  Rows of A / D = TMEM lanes.
__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bimg,
                                                            const float* __restrict__ Blo, float* __restrict__ D,
                                                            int N, int K, int mn_major, int split3, int lbo_o, int sbo_o, int kstep_o) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sB = reinterpret_cast<float*>(smem);              // N*K floats
  float* sBlo = sB + N * K;                                // N*K floats (split3)
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // contraction length / output width of this run
  const int KK = mn_major ? N : K;   // A has KK columns
  const int NN = mn_major ? K : N;   // D has NN columns

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  const uint32_t colA = 0, colAlo = 128, colD = 256;

  if (tid == 0) {
    const uint32_t bytes = uint32_t(N) * K * 4;
    mbar_expect_tx(&bar_b, split3 ? 2 * bytes : bytes);
    bulk_g2s(sB, Bimg, bytes, &bar_b);
    if (split3) bulk_g2s(sBlo, Blo, bytes, &bar_b);
  }
  // A row -> TMEM (hi and, for split3, the residual)
  for (int c0 = 0; c0 < KK; c0 += 16) {
    uint32_t v[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = A[size_t(tid) * KK + c0 + j];
      if (split3) {
        const uint32_t hi = __float_as_uint(a) & 0xFFFFE000u;
        v[j] = hi;
        lo[j] = __float_as_uint(a - __uint_as_float(hi));
      } else {
        v[j] = __float_as_uint(a);
      }
    }
    tmem_st16(tmem + lane_base + colA + c0, v);
    if (split3) tmem_st16(tmem + lane_base + colAlo + c0, lo);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    mbar_wait(&bar_b, 0);
    const uint32_t idesc = umma_idesc_tf32(NN, mn_major);
    const int nk = KK / 8;
    for (int pass = 0; pass < (split3 ? 3 : 1); ++pass) {
      // pass 0: Ahi*Bhi, pass 1: Alo*Bhi, pass 2: Ahi*Blo
      const uint32_t a_col = (pass == 1) ? colAlo : colA;
      const float* b_src = (pass == 2) ? sBlo : sB;
      for (int ks = 0; ks < nk; ++ks) {
        uint64_t desc;
        if (!mn_major) {
          // K-major: two 16B k-chunks per MMA, LBO = N*16 (next k-chunk), SBO = 128 (next 8 rows)
          desc = umma_smem_desc(smem_u32(b_src) + uint32_t(ks) * 2u * uint32_t(N) * 16u, uint32_t(N) * 16u, 128u);
        } else {
          // MN-major: k (=image row n) advances by 16 B; 8 k per MMA = 128 B; MN groups of 4 at SBO = N*16
          desc = umma_smem_desc(smem_u32(b_src) + uint32_t(ks) * (kstep_o ? uint32_t(kstep_o) : 128u), lbo_o ? uint32_t(lbo_o) : 128u,
                                sbo_o ? uint32_t(sbo_o) : uint32_t(N) * 16u);
        }
        umma_tf32_ts(tmem + colD, tmem + a_col + uint32_t(ks) * 8u, desc, idesc, (pass | ks) ? 1u : 0u);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < NN; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_base + colD + c0, v);  // ld + wait in one asm statement
#pragma unroll
    for (int j = 0; j < 16; ++j) D[size_t(tid) * NN + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace vlg

extern "C" int vlg_selftest_umma_ex(const float* A, const float* Bimg, const float* Blo, float* D, int N, int K,
                                    int b_mn_major, int split3, int lbo, int sbo, int kstep, void* stream);

extern "C" int vlg_selftest_umma(const float* A, const float* Bimg, const float* Blo, float* D, int N, int K,
                                 int b_mn_major, int split3, void* stream) {
  return vlg_selftest_umma_ex(A, Bimg, Blo, D, N, K, b_mn_major, split3, 0, 0, 0, stream);
}

extern "C" int vlg_selftest_umma_ex(const float* A, const float* Bimg, const float* Blo, float* D, int N, int K,
                                    int b_mn_major, int split3, int lbo, int sbo, int kstep, void* stream) {
  if (!A || !Bimg || !D || (split3 && !Blo)) return VLG_ERR_INVALID_ARGUMENT;
  if (N % 16 || K % 16 || N < 16 || N > 128 || K < 16 || K > 128) return VLG_ERR_UNSUPPORTED;
  const size_t smem = size_t(N) * K * 4 * 2;
  cudaError_t e = cudaFuncSetAttribute(vlg::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return VLG_ERR_CUDA;
  vlg::umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, Bimg, Blo, D, N, K, b_mn_major, split3, lbo, sbo, kstep);
  return cudaGetLastError() == cudaSuccess ? VLG_OK : VLG_ERR_CUDA;
}

// ---- tensor-pipe rate probe: `iters` back-to-back kind::tf32 MMAs (M=128, K=8, A in TMEM,
// B in shared memory), one commit at the end; reports clock64 cycles per CTA. ----
namespace vlg {
namespace {
__global__ void __launch_bounds__(384) mma_rate_kernel(int N, int iters, int lbo, int sbo, int mode, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_mma, bar_scratch;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 144 * 128; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) {
    tc::mbar_init(&bar_mma, 1);
    tc::mbar_init(&bar_scratch, 1);
    tc::fence_mbar_init();
    done = 0;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    // The whole warp runs the issue loop convergently, one elected lane's instructions take effect (as in the
    // product kernel: a divergent `if (tid == 0)` loop is itself the bottleneck at ~100+ cycles per MMA).
    // mode bit8 (256): M = 64 instead of 128; bit9 (512): kind::f16 (K = 16) instead of kind::tf32
    const uint32_t leader = tc::elect_one();
    uint32_t idesc = (mode & 512) ? tc::umma_idesc_f16(N) : tc::umma_idesc_tf32(N, 0);
    if (mode & 256) idesc = (idesc & ~(0x1Fu << 24)) | (uint32_t(64 >> 4) << 24);
    const bool f16k = (mode & 512) != 0;
    const uint32_t sb = tc::smem_u32(smem);
    const long long t0 = clock64();
    // mode bits 5..7: commit to a scratch mbarrier every 4 MMAs (32), switch the accumulator between two
    // column ranges every 8 MMAs (64), start a fresh accumulation (accumulate = 0) every 16 MMAs (128)
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 7;
      const uint64_t desc = tc::umma_smem_desc(sb + uint32_t(ks) * 2u * uint32_t(lbo), uint32_t(lbo), uint32_t(sbo));
      const uint32_t dcol = ((mode & 64) && ((i >> 3) & 1)) ? 256u : 128u;
      const uint32_t accum = ((mode & 128) && (i & 15) == 0) ? 0u : 1u;
      if (f16k)
        tc::umma_f16_ts_elect(tmem + dcol, tmem + uint32_t(ks) * 8u, desc, idesc, accum, leader);
      else
        tc::umma_tf32_ts_elect(tmem + dcol, tmem + uint32_t(ks) * 8u, desc, idesc, accum, leader);
      if ((mode & 32) && (i & 3) == 3) tc::umma_commit_elect(&bar_scratch, leader);
    }
    tc::umma_commit_elect(&bar_mma, leader);
    tc::mbar_wait(&bar_mma, 0);
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
    if (iters > 0 && (mode & 16)) {
      // completion latency of one MMA + commit issued into an idle pipe
      const long long t1 = clock64();
      tc::umma_tf32_ts_elect(tmem + 128u, tmem, tc::umma_smem_desc(sb, uint32_t(lbo), uint32_t(sbo)), idesc, 1u, leader);
      tc::umma_commit_elect(&bar_mma, leader);
      tc::mbar_wait(&bar_mma, 1);
      if (tid == 0) out[2 * gridDim.x + blockIdx.x] = clock64() - t1;
    }
    if (tid == 0) done = 1;
  } else if (warp >= 4) {
    // interference generators: mode bit0 = TMEM ld/st traffic (columns 384..447), bit1 = shared-memory loads
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    float acc = 0.f;
    if (mode & 4) {
      // latency of a TMEM load+store round trip while the MMA stream runs (or not: iters = 0):
      // 256 back-to-back {ld x32, wait, st x32, wait}; warps 4..7 only when bit3 is set
      if (!(mode & 8) || warp < 8) {
        const uint32_t col = tmem + lane_addr + 384u + uint32_t((warp >> 2) & 1) * 32u;
        const long long t0 = clock64();
        for (int it = 0; it < 256; ++it) {
          uint32_t v[32];
          tc::tmem_ld32_sync(col, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += 1u;
          tc::tmem_st32(col, v);
          tc::tmem_wait_st();
        }
        if (tid == 128) out[gridDim.x + blockIdx.x] = clock64() - t0;
      }
    } else
    while (!done) {
      if (mode & 1) {
        uint32_t v[32];
        tc::tmem_ld32_sync(tmem + lane_addr + 384u + uint32_t((warp >> 2) & 1) * 32u, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += 1u;
        tc::tmem_st32(tmem + lane_addr + 384u + uint32_t((warp >> 2) & 1) * 32u, v);
        tc::tmem_wait_st();
      }
      if (mode & 2) {
        const float4* p4 = reinterpret_cast<const float4*>(smem) + (tid & 31);
#pragma unroll
        for (int j = 0; j < 64; ++j) { float4 q = p4[j * 32]; acc += q.x + q.y + q.z + q.w; }
      }
      if (mode == 0) break;
    }
    if (acc == 123.456f) out[0] = 0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace vlg

extern "C" int vlg_selftest_mma_rate(int N, int iters, int lbo, int sbo, int ctas, int mode, long long* out, void* stream) {
  const size_t smem = 144 * 128 * 4;
  cudaError_t e = cudaFuncSetAttribute(vlg::mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return VLG_ERR_CUDA;
  vlg::mma_rate_kernel<<<ctas, 384, smem, static_cast<cudaStream_t>(stream)>>>(N, iters, lbo, sbo, mode, out);
  return cudaGetLastError() == cudaSuccess ? VLG_OK : VLG_ERR_CUDA;
}

// ---- kind::f16 building block: A[128,K] fp32 -> fp16 pairs in TMEM (two k per column, even k in the
// low half), B = fp16 image img16[k/8][n][k%8] bulk-copied to shared memory, fp32 accumulate ----
namespace vlg {
namespace {
__global__ void __launch_bounds__(128) umma_f16_selftest_kernel(const float* __restrict__ A, const void* __restrict__ Bimg,
                                                                float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&bar_b, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  const uint32_t colA = 0, colD = 256;
  if (tid == 0) {
    const uint32_t bytes = uint32_t(N) * K * 2;
    tc::mbar_expect_tx(&bar_b, bytes);
    tc::bulk_g2s(smem, Bimg, bytes, &bar_b);
  }
  for (int c0 = 0; c0 < K / 2; c0 += 16) {
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const __half2 h = __floats2half2_rn(A[size_t(tid) * K + 2 * (c0 + j)], A[size_t(tid) * K + 2 * (c0 + j) + 1]);
      v[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tc::tmem_st16(tmem + lane_base + colA + c0, v);
  }
  tc::tmem_wait_st();
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    const uint32_t leader = tc::elect_one();
    tc::tc_fence_after();
    tc::mbar_wait(&bar_b, 0);
    const uint32_t idesc = tc::umma_idesc_f16(N);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t desc = tc::umma_smem_desc(tc::smem_u32(smem) + uint32_t(ks) * 2u * uint32_t(N) * 16u, uint32_t(N) * 16u, 128u);
      tc::umma_f16_ts_elect(tmem + colD, tmem + colA + uint32_t(ks) * 8u, desc, idesc, ks ? 1u : 0u, leader);
    }
    tc::umma_commit_elect(&bar_mma, leader);
  }
  tc::mbar_wait(&bar_mma, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tc::tmem_ld16(tmem + lane_base + colD + c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) D[size_t(tid) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace vlg

extern "C" int vlg_selftest_umma_f16(const float* A, const void* Bimg16, float* D, int N, int K, void* stream) {
  if (!A || !Bimg16 || !D) return VLG_ERR_INVALID_ARGUMENT;
  if (N % 16 || K % 16 || N < 16 || N > 128 || K < 16 || K > 128) return VLG_ERR_UNSUPPORTED;
  const size_t smem = size_t(N) * K * 2;
  vlg::umma_f16_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, Bimg16, D, N, K);
  return cudaGetLastError() == cudaSuccess ? VLG_OK : VLG_ERR_CUDA;
}
