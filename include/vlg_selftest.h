/*
 * vlg_selftest.h -- test hooks of libvlg_b200.so (not part of the drop-in boundary).
 * They exercise the tcgen05 building blocks in isolation so that descriptor / layout
 * mistakes show up as a failing unit test, not as a wrong energy.
 */
#ifndef VLG_SELFTEST_H_
#define VLG_SELFTEST_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* D[128,N] = A[128,K] * B^T with tcgen05.mma kind::tf32, A staged in TMEM (tcgen05.st),
 * B streamed into shared memory by a 1-D bulk async copy from the canonical no-swizzle image
 *   Bimg[(k/4)*N + n][k%4]            (b_mn_major = 0: B[n][k], K-major descriptor)
 * or, the SAME image read MN-major   (b_mn_major = 1: computes D[m][j] = sum_n A[m][n] * B[n][j],
 *   i.e. the transposed use needed by the backward pass; then A is [128,N] and D is [128,K]).
 * split3 != 0 runs the 3xTF32 scheme with Blo (same layout, residuals).
 * All pointers are device pointers; returns 0 or a negative VLG_ERR_* code. */
int vlg_selftest_umma(const float* A, const float* Bimg, const float* Blo, float* D, int N, int K,
                      int b_mn_major, int split3, void* stream);

/* Same with explicit descriptor strides (bytes; 0 = default) -- used to decode how the tensor core
 * interprets a shared-memory matrix descriptor. */
int vlg_selftest_umma_ex(const float* A, const float* Bimg, const float* Blo, float* D, int N, int K,
                         int b_mn_major, int split3, int lbo, int sbo, int kstep, void* stream);

/* D[128,N] = fp16(A[128,K]) * fp16(B)^T with tcgen05.mma kind::f16, fp32 accumulate.  A (fp32, device) is
 * converted to fp16 pairs in TMEM (two k per 32-bit column); Bimg16 is the fp16 image img16[(k/8)*N + n][k%8]. */
int vlg_selftest_umma_f16(const float* A, const void* Bimg16, float* D, int N, int K, void* stream);

/* Tensor-pipe rate probe: `iters` back-to-back kind::tf32 MMAs (M=128, K=8, A in TMEM) per CTA on `ctas`
 * CTAs; out[cta] = clock64 cycles from first issue to mbarrier-observed completion.  mode bit0 adds
 * concurrent TMEM ld/st traffic, bit1 concurrent shared-memory loads (interference study). */
int vlg_selftest_mma_rate(int N, int iters, int lbo, int sbo, int ctas, int mode, long long* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
