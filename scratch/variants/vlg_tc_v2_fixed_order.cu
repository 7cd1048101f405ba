// tcgen05 (TF32) geodesic step kernel -- the tensor-core variant (<=1e-3 relative on lengths).
//
// One CTA (576 threads) = one curve, persistent over `steps` Adam steps.  The two 128-wide
// decoder layers and their transposes run as tcgen05.mma kind::tf32 with
//   * M = 128 curve points = the 128 TMEM lanes,
//   * the A operand (activations) living in TENSOR MEMORY: the epilogue threads write the
//     next layer's input back with tcgen05.st, in place of the accumulator they just read,
//   * the B operand (weights) streamed from L2 into a shared-memory ring by the TMA engine
//     (1-D bulk copies of pre-packed no-swizzle K-major images, mbarrier complete_tx),
//   * fp32 accumulators in TMEM, read back with tcgen05.ld.
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer (one lane), warps 2-9 and 10-17 =
// two epilogue groups of 8 warps.  Each group owns a "chain" of 256 TMEM columns and alternate
// decoders, so one chain's CUDA-core epilogue overlaps the other chain's MMAs; inside a group
// two threads share a curve point (TMEM lane) and split its columns, which gives the SM four
// epilogue warps per scheduler to hide TMEM / shared-memory latency.
//
// Per 128-point tile: forward of all K decoders (layer 1 on CUDA cores, exact fp32) ->
// selected outputs accumulate into Diff[m][segment] (shared memory, fp32) -> energy ->
// backward of all K decoders (input gradient only; layer-2 ReLU mask as bits in shared memory,
// layer-1 mask recomputed) -> dz -> d(omega).  Penalty gradient and Adam as in vlg_simt.cu.
#include "vlg_common.cuh"
#include "vlg_kernels.h"
#include "vlg_tcgen05.cuh"

namespace vlg {

namespace {

using namespace tc;

constexpr int TC_THREADS = 576;       // 2 + 16 warps
constexpr int GROUP_THREADS = 256;    // one epilogue group (chain)
constexpr int EPI_THREADS = 512;
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGES = 6;
constexpr int DIFF_STRIDE = 52;

// the four tensor-core GEMMs of one decoder
struct OpInfo {
  int img_off;   // float offset of the B image inside the decoder record
  int nstages;   // 16 KB stages
  int n;         // MMA N
  int kper;      // contraction length per stage
  int a_col;     // chain-relative TMEM column of A
  int d_col;     // chain-relative TMEM column of D
};
__device__ __forceinline__ OpInfo op_info(int op) {
  switch (op) {
    case 0: return {OFF_W2_UMMA, 4, 128, 32, 0, 128};    // F2: D2(Y) = A1(X) * W2^T
    case 1: return {OFF_W3_UMMA, 2, 64, 64, 128, 0};     // F3: D3(X[0:64]) = A2(Y) * W3^T
    case 2: return {OFF_W3T_UMMA, 2, 128, 32, 0, 128};   // B3: D4(Y) = G(X[0:64]) * W3
    default: return {OFF_W2T_UMMA, 4, 128, 32, 128, 0};  // B2: D5(X) = A4(Y) * W2
  }
}

// round-to-nearest to TF32 for finite values: the tensor core ignores the 13 low mantissa bits
__device__ __forceinline__ uint32_t tf32_round_bits(uint32_t b) { return b + 0x1000u; }
// relu, then TF32 round-to-nearest on the bit pattern.  (An integer-max formulation,
// max(int(bits + 0x1000), 0), produced wrong results when combined with the packed f32x2
// intrinsics under nvcc 12.9 -- keep the float max.)
__device__ __forceinline__ uint32_t relu_tf32(float v) { return __float_as_uint(fmaxf(v, 0.f)) + 0x1000u; }
#ifdef VLG_NO_PACKED
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#else
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
#endif

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct TcSmem {
  unsigned char* ring;  // NSTAGES * 16 KB
  float* Diff;          // M*128*52
  uint32_t* mask2;      // K*128*4 words
  uint8_t* sel;         // MAX_M*2*128
  float* sw;            // [chain][buf] 576 floats
  float2* zs;           // 128
  float* ts;            // 128
  float2* dzs;          // 4*128
  float* coef;          // 64
  float* basis;         // 288
  float* om;            // 56
  float* gacc;          // 20
  float* red;           // 4*20 + 32
  uint64_t* bars;       // full[NSTAGES], empty[NSTAGES], a_ready[2], acc_ready[2]
  uint32_t* tmem_base;
  volatile int* turn;
};

__device__ __forceinline__ TcSmem tc_carve(unsigned char* base, int M, int K) {
  TcSmem s;
  s.ring = base;
  float* f = reinterpret_cast<float*>(base + NSTAGES * STAGE_BYTES);
  s.Diff = f; f += M * 128 * DIFF_STRIDE;
  s.sw = f; f += 4 * 576;
  s.zs = reinterpret_cast<float2*>(f); f += 256;
  s.ts = f; f += 128;
  s.dzs = reinterpret_cast<float2*>(f); f += 1024;
  s.coef = f; f += 64;
  s.basis = f; f += 4 * MAX_NPOLY * MAX_KB;
  s.om = f; f += 3 * 2 * MAX_KB + 2;
  s.gacc = f; f += 2 * MAX_KB + 2;
  s.red = f; f += 112;
  s.bars = reinterpret_cast<uint64_t*>(f); f += 2 * (2 * NSTAGES + 4);
  s.tmem_base = reinterpret_cast<uint32_t*>(f); f += 2;
  s.turn = reinterpret_cast<volatile int*>(f); f += 2;
  s.sel = reinterpret_cast<uint8_t*>(f); f += MAX_M * 2 * 128 / 4;
  s.mask2 = reinterpret_cast<uint32_t*>(f);
  (void)K;
  return s;
}

}  // namespace

static size_t tc_smem_bytes(int M, int K) {
  size_t fl = size_t(M) * 128 * DIFF_STRIDE + 4 * 576 + 256 + 128 + 1024 + 64 + 4 * MAX_NPOLY * MAX_KB +
              (3 * 2 * MAX_KB + 2) + (2 * MAX_KB + 2) + 112 + 2 * (2 * NSTAGES + 4) + 2 + 2 + MAX_M * 2 * 128 / 4;
  return size_t(NSTAGES) * STAGE_BYTES + fl * 4 + size_t(K) * 128 * 16;
}

template <bool GRAD>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_curve_kernel(StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x;
  const int M = p.M, K = p.K, T = p.T, n_poly = p.n_poly, Kb = p.Kb;
  TcSmem s = tc_carve(smem_raw, M, K);
  uint64_t* full = s.bars;
  uint64_t* empty = s.bars + NSTAGES;
  uint64_t* a_ready = s.bars + 2 * NSTAGES;
  uint64_t* acc_ready = s.bars + 2 * NSTAGES + 2;
  const int ntiles = (T - 1 + TILE_SEGS - 1) / TILE_SEGS;

  if (tid == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&a_ready[0], GROUP_THREADS);
    mbar_init(&a_ready[1], GROUP_THREADS);
    mbar_init(&acc_ready[0], 1);
    mbar_init(&acc_ready[1], 1);
    *s.turn = 0;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_base;

  // The sequence of tensor-core ops is identical for every tile: for each decoder pair
  // (kA = 2p on chain 0, kB = 2p+1 on chain 1): forward F2(kA) F2(kB) F3(kA) F3(kB), and after
  // all pairs the backward B3(kA) B3(kB) B2(kA) B2(kB).  Producer and MMA issuer walk it in
  // lock step through the ring; the epilogue groups follow through a_ready / acc_ready.
  const int npairs = (K + 1) / 2;
  const long total_tiles = long(p.steps) * ntiles;

  if (warp == 0) {
    // ================= weight producer (TMA bulk copies) =================
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (long tl = 0; tl < total_tiles; ++tl)
        for (int phase = 0; phase < (GRAD ? 2 : 1); ++phase)
          for (int pr = 0; pr < npairs; ++pr)
            for (int o = 0; o < 2; ++o)
              for (int c = 0; c < 2; ++c) {
                const int k = 2 * pr + c;
                if (k >= K) continue;
                const OpInfo oi = op_info(phase * 2 + o);
                const char* src = reinterpret_cast<const char*>(dec_ptr(p.packed, k) + oi.img_off);
                for (int st = 0; st < oi.nstages; ++st) {
                  mbar_wait(&empty[slot], ph ^ 1);
                  mbar_expect_tx(&full[slot], STAGE_BYTES);
                  bulk_g2s(s.ring + slot * STAGE_BYTES, src + size_t(st) * STAGE_BYTES, STAGE_BYTES, &full[slot]);
                  if (++slot == NSTAGES) { slot = 0; ph ^= 1; }
                }
              }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      uint32_t ph_a[2] = {0, 0};
      for (long tl = 0; tl < total_tiles; ++tl)
        for (int phase = 0; phase < (GRAD ? 2 : 1); ++phase)
          for (int pr = 0; pr < npairs; ++pr)
            for (int o = 0; o < 2; ++o)
              for (int c = 0; c < 2; ++c) {
                const int k = 2 * pr + c;
                if (k >= K) continue;
                const OpInfo oi = op_info(phase * 2 + o);
                const uint32_t idesc = umma_idesc_tf32(oi.n, 0);
                const uint32_t chain = tmem + uint32_t(c) * 256u;
                mbar_wait(&a_ready[c], ph_a[c]);
                ph_a[c] ^= 1;
                tc_fence_after();
                for (int st = 0; st < oi.nstages; ++st) {
                  mbar_wait(&full[slot], ph);
                  tc_fence_after();
                  const uint32_t sbase = smem_u32(s.ring + slot * STAGE_BYTES);
                  const int nk = oi.kper / 8;
                  for (int ks = 0; ks < nk; ++ks) {
                    const uint64_t desc =
                        umma_smem_desc(sbase + uint32_t(ks) * 2u * uint32_t(oi.n) * 16u, uint32_t(oi.n) * 16u, 128u);
                    umma_tf32_ts(chain + oi.d_col, chain + oi.a_col + uint32_t(st * oi.kper + ks * 8), desc, idesc,
                                 (st | ks) ? 1u : 0u);
                  }
                  umma_commit(&empty[slot]);
                  if (++slot == NSTAGES) { slot = 0; ph ^= 1; }
                }
                umma_commit(&acc_ready[c]);
              }
    }
  } else {
    // ================= epilogue groups =================
    const int ew = warp - 2;                    // 0..15
    const int chain_id = ew >> 3;               // group / chain
    const int half = (ew >> 2) & 1;             // which 64 of the 128 columns
    const int row = (warp & 3) * 32 + lane;     // TMEM lane = curve point of the tile
    const int tg = half * 128 + row;            // 0..255 inside the group
    const int t512 = chain_id * 256 + tg;       // 0..511 over both groups
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    const uint32_t chain = tmem + lane_addr + uint32_t(chain_id) * 256u;
    const uint32_t colX = chain, colY = chain + 128u;
    const int col0 = half * 64;                 // this thread's hidden units
    const int xc0 = half * 32;                  // this thread's output / G columns
    const int bar_id = 1 + chain_id;
    float* swbuf = s.sw + chain_id * 2 * 576;
    int swsel = 0;
    uint32_t ph_acc = 0;
    const float coefm = 2.0f / float(M);

    for (int i = t512; i < 4 * n_poly * Kb; i += EPI_THREADS) s.basis[i] = p.basis[i];
    if (t512 < 2 * Kb) {
      s.om[t512] = p.omega[size_t(n) * 2 * Kb + t512];
      if (GRAD) {
        s.om[2 * MAX_KB + t512] = p.adam_m[size_t(n) * 2 * Kb + t512];
        s.om[4 * MAX_KB + t512] = p.adam_v[size_t(n) * 2 * Kb + t512];
      }
    }
    const float2 pa = make_float2(p.a[2 * n], p.a[2 * n + 1]);
    const float2 pb = make_float2(p.b[2 * n], p.b[2 * n + 1]);
    // small weights (W1, b1, b2, b3) of this group's first decoder
    if (chain_id < K && tg < 144) cp_async16(swbuf + tg * 4, dec_ptr(p.packed, chain_id) + tg * 4);
    named_bar(3, EPI_THREADS);

    for (int step = 0; step < p.steps; ++step) {
      if (t512 < 8 * n_poly) {
        const int r = t512 >> 1, d = t512 & 1;
        float acc = 0.f;
        for (int k = 0; k < Kb; ++k) acc = fmaf(s.basis[r * Kb + k], s.om[2 * k + d], acc);
        s.coef[t512] = acc;
      }
      if (t512 < 2 * MAX_KB) s.gacc[t512] = 0.f;
      float e_tot = 0.f, l_tot = 0.f;  // meaningful in t512 == 0
      named_bar(3, EPI_THREADS);

      for (int tile = 0; tile < ntiles; ++tile) {
        const int seg0 = tile * TILE_SEGS;
        const int nseg = min(TILE_SEGS, T - 1 - seg0);
        const int turn0 = (step * ntiles + tile) * K;
        const bool last_tile = (step == p.steps - 1) && (tile == ntiles - 1);
        // ---- tile setup ----
        if (chain_id == 0 && half == 0) {
          const int ti = min(seg0 + row, T - 1);
          const float t = p.t[ti];
          s.ts[row] = t;
          s.zs[row] = spline_point(t, n_poly, s.coef, pa, pb);
        } else if (chain_id == 1 && half == 0) {
          if (p.draws != nullptr) {
            for (int m = 0; m < M; ++m)
              for (int role = 0; role < 2; ++role) {
                uint8_t v = 255;
                if (row < nseg)
                  v = p.draws[(((size_t(n) * p.steps + step) * M + m) * 2 + role) * size_t(T - 1) + seg0 + row];
                s.sel[(m * 2 + role) * 128 + row] = v;
              }
          } else {
            for (int jp = 0; jp < (M + 1) / 2; ++jp) {
              uint32_t d[4] = {255u, 255u, 255u, 255u};
              if (row < nseg)
                counter_draws4(p.seed, uint32_t(p.curve_id0 + n), uint32_t(p.step0 + step), uint32_t(seg0 + row),
                               uint32_t(jp), uint32_t(K), d);
              for (int q = 0; q < 4; ++q) {
                const int m = 2 * jp + (q >> 1);
                if (m < M) s.sel[(m * 2 + (q & 1)) * 128 + row] = uint8_t(d[q]);
              }
            }
          }
        }
        for (int i = t512; i < M * 128 * DIFF_STRIDE / 4; i += EPI_THREADS)
          reinterpret_cast<float4*>(s.Diff)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float dzx = 0.f, dzy = 0.f;
        named_bar(3, EPI_THREADS);
        const float2 z = s.zs[row];
        const float2 zx2 = make_float2(z.x, z.x), zy2 = make_float2(z.y, z.y);

        // =============================== forward ===============================
        for (int k = chain_id; k < K; k += 2) {
          // small weights of decoder k were prefetched into swbuf[swsel]; prefetch the next item's
          cp_async_wait_all();
          named_bar(bar_id, GROUP_THREADS);
          const float* sw = swbuf + swsel * 576;
          {
            int kn = k + 2;
            if (kn >= K) kn = GRAD ? chain_id : (last_tile ? -1 : chain_id);
            if (kn >= 0 && kn < K && tg < 144)
              cp_async16(swbuf + (swsel ^ 1) * 576 + tg * 4, dec_ptr(p.packed, kn) + tg * 4);
          }
          swsel ^= 1;
          // layer 1 (CUDA cores, fp32) -> A1 in X[col0 : col0+64]
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int c = col0 + c0 + j;
              const float2 wx = *reinterpret_cast<const float2*>(sw + OFF_W1X + c);
              const float2 wy = *reinterpret_cast<const float2*>(sw + OFF_W1Y + c);
              const float2 bb = *reinterpret_cast<const float2*>(sw + OFF_B1 + c);
              const float2 h = ffma2(wy, zy2, ffma2(wx, zx2, bb));
              v[j] = relu_tf32(h.x);
              v[j + 1] = relu_tf32(h.y);
            }
            tmem_st32(colX + col0 + c0, v);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&a_ready[chain_id]);
          // layer 2 epilogue: D2 (Y) -> relu(+b2) -> A2 (Y, in place), mask bits
          mbar_wait(&acc_ready[chain_id], ph_acc);
          ph_acc ^= 1;
          tc_fence_after();
          {
            uint32_t v0[32], v1[32];
            tmem_ld32x2_sync(colY + col0, colY + col0 + 32, v0, v1);
            uint32_t bits0 = 0, bits1 = 0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 b0 = *reinterpret_cast<const float2*>(sw + OFF_B2 + col0 + j);
              const float2 b1 = *reinterpret_cast<const float2*>(sw + OFF_B2 + col0 + 32 + j);
              const float2 h0 = fadd2(make_float2(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1])), b0);
              const float2 h1 = fadd2(make_float2(__uint_as_float(v1[j]), __uint_as_float(v1[j + 1])), b1);
              if (h0.x > 0.f) bits0 |= 1u << j;
              if (h0.y > 0.f) bits0 |= 2u << j;
              if (h1.x > 0.f) bits1 |= 1u << j;
              if (h1.y > 0.f) bits1 |= 2u << j;
              v0[j] = relu_tf32(h0.x);
              v0[j + 1] = relu_tf32(h0.y);
              v1[j] = relu_tf32(h1.x);
              v1[j + 1] = relu_tf32(h1.y);
            }
            tmem_st32(colY + col0, v0);
            tmem_st32(colY + col0 + 32, v1);
            if (GRAD)
              *reinterpret_cast<uint2*>(s.mask2 + (k * 128 + row) * 4 + half * 2) = make_uint2(bits0, bits1);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&a_ready[chain_id]);
          // layer 3 epilogue: D3 (X[0:64]) + b3 -> Diff (this thread: columns xc0 .. xc0+31)
          mbar_wait(&acc_ready[chain_id], ph_acc);
          ph_acc ^= 1;
          tc_fence_after();
          float x[32];
          {
            uint32_t xv[32];
            tmem_ld32_sync(colX + xc0, xv);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(sw + OFF_B3 + xc0 + j);
              x[j] = __uint_as_float(xv[j]) + bb.x;
              x[j + 1] = __uint_as_float(xv[j + 1]) + bb.y;
              x[j + 2] = __uint_as_float(xv[j + 2]) + bb.z;
              x[j + 3] = __uint_as_float(xv[j + 3]) + bb.w;
            }
          }
          // Diff is shared by both groups: updates are serialised in decoder order
          if (tg == 0) {
            while (*s.turn != turn0 + k) {
            }
            __threadfence_block();
          }
          named_bar(bar_id, GROUP_THREADS);
          const int nq = half ? (DIFF_STRIDE - 32) / 4 : 8;  // float4 groups of this thread's columns
          // role 0: this point is the left end of its segment
          for (int m = 0; m < M; ++m)
            if (s.sel[(m * 2 + 0) * 128 + row] == k) {
              float4* d = reinterpret_cast<float4*>(s.Diff + (m * 128 + row) * DIFF_STRIDE + xc0);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < nq) {
                  float4 v = d[q];
                  v.x -= x[4 * q]; v.y -= x[4 * q + 1]; v.z -= x[4 * q + 2]; v.w -= x[4 * q + 3];
                  d[q] = v;
                }
            }
          named_bar(bar_id, GROUP_THREADS);
          // role 1: this point is the right end of the previous segment
          if (row >= 1)
            for (int m = 0; m < M; ++m)
              if (s.sel[(m * 2 + 1) * 128 + row - 1] == k) {
                float4* d = reinterpret_cast<float4*>(s.Diff + (m * 128 + row - 1) * DIFF_STRIDE + xc0);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (q < nq) {
                    float4 v = d[q];
                    v.x += x[4 * q]; v.y += x[4 * q + 1]; v.z += x[4 * q + 2]; v.w += x[4 * q + 3];
                    d[q] = v;
                  }
              }
          __threadfence_block();
          named_bar(bar_id, GROUP_THREADS);
          if (tg == 0) *s.turn = turn0 + k + 1;
        }
        named_bar(3, EPI_THREADS);

        // =============================== energy ===============================
        // (columns >= X of a Diff row are never written and stay zero)
        {
          float e = 0.f, l = 0.f;
          for (int idx = t512; idx < M * 128; idx += EPI_THREADS) {
            const int r = idx & 127;
            if (r < nseg) {
              const float4* d = reinterpret_cast<const float4*>(s.Diff + idx * DIFF_STRIDE);
              float q = 0.f;
#pragma unroll
              for (int c = 0; c < DIFF_STRIDE / 4; ++c) {
                const float4 v = d[c];
                q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
              }
              e += q;
              l += sqrtf(q);
            }
          }
          e = warp_sum(e);
          l = warp_sum(l);
          if (lane == 0) { s.red[80 + ew] = e; s.red[96 + ew] = l; }
        }

        if (GRAD) {
          // =============================== backward ===============================
          for (int k = chain_id; k < K; k += 2) {
            cp_async_wait_all();
            named_bar(bar_id, GROUP_THREADS);
            const float* sw = swbuf + swsel * 576;
            {
              int kn = k + 2;
              if (kn >= K) kn = last_tile ? -1 : chain_id;
              if (kn >= 0 && kn < K && tg < 144)
                cp_async16(swbuf + (swsel ^ 1) * 576 + tg * 4, dec_ptr(p.packed, kn) + tg * 4);
            }
            swsel ^= 1;
            // G = dE/dx_k (this point), columns xc0 .. xc0+31 -> X[xc0 : xc0+32]
            {
              float g[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = 0.f;
              const int nq = half ? (DIFF_STRIDE - 32) / 4 : 8;
              for (int m = 0; m < M; ++m) {
                if (row >= 1 && s.sel[(m * 2 + 1) * 128 + row - 1] == k) {
                  const float4* d = reinterpret_cast<const float4*>(s.Diff + (m * 128 + row - 1) * DIFF_STRIDE + xc0);
#pragma unroll
                  for (int q = 0; q < 8; ++q)
                    if (q < nq) {
                      const float4 v = d[q];
                      g[4 * q] += v.x; g[4 * q + 1] += v.y; g[4 * q + 2] += v.z; g[4 * q + 3] += v.w;
                    }
                }
                if (s.sel[(m * 2 + 0) * 128 + row] == k) {
                  const float4* d = reinterpret_cast<const float4*>(s.Diff + (m * 128 + row) * DIFF_STRIDE + xc0);
#pragma unroll
                  for (int q = 0; q < 8; ++q)
                    if (q < nq) {
                      const float4 v = d[q];
                      g[4 * q] -= v.x; g[4 * q + 1] -= v.y; g[4 * q + 2] -= v.z; g[4 * q + 3] -= v.w;
                    }
                }
              }
              uint32_t v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = tf32_round_bits(__float_as_uint(coefm * g[j]));
              tmem_st32(colX + xc0, v);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&a_ready[chain_id]);
            // dh2 = (G W3) * mask2 -> A4 (Y, in place)
            mbar_wait(&acc_ready[chain_id], ph_acc);
            ph_acc ^= 1;
            tc_fence_after();
            {
              uint32_t v0[32], v1[32];
              tmem_ld32x2_sync(colY + col0, colY + col0 + 32, v0, v1);
              const uint2 bits = *reinterpret_cast<const uint2*>(s.mask2 + (k * 128 + row) * 4 + half * 2);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v0[j] = ((bits.x >> j) & 1u) ? tf32_round_bits(v0[j]) : 0u;
                v1[j] = ((bits.y >> j) & 1u) ? tf32_round_bits(v1[j]) : 0u;
              }
              tmem_st32(colY + col0, v0);
              tmem_st32(colY + col0 + 32, v1);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&a_ready[chain_id]);
            // dh1 = (dh2 W2) * mask1 (recomputed); dz += dh1 W1 over this thread's 64 hidden units
            mbar_wait(&acc_ready[chain_id], ph_acc);
            ph_acc ^= 1;
            tc_fence_after();
            {
              uint32_t v0[32], v1[32];
              tmem_ld32x2_sync(colX + col0, colX + col0 + 32, v0, v1);
              float2 ax = make_float2(0.f, 0.f), ay = make_float2(0.f, 0.f);
#pragma unroll
              for (int j = 0; j < 64; j += 2) {
                const int c = col0 + j;
                const float2 wx = *reinterpret_cast<const float2*>(sw + OFF_W1X + c);
                const float2 wy = *reinterpret_cast<const float2*>(sw + OFF_W1Y + c);
                const float2 bb = *reinterpret_cast<const float2*>(sw + OFF_B1 + c);
                const float2 h = ffma2(wy, zy2, ffma2(wx, zx2, bb));
                const uint32_t r0 = j < 32 ? v0[j] : v1[j - 32];
                const uint32_t r1 = j < 32 ? v0[j + 1] : v1[j - 31];
                const float2 dh = make_float2(h.x > 0.f ? __uint_as_float(r0) : 0.f, h.y > 0.f ? __uint_as_float(r1) : 0.f);
                ax = ffma2(dh, wx, ax);
                ay = ffma2(dh, wy, ay);
              }
              dzx += ax.x + ax.y;
              dzy += ay.x + ay.y;
            }
          }
          s.dzs[(chain_id * 2 + half) * 128 + row] = make_float2(dzx, dzy);
        }
        named_bar(3, EPI_THREADS);
        // ---- d(omega) += P^T dz, energy partials ----
        if (GRAD && chain_id == 0 && half == 0) {
          float P[MAX_KB];
          design_row(s.ts[row], n_poly, Kb, s.basis, P);
          const float2 d0 = s.dzs[row], d1 = s.dzs[128 + row], d2 = s.dzs[256 + row], d3 = s.dzs[384 + row];
          const float dx = (d0.x + d1.x) + (d2.x + d3.x), dy = (d0.y + d1.y) + (d2.y + d3.y);
#pragma unroll
          for (int k = 0; k < MAX_KB; ++k)
            if (k < Kb) {
              const float cx = warp_sum(P[k] * dx), cy = warp_sum(P[k] * dy);
              if (lane == 0) { s.red[(warp & 3) * 20 + 2 * k] = cx; s.red[(warp & 3) * 20 + 2 * k + 1] = cy; }
            }
        }
        if (t512 == 0) {
          float ee = 0.f, ll = 0.f;
          for (int w = 0; w < 16; ++w) { ee += s.red[80 + w]; ll += s.red[96 + w]; }
          e_tot += ee;
          l_tot += ll;
        }
        named_bar(3, EPI_THREADS);
        if (GRAD && t512 < 2 * Kb)
          s.gacc[t512] += (s.red[t512] + s.red[20 + t512]) + (s.red[40 + t512] + s.red[60 + t512]);
      }  // tiles

      named_bar(3, EPI_THREADS);
      if (t512 == 0) {
        const float E = e_tot / float(M);
        if (p.energy_trace) p.energy_trace[size_t(step) * p.N + n] = E;
        if (step == p.steps - 1) {
          if (p.energy_last) p.energy_last[n] = E;
          if (p.length_out) p.length_out[n] = l_tot / float(M);
        }
      }
      if (GRAD && t512 < 2 * Kb) {
        const int k = t512 >> 1, d = t512 & 1;
        const float tend = p.t[T - 1];
        float P[MAX_KB];
        design_row(tend, n_poly, Kb, s.basis, P);
        const float2 ze = spline_point(tend, n_poly, s.coef, pa, pb);
        const float err = d == 0 ? ze.x - pb.x : ze.y - pb.y;
        const float g = s.gacc[t512] + (2.0f * p.penalty_w) * err * P[k];
        AdamScalars sc = adam_scalars(p.step0 + step + 1, p.lr, p.beta1, p.beta2);
        float om = s.om[t512], mm = s.om[2 * MAX_KB + t512], vv = s.om[4 * MAX_KB + t512];
        adam_update(om, mm, vv, g, sc, p.one_minus_b1, p.beta2f, p.one_minus_b2, p.eps);
        s.om[t512] = om;
        s.om[2 * MAX_KB + t512] = mm;
        s.om[4 * MAX_KB + t512] = vv;
      }
      named_bar(3, EPI_THREADS);
    }  // steps

    cp_async_wait_all();
    if (GRAD && t512 < 2 * Kb) {
      p.omega[size_t(n) * 2 * Kb + t512] = s.om[t512];
      p.adam_m[size_t(n) * 2 * Kb + t512] = s.om[2 * MAX_KB + t512];
      p.adam_v[size_t(n) * 2 * Kb + t512] = s.om[4 * MAX_KB + t512];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

size_t tc_workspace_bytes(int, int, int, int) { return 0; }

cudaError_t launch_tc(const StepParams& p, bool grad, cudaStream_t stream) {
  if (p.precision != 1) return cudaErrorNotSupported;  // 3xTF32 not built yet
  const size_t smem = tc_smem_bytes(p.M, p.K);
  if (smem > 232448) return cudaErrorNotSupported;
  cudaError_t e;
  if (grad) {
    e = cudaFuncSetAttribute(tc_curve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    tc_curve_kernel<true><<<p.N, TC_THREADS, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(tc_curve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    tc_curve_kernel<false><<<p.N, TC_THREADS, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace vlg
