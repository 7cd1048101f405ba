// tcgen05 (TF32) geodesic step kernel -- the tensor-core variant (<=1e-3 relative on lengths).
//
// One CTA (320 threads) = one curve, persistent over `steps` Adam steps.  The two 128-wide
// decoder layers and their transposes run as tcgen05.mma kind::tf32 with
//   * M = 128 curve points = the 128 TMEM lanes (one thread owns one point / one lane),
//   * the A operand (activations) living in TENSOR MEMORY: the epilogue threads write the
//     next layer's input back with tcgen05.st, in place of the accumulator they just read,
//   * the B operand (weights) streamed from L2 into a shared-memory ring by the TMA engine
//     (1-D bulk copies of pre-packed no-swizzle K-major images, mbarrier complete_tx),
//   * fp32 accumulators in TMEM, read back with tcgen05.ld.
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer (one lane), warps 2-5 and 6-9 =
// two epilogue warpgroups.  Each warpgroup owns a "chain" of 256 TMEM columns and alternate
// decoders, so that one chain's CUDA-core epilogue overlaps the other chain's MMAs.
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

constexpr int TC_THREADS = 320;
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

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

struct TcSmem {
  unsigned char* ring;  // NSTAGES * 16 KB
  float* Diff;          // M*128*52
  uint32_t* mask2;      // K*128*4 words
  uint8_t* sel;         // MAX_M*2*128
  float* sw;            // 2 * 576
  float2* zs;           // 128
  float* ts;            // 128
  float2* dzs;          // 2*128
  float* coef;          // 64
  float* basis;         // 288
  float* om;            // 56
  float* gacc;          // 20
  float* red;           // 4*20 + 16
  uint64_t* bars;       // full[NSTAGES], empty[NSTAGES], a_ready[2], acc_ready[2]
  uint32_t* tmem_base;
  volatile int* turn;
};

__device__ __forceinline__ TcSmem tc_carve(unsigned char* base, int M, int K) {
  TcSmem s;
  s.ring = base;
  float* f = reinterpret_cast<float*>(base + NSTAGES * STAGE_BYTES);
  s.Diff = f; f += M * 128 * DIFF_STRIDE;
  s.sw = f; f += 2 * 576;
  s.zs = reinterpret_cast<float2*>(f); f += 256;
  s.ts = f; f += 128;
  s.dzs = reinterpret_cast<float2*>(f); f += 512;
  s.coef = f; f += 64;
  s.basis = f; f += 4 * MAX_NPOLY * MAX_KB;
  s.om = f; f += 3 * 2 * MAX_KB + 2;
  s.gacc = f; f += 2 * MAX_KB + 2;
  s.red = f; f += 96;
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
  size_t fl = size_t(M) * 128 * DIFF_STRIDE + 2 * 576 + 256 + 128 + 512 + 64 + 4 * MAX_NPOLY * MAX_KB +
              (3 * 2 * MAX_KB + 2) + (2 * MAX_KB + 2) + 96 + 2 * (2 * NSTAGES + 4) + 2 + 2 + MAX_M * 2 * 128 / 4;
  return size_t(NSTAGES) * STAGE_BYTES + fl * 4 + size_t(K) * 128 * 16;
}

template <bool GRAD>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_curve_kernel(StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x;
  const int M = p.M, K = p.K, T = p.T, n_poly = p.n_poly, Kb = p.Kb, X = p.X;
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
    mbar_init(&a_ready[0], 128);
    mbar_init(&a_ready[1], 128);
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
  // lock step through the ring; the epilogue warpgroups follow through a_ready / acc_ready.
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
    // ================= epilogue warpgroups =================
    const int wg = (warp - 2) >> 2;             // chain
    const int row = (warp & 3) * 32 + lane;     // TMEM lane = curve point of the tile
    const int t256 = wg * 128 + row;            // 0..255 over both warpgroups
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    const uint32_t chain = tmem + lane_addr + uint32_t(wg) * 256u;
    const uint32_t colX = chain, colY = chain + 128u;
    float* sw = s.sw + wg * 576;
    uint32_t ph_acc = 0;
    const float coefm = 2.0f / float(M);

    for (int i = t256; i < 4 * n_poly * Kb; i += 256) s.basis[i] = p.basis[i];
    if (t256 < 2 * Kb) {
      s.om[t256] = p.omega[size_t(n) * 2 * Kb + t256];
      if (GRAD) {
        s.om[2 * MAX_KB + t256] = p.adam_m[size_t(n) * 2 * Kb + t256];
        s.om[4 * MAX_KB + t256] = p.adam_v[size_t(n) * 2 * Kb + t256];
      }
    }
    const float2 pa = make_float2(p.a[2 * n], p.a[2 * n + 1]);
    const float2 pb = make_float2(p.b[2 * n], p.b[2 * n + 1]);
    named_bar(3, 256);

    for (int step = 0; step < p.steps; ++step) {
      if (t256 < 8 * n_poly) {
        const int r = t256 >> 1, d = t256 & 1;
        float acc = 0.f;
        for (int k = 0; k < Kb; ++k) acc = fmaf(s.basis[r * Kb + k], s.om[2 * k + d], acc);
        s.coef[t256] = acc;
      }
      if (t256 < 2 * MAX_KB) s.gacc[t256] = 0.f;
      float e_tot = 0.f, l_tot = 0.f;  // meaningful in t256 == 0
      named_bar(3, 256);

      for (int tile = 0; tile < ntiles; ++tile) {
        const int seg0 = tile * TILE_SEGS;
        const int nseg = min(TILE_SEGS, T - 1 - seg0);
        const int turn0 = (step * ntiles + tile) * K;
        // ---- tile setup ----
        if (wg == 0) {
          const int ti = min(seg0 + row, T - 1);
          const float t = p.t[ti];
          s.ts[row] = t;
          s.zs[row] = spline_point(t, n_poly, s.coef, pa, pb);
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
        for (int i = t256; i < M * 128 * DIFF_STRIDE / 4; i += 256)
          reinterpret_cast<float4*>(s.Diff)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float dzx = 0.f, dzy = 0.f;
        named_bar(3, 256);
        const float2 z = s.zs[row];

        // =============================== forward ===============================
        for (int k = wg; k < K; k += 2) {
          const float* dec = dec_ptr(p.packed, k);
          for (int i = row; i < 576; i += 128) sw[i] = __ldg(dec + i);
          named_bar(1 + wg, 128);
          // layer 1 (CUDA cores, fp32) -> A1 in X
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = c0 + j;
              const float h = fmaf(sw[OFF_W1 + 2 * c + 1], z.y, fmaf(sw[OFF_W1 + 2 * c], z.x, sw[OFF_B1 + c]));
              v[j] = to_tf32(fmaxf(h, 0.f));
            }
            tmem_st32(colX + c0, v);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&a_ready[wg]);
          // layer 2 epilogue: D2 (Y) -> relu(+b2) -> A2 (Y, in place), mask bits
          mbar_wait(&acc_ready[wg], ph_acc);
          ph_acc ^= 1;
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(colY + c0, v);
            tmem_wait_ld();
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float h = __uint_as_float(v[j]) + sw[OFF_B2 + c0 + j];
              if (h > 0.f) bits |= 1u << j;
              v[j] = to_tf32(fmaxf(h, 0.f));
            }
            tmem_st32(colY + c0, v);
            if (GRAD) s.mask2[(k * 128 + row) * 4 + (c0 >> 5)] = bits;
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&a_ready[wg]);
          // layer 3 epilogue: D3 (X[0:64]) + b3 -> Diff
          mbar_wait(&acc_ready[wg], ph_acc);
          ph_acc ^= 1;
          tc_fence_after();
          uint32_t x0[32], x1[32];
          tmem_ld32(colX, x0);
          tmem_ld32(colX + 32, x1);
          tmem_wait_ld();
          // Diff is shared by both warpgroups: updates are serialised in decoder order
          if ((row & 127) == 0) {
            while (*s.turn != turn0 + k) {
            }
            __threadfence_block();
          }
          named_bar(1 + wg, 128);
          // role 0: this point is the left end of its segment
          for (int m = 0; m < M; ++m)
            if (s.sel[(m * 2 + 0) * 128 + row] == k) {
              float* d = s.Diff + (m * 128 + row) * DIFF_STRIDE;
#pragma unroll
              for (int j = 0; j < 32; ++j) d[j] -= __uint_as_float(x0[j]) + sw[OFF_B3 + j];
#pragma unroll
              for (int j = 0; j < 20; ++j)
                if (32 + j < X) d[32 + j] -= __uint_as_float(x1[j]) + sw[OFF_B3 + 32 + j];
            }
          named_bar(1 + wg, 128);
          // role 1: this point is the right end of the previous segment
          if (row >= 1)
            for (int m = 0; m < M; ++m)
              if (s.sel[(m * 2 + 1) * 128 + row - 1] == k) {
                float* d = s.Diff + (m * 128 + row - 1) * DIFF_STRIDE;
#pragma unroll
                for (int j = 0; j < 32; ++j) d[j] += __uint_as_float(x0[j]) + sw[OFF_B3 + j];
#pragma unroll
                for (int j = 0; j < 20; ++j)
                  if (32 + j < X) d[32 + j] += __uint_as_float(x1[j]) + sw[OFF_B3 + 32 + j];
              }
          __threadfence_block();
          named_bar(1 + wg, 128);
          if ((row & 127) == 0) *s.turn = turn0 + k + 1;
        }
        named_bar(3, 256);

        // =============================== energy ===============================
        {
          float e = 0.f, l = 0.f;
          for (int idx = t256; idx < M * 128; idx += 256) {
            const int r = idx & 127;
            if (r < nseg) {
              const float* d = s.Diff + idx * DIFF_STRIDE;
              float q = 0.f;
              for (int c = 0; c < X; ++c) q = fmaf(d[c], d[c], q);
              e += q;
              l += sqrtf(q);
            }
          }
          e = warp_sum(e);
          l = warp_sum(l);
          if (lane == 0) { s.red[80 + (warp - 2)] = e; s.red[88 + (warp - 2)] = l; }
        }

        if (GRAD) {
          // =============================== backward ===============================
          for (int k = wg; k < K; k += 2) {
            const float* dec = dec_ptr(p.packed, k);
            named_bar(1 + wg, 128);  // previous item's readers of sw are done
            for (int i = row; i < 384; i += 128) sw[i] = __ldg(dec + i);
            // G = dE/dx_k (this point) -> X[0:64]
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
              float g[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = 0.f;
              for (int m = 0; m < M; ++m) {
                if (row >= 1 && s.sel[(m * 2 + 1) * 128 + row - 1] == k) {
                  const float* d = s.Diff + (m * 128 + row - 1) * DIFF_STRIDE + c0;
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (c0 + j < DIFF_STRIDE) g[j] += d[j];
                }
                if (s.sel[(m * 2 + 0) * 128 + row] == k) {
                  const float* d = s.Diff + (m * 128 + row) * DIFF_STRIDE + c0;
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (c0 + j < DIFF_STRIDE) g[j] -= d[j];
                }
              }
              uint32_t v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (c0 + j < X) ? to_tf32(coefm * g[j]) : 0u;
              tmem_st32(colX + c0, v);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&a_ready[wg]);
            named_bar(1 + wg, 128);  // sw visible
            // dh2 = (G W3) * mask2 -> A4 (Y, in place)
            mbar_wait(&acc_ready[wg], ph_acc);
            ph_acc ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(colY + c0, v);
              tmem_wait_ld();
              const uint32_t bits = s.mask2[(k * 128 + row) * 4 + (c0 >> 5)];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? to_tf32(__uint_as_float(v[j])) : 0u;
              tmem_st32(colY + c0, v);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&a_ready[wg]);
            // dh1 = (dh2 W2) * mask1 (recomputed); dz += dh1 W1
            mbar_wait(&acc_ready[wg], ph_acc);
            ph_acc ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(colX + c0, v);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;
                const float wa = sw[OFF_W1 + 2 * c], wb = sw[OFF_W1 + 2 * c + 1];
                const float h = fmaf(wb, z.y, fmaf(wa, z.x, sw[OFF_B1 + c]));
                if (h > 0.f) {
                  dzx = fmaf(__uint_as_float(v[j]), wa, dzx);
                  dzy = fmaf(__uint_as_float(v[j]), wb, dzy);
                }
              }
            }
          }
          s.dzs[wg * 128 + row] = make_float2(dzx, dzy);
        }
        named_bar(3, 256);
        // ---- d(omega) += P^T dz, energy partials ----
        if (GRAD && wg == 0) {
          float P[MAX_KB];
          design_row(s.ts[row], n_poly, Kb, s.basis, P);
          const float2 d0 = s.dzs[row], d1 = s.dzs[128 + row];
          const float dx = d0.x + d1.x, dy = d0.y + d1.y;
#pragma unroll
          for (int k = 0; k < MAX_KB; ++k)
            if (k < Kb) {
              const float cx = warp_sum(P[k] * dx), cy = warp_sum(P[k] * dy);
              if (lane == 0) { s.red[(warp & 3) * 20 + 2 * k] = cx; s.red[(warp & 3) * 20 + 2 * k + 1] = cy; }
            }
        }
        if (t256 == 0) {
          float ee = 0.f, ll = 0.f;
          for (int w = 0; w < 8; ++w) { ee += s.red[80 + w]; ll += s.red[88 + w]; }
          e_tot += ee;
          l_tot += ll;
        }
        named_bar(3, 256);
        if (GRAD && t256 < 2 * Kb)
          s.gacc[t256] += (s.red[t256] + s.red[20 + t256]) + (s.red[40 + t256] + s.red[60 + t256]);
      }  // tiles

      named_bar(3, 256);
      if (t256 == 0) {
        const float E = e_tot / float(M);
        if (p.energy_trace) p.energy_trace[size_t(step) * p.N + n] = E;
        if (step == p.steps - 1) {
          if (p.energy_last) p.energy_last[n] = E;
          if (p.length_out) p.length_out[n] = l_tot / float(M);
        }
      }
      if (GRAD && t256 < 2 * Kb) {
        const int k = t256 >> 1, d = t256 & 1;
        const float tend = p.t[T - 1];
        float P[MAX_KB];
        design_row(tend, n_poly, Kb, s.basis, P);
        const float2 ze = spline_point(tend, n_poly, s.coef, pa, pb);
        const float err = d == 0 ? ze.x - pb.x : ze.y - pb.y;
        const float g = s.gacc[t256] + (2.0f * p.penalty_w) * err * P[k];
        AdamScalars sc = adam_scalars(p.step0 + step + 1, p.lr, p.beta1, p.beta2);
        float om = s.om[t256], mm = s.om[2 * MAX_KB + t256], vv = s.om[4 * MAX_KB + t256];
        adam_update(om, mm, vv, g, sc, p.one_minus_b1, p.beta2f, p.one_minus_b2, p.eps);
        s.om[t256] = om;
        s.om[2 * MAX_KB + t256] = mm;
        s.om[4 * MAX_KB + t256] = vv;
      }
      named_bar(3, 256);
    }  // steps

    if (GRAD && t256 < 2 * Kb) {
      p.omega[size_t(n) * 2 * Kb + t256] = s.om[t256];
      p.adam_m[size_t(n) * 2 * Kb + t256] = s.om[2 * MAX_KB + t256];
      p.adam_v[size_t(n) * 2 * Kb + t256] = s.om[4 * MAX_KB + t256];
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
