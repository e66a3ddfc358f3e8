// Flash attention forward, second generation, for sm_100a (head_dim 128, GQA-native; causal with optional sliding
// window, or full attention).  This is the kernel on the training path (reference call site: the NKI flash kernel at
// src/neuronx_distributed_training/models/hf_models/modeling_llama.py:482-489).
//
// One CTA = 256 query rows (two 128-row tiles) of one (batch, head).  The two tiles ping-pong over the tensor core: while
// the softmax warps of one tile work on S_j, the MMA warp computes the other tile's scores / P·V products.
//
//   warps 0..3   softmax + epilogue of tile 0   (thread r owns query row r = TMEM lane r)
//   warps 4..7   softmax + epilogue of tile 1
//   warp  8      TMA producer: Q0, Q1, then K_j / V_j tiles (128 keys each) through ONE five-slot ring
//                (order K_0 V_0 K_1 V_1 ...; every tile is used by both query tiles before its slot is released)
//   warp  9      tcgen05 issuer          (warps 10, 11 only pad the third warpgroup for setmaxnreg)
//
// Tensor memory (512 columns): [0,128) S0, [128,256) S1, [256,384) O0, [384,512) O1.  P never touches shared memory:
// the softmax thread overwrites the first 64 columns of its own S row with the bf16-packed probabilities
// (tcgen05.st) and the P·V product reads its A operand straight from tensor memory (tcgen05.mma with [tmem] A).
// Since QK(t, j+1) is issued after PV(t, j) and tcgen05.mma executes in issue order, "S_{j+1} is ready" also means
// "PV_j has retired": the softmax thread may rescale O in place (LAZY rescaling: only when the running maximum grows
// by more than 2^8, which is exact after the final 1/l normalisation) without any extra barrier.
// The softmax inner loop uses the packed fp32 pipe (FFMA2 / FADD2), three-input maxima (FMNMX3) and ex2.approx.
//
// Output: O [B,S,Hq,D] bf16 and LSE [B,Hq,S] fp32 (natural log), consumed by the backward kernels.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "sm100_ptx.cuh"
#include "tma_utils.h"

namespace fa2 {

using namespace sm100;

constexpr int BM = 128;          // query rows per tile (two tiles per CTA)
constexpr int BN = 128;          // keys per block
constexpr int D = 128;           // head dim
constexpr int TILE_BYTES = 128 * 128 * 2;     // one [128 x 128] bf16 tile = two 16 KB swizzle chunks
constexpr int CHUNK_BYTES = 128 * 64 * 2;     // [128 rows x 64 cols] bf16
constexpr int KV_SLOTS = 5;
constexpr int NUM_THREADS = 384;           // three warpgroups: softmax tile 0, softmax tile 1, {TMA, MMA, 2 idle warps}
constexpr int SMEM_BYTES = (2 + KV_SLOTS) * TILE_BYTES + 1024 + 256;
constexpr float RESCALE_THRESHOLD = 8.0f;

struct Params {
  int B, S, Hq, Hk;      // S = number of (local) queries per batch entry
  int Sk;                // number of keys per batch entry (>= S under context parallelism: the gathered sequence)
  int q_off;             // absolute position of query row 0 relative to key 0 (multiple of 128; context-parallel rank offset)
  int num_q_pairs;       // S / 256
  int window;            // 0 = none; otherwise keys with k <= q - window are masked
  int causal;            // 1: keys after the query are masked; 0: full attention
  float scale_log2;      // softmax scale * log2(e)
  float* lse;            // [B, Hq, S]
  long long* dbg;        // optional in-kernel timeline of CTA (0, 0): [3 roles][128 blocks][4 stamps] of clock64 (else nullptr)
};

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int kk) {
  // K-major [128 rows x 128 cols] tile stored as two [128 x 64] swizzle chunks; kk = 16-column step (0..7)
  return make_smem_desc_sw128(tile_addr + (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int kk) {
  // the same physical tile read as an MN-major operand: rows are the K index, the two 64-column chunks are N groups
  return make_smem_desc_sw128(tile_addr + kk * (16 * 128), CHUNK_BYTES, 1024);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// packed fp32 pairs live in 64-bit registers (FFMA2 / FADD2 / FMUL2 operate on aligned register pairs)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_fwd2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                  const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [2] tiles; reused as the O staging buffers in the epilogue
  uint8_t* sKV = sQ + 2 * TILE_BYTES;               // [KV_SLOTS] tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + KV_SLOTS * TILE_BYTES);
  uint64_t* q_full = bars;                          // [2]
  uint64_t* kv_full = bars + 2;                     // [5]
  uint64_t* kv_empty = bars + 7;                    // [5]
  uint64_t* s_full = bars + 12;                     // [2]  S_t(j) is in TMEM (and PV_t(j-1) has retired)
  uint64_t* p_full = bars + 14;                     // [2]  P_t(j) is in TMEM, O_t has been rescaled if necessary
  uint64_t* o_full = bars + 16;                     // [2]  last PV of tile t has retired
  uint64_t* turn = bars + 18;                       // [2]  exp-phase hand-over between the two softmax warpgroups
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);

  // heavy (late) query blocks first; heads vary fastest so that concurrently running CTAs have equal lengths and the
  // heads of one GQA group share their K/V tiles in L2
  const int qp = p.num_q_pairs - 1 - (int)blockIdx.y;
  const int head = blockIdx.x % p.Hq, batch = blockIdx.x / p.Hq;
  const int kv_head = head / (p.Hq / p.Hk);
  const int q0 = qp * 2 * BM;
  const int qa0 = p.q_off + q0;                                        // absolute position of the CTA's first query
  const int j_lo = (p.window > 0) ? max(0, (qa0 - p.window + 1) / BN) : 0;
  const int j_hi = p.causal ? qa0 / BN + 1 : p.Sk / BN - 1;            // inclusive, for tile 1
  const int n = j_hi - j_lo + 1;                                       // key blocks of tile 1
  const int n0 = p.causal ? n - 1 : n;                                 // key blocks of tile 0 (its diagonal is one earlier)
  const int row_base = batch * p.S;                                    // row of query 0 of this batch in the [B*S, Hq*D] views
  const int krow_base = batch * p.Sk;                                  // row of key 0 of this batch in the [B*Sk, Hk*D] views

  if (warp_idx == 8 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&turn[i], 4);
    }
    for (int i = 0; i < KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 9) {
    tmem_alloc<1>(tmem_ptr, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // register reallocation (the launch gives every thread 168): the producer / issuer warpgroup keeps 80, the softmax
  // warpgroups (one fp32 score row of 128 values per thread) take 208
  if (warp_idx == 8) {
    // =========================== TMA producer ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (elect_one()) {
      int u = 0;
      auto load_kv = [&](const CUtensorMap* map, int j) {
        const int slot = u % KV_SLOTS;
        const uint32_t ph = (u / KV_SLOTS) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        mbar_expect_tx(&kv_full[slot], TILE_BYTES);
        uint8_t* dst = sKV + slot * TILE_BYTES;
        tma_load_2d(dst, map, &kv_full[slot], kv_head * D, krow_base + j * BN);
        tma_load_2d(dst + CHUNK_BYTES, map, &kv_full[slot], kv_head * D + 64, krow_base + j * BN);
        ++u;
      };
      auto load_q = [&](int t) {
        mbar_expect_tx(&q_full[t], TILE_BYTES);
        tma_load_2d(sQ + t * TILE_BYTES, &tma_q, &q_full[t], head * D, row_base + q0 + t * BM);
        tma_load_2d(sQ + t * TILE_BYTES + CHUNK_BYTES, &tma_q, &q_full[t], head * D + 64, row_base + q0 + t * BM);
      };
      load_q(0);
      load_kv(&tma_k, j_lo);
      load_q(1);
      load_kv(&tma_v, j_lo);
      for (int i = 1; i < n; ++i) {
        load_kv(&tma_k, j_lo + i);
        load_kv(&tma_v, j_lo + i);
      }
    }
  } else if (warp_idx == 9) {
    // =========================== MMA issuer ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    constexpr uint32_t idesc_qk = make_idesc_bf16(BM, BN, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D, false, true);
    const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV);
    // ring position of K_i is 2i, of V_i is 2i + 1
    auto wait_kv = [&](int u) { mbar_wait(&kv_full[u % KV_SLOTS], (u / KV_SLOTS) & 1); };
    auto issue_qk = [&](int t, int i, bool release_k) {
      const int u = 2 * i;
      wait_kv(u);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t k_addr = kv_addr + (u % KV_SLOTS) * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16<1>(tmem_base + t * BN, desc_kmajor(q_addr + t * TILE_BYTES, kk), desc_kmajor(k_addr, kk), idesc_qk,
                      kk != 0);
        umma_commit(&s_full[t]);
        if (release_k) umma_commit(&kv_empty[u % KV_SLOTS]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int i, bool release_v, bool last) {
      const int u = 2 * i + 1;
      mbar_wait(&p_full[t], i & 1);
      const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && i < 128;
      if (stamp) p.dbg[(2 * 128 + i) * 4 + 2 * t] = clock64();        // MMA warp saw P_t(i)
      wait_kv(u);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = kv_addr + (u % KV_SLOTS) * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          umma_f16_ts(tmem_base + 256 + t * D, tmem_base + t * BN + kk * 8, desc_mnmajor(v_addr, kk), idesc_pv,
                      (i | kk) != 0);
        if (release_v) umma_commit(&kv_empty[u % KV_SLOTS]);
        if (last) umma_commit(&o_full[t]);
      }
      __syncwarp();
      if (stamp) p.dbg[(2 * 128 + i) * 4 + 2 * t + 1] = clock64();    // PV_t(i) issued
    };
    mbar_wait(&q_full[0], 0);
    issue_qk(0, 0, false);
    mbar_wait(&q_full[1], 0);
    issue_qk(1, 0, true);
    for (int i = 0; i < n; ++i) {
      if (i < n0) {
        issue_pv(0, i, false, i + 1 == n0);
        if (i + 1 < n0) issue_qk(0, i + 1, false);
      }
      issue_pv(1, i, true, i + 1 == n);
      if (i + 1 < n) issue_qk(1, i + 1, true);
    }
  } else if (warp_idx >= 10) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
  } else {
    // =========================== softmax + epilogue (thread = query row) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = warp_idx >> 2;                     // query tile
    const uint32_t qd = warp_idx & 3;                // TMEM lane quadrant
    const uint32_t row = qd * 32 + lane;
    const uint32_t lane_off = (qd * 32u) << 16;
    const uint32_t tS = tmem_base + lane_off + t * BN;
    const uint32_t tO = tmem_base + lane_off + 256 + t * D;
    const int q_pos = qa0 + t * BM + (int)row;        // absolute position (mask arithmetic)
    const int nt = t == 0 ? n0 : n;
    const int j_diag = qa0 / BN + t;
    const float sc = p.scale_log2;
    float m_used = -INFINITY;                        // raw (unscaled) score domain
    uint64_t l2 = pack2(0.f, 0.f);                   // two partial row sums
    for (int i = 0; i < nt; ++i) {
      const int j = j_lo + i;
      mbar_wait(&s_full[t], i & 1);
      tcgen05_fence_after();
      const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && qd == 0 && i < 128;
      if (stamp) p.dbg[(t * 128 + i) * 4 + 0] = clock64();            // S_t(i) observed
      uint32_t sr[4][32];                            // the fp32 score row (raw bits), 4 x 32 columns
      tmem_ld_32x32b_x32(tS + 0, sr[0]);
      tmem_ld_32x32b_x32(tS + 32, sr[1]);
      tmem_ld_32x32b_x32(tS + 64, sr[2]);
      tmem_ld_32x32b_x32(tS + 96, sr[3]);
      tmem_ld_wait();
      const bool need_mask = (p.causal && j == j_diag) || (p.window > 0 && j * BN <= q_pos - p.window);
      if (need_mask) {
        const int k0 = j * BN;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int kpos = k0 + c * 32 + e;
            if ((p.causal && kpos > q_pos) || (p.window > 0 && kpos <= q_pos - p.window)) sr[c][e] = 0xff800000u;  // -inf
          }
      }
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a0 = fmax3(__uint_as_float(sr[c][0]), __uint_as_float(sr[c][1]), __uint_as_float(sr[c][2]));
        float a1 = fmax3(__uint_as_float(sr[c][3]), __uint_as_float(sr[c][4]), __uint_as_float(sr[c][5]));
#pragma unroll
        for (int e = 6; e + 4 <= 32; e += 4) {
          a0 = fmax3(a0, __uint_as_float(sr[c][e]), __uint_as_float(sr[c][e + 1]));
          a1 = fmax3(a1, __uint_as_float(sr[c][e + 2]), __uint_as_float(sr[c][e + 3]));
        }
        // columns 30, 31 are left by the loop above (6 + 4*6 = 30)
        mx[c] = fmax3(a0, a1, fmaxf(__uint_as_float(sr[c][30]), __uint_as_float(sr[c][31])));
      }
      const float m_new = fmaxf(fmax3(m_used, mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const bool grow = (m_new - m_used) * sc > RESCALE_THRESHOLD;      // true on the first block (m_used = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float alpha = (m_used == -INFINITY) ? 0.f : ex2((m_used - m_new) * sc);
        const uint64_t a2 = pack2(alpha, alpha);
        l2 = fmul2(l2, a2);
        m_used = m_new;
        if (i > 0) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tO + c * 32, v);
            tmem_ld_wait();
            uint32_t w0[16], w1[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float lo, hi;
              unpack2(fmul2(pack2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), a2), lo, hi);
              w0[2 * e] = __float_as_uint(lo); w0[2 * e + 1] = __float_as_uint(hi);
              unpack2(fmul2(pack2(__uint_as_float(v[16 + 2 * e]), __uint_as_float(v[16 + 2 * e + 1])), a2), lo, hi);
              w1[2 * e] = __float_as_uint(lo); w1[2 * e + 1] = __float_as_uint(hi);
            }
            tmem_st_32x32b_x16(tO + c * 32, w0);
            tmem_st_32x32b_x16(tO + c * 32 + 16, w1);
          }
        }
      }
      // The exponentials of the two query tiles take turns (tile 0 block i, tile 1 block i, tile 0 block i+1, ...), the
      // order in which the MMA warp serves them.  Without the hand-over both warpgroups share the MUFU pipe, the
      // higher-numbered warps win the issue arbitration and overtake tile 0, and the tensor core -- which needs P0 first
      // -- idles until the preempted tile finishes.  Loading S and the row maximum above stay outside the turn.
      if (t == 0) {
        if (i > 0) mbar_wait(&turn[0], (i - 1) & 1);
      } else if (i < n0) {
        mbar_wait(&turn[1], i & 1);
      }
      if (stamp) p.dbg[(t * 128 + i) * 4 + 1] = clock64();            // turn acquired: exponentials start
      // a row can still be fully masked so far (sliding window): keep exp2(-inf - 0) = 0 instead of NaN
      const float m_sc = (m_used == -INFINITY) ? 0.f : m_used * sc;
      const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_sc, -m_sc);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float x0, x1;
          unpack2(ffma2(pack2(__uint_as_float(sr[c][2 * e]), __uint_as_float(sr[c][2 * e + 1])), sc2, nm2), x0, x1);
          x0 = ex2(x0);
          x1 = ex2(x1);
          l2 = fadd2(l2, pack2(x0, x1));
          pk[e] = pack_bf16x2(x0, x1);
        }
        tmem_st_32x32b_x16(tS + c * 16, pk);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&turn[t ^ 1]);
      if (stamp) p.dbg[(t * 128 + i) * 4 + 2] = clock64();            // exponentials issued, turn handed over
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (stamp) p.dbg[(t * 128 + i) * 4 + 3] = clock64();            // P_t(i) published
    }
    // ---------------- epilogue: O / l -> bf16 -> smem (the Q tile is free now) -> TMA store; LSE ----------------
    float l_lo, l_hi;
    unpack2(l2, l_lo, l_hi);
    const float l = l_lo + l_hi;
    mbar_wait(&o_full[t], 0);
    tcgen05_fence_after();
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    const uint32_t orow = smem_u32(sQ) + t * TILE_BYTES + row * 128;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tO + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int t8 = 0; t8 < 4; ++t8) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[t8 * 8 + 0]) * inv_l, __uint_as_float(v[t8 * 8 + 1]) * inv_l);
        o.y = pack_bf16x2(__uint_as_float(v[t8 * 8 + 2]) * inv_l, __uint_as_float(v[t8 * 8 + 3]) * inv_l);
        o.z = pack_bf16x2(__uint_as_float(v[t8 * 8 + 4]) * inv_l, __uint_as_float(v[t8 * 8 + 5]) * inv_l);
        o.w = pack_bf16x2(__uint_as_float(v[t8 * 8 + 6]) * inv_l, __uint_as_float(v[t8 * 8 + 7]) * inv_l);
        const int c16 = c * 4 + t8;      // 16-byte column group 0..15
        const int chunk = c16 >> 3, j16 = c16 & 7;
        st_shared_v4(orow + chunk * CHUNK_BYTES + ((j16 ^ (row & 7)) * 16), o);
      }
    }
    p.lse[((size_t)batch * p.Hq + head) * p.S + q0 + t * BM + row] =
        l > 0.f ? (m_used * sc + log2f(l)) * 0.6931471805599453f : -INFINITY;
    tcgen05_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
    if ((threadIdx.x & 127) == 0) {
      tma_store_2d(&tma_o, sQ + t * TILE_BYTES, head * D, row_base + q0 + t * BM);
      tma_store_2d(&tma_o, sQ + t * TILE_BYTES + CHUNK_BYTES, head * D + 64, row_base + q0 + t * BM);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 9) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace fa2

static long long* g_fwd2_dbg = nullptr;
// int64 tensor of >= 3 * 128 * 4 elements on the device: the next launches record the timeline of CTA (0, 0); empty -> off
void flash_fwd2_set_debug(const at::Tensor& t) {
  g_fwd2_dbg = t.numel() >= 3 * 128 * 4 ? reinterpret_cast<long long*>(t.data_ptr<int64_t>()) : nullptr;
}

// q [B,S,Hq,128], k/v [B,Sk,Hk,128] bf16 contiguous, S a multiple of 256, Sk >= q_offset + S (context parallelism: the
// gathered keys; query i sits at absolute position q_offset + i); returns (out [B,S,Hq,128] bf16, lse [B,Hq,S] fp32)
std::vector<at::Tensor> flash_attn_fwd2(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, bool causal,
                                        int64_t window, double scale, int64_t q_offset) {
  using namespace fa2;
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && k.scalar_type() == at::kBFloat16 && v.scalar_type() == at::kBFloat16);
  TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4 && q.size(3) == D && k.size(3) == D && v.size(3) == D,
              "flash_attn_fwd2: head_dim must be 128");
  TORCH_CHECK(causal || window == 0, "flash_attn_fwd2: a sliding window needs causal attention");
  TORCH_CHECK(q.is_contiguous() && k.is_contiguous() && v.is_contiguous(), "flash_attn_fwd2: contiguous [B,S,H,D] inputs required");
  const int B = q.size(0), S = q.size(1), Hq = q.size(2), Hk = k.size(2);
  const int Sk = k.size(1);
  TORCH_CHECK(v.size(1) == Sk && v.size(2) == Hk && Hq % Hk == 0 && k.size(0) == B && v.size(0) == B);
  TORCH_CHECK(S % (2 * BM) == 0, "flash_attn_fwd2: the query length must be a multiple of 256");
  TORCH_CHECK(Sk % BN == 0 && q_offset % BN == 0 && q_offset >= 0 && q_offset + S <= Sk,
              "flash_attn_fwd2: keys must cover [0, q_offset + S) in whole blocks of 128");
  c10::cuda::CUDAGuard guard(q.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto out = at::empty_like(q);
  auto lse = at::empty({B, Hq, S}, q.options().dtype(at::kFloat));
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tq = tma::make_2d(q.data_ptr(), BF, 2, (uint64_t)B * S, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, 128, 64);
  CUtensorMap tk = tma::make_2d(k.data_ptr(), BF, 2, (uint64_t)B * Sk, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tv = tma::make_2d(v.data_ptr(), BF, 2, (uint64_t)B * Sk, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap to = tma::make_2d(out.data_ptr(), BF, 2, (uint64_t)B * S, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, 128, 64);
  Params p;
  p.B = B; p.S = S; p.Sk = Sk; p.q_off = (int)q_offset; p.Hq = Hq; p.Hk = Hk; p.num_q_pairs = S / (2 * BM); p.window = (int)window;
  p.causal = causal ? 1 : 0;
  p.scale_log2 = (float)(scale * 1.4426950408889634);
  p.lse = lse.data_ptr<float>();
  p.dbg = g_fwd2_dbg;
  static bool once = false;
  if (!once) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  dim3 grid(Hq * B, p.num_q_pairs);
  flash_fwd2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {out, lse};
}
