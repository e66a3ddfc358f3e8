// Flash attention forward, second generation, for sm_100a (head_dim 128, GQA-native; causal with optional sliding
// window, or full attention).  This is the kernel on the training path (reference call site: the NKI flash kernel at
// src/neuronx_distributed_training/models/hf_models/modeling_llama.py:482-489).
//
// One CTA = 256 query rows (two 128-row tiles) of one (batch, head), walking over the keys in blocks of 64.
//
//   warps 0..3   softmax + epilogue of tile 0   (thread r owns query row r = TMEM lane r)
//   warps 4..7   softmax + epilogue of tile 1
//   warp  8      TMA producer: Q0, Q1, then K_j / V_j tiles (128 keys each) through ONE five-slot ring
//                (order K_0 V_0 K_1 V_1 ...; every tile is used by both query tiles before its slot is released)
//   warps 9, 10  tcgen05 issuers, one per query tile   (warp 11 only pads the third warpgroup for setmaxnreg)
//
// Tensor memory (512 columns): per tile t two score buffers S_t[0], S_t[1] of 64 columns ([128 t, 128 t + 128)) and the
// output accumulator O_t ([256 + 128 t, ...)).  P never touches shared memory: the softmax thread overwrites the first
// 32 columns of the score buffer it just read with the bf16-packed probabilities (tcgen05.st) and the P·V product reads
// its A operand straight from tensor memory (tcgen05.mma with [tmem] A).
// Because the scores are DOUBLE BUFFERED, QK(t, i+2) only has to follow PV(t, i) (which frees the buffer), so the tensor
// core computes the scores of the next two key blocks while the softmax warps are still exponentiating block i: the
// serial chain "softmax -> P·V -> next QK^T -> softmax" of a single-buffer design (measured with the in-kernel timeline:
// 1450 of 3450 cycles per block pair) is off the critical path.  Each query tile has its own issuer thread, so neither
// tile ever waits for the other's probabilities.  A quarter of the exponentials run as a cubic polynomial on the FMA
// pipe (exp2_poly2) to take load off the MUFU unit, which both softmax warpgroups share.
// LAZY rescaling: O is only rescaled when the running maximum grows by more than 2^8 (exact after the final 1/l
// normalisation); the rescaling thread first waits for the previous P·V of its tile (pv_done).
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
constexpr int BN = 128;          // keys per K / V tile in shared memory
constexpr int BK = 64;           // keys per softmax block (half a tile)
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
  int num_q_units;       // S / (128 * tiles per CTA)
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

// exp2 of a packed pair on the FMA / ALU pipes (Cody-Waite range reduction + cubic minimax polynomial on [-0.5, 0.5],
// max relative error 7.7e-5 -- far below the bf16 rounding of P).  Inputs are clamped at -126 (also maps -inf to ~0).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& r0, float& r1) {
  float a, b;
  unpack2(x2, a, b);
  x2 = pack2(fmaxf(a, -126.f), fmaxf(b, -126.f));
  const uint64_t magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f);
  const uint64_t one = pack2(1.f, 1.f), none = pack2(-1.f, -1.f);
  const uint64_t xr = fadd2(x2, magic);                 // round(x) in the low mantissa bits
  const uint64_t n = ffma2(xr, one, nmagic);            // round(x) as a float
  const uint64_t f = ffma2(n, none, x2);                // x - round(x)  in [-0.5, 0.5]
  uint64_t q = ffma2(f, pack2(0.05508868f, 0.05508868f), pack2(0.24260405f, 0.24260405f));
  q = ffma2(q, f, pack2(0.69327623f, 0.69327623f));
  q = ffma2(q, f, pack2(0.99992895f, 0.99992895f));
  float p0, p1, x0, x1;
  unpack2(q, p0, p1);
  unpack2(xr, x0, x1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(x0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(x1) << 23));
}

// EMU = how many of every four column pairs take the polynomial exp2 (FMA pipe) instead of MUFU.EX2
// NT  = query tiles per CTA: 2 (ping-pong, the fast configuration) or 1.  With few heads per GPU (tensor parallelism) the
//       grid of two-tile CTAs is a single unbalanced wave -- 128 CTAs on 148 SMs whose longest member is 2.2x the mean --
//       so small grids launch twice as many one-tile CTAs (one softmax warpgroup and one issuer stay idle).
template <int EMU, int NT>
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
  uint64_t* s_full = bars + 12;                     // [2 tiles][2 buffers]  S_t(i) is in TMEM buffer i & 1
  // P_t(i) is in TMEM buffer i & 1 (O_t rescaled if necessary).  One barrier PER BUFFER: the softmax warps may publish
  // block i+1 before the issuer has looked at block i, and a parity wait is only sound while the waiter lags its
  // barrier by at most one phase.
  uint64_t* p_full = bars + 16;                     // [2 tiles][2 buffers]
  uint64_t* pv_done = bars + 20;                    // [2]  PV_t(i) has retired (one phase per block)
  uint64_t* o_full = bars + 22;                     // [2]  the LAST P·V of tile t has retired (single use)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);

  // heavy (late) query blocks first; heads vary fastest so that concurrently running CTAs have equal lengths and the
  // heads of one GQA group share their K/V tiles in L2
  const int qp = p.num_q_units - 1 - (int)blockIdx.y;
  const int head = blockIdx.x % p.Hq, batch = blockIdx.x / p.Hq;
  const int kv_head = head / (p.Hq / p.Hk);
  const int q0 = qp * NT * BM;
  const int qa0 = p.q_off + q0;                                        // absolute position of the CTA's first query
  const int j_lo = (p.window > 0) ? max(0, (qa0 - p.window + 1) / BN) : 0;
  const int j_hi = p.causal ? qa0 / BN + (NT - 1) : p.Sk / BN - 1;     // inclusive, for the LAST query tile of the CTA
  const int n = j_hi - j_lo + 1;                                       // K / V tiles (128 keys) of the last query tile
  const int n0 = (p.causal && NT == 2) ? n - 1 : n;                    // ... of query tile 0 (its diagonal is one earlier)
  const int N1 = 2 * n, N0 = 2 * n0;                                   // 64-key softmax blocks per query tile
  const int row_base = batch * p.S;                                    // row of query 0 of this batch in the [B*S, Hq*D] views
  const int krow_base = batch * p.Sk;                                  // row of key 0 of this batch in the [B*Sk, Hk*D] views

  if (warp_idx == 8 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[2 * i], 1);
      mbar_init(&s_full[2 * i + 1], 1);
      mbar_init(&p_full[2 * i], 4);
      mbar_init(&p_full[2 * i + 1], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < KV_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], NT);      // one release per query tile
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
      if (NT == 2) load_q(1);
      load_kv(&tma_v, j_lo);
      for (int i = 1; i < n; ++i) {
        load_kv(&tma_k, j_lo + i);
        load_kv(&tma_v, j_lo + i);
      }
    }
  } else if (warp_idx == 9 || (NT == 2 && warp_idx == 10)) {
    // =========================== MMA issuers: warp 9 -> query tile 0, warp 10 -> query tile 1 ===========================
    // One elected thread per tile runs the whole chain  wait P_t(i) -> PV_t(i) -> QK_t(i+2).  Two independent issuers keep
    // the per-instruction overhead (barrier polls, descriptor arithmetic, commits) of the 64-key blocks off the critical
    // path: measured with one in-order issuer for both tiles, the issue loop itself took 1935 cycles per block pair
    // against 1024 cycles of tensor work.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(BM, BK, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D, false, true);
      const int t = warp_idx - 9;
      const int nt = (t == 0 && NT == 2) ? N0 : N1;
      const uint32_t q_addr = smem_u32(sQ) + t * TILE_BYTES, kv_addr = smem_u32(sKV);
      const uint32_t tS = tmem_base + t * 128, tO = tmem_base + 256 + t * D;
      const bool stamp_cta = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
      // a K / V tile that query tile 0 never touches (the last one under causal masking) is released twice by tile 1
      auto release = [&](int u) {
        umma_commit(&kv_empty[u % KV_SLOTS]);
        if (NT == 2 && t == 1 && (u >> 1) >= n0) umma_commit(&kv_empty[u % KV_SLOTS]);
      };
      auto issue_qk = [&](int i) {
        const int u = 2 * (i >> 1), half = i & 1;
        if (half == 0) mbar_wait(&kv_full[u % KV_SLOTS], (u / KV_SLOTS) & 1);
        // rows [64 half, 64 half + 64) of both 64-column chunks of the K tile
        const uint32_t k_addr = kv_addr + (u % KV_SLOTS) * TILE_BYTES + half * (BK * 128);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16<1>(tS + half * BK, desc_kmajor(q_addr, kk), desc_kmajor(k_addr, kk), idesc_qk, kk != 0);
        umma_commit(&s_full[2 * t + half]);
        if (half == 1) release(u);
      };
      mbar_wait(&q_full[t], 0);
      issue_qk(0);
      issue_qk(1);
      for (int i = 0; i < nt; ++i) {
        const int u = 2 * (i >> 1) + 1, half = i & 1;
        mbar_wait(&p_full[2 * t + half], (i >> 1) & 1);
        if (stamp_cta && i < 128) p.dbg[(2 * 128 + i) * 4 + 2 * t] = clock64();        // issuer saw P_t(i)
        if (half == 0) mbar_wait(&kv_full[u % KV_SLOTS], (u / KV_SLOTS) & 1);
        tcgen05_fence_after();
        const uint32_t v_addr = kv_addr + (u % KV_SLOTS) * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk)
          umma_f16_ts(tO, tS + half * BK + kk * 8, desc_mnmajor(v_addr, 4 * half + kk), idesc_pv, (i | kk) != 0);
        umma_commit(&pv_done[t]);
        if (i + 1 == nt) umma_commit(&o_full[t]);
        if (half == 1) release(u);
        if (i + 2 < nt) issue_qk(i + 2);
        if (stamp_cta && i < 128) p.dbg[(2 * 128 + i) * 4 + 2 * t + 1] = clock64();    // PV_t(i) and QK_t(i+2) issued
      }
    }
  } else if (warp_idx >= 10 || (NT == 1 && warp_idx >= 4)) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");      // idle warps (with NT == 1 also the second softmax warpgroup)
  } else {
    // =========================== softmax + epilogue (thread = query row) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = warp_idx >> 2;                     // query tile
    const uint32_t qd = warp_idx & 3;                // TMEM lane quadrant
    const uint32_t row = qd * 32 + lane;
    const uint32_t lane_off = (qd * 32u) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + t * D;
    const int q_tile0 = qa0 + t * BM;                 // absolute position of the tile's first query
    const int q_pos = q_tile0 + (int)row;             // absolute position (mask arithmetic)
    const int nt = (t == 0 && NT == 2) ? N0 : N1;
    const float sc = p.scale_log2;
    float m_used = -INFINITY;                        // raw (unscaled) score domain
    uint64_t l2 = pack2(0.f, 0.f);                   // two partial row sums
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int k0 = j_lo * BN + i * BK;             // first key of this block
      mbar_wait(&s_full[2 * t + b], (i >> 1) & 1);
      tcgen05_fence_after();
      const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && qd == 0 && i < 128;
      if (stamp) p.dbg[(t * 128 + i) * 4 + 0] = clock64();            // S_t(i) observed
      uint32_t sr[2][32];                            // the fp32 score row (raw bits), 2 x 32 columns
      tmem_ld_32x32b_x32(tS + b * BK, sr[0]);
      tmem_ld_32x32b_x32(tS + b * BK + 32, sr[1]);
      tmem_ld_wait();
      const bool need_mask = (p.causal && k0 + BK - 1 > q_tile0) || (p.window > 0 && k0 <= q_pos - p.window);
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int kpos = k0 + c * 32 + e;
            if ((p.causal && kpos > q_pos) || (p.window > 0 && kpos <= q_pos - p.window)) sr[c][e] = 0xff800000u;  // -inf
          }
      }
      float mx[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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
      const float m_new = fmax3(m_used, mx[0], mx[1]);
      const bool grow = (m_new - m_used) * sc > RESCALE_THRESHOLD;      // true on the first block (m_used = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float alpha = (m_used == -INFINITY) ? 0.f : ex2((m_used - m_new) * sc);
        const uint64_t a2 = pack2(alpha, alpha);
        l2 = fmul2(l2, a2);
        m_used = m_new;
        if (i > 0) {
          // O_t is only stable once the previous P·V of this tile has retired.  (S_t(i) has been seen, so PV_t(<= i-2) are
          // done and the barrier is at most one phase behind the parity waited for.)
          mbar_wait(&pv_done[t], (i - 1) & 1);
          tcgen05_fence_after();
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
      if (stamp) p.dbg[(t * 128 + i) * 4 + 1] = clock64();            // exponentials start
      // a row can still be fully masked so far (sliding window): keep exp2(-inf - 0) = 0 instead of NaN
      const float m_sc = (m_used == -INFINITY) ? 0.f : m_used * sc;
      const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_sc, -m_sc);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float x0, x1;
          const uint64_t xs = ffma2(pack2(__uint_as_float(sr[c][2 * e]), __uint_as_float(sr[c][2 * e + 1])), sc2, nm2);
          if ((e & 3) < EMU) {
            exp2_poly2(xs, x0, x1);
          } else {
            unpack2(xs, x0, x1);
            x0 = ex2(x0);
            x1 = ex2(x1);
          }
          l2 = fadd2(l2, pack2(x0, x1));
          pk[e] = pack_bf16x2(x0, x1);
        }
        tmem_st_32x32b_x16(tS + b * BK + c * 16, pk);
      }
      if (stamp) p.dbg[(t * 128 + i) * 4 + 2] = clock64();            // exponentials issued
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * t + b]);
      if (stamp) p.dbg[(t * 128 + i) * 4 + 3] = clock64();            // P_t(i) published
    }
    // ---------------- epilogue: O / l -> bf16 -> smem (the Q tile is free now) -> TMA store; LSE ----------------
    float l_lo, l_hi;
    unpack2(l2, l_lo, l_hi);
    const float l = l_lo + l_hi;
    // (pv_done cannot be used here: a parity wait is only sound when the waiter lags the barrier by at most one phase,
    // and after the loop this thread may be zero, one or two P·V products behind; o_full is armed once.)
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
  TORCH_CHECK(S % BM == 0, "flash_attn_fwd2: the query length must be a multiple of 128");
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
  p.B = B; p.S = S; p.Sk = Sk; p.q_off = (int)q_offset; p.Hq = Hq; p.Hk = Hk; p.window = (int)window;
  p.causal = causal ? 1 : 0;
  p.scale_log2 = (float)(scale * 1.4426950408889634);
  p.lse = lse.data_ptr<float>();
  p.dbg = g_fwd2_dbg;
  constexpr int EMU = 1;      // 25 % of the exponentials on the FMA pipe (measured: 0 / 25 / 50 % -> 1292 / 1297 / 1257 TFLOP/s)
  static bool once = false;
  if (!once) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_fwd2_kernel<EMU, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_fwd2_kernel<EMU, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  // two-tile CTAs need an even number of 128-row tiles and enough CTAs to balance: below two CTAs per SM the longest
  // (last) query pair dominates the single wave, and one-tile CTAs halve that critical path
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const bool two_tiles = S % (2 * BM) == 0 && (int64_t)Hq * B * (S / (2 * BM)) >= 2 * (int64_t)sms;
  p.num_q_units = S / (BM * (two_tiles ? 2 : 1));
  dim3 grid(Hq * B, p.num_q_units);
  if (two_tiles) flash_fwd2_kernel<EMU, 2><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  else flash_fwd2_kernel<EMU, 1><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {out, lse};
}
