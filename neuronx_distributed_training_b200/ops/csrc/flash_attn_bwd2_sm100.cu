// Flash attention BACKWARD, second generation, for sm_100a (head_dim 128, GQA-native; causal with optional sliding
// window, or full attention).  This is the kernel on the training path (reference call site: the NKI flash kernel's
// backward behind src/neuronx_distributed_training/models/hf_models/modeling_llama.py:482-489).
//
// One CTA owns one 128-key block j of one (batch, kv head) and walks over 64-query chunks of every query head of the GQA
// group, so dK_j / dV_j accumulate in tensor memory over the whole walk and are written once.  Everything is computed
// TRANSPOSED (TMEM lane = key), which lets P^T feed the dV product straight from tensor memory:
//     S^T  = K_j Q_i^T          A = K  (smem, K-major)     B = Q  (smem, K-major)     -> TMEM S buffer (double buffered)
//     dP^T = V_j dO_i^T         A = V  (smem, K-major)     B = dO (smem, K-major)     -> TMEM dP
//   compute warps (thread = key row):  P^T = exp2(S^T c - LSE_i) -> bf16, written over S^T in TMEM;
//                                      dS^T = P^T o (dP^T - Delta_i) scale -> bf16 tile in shared memory (double buffered)
//     dV  += P^T  dO_i          A = P^T (TMEM)             B = dO (smem, MN-major)
//     dK  += dS^T Q_i           A = dS^T (smem, K-major)   B = Q  (smem, MN-major)
//     dQ^T = K_j^T dS^T         A = K  (smem, MN-major)    B = dS^T (smem, MN-major)  -> TMEM dQ (lane = head-dim index)
//   writer warps: dQ^T TMEM -> registers -> fp32 staging tile [64 q x 128 d] -> TMA reduce-add into the fp32 dQ
//   accumulator (other key blocks add to the same rows).
// The five GEMMs of step t are software-pipelined against the compute warps: S^T(t+1) is issued before dV(t), and
// dP^T(t+1) as soon as dS^T(t) exists, so the tensor core always has the next scores ready while the exponentials run.
//
//   warps 0..7   compute (P, dS): two warpgroups, each takes 32 of the 64 query columns of every key row (the in-kernel
//                timeline showed ONE compute warpgroup to be the critical path: 2260 cycles per step against 1280 cycles
//                of tensor work); they also write dV (warpgroup 0) and dK (warpgroup 1) at the end
//   warps 8..11  dQ writers
//   warp  12     TMA producer (K, V once; Q_i, dO_i, LSE_i, Delta_i through a 3-stage ring)
//   warps 13, 14 tcgen05 issuers (A: S^T, dP^T, dV;  B: dK, dQ^T)  warp 15       pads the fourth warpgroup (setmaxnreg)
// Tensor memory (512 columns): [0,64) S^T/P^T even, [64,128) S^T/P^T odd, [128,192) dP^T, [192,256) dQ^T,
// [256,384) dV, [384,512) dK.
// A small pre-pass produces, per (batch, head, query):  -LSE log2(e)  and  -Delta scale  with Delta = rowsum(dO o O).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "sm100_ptx.cuh"
#include "tma_utils.h"

namespace fab2 {

using namespace sm100;

constexpr int BQ = 64;           // query rows per step
constexpr int BN = 128;          // keys per CTA
constexpr int D = 128;           // head dim
constexpr int KV_TILE = 128 * 128 * 2;        // K / V tile: two [128 x 64] swizzle chunks of 16 KB
constexpr int KV_CHUNK = 128 * 64 * 2;
constexpr int Q_TILE = 64 * 128 * 2;          // Q / dO tile: two [64 x 64] swizzle chunks of 8 KB
constexpr int Q_CHUNK = 64 * 64 * 2;
constexpr int DS_TILE = 128 * 64 * 2;         // dS^T tile: one [128 keys x 64 q] swizzle chunk
constexpr int DQ_STAGE = 64 * 128 * 4;        // fp32 [64 q x 128 d] as four [64 x 32] chunks of 8 KB
constexpr int DQ_CHUNK = 64 * 32 * 4;
constexpr int NS = 3;                         // Q / dO ring stages
constexpr int STAT_BYTES = 2 * BQ * 4;        // per stage: 64 x (-LSE log2 e), 64 x (-Delta scale)
constexpr int NUM_THREADS = 512;
constexpr int SMEM_BYTES = 2 * KV_TILE + NS * 2 * Q_TILE + 2 * DS_TILE + DQ_STAGE + NS * STAT_BYTES + 256 + 1024;

struct Params {
  int B, S, Hq, Hk;        // S = number of (local) queries per batch entry
  int Sk;                  // number of keys per batch entry (the gathered sequence under context parallelism)
  int q_off;               // absolute position of query 0 relative to key 0 (multiple of 128)
  int num_kblocks;         // Sk / 128
  int window;              // 0 = none
  int causal;
  float scale;             // softmax scale
  float scale_log2;        // scale * log2(e)
  const float* nlse2;      // [B, Hq, S]  -LSE * log2(e)   (-inf for rows without any visible key)
  const float* nds;        // [B, Hq, S]  -Delta * scale
  int split_heads;         // 1: one CTA per (key block, QUERY head); dK / dV are written per query head and summed over the
                           //    GQA group afterwards.  Used when (KV heads x batches x key blocks) would not fill the GPU.
  long long* dbg;          // optional in-kernel timeline of CTA (0, 0): [4 roles][256 steps][4 stamps] of clock64 (else nullptr)
};

// ---- shared-memory matrix descriptors (128B swizzle); `rows_chunk_bytes` = bytes of one 64-column chunk of the tile ----
__device__ __forceinline__ uint64_t desc_k(uint32_t tile, int kk, int chunk_bytes) {
  // K-major operand: kk = 16-element step along the contiguous dimension (4 steps per 64-column chunk)
  return make_smem_desc_sw128(tile + (kk >> 2) * chunk_bytes + (kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t tile, int kk, int chunk_bytes) {
  // MN-major operand: rows of the tile are the reduction index (16 rows = 2048 bytes per step), the 64-column chunks
  // are groups along M / N
  return make_smem_desc_sw128(tile + kk * 2048, chunk_bytes, 1024);
}

__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
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
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// pre-pass: one warp per (b, s, h) row of 128 elements
//   nds[b, h, s]   = -scale * sum_d dO[b, s, h, d] * O[b, s, h, d]
//   nlse2[b, h, s] = -lse[b, h, s] * log2(e)        (-inf when lse = -inf)
__global__ void __launch_bounds__(256) flash_bwd2_prep_kernel(const __nv_bfloat16* __restrict__ dout,
                                                              const __nv_bfloat16* __restrict__ out,
                                                              const float* __restrict__ lse, float* __restrict__ nds,
                                                              float* __restrict__ nlse2, int B, int S, int H, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)B * S * H;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const uint2 a = *reinterpret_cast<const uint2*>(dout + r * D + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(out + r * D + lane * 4);
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 x = __bfloat1622float2(pa[i]), y = __bfloat1622float2(pb[i]);
      acc += x.x * y.x + x.y * y.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const int64_t h = r % H, bs = r / H;
      const int64_t s = bs % S, b_ = bs / S;
      const int64_t idx = (b_ * H + h) * S + s;
      nds[idx] = -acc * scale;
      const float l = lse[idx];
      nlse2[idx] = (l == -INFINITY) ? -INFINITY : -l * 1.4426950408889634f;
    }
  }
}

// fp32 dQ accumulator -> bf16
__global__ void __launch_bounds__(256) flash_bwd2_cast_kernel(const float4* __restrict__ in, uint2* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    out[i] = o;
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_bwd2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                  const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                  const __grid_constant__ CUtensorMap tma_dqh, const __grid_constant__ CUtensorMap tma_dk,
                  const __grid_constant__ CUtensorMap tma_dv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_TILE;
  uint8_t* sQ = sV + KV_TILE;                   // [NS]
  uint8_t* sdO = sQ + NS * Q_TILE;              // [NS]
  uint8_t* sdS = sdO + NS * Q_TILE;             // [2]
  uint8_t* sdQ = sdS + 2 * DS_TILE;             // fp32 staging
  float* sStat = reinterpret_cast<float*>(sdQ + DQ_STAGE);       // [NS][2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + NS * STAT_BYTES);
  uint64_t* kv_full = bars;            // 1
  uint64_t* ring_full = bars + 1;      // [NS]
  uint64_t* ring_empty = bars + 4;     // [NS]
  uint64_t* s_full = bars + 7;         // [2]
  uint64_t* dp_full = bars + 9;        // 1
  uint64_t* p_full = bars + 10;        // [2] (4 warps arrive; one per S^T buffer so that no waiter lags two phases)
  uint64_t* ds_full = bars + 12;       // [2] (4 warps arrive)
  uint64_t* ds_free = bars + 14;       // [2]
  uint64_t* dq_full = bars + 16;       // 1
  uint64_t* dq_free = bars + 17;       // 1   (4 warps arrive)
  uint64_t* dkv_full = bars + 18;      // 1   (both MMA issuers commit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);

  // key block 0 has the longest walk: key blocks vary SLOWEST in launch order (longest-processing-time-first over the
  // whole grid), heads and batches fastest
  const int j = blockIdx.y;
  const int groups = p.split_heads ? p.Hq : p.Hk;             // CTAs per (batch, key block)
  const int grp = blockIdx.x % groups, batch = blockIdx.x / groups;
  const int kv_head = p.split_heads ? grp / (p.Hq / p.Hk) : grp;
  const int rep = p.split_heads ? 1 : p.Hq / p.Hk;            // query heads this CTA walks over
  const int head0 = p.split_heads ? grp : kv_head * rep;      // first of them
  const int out_col = grp * D;                                // column of this CTA's dK / dV tile in the output views
  const int k0 = j * BN;
  const int n_chunks = p.S / BQ;
  // local 64-query chunks whose (absolute) positions q_off + 64 i .. can see keys of this block
  const int i_lo = p.causal ? max(0, (k0 - p.q_off) / BQ) : 0;
  int i_hi = n_chunks - 1;
  if (p.window > 0) i_hi = min(i_hi, (k0 + BN - 1 + p.window - 1 - p.q_off) / BQ - ((k0 + BN - 1 + p.window - 1 - p.q_off) < 0 ? 1 : 0));
  const int n_i = max(0, i_hi - i_lo + 1);
  const int steps = rep * n_i;                   // 0: no local query sees this key block (dK = dV = 0)
  const int row_base = batch * p.S;
  const int krow_base = batch * p.Sk;

  if (warp_idx == 12 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dqh); tma_prefetch_desc(&tma_dk); tma_prefetch_desc(&tma_dv);
    mbar_init(kv_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 2); }   // both MMA issuers release a stage
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&ds_full[i], 8); mbar_init(&ds_free[i], 1); }
    mbar_init(dp_full, 1);
    mbar_init(&p_full[0], 8);
    mbar_init(&p_full[1], 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(dkv_full, 2);                      // the last dV (issuer A) and the last dK (issuer B)
    fence_barrier_init();
  }
  if (warp_idx == 13) {
    tmem_alloc<1>(tmem_ptr, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_dp = tmem_base + 128, tmem_dq = tmem_base + 192, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 384;

  if (warp_idx == 12) {
    // =========================== TMA producer ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * KV_TILE);
      tma_load_2d(sK, &tma_k, kv_full, kv_head * D, krow_base + k0);
      tma_load_2d(sK + KV_CHUNK, &tma_k, kv_full, kv_head * D + 64, krow_base + k0);
      tma_load_2d(sV, &tma_v, kv_full, kv_head * D, krow_base + k0);
      tma_load_2d(sV + KV_CHUNK, &tma_v, kv_full, kv_head * D + 64, krow_base + k0);
      for (int t = 0; t < steps; ++t) {
        const int st = t % NS;
        const int head = head0 + t / n_i;
        const int q0 = (i_lo + t % n_i) * BQ;
        mbar_wait(&ring_empty[st], ((t / NS) & 1) ^ 1);
        mbar_expect_tx(&ring_full[st], 2 * Q_TILE + STAT_BYTES);
        tma_load_2d(sQ + st * Q_TILE, &tma_q, &ring_full[st], head * D, row_base + q0);
        tma_load_2d(sQ + st * Q_TILE + Q_CHUNK, &tma_q, &ring_full[st], head * D + 64, row_base + q0);
        tma_load_2d(sdO + st * Q_TILE, &tma_do, &ring_full[st], head * D, row_base + q0);
        tma_load_2d(sdO + st * Q_TILE + Q_CHUNK, &tma_do, &ring_full[st], head * D + 64, row_base + q0);
        const size_t stat = ((size_t)batch * p.Hq + head) * p.S + q0;
        bulk_g2s(sStat + st * 2 * BQ, p.nlse2 + stat, BQ * 4, &ring_full[st]);
        bulk_g2s(sStat + st * 2 * BQ + BQ, p.nds + stat, BQ * 4, &ring_full[st]);
      }
    }
  } else if (warp_idx == 13) {
    // =========================== MMA issuer A: S^T, dP^T, dV (one elected thread runs the whole loop) ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one() && steps > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, BQ, false, false);     // S^T, dP^T   [128 x 64], K = 128
      constexpr uint32_t idesc_dv = make_idesc_bf16(128, D, false, true);      // dV (A in TMEM)  [128 x 128], K = 64
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_base = smem_u32(sQ), do_base = smem_u32(sdO);
      auto issue_s = [&](int t) {          // S^T(t) -> buffer t & 1
        const int st = t % NS;
        mbar_wait(&ring_full[st], (t / NS) & 1);
        const uint32_t q_addr = q_base + st * Q_TILE;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16<1>(tmem_base + (t & 1) * BQ, desc_k(k_addr, kk, KV_CHUNK), desc_k(q_addr, kk, Q_CHUNK), idesc_s, kk != 0);
        umma_commit(&s_full[t & 1]);
      };
      auto issue_dp = [&](int t) {         // dP^T(t); the ring stage has been waited for by issue_s(t)
        const uint32_t do_addr = do_base + (t % NS) * Q_TILE;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16<1>(tmem_dp, desc_k(v_addr, kk, KV_CHUNK), desc_k(do_addr, kk, Q_CHUNK), idesc_s, kk != 0);
        umma_commit(dp_full);
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      issue_dp(0);
      for (int t = 0; t < steps; ++t) {
        const int st = t % NS, b = t & 1;
        if (t + 1 < steps) issue_s(t + 1);         // buffer (t+1) & 1: its P^T(t-1) was consumed by dV(t-1), issued earlier by this thread
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && t < 256;
        if (stamp) p.dbg[(2 * 256 + t) * 4 + 0] = clock64();        // A: S^T(t+1) issued, waiting for P^T(t)
        mbar_wait(&p_full[b], (t >> 1) & 1);
        tcgen05_fence_after();
        if (stamp) p.dbg[(2 * 256 + t) * 4 + 1] = clock64();        // A: P^T(t) seen
        const uint32_t do_addr = do_base + st * Q_TILE;
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk)       // dV += P^T dO
          umma_f16_ts(tmem_dv, tmem_base + b * BQ + (kk >> 1) * 32 + (kk & 1) * 8, desc_mn(do_addr, kk, Q_CHUNK), idesc_dv,
                      (t | kk) != 0);     // P^T of query half h sits at columns [32 h, 32 h + 16) of the buffer
        umma_commit(&ring_empty[st]);              // this issuer's reads of ring stage st (S, dP, dV) are done when this fires
        if (t + 1 == steps) umma_commit(dkv_full);
        if (stamp) p.dbg[(2 * 256 + t) * 4 + 2] = clock64();        // A: dV(t) issued
        if (t + 1 < steps) {
          mbar_wait(&ds_full[b], (t >> 1) & 1);    // the compute warps have read dP^T(t)
          tcgen05_fence_after();
          issue_dp(t + 1);
        }
        if (stamp) p.dbg[(2 * 256 + t) * 4 + 3] = clock64();        // A: dP^T(t+1) issued
      }
    }
  } else if (warp_idx == 14) {
    // =========================== MMA issuer B: dK, dQ^T ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one() && steps > 0) {
      constexpr uint32_t idesc_dk = make_idesc_bf16(128, D, false, true);      // dK   [128 x 128], K = 64
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, BQ, true, true);      // dQ^T [128 x 64], K = 128
      const uint32_t k_addr = smem_u32(sK), q_base = smem_u32(sQ), ds_base = smem_u32(sdS);
      mbar_wait(kv_full, 0);
      for (int t = 0; t < steps; ++t) {
        const int st = t % NS, b = t & 1;
        mbar_wait(&ds_full[b], (t >> 1) & 1);      // dS^T(t) is in shared memory (its ring stage landed long before)
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && t < 256;
        if (stamp) p.dbg[(3 * 256 + t) * 4 + 0] = clock64();        // B: dS^T(t) seen
        const uint32_t q_addr = q_base + st * Q_TILE, ds_addr = ds_base + b * DS_TILE;
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk)       // dK += dS^T Q   (reduction over the 64 queries)
          umma_f16<1>(tmem_dk, desc_k(ds_addr, kk, DS_TILE), desc_mn(q_addr, kk, Q_CHUNK), idesc_dk, (t | kk) != 0);
        if (stamp) p.dbg[(3 * 256 + t) * 4 + 1] = clock64();        // B: dK(t) issued
        mbar_wait(dq_free, (t & 1) ^ 1);           // dQ^T(t-1) has left tensor memory
        tcgen05_fence_after();
        if (stamp) p.dbg[(3 * 256 + t) * 4 + 2] = clock64();        // B: dQ^T buffer free
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)       // dQ^T = K^T dS^T   (reduction over the 128 keys)
          umma_f16<1>(tmem_dq, desc_mn(k_addr, kk, KV_CHUNK), desc_mn(ds_addr, kk, DS_TILE), idesc_dq, kk != 0);
        umma_commit(dq_full);
        umma_commit(&ring_empty[st]);
        umma_commit(&ds_free[b]);
        if (t + 1 == steps) umma_commit(dkv_full);
        if (stamp) p.dbg[(3 * 256 + t) * 4 + 3] = clock64();        // B: dQ^T(t) issued
      }
    }
  } else if (warp_idx >= 15) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else if (warp_idx >= 8) {
    // =========================== dQ writers (thread = head-dim index) ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    const uint32_t qd = warp_idx & 3;
    const uint32_t lane_off = (qd * 32u) << 16;
    // every warp owns one staging chunk (its 32 head-dim columns x 64 queries, 8 KB) and reduces it on its own: no
    // cross-warp barrier.  The chunk is written and reduced in two halves of 32 queries so that the stores of one half
    // overlap the TMA read of the other.
    const uint32_t my = smem_u32(sdQ) + qd * DQ_CHUNK + lane * 4;     // row = query (128 bytes), column = lane
    const uint8_t* chunk = sdQ + qd * DQ_CHUNK;
    for (int t = 0; t < steps; ++t) {
      const int head = head0 + t / n_i;
      const int q0 = (i_lo + t % n_i) * BQ;
      mbar_wait(dq_full, t & 1);
      tcgen05_fence_after();
      const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 256 && t < 256;
      if (stamp) p.dbg[(1 * 256 + t) * 4 + 0] = clock64();          // dQ^T(t) seen
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(tmem_dq + lane_off, v[0]);
      tmem_ld_32x32b_x32(tmem_dq + lane_off + 32, v[1]);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      if (stamp) p.dbg[(1 * 256 + t) * 4 + 1] = clock64();          // dQ^T(t) in registers, TMEM released
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // the reduce that read this half one step ago must have finished with it (at most one newer group in flight)
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; ++e) st_shared_b32(my + (h * 32 + e) * 128, v[h][e]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tma_dqh, chunk + h * (DQ_CHUNK / 2), head * D + qd * 32, row_base + q0 + h * 32);
          tma_store_commit();
        }
        if (stamp) p.dbg[(1 * 256 + t) * 4 + 2 + h] = clock64();    // half h staged and its reduce issued
      }
    }
    if (lane == 0) tma_store_wait<0>();
  } else {
    // =========================== compute: P^T, dS^T (thread = key row x half of the query columns) ===========================
    // Two warpgroups split the 64 query columns of a step: two warps per scheduler hide each other's TMEM / shared-memory
    // / MUFU latencies.  (Measured alternatives: one warpgroup doing all 64 columns 996 TFLOP/s; a two-stage pipeline
    // P -> dS over the two warpgroups 735 TFLOP/s; this column split 1034 TFLOP/s.)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    constexpr int HQ = BQ / 2;                       // query columns per thread
    const uint32_t h = warp_idx >> 2;                // which half of the query columns (warpgroup 0 / 1)
    const uint32_t qd = warp_idx & 3;
    const uint32_t row = qd * 32 + lane;
    const uint32_t lane_off = (qd * 32u) << 16;
    const int kpos = k0 + (int)row;
    const uint64_t sc2 = pack2(p.scale_log2, p.scale_log2), scl2 = pack2(p.scale, p.scale);
    for (int t = 0; t < steps; ++t) {
      const int st = t % NS, b = t & 1;
      const int q0 = (i_lo + t % n_i) * BQ;
      const uint32_t nl_addr = smem_u32(sStat) + st * STAT_BYTES + h * (HQ * 4), nd_addr = nl_addr + BQ * 4;
      mbar_wait(&ring_full[st], (t / NS) & 1);              // LSE / Delta of this chunk are in shared memory
      mbar_wait(&s_full[b], (t >> 1) & 1);
      tcgen05_fence_after();
      const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && t < 256;
      if (stamp) p.dbg[(0 * 256 + t) * 4 + 0] = clock64();          // S^T(t) seen
      uint32_t sr[HQ];
      tmem_ld_32x32b_x32(tmem_base + lane_off + b * BQ + h * HQ, sr);
      tmem_ld_wait();
      // masks: causal -> key after the query; window -> key at or before (query - window)
      const int qa0 = p.q_off + q0 + (int)h * HQ;            // absolute position of this thread's first query
      const bool need_mask = (p.causal && qa0 < k0 + BN) || (p.window > 0 && k0 <= qa0 + HQ - 1 - p.window);
      float pf[HQ];
#pragma unroll
      for (int g = 0; g < HQ / 4; ++g) {
        const float4 nl = ld_shared_f4(nl_addr + g * 16);
        float x0, x1, x2, x3;
        unpack2(ffma2(pack2(__uint_as_float(sr[4 * g]), __uint_as_float(sr[4 * g + 1])), sc2, pack2(nl.x, nl.y)), x0, x1);
        unpack2(ffma2(pack2(__uint_as_float(sr[4 * g + 2]), __uint_as_float(sr[4 * g + 3])), sc2, pack2(nl.z, nl.w)), x2, x3);
        pf[4 * g + 0] = ex2(x0);
        pf[4 * g + 1] = ex2(x1);
        pf[4 * g + 2] = ex2(x2);
        pf[4 * g + 3] = ex2(x3);
      }
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < HQ; ++c) {
          const int q_pos = qa0 + c;
          if ((p.causal && kpos > q_pos) || (p.window > 0 && kpos <= q_pos - p.window)) pf[c] = 0.f;
        }
      }
      {
        uint32_t pk[HQ / 2];
#pragma unroll
        for (int e = 0; e < HQ / 2; ++e) pk[e] = pack_bf16x2(pf[2 * e], pf[2 * e + 1]);
        // The packed P^T of this half goes over the first 16 of ITS OWN 32 score columns ([32 h, 32 h + 16)): the other
        // warpgroup may still be reading its half of S^T.  The dV product addresses the two halves separately.
        tmem_st_32x32b_x16(tmem_base + lane_off + b * BQ + h * HQ, pk);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      if (stamp) p.dbg[(0 * 256 + t) * 4 + 1] = clock64();          // P^T(t) published
      // ---- dS^T = P^T o (dP^T - Delta) * scale ----
      mbar_wait(dp_full, t & 1);
      tcgen05_fence_after();
      if (stamp) p.dbg[(0 * 256 + t) * 4 + 2] = clock64();          // dP^T(t) seen
      uint32_t dr[HQ];
      tmem_ld_32x32b_x32(tmem_dp + lane_off + h * HQ, dr);
      tmem_ld_wait();
      mbar_wait(&ds_free[b], ((t >> 1) & 1) ^ 1);          // the MMAs that read this dS^T buffer two steps ago have retired
      const uint32_t dsrow = smem_u32(sdS) + b * DS_TILE + row * 128;
#pragma unroll
      for (int g = 0; g < HQ / 8; ++g) {                    // 8 queries = one 16-byte group
        const float4 nda = ld_shared_f4(nd_addr + (2 * g) * 16), ndb = ld_shared_f4(nd_addr + (2 * g + 1) * 16);
        const int c = 8 * g;
        float e[8];
        unpack2(fmul2(ffma2(pack2(__uint_as_float(dr[c + 0]), __uint_as_float(dr[c + 1])), scl2, pack2(nda.x, nda.y)),
                      pack2(pf[c + 0], pf[c + 1])), e[0], e[1]);
        unpack2(fmul2(ffma2(pack2(__uint_as_float(dr[c + 2]), __uint_as_float(dr[c + 3])), scl2, pack2(nda.z, nda.w)),
                      pack2(pf[c + 2], pf[c + 3])), e[2], e[3]);
        unpack2(fmul2(ffma2(pack2(__uint_as_float(dr[c + 4]), __uint_as_float(dr[c + 5])), scl2, pack2(ndb.x, ndb.y)),
                      pack2(pf[c + 4], pf[c + 5])), e[4], e[5]);
        unpack2(fmul2(ffma2(pack2(__uint_as_float(dr[c + 6]), __uint_as_float(dr[c + 7])), scl2, pack2(ndb.z, ndb.w)),
                      pack2(pf[c + 6], pf[c + 7])), e[6], e[7]);
        uint4 o;
        o.x = pack_bf16x2(e[0], e[1]); o.y = pack_bf16x2(e[2], e[3]);
        o.z = pack_bf16x2(e[4], e[5]); o.w = pack_bf16x2(e[6], e[7]);
        const int g16 = (int)h * (HQ / 8) + g;              // 16-byte group 0..7 within the 128-byte row
        st_shared_v4(dsrow + ((g16 ^ (row & 7)) * 16), o);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[b]);
      if (stamp) p.dbg[(0 * 256 + t) * 4 + 3] = clock64();          // dS^T(t) published
    }
    // ---------------- warpgroup 0: dV_j, warpgroup 1: dK_j: TMEM -> bf16 -> smem (the V / K tile is free) -> TMA store ----------------
    if (steps > 0) mbar_wait(dkv_full, 0);
    else mbar_wait(kv_full, 0);                  // no step ran: the accumulators were never written, the result is zero
    tcgen05_fence_after();
    {
      const uint32_t src = h == 0 ? tmem_dv : tmem_dk;
      const uint32_t orow = smem_u32(h == 0 ? sV : sK) + row * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(src + lane_off + c * 32, v);
        tmem_ld_wait();
        if (steps == 0) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
          const int c16 = c * 4 + g;
          const int chunk = c16 >> 3, j16 = c16 & 7;
          st_shared_v4(orow + chunk * KV_CHUNK + ((j16 ^ (row & 7)) * 16), o);
        }
      }
    }
    tcgen05_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
    if ((threadIdx.x & 127) == 0) {
      if (h == 0) {
        tma_store_2d(&tma_dv, sV, out_col, krow_base + k0);
        tma_store_2d(&tma_dv, sV + KV_CHUNK, out_col + 64, krow_base + k0);
      } else {
        tma_store_2d(&tma_dk, sK, out_col, krow_base + k0);
        tma_store_2d(&tma_dk, sK + KV_CHUNK, out_col + 64, krow_base + k0);
      }
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 13) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace fab2

static long long* g_bwd2_dbg = nullptr;
// int64 tensor of >= 4 * 256 * 4 elements on the device: the next launches record the timeline of CTA (0, 0); empty -> off
void flash_bwd2_set_debug(const at::Tensor& t) {
  g_bwd2_dbg = t.numel() >= 4 * 256 * 4 ? reinterpret_cast<long long*>(t.data_ptr<int64_t>()) : nullptr;
}

// dout / q / out [B,S,Hq,128], k / v [B,Sk,Hk,128] bf16 contiguous (Sk >= q_offset + S: gathered keys under context
// parallelism, query i at absolute position q_offset + i), lse [B,Hq,S] fp32 (natural log, from the forward).
// Returns (dq [B,S,Hq,128], dk, dv [B,S,Hk,128]) in bf16.
std::vector<at::Tensor> flash_attn_bwd(const at::Tensor& dout, const at::Tensor& q, const at::Tensor& k,
                                       const at::Tensor& v, const at::Tensor& out, const at::Tensor& lse,
                                       bool causal, int64_t window, double scale, int64_t q_offset) {
  using namespace fab2;
  TORCH_CHECK(causal || window == 0, "flash_attn_bwd: a sliding window needs causal attention");
  for (const at::Tensor* t : {&dout, &q, &k, &v, &out})
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kBFloat16 && t->is_contiguous() && t->dim() == 4 && t->size(3) == D,
                "flash_attn_bwd: contiguous bf16 [B,S,H,128] tensors required");
  const int B = q.size(0), S = q.size(1), Hq = q.size(2), Hk = k.size(2);
  const int Sk = k.size(1);
  TORCH_CHECK(v.size(1) == Sk && v.size(2) == Hk && Hq % Hk == 0 && S % BN == 0 && Sk % BN == 0 && k.size(0) == B && v.size(0) == B);
  TORCH_CHECK(q_offset % BN == 0 && q_offset >= 0 && q_offset + S <= Sk, "flash_attn_bwd: keys must cover [0, q_offset + S)");
  TORCH_CHECK(dout.sizes() == q.sizes() && out.sizes() == q.sizes());
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == (int64_t)B * Hq * S);
  c10::cuda::CUDAGuard guard(q.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  auto stats = at::empty({2, B, Hq, S}, q.options().dtype(at::kFloat));
  float* nds = stats.data_ptr<float>();
  float* nlse2 = nds + (int64_t)B * Hq * S;
  {
    const int64_t rows = (int64_t)B * S * Hq;
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
    flash_bwd2_prep_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)dout.data_ptr(), (const __nv_bfloat16*)out.data_ptr(),
                                                     lse.data_ptr<float>(), nds, nlse2, B, S, Hq, (float)scale);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  auto dq_acc = at::empty({B, S, Hq, D}, q.options().dtype(at::kFloat));
  C10_CUDA_CHECK(cudaMemsetAsync(dq_acc.data_ptr(), 0, dq_acc.numel() * sizeof(float), stream));
  auto dq = at::empty_like(q);
  // few KV heads per GPU (tensor parallelism): one CTA per (key block, KV head) would leave most SMs idle and the longest
  // walk (key block 0, all heads of the group) would be the whole kernel: split the walks per query head instead
  const bool split = Hq > Hk && (int64_t)Hk * B * (Sk / BN) < 2 * (int64_t)sms;
  auto dk = split ? at::empty({B, Sk, Hq, D}, k.options()) : at::empty_like(k);
  auto dv = split ? at::empty({B, Sk, Hq, D}, v.options()) : at::empty_like(v);
  const int Ho = split ? Hq : Hk;             // heads of the dK / dV views the kernel writes
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const uint64_t rows = (uint64_t)B * S, krows = (uint64_t)B * Sk;
  CUtensorMap tq = tma::make_2d(q.data_ptr(), BF, 2, rows, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, BQ, 64);
  CUtensorMap tdo = tma::make_2d(dout.data_ptr(), BF, 2, rows, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, BQ, 64);
  CUtensorMap tk = tma::make_2d(k.data_ptr(), BF, 2, krows, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tv = tma::make_2d(v.data_ptr(), BF, 2, krows, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tdk = tma::make_2d(dk.data_ptr(), BF, 2, krows, (uint64_t)Ho * D, (uint64_t)Ho * D * 2, 128, 64);
  CUtensorMap tdv = tma::make_2d(dv.data_ptr(), BF, 2, krows, (uint64_t)Ho * D, (uint64_t)Ho * D * 2, 128, 64);
  CUtensorMap tdq = tma::make_2d(dq_acc.data_ptr(), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, (uint64_t)Hq * D,
                                 (uint64_t)Hq * D * 4, BQ / 2, 32, /*swizzle128=*/false);   // half chunk: 32 q x 32 d
  Params p;
  p.B = B; p.S = S; p.Sk = Sk; p.q_off = (int)q_offset; p.Hq = Hq; p.Hk = Hk; p.num_kblocks = Sk / BN; p.window = (int)window;
  p.causal = causal ? 1 : 0;
  p.scale = (float)scale;
  p.scale_log2 = (float)(scale * 1.4426950408889634);
  p.nlse2 = nlse2;
  p.nds = nds;
  p.split_heads = split ? 1 : 0;
  p.dbg = g_bwd2_dbg;
  static bool once = false;
  if (!once) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  dim3 grid(Ho * B, p.num_kblocks);     // heads x batches fastest, key block 0 (longest query walk) first
  flash_bwd2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  {
    const int64_t n4 = dq_acc.numel() / 4;
    const int grid2 = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sms * 16);
    flash_bwd2_cast_kernel<<<grid2, 256, 0, stream>>>((const float4*)dq_acc.data_ptr(), (uint2*)dq.data_ptr(), n4);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  if (split) {      // sum the per-query-head partials of every GQA group (fp32 accumulation)
    const int r = Hq / Hk;
    dk = dk.view({B, Sk, Hk, r, D}).sum(3, /*keepdim=*/false, at::kFloat).to(at::kBFloat16);
    dv = dv.view({B, Sk, Hk, r, D}).sum(3, /*keepdim=*/false, at::kFloat).to(at::kBFloat16);
  }
  return {dq, dk, dv};
}
