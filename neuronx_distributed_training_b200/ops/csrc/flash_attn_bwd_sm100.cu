// Causal flash attention BACKWARD for sm_100a (head_dim 128, GQA-native, optional sliding window) -- PREVIEW.
//
// Status: written and compiled in round 1, NOT yet validated on hardware and therefore not on the training path
// (`ops.flash_attention` keeps using torch SDPA for the backward).  Exposed as `flash_attn_bwd_preview`;
// `tests/gpu_checks/check_flash_bwd.py` is the numerics / timing driver for it.
//
// One CTA owns one 128-key block j of one (batch, kv head) and loops over the query blocks i >= j of every query
// head of the GQA group, so dK_j / dV_j accumulate in TMEM over the whole loop and are written once (no atomics,
// no second pass over the group).  Per (head, i) step, all five GEMMs run on tcgen05 with fp32 accumulators in TMEM:
//     S   = Q_i K_j^T                 A = Q  (K-major)   B = K  (K-major)     -> TMEM [0,128)
//     dP  = dO_i V_j^T                A = dO (K-major)   B = V  (K-major)     -> TMEM [128,256)
//   softmax warps (thread t = query row t):  P = exp2(S*c - LSE_i),  dS = P o (dP - Delta_i) * scale, both written
//   as bf16 tiles in the 128B-swizzled layout
//     dV += P^T  dO_i                 A = P  (MN-major)  B = dO (MN-major)    -> TMEM [256,384)
//     dK += dS^T Q_i                  A = dS (MN-major)  B = Q  (MN-major)    -> TMEM [384,512)
//     dQ_i = dS K_j                   A = dS (K-major)   B = K  (MN-major)    -> TMEM [0,128)  (S is consumed by then)
//   dQ_i is staged through shared memory (the P/dS tiles are free at that point) and TMA reduce-added into an fp32
//   dQ accumulator in global memory (other key blocks add to the same rows).
// Every operand tile is the same physical object: [128 rows x 128 cols] bf16 as two [128 x 64] swizzle chunks, read
// through a K-major or an MN-major shared-memory descriptor depending on its role -- no transposes anywhere.
// Delta = rowsum(dO o O) comes from a small pre-pass kernel.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "sm100_ptx.cuh"
#include "tma_utils.h"

namespace fab100 {

using namespace sm100;

constexpr int BM = 128;          // query rows per step
constexpr int BN = 128;          // keys per CTA
constexpr int D = 128;           // head dim
constexpr int TILE_BYTES = 128 * 128 * 2;
constexpr int CHUNK_BYTES = 128 * 64 * 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = 6 * TILE_BYTES + 1024 + 256;     // K, V, Q, dO, P, dS
constexpr int DQ_CHUNK_BYTES = 128 * 128;                   // [128 rows x 32 fp32] staging chunk

struct Params {
  int B, S, Hq, Hk;
  int num_blocks;          // S / 128
  int window;              // 0 = none
  float scale;             // softmax scale
  float scale_log2;        // scale * log2(e)
  const float* lse;        // [B, Hq, S] natural log
  const float* delta;      // [B, Hq, S]
};

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int kk) {
  return make_smem_desc_sw128(tile_addr + (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int kk) {
  return make_smem_desc_sw128(tile_addr + kk * (16 * 128), CHUNK_BYTES, 1024);
}

// delta[b, h, s] = sum_d dO[b, s, h, d] * O[b, s, h, d]      (one warp per row of 128 elements)
__global__ void __launch_bounds__(256) flash_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout,
                                                              const __nv_bfloat16* __restrict__ out,
                                                              float* __restrict__ delta, int B, int S, int H) {
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
      delta[(b_ * H + h) * S + s] = acc;
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                 const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                 const __grid_constant__ CUtensorMap tma_dq, const __grid_constant__ CUtensorMap tma_dk,
                 const __grid_constant__ CUtensorMap tma_dv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sQ = sV + TILE_BYTES;
  uint8_t* sdO = sQ + TILE_BYTES;
  uint8_t* sP = sdO + TILE_BYTES;
  uint8_t* sdS = sP + TILE_BYTES;            // sP..sdS+TILE is also the 64 KB fp32 dQ staging area
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + TILE_BYTES);
  uint64_t* kv_full = bars;          // K_j, V_j landed
  uint64_t* qdo_full = bars + 1;     // Q_i, dO_i landed
  uint64_t* qdo_empty = bars + 2;    // second MMA group of the step retired: Q/dO (and P/dS) may be overwritten
  uint64_t* sdp_full = bars + 3;     // S and dP are in TMEM
  uint64_t* pds_full = bars + 4;     // P and dS are in shared memory, S / dP have been read out of TMEM
  uint64_t* mma2_done = bars + 5;    // dV, dK, dQ MMAs of the step retired
  uint64_t* s_free = bars + 6;       // dQ has been read out of TMEM (4 warps arrive)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int j = blockIdx.x;                      // key block
  const int kv_head = blockIdx.y, batch = blockIdx.z;
  const int rep = p.Hq / p.Hk;
  const int k0 = j * BN;
  const int i_lo = j;                            // causal: only queries at or after the block
  int i_hi = p.num_blocks - 1;
  if (p.window > 0) i_hi = min(i_hi, (k0 + BN - 1 + p.window - 1) / BM);
  const int n_i = i_hi - i_lo + 1;
  const int steps = rep * n_i;
  const int row_base = batch * p.S;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dq); tma_prefetch_desc(&tma_dk); tma_prefetch_desc(&tma_dv);
    mbar_init(kv_full, 1);
    mbar_init(qdo_full, 1);
    mbar_init(qdo_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 1);
    mbar_init(mma2_done, 1);
    mbar_init(s_free, 4);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc<1>(tmem_ptr, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 384;

  if (warp_idx == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_2d(sK, &tma_k, kv_full, kv_head * D, row_base + k0);
      tma_load_2d(sK + CHUNK_BYTES, &tma_k, kv_full, kv_head * D + 64, row_base + k0);
      tma_load_2d(sV, &tma_v, kv_full, kv_head * D, row_base + k0);
      tma_load_2d(sV + CHUNK_BYTES, &tma_v, kv_full, kv_head * D + 64, row_base + k0);
      for (int t = 0; t < steps; ++t) {
        const int head = kv_head * rep + t / n_i;
        const int q0 = (i_lo + t % n_i) * BM;
        mbar_wait(qdo_empty, (t & 1) ^ 1);
        mbar_expect_tx(qdo_full, 2 * TILE_BYTES);
        tma_load_2d(sQ, &tma_q, qdo_full, head * D, row_base + q0);
        tma_load_2d(sQ + CHUNK_BYTES, &tma_q, qdo_full, head * D + 64, row_base + q0);
        tma_load_2d(sdO, &tma_do, qdo_full, head * D, row_base + q0);
        tma_load_2d(sdO + CHUNK_BYTES, &tma_do, qdo_full, head * D + 64, row_base + q0);
      }
    }
  } else if (warp_idx == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, false, false);   // S, dP
    constexpr uint32_t idesc_mm = make_idesc_bf16(128, 128, true, true);     // dV, dK
    constexpr uint32_t idesc_km = make_idesc_bf16(128, 128, false, true);    // dQ
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
    const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
    mbar_wait(kv_full, 0);
    for (int t = 0; t < steps; ++t) {
      mbar_wait(qdo_full, t & 1);
      mbar_wait(s_free, (t & 1) ^ 1);            // the previous step's dQ has left TMEM [0,128)
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16<1>(tmem_s, desc_kmajor(q_addr, kk), desc_kmajor(k_addr, kk), idesc_kk, kk != 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16<1>(tmem_dp, desc_kmajor(do_addr, kk), desc_kmajor(v_addr, kk), idesc_kk, kk != 0);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, t & 1);
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BM / 16; ++kk)     // reduction over the 128 query rows
          umma_f16<1>(tmem_dv, desc_mnmajor(p_addr, kk), desc_mnmajor(do_addr, kk), idesc_mm, (t | kk) != 0);
#pragma unroll
        for (int kk = 0; kk < BM / 16; ++kk)
          umma_f16<1>(tmem_dk, desc_mnmajor(ds_addr, kk), desc_mnmajor(q_addr, kk), idesc_mm, (t | kk) != 0);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)     // reduction over the 128 keys
          umma_f16<1>(tmem_s, desc_kmajor(ds_addr, kk), desc_mnmajor(k_addr, kk), idesc_km, kk != 0);
        umma_commit(mma2_done);
        umma_commit(qdo_empty);
      }
      __syncwarp();
    }
  } else {
    // =========================== softmax / dS / dQ epilogue (thread = query row) ===========================
    const uint32_t qd = warp_idx & 3;
    const uint32_t row = qd * 32 + lane;
    const uint32_t lane_off = (qd * 32u) << 16;
    const bool issuer = threadIdx.x == 64;
    for (int t = 0; t < steps; ++t) {
      const int head = kv_head * rep + t / n_i;
      const int i = i_lo + t % n_i;
      const int q_pos = i * BM + (int)row;
      const size_t stat = ((size_t)batch * p.Hq + head) * p.S + q_pos;
      const float lse2 = p.lse[stat] * 1.4426950408889634f;
      const float dl = p.delta[stat];
      const bool dead_row = !(lse2 > -INFINITY);
      const bool need_mask = (i == j) || (p.window > 0 && k0 <= q_pos - p.window);
      mbar_wait(sdp_full, t & 1);
      tcgen05_fence_after();
      uint8_t* prow = sP + row * 128;
      uint8_t* dsrow = sdS + row * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + c * 32, sv);
        tmem_ld_32x32b_x32(tmem_dp + lane_off + c * 32, dv);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pe[8], de[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = c * 32 + g * 8 + e;
            float pv = exp2f(__uint_as_float(sv[g * 8 + e]) * p.scale_log2 - lse2);
            if (need_mask) {
              const int kpos = k0 + col;
              if (kpos > q_pos || (p.window > 0 && kpos <= q_pos - p.window)) pv = 0.f;
            }
            if (dead_row) pv = 0.f;
            pe[e] = pv;
            de[e] = pv * (__uint_as_float(dv[g * 8 + e]) - dl) * p.scale;
          }
          uint4 po, dso;
          po.x = pack_bf16x2(pe[0], pe[1]); po.y = pack_bf16x2(pe[2], pe[3]);
          po.z = pack_bf16x2(pe[4], pe[5]); po.w = pack_bf16x2(pe[6], pe[7]);
          dso.x = pack_bf16x2(de[0], de[1]); dso.y = pack_bf16x2(de[2], de[3]);
          dso.z = pack_bf16x2(de[4], de[5]); dso.w = pack_bf16x2(de[6], de[7]);
          const int c16 = c * 4 + g;                      // 16-byte column group 0..15
          const int chunk = c16 >> 3, j16 = c16 & 7;
          const int off = chunk * CHUNK_BYTES + ((j16 ^ (row & 7)) * 16);
          *reinterpret_cast<uint4*>(prow + off) = po;
          *reinterpret_cast<uint4*>(dsrow + off) = dso;
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) mbar_arrive(pds_full);
      // ---- dQ_i: TMEM -> fp32 staging (P / dS tiles are free once the second MMA group retired) -> reduce-add ----
      mbar_wait(mma2_done, t & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + c * 32, v);
        tmem_ld_wait();
        uint8_t* cb = sP + c * DQ_CHUNK_BYTES + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint4 o = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          *reinterpret_cast<uint4*>(cb + ((g ^ (row & 7)) * 16)) = o;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_reduce_add_2d(&tma_dq, sP + c * DQ_CHUNK_BYTES, head * D + c * 32, row_base + i * BM);
        tma_store_commit();
        tma_store_wait_read<0>();               // staging may be rewritten by the next step's P / dS
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // ---------------- dV_j, dK_j: TMEM -> bf16 -> smem (Q / dO buffers are free) -> TMA store ----------------
    // (the last step's mma2_done has been waited for above, so every accumulation has retired)
    tcgen05_fence_after();
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t src = which == 0 ? tmem_dv : tmem_dk;
      uint8_t* orow = (which == 0 ? sQ : sdO) + row * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(src + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
          const int c16 = c * 4 + g;
          const int chunk = c16 >> 3, j16 = c16 & 7;
          *reinterpret_cast<uint4*>(orow + chunk * CHUNK_BYTES + ((j16 ^ (row & 7)) * 16)) = o;
        }
      }
    }
    tcgen05_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (issuer) {
      tma_store_2d(&tma_dv, sQ, kv_head * D, row_base + k0);
      tma_store_2d(&tma_dv, sQ + CHUNK_BYTES, kv_head * D + 64, row_base + k0);
      tma_store_2d(&tma_dk, sdO, kv_head * D, row_base + k0);
      tma_store_2d(&tma_dk, sdO + CHUNK_BYTES, kv_head * D + 64, row_base + k0);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace fab100

// dout / q / out [B,S,Hq,128], k / v [B,S,Hk,128] bf16 contiguous, lse [B,Hq,S] fp32 (natural log, from the forward).
// Returns (dq [B,S,Hq,128], dk, dv [B,S,Hk,128]) in bf16.
std::vector<at::Tensor> flash_attn_bwd_preview(const at::Tensor& dout, const at::Tensor& q, const at::Tensor& k,
                                               const at::Tensor& v, const at::Tensor& out, const at::Tensor& lse,
                                               bool causal, int64_t window, double scale) {
  using namespace fab100;
  TORCH_CHECK(causal, "flash_attn_bwd: only causal attention is implemented");
  for (const at::Tensor* t : {&dout, &q, &k, &v, &out})
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kBFloat16 && t->is_contiguous() && t->dim() == 4 && t->size(3) == D,
                "flash_attn_bwd: contiguous bf16 [B,S,H,128] tensors required");
  const int B = q.size(0), S = q.size(1), Hq = q.size(2), Hk = k.size(2);
  TORCH_CHECK(k.size(1) == S && v.size(1) == S && v.size(2) == Hk && Hq % Hk == 0 && S % BM == 0);
  TORCH_CHECK(dout.sizes() == q.sizes() && out.sizes() == q.sizes());
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == (int64_t)B * Hq * S);
  c10::cuda::CUDAGuard guard(q.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto delta = at::empty({B, Hq, S}, q.options().dtype(at::kFloat));
  {
    const int64_t rows = (int64_t)B * S * Hq;
    const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
    flash_bwd_delta_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)dout.data_ptr(),
                                                     (const __nv_bfloat16*)out.data_ptr(), delta.data_ptr<float>(),
                                                     B, S, Hq);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  auto dq_acc = at::zeros({B, S, Hq, D}, q.options().dtype(at::kFloat));
  auto dk = at::empty_like(k), dv = at::empty_like(v);
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const uint64_t rows = (uint64_t)B * S;
  CUtensorMap tq = tma::make_2d(q.data_ptr(), BF, 2, rows, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, 128, 64);
  CUtensorMap tdo = tma::make_2d(dout.data_ptr(), BF, 2, rows, (uint64_t)Hq * D, (uint64_t)Hq * D * 2, 128, 64);
  CUtensorMap tk = tma::make_2d(k.data_ptr(), BF, 2, rows, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tv = tma::make_2d(v.data_ptr(), BF, 2, rows, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tdk = tma::make_2d(dk.data_ptr(), BF, 2, rows, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tdv = tma::make_2d(dv.data_ptr(), BF, 2, rows, (uint64_t)Hk * D, (uint64_t)Hk * D * 2, 128, 64);
  CUtensorMap tdq = tma::make_2d(dq_acc.data_ptr(), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, (uint64_t)Hq * D,
                                 (uint64_t)Hq * D * 4, 128, 32);
  Params p;
  p.B = B; p.S = S; p.Hq = Hq; p.Hk = Hk; p.num_blocks = S / BM; p.window = (int)window;
  p.scale = (float)scale;
  p.scale_log2 = (float)(scale * 1.4426950408889634);
  p.lse = lse.data_ptr<float>();
  p.delta = delta.data_ptr<float>();
  static bool once = false;
  if (!once) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  dim3 grid(p.num_blocks, Hk, B);      // key block 0 has the longest query loop and is scheduled first
  flash_bwd_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {dq_acc.to(at::kBFloat16), dk, dv};
}
