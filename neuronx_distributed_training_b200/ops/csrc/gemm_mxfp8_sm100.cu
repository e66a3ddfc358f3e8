// MX block-scaled fp8 GEMM for sm_100a: `tcgen05.mma.kind::mxf8f6f4.block_scale` with 1 x 32 E8M0 scale factors.
//
//   out[M, N] (+)= sum_kb  2^(sfa[m, kb] - 127) * 2^(sfb[n, kb] - 127) * sum_{k in block kb} A_q[m, k] * B_q[n, k]
//
// A_q [M, K], B_q [N, K]: e4m3 bytes, K-major.  sfa / sfb: one E8M0 byte per row and 32-element K block, stored in the
// layout the tensor core consumes: tiles of 128 rows x 4 K blocks = 512 bytes, byte (r % 32) * 16 + (r / 32) * 4 + (kb % 4),
// tiles ordered [row tile][K tile] (`ops/fp8.py:pack_mx_scales`, the same atom CUTLASS calls Sm1xxBlockScaledBasicChunk).
// The scaling happens INSIDE the tensor core: per K block the issuer copies the 512-byte scale tiles of the stage from shared
// memory into tensor memory (`tcgen05.cp.32x128b.warpx4`, one 16-byte line per lane, broadcast to the four lane quadrants)
// and the four K = 32 MMAs of the block name their scale byte through the scale-factor id of the instruction descriptor.
// BASELINE config 4 ("block-scaled fp8 on the GEMM"); the reference has no fp8 mode.
//
// Persistent, warp specialised: warp 0 TMA producer (A 16 KB + B 32 KB + 1.5 KB of scales per stage, 4 stages), warp 1
// TMEM allocator + single-thread issuer, warps 2..5 epilogue (TMEM -> registers -> swizzled staging -> TMA store / reduce).
// Tile 128 x 256 x 128; one accumulator buffer (256 TMEM columns) + 4 x (4 + 8) columns of scale factors.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "sm100_ptx.cuh"
#include "tma_utils.h"

namespace mxgemm {

using namespace sm100;

constexpr int BM = 128, BN = 256, BK = 128;           // BK in fp8 elements = bytes per swizzle row
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK;   // 16 KB, 32 KB
constexpr int SF_ATOM = 512;                          // 128 rows x 4 K blocks
constexpr int SFA_BYTES = SF_ATOM, SFB_BYTES = 2 * SF_ATOM;
constexpr int C_BYTES = BM * 128;                     // staging: 128 rows x 128 bytes (64 bf16 or 32 fp32 columns)
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES + SFA_BYTES + SFB_BYTES) + C_BYTES + 1024 + 256;
constexpr int TMEM_SFA = 256, TMEM_SFB = 256 + 4 * STAGES;      // column offsets; 4 (SFA) / 8 (SFB) columns per stage
constexpr int TMEM_COLS = 512;

enum OutKind : int { OUT_BF16_STORE = 0, OUT_F32_STORE = 1, OUT_F32_ADD = 2 };

struct Params {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_tiles, num_k_blocks;
  int group_m;
  const uint8_t* sfa;      // [M / 128][K / 128][512]
  const uint8_t* sfb;      // [ceil(N / 128)][K / 128][512]
};

// instruction descriptor of kind::mxf8f6f4.block_scale: e4m3 x e4m3, E8M0 scales, K-major operands
__device__ __forceinline__ uint32_t idesc_mx(uint32_t sf_id) {
  return (sf_id << 4)                                   // B scale-factor id (byte of the scale word)
         | (0u << 7) | (0u << 10)                       // A, B format: E4M3
         | (static_cast<uint32_t>(BN >> 3) << 17) | (1u << 23)   // N, scale format E8M0
         | (static_cast<uint32_t>(BM >> 4) << 24) | (sf_id << 29);   // M, A scale-factor id
}
__device__ __forceinline__ void umma_mx(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                        uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// shared memory -> tensor memory: 32 lines of 16 bytes, replicated to the four 32-lane quadrants
__device__ __forceinline__ void utccp_32x128b_warpx4(uint32_t tmem_dst, uint64_t smem_desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}
// K-major, no swizzle: 8-row core matrices of 16-byte lines, 128 bytes apart
__device__ __forceinline__ uint64_t sf_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(16 >> 4) << 16;       // leading byte offset (unused: one 16-byte column)
  d |= static_cast<uint64_t>(128 >> 4) << 32;      // stride between 8-line groups
  d |= 1ull << 46;                                 // descriptor version (sm_100)
  return d;                                        // layout type 0 = no swizzle
}

template <int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_BYTES;
  uint8_t* sC = sB + STAGES * B_BYTES;
  uint8_t* sSFA = sC + C_BYTES;
  uint8_t* sSFB = sSFA + STAGES * SFA_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSFB + STAGES * SFB_BYTES);
  uint64_t* full = bars;                      // [STAGES]
  uint64_t* empty = bars + STAGES;            // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;     // 1
  uint64_t* acc_empty = acc_full + 1;         // 1 (4 epilogue warps arrive)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b); tma_prefetch_desc(&tma_c);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc<1>(tmem_ptr, TMEM_COLS);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
    const int per_group = p.group_m * p.num_n_blocks;
    const int g = t / per_group, first_m = g * p.group_m;
    const int gsz = min(p.group_m, p.num_m_blocks - first_m);
    const int r = t - g * per_group;
    m_blk = first_m + r % gsz;
    n_blk = r / gsz;
  };

  if (warp_idx == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(t, m_blk, n_blk);
        const bool second_atom = n_blk * 2 + 1 < (p.N + 127) / 128;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], A_BYTES + B_BYTES + SFA_BYTES + (second_atom ? SFB_BYTES : SF_ATOM));
          tma_load_2d(sA + stage * A_BYTES, &tma_a, &full[stage], kb * BK, m_blk * BM);
          tma_load_2d(sB + stage * B_BYTES, &tma_b, &full[stage], kb * BK, n_blk * BN);
          bulk_g2s(sSFA + stage * SFA_BYTES, p.sfa + ((size_t)m_blk * p.num_k_blocks + kb) * SF_ATOM, SF_ATOM, &full[stage]);
          bulk_g2s(sSFB + stage * SFB_BYTES, p.sfb + ((size_t)(n_blk * 2) * p.num_k_blocks + kb) * SF_ATOM, SF_ATOM, &full[stage]);
          if (second_atom)
            bulk_g2s(sSFB + stage * SFB_BYTES + SF_ATOM, p.sfb + ((size_t)(n_blk * 2 + 1) * p.num_k_blocks + kb) * SF_ATOM,
                     SF_ATOM, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // =========================== issuer (one thread) ===========================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, tile_i = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile_i) {
        mbar_wait(acc_empty, (tile_i & 1) ^ 1);          // the epilogue has read the previous tile out of tensor memory
        tcgen05_fence_after();
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * A_BYTES), b_addr = smem_u32(sB + stage * B_BYTES);
          const uint32_t t_sfa = tmem_base + TMEM_SFA + stage * 4, t_sfb = tmem_base + TMEM_SFB + stage * 8;
          utccp_32x128b_warpx4(t_sfa, sf_smem_desc(smem_u32(sSFA + stage * SFA_BYTES)));
          utccp_32x128b_warpx4(t_sfb, sf_smem_desc(smem_u32(sSFB + stage * SFB_BYTES)));
          utccp_32x128b_warpx4(t_sfb + 4, sf_smem_desc(smem_u32(sSFB + stage * SFB_BYTES + SF_ATOM)));
#pragma unroll
          for (uint32_t kk = 0; kk < BK / 32; ++kk) {
            // 32 fp8 elements = 32 bytes along the 128-byte swizzle row; scale byte kk of the stage's scale words
            umma_mx(tmem_base, make_smem_desc_sw128(a_addr + kk * 32, 16, 1024), make_smem_desc_sw128(b_addr + kk * 32, 16, 1024),
                    idesc_mx(kk), (kb | (int)kk) != 0, t_sfa + (kk << 30), t_sfb + (kk << 30));
          }
          umma_commit(&empty[stage]);                    // operands AND scale tiles of the stage are free once these retire
          if (kb + 1 == p.num_k_blocks) umma_commit(acc_full);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    const uint32_t qd = warp_idx & 3;
    const uint32_t row = qd * 32 + lane;
    const uint32_t lane_off = (qd * 32u) << 16;
    const uint32_t crow = smem_u32(sC) + row * 128;
    uint32_t tile_i = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile_i) {
      int m_blk, n_blk;
      tile_coords(t, m_blk, n_blk);
      mbar_wait(acc_full, tile_i & 1);
      tcgen05_fence_after();
      constexpr int CHUNK = OUT == OUT_BF16_STORE ? 64 : 32;      // columns per 128-byte staging row
#pragma unroll 1
      for (int c = 0; c < BN / CHUNK; ++c) {
        if (n_blk * BN + c * CHUNK >= p.N) break;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + c * CHUNK, v0);
        if constexpr (CHUNK == 64) tmem_ld_32x32b_x32(tmem_base + lane_off + c * CHUNK + 32, v1);
        tmem_ld_wait();
        if (c + 1 == BN / CHUNK || n_blk * BN + (c + 1) * CHUNK >= p.N) {   // last read of this accumulator
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        // the previous TMA store must have finished reading the staging tile
        if (threadIdx.x == 64) tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint4 o;
          if constexpr (OUT == OUT_BF16_STORE) {
            const uint32_t* v = g < 4 ? v0 : v1;
            const int b = (g & 3) * 8;
            o.x = pack_bf16x2(__uint_as_float(v[b + 0]), __uint_as_float(v[b + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[b + 2]), __uint_as_float(v[b + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[b + 4]), __uint_as_float(v[b + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[b + 6]), __uint_as_float(v[b + 7]));
          } else {
            o = make_uint4(v0[4 * g], v0[4 * g + 1], v0[4 * g + 2], v0[4 * g + 3]);
          }
          st_shared_v4(crow + ((g ^ (row & 7)) * 16), o);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          if constexpr (OUT == OUT_F32_ADD) tma_reduce_add_2d(&tma_c, sC, n_blk * BN + c * CHUNK, m_blk * BM);
          else tma_store_2d(&tma_c, sC, n_blk * BN + c * CHUNK, m_blk * BM);
          tma_store_commit();
        }
      }
    }
    if (threadIdx.x == 64) tma_store_wait<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

}  // namespace mxgemm

// a_q [M, K], b_q [N, K] uint8 (e4m3), sfa / sfb packed E8M0 scale tiles ([rows/128][K/128][512] uint8); out [M, N] bf16
// (store) or fp32 (store / accumulate).  M % 128 == 0, K % 128 == 0, N % 8 == 0.
void gemm_mxfp8(const at::Tensor& a_q, const at::Tensor& sfa, const at::Tensor& b_q, const at::Tensor& sfb, at::Tensor& out,
                bool accumulate) {
  using namespace mxgemm;
  TORCH_CHECK(a_q.is_cuda() && a_q.scalar_type() == at::kByte && b_q.scalar_type() == at::kByte && a_q.is_contiguous() &&
              b_q.is_contiguous() && sfa.is_contiguous() && sfb.is_contiguous() && out.is_contiguous());
  const int M = a_q.size(0), K = a_q.size(1), N = b_q.size(0);
  TORCH_CHECK(b_q.size(1) == K && out.size(0) == M && out.size(1) == N);
  TORCH_CHECK(M % BM == 0 && K % BK == 0 && N % 8 == 0, "gemm_mxfp8: M % 128 == 0, K % 128 == 0, N % 8 == 0 required");
  TORCH_CHECK(sfa.scalar_type() == at::kByte && sfa.numel() == (int64_t)(M / 128) * (K / 128) * 512, "gemm_mxfp8: sfa size");
  TORCH_CHECK(sfb.scalar_type() == at::kByte && sfb.numel() == (int64_t)((N + 127) / 128) * (K / 128) * 512, "gemm_mxfp8: sfb size");
  const bool f32 = out.scalar_type() == at::kFloat;
  TORCH_CHECK(f32 || (out.scalar_type() == at::kBFloat16 && !accumulate), "gemm_mxfp8: bf16 store or fp32 store / accumulate");
  c10::cuda::CUDAGuard guard(a_q.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const auto U8 = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUtensorMap ta = tma::make_2d(a_q.data_ptr(), U8, 1, M, K, (uint64_t)K, BM, BK);
  CUtensorMap tb = tma::make_2d(b_q.data_ptr(), U8, 1, N, K, (uint64_t)K, BN, BK);
  CUtensorMap tc = f32 ? tma::make_2d(out.data_ptr(), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, (uint64_t)N * 4, BM, 32)
                       : tma::make_2d(out.data_ptr(), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, (uint64_t)N * 2, BM, 64);
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = M / BM; p.num_n_blocks = (N + BN - 1) / BN; p.num_tiles = p.num_m_blocks * p.num_n_blocks;
  p.num_k_blocks = K / BK;
  p.group_m = 8;
  p.sfa = sfa.data_ptr<uint8_t>(); p.sfb = sfb.data_ptr<uint8_t>();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = std::min(p.num_tiles, sms);
  auto launch = [&](auto kern) {
    static bool once = false;
    (void)once;
    C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p);
  };
  if (!f32) launch(gemm_mxfp8_kernel<OUT_BF16_STORE>);
  else if (accumulate) launch(gemm_mxfp8_kernel<OUT_F32_ADD>);
  else launch(gemm_mxfp8_kernel<OUT_F32_STORE>);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------------
// MX quantisation kernels: bf16 -> e4m3 payload + E8M0 scale per 1 x 32 block, scales written straight into the packed
// tile layout the GEMM consumes.  scale exponent e = ceil(log2(amax / 448)) (no saturation), stored as e + 127; an all-zero
// block gets the neutral 2^0.  `ops/fp8.py:quantize_mx` + `pack_mx_scales` is the bit-exact PyTorch reference.
// ---------------------------------------------------------------------------------------------------------------------
#include <cuda_fp8.h>

namespace mxgemm {

__device__ __forceinline__ uint32_t e8m0_for_amax(float amax) {
  if (!(amax > 0.f)) return 127u;
  const float s = amax * (1.f / 448.f);
  const uint32_t bits = __float_as_uint(s);
  uint32_t e = (bits >> 23) & 0xffu;                 // floor(log2(s)) + 127   (s is normal for every bf16 amax > 0)
  if (bits & 0x7fffffu) e += 1;                      // ceil
  return min(max(e, 1u), 254u);
}
__device__ __forceinline__ float inv_scale_of(uint32_t e8m0) { return __uint_as_float((254u - e8m0) << 23); }   // 2^(127 - e)
__device__ __forceinline__ size_t sf_offset(int r, int kb, int k_tiles) {
  return ((size_t)(r >> 7) * k_tiles + (kb >> 2)) * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (kb & 3);
}

// x [R, C] -> q [R, C], scales over (R, C / 32); one warp per 256-column segment of a row, 4 lanes per 32-block
__global__ void __launch_bounds__(256) mx_quant_rows_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                            uint8_t* __restrict__ sf, int R, int C) {
  const int lane = threadIdx.x & 31;
  const int segs = C / 256 + (C % 256 != 0);
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int k_tiles = C / 128;
  for (int64_t item = warp; item < (int64_t)R * segs; item += nwarps) {
    const int r = (int)(item / segs), c = (int)(item % segs) * 256 + lane * 8;
    float f[8];
    float amax = 0.f;
    if (c < C) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + (size_t)r * C + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
        amax = fmaxf(amax, fmaxf(fabsf(t.x), fabsf(t.y)));
      }
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    if (c < C) {
      const uint32_t e = e8m0_for_amax(amax);
      const float inv = inv_scale_of(e);
      alignas(8) __nv_fp8_e4m3 o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = __nv_fp8_e4m3(f[i] * inv);
      *reinterpret_cast<uint2*>(q + (size_t)r * C + c) = *reinterpret_cast<const uint2*>(o);
      if ((lane & 3) == 0) sf[sf_offset(r, c >> 5, k_tiles)] = (uint8_t)e;
    }
  }
}

// x [R, C] -> q = MX(x^T) [C, R] with blocks along R, scales over (C, R / 32).  One CTA per [32 rows x 128 columns] tile.
__global__ void __launch_bounds__(256) mx_quant_cols_transposed_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                                       uint8_t* __restrict__ sf, int R, int C) {
  __shared__ float tile[32][129];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 128;
  for (int u = threadIdx.x; u < 32 * 16; u += 256) {          // 16 eight-element units per row
    const int rr = u >> 4, cc = (u & 15) * 8;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 + cc < C) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + (size_t)(r0 + rr) * C + c0 + cc);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[rr][cc + i] = f[i];
  }
  __syncthreads();
  const int c = threadIdx.x >> 1, half = threadIdx.x & 1;     // two threads per output row (column c of x), 16 values each
  float v[16];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = tile[half * 16 + i][c];
    amax = fmaxf(amax, fabsf(v[i]));
  }
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  if (c0 + c < C) {
    const uint32_t e = e8m0_for_amax(amax);
    const float inv = inv_scale_of(e);
    alignas(16) __nv_fp8_e4m3 o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __nv_fp8_e4m3(v[i] * inv);
    *reinterpret_cast<uint4*>(q + (size_t)(c0 + c) * R + r0 + half * 16) = *reinterpret_cast<const uint4*>(o);
    if (half == 0) sf[sf_offset(c0 + c, r0 >> 5, R / 128)] = (uint8_t)e;
  }
}

}  // namespace mxgemm

// bf16 [R, C] (C % 128 == 0) -> (e4m3 bytes [R, C], packed E8M0 scale tiles [ceil(R / 128), C / 128, 512])
std::vector<at::Tensor> mx_quantize_rows(const at::Tensor& x) {
  using namespace mxgemm;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.is_contiguous() && x.dim() == 2);
  const int R = x.size(0), C = x.size(1);
  TORCH_CHECK(C % 128 == 0, "mx_quantize_rows: the reduction dimension must be a multiple of 128");
  c10::cuda::CUDAGuard guard(x.device());
  auto q = at::empty({R, C}, x.options().dtype(at::kByte));
  auto sf = at::full({(R + 127) / 128, C / 128, 512}, 127, x.options().dtype(at::kByte));
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int64_t items = (int64_t)R * ((C + 255) / 256);
  const int grid = (int)std::min<int64_t>((items + 7) / 8, (int64_t)sms * 16);
  if (R > 0)
    mx_quant_rows_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>((const __nv_bfloat16*)x.data_ptr(), q.data_ptr<uint8_t>(),
                                                                              sf.data_ptr<uint8_t>(), R, C);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {q, sf};
}

// bf16 [R, C] (R % 128 == 0) -> (e4m3 bytes of x^T [C, R] quantised along R, packed scale tiles [ceil(C / 128), R / 128, 512])
std::vector<at::Tensor> mx_quantize_cols_transposed(const at::Tensor& x) {
  using namespace mxgemm;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.is_contiguous() && x.dim() == 2);
  const int R = x.size(0), C = x.size(1);
  TORCH_CHECK(R % 128 == 0 && C % 8 == 0, "mx_quantize_cols_transposed: rows must be a multiple of 128, columns of 8");
  c10::cuda::CUDAGuard guard(x.device());
  auto q = at::empty({C, R}, x.options().dtype(at::kByte));
  auto sf = at::full({(C + 127) / 128, R / 128, 512}, 127, x.options().dtype(at::kByte));
  dim3 grid((C + 127) / 128, R / 32);
  mx_quant_cols_transposed_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>((const __nv_bfloat16*)x.data_ptr(),
                                                                                      q.data_ptr<uint8_t>(), sf.data_ptr<uint8_t>(), R, C);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {q, sf};
}
