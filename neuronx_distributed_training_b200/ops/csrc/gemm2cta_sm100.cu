// Two-SM (cta_group::2) variant of the persistent bf16 GEMM: a cluster of two CTAs computes one 256 x 256 output tile.
//
//   D[M,N] (+)= A * B      same operand layouts / epilogues as gemm_sm100.cuh
//
// Why: with one CTA per tile (128 x 256 x 64 per stage) every SM stages 48 KB of operands per 4.2 MFLOP; with the pair each
// SM stages its 128 rows of A and only HALF of the B tile (the tensor core of either SM reads the other half out of the
// peer's shared memory), 32 KB per stage for the same math: 1/3 less shared-memory fill traffic and L2 -> SM traffic, room for
// a 6-deep ring, and one instruction stream (the leader's) drives both tensor cores.
//
// Roles per CTA (192 threads):  warp 0 TMA producer (each CTA loads its own halves; the transaction bytes of BOTH CTAs are
// counted on the LEADER's full barrier);  warp 1 TMEM allocator (both CTAs, cta_group::2) + MMA issuer (leader only; its
// commits are multicast to the barriers of both CTAs);  warps 2..5 epilogue (each CTA drains its own 128 accumulator lanes
// and tells the leader's MMA warp when the accumulator stage is free).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "gemm_sm100.cuh"
#include "tma_utils.h"

namespace gemm2cta {

using namespace sm100;
using gemm100::OutKind;
using gemm100::OUT_BF16_STORE;
using gemm100::OUT_F32_STORE;
using gemm100::OUT_F32_ADD;
using gemm100::OUT_BF16_ADD;

constexpr int CTA_M = 128;               // accumulator rows per CTA
constexpr int TILE_M = 256;              // per cluster
constexpr int TILE_N = 256;
constexpr int HALF_N = TILE_N / 2;       // B rows staged per CTA
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 6;
constexpr int ACC_STAGES = 2;
constexpr int A_STAGE_BYTES = CTA_M * BLOCK_K * 2;    // 16 KB
constexpr int B_STAGE_BYTES = HALF_N * BLOCK_K * 2;   // 16 KB
constexpr int C_STAGE_BYTES = CTA_M * 128;
constexpr int C_BUFS = 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + C_BUFS * C_STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = ACC_STAGES * TILE_N;        // 512
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct Params {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_tiles;
  int group_m;
};

__device__ __forceinline__ void tile_coords(const Params& p, int t, int& m_blk, int& n_blk) {
  const int per_group = p.group_m * p.num_n_blocks;
  const int g = t / per_group;
  const int first_m = g * p.group_m;
  const int gsz = min(p.group_m, p.num_m_blocks - first_m);
  const int r = t - g * per_group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(p)), "r"(cta));
  return remote;
}

template <bool A_MN, bool B_MN, int OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2cta_kernel(const __grid_constant__ CUtensorMap tma_a_, const __grid_constant__ CUtensorMap tma_b_,
                const __grid_constant__ CUtensorMap tma_c_, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const CUtensorMap* tma_a = &tma_a_;
  const CUtensorMap* tma_b = &tma_b_;
  const CUtensorMap* tma_c = &tma_c_;
  const uint32_t warp_idx = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_c = smem_b + STAGES * B_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + C_BUFS * C_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]     used in the leader only
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]     per CTA, multicast commit
  uint64_t* tmem_full = bars + 2 * STAGES;         // [ACC_STAGES] per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [ACC_STAGES] used in the leader only: 8 epilogue warps arrive
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(tma_a);
    tma_prefetch_desc(tma_b);
    tma_prefetch_desc(tma_c);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc<2>(tmem_ptr, TMEM_COLS);
    tmem_relinquish<2>();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                              // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int first_tile = blockIdx.x >> 1;
  const int tile_stride = gridDim.x >> 1;

  if (warp_idx == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int t = first_tile; t < p.num_tiles; t += tile_stride) {
        int m_blk, n_blk;
        tile_coords(p, t, m_blk, n_blk);
        const int m0 = m_blk * TILE_M + (int)cta_rank * CTA_M;
        const int n0 = n_blk * TILE_N + (int)cta_rank * HALF_N;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // the leader expects the bytes of both CTAs; the peer's copies complete on the leader's barrier
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + B_STAGE_BYTES));
          const uint32_t bar = map_to_cta(&full_bar[stage], 0);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          const int k0 = kb * BLOCK_K;
          if constexpr (!A_MN) {
            tma_load_2d_cta_group2(sa, tma_a, bar, k0, m0);                              // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int i = 0; i < CTA_M / 64; ++i)                                         // boxes {64 m, 64 k}
              tma_load_2d_cta_group2(sa + i * (64 * BLOCK_K * 2), tma_a, bar, m0 + i * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_cta_group2(sb, tma_b, bar, k0, n0);                              // box {64 k, 128 n}
          } else {
#pragma unroll
            for (int i = 0; i < HALF_N / 64; ++i)                                        // boxes {64 n, 64 k}
              tma_load_2d_cta_group2(sb + i * (64 * BLOCK_K * 2), tma_b, bar, n0 + i * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, TILE_N, A_MN, B_MN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = first_tile; t < p.num_tiles; t += tile_stride) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TILE_N;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
            const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024)
                                       : make_smem_desc_sw128(a_addr + k * (UMMA_K * 2), 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024)
                                       : make_smem_desc_sw128(b_addr + k * (UMMA_K * 2), 16, 1024);
              umma_f16<2>(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage], 3);                  // the slot is free in BOTH CTAs
            if (kb == num_k_blocks - 1) umma_commit_2cta(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =========================== epilogue (both CTAs, own 128 rows) ===========================
    const uint32_t q = warp_idx & 3;
    const uint32_t row = q * 32 + lane;
    const bool issuer = (warp_idx == 2 && lane == 0);
    constexpr bool F32 = (OUT == OUT_F32_STORE || OUT == OUT_F32_ADD);
    constexpr int COLS_PER_CHUNK = F32 ? 32 : 64;
    constexpr int NCHUNK = TILE_N / COLS_PER_CHUNK;
    uint32_t acc = 0, acc_phase = 0, cbuf = 0;
    for (int t = first_tile; t < p.num_tiles; t += tile_stride) {
      int m_blk, n_blk;
      tile_coords(p, t, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * TILE_N;
      const int cm = m_blk * TILE_M + (int)cta_rank * CTA_M;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        if (issuer) tma_store_wait_read<C_BUFS - 1>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t cb = smem_u32(smem_c + cbuf * C_STAGE_BYTES + row * 128);
        if constexpr (F32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(cb + ((j ^ (row & 7)) * 16), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        } else {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(taddr + c * 64, v0);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v0[8 * j + 0]), __uint_as_float(v0[8 * j + 1]));
            o.y = pack_bf16x2(__uint_as_float(v0[8 * j + 2]), __uint_as_float(v0[8 * j + 3]));
            o.z = pack_bf16x2(__uint_as_float(v0[8 * j + 4]), __uint_as_float(v0[8 * j + 5]));
            o.w = pack_bf16x2(__uint_as_float(v0[8 * j + 6]), __uint_as_float(v0[8 * j + 7]));
            st_shared_v4(cb + ((j ^ (row & 7)) * 16), o);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v1[8 * j + 0]), __uint_as_float(v1[8 * j + 1]));
            o.y = pack_bf16x2(__uint_as_float(v1[8 * j + 2]), __uint_as_float(v1[8 * j + 3]));
            o.z = pack_bf16x2(__uint_as_float(v1[8 * j + 4]), __uint_as_float(v1[8 * j + 5]));
            o.w = pack_bf16x2(__uint_as_float(v1[8 * j + 6]), __uint_as_float(v1[8 * j + 7]));
            st_shared_v4(cb + (((j + 4) ^ (row & 7)) * 16), o);
          }
        }
        if (c == NCHUNK - 1) {
          // this CTA's half of the accumulator stage is drained: tell the leader's MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          const int cn = n_blk * TILE_N + c * COLS_PER_CHUNK;
          if (cn < p.N && cm < p.M) {
            if constexpr (OUT == OUT_F32_ADD || OUT == OUT_BF16_ADD)
              tma_reduce_add_2d(tma_c, smem_c + cbuf * C_STAGE_BYTES, cn, cm);
            else
              tma_store_2d(tma_c, smem_c + cbuf * C_STAGE_BYTES, cn, cm);
          }
          tma_store_commit();
        }
        cbuf ^= 1;
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) tma_store_wait<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                              // nobody leaves while the pair still reads its smem / signals its barriers
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<2>(tmem_base, TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN, int OUT>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p, int grid,
                   cudaStream_t stream) {
  auto kern = gemm2cta_kernel<A_MN, B_MN, OUT>;
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace gemm2cta

int64_t gemm_get_cta_limit();

// out (+)= op(a) op(b) on SM pairs.  Same contract as gemm(); the caller decides when the pair kernel pays (large M, N).
void gemm_2cta(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, bool a_kmajor, bool b_kmajor, bool accumulate) {
  using namespace gemm2cta;
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && out.is_cuda(), "gemm_2cta: CUDA tensors required");
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16, "gemm_2cta: bf16 inputs required");
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && out.is_contiguous(), "gemm_2cta: contiguous tensors required");
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && out.dim() == 2);
  const int M = a_kmajor ? a.size(0) : a.size(1);
  const int K = a_kmajor ? a.size(1) : a.size(0);
  const int N = b_kmajor ? b.size(0) : b.size(1);
  TORCH_CHECK((b_kmajor ? b.size(1) : b.size(0)) == K, "gemm_2cta: K mismatch");
  TORCH_CHECK(out.size(0) == M && out.size(1) == N, "gemm_2cta: output shape mismatch");
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0 && M % 8 == 0, "gemm_2cta: dims must be multiples of 8 (16-byte TMA pitch)");
  const bool f32 = out.scalar_type() == at::kFloat;
  TORCH_CHECK(f32 || out.scalar_type() == at::kBFloat16, "gemm_2cta: out must be bf16 or fp32");
  c10::cuda::CUDAGuard guard(a.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta = a_kmajor ? tma::make_2d(a.data_ptr(), BF, 2, M, K, (uint64_t)K * 2, CTA_M, 64)
                            : tma::make_2d(a.data_ptr(), BF, 2, K, M, (uint64_t)M * 2, BLOCK_K, 64);
  CUtensorMap tb = b_kmajor ? tma::make_2d(b.data_ptr(), BF, 2, N, K, (uint64_t)K * 2, HALF_N, 64)
                            : tma::make_2d(b.data_ptr(), BF, 2, K, N, (uint64_t)N * 2, BLOCK_K, 64);
  CUtensorMap tc = f32 ? tma::make_2d(out.data_ptr(), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, (uint64_t)N * 4, CTA_M, 32)
                       : tma::make_2d(out.data_ptr(), BF, 2, M, N, (uint64_t)N * 2, CTA_M, 64);
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + TILE_M - 1) / TILE_M;
  p.num_n_blocks = (N + TILE_N - 1) / TILE_N;
  p.num_tiles = p.num_m_blocks * p.num_n_blocks;
  p.group_m = 8;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int limit = (int)gemm_get_cta_limit();
  const int pairs = std::max(1, (limit > 0 ? std::min(limit, sms) : sms) / 2);
  const int grid = 2 * std::min(p.num_tiles, pairs);
  const int kind = f32 ? (accumulate ? OUT_F32_ADD : OUT_F32_STORE) : (accumulate ? OUT_BF16_ADD : OUT_BF16_STORE);
#define DISPATCH(AMN, BMN)                                                                      \
  switch (kind) {                                                                               \
    case OUT_BF16_STORE: launch<AMN, BMN, OUT_BF16_STORE>(ta, tb, tc, p, grid, stream); break;  \
    case OUT_F32_STORE: launch<AMN, BMN, OUT_F32_STORE>(ta, tb, tc, p, grid, stream); break;    \
    case OUT_F32_ADD: launch<AMN, BMN, OUT_F32_ADD>(ta, tb, tc, p, grid, stream); break;        \
    default: launch<AMN, BMN, OUT_BF16_ADD>(ta, tb, tc, p, grid, stream); break;                \
  }
  if (a_kmajor && b_kmajor) { DISPATCH(false, false) }
  else if (a_kmajor && !b_kmajor) { DISPATCH(false, true) }
  else if (!a_kmajor && !b_kmajor) { DISPATCH(true, true) }
  else { DISPATCH(true, false) }
#undef DISPATCH
}
