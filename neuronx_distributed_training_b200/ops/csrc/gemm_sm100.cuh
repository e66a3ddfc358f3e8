// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (fp32 accumulators
// in TMEM, double buffered) -> tcgen05.ld epilogue -> swizzled smem -> TMA store / TMA reduce-add.
//
//   D[M,N] (+)= A * B      A: [M,K] (K-major) or stored [K,M] (MN-major)
//                          B: stored [N,K] (K-major) or stored [K,N] (MN-major)
//
// One kernel serves the three GEMMs of a linear layer without any transposes:
//   forward  Y  = X  W^T   (A K-major,  B K-major)
//   dgrad    dX = dY W     (A K-major,  B MN-major)
//   wgrad    dW += dY^T X  (A MN-major, B MN-major, fp32 reduce-add epilogue into main_grad)
//
// Warp roles (192 threads, 1 CTA / SM, grid = #SMs, static round-robin tile schedule with M-grouping for L2):
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM lane quarter = warp_idx % 4)
// This one-CTA-per-tile body is what the fused TP kernels (fused_tp.cu), the grouped expert GEMMs and the fp8 GEMM instantiate;
// plain `gemm()` calls with more than one 128-row / 128-column block go to the SM-pair kernel (gemm2cta_sm100.cu).
// The optional `TileGate` hook lets the fused all-gather->GEMM kernel make the producer wait (acquire at
// system scope) until the rows of A that a tile needs have arrived from the peer GPU.
#pragma once
#include "sm100_ptx.cuh"

namespace gemm100 {

using namespace sm100;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int C_STAGE_BYTES = BLOCK_M * 128;           // 128 rows x 128 B (64 bf16 or 32 fp32 columns)
constexpr int C_BUFS = 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + C_BUFS * C_STAGE_BYTES + 1024 /*align slack*/ + 256;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;        // 512

enum OutKind : int { OUT_BF16_STORE = 0, OUT_F32_STORE = 1, OUT_F32_ADD = 2, OUT_BF16_ADD = 3, OUT_BF16_PEER = 4 };

struct Params {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_tiles;
  int group_m;   // tile-order grouping along M (L2 reuse of B)
  int m_first;   // rotation of the m-block order (fused collectives start at a rank-dependent shard)
  int k_first;   // rotation of the k-block order (all-gather fused into the K loop of wgrad)
};

// Hook object threaded through the kernel.  The plain GEMM uses NoGate; the fused collectives provide
//   wait_m / wait_k : block the TMA producer until the rows of the gathered operand have arrived
//   out_map/out_row : tensor map + row of a tile in the owner rank's staging buffer (GEMM -> reduce-scatter)
//   tile_done       : publish a finished tile to its owner rank
struct NoGate {
  __device__ __forceinline__ void wait_m(int /*m_blk*/) const {}
  __device__ __forceinline__ void wait_k(int /*k_blk*/) const {}
  __device__ __forceinline__ void before_first_store() const {}
  __device__ __forceinline__ const CUtensorMap* out_map(int /*m_blk*/) const { return nullptr; }
  __device__ __forceinline__ int out_row(int m_blk) const { return m_blk * BLOCK_M; }
  __device__ __forceinline__ void tile_done(int /*m_blk*/, int /*n_blk*/) const {}
};

// Grouped (per-expert) GEMMs specialise this trait.  A grouped gate provides
//   bool tile_info(m_blk, a_m0, b_group, kb_begin, kb_count)
// returning false for tiles that must be skipped (padding blocks, experts without tokens), the row coordinate
// of the A tile, the expert whose B matrix the tile multiplies (3-D tensor map when `b_3d`) and the K range.
template <class Gate>
struct gate_traits {
  static constexpr bool grouped = false;
  static constexpr bool b_3d = false;
  // fp8 (e4m3) K-major operands: 128 elements per 128-byte swizzle row, tcgen05.mma.kind::f8f6f4 (K = 32 per
  // instruction); the gate then also provides the dequantisation scales  out[m, n] = acc * row_scale(m) * col_scale(n)
  static constexpr bool fp8 = false;
};

__device__ __forceinline__ void tile_coords(const Params& p, int t, int& m_blk, int& n_blk) {
  const int per_group = p.group_m * p.num_n_blocks;
  const int g = t / per_group;
  const int first_m = g * p.group_m;
  const int gsz = min(p.group_m, p.num_m_blocks - first_m);
  const int r = t - g * per_group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
  m_blk += p.m_first;
  if (m_blk >= p.num_m_blocks) m_blk -= p.num_m_blocks;
}

template <bool A_MN, bool B_MN, int OUT, class TileGate = NoGate>
__device__ __forceinline__ void gemm_body(const CUtensorMap* __restrict__ tma_a, const CUtensorMap* __restrict__ tma_b,
                                          const CUtensorMap* __restrict__ tma_c, const Params& p, uint8_t* smem_raw,
                                          const TileGate& gate, int cta_first_tile, int cta_tile_stride) {
  using GT = gate_traits<TileGate>;
  const uint32_t warp_idx = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_c = smem_b + STAGES * B_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + C_BUFS * C_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;         // [ACC_STAGES]
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [ACC_STAGES]
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
      mbar_init(&tmem_empty[i], 4);   // one arrive per epilogue warp
    }
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
  constexpr int K_ELEMS = GT::fp8 ? 2 * BLOCK_K : BLOCK_K;      // elements per 128-byte operand row
  const int num_k_blocks = (p.K + K_ELEMS - 1) / K_ELEMS;

  if (warp_idx == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int t = cta_first_tile; t < p.num_tiles; t += cta_tile_stride) {
        int m_blk, n_blk;
        tile_coords(p, t, m_blk, n_blk);
        gate.wait_m(m_blk);
        int m0 = m_blk * BLOCK_M;
        const int n0 = n_blk * BLOCK_N;
        int b_group = 0, kb_begin = 0, tile_k_blocks = num_k_blocks;
        if constexpr (GT::grouped) {
          if (!gate.tile_info(m_blk, m0, b_group, kb_begin, tile_k_blocks)) continue;
        }
        for (int kbi = 0; kbi < tile_k_blocks; ++kbi) {
          int kb;
          if constexpr (GT::grouped) {
            kb = kb_begin + kbi;
          } else {
            kb = kbi + p.k_first;
            if (kb >= num_k_blocks) kb -= num_k_blocks;
          }
          gate.wait_k(kb);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          const int k0 = kb * K_ELEMS;
          if constexpr (!A_MN) {
            tma_load_2d(sa, tma_a, &full_bar[stage], k0, m0);                       // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i)                                  // boxes {64 m, 64 k}
              tma_load_2d(sa + i * (64 * BLOCK_K * 2), tma_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (GT::b_3d) {
            // stacked per-expert B: third coordinate selects the expert, OOB rows/cols zero-fill per expert
            if constexpr (!B_MN) {
              tma_load_3d(sb, tma_b, &full_bar[stage], k0, n0, b_group);
            } else {
#pragma unroll
              for (int i = 0; i < BLOCK_N / 64; ++i)
                tma_load_3d(sb + i * (64 * BLOCK_K * 2), tma_b, &full_bar[stage], n0 + i * 64, k0, b_group);
            }
          } else if constexpr (!B_MN) {
            tma_load_2d(sb, tma_b, &full_bar[stage], k0, n0);                       // box {64 k, 256 n}
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i)                                  // boxes {64 n, 64 k}
              tma_load_2d(sb + i * (64 * BLOCK_K * 2), tma_b, &full_bar[stage], n0 + i * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // =========================== MMA issuer ===========================
    static_assert(!GT::fp8 || (!A_MN && !B_MN), "the fp8 path takes K-major operands only");
    constexpr uint32_t idesc = GT::fp8 ? make_idesc_f8(BLOCK_M, BLOCK_N, false, false)
                                       : make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = cta_first_tile; t < p.num_tiles; t += cta_tile_stride) {
      int tile_k_blocks = num_k_blocks;
      if constexpr (GT::grouped) {
        int m_blk, n_blk, m0, b_group, kb_begin;
        tile_coords(p, t, m_blk, n_blk);
        if (!gate.tile_info(m_blk, m0, b_group, kb_begin, tile_k_blocks)) continue;
      }
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < tile_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 16 bf16 = 32 B inside the 128 B swizzle row.  MN-major: 16 k-rows of 128 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024)
                                     : make_smem_desc_sw128(a_addr + k * (UMMA_K * 2), 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * (UMMA_K * 128), 64 * BLOCK_K * 2, 1024)
                                     : make_smem_desc_sw128(b_addr + k * (UMMA_K * 2), 16, 1024);
            if constexpr (GT::fp8) umma_f8(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16<1>(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // smem slot reusable once these MMAs retire
          if (kb == tile_k_blocks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================== epilogue ===========================
    const uint32_t q = warp_idx & 3;                 // TMEM lane quarter this warp may touch
    const uint32_t row = q * 32 + lane;              // row inside the 128-row tile
    const bool issuer = (warp_idx == 2 && lane == 0);
    constexpr bool F32 = (OUT == OUT_F32_STORE || OUT == OUT_F32_ADD);
    constexpr int COLS_PER_CHUNK = F32 ? 32 : 64;
    constexpr int NCHUNK = BLOCK_N / COLS_PER_CHUNK;
    uint32_t acc = 0, acc_phase = 0, cbuf = 0;
    bool first_store = true;
    int prev_tile_m = -1, prev_tile_n = -1;
    for (int t = cta_first_tile; t < p.num_tiles; t += cta_tile_stride) {
      int m_blk, n_blk;
      tile_coords(p, t, m_blk, n_blk);
      if constexpr (GT::grouped) {
        int m0, b_group, kb_begin, nkb;
        if (!gate.tile_info(m_blk, m0, b_group, kb_begin, nkb)) continue;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * BLOCK_N;
      const CUtensorMap* cmap = tma_c;
      int cm = m_blk * BLOCK_M;
      if constexpr (OUT == OUT_BF16_PEER) {
        // GEMM -> reduce-scatter: the tile is TMA-stored into the staging buffer of the rank that owns these rows
        // (a tensor map over peer-mapped memory), full 128 B lines over NVLink.
        if (first_store) { gate.before_first_store(); first_store = false; }
        cmap = gate.out_map(m_blk);
        cm = gate.out_row(m_blk);
      }
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        if (issuer) tma_store_wait_read<C_BUFS - 1>();          // the buffer we are about to overwrite is drained
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint8_t* cb = smem_c + cbuf * C_STAGE_BYTES + row * 128;
        if constexpr (F32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
          if constexpr (GT::fp8) {
            const float rs = gate.row_scale(m_blk * BLOCK_M + (int)row);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(__uint_as_float(v[j]) * rs * gate.col_scale(n_blk * BLOCK_N + c * 32 + j));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 o = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<uint4*>(cb + ((j ^ (row & 7)) * 16)) = o;
          }
        } else {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(taddr + c * 64, v0);
          tmem_ld_32x32b_x32(taddr + c * 64 + 32, v1);
          tmem_ld_wait();
          if constexpr (GT::fp8) {
            const float rs = gate.row_scale(m_blk * BLOCK_M + (int)row);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v0[j] = __float_as_uint(__uint_as_float(v0[j]) * rs * gate.col_scale(n_blk * BLOCK_N + c * 64 + j));
              v1[j] = __float_as_uint(__uint_as_float(v1[j]) * rs * gate.col_scale(n_blk * BLOCK_N + c * 64 + 32 + j));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v0[8 * j + 0]), __uint_as_float(v0[8 * j + 1]));
            o.y = pack_bf16x2(__uint_as_float(v0[8 * j + 2]), __uint_as_float(v0[8 * j + 3]));
            o.z = pack_bf16x2(__uint_as_float(v0[8 * j + 4]), __uint_as_float(v0[8 * j + 5]));
            o.w = pack_bf16x2(__uint_as_float(v0[8 * j + 6]), __uint_as_float(v0[8 * j + 7]));
            *reinterpret_cast<uint4*>(cb + ((j ^ (row & 7)) * 16)) = o;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v1[8 * j + 0]), __uint_as_float(v1[8 * j + 1]));
            o.y = pack_bf16x2(__uint_as_float(v1[8 * j + 2]), __uint_as_float(v1[8 * j + 3]));
            o.z = pack_bf16x2(__uint_as_float(v1[8 * j + 4]), __uint_as_float(v1[8 * j + 5]));
            o.w = pack_bf16x2(__uint_as_float(v1[8 * j + 6]), __uint_as_float(v1[8 * j + 7]));
            *reinterpret_cast<uint4*>(cb + (((j + 4) ^ (row & 7)) * 16)) = o;
          }
        }
        if (c == NCHUNK - 1) {
          // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          const int cn = n_blk * BLOCK_N + c * COLS_PER_CHUNK;
          if (cn < p.N) {
            if constexpr (OUT == OUT_F32_ADD || OUT == OUT_BF16_ADD) {
              tma_reduce_add_2d(cmap, smem_c + cbuf * C_STAGE_BYTES, cn, cm);
            } else if constexpr (OUT == OUT_BF16_PEER) {
              // blocked staging layout (gate returns a destination): the [128 x 64] chunk leaves as ONE contiguous 16 KB
              // bulk copy, swizzle and all, instead of 128 separate 128-byte rows of a 2-D box; the owner un-swizzles
              uint8_t* dst = gate.peer_chunk(m_blk, cn);
              if (dst != nullptr) bulk_s2g(dst, smem_c + cbuf * C_STAGE_BYTES, C_STAGE_BYTES);
              else tma_store_2d(cmap, smem_c + cbuf * C_STAGE_BYTES, cn, cm);
            } else {
              tma_store_2d(cmap, smem_c + cbuf * C_STAGE_BYTES, cn, cm);
            }
          }
          tma_store_commit();
        }
        cbuf ^= 1;
      }
      if constexpr (OUT == OUT_BF16_PEER) {
        if (issuer) {
          // completion of the PREVIOUS tile's stores is awaited while this tile's are in flight (NCHUNK groups),
          // so the NVLink write latency hides behind the next tile's math
          if (prev_tile_m >= 0) {
            tma_store_wait<NCHUNK>();
            fence_proxy_async_all();
            __threadfence_system();
            gate.tile_done(prev_tile_m, prev_tile_n);    // publish the tile to the owner rank
          }
          prev_tile_m = m_blk;
          prev_tile_n = n_blk;
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) {
      tma_store_wait<0>();
      if constexpr (OUT == OUT_BF16_PEER) {
        if (prev_tile_m >= 0) {
          fence_proxy_async_all();
          __threadfence_system();
          gate.tile_done(prev_tile_m, prev_tile_n);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

}  // namespace gemm100
