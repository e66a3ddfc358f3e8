// Fused tensor-parallel kernels for sm_100a over NVSwitch peer memory.
//
//   all-gather -> GEMM     : a few "comm" CTAs pull the peers' sequence shards with TMA bulk copies
//                            (peer HBM -> smem ring -> local gathered buffer) and publish per-row-block ready flags;
//                            the remaining CTAs run the tcgen05 GEMM, whose TMA producer waits (acquire) on the
//                            flag of the rows a tile / k-block needs.  Tile order is rotated so every rank starts
//                            on its own shard and walks the ring, i.e. transfer and math overlap block by block.
//   GEMM -> reduce-scatter : the GEMM epilogue TMA-stores each finished partial tile straight into the staging
//                            buffer of the rank that owns those rows (one tensor map per destination over
//                            peer-mapped memory) and, one tile later, bumps that rank's per-row-block arrival
//                            counter (red.release.sys); after its last tile every CTA turns into a reducer that
//                            sums the TP partial slabs of the row-blocks that have fully arrived (fp32 accumulate)
//                            into the local output.
// Cross-GPU ordering: one monotonically increasing epoch per call (flags and handshake slots compare >=).  Kernels
// start with an all-peer handshake (st.release.sys of the epoch into every peer's slot, ld.acquire.sys on the own
// slots) which both publishes "my input shard is ready" and proves "the peer has finished the call that last used
// this double-buffer parity", so buffers are reused safely two calls later.  The epoch can live on the device
// (epoch_ptr, advanced by bump_epoch_kernel in stream order) so that the kernels replay correctly inside a CUDA
// graph; tp_barrier_kernel closes a captured region.  All cross-GPU spins are bounded and trap instead of hanging.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "gemm_sm100.cuh"
#include "tma_utils.h"
#include <cstdlib>
#include <set>

namespace gemm100 {
Params make_params(int M, int N, int K);
void make_operand_maps(const void* a, const void* b, int M, int N, int K, bool a_kmajor, bool b_kmajor,
                       CUtensorMap& ta, CUtensorMap& tb);
CUtensorMap make_out_map(void* out, int M, int N, bool f32);
}  // namespace gemm100

namespace fused {

using namespace gemm100;
constexpr int MAX_TP = 8;
constexpr int RING_SLOTS = 7;
constexpr int PIECE_BYTES = 32 * 1024;
constexpr int STORE_LAG = 3;

// Optional timeline capture: when `g_dbg` is non-null every CTA records globaltimer stamps
// [cta][0]=kernel start, [1]=handshake done, [2]=main phase done, [3]=kernel end.
__device__ unsigned long long* g_dbg = nullptr;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void dbg_stamp(int slot) {
  if (g_dbg != nullptr && (threadIdx.x == 0 || (slot == 1 && threadIdx.x == 64))) g_dbg[blockIdx.x * 4 + slot] = gtimer();
}

struct CommArgs {
  const uint8_t* x_local;          // my input shard [rows_per_rank, row_bytes] (ordinary device memory)
  uint8_t* xg_peer[MAX_TP];        // every rank's gathered buffer [tp * rows_per_rank, row_bytes] (symmetric, this parity)
  uint32_t* flags_peer[MAX_TP];    // every rank's ready flags, one per 128-row block of its gathered buffer
  uint32_t* hs_peer[MAX_TP];       // every rank's handshake array
  uint32_t* hs_local;              // my handshake array [tp]
  uint32_t* flags;                 // my ready flags (== flags_peer[rank])
  const uint8_t* x_peer[MAX_TP];   // pull mode: every rank's shard staged in its symmetric buffer (this parity)
  uint32_t epoch;
  const uint32_t* epoch_ptr;       // != nullptr: the epoch lives on the device (CUDA-graph replay safe), see bump_epoch
  int rank, tp, rows_per_rank, row_bytes, num_comm_ctas;
  int push;                        // 1: owners push their shard to everybody; 0: consumers pull the peers' shards
};

struct AgGate {
  const uint32_t* flags;
  uint32_t epoch;
  int mode;   // 0: gathered operand indexed by m-block, 1: by k-block (wgrad)
  int local_lo, local_hi;   // row-blocks of my own shard: already in place, never waited on
  int total_blocks;
  mutable unsigned long long ready_lo = 0, ready_hi = 0;   // per-thread cache: blocks already observed ready
  __device__ __forceinline__ void wait_flag(int blk) const {
    if (blk >= local_lo && blk < local_hi) return;
    // a block that has been seen ready stays ready for the rest of the call: poll each flag at most once
    const unsigned long long bit = 1ull << (blk & 63);
    unsigned long long& word = (blk < 64) ? ready_lo : ready_hi;
    if (blk < 128 && (word & bit)) return;
    spin_until_ge_sys(flags + blk, epoch, 64);
    fence_proxy_async_all();   // generic-proxy acquire -> subsequent async-proxy (TMA) reads
    if (blk < 128) word |= bit;
  }
  __device__ __forceinline__ void wait_m(int m_blk) const {
    if (mode == 0) wait_flag(m_blk);
  }
  __device__ __forceinline__ void wait_k(int kb) const {
    if (mode == 1) wait_flag(kb >> 1);   // BLOCK_K = 64 rows, flags are per 128 rows
  }
  __device__ __forceinline__ void before_first_store() const {}
  __device__ __forceinline__ const CUtensorMap* out_map(int) const { return nullptr; }
  __device__ __forceinline__ int out_row(int m_blk) const { return m_blk * BLOCK_M; }
  __device__ __forceinline__ void tile_done(int, int) const {}
};

__device__ __forceinline__ void send_handshake(uint32_t* const* hs_peer, int rank, int tp, uint32_t epoch) {
  for (int p = 0; p < tp; ++p) st_release_sys(hs_peer[p] + rank, epoch);
}

// PUSH-mode all-gather.  One thread per comm CTA drives: my shard --bulk--> smem ring --bulk--> every peer's gathered
// buffer.  Remote *stores* are posted (fire and forget), so the transfer runs at link rate instead of being bound by
// NVLink round-trip latency x bytes in flight as a pull (remote loads) would be.  After a 128-row block has been
// written everywhere, its ready flag is released (st.release.sys) on every peer.
__device__ void comm_cta_push(const CommArgs& a, uint8_t* smem_raw) {
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + RING_SLOTS * PIECE_BYTES);
  if (threadIdx.x != 0) return;
  dbg_stamp(0);
  if (blockIdx.x == 0) send_handshake(a.hs_peer, a.rank, a.tp, a.epoch);
  for (int i = 0; i < RING_SLOTS; ++i) mbar_init(&full[i], 1);
  fence_barrier_init();
  // a peer's gathered buffer of this parity was last read by its call (epoch - 2): it is free once that peer has
  // entered call (epoch - 1)
  for (int p = 0; p < a.tp; ++p) {
    if (p == a.rank) continue;
    spin_until_ge_sys(a.hs_local + p, (uint32_t)(a.epoch - 1), 100);
  }
  dbg_stamp(1);
  const int blocks_per_rank = a.rows_per_rank / 128;
  const size_t block_bytes = (size_t)128 * a.row_bytes;
  const int pieces = (int)((block_bytes + PIECE_BYTES - 1) / PIECE_BYTES);
  uint32_t slot_phase = 0;   // bit i = phase parity of ring slot i
  for (int lb = blockIdx.x; lb < blocks_per_rank; lb += a.num_comm_ctas) {
    const int blk = a.rank * blocks_per_rank + lb;
    const uint8_t* src = a.x_local + (size_t)lb * block_bytes;
    const size_t dst_off = (size_t)blk * block_bytes;
    auto piece_bytes = [&](int p) -> uint32_t {
      const size_t off = (size_t)p * PIECE_BYTES;
      return (uint32_t)min((size_t)PIECE_BYTES, block_bytes - off);
    };
    const int prime = min(pieces, RING_SLOTS);
    for (int p = 0; p < prime; ++p) {
      mbar_expect_tx(&full[p], piece_bytes(p));
      bulk_g2s(ring + p * PIECE_BYTES, src + (size_t)p * PIECE_BYTES, piece_bytes(p), &full[p]);
    }
    for (int p = 0; p < pieces; ++p) {
      const int s = p % RING_SLOTS;
      mbar_wait(&full[s], (slot_phase >> s) & 1);
      slot_phase ^= 1u << s;
      // the rank right below me consumes my rows first (its tile order walks rank+1, rank+2, ...)
      for (int d = 1; d < a.tp; ++d) {
        int peer = a.rank - d;
        if (peer < 0) peer += a.tp;
        bulk_s2g(a.xg_peer[peer] + dst_off + (size_t)p * PIECE_BYTES, ring + s * PIECE_BYTES, piece_bytes(p));
      }
      tma_store_commit();
      // refill the slot of piece (p - LAG): its stores have finished reading shared memory once at most LAG newer
      // store groups are still pending, so LAG stores and RING_SLOTS - LAG loads stay in flight
      const int r = p - STORE_LAG;
      if (r >= 0 && r + RING_SLOTS < pieces) {
        tma_store_wait_read<STORE_LAG>();
        const int rs = r % RING_SLOTS, np = r + RING_SLOTS;
        mbar_expect_tx(&full[rs], piece_bytes(np));
        bulk_g2s(ring + rs * PIECE_BYTES, src + (size_t)np * PIECE_BYTES, piece_bytes(np), &full[rs]);
      }
    }
    tma_store_wait<0>();           // every byte of this row-block has been written to every peer
    fence_proxy_async_all();
    __threadfence_system();
    for (int d = 1; d < a.tp; ++d) {
      int peer = a.rank - d;
      if (peer < 0) peer += a.tp;
      st_release_sys(a.flags_peer[peer] + blk, a.epoch);
    }
  }
  dbg_stamp(3);
}

// PULL-mode all-gather: every comm CTA copies peers' staged shards into MY gathered buffer
// (peer HBM --bulk--> smem ring --bulk--> local), publishing a local ready flag per 128-row block.
__device__ void comm_cta_pull(const CommArgs& a, uint8_t* smem_raw) {
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + RING_SLOTS * PIECE_BYTES);
  if (threadIdx.x != 0) return;
  dbg_stamp(0);
  if (blockIdx.x == 0) send_handshake(a.hs_peer, a.rank, a.tp, a.epoch);
  for (int i = 0; i < RING_SLOTS; ++i) mbar_init(&full[i], 1);
  fence_barrier_init();
  const int blocks_per_rank = a.rows_per_rank / 128;
  const int total_blocks = blocks_per_rank * a.tp;
  const size_t block_bytes = (size_t)128 * a.row_bytes;
  const int pieces = (int)((block_bytes + PIECE_BYTES - 1) / PIECE_BYTES);
  uint32_t slot_phase = 0;
  int waited_rank_mask = 0;
  uint8_t* xg = a.xg_peer[a.rank];
  bool stamped = false;
  for (int idx = blockIdx.x; idx < total_blocks; idx += a.num_comm_ctas) {
    int blk = idx + a.rank * blocks_per_rank;
    if (blk >= total_blocks) blk -= total_blocks;
    const int src_rank = blk / blocks_per_rank;
    if (src_rank == a.rank) continue;
    if (!((waited_rank_mask >> src_rank) & 1)) {
      spin_until_ge_sys(a.hs_local + src_rank, a.epoch, 100, 0x100u | (uint32_t)a.rank);
      waited_rank_mask |= 1 << src_rank;
    }
    if (!stamped) { dbg_stamp(1); stamped = true; }
    const uint8_t* src = a.x_peer[src_rank] + (size_t)(blk - src_rank * blocks_per_rank) * block_bytes;
    uint8_t* dst = xg + (size_t)blk * block_bytes;
    auto piece_bytes = [&](int p) -> uint32_t {
      const size_t off = (size_t)p * PIECE_BYTES;
      return (uint32_t)min((size_t)PIECE_BYTES, block_bytes - off);
    };
    const int prime = min(pieces, RING_SLOTS);
    for (int p = 0; p < prime; ++p) {
      mbar_expect_tx(&full[p], piece_bytes(p));
      bulk_g2s(ring + p * PIECE_BYTES, src + (size_t)p * PIECE_BYTES, piece_bytes(p), &full[p]);
    }
    for (int p = 0; p < pieces; ++p) {
      const int s = p % RING_SLOTS;
      mbar_wait(&full[s], (slot_phase >> s) & 1);
      slot_phase ^= 1u << s;
      bulk_s2g(dst + (size_t)p * PIECE_BYTES, ring + s * PIECE_BYTES, piece_bytes(p));
      tma_store_commit();
      const int r = p - STORE_LAG;
      if (r >= 0 && r + RING_SLOTS < pieces) {
        tma_store_wait_read<STORE_LAG>();
        const int rs = r % RING_SLOTS, np = r + RING_SLOTS;
        mbar_expect_tx(&full[rs], piece_bytes(np));
        bulk_g2s(ring + rs * PIECE_BYTES, src + (size_t)np * PIECE_BYTES, piece_bytes(np), &full[rs]);
      }
    }
    tma_store_wait<0>();
    fence_proxy_async_all();
    __threadfence();
    st_release_sys(a.flags + blk, a.epoch);
  }
  dbg_stamp(3);
}

__device__ __forceinline__ void comm_cta(const CommArgs& a, uint8_t* smem_raw) {
  if (a.push) comm_cta_push(a, smem_raw); else comm_cta_pull(a, smem_raw);
}

template <bool A_MN, bool B_MN, int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ag_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const Params p, const CommArgs comm_in, const int mode) {
  extern __shared__ uint8_t smem_raw[];
  CommArgs comm = comm_in;
  if (comm.epoch_ptr != nullptr) comm.epoch = ld_acquire_gpu(comm.epoch_ptr);
  if ((int)blockIdx.x < comm.num_comm_ctas) {
    comm_cta(comm, smem_raw);
    return;
  }
  const int bpr = comm.rows_per_rank / 128;
  AgGate gate{comm.flags, comm.epoch, mode, comm.rank * bpr, (comm.rank + 1) * bpr, bpr * comm.tp};
  dbg_stamp(0);
  gemm_body<A_MN, B_MN, OUT, AgGate>(&tma_a, &tma_b, &tma_c, p, smem_raw, gate, (int)blockIdx.x - comm.num_comm_ctas,
                                     (int)gridDim.x - comm.num_comm_ctas);
  dbg_stamp(3);
}

// pure all-gather (same comm CTAs, no GEMM): used where the gathered tensor is needed on its own
__global__ void __launch_bounds__(NUM_THREADS, 1) ag_only_kernel(const CommArgs comm_in) {
  extern __shared__ uint8_t smem_raw[];
  CommArgs comm = comm_in;
  if (comm.epoch_ptr != nullptr) comm.epoch = ld_acquire_gpu(comm.epoch_ptr);
  comm_cta(comm, smem_raw);
  // the kernel completes only when every peer's rows have landed in my gathered buffer
  const int bpr = comm.rows_per_rank / 128;
  for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < bpr * comm.tp; blk += gridDim.x * blockDim.x) {
    if (blk / bpr == comm.rank) continue;
    spin_until_ge_sys(comm.flags + blk, comm.epoch, 200, 0x200u | (uint32_t)comm.rank);
  }
}

// --------------------------------------------------------------------------------------------------
struct RsArgs {
  __nv_bfloat16* stage_peer[MAX_TP];   // every rank's staging buffer [tp][rows_per_rank][N] (this parity)
  uint32_t* counter_peer[MAX_TP];      // every rank's arrival counters [blocks_per_rank] (this parity)
  uint32_t* hs_peer[MAX_TP];
  uint32_t* hs_local;
  const __nv_bfloat16* stage_local;
  uint32_t* counter_local;
  __nv_bfloat16* out;                  // [rows_per_rank, N]
  uint32_t epoch;
  const uint32_t* epoch_ptr;
  int rank, tp, rows_per_rank, N, num_n_blocks;
};

struct RsMaps {
  CUtensorMap m[MAX_TP];   // one store map per destination rank: its staging buffer [tp * rows_per_rank, N]
};

// Protocol: arrival counters (one per row-block of the owner) cleared in front of the kernel, partial tiles stored with
// 2-D TMA into the owner's row-major staging slab, `red.release.sys` on the counter once a tile's stores have drained.
struct RsGate {
  RsArgs a;
  const CUtensorMap* maps;
  __device__ __forceinline__ void wait_m(int) const {}
  __device__ __forceinline__ void wait_k(int) const {}
  __device__ __forceinline__ void before_first_store() const {
    if (threadIdx.x == 64) {
      for (int p = 0; p < a.tp; ++p)
        spin_until_ge_sys(a.hs_local + p, a.epoch, 100, 0x300u | (uint32_t)a.rank);
      dbg_stamp(1);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  __device__ __forceinline__ const CUtensorMap* out_map(int m_blk) const {
    return maps + m_blk / (a.rows_per_rank / 128);
  }
  __device__ __forceinline__ int out_row(int m_blk) const {
    const int bpr = a.rows_per_rank / 128;
    const int d = m_blk / bpr;
    return a.rank * a.rows_per_rank + (m_blk - d * bpr) * 128;
  }
  __device__ __forceinline__ uint8_t* peer_chunk(int, int) const { return nullptr; }     // always the TMA-store path
  __device__ __forceinline__ void tile_done(int m_blk, int n_blk) const {
    const int bpr = a.rows_per_rank / 128;
    const int d = m_blk / bpr;
    (void)n_blk;
    red_add_release_sys(a.counter_peer[d] + (m_blk - d * bpr), 1u);
  }
};

template <bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_rs_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ RsMaps maps, const Params p, const RsArgs rs_in) {
  extern __shared__ uint8_t smem_raw[];
  RsArgs rs = rs_in;
  if (rs.epoch_ptr != nullptr) rs.epoch = ld_acquire_gpu(rs.epoch_ptr);
  if (blockIdx.x == 0 && threadIdx.x == 0) send_handshake(rs.hs_peer, rs.rank, rs.tp, rs.epoch);
  RsGate gate{rs, maps.m};
  const CUtensorMap& tma_c = maps.m[0];
  dbg_stamp(0);
  gemm_body<false, B_MN, OUT_BF16_PEER, RsGate>(&tma_a, &tma_b, &tma_c, p, smem_raw, gate, blockIdx.x, gridDim.x);
  dbg_stamp(2);
  // ---------------- reduce phase: sum the TP partial slabs of my rows ----------------
  const int bpr = rs.rows_per_rank / 128;
  const int n_chunks = (rs.N + 255) / 256;
  const uint32_t target = (uint32_t)(rs.tp * rs.num_n_blocks);
  for (int item = blockIdx.x; item < bpr * n_chunks; item += gridDim.x) {
    const int rb = item / n_chunks, nc = item - rb * n_chunks;
    if (threadIdx.x == 0) spin_until_ge_sys(rs.counter_local + rb, target, 200, 0x400u | (uint32_t)rs.rank);
    __syncthreads();
    // 128 rows x 32 sixteen-byte columns per item; every thread keeps 4 independent outputs in flight and issues all
    // tp x 4 slab loads before the adds (the loads are L2 / HBM latency bound otherwise)
    constexpr int ILP = 4;
    for (int u0 = threadIdx.x; u0 < 128 * 32; u0 += NUM_THREADS * ILP) {
      uint4 v[ILP][MAX_TP];
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        const int u = u0 + i * NUM_THREADS;
        const int r = u >> 5, col = nc * 256 + (u & 31) * 8;
        const bool ok = u < 128 * 32 && col < rs.N;
#pragma unroll
        for (int sl = 0; sl < MAX_TP; ++sl)
          if (ok && sl < rs.tp)
            v[i][sl] = __ldcg(reinterpret_cast<const uint4*>(rs.stage_local + ((size_t)sl * rs.rows_per_rank + rb * 128 + r) * rs.N + col));
      }
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        const int u = u0 + i * NUM_THREADS;
        const int r = u >> 5, col = nc * 256 + (u & 31) * 8;
        if (u >= 128 * 32 || col >= rs.N) continue;
        float acc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.f;
#pragma unroll
        for (int sl = 0; sl < MAX_TP; ++sl) {
          if (sl < rs.tp) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[i][sl]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(h[t]);
              acc[2 * t] += f.x;
              acc[2 * t + 1] += f.y;
            }
          }
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(rs.out + ((size_t)rb * 128 + r) * rs.N + col) = o;
      }
    }
    __syncthreads();
  }
  dbg_stamp(3);
}

// Device-resident epoch: incremented by a one-thread kernel in stream order right before every fused op, so that a
// captured CUDA graph advances the cross-GPU protocol on every replay (a host-passed epoch would be baked in).
__global__ void bump_epoch_kernel(uint32_t* epoch) { *epoch = *epoch + 1u; }

// All-peer barrier at the epoch held in `epoch_ptr`: every rank has finished everything it enqueued before this
// kernel once the kernel completes on all ranks.  Placed at CUDA-graph boundaries, where the double-buffer parity of
// the staged shards restarts.
__global__ void tp_barrier_kernel(const RsArgs a) {
  const uint32_t epoch = ld_acquire_gpu(a.epoch_ptr);
  const int t = threadIdx.x;
  if (t < a.tp) {
    st_release_sys(a.hs_peer[t] + a.rank, epoch);
    spin_until_ge_sys(a.hs_local + t, epoch, 100, 0x500u | (uint32_t)a.rank);
  }
}

template <class K>
static void set_smem(K kern) {
  // keyed by the function address: kernels of different template arguments share one signature type
  static std::set<const void*> done;
  const void* key = reinterpret_cast<const void*>(kern);
  if (done.insert(key).second)
    C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

static void fill_ptrs(const std::vector<int64_t>& v, int tp, void** dst) {
  TORCH_CHECK((int)v.size() == tp, "expected one pointer per TP rank");
  for (int i = 0; i < MAX_TP; ++i) dst[i] = i < tp ? reinterpret_cast<void*>(v[i]) : nullptr;
}

}  // namespace fused

// mode 0: out[M_total,N] = xg[M_total,K] * op(b)            (b: [N,K] if b_kmajor else [K,N])
// mode 1: out[N',K'] (+)= other[M_total,N']^T * xg[M_total,K']   (wgrad; `b` unused)
// xg is filled inside the kernel from the peers' shards `peer_x` ([rows_per_rank, K] each).
void ag_gemm(at::Tensor xg, const at::Tensor& x_local, const at::Tensor& b, at::Tensor out, const at::Tensor& other,
             std::vector<int64_t> xg_peer, std::vector<int64_t> flags_peer, int64_t hs_local_ptr,
             std::vector<int64_t> hs_peer, int64_t epoch, int64_t rank, int64_t tp, int64_t mode, bool b_kmajor,
             bool accumulate, int64_t num_comm_ctas, std::vector<int64_t> x_peer, bool push, int64_t epoch_ptr) {
  using namespace fused;
  TORCH_CHECK(xg.is_cuda() && xg.is_contiguous() && xg.scalar_type() == at::kBFloat16 && xg.dim() == 2);
  TORCH_CHECK(tp >= 2 && tp <= MAX_TP);
  const int M_total = xg.size(0), Kx = xg.size(1);
  const int rows_per_rank = M_total / tp;
  TORCH_CHECK(rows_per_rank * tp == M_total && rows_per_rank % 128 == 0, "ag_gemm: rows per rank must be a multiple of 128");
  c10::cuda::CUDAGuard guard(xg.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  CommArgs comm;
  TORCH_CHECK(x_local.is_cuda() && x_local.is_contiguous() && x_local.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x_local.size(0) * tp == xg.size(0) && x_local.size(1) == xg.size(1), "shard / gathered shape mismatch");
  comm.x_local = reinterpret_cast<const uint8_t*>(x_local.data_ptr());
  fill_ptrs(xg_peer, tp, (void**)comm.xg_peer);
  fill_ptrs(flags_peer, tp, (void**)comm.flags_peer);
  fill_ptrs(x_peer, tp, (void**)comm.x_peer);
  comm.push = push ? 1 : 0;
  fill_ptrs(hs_peer, tp, (void**)comm.hs_peer);
  comm.hs_local = reinterpret_cast<uint32_t*>(hs_local_ptr);
  comm.flags = comm.flags_peer[rank];
  TORCH_CHECK(reinterpret_cast<uint8_t*>(xg.data_ptr()) == comm.xg_peer[rank], "xg must be my symmetric gathered buffer");
  comm.epoch = (uint32_t)epoch;
  comm.epoch_ptr = reinterpret_cast<const uint32_t*>(epoch_ptr);
  if (epoch_ptr != 0) bump_epoch_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<uint32_t*>(epoch_ptr));
  comm.rank = rank; comm.tp = tp; comm.rows_per_rank = rows_per_rank; comm.row_bytes = Kx * 2;
  num_comm_ctas = std::max<int64_t>(1, std::min<int64_t>(num_comm_ctas, (push ? 1 : (tp - 1)) * (rows_per_rank / 128)));
  comm.num_comm_ctas = num_comm_ctas;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  TORCH_CHECK(num_comm_ctas >= 1 && num_comm_ctas < sms);
  CUtensorMap ta, tb, tc;
  Params p;
  if (mode == 0) {
    const int N = b_kmajor ? b.size(0) : b.size(1);
    TORCH_CHECK((b_kmajor ? b.size(1) : b.size(0)) == Kx && b.is_contiguous());
    TORCH_CHECK(out.size(0) == M_total && out.size(1) == N && out.scalar_type() == at::kBFloat16 && out.is_contiguous());
    make_operand_maps(xg.data_ptr(), b.data_ptr(), M_total, N, Kx, true, b_kmajor, ta, tb);
    tc = make_out_map(out.data_ptr(), M_total, N, false);
    p = make_params(M_total, N, Kx);
    p.group_m = 1;                                         // all n-tiles of one row-block before the next
    p.m_first = rank * (rows_per_rank / 128);
    auto run = [&](auto kern) {
      set_smem(kern);
      kern<<<sms, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, comm, 0);
    };
    if (b_kmajor) run(ag_gemm_kernel<false, false, OUT_BF16_STORE>);
    else run(ag_gemm_kernel<false, true, OUT_BF16_STORE>);
  } else {
    // wgrad: A = other (dY, stored [M_total, N'] => MN-major), B = xg (stored [M_total, K'] => MN-major)
    const int Np = other.size(1);
    TORCH_CHECK(other.size(0) == M_total && other.is_contiguous() && other.scalar_type() == at::kBFloat16);
    TORCH_CHECK(out.size(0) == Np && out.size(1) == Kx && out.is_contiguous());
    const bool f32 = out.scalar_type() == at::kFloat;
    make_operand_maps(other.data_ptr(), xg.data_ptr(), Np, Kx, M_total, false, false, ta, tb);
    tc = make_out_map(out.data_ptr(), Np, Kx, f32);
    p = make_params(Np, Kx, M_total);
    p.k_first = rank * (rows_per_rank / 64);
    auto run = [&](auto kern) {
      set_smem(kern);
      kern<<<sms, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, comm, 1);
    };
    if (f32) { if (accumulate) run(ag_gemm_kernel<true, true, OUT_F32_ADD>); else run(ag_gemm_kernel<true, true, OUT_F32_STORE>); }
    else { if (accumulate) run(ag_gemm_kernel<true, true, OUT_BF16_ADD>); else run(ag_gemm_kernel<true, true, OUT_BF16_STORE>); }
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void ag_only(at::Tensor xg, const at::Tensor& x_local, std::vector<int64_t> xg_peer, std::vector<int64_t> flags_peer,
             int64_t hs_local_ptr, std::vector<int64_t> hs_peer, int64_t epoch, int64_t rank, int64_t tp,
             int64_t num_comm_ctas, std::vector<int64_t> x_peer, bool push, int64_t epoch_ptr) {
  using namespace fused;
  const int M_total = xg.size(0), Kx = xg.size(1);
  const int rows_per_rank = M_total / tp;
  TORCH_CHECK(rows_per_rank % 128 == 0);
  c10::cuda::CUDAGuard guard(xg.device());
  CommArgs comm;
  TORCH_CHECK(x_local.is_cuda() && x_local.is_contiguous() && x_local.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x_local.size(0) * tp == xg.size(0) && x_local.size(1) == xg.size(1), "shard / gathered shape mismatch");
  comm.x_local = reinterpret_cast<const uint8_t*>(x_local.data_ptr());
  fill_ptrs(xg_peer, tp, (void**)comm.xg_peer);
  fill_ptrs(flags_peer, tp, (void**)comm.flags_peer);
  fill_ptrs(x_peer, tp, (void**)comm.x_peer);
  comm.push = push ? 1 : 0;
  fill_ptrs(hs_peer, tp, (void**)comm.hs_peer);
  comm.hs_local = reinterpret_cast<uint32_t*>(hs_local_ptr);
  comm.flags = comm.flags_peer[rank];
  TORCH_CHECK(reinterpret_cast<uint8_t*>(xg.data_ptr()) == comm.xg_peer[rank], "xg must be my symmetric gathered buffer");
  comm.epoch = (uint32_t)epoch;
  comm.epoch_ptr = reinterpret_cast<const uint32_t*>(epoch_ptr);
  if (epoch_ptr != 0)
    bump_epoch_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>(reinterpret_cast<uint32_t*>(epoch_ptr));
  comm.rank = rank; comm.tp = tp; comm.rows_per_rank = rows_per_rank; comm.row_bytes = Kx * 2;
  num_comm_ctas = std::max<int64_t>(1, std::min<int64_t>(num_comm_ctas, (push ? 1 : (tp - 1)) * (rows_per_rank / 128)));
  comm.num_comm_ctas = num_comm_ctas;
  set_smem(ag_only_kernel);
  ag_only_kernel<<<num_comm_ctas, NUM_THREADS, SMEM_BYTES, at::cuda::getCurrentCUDAStream()>>>(comm);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// out[rows_per_rank, N] = reduce_scatter_rows( a[M_total, K] * op(b) )
void gemm_rs(const at::Tensor& a, const at::Tensor& b, at::Tensor out, std::vector<int64_t> stage_peer,
             std::vector<int64_t> counter_peer, int64_t hs_local_ptr, std::vector<int64_t> hs_peer, int64_t epoch,
             int64_t rank, int64_t tp, bool b_kmajor, int64_t epoch_ptr) {
  using namespace fused;
  TORCH_CHECK(a.is_cuda() && a.is_contiguous() && b.is_contiguous() && out.is_contiguous());
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && out.scalar_type() == at::kBFloat16);
  const int M_total = a.size(0), K = a.size(1);
  const int N = b_kmajor ? b.size(0) : b.size(1);
  TORCH_CHECK((b_kmajor ? b.size(1) : b.size(0)) == K);
  const int rows_per_rank = M_total / tp;
  TORCH_CHECK(rows_per_rank * tp == M_total && rows_per_rank % 128 == 0 && N % 8 == 0);
  TORCH_CHECK(out.size(0) == rows_per_rank && out.size(1) == N);
  c10::cuda::CUDAGuard guard(a.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  RsArgs rs;
  fill_ptrs(stage_peer, tp, (void**)rs.stage_peer);
  fill_ptrs(counter_peer, tp, (void**)rs.counter_peer);
  fill_ptrs(hs_peer, tp, (void**)rs.hs_peer);
  rs.hs_local = reinterpret_cast<uint32_t*>(hs_local_ptr);
  rs.stage_local = rs.stage_peer[rank];
  rs.counter_local = rs.counter_peer[rank];
  rs.out = reinterpret_cast<__nv_bfloat16*>(out.data_ptr());
  rs.epoch = (uint32_t)epoch;
  rs.epoch_ptr = reinterpret_cast<const uint32_t*>(epoch_ptr);
  rs.rank = rank; rs.tp = tp; rs.rows_per_rank = rows_per_rank; rs.N = N;
  CUtensorMap ta, tb;
  make_operand_maps(a.data_ptr(), b.data_ptr(), M_total, N, K, true, b_kmajor, ta, tb);
  RsMaps maps;
  for (int d = 0; d < MAX_TP; ++d)
    maps.m[d] = make_out_map(rs.stage_peer[d < tp ? d : 0], M_total, N, false);
  Params p = make_params(M_total, N, K);
  rs.num_n_blocks = p.num_n_blocks;
  p.group_m = 1;
  // start with the rows owned by rank+1 and finish with my own rows: what I need from the others arrives
  // while I am still computing, and what the others need from me is produced in ring order
  p.m_first = ((rank + 1) % tp) * (rows_per_rank / 128);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  if (epoch_ptr != 0) bump_epoch_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<uint32_t*>(epoch_ptr));
  C10_CUDA_CHECK(cudaMemsetAsync(rs.counter_local, 0, sizeof(uint32_t) * (rows_per_rank / 128), stream));
  auto run = [&](auto kern) {
    set_smem(kern);
    kern<<<sms, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, maps, p, rs);
  };
  if (b_kmajor) run(gemm_rs_kernel<false>); else run(gemm_rs_kernel<true>);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// all-peer barrier on the current stream (advances the device epoch); see tp_barrier_kernel
void tp_barrier(int64_t hs_local_ptr, std::vector<int64_t> hs_peer, int64_t rank, int64_t tp, int64_t epoch_ptr) {
  using namespace fused;
  TORCH_CHECK(epoch_ptr != 0 && tp >= 2 && tp <= MAX_TP);
  auto stream = at::cuda::getCurrentCUDAStream();
  RsArgs a{};
  fill_ptrs(hs_peer, tp, (void**)a.hs_peer);
  a.hs_local = reinterpret_cast<uint32_t*>(hs_local_ptr);
  a.epoch_ptr = reinterpret_cast<const uint32_t*>(epoch_ptr);
  a.rank = rank; a.tp = tp;
  bump_epoch_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<uint32_t*>(epoch_ptr));
  tp_barrier_kernel<<<1, 32, 0, stream>>>(a);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// debug: point the in-kernel timeline capture at `buf` ([num_ctas, 4] int64) or disable it with an empty tensor
void fused_tp_set_debug(at::Tensor buf) {
  unsigned long long* p = buf.numel() ? reinterpret_cast<unsigned long long*>(buf.data_ptr()) : nullptr;
  C10_CUDA_CHECK(cudaMemcpyToSymbol(fused::g_dbg, &p, sizeof(p)));
}
