// Grouped (per-expert) bf16 GEMMs for dropless MoE on sm_100a, built on the persistent tcgen05 GEMM body.
//
// Tokens are permuted so that every expert owns a contiguous run of rows whose length is padded to a multiple of
// 128 (one BLOCK_M); padding rows are zero.  With that layout
//   forward / dgrad : Y[rows of e] = X[rows of e] * W_e      one launch; the m-block -> expert table lives on the
//                     device (no host sync on the routing result); W is a 3-D tensor map [E, ., .] so out-of-range
//                     rows / columns zero-fill per expert
//   wgrad           : dW_e (+)= X[rows of e]^T * dY[rows of e]   one launch; the reduction range of an output tile is
//                     the row range of its expert, read from a device offsets table
// This is the role the reference delegates to the NKI "blockwise" dropless-MoE kernel.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "gemm_sm100.cuh"
#include "tma_utils.h"

namespace gemm100 {

struct GroupedRowsGate : NoGate {
  const int* __restrict__ block_expert;   // [num_m_blocks]  expert of each 128-row block, -1 = unused block
  __device__ __forceinline__ bool tile_info(int m_blk, int& /*a_m0*/, int& b_group, int& /*kb_begin*/,
                                            int& /*kb_count*/) const {
    const int e = __ldg(block_expert + m_blk);
    b_group = e;
    return e >= 0;
  }
};
template <>
struct gate_traits<GroupedRowsGate> {
  static constexpr bool grouped = true;
  static constexpr bool b_3d = true;
  static constexpr bool fp8 = false;
};

struct GroupedWgradGate : NoGate {
  const int* __restrict__ offsets;        // [E + 1] padded row offsets (multiples of 128)
  int blocks_per_expert;                  // output row blocks per expert (= rows_per_expert / 128)
  __device__ __forceinline__ bool tile_info(int m_blk, int& a_m0, int& b_group, int& kb_begin, int& kb_count) const {
    const int e = m_blk / blocks_per_expert;
    a_m0 = (m_blk - e * blocks_per_expert) * BLOCK_M;
    b_group = 0;
    const int lo = __ldg(offsets + e), hi = __ldg(offsets + e + 1);
    kb_begin = lo / BLOCK_K;
    kb_count = (hi - lo) / BLOCK_K;
    return kb_count > 0;
  }
};
template <>
struct gate_traits<GroupedWgradGate> {
  static constexpr bool grouped = true;
  static constexpr bool b_3d = false;
  static constexpr bool fp8 = false;
};

template <bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
grouped_rows_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, const Params p, const int* block_expert) {
  extern __shared__ uint8_t smem_raw[];
  GroupedRowsGate g;
  g.block_expert = block_expert;
  gemm_body<false, B_MN, OUT_BF16_STORE, GroupedRowsGate>(&tma_a, &tma_b, &tma_c, p, smem_raw, g, blockIdx.x, gridDim.x);
}

template <int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
grouped_wgrad_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_c, const Params p, const int* offsets,
                     int blocks_per_expert) {
  extern __shared__ uint8_t smem_raw[];
  GroupedWgradGate g;
  g.offsets = offsets;
  g.blocks_per_expert = blocks_per_expert;
  gemm_body<true, true, OUT, GroupedWgradGate>(&tma_a, &tma_b, &tma_c, p, smem_raw, g, blockIdx.x, gridDim.x);
}

Params make_params(int M, int N, int K);
CUtensorMap make_out_map(void* out, int M, int N, bool f32);

template <class Kern>
static void set_smem(Kern kern) {
  C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

}  // namespace gemm100

// out[T_pad, N] = a[T_pad, K] x W[expert of the row block];  b is [E, N, K] (b_kmajor) or [E, K, N].
// block_expert: int32 [T_pad / 128] on the device, -1 for blocks that hold no tokens (left untouched in `out`).
void grouped_gemm(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, const at::Tensor& block_expert,
                  bool b_kmajor) {
  using namespace gemm100;
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && out.is_cuda() && block_expert.is_cuda(), "grouped_gemm: CUDA tensors");
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 &&
              out.scalar_type() == at::kBFloat16, "grouped_gemm: bf16 tensors required");
  TORCH_CHECK(block_expert.scalar_type() == at::kInt, "grouped_gemm: block_expert must be int32");
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && out.is_contiguous() && block_expert.is_contiguous());
  TORCH_CHECK(a.dim() == 2 && b.dim() == 3 && out.dim() == 2);
  const int64_t T = a.size(0), K = a.size(1), E = b.size(0);
  const int64_t N = b_kmajor ? b.size(1) : b.size(2);
  TORCH_CHECK((b_kmajor ? b.size(2) : b.size(1)) == K, "grouped_gemm: K mismatch");
  TORCH_CHECK(out.size(0) == T && out.size(1) == N, "grouped_gemm: output shape mismatch");
  TORCH_CHECK(T % BLOCK_M == 0, "grouped_gemm: permuted rows must be padded to a multiple of 128");
  TORCH_CHECK(block_expert.numel() == T / BLOCK_M, "grouped_gemm: block_expert needs one entry per 128 rows");
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0, "grouped_gemm: K and N must be multiples of 8");
  if (T == 0) return;
  c10::cuda::CUDAGuard guard(a.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ta = tma::make_2d(a.data_ptr(), BF, 2, T, K, (uint64_t)K * 2, BLOCK_M, 64);
  CUtensorMap tb = b_kmajor
      ? tma::make_3d(b.data_ptr(), BF, K, N, E, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, BLOCK_N, 1)
      : tma::make_3d(b.data_ptr(), BF, N, K, E, (uint64_t)N * 2, (uint64_t)N * K * 2, 64, BLOCK_K, 1);
  CUtensorMap tc = make_out_map(out.data_ptr(), (int)T, (int)N, false);
  Params p = make_params((int)T, (int)N, (int)K);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = std::min(p.num_tiles, sms);
  static bool configured = false;
  if (!configured) {
    set_smem(grouped_rows_kernel<false>);
    set_smem(grouped_rows_kernel<true>);
    configured = true;
  }
  if (b_kmajor)
    grouped_rows_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, block_expert.data_ptr<int>());
  else
    grouped_rows_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, block_expert.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// out[E, M, N] (+)= x[rows of e, M]^T * dy[rows of e, N];  offsets: int32 [E + 1] padded row offsets on the device.
// `out` is fp32 (store or reduce-add, e.g. straight into main_grad) or bf16.  Experts without tokens are skipped,
// so with accumulate = false the caller must pass a zero-initialised `out`.
void grouped_wgrad(const at::Tensor& x, const at::Tensor& dy, at::Tensor& out, const at::Tensor& offsets,
                   bool accumulate) {
  using namespace gemm100;
  TORCH_CHECK(x.is_cuda() && dy.is_cuda() && out.is_cuda() && offsets.is_cuda(), "grouped_wgrad: CUDA tensors");
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && dy.scalar_type() == at::kBFloat16, "grouped_wgrad: bf16 inputs");
  TORCH_CHECK(offsets.scalar_type() == at::kInt, "grouped_wgrad: offsets must be int32");
  TORCH_CHECK(x.is_contiguous() && dy.is_contiguous() && out.is_contiguous() && offsets.is_contiguous());
  TORCH_CHECK(x.dim() == 2 && dy.dim() == 2 && out.dim() == 3 && x.size(0) == dy.size(0));
  const int64_t T = x.size(0), M = x.size(1), N = dy.size(1), E = out.size(0);
  TORCH_CHECK(out.size(1) == M && out.size(2) == N, "grouped_wgrad: output shape mismatch");
  TORCH_CHECK(offsets.numel() == E + 1, "grouped_wgrad: offsets must have E + 1 entries");
  TORCH_CHECK(M % BLOCK_M == 0, "grouped_wgrad: rows per expert must be a multiple of 128");
  TORCH_CHECK(N % 8 == 0 && T % BLOCK_M == 0);
  const bool f32 = out.scalar_type() == at::kFloat;
  TORCH_CHECK(f32 || out.scalar_type() == at::kBFloat16, "grouped_wgrad: out must be fp32 or bf16");
  if (T == 0) return;
  c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  // both operands are MN-major: stored [T (reduction), M] and [T, N]; boxes {64 mn, 64 k}
  CUtensorMap ta = tma::make_2d(x.data_ptr(), BF, 2, T, M, (uint64_t)M * 2, BLOCK_K, 64);
  CUtensorMap tb = tma::make_2d(dy.data_ptr(), BF, 2, T, N, (uint64_t)N * 2, BLOCK_K, 64);
  CUtensorMap tc = make_out_map(out.data_ptr(), (int)(E * M), (int)N, f32);
  Params p = make_params((int)(E * M), (int)N, (int)T);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = std::min(p.num_tiles, sms);
  static bool configured = false;
  if (!configured) {
    set_smem(grouped_wgrad_kernel<OUT_F32_ADD>);
    set_smem(grouped_wgrad_kernel<OUT_F32_STORE>);
    set_smem(grouped_wgrad_kernel<OUT_BF16_STORE>);
    set_smem(grouped_wgrad_kernel<OUT_BF16_ADD>);
    configured = true;
  }
  const int bpe = (int)(M / BLOCK_M);
  const int* off = offsets.data_ptr<int>();
  if (f32 && accumulate)
    grouped_wgrad_kernel<OUT_F32_ADD><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, off, bpe);
  else if (f32)
    grouped_wgrad_kernel<OUT_F32_STORE><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, off, bpe);
  else if (accumulate)
    grouped_wgrad_kernel<OUT_BF16_ADD><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, off, bpe);
  else
    grouped_wgrad_kernel<OUT_BF16_STORE><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, off, bpe);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
