// Host launcher + kernel instantiations for the sm_100a tcgen05 GEMM (see gemm_sm100.cuh).
#include <cstdlib>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "gemm_sm100.cuh"
#include "tma_utils.h"

namespace gemm100 {

template <bool A_MN, bool B_MN, int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  gemm_body<A_MN, B_MN, OUT, NoGate>(&tma_a, &tma_b, &tma_c, p, smem_raw, NoGate{}, blockIdx.x, gridDim.x);
}

template <bool A_MN, bool B_MN, int OUT>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Params& p, int grid,
                   cudaStream_t stream) {
  auto kern = gemm_kernel<A_MN, B_MN, OUT>;
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

Params make_params(int M, int N, int K);
void make_operand_maps(const void* a, const void* b, int M, int N, int K, bool a_kmajor, bool b_kmajor,
                       CUtensorMap& ta, CUtensorMap& tb);
CUtensorMap make_out_map(void* out, int M, int N, bool f32);

Params make_params(int M, int N, int K) {
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  p.num_tiles = p.num_m_blocks * p.num_n_blocks;
  p.group_m = 8;
  p.m_first = 0;
  p.k_first = 0;
  return p;
}

void make_operand_maps(const void* a, const void* b, int M, int N, int K, bool a_kmajor, bool b_kmajor,
                       CUtensorMap& ta, CUtensorMap& tb) {
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  // K-major operand stored [rows=MN, cols=K]: box {64 k, BLOCK_MN rows}.  MN-major stored [rows=K, cols=MN]:
  // box {64 mn, 64 k}.
  ta = a_kmajor ? tma::make_2d(a, BF, 2, M, K, (uint64_t)K * 2, BLOCK_M, 64)
                : tma::make_2d(a, BF, 2, K, M, (uint64_t)M * 2, BLOCK_K, 64);
  tb = b_kmajor ? tma::make_2d(b, BF, 2, N, K, (uint64_t)K * 2, BLOCK_N, 64)
                : tma::make_2d(b, BF, 2, K, N, (uint64_t)N * 2, BLOCK_K, 64);
}

CUtensorMap make_out_map(void* out, int M, int N, bool f32) {
  return f32 ? tma::make_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, (uint64_t)N * 4, BLOCK_M, 32)
             : tma::make_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, (uint64_t)N * 2, BLOCK_M, 64);
}

}  // namespace gemm100

// Upper bound on the CTAs of the persistent GEMM grid (0 = one per SM).  The two-stream micro-batch mode lowers it so
// that a collective of the other stream finds free SMs WHILE a GEMM runs instead of queueing behind its 148 resident CTAs.
static int g_gemm_cta_limit = 0;
void gemm_set_cta_limit(int64_t n) { g_gemm_cta_limit = (int)n; }
int64_t gemm_get_cta_limit() { return g_gemm_cta_limit; }

// SM-pair kernel (gemm2cta_sm100.cu) for every shape with more than one 128-row / 128-column block: measured +9..18 % on the
// K-major (forward) shapes of Llama-3-8B, equal on the MN-major ones.  gemm_set_2cta(false) forces the one-CTA kernel (A/B runs).
void gemm_2cta(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, bool a_kmajor, bool b_kmajor, bool accumulate);
static bool g_gemm_2cta = [] { const char* e = std::getenv("NXDT_GEMM_2CTA"); return !(e && e[0] == '0'); }();
void gemm_set_2cta(bool on) { g_gemm_2cta = on; }

// out (+)= op(a) op(b).  a_kmajor: a is [M,K] else [K,M];  b_kmajor: b is [N,K] else [K,N]; out is [M,N].
void gemm(const at::Tensor& a, const at::Tensor& b, at::Tensor& out, bool a_kmajor, bool b_kmajor, bool accumulate) {
  using namespace gemm100;
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && out.is_cuda(), "gemm: CUDA tensors required");
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16, "gemm: bf16 inputs required");
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && out.is_contiguous(), "gemm: contiguous tensors required");
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && out.dim() == 2);
  const int M = a_kmajor ? a.size(0) : a.size(1);
  const int K = a_kmajor ? a.size(1) : a.size(0);
  const int N = b_kmajor ? b.size(0) : b.size(1);
  TORCH_CHECK((b_kmajor ? b.size(1) : b.size(0)) == K, "gemm: K mismatch");
  TORCH_CHECK(out.size(0) == M && out.size(1) == N, "gemm: output shape mismatch");
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0 && M % 8 == 0, "gemm: dims must be multiples of 8 (16-byte TMA pitch)");
  const bool f32 = out.scalar_type() == at::kFloat;
  TORCH_CHECK(f32 || out.scalar_type() == at::kBFloat16, "gemm: out must be bf16 or fp32");
  if (g_gemm_2cta && M > BLOCK_M && N > 128) return gemm_2cta(a, b, out, a_kmajor, b_kmajor, accumulate);
  c10::cuda::CUDAGuard guard(a.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  CUtensorMap ta, tb;
  make_operand_maps(a.data_ptr(), b.data_ptr(), M, N, K, a_kmajor, b_kmajor, ta, tb);
  CUtensorMap tc = make_out_map(out.data_ptr(), M, N, f32);
  Params p = make_params(M, N, K);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = std::min(p.num_tiles, g_gemm_cta_limit > 0 ? std::min(g_gemm_cta_limit, sms) : sms);
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
