// fp8 (e4m3) GEMM for sm_100a with row / column block scales -- PREVIEW (compiled in round 1, not yet validated on
// hardware; `tests/gpu_checks/check_fp8.py` is its numerics / timing driver).
//
//   out[M, N] (+)= (A_q[M, K] * B_q[N, K]^T) * a_scale[M] (outer) b_scale[N]
//
// Both operands are K-major e4m3 with ONE scale per row, i.e. per 1 x K block of the reduction dimension -- the
// quantisation kernels below emit a row-scaled copy and, for the tensors that are reduced along their other dimension
// in the backward GEMMs, a transposed copy scaled per original column, so forward, dgrad and wgrad all run "nt":
//     fwd    Y  = Xq  Wq^T          Xq  [M,K] row scales (per token),      Wq  [N,K] row scales (per output feature)
//     dgrad  dX = dYq (W^T)q^T      dYq [M,N] row scales (per token),      (W^T)q [K,N] row scales (per input feature)
//     wgrad  dW = (dY^T)q (X^T)q^T  (dY^T)q [N,M] per output feature,      (X^T)q [K,M] per input feature
// The kernel is the persistent tcgen05 body of gemm_sm100.cuh with `gate_traits::fp8`: 128 fp8 elements per 128-byte
// swizzle row, `tcgen05.mma.kind::f8f6f4` (K = 32 per instruction, twice the bf16 rate), dequantisation in the
// epilogue right after the TMEM load.  The hardware micro-scaling mode (`kind::mxf8f6f4.block_scale`, 1 x 32 E8M0
// scales staged in TMEM) is the planned follow-up; this recipe needs no scale traffic inside the main loop.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_fp8.h>

#include "gemm_sm100.cuh"
#include "tma_utils.h"

namespace gemm100 {

struct Fp8Gate : NoGate {
  const float* __restrict__ a_scale;   // [M]
  const float* __restrict__ b_scale;   // [N]
  int M, N;
  __device__ __forceinline__ float row_scale(int m) const { return m < M ? __ldg(a_scale + m) : 0.f; }
  __device__ __forceinline__ float col_scale(int n) const { return n < N ? __ldg(b_scale + n) : 0.f; }
};
template <>
struct gate_traits<Fp8Gate> {
  static constexpr bool grouped = false;
  static constexpr bool b_3d = false;
  static constexpr bool fp8 = true;
};

template <int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ CUtensorMap tma_c, const Params p, const float* a_scale, const float* b_scale) {
  extern __shared__ uint8_t smem_raw[];
  Fp8Gate g;
  g.a_scale = a_scale; g.b_scale = b_scale; g.M = p.M; g.N = p.N;
  gemm_body<false, false, OUT, Fp8Gate>(&tma_a, &tma_b, &tma_c, p, smem_raw, g, blockIdx.x, gridDim.x);
}

Params make_params(int M, int N, int K);
CUtensorMap make_out_map(void* out, int M, int N, bool f32);

// ---------------------------------------------------------------------------------------------------
// quantisation: bf16 [R, C] -> e4m3 with one fp32 scale per row (scale = amax / 448, x_q = x / scale)
// ---------------------------------------------------------------------------------------------------
constexpr float E4M3_MAX = 448.f;

__global__ void __launch_bounds__(256) quant_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_fp8_e4m3* __restrict__ q,
                                                         float* __restrict__ scale, int R, int C) {
  __shared__ float sm[8];
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const __nv_bfloat16* xr = x + (size_t)r * C;
    float amax = 0.f;
    for (int c = threadIdx.x * 8; c < C; c += 256 * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = sm[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) amax = fmaxf(amax, sm[i]);
    __syncthreads();
    const float sc = amax > 0.f ? amax / E4M3_MAX : 1.f;
    const float inv = 1.f / sc;
    if (threadIdx.x == 0) scale[r] = sc;
    __nv_fp8_e4m3* qr = q + (size_t)r * C;
    for (int c = threadIdx.x * 8; c < C; c += 256 * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
      alignas(8) __nv_fp8_e4m3 o[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        o[2 * i] = __nv_fp8_e4m3(f.x * inv);
        o[2 * i + 1] = __nv_fp8_e4m3(f.y * inv);
      }
      *reinterpret_cast<uint2*>(qr + c) = *reinterpret_cast<const uint2*>(o);
    }
  }
}

// column statistics of a bf16 [R, C] matrix: amax[c] = max_r |x[r, c]|  (atomicMax on the float bit pattern, x >= 0)
__global__ void __launch_bounds__(256) col_amax_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ amax,
                                                       int R, int C) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (c >= C) return;
  float a0 = 0.f, a1 = 0.f;
  for (int r = blockIdx.y; r < R; r += gridDim.y) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (size_t)r * C + c));
    a0 = fmaxf(a0, fabsf(f.x));
    a1 = fmaxf(a1, fabsf(f.y));
  }
  atomicMax(reinterpret_cast<int*>(amax + c), __float_as_int(a0));
  atomicMax(reinterpret_cast<int*>(amax + c + 1), __float_as_int(a1));
}

// transposed quantisation: x [R, C] bf16 -> q [C, R] e4m3 with scale[c] = amax[c] / 448   (32 x 32 smem tiles)
__global__ void __launch_bounds__(256) quant_cols_transpose_kernel(const __nv_bfloat16* __restrict__ x,
                                                                   const float* __restrict__ amax,
                                                                   __nv_fp8_e4m3* __restrict__ q,
                                                                   float* __restrict__ scale, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? __bfloat162float(x[(size_t)r * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < R) {
      const float a = amax[c];
      const float sc = a > 0.f ? a / E4M3_MAX : 1.f;
      q[(size_t)c * R + r] = __nv_fp8_e4m3(tile[tx][i] / sc);
      if (blockIdx.y == 0 && tx == 0) scale[c] = sc;
    }
  }
}

}  // namespace gemm100

// x [R, C] bf16 -> (q [R, C] e4m3 as uint8, scale [R] fp32)
std::vector<at::Tensor> fp8_quantize_rows(const at::Tensor& x) {
  using namespace gemm100;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous() && x.size(1) % 8 == 0);
  const int R = x.size(0), C = x.size(1);
  c10::cuda::CUDAGuard guard(x.device());
  auto q = at::empty({R, C}, x.options().dtype(at::kByte));
  auto scale = at::empty({R}, x.options().dtype(at::kFloat));
  if (R == 0) return {q, scale};
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  quant_rows_kernel<<<std::min(R, sms * 8), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      (const __nv_bfloat16*)x.data_ptr(), (__nv_fp8_e4m3*)q.data_ptr(), scale.data_ptr<float>(), R, C);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {q, scale};
}

// x [R, C] bf16 -> (q [C, R] e4m3 as uint8 = quantised x^T, scale [C] fp32 per original column)
std::vector<at::Tensor> fp8_quantize_cols_transposed(const at::Tensor& x) {
  using namespace gemm100;
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous() && x.size(1) % 2 == 0);
  const int R = x.size(0), C = x.size(1);
  c10::cuda::CUDAGuard guard(x.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  auto q = at::empty({C, R}, x.options().dtype(at::kByte));
  auto scale = at::empty({C}, x.options().dtype(at::kFloat));
  auto amax = at::zeros({C}, x.options().dtype(at::kFloat));
  if (R == 0 || C == 0) return {q, scale};
  dim3 g1((C / 2 + 255) / 256, std::min(R, 64));
  col_amax_kernel<<<g1, 256, 0, stream>>>((const __nv_bfloat16*)x.data_ptr(), amax.data_ptr<float>(), R, C);
  dim3 g2((C + 31) / 32, (R + 31) / 32);
  quant_cols_transpose_kernel<<<g2, 256, 0, stream>>>((const __nv_bfloat16*)x.data_ptr(), amax.data_ptr<float>(),
                                                      (__nv_fp8_e4m3*)q.data_ptr(), scale.data_ptr<float>(), R, C);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {q, scale};
}

// out[M, N] (+)= (a_q[M, K] b_q[N, K]^T) * a_scale[M] x b_scale[N];  a_q / b_q are e4m3 bytes (uint8 tensors)
void gemm_fp8(const at::Tensor& a_q, const at::Tensor& a_scale, const at::Tensor& b_q, const at::Tensor& b_scale,
              at::Tensor& out, bool accumulate) {
  using namespace gemm100;
  TORCH_CHECK(a_q.is_cuda() && b_q.is_cuda() && out.is_cuda());
  TORCH_CHECK(a_q.scalar_type() == at::kByte && b_q.scalar_type() == at::kByte, "gemm_fp8: e4m3 bytes (uint8) expected");
  TORCH_CHECK(a_q.is_contiguous() && b_q.is_contiguous() && out.is_contiguous() && a_q.dim() == 2 && b_q.dim() == 2);
  const int M = a_q.size(0), K = a_q.size(1), N = b_q.size(0);
  TORCH_CHECK(b_q.size(1) == K && out.size(0) == M && out.size(1) == N);
  TORCH_CHECK(K % 16 == 0 && N % 8 == 0, "gemm_fp8: K must be a multiple of 16 and N of 8 (16-byte TMA pitch)");
  TORCH_CHECK(a_scale.scalar_type() == at::kFloat && b_scale.scalar_type() == at::kFloat && a_scale.numel() == M &&
              b_scale.numel() == N && a_scale.is_contiguous() && b_scale.is_contiguous());
  const bool f32 = out.scalar_type() == at::kFloat;
  TORCH_CHECK(f32 || out.scalar_type() == at::kBFloat16);
  if (M == 0 || N == 0) return;
  c10::cuda::CUDAGuard guard(a_q.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const auto U8 = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUtensorMap ta = tma::make_2d(a_q.data_ptr(), U8, 1, M, K, (uint64_t)K, BLOCK_M, 128);
  CUtensorMap tb = tma::make_2d(b_q.data_ptr(), U8, 1, N, K, (uint64_t)K, BLOCK_N, 128);
  CUtensorMap tc = make_out_map(out.data_ptr(), M, N, f32);
  Params p = make_params(M, N, K);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = std::min(p.num_tiles, sms);
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_fp8_kernel<OUT_BF16_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_fp8_kernel<OUT_BF16_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_fp8_kernel<OUT_F32_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_fp8_kernel<OUT_F32_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const float* sa = a_scale.data_ptr<float>();
  const float* sb = b_scale.data_ptr<float>();
  if (f32 && accumulate) gemm_fp8_kernel<OUT_F32_ADD><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, sa, sb);
  else if (f32) gemm_fp8_kernel<OUT_F32_STORE><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, sa, sb);
  else if (accumulate) gemm_fp8_kernel<OUT_BF16_ADD><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, sa, sb);
  else gemm_fp8_kernel<OUT_BF16_STORE><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p, sa, sb);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
