// Vocab-parallel embedding for sm_100a.
//   forward : out[row(t), :] = W[id_t - vocab_start, :] if id_t is in this rank's vocab range, else 0
//             (partial sums over the TP group are then reduce-scattered / all-reduced); with `seq_first` the rows
//             are written directly in the sequence-first [S, B, H] layout sequence parallelism uses, so no
//             transpose copy is needed before the reduce-scatter
//   backward: main_grad[id_t - vocab_start, :] += g[row(t), :]   fp32 atomics straight into the flat gradient buffer
//             (no [V, H] temporary, no dtype round trip)
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace {

__device__ __forceinline__ int64_t out_row(int64_t t, int B, int S, int seq_first) {
  if (!seq_first) return t;
  const int64_t b = t / S, s = t - b * S;
  return s * B + b;
}

// one warp per token row chunk: 32 lanes x 8 bf16 = 256 elements per iteration
__global__ void __launch_bounds__(256) embedding_fwd_kernel(const int64_t* __restrict__ ids,
                                                            const __nv_bfloat16* __restrict__ w,
                                                            __nv_bfloat16* __restrict__ out, int64_t T, int H, int B,
                                                            int S, int64_t vocab_start, int64_t vocab_rows,
                                                            int seq_first) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < T; t += nwarps) {
    const int64_t id = ids[t] - vocab_start;
    const bool hit = id >= 0 && id < vocab_rows;
    const uint4* src = reinterpret_cast<const uint4*>(w + (hit ? id : 0) * (int64_t)H);
    uint4* dst = reinterpret_cast<uint4*>(out + out_row(t, B, S, seq_first) * (int64_t)H);
    for (int c = lane; c < H / 8; c += 32) dst[c] = hit ? src[c] : make_uint4(0u, 0u, 0u, 0u);
  }
}

__global__ void __launch_bounds__(256) embedding_bwd_kernel(const int64_t* __restrict__ ids,
                                                            const __nv_bfloat16* __restrict__ g,
                                                            float* __restrict__ dw, int64_t T, int H, int B, int S,
                                                            int64_t vocab_start, int64_t vocab_rows, int seq_first) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < T; t += nwarps) {
    const int64_t id = ids[t] - vocab_start;
    if (id < 0 || id >= vocab_rows) continue;
    const uint4* src = reinterpret_cast<const uint4*>(g + out_row(t, B, S, seq_first) * (int64_t)H);
    float* dst = dw + id * (int64_t)H;
    for (int c = lane; c < H / 8; c += 32) {
      const uint4 v = src[c];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        atomicAdd(dst + c * 8 + 2 * i, f.x);
        atomicAdd(dst + c * 8 + 2 * i + 1, f.y);
      }
    }
  }
}

}  // namespace

// ids [B, S] int64, w [V_local, H] bf16 -> out [B, S, H] (or [S, B, H] when seq_first)
at::Tensor embedding_fwd(const at::Tensor& ids, const at::Tensor& w, int64_t vocab_start, bool seq_first) {
  TORCH_CHECK(ids.is_cuda() && w.is_cuda() && ids.is_contiguous() && w.is_contiguous());
  TORCH_CHECK(ids.scalar_type() == at::kLong && w.scalar_type() == at::kBFloat16 && ids.dim() == 2 && w.dim() == 2);
  const int B = ids.size(0), S = ids.size(1), H = w.size(1);
  TORCH_CHECK(H % 8 == 0);
  c10::cuda::CUDAGuard guard(w.device());
  auto out = seq_first ? at::empty({S, B, H}, w.options()) : at::empty({B, S, H}, w.options());
  const int64_t T = (int64_t)B * S;
  if (T == 0) return out;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<int64_t>((T + 7) / 8, (int64_t)sms * 8);
  embedding_fwd_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      ids.data_ptr<int64_t>(), (const __nv_bfloat16*)w.data_ptr(), (__nv_bfloat16*)out.data_ptr(), T, H, B, S,
      vocab_start, w.size(0), seq_first ? 1 : 0);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

// dw [V_local, H] fp32 += scatter(g);  g [B, S, H] (or [S, B, H] when seq_first) bf16
void embedding_bwd_(at::Tensor dw, const at::Tensor& ids, const at::Tensor& g, int64_t vocab_start, bool seq_first) {
  TORCH_CHECK(dw.is_cuda() && g.is_cuda() && ids.is_cuda() && dw.is_contiguous() && g.is_contiguous() && ids.is_contiguous());
  TORCH_CHECK(dw.scalar_type() == at::kFloat && g.scalar_type() == at::kBFloat16 && ids.scalar_type() == at::kLong);
  const int B = ids.size(0), S = ids.size(1), H = dw.size(1);
  TORCH_CHECK(g.numel() == (int64_t)B * S * H && H % 8 == 0);
  c10::cuda::CUDAGuard guard(dw.device());
  const int64_t T = (int64_t)B * S;
  if (T == 0) return;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<int64_t>((T + 7) / 8, (int64_t)sms * 8);
  embedding_bwd_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      ids.data_ptr<int64_t>(), (const __nv_bfloat16*)g.data_ptr(), dw.data_ptr<float>(), T, H, B, S, vocab_start,
      dw.size(0), seq_first ? 1 : 0);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
