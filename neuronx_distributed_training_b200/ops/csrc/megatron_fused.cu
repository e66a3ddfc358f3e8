// Fused bandwidth-bound kernels of the Megatron-GPT model family for sm_100a:
//   * LayerNorm / LayerNorm1P forward + backward (fp32 statistics, dW/dB reduced in registers + one atomic pass)
//   * bias + activation (gelu / geglu / reglu / swiglu / relu / squared-relu / silu) forward + backward
//   * scale + causal / sliding-window mask + softmax forward + backward (the eager-attention path that the
//     reference runs through apex FusedScaleMaskSoftmax / the Neuron fused-softmax custom op)
//   * bias + dropout + residual add (Philox mask kept as one byte per element) and the dropout backward
// All kernels read/write 128-bit vectors and keep the math in fp32.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/CUDAGeneratorImpl.h>
#include <ATen/cuda/CUDAGraphsUtils.cuh>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int THREADS, bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, float* sm) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (threadIdx.x < THREADS / 32) ? sm[threadIdx.x] : (IS_MAX ? -INFINITY : 0.f);
  if (w == 0) {
    r = IS_MAX ? warp_max(r) : warp_sum(r);
    if (l == 0) sm[0] = r;
  }
  __syncthreads();
  r = sm[0];
  __syncthreads();
  return r;
}

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  bf16x8 v = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(v.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  bf16x8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<bf16x8*>(p) = v;
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

inline int grid_for(size_t units, int threads, int per_sm) {
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const size_t want = (units + threads - 1) / threads;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sms * per_sm));
}

// ------------------------------------------------------------------------------------------------
// LayerNorm
// ------------------------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                const T* __restrict__ b, T* __restrict__ y,
                                                                float* __restrict__ mean_out,
                                                                float* __restrict__ rstd_out, int rows, int H,
                                                                float eps, float w_shift) {
  __shared__ float sm[32];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + (size_t)row * H;
    float s = 0.f;
    for (int c = threadIdx.x * 8; c < H; c += THREADS * 8) {
      float f[8];
      load8<T>(xr + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[i];
    }
    const float mean = block_reduce<THREADS, false>(s, sm) / H;
    float ss = 0.f;
    for (int c = threadIdx.x * 8; c < H; c += THREADS * 8) {
      float f[8];
      load8<T>(xr + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += (f[i] - mean) * (f[i] - mean);
    }
    const float rstd = rsqrtf(block_reduce<THREADS, false>(ss, sm) / H + eps);
    if (threadIdx.x == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    T* yr = y + (size_t)row * H;
    for (int c = threadIdx.x * 8; c < H; c += THREADS * 8) {
      float f[8], g[8], bb[8];
      load8<T>(xr + c, f);
      load8<T>(w + c, g);
      if (b != nullptr) load8<T>(b + c, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * rstd * (g[i] + w_shift) + (b != nullptr ? bb[i] : 0.f);
      store8<T>(yr + c, f);
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w;   dw += dy * xhat;  db += dy
template <typename T, int THREADS, int MAXV>
__global__ void __launch_bounds__(THREADS) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                const T* __restrict__ w, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, T* __restrict__ dx,
                                                                float* __restrict__ dw, float* __restrict__ db, int rows,
                                                                int H, float w_shift) {
  __shared__ float sm[32];
  float dw_acc[MAXV][8], db_acc[MAXV][8];
#pragma unroll
  for (int v = 0; v < MAXV; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) dw_acc[v][i] = db_acc[v][i] = 0.f;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + (size_t)row * H;
    const T* gr = dy + (size_t)row * H;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      const int c = (v * THREADS + threadIdx.x) * 8;
      if (c < H) {
        float fx[8], fg[8], fw[8];
        load8<T>(xr + c, fx);
        load8<T>(gr + c, fg);
        load8<T>(w + c, fw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (fx[i] - mu) * rs;
          const float g = fg[i] * (fw[i] + w_shift);
          s1 += g;
          s2 += g * xh;
          dw_acc[v][i] += fg[i] * xh;
          db_acc[v][i] += fg[i];
        }
      }
    }
    s1 = block_reduce<THREADS, false>(s1, sm) / H;
    s2 = block_reduce<THREADS, false>(s2, sm) / H;
    T* dxr = dx + (size_t)row * H;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      const int c = (v * THREADS + threadIdx.x) * 8;
      if (c < H) {
        float fx[8], fg[8], fw[8], o[8];
        load8<T>(xr + c, fx);
        load8<T>(gr + c, fg);
        load8<T>(w + c, fw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (fx[i] - mu) * rs;
          o[i] = rs * (fg[i] * (fw[i] + w_shift) - s1 - xh * s2);
        }
        store8<T>(dxr + c, o);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < MAXV; ++v) {
    const int c = (v * THREADS + threadIdx.x) * 8;
    if (c < H) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(dw + c + i, dw_acc[v][i]);
        atomicAdd(db + c + i, db_acc[v][i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bias + activation.  GLU modes read [rows, 2I] (gate || up) and write [rows, I].
// ------------------------------------------------------------------------------------------------
enum ActMode : int { ACT_GELU = 0, ACT_GEGLU = 1, ACT_REGLU = 2, ACT_SWIGLU = 3, ACT_RELU = 4, ACT_SQRELU = 5, ACT_SILU = 6 };

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k * (x + c * x * x * x)));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  const float t = tanhf(k * (x + c * x * x * x));
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * (1.f + 3.f * c * x * x);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

template <int MODE>
__device__ __forceinline__ float act_f(float x) {
  if constexpr (MODE == ACT_GELU || MODE == ACT_GEGLU) return gelu_tanh(x);
  if constexpr (MODE == ACT_REGLU || MODE == ACT_RELU) return fmaxf(x, 0.f);
  if constexpr (MODE == ACT_SQRELU) { const float r = fmaxf(x, 0.f); return r * r; }
  return x * sigmoidf_(x);   // swiglu / silu
}
template <int MODE>
__device__ __forceinline__ float act_df(float x) {
  if constexpr (MODE == ACT_GELU || MODE == ACT_GEGLU) return gelu_tanh_grad(x);
  if constexpr (MODE == ACT_REGLU || MODE == ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if constexpr (MODE == ACT_SQRELU) return x > 0.f ? 2.f * x : 0.f;
  const float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}
template <int MODE>
constexpr bool is_glu() { return MODE == ACT_GEGLU || MODE == ACT_REGLU || MODE == ACT_SWIGLU; }

template <typename T, int MODE>
__global__ void __launch_bounds__(256) bias_act_fwd_kernel(const T* __restrict__ x, const T* __restrict__ bias,
                                                           T* __restrict__ out, size_t rows, int C) {
  constexpr bool GLU = is_glu<MODE>();
  const int I = GLU ? C / 2 : C;
  const size_t units = rows * (size_t)(I / 8);
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < units; u += (size_t)gridDim.x * blockDim.x) {
    const size_t r = u / (I / 8);
    const int c = (int)(u % (I / 8)) * 8;
    float g[8], o[8];
    load8<T>(x + r * C + c, g);
    if (bias != nullptr) {
      float bb[8];
      load8<T>(bias + c, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += bb[i];
    }
    if constexpr (GLU) {
      float up[8];
      load8<T>(x + r * C + I + c, up);
      if (bias != nullptr) {
        float bb[8];
        load8<T>(bias + I + c, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) up[i] += bb[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = act_f<MODE>(g[i]) * up[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = act_f<MODE>(g[i]);
    }
    store8<T>(out + r * I + c, o);
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) bias_act_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                                                           const T* __restrict__ bias, T* __restrict__ dx, size_t rows,
                                                           int C) {
  constexpr bool GLU = is_glu<MODE>();
  const int I = GLU ? C / 2 : C;
  const size_t units = rows * (size_t)(I / 8);
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < units; u += (size_t)gridDim.x * blockDim.x) {
    const size_t r = u / (I / 8);
    const int c = (int)(u % (I / 8)) * 8;
    float g[8], d[8], o[8];
    load8<T>(x + r * C + c, g);
    load8<T>(dout + r * I + c, d);
    if (bias != nullptr) {
      float bb[8];
      load8<T>(bias + c, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += bb[i];
    }
    if constexpr (GLU) {
      float up[8], o2[8];
      load8<T>(x + r * C + I + c, up);
      if (bias != nullptr) {
        float bb[8];
        load8<T>(bias + I + c, bb);
#pragma unroll
        for (int i = 0; i < 8; ++i) up[i] += bb[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = d[i] * up[i] * act_df<MODE>(g[i]);
        o2[i] = d[i] * act_f<MODE>(g[i]);
      }
      store8<T>(dx + r * C + c, o);
      store8<T>(dx + r * C + I + c, o2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = d[i] * act_df<MODE>(g[i]);
      store8<T>(dx + r * C + c, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// scale + mask + softmax over the last dim of [R, Sq, Sk] (R = batch * heads).  Query row i (absolute position
// i + q_offset) sees keys k <= pos (causal) and k > pos - window (window > 0).  Fully masked rows give zeros.
// ------------------------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) softmax_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, size_t rows,
                                                              int Sq, int Sk, float scale, int causal, int window,
                                                              int q_offset) {
  __shared__ float sm[32];
  for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int qi = (int)(row % Sq) + q_offset;
    int hi = causal ? min(Sk, qi + 1) : Sk;                 // exclusive
    int lo = window > 0 ? max(0, qi - window + 1) : 0;      // inclusive
    if (hi < 0) hi = 0;
    const T* xr = x + row * Sk;
    T* yr = y + row * Sk;
    float mx = -INFINITY;
    for (int c = threadIdx.x * 8; c < Sk; c += THREADS * 8) {
      if (c + 8 <= lo || c >= hi) continue;
      float f[8];
      load8<T>(xr + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c + i >= lo && c + i < hi) mx = fmaxf(mx, f[i] * scale);
    }
    mx = block_reduce<THREADS, true>(mx, sm);
    float sum = 0.f;
    for (int c = threadIdx.x * 8; c < Sk; c += THREADS * 8) {
      if (c + 8 <= lo || c >= hi) continue;
      float f[8];
      load8<T>(xr + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c + i >= lo && c + i < hi) sum += __expf(f[i] * scale - mx);
    }
    sum = block_reduce<THREADS, false>(sum, sm);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    for (int c = threadIdx.x * 8; c < Sk; c += THREADS * 8) {
      float f[8], o[8];
      const bool any = !(c + 8 <= lo || c >= hi);
      if (any) load8<T>(xr + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (any && c + i >= lo && c + i < hi) ? __expf(f[i] * scale - mx) * inv : 0.f;
      store8<T>(yr + c, o);
    }
  }
}

// dx = scale * y * (dy - sum(dy * y))
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) softmax_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                              T* __restrict__ dx, size_t rows, int Sk, float scale) {
  __shared__ float sm[32];
  for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* gr = dy + row * Sk;
    const T* yr = y + row * Sk;
    float dot = 0.f;
    for (int c = threadIdx.x * 8; c < Sk; c += THREADS * 8) {
      float g[8], p[8];
      load8<T>(gr + c, g);
      load8<T>(yr + c, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) dot += g[i] * p[i];
    }
    dot = block_reduce<THREADS, false>(dot, sm);
    T* dr = dx + row * Sk;
    for (int c = threadIdx.x * 8; c < Sk; c += THREADS * 8) {
      float g[8], p[8], o[8];
      load8<T>(gr + c, g);
      load8<T>(yr + c, p);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = scale * p[i] * (g[i] - dot);
      store8<T>(dr + c, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// out = residual + dropout(x + bias, p);  mask (1 byte / element) is kept for the backward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bias_dropout_add_kernel(const T* __restrict__ x, const T* __restrict__ bias,
                                                               const T* __restrict__ residual, T* __restrict__ out,
                                                               uint8_t* __restrict__ mask, size_t numel, int H, float p,
                                                               at::PhiloxCudaState philox) {
  const auto seeds = at::cuda::philox::unpack(philox);
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  curandStatePhilox4_32_10_t state;
  curand_init(std::get<0>(seeds), tid, std::get<1>(seeds), &state);
  const float keep_scale = 1.f / (1.f - p);
  const size_t units = numel / 8;
  for (size_t u = tid; u < units; u += (size_t)gridDim.x * blockDim.x) {
    const size_t e = u * 8;
    float f[8], r[8], o[8];
    load8<T>(x + e, f);
    load8<T>(residual + e, r);
    if (bias != nullptr) {
      float bb[8];
      load8<T>(bias + (e % H), bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += bb[i];
    }
    const float4 r0 = curand_uniform4(&state), r1 = curand_uniform4(&state);
    const float rnd[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    uint8_t m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      m[i] = rnd[i] >= p ? 1 : 0;
      o[i] = r[i] + (m[i] ? f[i] * keep_scale : 0.f);
    }
    store8<T>(out + e, o);
    *reinterpret_cast<uint2*>(mask + e) = *reinterpret_cast<const uint2*>(m);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const T* __restrict__ dout, const uint8_t* __restrict__ mask,
                                                          T* __restrict__ dx, size_t numel, float keep_scale) {
  const size_t units = numel / 8;
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < units; u += (size_t)gridDim.x * blockDim.x) {
    const size_t e = u * 8;
    float d[8];
    load8<T>(dout + e, d);
    const uint2 mm = *reinterpret_cast<const uint2*>(mask + e);
    const uint8_t* m = reinterpret_cast<const uint8_t*>(&mm);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = m[i] ? d[i] * keep_scale : 0.f;
    store8<T>(dx + e, d);
  }
}

}  // namespace

// ====================================================================================================
#define DISPATCH_FLOAT_BF16(T_, ...)                          \
  if ((T_) == at::kBFloat16) { using scalar_t = __nv_bfloat16; __VA_ARGS__; } \
  else if ((T_) == at::kFloat) { using scalar_t = float; __VA_ARGS__; }       \
  else { TORCH_CHECK(false, "bf16 or fp32 tensors required"); }

std::vector<at::Tensor> layernorm_fwd(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& b,
                                      double eps, bool one_plus) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.is_contiguous() && w.is_contiguous());
  TORCH_CHECK(x.scalar_type() == w.scalar_type(), "layernorm: x and weight dtypes must match");
  const int rows = x.size(0), H = x.size(1);
  TORCH_CHECK(H % 8 == 0);
  if (b.has_value()) TORCH_CHECK(b->scalar_type() == x.scalar_type() && b->is_contiguous() && b->numel() == H);
  c10::cuda::CUDAGuard guard(x.device());
  auto y = at::empty_like(x);
  auto mean = at::empty({rows}, x.options().dtype(at::kFloat));
  auto rstd = at::empty({rows}, x.options().dtype(at::kFloat));
  if (rows == 0) return {y, mean, rstd};
  auto stream = at::cuda::getCurrentCUDAStream();
  constexpr int TH = 256;
  const int grid = std::min(rows, at::cuda::getCurrentDeviceProperties()->multiProcessorCount * 8);
  DISPATCH_FLOAT_BF16(x.scalar_type(), layernorm_fwd_kernel<scalar_t, TH><<<grid, TH, 0, stream>>>(
      (const scalar_t*)x.data_ptr(), (const scalar_t*)w.data_ptr(),
      b.has_value() ? (const scalar_t*)b->data_ptr() : nullptr, (scalar_t*)y.data_ptr(), mean.data_ptr<float>(),
      rstd.data_ptr<float>(), rows, H, (float)eps, one_plus ? 1.f : 0.f));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {y, mean, rstd};
}

std::vector<at::Tensor> layernorm_bwd(const at::Tensor& dy, const at::Tensor& x, const at::Tensor& w,
                                      const at::Tensor& mean, const at::Tensor& rstd, bool one_plus) {
  TORCH_CHECK(dy.is_contiguous() && x.is_contiguous() && dy.scalar_type() == x.scalar_type());
  const int rows = x.size(0), H = x.size(1);
  c10::cuda::CUDAGuard guard(x.device());
  auto dx = at::empty_like(x);
  auto dw = at::zeros({H}, x.options().dtype(at::kFloat));
  auto db = at::zeros({H}, x.options().dtype(at::kFloat));
  if (rows == 0) return {dx, dw, db};
  auto stream = at::cuda::getCurrentCUDAStream();
  constexpr int TH = 256;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = std::min(rows, sms * 4);
  const int need_v = (H + TH * 8 - 1) / (TH * 8);
  TORCH_CHECK(need_v <= 4, "layernorm_bwd: hidden size > 8192 not supported by the native kernel");
  const float shift = one_plus ? 1.f : 0.f;
#define RUN(MAXV)                                                                                                    \
  DISPATCH_FLOAT_BF16(x.scalar_type(), layernorm_bwd_kernel<scalar_t, TH, MAXV><<<grid, TH, 0, stream>>>(             \
      (const scalar_t*)dy.data_ptr(), (const scalar_t*)x.data_ptr(), (const scalar_t*)w.data_ptr(),                  \
      mean.data_ptr<float>(), rstd.data_ptr<float>(), (scalar_t*)dx.data_ptr(), dw.data_ptr<float>(),                \
      db.data_ptr<float>(), rows, H, shift))
  if (need_v <= 1) { RUN(1); } else if (need_v <= 2) { RUN(2); } else { RUN(4); }
#undef RUN
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {dx, dw, db};
}

#define DISPATCH_ACT(MODE_, ...)                                              \
  switch (MODE_) {                                                            \
    case ACT_GELU: { constexpr int kMode = ACT_GELU; __VA_ARGS__; break; }    \
    case ACT_GEGLU: { constexpr int kMode = ACT_GEGLU; __VA_ARGS__; break; }  \
    case ACT_REGLU: { constexpr int kMode = ACT_REGLU; __VA_ARGS__; break; }  \
    case ACT_SWIGLU: { constexpr int kMode = ACT_SWIGLU; __VA_ARGS__; break; }\
    case ACT_RELU: { constexpr int kMode = ACT_RELU; __VA_ARGS__; break; }    \
    case ACT_SQRELU: { constexpr int kMode = ACT_SQRELU; __VA_ARGS__; break; }\
    case ACT_SILU: { constexpr int kMode = ACT_SILU; __VA_ARGS__; break; }    \
    default: TORCH_CHECK(false, "bias_act: unknown activation mode ", MODE_); \
  }

static bool mode_is_glu(int64_t mode) { return mode == ACT_GEGLU || mode == ACT_REGLU || mode == ACT_SWIGLU; }

at::Tensor bias_act_fwd(const at::Tensor& x, const c10::optional<at::Tensor>& bias, int64_t mode) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.dim() >= 1);
  const int C = x.size(-1);
  const bool glu = mode_is_glu(mode);
  const int I = glu ? C / 2 : C;
  TORCH_CHECK(I % 8 == 0 && (!glu || C % 2 == 0), "bias_act: feature dim must be a multiple of 8");
  if (bias.has_value()) TORCH_CHECK(bias->scalar_type() == x.scalar_type() && bias->is_contiguous() && bias->numel() == C);
  const size_t rows = x.numel() / C;
  c10::cuda::CUDAGuard guard(x.device());
  auto sizes = x.sizes().vec();
  sizes.back() = I;
  auto out = at::empty(sizes, x.options());
  if (rows == 0) return out;
  auto stream = at::cuda::getCurrentCUDAStream();
  const int grid = grid_for(rows * (size_t)(I / 8), 256, 8);
  DISPATCH_ACT(mode, DISPATCH_FLOAT_BF16(x.scalar_type(), bias_act_fwd_kernel<scalar_t, kMode><<<grid, 256, 0, stream>>>(
      (const scalar_t*)x.data_ptr(), bias.has_value() ? (const scalar_t*)bias->data_ptr() : nullptr,
      (scalar_t*)out.data_ptr(), rows, C)));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

at::Tensor bias_act_bwd(const at::Tensor& dout, const at::Tensor& x, const c10::optional<at::Tensor>& bias, int64_t mode) {
  TORCH_CHECK(dout.is_cuda() && dout.is_contiguous() && x.is_contiguous() && dout.scalar_type() == x.scalar_type());
  const int C = x.size(-1);
  const bool glu = mode_is_glu(mode);
  const int I = glu ? C / 2 : C;
  TORCH_CHECK(dout.size(-1) == I && dout.numel() / I == x.numel() / C);
  const size_t rows = x.numel() / C;
  c10::cuda::CUDAGuard guard(x.device());
  auto dx = at::empty_like(x);
  if (rows == 0) return dx;
  auto stream = at::cuda::getCurrentCUDAStream();
  const int grid = grid_for(rows * (size_t)(I / 8), 256, 8);
  DISPATCH_ACT(mode, DISPATCH_FLOAT_BF16(x.scalar_type(), bias_act_bwd_kernel<scalar_t, kMode><<<grid, 256, 0, stream>>>(
      (const scalar_t*)dout.data_ptr(), (const scalar_t*)x.data_ptr(),
      bias.has_value() ? (const scalar_t*)bias->data_ptr() : nullptr, (scalar_t*)dx.data_ptr(), rows, C)));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return dx;
}

at::Tensor scale_mask_softmax_fwd(const at::Tensor& x, double scale, bool causal, int64_t window, int64_t q_offset) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && x.dim() >= 2);
  const int Sk = x.size(-1), Sq = x.size(-2);
  TORCH_CHECK(Sk % 8 == 0, "softmax: key length must be a multiple of 8");
  const size_t rows = x.numel() / Sk;
  c10::cuda::CUDAGuard guard(x.device());
  auto y = at::empty_like(x);
  if (rows == 0) return y;
  auto stream = at::cuda::getCurrentCUDAStream();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<size_t>(rows, (size_t)sms * 16);
  if (Sk <= 2048) {
    DISPATCH_FLOAT_BF16(x.scalar_type(), softmax_fwd_kernel<scalar_t, 128><<<grid, 128, 0, stream>>>(
        (const scalar_t*)x.data_ptr(), (scalar_t*)y.data_ptr(), rows, Sq, Sk, (float)scale, causal ? 1 : 0, (int)window,
        (int)q_offset));
  } else {
    DISPATCH_FLOAT_BF16(x.scalar_type(), softmax_fwd_kernel<scalar_t, 256><<<grid, 256, 0, stream>>>(
        (const scalar_t*)x.data_ptr(), (scalar_t*)y.data_ptr(), rows, Sq, Sk, (float)scale, causal ? 1 : 0, (int)window,
        (int)q_offset));
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return y;
}

at::Tensor scale_mask_softmax_bwd(const at::Tensor& dy, const at::Tensor& y, double scale) {
  TORCH_CHECK(dy.is_cuda() && dy.is_contiguous() && y.is_contiguous() && dy.scalar_type() == y.scalar_type());
  const int Sk = y.size(-1);
  const size_t rows = y.numel() / Sk;
  c10::cuda::CUDAGuard guard(y.device());
  auto dx = at::empty_like(y);
  if (rows == 0) return dx;
  auto stream = at::cuda::getCurrentCUDAStream();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<size_t>(rows, (size_t)sms * 16);
  if (Sk <= 2048) {
    DISPATCH_FLOAT_BF16(y.scalar_type(), softmax_bwd_kernel<scalar_t, 128><<<grid, 128, 0, stream>>>(
        (const scalar_t*)dy.data_ptr(), (const scalar_t*)y.data_ptr(), (scalar_t*)dx.data_ptr(), rows, Sk, (float)scale));
  } else {
    DISPATCH_FLOAT_BF16(y.scalar_type(), softmax_bwd_kernel<scalar_t, 256><<<grid, 256, 0, stream>>>(
        (const scalar_t*)dy.data_ptr(), (const scalar_t*)y.data_ptr(), (scalar_t*)dx.data_ptr(), rows, Sk, (float)scale));
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return dx;
}

std::vector<at::Tensor> bias_dropout_add_fwd(const at::Tensor& x, const c10::optional<at::Tensor>& bias,
                                             const at::Tensor& residual, double p) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && residual.is_contiguous() && x.sizes() == residual.sizes());
  TORCH_CHECK(x.scalar_type() == residual.scalar_type() && p >= 0.0 && p < 1.0);
  const int H = x.size(-1);
  TORCH_CHECK(H % 8 == 0);
  if (bias.has_value()) TORCH_CHECK(bias->scalar_type() == x.scalar_type() && bias->is_contiguous() && bias->numel() == H);
  const size_t numel = x.numel();
  c10::cuda::CUDAGuard guard(x.device());
  auto out = at::empty_like(x);
  auto mask = at::empty(x.sizes(), x.options().dtype(at::kByte));
  if (numel == 0) return {out, mask};
  auto stream = at::cuda::getCurrentCUDAStream();
  const int grid = grid_for(numel / 8, 256, 8);
  // every thread draws 8 randoms per grid-stride iteration
  const uint64_t iters = (numel / 8 + (size_t)grid * 256 - 1) / ((size_t)grid * 256);
  auto gen = at::get_generator_or_default<at::CUDAGeneratorImpl>(c10::nullopt, at::cuda::detail::getDefaultCUDAGenerator());
  at::PhiloxCudaState philox;
  {
    std::lock_guard<std::mutex> lock(gen->mutex_);
    philox = gen->philox_cuda_state(iters * 8);
  }
  DISPATCH_FLOAT_BF16(x.scalar_type(), bias_dropout_add_kernel<scalar_t><<<grid, 256, 0, stream>>>(
      (const scalar_t*)x.data_ptr(), bias.has_value() ? (const scalar_t*)bias->data_ptr() : nullptr,
      (const scalar_t*)residual.data_ptr(), (scalar_t*)out.data_ptr(), mask.data_ptr<uint8_t>(), numel, H, (float)p,
      philox));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {out, mask};
}

at::Tensor dropout_bwd(const at::Tensor& dout, const at::Tensor& mask, double p) {
  TORCH_CHECK(dout.is_cuda() && dout.is_contiguous() && mask.is_contiguous() && mask.scalar_type() == at::kByte);
  TORCH_CHECK(dout.numel() == mask.numel() && dout.numel() % 8 == 0);
  c10::cuda::CUDAGuard guard(dout.device());
  auto dx = at::empty_like(dout);
  const size_t numel = dout.numel();
  if (numel == 0) return dx;
  auto stream = at::cuda::getCurrentCUDAStream();
  const int grid = grid_for(numel / 8, 256, 8);
  DISPATCH_FLOAT_BF16(dout.scalar_type(), dropout_bwd_kernel<scalar_t><<<grid, 256, 0, stream>>>(
      (const scalar_t*)dout.data_ptr(), mask.data_ptr<uint8_t>(), (scalar_t*)dx.data_ptr(), numel,
      (float)(1.0 / (1.0 - p))));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return dx;
}
