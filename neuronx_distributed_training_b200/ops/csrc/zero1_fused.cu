// Fused ZeRO-1 step over NVSwitch peer memory (data-parallel group), two kernels:
//
//   zero1_reduce   : reduce-scatter by PULL - every rank sums its own shard of the flat gradient buffer straight out of
//                    all DP peers' (symmetric) gradient buffers (128-bit peer loads, fp32 accumulate), keeps the
//                    reduced fp32 shard locally and produces its squared norm (for the global clip coefficient).
//   zero1_adam_push: partitioned AdamW on the shard (fp32 master / moments, per-segment lr and weight decay, device-side
//                    clip scale, optional stochastic rounding) and all-gather by PUSH - the updated bf16 parameters are
//                    written directly into every peer's (symmetric) parameter buffer.
// Cross-GPU ordering uses the same epoch handshake as the fused TP kernels: `reduce` starts once every peer has
// entered the step (its backward is complete, gradients final); `adam_push` ends with an all-peer "my pushes are
// visible" handshake, so when the kernel retires on a rank its parameter buffer is complete.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"

namespace z1 {

using namespace sm100;
constexpr int MAX_DP = 8;
constexpr int MAX_SEG = 8;

struct Peers {
  const void* grads[MAX_DP];
  __nv_bfloat16* params[MAX_DP];
  uint32_t* hs[MAX_DP];      // handshake arrays (one slot per source rank, two words: [2*r] enter, [2*r+1] done)
  uint32_t* hs_local;
  int rank, dp;
  uint32_t epoch;
};

struct Segs {
  long long lo[MAX_SEG], hi[MAX_SEG];      // shard-local element ranges
  float lr[MAX_SEG], wd[MAX_SEG];
  int n;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename GT>
__device__ __forceinline__ void load4(const GT* p, float (&f)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&f)[4]) {
  const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 v = __ldcg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

template <typename GT>
__global__ void __launch_bounds__(256) zero1_reduce_kernel(const Peers pr, float* __restrict__ gshard, long long shard_lo,
                                                           long long n, float scale, float* __restrict__ sq_out,
                                                           const unsigned char* __restrict__ count_mask) {
  // every rank has finished its backward (gradients are final) once it has entered this step
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int p = 0; p < pr.dp; ++p) st_release_sys(pr.hs[p] + 2 * pr.rank, pr.epoch);
  if (threadIdx.x == 0)
    for (int p = 0; p < pr.dp; ++p)
      spin_until_ge_sys(pr.hs_local + 2 * p, pr.epoch, 200);
  __syncthreads();
  float sq = 0.f;
  const long long n4 = n / 4;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n4; u += (long long)gridDim.x * blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // start at my own buffer and walk the ring so the peers are not all hit in the same order
    for (int d = 0; d < pr.dp; ++d) {
      int p = pr.rank + d;
      if (p >= pr.dp) p -= pr.dp;
      float f[4];
      load4<GT>(reinterpret_cast<const GT*>(pr.grads[p]) + shard_lo + u * 4, f);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += f[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] *= scale;
    *reinterpret_cast<float4*>(gshard + u * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    const float w = count_mask ? (float)count_mask[u * 4 / 64] : 1.f;     // 1 flag per 64 elements: counted in the norm?
    sq += w * (acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2] + acc[3] * acc[3]);
  }
  __shared__ float sm[8];
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    atomicAdd(sq_out, t);
  }
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(256) zero1_adam_push_kernel(const Peers pr, float* __restrict__ master, float* __restrict__ m,
                                                              float* __restrict__ v, const float* __restrict__ gshard,
                                                              long long shard_lo, const Segs segs, float b1, float b2,
                                                              float eps, float bc1, float bc2, const float* __restrict__ clip,
                                                              int sr, uint32_t seed, unsigned int* __restrict__ done_ctr) {
  const float gs = clip ? *clip : 1.f;
  const float inv_bc1 = 1.f / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int sidx = 0; sidx < segs.n; ++sidx) {
    const long long lo = segs.lo[sidx], n4 = (segs.hi[sidx] - lo) / 4;
    const float lr = segs.lr[sidx], decay = 1.f - lr * segs.wd[sidx];
    for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n4; u += (long long)gridDim.x * blockDim.x) {
      const long long i0 = lo + u * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(gshard + i0);
      float4 m4 = *reinterpret_cast<float4*>(m + i0), v4 = *reinterpret_cast<float4*>(v + i0);
      float4 p4 = *reinterpret_cast<float4*>(master + i0);
      float g[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
      float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pp[i] *= decay;
        mm[i] = b1 * mm[i] + (1.f - b1) * g[i];
        vv[i] = b2 * vv[i] + (1.f - b2) * g[i] * g[i];
        pp[i] -= lr * (mm[i] * inv_bc1) / (sqrtf(vv[i]) * inv_sqrt_bc2 + eps);
      }
      *reinterpret_cast<float4*>(m + i0) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      *reinterpret_cast<float4*>(master + i0) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      uint2 o;
      if (sr) {
        uint32_t b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = ((__float_as_uint(pp[i]) + (hash32(seed ^ hash32((uint32_t)(i0 + i))) & 0xFFFFu)) >> 16);
        o.x = b[0] | (b[1] << 16); o.y = b[2] | (b[3] << 16);
      } else {
        o.x = pack_bf16x2(pp[0], pp[1]); o.y = pack_bf16x2(pp[2], pp[3]);
      }
      // all-gather by push: the same 8 bytes go to every rank's parameter buffer (mine included)
      for (int d = 0; d < pr.dp; ++d) {
        int p = pr.rank + d;
        if (p >= pr.dp) p -= pr.dp;
        *reinterpret_cast<uint2*>(pr.params[p] + shard_lo + i0) = o;
      }
    }
  }
  // make my pushes visible everywhere, then the last CTA tells every peer; nobody leaves before all peers are done
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_ctr, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *done_ctr = 0;
    __threadfence_system();
    for (int p = 0; p < pr.dp; ++p) st_release_sys(pr.hs[p] + 2 * pr.rank + 1, pr.epoch);
    for (int p = 0; p < pr.dp; ++p)
      spin_until_ge_sys(pr.hs_local + 2 * p + 1, pr.epoch, 200);
  }
}

static void fill(Peers& pr, const std::vector<int64_t>& grads, const std::vector<int64_t>& params,
                 const std::vector<int64_t>& hs, int64_t rank, int64_t epoch) {
  const int dp = (int)grads.size();
  TORCH_CHECK(dp >= 1 && dp <= MAX_DP && (int)params.size() == dp && (int)hs.size() == dp);
  for (int i = 0; i < MAX_DP; ++i) {
    pr.grads[i] = i < dp ? reinterpret_cast<const void*>(grads[i]) : nullptr;
    pr.params[i] = i < dp ? reinterpret_cast<__nv_bfloat16*>(params[i]) : nullptr;
    pr.hs[i] = i < dp ? reinterpret_cast<uint32_t*>(hs[i]) : nullptr;
  }
  pr.hs_local = pr.hs[rank];
  pr.rank = (int)rank; pr.dp = dp; pr.epoch = (uint32_t)epoch;
}

}  // namespace z1

// gshard[n] (fp32) = scale * sum_p grads_p[shard_lo : shard_lo+n];  returns sum of squares (count_mask: 1 byte per 64 elems)
at::Tensor zero1_reduce(std::vector<int64_t> grads, std::vector<int64_t> params, std::vector<int64_t> hs, int64_t rank,
                        int64_t epoch, at::Tensor gshard, int64_t shard_lo, bool grads_fp32, double scale,
                        const at::Tensor& count_mask) {
  using namespace z1;
  TORCH_CHECK(gshard.is_cuda() && gshard.is_contiguous() && gshard.scalar_type() == at::kFloat);
  const long long n = gshard.numel();
  TORCH_CHECK(n % 64 == 0 && shard_lo % 64 == 0, "zero1: shard must be 64-element aligned");
  c10::cuda::CUDAGuard guard(gshard.device());
  Peers pr;
  fill(pr, grads, params, hs, rank, epoch);
  auto sq = at::zeros({}, gshard.options());
  const unsigned char* cm = count_mask.numel() ? count_mask.data_ptr<unsigned char>() : nullptr;
  const int grid = at::cuda::getCurrentDeviceProperties()->multiProcessorCount * 4;
  auto stream = at::cuda::getCurrentCUDAStream();
  if (grads_fp32)
    zero1_reduce_kernel<float><<<grid, 256, 0, stream>>>(pr, gshard.data_ptr<float>(), shard_lo, n, (float)scale, sq.data_ptr<float>(), cm);
  else
    zero1_reduce_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(pr, gshard.data_ptr<float>(), shard_lo, n, (float)scale, sq.data_ptr<float>(), cm);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return sq;
}

void zero1_adam_push(std::vector<int64_t> grads, std::vector<int64_t> params, std::vector<int64_t> hs, int64_t rank,
                     int64_t epoch, at::Tensor master, at::Tensor m, at::Tensor v, const at::Tensor& gshard, int64_t shard_lo,
                     std::vector<int64_t> seg_lo, std::vector<int64_t> seg_hi, std::vector<double> seg_lr,
                     std::vector<double> seg_wd, double b1, double b2, double eps, double bc1, double bc2,
                     const at::Tensor& clip, bool sr, int64_t seed, at::Tensor done_ctr) {
  using namespace z1;
  TORCH_CHECK(master.scalar_type() == at::kFloat && m.scalar_type() == at::kFloat && v.scalar_type() == at::kFloat);
  TORCH_CHECK(seg_lo.size() <= MAX_SEG && seg_lo.size() == seg_hi.size() && seg_lr.size() == seg_lo.size() && seg_wd.size() == seg_lo.size());
  c10::cuda::CUDAGuard guard(master.device());
  Peers pr;
  fill(pr, grads, params, hs, rank, epoch);
  Segs sg;
  sg.n = (int)seg_lo.size();
  for (int i = 0; i < sg.n; ++i) {
    TORCH_CHECK(seg_lo[i] % 4 == 0 && seg_hi[i] % 4 == 0, "zero1: segments must be 4-element aligned");
    sg.lo[i] = seg_lo[i]; sg.hi[i] = seg_hi[i]; sg.lr[i] = (float)seg_lr[i]; sg.wd[i] = (float)seg_wd[i];
  }
  const int grid = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;   // all CTAs resident (end-of-kernel handshake)
  zero1_adam_push_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      pr, master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), gshard.data_ptr<float>(), shard_lo, sg, (float)b1,
      (float)b2, (float)eps, (float)bc1, (float)bc2, clip.numel() ? clip.data_ptr<float>() : nullptr, sr ? 1 : 0, (uint32_t)seed,
      reinterpret_cast<unsigned int*>(done_ctr.data_ptr()));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
