// NVLS kernels: collectives through the NVSwitch MULTICAST mapping of a symmetric allocation (multimem.* PTX).
//
//   multimem.st         one store from this GPU is replicated by the switch into the same offset of EVERY rank's buffer
//   multimem.ld_reduce  one load returns the SUM over every rank's buffer at that offset (the switch reduces in flight)
//   multimem.red        one reduction-store is applied to every rank's copy (used here for arrival counters)
//
// 1. rmsnorm_allgather_mc: sequence-parallel RMSNorm whose output IS the all-gather.  Each rank normalises its [m, H]
//    rows and stores them ONCE through the multicast address into rows [rank*m, (rank+1)*m) of the gathered [tp*m, H]
//    buffer of all ranks -- the norm's write and the collective are the same instruction stream, the gathered activation
//    never exists as a local-only tensor and no all-gather kernel runs.  The kernel returns when the rows of ALL ranks
//    have landed locally (arrival counter bumped by every peer CTA with multimem.red, one thread spins on the local copy),
//    so plain stream order makes the consumer GEMM safe.  Reference call sites of the op pair it replaces:
//    modeling_llama.py:145-161 (norm) + the sequence-parallel gather inside ColumnParallelLinear.
// 2. nvls_allreduce: one-shot sum of a small fp32 tensor: every rank stores its operand into a symmetric staging region, one
//    barrier, every rank reads the whole reduced vector with multimem.ld_reduce.  One launch, no NCCL kernel: for the
//    latency-bound tensor-parallel reductions.
//
// All waits are bounded (≈20 s of globaltimer) and trap instead of hanging the GPU.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_bf16.h>

namespace nvls {

__device__ __forceinline__ void mc_st_v4(void* mc_addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const void* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st_f32x4(void* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_red_add_release(uint32_t* mc_addr, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until the local copy of an arrival counter reaches `target` (wrap-safe), trap after ~20 s (same bound as the
// fused TP kernels: long enough for a rank delayed by data loading, short enough that a lost peer ends the job)
__device__ __forceinline__ void wait_counter(const uint32_t* local_flag, uint32_t target) {
  const uint64_t t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys_u32(local_flag) - target) < 0) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > 20000000000ull) {
      printf("nvls: arrival counter timeout (have %u, want %u)\n", ld_acquire_sys_u32(local_flag), target);
      __trap();
    }
  }
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// One warp per row.  CH = 16-byte chunks per lane (H = 256 * CH): the row stays in registers between the sum of squares
// and the stores, the CH multicast stores of a lane are issued back to back.  CH = 0: any H % 256 == 0 (row re-read from L1).
template <int WARPS, int CH>
__global__ void __launch_bounds__(WARPS * 32) rmsnorm_allgather_mc_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, uint8_t* mc_out, float* __restrict__ rstd_out,
    int m, int H, float eps, int rank, uint32_t* mc_flag, const uint32_t* local_flag, uint32_t target) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  for (int row = blockIdx.x * WARPS + warp; row < m; row += gridDim.x * WARPS) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * H);
    uint8_t* dst = mc_out + ((size_t)rank * m + row) * H * 2;
    float ss = 0.f;
    if constexpr (CH > 0) {
      uint4 v[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) v[i] = xr[lane + 32 * i];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float f[8];
        unpack8(v[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss / H + eps);
      if (lane == 0) rstd_out[row] = rstd;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float f[8], g[8];
        unpack8(v[i], f);
        unpack8(wr[lane + 32 * i], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] * rstd * g[j];
        mc_st_v4(dst + (size_t)(lane + 32 * i) * 16, pack8(f));     // lands in every rank's gathered buffer
      }
    } else {
      const int chunks = H / 8;
      for (int c = lane; c < chunks; c += 32) {
        float f[8];
        unpack8(xr[c], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss / H + eps);
      if (lane == 0) rstd_out[row] = rstd;
      for (int c = lane; c < chunks; c += 32) {
        float f[8], g[8];
        unpack8(xr[c], f);
        unpack8(wr[c], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] * rstd * g[j];
        mc_st_v4(dst + (size_t)c * 16, pack8(f));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    mc_red_add_release(mc_flag, 1u);                              // +1 on every rank's counter
    if (blockIdx.x == 0) wait_counter(local_flag, target);        // all CTAs of all ranks have published their rows
  }
}

// One-shot all-reduce (sum) of n fp32 values: copy the operand into this rank's slot of the symmetric region, one barrier,
// then EVERY rank reads the whole reduced vector with multimem.ld_reduce (the switch adds the ranks' slots in flight) straight
// into its output.  The caller alternates between two regions, which makes a trailing barrier unnecessary (a rank can only
// be one call ahead of the slowest one).  n % 4 == 0.
__global__ void __launch_bounds__(256) nvls_allreduce_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             uint8_t* local_region, const uint8_t* mc_region, int64_t n,
                                                             uint32_t* mc_flag, const uint32_t* local_flag, uint32_t target) {
  const int64_t n4 = n / 4;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  float4* loc = reinterpret_cast<float4*>(local_region);
  for (int64_t i = tid; i < n4; i += nth) loc[i] = reinterpret_cast<const float4*>(in)[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    mc_red_add_release(mc_flag, 1u);
    wait_counter(local_flag, target);          // every CTA of every rank has stored its part of the operand
  }
  __syncthreads();
  const float4* mc = reinterpret_cast<const float4*>(mc_region);
  for (int64_t i = tid; i < n4; i += nth) reinterpret_cast<float4*>(out)[i] = mc_ld_reduce_f32x4(mc + i);
}

}  // namespace nvls

// CTAs of rmsnorm_allgather_mc for m local rows: a function of m ONLY (every rank must compute the same arrival target)
int64_t nvls_grid_for_rows(int64_t m) { return std::max<int64_t>(1, std::min<int64_t>((m + 3) / 4, 2368)); }

// x [m, H] bf16 local rows, w [H]; gathered rows are written through `mc_out_ptr` (multicast address of the gathered
// [world*m, H] buffer inside a symmetric allocation).  `mc_flag_ptr` / `local_flag_ptr`: multicast and local address of one
// u32 arrival counter; `target` = counter value after this call completes on every rank (the caller adds world * grid per
// call; grid is returned so the caller can do that bookkeeping).  Returns rstd [m].
at::Tensor rmsnorm_allgather_mc(const at::Tensor& x, const at::Tensor& w, double eps, int64_t mc_out_ptr, int64_t rank,
                                int64_t world, int64_t mc_flag_ptr, int64_t local_flag_ptr, int64_t counter_before) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.is_contiguous() && w.is_contiguous());
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16, "rmsnorm_allgather_mc: bf16 only");
  const int m = x.size(0), H = x.size(1);
  TORCH_CHECK(H % 256 == 0, "rmsnorm_allgather_mc: hidden size must be a multiple of 256");
  c10::cuda::CUDAGuard guard(x.device());
  auto rstd = at::empty({m}, x.options().dtype(at::kFloat));
  auto stream = at::cuda::getCurrentCUDAStream();
  constexpr int WARPS = 4;
  const int grid = (int)nvls_grid_for_rows(m);
  const uint32_t target = (uint32_t)counter_before + (uint32_t)(world * grid);
#define NVLS_NORM(CH)                                                                                                       \
  nvls::rmsnorm_allgather_mc_kernel<WARPS, CH><<<grid, WARPS * 32, 0, stream>>>(                                            \
      (const __nv_bfloat16*)x.data_ptr(), (const __nv_bfloat16*)w.data_ptr(), reinterpret_cast<uint8_t*>(mc_out_ptr),       \
      rstd.data_ptr<float>(), m, H, (float)eps, (int)rank, reinterpret_cast<uint32_t*>(mc_flag_ptr),                        \
      reinterpret_cast<const uint32_t*>(local_flag_ptr), target)
  switch (H / 256) {
    case 8: NVLS_NORM(8); break;
    case 16: NVLS_NORM(16); break;
    default: NVLS_NORM(0); break;
  }
#undef NVLS_NORM
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return rstd;
}

// out = sum over ranks of in (n fp32 values, n % 4 == 0, n * 4 <= region bytes).  `local_region_ptr` / `mc_region_ptr`: local and
// multicast address of the staging region inside the symmetric allocation (the caller alternates between two regions);
// one barrier event with 8 arrivals (CTAs) per rank.
void nvls_allreduce(const at::Tensor& in, at::Tensor& out, int64_t local_region_ptr, int64_t mc_region_ptr, int64_t world,
                    int64_t mc_flag_ptr, int64_t local_flag_ptr, int64_t counter_before) {
  TORCH_CHECK(in.is_cuda() && out.is_cuda() && in.scalar_type() == at::kFloat && out.scalar_type() == at::kFloat);
  TORCH_CHECK(in.is_contiguous() && out.is_contiguous() && in.numel() == out.numel() && in.numel() % 4 == 0);
  c10::cuda::CUDAGuard guard(in.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  constexpr int GRID = 8;
  const uint32_t target = (uint32_t)counter_before + (uint32_t)(world * GRID);
  nvls::nvls_allreduce_kernel<<<GRID, 256, 0, stream>>>(in.data_ptr<float>(), out.data_ptr<float>(),
                                                        reinterpret_cast<uint8_t*>(local_region_ptr),
                                                        reinterpret_cast<const uint8_t*>(mc_region_ptr), in.numel(),
                                                        reinterpret_cast<uint32_t*>(mc_flag_ptr),
                                                        reinterpret_cast<const uint32_t*>(local_flag_ptr), target);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
int64_t nvls_allreduce_arrivals() { return 8; }
