// RMSNorm ("T5 layer norm": no mean subtraction, no bias) for the decoder layers around the W4A16 linears.
//
// Replaces the reference's generalT5LayerNorm (awq/kernels/csrc/layernorm/layernorm.cu:38-64, host :111-131,
// exported as `layernorm_forward_cuda` at csrc/pybind.cpp:17, called from tinychat/models/llama.py:33-36 and
// tinychat/modules/fused_norm.py:16-21).  Numerical contract of the reference, kept here:
//     var  = sum_i float(x_i)^2            (fp32)
//     r    = rsqrtf(var / n + eps)
//     y_i  = rn_T( (float(x_i) * r) * float(gamma_i) )     fp16 additionally clamped to +-(65504 - 1000)
// (layernorm.cu:50-62, reduction.cuh:72-82).  Only the summation order of `var` differs.
//
// HBM-bound: algorithmic bytes = 2 * m * n * sizeof(T) (+ n * sizeof(T) of gamma, L2-resident).  One CTA per row;
// the row is read ONCE with 16-byte loads and kept in registers between the reduction and the scaling (the
// reference reads it twice with 2-byte loads); all loads of a thread are issued before the first use.  Rows wider
// than the register budget, or not a multiple of 8 elements, take the generic two-pass kernel.
#include "w4_common.cuh"
#include "w4a16_kernels.h"

namespace b200awq {
namespace {

constexpr int kNormThreads = 256;

template <typename T>
__device__ __forceinline__ float2 pair_to_float2(uint32_t v) {
  if constexpr (TypeTraits<T>::kIsBf16) return __bfloat1622float2(u32_as_b2(v));
  else return __half22float2(u32_as_h2(v));
}
template <typename T>
__device__ __forceinline__ float elem_to_float(T v) {
  if constexpr (TypeTraits<T>::kIsBf16) return __bfloat162float(v);
  else return __half2float(v);
}
// reduction.cuh:72-82: finite clamp for fp16 only
template <typename T>
__device__ __forceinline__ T norm_round(float v) {
  if constexpr (TypeTraits<T>::kIsBf16) {
    return __float2bfloat16_rn(v);
  } else {
    constexpr float kLim = 65504.f - 1000.f;
    return v > 0.f ? __float2half_rn(fminf(v, kLim)) : __float2half_rn(fmaxf(v, -kLim));
  }
}
template <typename T>
__device__ __forceinline__ uint32_t norm_pair(uint32_t x2, uint32_t g2, float r) {
  const float2 x = pair_to_float2<T>(x2), g = pair_to_float2<T>(g2);
  const T lo = norm_round<T>((x.x * r) * g.x), hi = norm_round<T>((x.y * r) * g.y);
  return (uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) << 16);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += red[w];  // every thread, same fixed order: deterministic, no second barrier
  return t;
}

// NCH 16-byte chunks per thread live in registers.
template <typename T, int NCH>
__global__ void __launch_bounds__(kNormThreads) rmsnorm_rows_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                                    T* __restrict__ y, float eps, int n) {
  __shared__ float red[kNormThreads / 32];
  const int nchunk = n >> 3, tid = threadIdx.x;
  const uint4* gv = reinterpret_cast<const uint4*>(gamma);
  uint4 g[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {  // gamma is a constant of the model: fetched before the dependency wait
    const int c = tid + j * kNormThreads;
    g[j] = c < nchunk ? __ldg(gv + c) : make_uint4(0, 0, 0, 0);
  }
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const uint4* xv = reinterpret_cast<const uint4*>(x + (size_t)blockIdx.x * n);
  uint4 v[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c = tid + j * kNormThreads;
    v[j] = c < nchunk ? xv[c] : make_uint4(0, 0, 0, 0);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 p = pair_to_float2<T>(w[e]);
      s = fmaf(p.x, p.x, s);
      s = fmaf(p.y, p.y, s);
    }
  }
  const float r = rsqrtf(block_sum(s, red) / (float)n + eps);
  uint4* yv = reinterpret_cast<uint4*>(y + (size_t)blockIdx.x * n);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c = tid + j * kNormThreads;
    if (c < nchunk)
      yv[c] = make_uint4(norm_pair<T>(v[j].x, g[j].x, r), norm_pair<T>(v[j].y, g[j].y, r), norm_pair<T>(v[j].z, g[j].z, r),
                         norm_pair<T>(v[j].w, g[j].w, r));
  }
}

// any n: two passes over the row (the second one hits L1 / L2)
template <typename T>
__global__ void __launch_bounds__(kNormThreads) rmsnorm_generic_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                                       T* __restrict__ y, float eps, int n) {
  __shared__ float red[kNormThreads / 32];
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const T* xr = x + (size_t)blockIdx.x * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += kNormThreads) {
    const float f = elem_to_float<T>(xr[i]);
    s = fmaf(f, f, s);
  }
  const float r = rsqrtf(block_sum(s, red) / (float)n + eps);
  T* yr = y + (size_t)blockIdx.x * n;
  for (int i = threadIdx.x; i < n; i += kNormThreads)
    yr[i] = norm_round<T>((elem_to_float<T>(xr[i]) * r) * elem_to_float<T>(gamma[i]));
}

template <typename T>
int launch_t(const void* x, const void* gamma, void* y, float eps, int m, int n, bool pdl, cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)m);
  cfg.blockDim = dim3(kNormThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const T* xp = static_cast<const T*>(x);
  const T* gp = static_cast<const T*>(gamma);
  T* yp = static_cast<T*>(y);
  const bool vec = n % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
  const int per = (n / 8 + kNormThreads - 1) / kNormThreads;
  cudaError_t e;
  if (vec && per <= 1) e = cudaLaunchKernelEx(&cfg, rmsnorm_rows_kernel<T, 1>, xp, gp, yp, eps, n);
  else if (vec && per <= 2) e = cudaLaunchKernelEx(&cfg, rmsnorm_rows_kernel<T, 2>, xp, gp, yp, eps, n);
  else if (vec && per <= 4) e = cudaLaunchKernelEx(&cfg, rmsnorm_rows_kernel<T, 4>, xp, gp, yp, eps, n);
  else if (vec && per <= 8) e = cudaLaunchKernelEx(&cfg, rmsnorm_rows_kernel<T, 8>, xp, gp, yp, eps, n);
  else e = cudaLaunchKernelEx(&cfg, rmsnorm_generic_kernel<T>, xp, gp, yp, eps, n);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

int launch_rmsnorm(const void* x, const void* gamma, void* y, float eps, int m, int n, int dtype, bool pdl, cudaStream_t stream) {
  if (dtype == B200AWQ_DTYPE_F16) return launch_t<__half>(x, gamma, y, eps, m, n, pdl, stream);
  return launch_t<__nv_bfloat16>(x, gamma, y, eps, m, n, pdl, stream);
}

}  // namespace b200awq
