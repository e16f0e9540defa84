// How fast can one B200 stream a large read-only buffer from HBM through each load path, as a function of the
// bytes in flight per SM?  (The decode GEMV needs ~44 B/ns per SM = 6.5 TB/s / 148.)
//   A. cp.async.bulk (UBLKCP) global -> shared, P issuing threads per CTA, ring of S slots of B bytes, no consumer
//      (the issuer waits for a slot's previous copy to land, then re-issues into it)
//   B. cp.async 16 B (LDGSTS) from T threads, D commit groups in flight per thread, U copies per group
//   C. ld.global.nc.v4 (LDG.128) from T threads, U independent loads per iteration, xor-reduced
//   D. A + C together in one CTA (does the bulk-copy path add to the LDG path?)
// Persistent CTAs (148 x ctas_per_sm), each CTA walks a disjoint contiguous slab; total 2 GB per run (>> L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o hbm_probe.bin hbm_probe.cu && ./hbm_probe.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- A: bulk copies
// P issuing threads (one per warp), each with its own S slots of B bytes.  The CTA's slab is dealt to the issuers
// chunk by chunk (chunk i -> issuer i % P) so the CTA reads its slab sequentially.
__global__ void bulk_kernel(const uint8_t* src, size_t slab, int P, int S, int B, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [P][S]
  uint8_t* ring = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P * S; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane != 0 || warp >= P) return;
  const uint8_t* base = src + (size_t)blockIdx.x * slab;
  const size_t nchunk = slab / B;
  int slot = 0, use = 0;
  for (size_t c = warp; c < nchunk; c += P) {
    uint64_t* bar = &bars[warp * S + slot];
    if (use > 0) mbar_wait(bar, (uint32_t)((use - 1) & 1));
    mbar_expect_tx(bar, (uint32_t)B);
    bulk_g2s(smem_u32(ring + (size_t)(warp * S + slot) * B), base + c * B, (uint32_t)B, bar);
    if (++slot == S) slot = 0, ++use;
  }
  for (int s = 0; s < S; ++s) {  // drain
    const int u = use + (s < slot ? 1 : 0);
    if (u > 0) mbar_wait(&bars[warp * S + s], (uint32_t)((u - 1) & 1));
  }
  if (sink && base[0] == 255 && base[1] == 254 && base[77] == 3) sink[0] = 1;
}

// ---------------------------------------------------------------- B: cp.async 16 B
template <int U, int D>
__global__ void cpasync_kernel(const uint8_t* src, size_t slab, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int T = blockDim.x;
  const uint8_t* base = src + (size_t)blockIdx.x * slab;
  const size_t step = (size_t)T * 16 * U;  // bytes per group per CTA
  const size_t ngroups = slab / step;
  const uint32_t sbase = smem_u32(smem) + threadIdx.x * 16;
  for (size_t g = 0; g < ngroups + D; ++g) {
    if (g < ngroups) {
      const uint8_t* p = base + g * step + (size_t)threadIdx.x * 16;
      const uint32_t d = sbase + (uint32_t)((g % D) * step);
#pragma unroll
      for (int u = 0; u < U; ++u)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + u * T * 16), "l"(p + (size_t)u * T * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
  }
  if (sink && smem[threadIdx.x] == 255 && smem[threadIdx.x + 1] == 254 && smem[77] == 3) sink[0] = 1;
}

// ---------------------------------------------------------------- C: LDG.128
template <int U>
__global__ void ldg_kernel(const uint8_t* src, size_t slab, unsigned long long* sink) {
  const int T = blockDim.x;
  const uint4* base = reinterpret_cast<const uint4*>(src + (size_t)blockIdx.x * slab);
  const size_t n16 = slab / 16, step = (size_t)T * U;
  uint32_t acc = 0;
  for (size_t i = threadIdx.x; i + (size_t)(U - 1) * T < n16; i += step) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                   : "l"(base + i + (size_t)u * T));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (sink && acc == 0x12345678u) sink[0] = acc;
}

// ---------------------------------------------------------------- D: bulk + LDG in one CTA: warps [0, P) issue bulk copies
// over the first `frac_bulk`/16 of the slab, the remaining warps LDG the rest
template <int U>
__global__ void mixed_kernel(const uint8_t* src, size_t slab, int P, int S, int B, int bulk16, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < P * S; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t slab_a = (slab / 16 * bulk16) / B * B;
  const uint8_t* base = src + (size_t)blockIdx.x * slab;
  if (warp < P) {
    if (lane != 0) return;
    const size_t nchunk = slab_a / B;
    int slot = 0, use = 0;
    for (size_t c = warp; c < nchunk; c += P) {
      uint64_t* bar = &bars[warp * S + slot];
      if (use > 0) mbar_wait(bar, (uint32_t)((use - 1) & 1));
      mbar_expect_tx(bar, (uint32_t)B);
      bulk_g2s(smem_u32(ring + (size_t)(warp * S + slot) * B), base + c * B, (uint32_t)B, bar);
      if (++slot == S) slot = 0, ++use;
    }
    for (int s = 0; s < S; ++s) {
      const int u = use + (s < slot ? 1 : 0);
      if (u > 0) mbar_wait(&bars[warp * S + s], (uint32_t)((u - 1) & 1));
    }
    return;
  }
  const int T = blockDim.x - P * 32, t = threadIdx.x - P * 32;
  const uint4* b4 = reinterpret_cast<const uint4*>(base + slab_a);
  const size_t n16 = (slab - slab_a) / 16, step = (size_t)T * U;
  uint32_t acc = 0;
  for (size_t i = t; i + (size_t)(U - 1) * T < n16; i += step) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                   : "l"(b4 + i + (size_t)u * T));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (sink && acc == 0x12345678u) sink[0] = acc;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) {
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int main() {
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  const size_t total = (size_t)2 << 30;
  uint8_t* buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("SMs %d, buffer %.1f GB\n", nsm, total / 1e9);
  auto report = [&](const char* what, size_t bytes, float ms, size_t inflight_sm) {
    printf("%-64s %7.0f GB/s   (%5.1f B/ns/SM, %6.1f KB in flight / SM)\n", what, bytes / ms / 1e6, bytes / ms / 1e6 / nsm,
           inflight_sm / 1024.0);
    fflush(stdout);
  };
  char name[160];

  // ---------------- A
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cfgA[][4] = {  // ctas/SM, P, S, B
      {1, 1, 5, 4096},  {1, 1, 5, 16384}, {1, 1, 10, 16384}, {1, 2, 5, 16384}, {1, 4, 3, 16384}, {1, 4, 6, 8192},
      {1, 8, 3, 8192},  {1, 8, 6, 4096},  {1, 16, 3, 4096},  {1, 1, 6, 32768}, {1, 2, 3, 32768}, {2, 1, 5, 16384},
      {2, 2, 3, 16384}, {2, 4, 3, 8192},  {2, 8, 3, 4096},   {4, 2, 3, 8192},  {4, 4, 3, 4096},  {1, 12, 2, 8192},
      {1, 24, 2, 4096}, {2, 12, 2, 4096}, {1, 6, 2, 16384},  {1, 3, 2, 32768}, {1, 1, 2, 98304 - 2048}};
  for (auto& c : cfgA) {
    const int cps = c[0], P = c[1], S = c[2], B = c[3] / 16 * 16;
    const int grid = nsm * cps;
    const size_t slab = total / grid / ((size_t)B * P) * ((size_t)B * P);
    const size_t smem = 1024 + (size_t)P * S * B;
    if (smem * cps > 220 * 1024) continue;
    bulk_kernel<<<grid, P * 32, smem>>>(buf, slab, P, S, B, nullptr);
    CK(cudaEventRecord(e0));
    bulk_kernel<<<grid, P * 32, smem>>>(buf, slab, P, S, B, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    snprintf(name, sizeof name, "A bulk   ctas/SM %d  issuers %2d  slots %2d x %6d B", cps, P, S, B);
    report(name, slab * grid, time_ms(e0, e1), (size_t)cps * P * S * B);
  }

  // ---------------- B
  auto runB = [&](auto kern, int U, int D, int cps, int T) {
    const int grid = nsm * cps;
    const size_t step = (size_t)T * 16 * U;
    const size_t slab = total / grid / step * step;
    const size_t smem = step * D;
    if (smem * cps > 220 * 1024) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, T, smem>>>(buf, slab, nullptr);
    CK(cudaEventRecord(e0));
    kern<<<grid, T, smem>>>(buf, slab, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    snprintf(name, sizeof name, "B cp.async16  ctas/SM %d  threads %4d  groups in flight %d x %d copies", cps, T, D, U);
    report(name, slab * grid, time_ms(e0, e1), (size_t)cps * smem);
  };
  runB(cpasync_kernel<1, 4>, 1, 4, 1, 1024);
  runB(cpasync_kernel<2, 4>, 2, 4, 1, 1024);
  runB(cpasync_kernel<4, 3>, 4, 3, 1, 1024);
  runB(cpasync_kernel<4, 3>, 4, 3, 1, 512);
  runB(cpasync_kernel<4, 4>, 4, 4, 1, 256);
  runB(cpasync_kernel<8, 4>, 8, 4, 1, 128);
  runB(cpasync_kernel<8, 4>, 8, 4, 1, 64);
  runB(cpasync_kernel<8, 4>, 8, 4, 1, 32);
  runB(cpasync_kernel<4, 3>, 4, 3, 2, 512);
  runB(cpasync_kernel<2, 4>, 2, 4, 2, 512);
  runB(cpasync_kernel<4, 4>, 4, 4, 2, 256);
  runB(cpasync_kernel<4, 4>, 4, 4, 4, 128);

  // ---------------- C
  auto runC = [&](auto kern, int U, int cps, int T) {
    const int grid = nsm * cps;
    const size_t step = (size_t)T * 16 * U;
    const size_t slab = total / grid / step * step;
    kern<<<grid, T>>>(buf, slab, nullptr);
    CK(cudaEventRecord(e0));
    kern<<<grid, T>>>(buf, slab, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    snprintf(name, sizeof name, "C LDG.128  ctas/SM %d  threads %4d  loads per thread in flight %d", cps, T, U);
    report(name, slab * grid, time_ms(e0, e1), (size_t)cps * step);
  };
  runC(ldg_kernel<1>, 1, 1, 1024);
  runC(ldg_kernel<2>, 2, 1, 1024);
  runC(ldg_kernel<4>, 4, 1, 1024);
  runC(ldg_kernel<8>, 8, 1, 1024);
  runC(ldg_kernel<4>, 4, 1, 512);
  runC(ldg_kernel<8>, 8, 1, 512);
  runC(ldg_kernel<8>, 8, 1, 256);
  runC(ldg_kernel<16>, 16, 1, 256);
  runC(ldg_kernel<16>, 16, 1, 128);
  runC(ldg_kernel<4>, 4, 2, 1024);
  runC(ldg_kernel<8>, 8, 2, 512);
  runC(ldg_kernel<8>, 8, 4, 256);

  // ---------------- D
  CK(cudaFuncSetAttribute(mixed_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cfgD[][6] = {  // ctas/SM, P, S, B, LDG threads, sixteenths of the slab through bulk copies
      {1, 2, 4, 16384, 256, 8}, {1, 2, 4, 16384, 512, 8}, {1, 4, 4, 8192, 512, 8}, {1, 2, 4, 16384, 512, 5},
      {1, 2, 4, 16384, 512, 11}, {2, 2, 3, 8192, 256, 8}};
  for (auto& c : cfgD) {
    const int cps = c[0], P = c[1], S = c[2], B = c[3], TL = c[4], f16 = c[5];
    const int grid = nsm * cps;
    const size_t unit = (size_t)B * P * 16 * 64;
    const size_t slab = total / grid / unit * unit;
    const size_t smem = 1024 + (size_t)P * S * B;
    mixed_kernel<8><<<grid, P * 32 + TL, smem>>>(buf, slab, P, S, B, f16, nullptr);
    CK(cudaEventRecord(e0));
    mixed_kernel<8><<<grid, P * 32 + TL, smem>>>(buf, slab, P, S, B, f16, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    snprintf(name, sizeof name, "D mixed  ctas/SM %d  bulk %d x %d x %5d B (%2d/16)  + LDG %3d thr x 8", cps, P, S, B, f16, TL);
    report(name, slab * grid, time_ms(e0, e1), (size_t)cps * (P * S * B + TL * 16 * 8));
  }
  return 0;
}
