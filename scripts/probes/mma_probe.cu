// Throughput / latency of the warp-level (legacy) tensor-core MACs on one SM of a B200, and of the ALU ops that
// feed them in the decode kernel:  HMMA.16816.F32 (f16 x f16 + f32), IMMA.16832 (u8 x s8 + s32), LOP3, SHF, PRMT,
// I2F, LDSM.  One CTA, W warps (W / 4 per scheduler); every warp runs R rounds of 8 independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe.bin mma_probe.cu && ./mma_probe.bin
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define HMMA(d, a, b)                                                                                              \
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])                                                    \
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]))
#define IMMA(d, a, b)                                                                                              \
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"   \
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])                                                    \
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]))

template <int KIND, int CHAINS>
__global__ void probe(long long* out, uint32_t seed, int rounds) {
  uint32_t a[4] = {seed, seed * 3, seed * 5, seed * 7}, b[2] = {seed * 11, seed * 13};
  float f[8][4];
  int q[8][4];
  uint32_t v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = seed + c * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) f[c][i] = 0.f, q[c][i] = 0;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (KIND == 0) HMMA(f[c], a, b);
      if (KIND == 1) IMMA(q[c], a, b);
      if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0xea;" : "+r"(v[c]) : "r"(a[0]), "r"(b[0]));
      if (KIND == 3) asm volatile("shr.u32 %0, %0, 1;" : "+r"(v[c]));
      if (KIND == 4) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(v[c]) : "r"(a[1]));
      if (KIND == 5) asm volatile("{ .reg .f32 t; cvt.rn.f32.s32 t, %0; mov.b32 %0, t; }" : "+r"(v[c]));
      if (KIND == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(a[1]), "r"(b[1]));
      if (KIND == 7) asm volatile("{ .reg .b32 t; mul.hi.u32 %0, %0, %1; }" : "+r"(v[c]) : "r"(a[1]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  int z = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    z += v[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += f[c][i], z += q[c][i];
  }
  if (threadIdx.x % 32 == 0) out[threadIdx.x / 32] = t1 - t0;
  out[64 + threadIdx.x] = (long long)z + (long long)s;  // keeps every chain live (ptxas removes dead inline PTX)
}

template <int KIND, int CHAINS>
void run(const char* name, long long* d) {
  const int rounds = 512;
  for (int warps : {4, 8, 16, 32}) {
    probe<KIND, CHAINS><<<1, warps * 32>>>(d, 12345u, rounds);
    probe<KIND, CHAINS><<<1, warps * 32>>>(d, 12345u, rounds);
    long long h[32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per_warp = (double)mx / (rounds * CHAINS);
    printf("  %-34s chains %d  warps/SM %2d: %7.2f cycles per instr per warp, %6.2f per instr per scheduler\n", name, CHAINS,
           warps, per_warp, per_warp / (warps / 4));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16384);
  printf("independent chains (throughput)\n");
  run<0, 8>("HMMA.16816.F32", d);
  run<1, 8>("IMMA.16832.U8.S8", d);
  run<2, 8>("LOP3", d);
  run<3, 8>("SHF (shr)", d);
  run<4, 8>("PRMT", d);
  run<5, 8>("I2F", d);
  run<6, 8>("IMAD (mad.lo)", d);
  run<7, 8>("IMAD.HI (mul.hi)", d);
  printf("one dependent chain (latency)\n");
  run<0, 1>("HMMA.16816.F32", d);
  run<1, 1>("IMMA.16832.U8.S8", d);
  run<2, 1>("LOP3", d);
  run<5, 1>("I2F", d);
  run<7, 1>("IMAD.HI (mul.hi)", d);
  printf("two dependent chains\n");
  run<0, 2>("HMMA.16816.F32", d);
  run<1, 2>("IMMA.16832.U8.S8", d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
