// Standalone sm_100a micro-probe: what the fixed costs of a tcgen05 streaming pipeline are, per instruction.
// Not part of the product; built by scripts/probes/build.sh into scripts/probes/tc_probe.bin and run on the GPU box:
//     scripts/probes/tc_probe.bin            (prints one table, a few milliseconds of GPU time)
// Every number is clock64 cycles on one SM, one CTA, nothing else running.
//   A. tcgen05.mma (TS form: A in TMEM, B = 128-byte-swizzled K-major tile in shared memory, M = 128, K = 16),
//      R back-to-back instructions from ONE thread: issue cycles per MMA, and cycles per MMA until the commit lands,
//      for N = 16 .. 256.  (Sizes the MMA thread of a decode kernel: 55 MMAs per SM per 4096x4096 GEMV.)
//   B. mbarrier: try_wait on an already completed phase; arrive -> wake-up of a waiting warp; tcgen05.commit with
//      nothing pending -> mbarrier complete.
//   C. tcgen05.st.32x32b.x32 + wait::st (one warp, then four warps at once); tcgen05.ld.32x32b.x32 + wait::ld.
//   D. cp.async.bulk (global -> shared) of 4 / 8 / 16 / 32 KB from one thread: issue cycles per copy, and latency
//      until complete_tx lands (data resident in L2).
//   E. tcgen05.mma kind::i8, TS form (u8 A in TMEM x s8 B in shared memory -> s32, K = 32 per instruction): a known-answer
//      check of the operand layouts (A: TMEM lane = row, 32-bit column c = bytes k = 4c .. 4c+3, little endian; B: the
//      same K-major 128-byte-swizzled tile as for f16, one byte per element) and the issue cost per instruction.
//   F. two warps issuing kind::f16 MMAs at the same time: is the ~70-cycle issue cost per thread or per SM?
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  while (!bar_try(b, parity)) {
  }
}
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
#define ST32(taddr, r)                                                                                                   \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"  \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                              \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),      \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),        \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),        \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                     \
      : "memory")
#define LD32(taddr, r)                                                                                                   \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"    \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
      : "r"(taddr)                                                                                                       \
      : "memory")

constexpr int kSlots = 64;  // result slots (long long each)

// out[0..] layout documented in main()
__global__ void __launch_bounds__(256, 1) probe_kernel(long long* out, const uint8_t* gsrc, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* btile = smem;                                        // 256 rows x 128 B (zeros): B operand, 32 KB
  uint8_t* dst = smem + 32768;                                  // bulk-copy landing zone, 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);   // [0] commit, [1] ping, [2] pong, [3] done-phase, [4] copy
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(btile)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) bar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tmem = *slot;
  const uint32_t d_tmem = tmem, a_tmem = tmem + 256;  // D: columns [0,256), A: [256, 512)

  // zero the A operand columns (32 x 8 = 256 columns) from warps 4..7 (lanes quarter q)
  if (warp >= 4) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int c = 0; c < 8; ++c) ST32(a_tmem + lane_base + c * 32, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_before();
  __syncthreads();
  tc_after();

  uint32_t cph = 0;  // phase of bars[0]
  // ---------------------------------------------------------------- A. MMA issue / throughput for N = 16 << i
  if (threadIdx.x == 0) {
    for (int ni = 0; ni < 5; ++ni) {
      const int N = 16 << ni;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
      const uint64_t bdesc = sw128_desc(s32(btile));
      // warm-up
      mma_ts(d_tmem, a_tmem, bdesc, idesc, 0);
      tc_commit(&bars[0]);
      bar_wait(&bars[0], cph), cph ^= 1;
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) mma_ts(d_tmem, a_tmem + (r & 7) * 8, bdesc + (uint64_t)((r & 3) * 2), idesc, 1);
      const long long t1 = clock64();
      tc_commit(&bars[0]);
      bar_wait(&bars[0], cph), cph ^= 1;
      const long long t2 = clock64();
      out[ni * 2 + 0] = t1 - t0;
      out[ni * 2 + 1] = t2 - t0;
    }
    // commit with nothing pending
    const long long t0 = clock64();
    tc_commit(&bars[0]);
    bar_wait(&bars[0], cph), cph ^= 1;
    out[10] = clock64() - t0;
    // try_wait on a completed phase (the phase that just completed has parity cph ^ 1)
    const long long t3 = clock64();
    for (int r = 0; r < 16; ++r) (void)bar_try(&bars[0], cph ^ 1);
    out[11] = (clock64() - t3) / 16;
  }
  __syncthreads();
  // ---------------------------------------------------------------- B. arrive -> wake-up ping-pong between warps 1 and 2
  if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < 32; ++r) {
      bar_arrive(&bars[1]);
      bar_wait(&bars[2], ph);
      ph ^= 1;
    }
    out[12] = (clock64() - t0) / 64;  // one hop
  } else if (warp == 2 && lane == 0) {
    uint32_t ph = 0;
    for (int r = 0; r < 32; ++r) {
      bar_wait(&bars[1], ph);
      ph ^= 1;
      bar_arrive(&bars[2]);
    }
  }
  __syncthreads();
  // ---------------------------------------------------------------- C. tcgen05.st / ld round trips
  if (warp >= 4) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    if (warp == 4) {  // one warp alone
      const long long t0 = clock64();
      for (int i = 0; i < 8; ++i) {
        ST32(a_tmem + lane_base + (i & 7) * 32, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      if (lane == 0) out[13] = (clock64() - t0) / 8;
      const long long t1 = clock64();
      for (int i = 0; i < 8; ++i) {
        LD32(a_tmem + lane_base + (i & 7) * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if (lane == 0) out[14] = (clock64() - t1) / 8;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const long long t0 = clock64();  // four warps at once
    for (int i = 0; i < 8; ++i) {
      ST32(a_tmem + lane_base + (i & 7) * 32, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 128) out[15] = (clock64() - t0) / 8;
    // 8 stores back to back, one wait
    const long long t1 = clock64();
    for (int i = 0; i < 8; ++i) ST32(a_tmem + lane_base + (i & 7) * 32, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (threadIdx.x == 128) out[16] = (clock64() - t1) / 8;
    if (r[0] == 0x12345) out[63] = r[1];  // keep r alive
  }
  tc_before();
  __syncthreads();
  tc_after();
  // ---------------------------------------------------------------- D. bulk copies global -> shared, one issuing thread
  if (threadIdx.x == 0) {
    uint32_t ph = 0;
    for (int si = 0; si < 4; ++si) {
      const uint32_t bytes = 4096u << si;
      const int n = 32768 / (int)bytes;  // copies that fill the 32 KB landing zone
      // warm the lines into L2 first
      bar_expect(&bars[4], 32768);
      for (int c = 0; c < n; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(dst + c * bytes)),
                     "l"(gsrc + (size_t)si * 32768 + (size_t)c * bytes), "r"(bytes), "r"(s32(&bars[4]))
                     : "memory");
      bar_wait(&bars[4], ph), ph ^= 1;
      bar_expect(&bars[4], 32768);
      const long long t0 = clock64();
      for (int c = 0; c < n; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(dst + c * bytes)),
                     "l"(gsrc + (size_t)si * 32768 + (size_t)c * bytes), "r"(bytes), "r"(s32(&bars[4]))
                     : "memory");
      const long long t1 = clock64();
      bar_wait(&bars[4], ph), ph ^= 1;
      const long long t2 = clock64();
      out[20 + si * 2] = (t1 - t0) / n;  // issue cycles per copy
      out[21 + si * 2] = t2 - t0;        // 32 KB landed (from L2)
    }
  }
  tc_before();
  __syncthreads();
  if (warp == 0) {
    tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__host__ __device__ inline int probe_a(int r, int k) { return (r + 3 * k) & 15; }          // u8 operand (a 4-bit weight)
__host__ __device__ inline int probe_b(int n, int k) { return ((n * 7 + k * 3) % 11) - 5; }  // s8 operand

// out2[0] = mismatches of D (128 x 16, K = 32) against the integer dot products, out2[1..16] = D[lane 1][0..15] as computed,
// out2[20] / out2[21] = issue cycles per i8 MMA (N = 16 / N = 128), out2[24] = cycles per f16 MMA (N = 16) when TWO
// warps issue at once (compare with section A), out2[25] the same for the second warp.
__global__ void __launch_bounds__(256, 1) probe2_kernel(long long* out2, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* btile = smem;  // 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(btile)[i] = 0;
  __syncthreads();
  // B[n][k], n < 16, k < 32: K-major rows of 128 B, 16-byte chunk index XOR (row & 7)  (SWIZZLE_128B)
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int n = i >> 5, k = i & 31;
    btile[n * 128 + (((k >> 4) ^ (n & 7)) << 4) + (k & 15)] = (uint8_t)(int8_t)probe_b(n, k);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) bar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tmem = *slot;
  const uint32_t d_tmem = tmem, a_tmem = tmem + 256, d2_tmem = tmem + 128;
  if (warp >= 4) {  // A[r][k]: lane r, column c = bytes 4c .. 4c+3
    const int r = (warp & 3) * 32 + lane;
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      uint32_t w = 0;
      if (c < 8)
        for (int b = 0; b < 4; ++b) w |= (uint32_t)probe_a(r, 4 * c + b) << (8 * b);
      v[c] = w;
    }
    ST32(a_tmem + ((uint32_t)((warp & 3) * 32) << 16), v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_before();
  __syncthreads();
  tc_after();
  uint32_t cph = 0;
  if (threadIdx.x == 0) {
    // u8 x s8 -> s32, K-major both, M = 128
    const uint32_t idesc16 = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = sw128_desc(s32(btile));
    mma_i8_ts(d_tmem, a_tmem, bdesc, idesc16, 0);
    tc_commit(&bars[0]);
    bar_wait(&bars[0], cph), cph ^= 1;
  }
  tc_before();
  __syncthreads();
  tc_after();
  if (warp >= 4) {  // check D[r][n]
    const int r = (warp & 3) * 32 + lane;
    uint32_t v[32];
    LD32(d_tmem + ((uint32_t)((warp & 3) * 32) << 16), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int bad = 0;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      int ref = 0;
      for (int k = 0; k < 32; ++k) ref += probe_a(r, k) * probe_b(n, k);
      bad += ((int)v[n] != ref);
      if (r == 1) out2[1 + n] = (int)v[n];
    }
    if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(&out2[0]), (unsigned long long)bad);
  }
  tc_before();
  __syncthreads();
  tc_after();
  if (threadIdx.x == 0) {
    const uint64_t bdesc = sw128_desc(s32(btile));
    for (int t = 0; t < 2; ++t) {
      const int N = t ? 128 : 16;
      const uint32_t idesc = (2u << 4) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) mma_i8_ts(d_tmem, a_tmem, bdesc + (uint64_t)((r & 3) * 2), idesc, 1);
      const long long t1 = clock64();
      tc_commit(&bars[0]);
      bar_wait(&bars[0], cph), cph ^= 1;
      out2[20 + t] = t1 - t0;
      out2[22 + t] = clock64() - t0;
    }
  }
  __syncthreads();
  // F: warps 1 and 2 issue f16 MMAs (N = 16) into different accumulators at the same time
  if ((warp == 1 || warp == 2) && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = sw128_desc(s32(btile));
    const uint32_t d = warp == 1 ? d_tmem : d2_tmem;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) mma_ts(d, a_tmem + (r & 7) * 8, bdesc + (uint64_t)((r & 3) * 2), idesc, 1);
    const long long t1 = clock64();
    tc_commit(&bars[warp]);
    bar_wait(&bars[warp], 0);
    out2[24 + (warp - 1)] = t1 - t0;
  }
  tc_before();
  __syncthreads();
  if (warp == 0) {
    tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? std::atoi(argv[1]) : 64;
  long long* out = nullptr;
  uint8_t* src = nullptr;
  CK(cudaMalloc(&out, kSlots * sizeof(long long)));
  CK(cudaMemset(out, 0, kSlots * sizeof(long long)));
  CK(cudaMalloc(&src, 4 * 32768));
  CK(cudaMemset(src, 1, 4 * 32768));
  const int smem = 65536 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int it = 0; it < 2; ++it) {  // second pass: instruction cache warm
    probe_kernel<<<1, 256, smem>>>(out, src, reps);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(kSlots);
  CK(cudaMemcpy(h.data(), out, kSlots * sizeof(long long), cudaMemcpyDeviceToHost));
  std::printf("A. tcgen05.mma TS, M=128 K=16, %d back-to-back from one thread (cycles per MMA)\n", reps);
  std::printf("   %6s %12s %18s %14s\n", "N", "issue", "issue..complete", "ideal (N/2)");
  for (int ni = 0; ni < 5; ++ni)
    std::printf("   %6d %12.1f %18.1f %14d\n", 16 << ni, (double)h[ni * 2] / reps, (double)h[ni * 2 + 1] / reps, (16 << ni) / 2);
  std::printf("B. tcgen05.commit with nothing pending -> wait returns: %lld cycles\n", h[10]);
  std::printf("   mbarrier.try_wait on a completed phase:              %lld cycles\n", h[11]);
  std::printf("   mbarrier arrive -> other warp wakes (one hop):       %lld cycles\n", h[12]);
  std::printf("C. tcgen05.st.x32 + wait::st, one warp:                 %lld cycles\n", h[13]);
  std::printf("   tcgen05.ld.x32 + wait::ld, one warp:                 %lld cycles\n", h[14]);
  std::printf("   tcgen05.st.x32 + wait::st, four warps at once:       %lld cycles\n", h[15]);
  std::printf("   tcgen05.st.x32 x8 then one wait, four warps (per st): %lld cycles\n", h[16]);
  std::printf("D. cp.async.bulk global(L2) -> shared from one thread, 32 KB in total\n");
  std::printf("   %8s %18s %22s\n", "bytes", "issue per copy", "32 KB landed (cycles)");
  for (int si = 0; si < 4; ++si) std::printf("   %8d %18lld %22lld\n", 4096 << si, h[20 + si * 2], h[21 + si * 2]);

  CK(cudaMemset(out, 0, kSlots * sizeof(long long)));
  const int smem2 = 16384 + 1024 + 256;
  for (int it = 0; it < 2; ++it) {
    if (it) CK(cudaMemset(out, 0, sizeof(long long)));
    probe2_kernel<<<1, 256, smem2>>>(out, reps);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h.data(), out, kSlots * sizeof(long long), cudaMemcpyDeviceToHost));
  std::printf("E. tcgen05.mma kind::i8 TS (u8 A in TMEM x s8 B in smem -> s32), M=128 N=16 K=32: %lld of 2048 outputs differ\n", h[0]);
  std::printf("   D[row 1][0..15] computed:");
  for (int n = 0; n < 16; ++n) std::printf(" %lld", h[1 + n]);
  std::printf("\n   D[row 1][0..15] expected:");
  for (int n = 0; n < 16; ++n) {
    int ref = 0;
    for (int k = 0; k < 32; ++k) ref += probe_a(1, k) * probe_b(n, k);
    std::printf(" %d", ref);
  }
  std::printf("\n   issue cycles per i8 MMA: N=16 %.1f (to complete %.1f), N=128 %.1f (to complete %.1f)\n", (double)h[20] / reps,
              (double)h[22] / reps, (double)h[21] / reps, (double)h[23] / reps);
  std::printf("F. two warps issuing f16 MMAs (N=16) at once: %.1f and %.1f cycles per MMA each\n", (double)h[24] / reps,
              (double)h[25] / reps);
  return 0;
}
