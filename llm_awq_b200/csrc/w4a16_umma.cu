// Tensor-core W4A16 kernel for sm_100a: TMA -> shared memory -> register dequant ->
// tensor memory -> tcgen05.mma with the fp32 accumulator in TMEM.
//
// Replaces the reference's gemm_w4a16_T1 / gemm_w4a16_T2 (quantization_new/gemm/
// gemm_cuda.cu:312-752, :922-1124; Ampere cp.async + ldmatrix + mma.sync).
//
// Orientation: the tile is computed TRANSPOSED, D[oc, tok] = W~[oc, k] . X[tok, k]^T.
//   * UMMA M = 128 output channels.  The dequantised weight is the A operand and lives
//     in TENSOR MEMORY (tcgen05.mma "TS" form): TMEM lane = output channel, one 32-bit
//     column = one k-adjacent fp16/bf16 pair.  That is exactly what one lane produces
//     from one packed 32-bit word (see w4_common.cuh), so a dequant thread owns one
//     output channel, reads its 2 x 16 B of packed nibbles per 64-k block from shared
//     memory, applies scale/zero with one exact FMA per pair and writes 32 columns with
//     a single tcgen05.st.  The dequantised weights never touch shared memory, which
//     keeps shared-memory bandwidth for the activation operand.
//   * UMMA N = TN tokens (32..256).  The activation tile X[tok, 64 k] arrives by TMA
//     with the 128-byte swizzle and is consumed in place as the K-major B operand.
//   * D (128 lanes x TN fp32 columns) stays in TMEM for the whole K loop.
// Warp roles (384 threads, persistent over output tiles):
//   warp 0 lane 0  TMA producer        (activation tile + packed-weight tile per stage)
//   warp 1 lane 0  tcgen05.mma issuer  (4 x K16 per stage, commits release the stage)
//   warp 2         TMEM allocation / release
//   warps 4-11     two groups of 4 (one warp per TMEM lane quarter each) that take alternate 64-k stages:
//                  dequantise -> TMEM; per tile each group converts alternate 32-token chunks of the epilogue
//                  (tcgen05.ld -> cvt -> smem transpose -> coalesced 16-byte stores).  With one group the kernel
//                  was bound by the dequantisation latency of a single warp per scheduler (~800 cycles per
//                  stage against 512 for the MMAs of a 256-token tile); see profiles/README.md.
// Small token counts: the k range is split over a thread-block cluster (2/4/8 CTAs share a tile) and the fp32
// partial tiles are reduced through distributed shared memory in the epilogue (mbarriers with cluster scope).
#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved at run time)

#include "w4_common.cuh"
#include "w4a16_kernels.h"

namespace b200awq {

#ifdef B200AWQ_TRACE
// Debug build only (scripts/umma_probe.py): bit 0 skip the dequant ALU work, bit 1 skip tcgen05.st, bit 2 skip the MMAs,
// bit 3 skip the activation TMA, bit 4 skip the weight TMA, bit 5 skip the epilogue.  Results are wrong by design; only the timing matters.
__device__ int g_umma_dbg;
extern "C" int b200awq_debug_set_umma(int v) { return (int)cudaMemcpyToSymbol(g_umma_dbg, &v, sizeof(int)); }
#define UMMA_DBG(bit) (dbg & (bit))
// per-stage clock64 stamps of CTA 0 (events: 0 producer past empty, 1 TMA issued, 2 dequant past full, 3 dequant arrives
// afull, 4 MMA past afull, 5 MMA committed, 6 tile's dfull seen by epilogue, 7 epilogue done)
constexpr int kUmmaTraceStages = 256;
__device__ long long g_umma_trace[8][kUmmaTraceStages];
extern "C" int b200awq_debug_read_umma_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_umma_trace, sizeof(long long) * 8 * kUmmaTraceStages);
}
#define UMMA_STAMP(ev, idx)                                                             \
  do {                                                                                  \
    if (blockIdx.x == 0 && (idx) < kUmmaTraceStages) g_umma_trace[ev][idx] = clock64(); \
  } while (0)
#else
#define UMMA_DBG(bit) false
#define UMMA_STAMP(ev, idx) \
  do {                      \
  } while (0)
#endif

constexpr int kBM = 128;      // output channels per tile (TMEM lanes)
constexpr int kBK = 64;       // k per pipeline stage
// The dequantised-A ring in TMEM has one 32-column slot PER shared-memory stage (slot index == stage index), so one
// tcgen05.commit per stage frees both and the per-stage hand-offs are: TMA -> full, dequant -> afull, MMA -> empty.
constexpr int kUmmaThreads = 384;  // 4 service warps + 2 groups of 4 dequantisation / epilogue warps
constexpr int kWBytes = kBM * kBK / 2;  // 4096: packed weights per stage

// ---------------------------------------------------------------- tcgen05 / TMA PTX
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart (what TMA SWIZZLE_128B writes)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
  d |= (uint64_t)0 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

#define B200AWQ_TMEM_ST32(taddr, r)                                                                                      \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"  \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                              \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),      \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),        \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),        \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                     \
      : "memory")

#define B200AWQ_TMEM_LD32(taddr, r)                                                                                      \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"    \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
      : "r"(taddr)                                                                                                       \
      : "memory")

// ---------------------------------------------------------------- cluster-scope mbarrier helpers (split-K)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t rank) {
  const uint32_t a = map_to_rank(smem_u32(local_bar), rank);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

template <int TN>
struct UmmaSmem {
  static constexpr int kXBytes = TN * 128;  // TN token rows x 64 k x 2 B
  static constexpr int kStageBytes = kXBytes + kWBytes;
  static constexpr int kOutRow = 256 + 16;                // 128 channels x 2 B (+ pad)
  static constexpr int kOutBytes = 2 * 32 * kOutRow;      // double-buffered 32-token chunks
  static constexpr int kStages = (200 * 1024 - kOutBytes - 1536) / kStageBytes >= 8
                                     ? 8
                                     : (200 * 1024 - kOutBytes - 1536) / kStageBytes;
  static constexpr int kPartBytes = TN <= 64 ? TN * kBM * 4 : 0;  // fp32 partial tile for split-K (TN <= 64 only)
  static constexpr int kTotal = kStages * kStageBytes + kOutBytes + 1024 /*align*/ + 512 /*barriers*/;
  static constexpr int kTotalSplit = kTotal + kPartBytes;
  static constexpr int kTmemCols = (TN + 32 * kStages) <= 64    ? 64
                                   : (TN + 32 * kStages) <= 128 ? 128
                                   : (TN + 32 * kStages) <= 256 ? 256
                                                                : 512;
  static_assert(TN + 32 * kStages <= 512, "accumulator + A ring must fit TMEM");
};

template <typename T, int TN>
__global__ void __launch_bounds__(kUmmaThreads, 1)
w4a16_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const T* __restrict__ scales, const T* __restrict__ szeros, T* __restrict__ y, int M, int N, int K, int S) {
  using L = UmmaSmem<TN>;
  constexpr int STAGES = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;                                   // [STAGES][TN * 128]   (1024-aligned each)
  uint8_t* ws = xs + STAGES * L::kXBytes;               // [STAGES][4096]
  uint8_t* outb = ws + STAGES * kWBytes;                // [2][32][kOutRow]
  uint64_t* bars = reinterpret_cast<uint64_t*>(outb + L::kOutBytes);
  uint64_t* full = bars;                  // [STAGES]  TMA -> dequant (and, transitively through afull, the MMA)
  uint64_t* empty = full + STAGES;        // [STAGES]  MMA commit -> TMA: smem stage AND TMEM A slot s are free
  uint64_t* afull = empty + STAGES;       // [STAGES]  dequant (4) -> MMA
  uint64_t* dfull = afull + STAGES;       // MMA commit -> epilogue
  uint64_t* dempty = dfull + 1;           // epilogue (4) -> MMA
  uint64_t* pfull = dempty + 1;           // split-K: every rank's partial tile is written   (4 S remote arrivals)
  uint64_t* pempty = pfull + 1;           // split-K: every rank has finished reading MY tile (4 S remote arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + 1);
  float* pbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [TN][128] fp32 (S > 1 only)

#ifdef B200AWQ_TRACE
  const int dbg = g_umma_dbg;
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok_tiles = (M + TN - 1) / TN;
  const int num_tiles = tok_tiles * (N / kBM);
  // split-K over a thread-block cluster (small token counts): rank r of S owns k blocks [kb0, kb0 + num_kb) of the
  // SAME tile; the cluster walks the tiles together.  (Activation multicast over CTA pairs was tried and removed:
  // -44 % L2 traffic but the per-stage pair handshake made it slower, profiles/README.md.)
  const int CS = S;
  const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const int num_kb = K / kBK / S;
  const int kb0 = crank * num_kb;
  const int num_units = num_tiles;
  const int tile0 = (CS > 1) ? (int)(blockIdx.x / CS) : (int)blockIdx.x;
  const int tile_step = (CS > 1) ? (int)(gridDim.x / CS) : (int)gridDim.x;
#define B200AWQ_UNIT_DECODE(unit, oc_tile, tok_tile) \
  const int oc_tile = (unit) / tok_tiles;            \
  const int tok_tile = (unit) % tok_tiles
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&afull[i], 4);
    }
    mbar_init(dfull, 1);
    mbar_init(dempty, 8);
    mbar_init(pfull, 4 * S);
    mbar_init(pempty, 4 * S);
    mbar_fence_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)L::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // remote mbarrier arrivals / multicast copies must not race the inits
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_tmem = tmem_base;             // columns [0, TN)
  const uint32_t a_tmem = tmem_base + TN;        // columns [TN, TN + 32 * STAGES)

  pdl_launch_dependents();

  if (warp == 0) {
    // =========================================================== TMA producer
    if (lane == 0) {
      pdl_wait_prior_grid();  // activations come from the previous kernel in the stream
      int s = 0;
      uint32_t ph = 0;
      [[maybe_unused]] int it = 0;
      for (int tile = tile0; tile < num_units; tile += tile_step) {
        B200AWQ_UNIT_DECODE(tile, oc_tile, tok_tile);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          mbar_wait(&empty[s], ph ^ 1);
          UMMA_STAMP(0, it);
          mbar_expect_tx(&full[s], (UMMA_DBG(8) ? 0 : L::kXBytes) + (UMMA_DBG(16) ? 0 : kWBytes));
          if (!UMMA_DBG(8)) tma_load_2d(xs + s * L::kXBytes, &tm_x, &full[s], (kb0 + kb) * kBK, tok_tile * TN);
          if (!UMMA_DBG(16)) tma_load_2d(ws + s * kWBytes, &tm_w, &full[s], (kb0 + kb) * kBK, oc_tile * (kBM / 4));
          UMMA_STAMP(1, it);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t kFmt = TypeTraits<T>::kIsBf16 ? 1u : 0u;
      // kind::f16 instruction descriptor: D = f32, A/B = f16|bf16, both K-major, N = TN, M = 128
      constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0, dph = 0;
      [[maybe_unused]] int it = 0;
      for (int tile = tile0; tile < num_units; tile += tile_step) {
        mbar_wait(dempty, dph ^ 1);  // accumulator drained by the epilogue of the previous tile
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          // dequantised weights are in TMEM slot s; the dequant warps saw full[s] first, so the activation tile of
          // the same barrier has landed as well
          mbar_wait(&afull[s], ph);
          UMMA_STAMP(4, it);
          tc_fence_after();
          const uint64_t bdesc = make_sw128_desc(smem_u32(xs + s * L::kXBytes));
          if (!UMMA_DBG(4)) {
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk) {
              umma_ts(d_tmem, a_tmem + s * 32 + kk * 8, bdesc + (uint64_t)(kk * 2) /* +32 B */, idesc,
                      (kb | kk) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty[s]);
          UMMA_STAMP(5, it);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        tc_commit(dfull);
        dph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // =========================================================== dequant + epilogue warps
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;      // output channel inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int grp = (warp - 4) >> 2;               // the two groups take alternate 64-k stages and alternate epilogue chunks
    const int et = ((warp - 4) & 3) * 32 + lane;  // 0..127 inside the group
    int s = 0;
    uint32_t ph = 0, dph = 0, pph = 0;
    [[maybe_unused]] int it0 = 0, tl = 0;
    for (int tile = tile0; tile < num_units; tile += tile_step, it0 += num_kb, ++tl) {
      B200AWQ_UNIT_DECODE(tile, oc_tile, tok_tile);
      const int n = oc_tile * kBM + row;
      const uint16_t* sp = reinterpret_cast<const uint16_t*>(scales) + (size_t)(kb0 / 2) * N + n;
      const uint16_t* zp = reinterpret_cast<const uint16_t*>(szeros) + (size_t)(kb0 / 2) * N + n;
      // scale / zero of this thread's channel, fetched FOUR 128-k groups ahead (a group = one stage of this warp
      // group; one group ahead left the global-load latency exposed at every stage: profiles/README.md)
      const int ngk = num_kb / 2;
      uint16_t sq[4], zq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sq[i] = (i < ngk) ? __ldg(sp + (size_t)i * N) : (uint16_t)0;
        zq[i] = (i < ngk) ? __ldg(zp + (size_t)i * N) : (uint16_t)0;
      }
      uint32_t s2 = 0, z2 = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if ((kb & 1) != grp) {  // the other group's stage (num_kb is even: a 128-k group is one stage of each)
          if (++s == STAGES) s = 0, ph ^= 1;
          continue;
        }
        s2 = splat16(sq[0]);
        z2 = splat16(zq[0]);
        sq[0] = sq[1], sq[1] = sq[2], sq[2] = sq[3];
        zq[0] = zq[1], zq[1] = zq[2], zq[2] = zq[3];
        if (kb / 2 + 4 < ngk) {
          sq[3] = __ldg(sp + (size_t)(kb / 2 + 4) * N);
          zq[3] = __ldg(zp + (size_t)(kb / 2 + 4) * N);
        }
        mbar_wait(&full[s], ph);
        if (threadIdx.x == 128 + grp * 128) UMMA_STAMP(2, it0 + kb);
        const uint8_t* wp = ws + s * kWBytes + (row >> 2) * 128 + (row & 3) * 32;
        const uint4 w0 = *reinterpret_cast<const uint4*>(wp);
        const uint4 w1 = *reinterpret_cast<const uint4*>(wp + 16);
        uint32_t r[32];
        if (UMMA_DBG(1)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = w0.x + i;
        } else {
          const uint32_t words[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t o[4];
              dequant_word<T>(words[b * 4 + u], s2, z2, o);
#pragma unroll
              for (int j = 0; j < 4; ++j) r[b * 16 + 4 * j + u] = o[j];  // column = (k in block) / 2
            }
        }
        // full[s] was armed only after the commit of the MMAs that last read TMEM slot s: the slot is free
        tc_fence_after();
        if (!UMMA_DBG(2)) {
          B200AWQ_TMEM_ST32(a_tmem + lane_base + s * 32, r);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&afull[s]);
        if (threadIdx.x == 128 + grp * 128) UMMA_STAMP(3, it0 + kb);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
      // ------------------------------------------------ epilogue for this tile
      mbar_wait(dfull, dph);
      dph ^= 1;
      tc_fence_after();
      if (threadIdx.x == 128) UMMA_STAMP(6, tl);
      if (S > 1 && grp == 1) {  // the partial-tile exchange is done by group 0 alone
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dempty);
        continue;
      }
      if (S > 1) {
        // split-K: fp32 partial tile -> own shared memory; every rank then reduces TN / S tokens of the tile
        // over distributed shared memory and stores them (no atomics, fixed summation order)
        if constexpr (TN <= 64) {
          mbar_wait_cluster(pempty, pph ^ 1);  // all ranks have finished reading my previous partial tile
#pragma unroll 1
          for (int c = 0; c < TN / 32; ++c) {
            uint32_t v[32];
            B200AWQ_TMEM_LD32(d_tmem + lane_base + c * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) pbuf[(c * 32 + j) * kBM + row] = __uint_as_float(v[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(dempty);  // the accumulator may be overwritten by the next tile
            for (int r = 0; r < S; ++r) mbar_arrive_remote(pfull, (uint32_t)r);
          }
          mbar_wait_cluster(pfull, pph);  // every rank's partial tile is in its shared memory
          const int tpr = TN / S;
          const uint32_t base = smem_u32(pbuf);
          for (int t = crank * tpr; t < (crank + 1) * tpr; ++t) {
            const int tok = tok_tile * TN + t;
            if (tok >= M) break;
            float acc = 0.f;
            for (int r = 0; r < S; ++r) acc += ld_cluster_f32(map_to_rank(base + (uint32_t)(t * kBM + et) * 4u, (uint32_t)r));
            y[(size_t)tok * N + oc_tile * kBM + et] = from_float<T>(acc);
          }
          __syncwarp();
          if (lane == 0)
            for (int r = 0; r < S; ++r) mbar_arrive_remote(pempty, (uint32_t)r);
          pph ^= 1;
        }
        continue;
      }
      // each group converts its own 32-token chunks (c & 1 == grp) through its own staging buffer
      uint8_t* ob = outb + grp * 32 * L::kOutRow;
#pragma unroll 1
      for (int c = grp; c < TN / 32; c += 2) {
        const int tok0 = tok_tile * TN + c * 32;
        if (tok0 >= M || UMMA_DBG(32)) break;  // uniform across the group
        uint32_t v[32];
        B200AWQ_TMEM_LD32(d_tmem + lane_base + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j)
          *reinterpret_cast<T*>(ob + j * L::kOutRow + row * 2) = from_float<T>(__uint_as_float(v[j]));
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = (et >> 4) + 8 * i, cc = et & 15;
          if (tok0 + rr < M) {
            const uint4 val = *reinterpret_cast<const uint4*>(ob + rr * L::kOutRow + cc * 16);
            *reinterpret_cast<uint4*>(y + (size_t)(tok0 + rr) * N + oc_tile * kBM + cc * 8) = val;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // buffer free for this group's next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dempty);
      if (threadIdx.x == 128) UMMA_STAMP(7, tl);
    }
    if (S > 1 && grp == 0) mbar_wait_cluster(pempty, pph ^ 1);  // nobody is still reading my last partial tile
  }

  __syncwarp();
  if (CS > 1) cluster_sync_all();  // no CTA of the cluster exits while a peer may still touch its shared memory / barriers
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::kTmemCols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T, int TN>
static int launch_umma_t(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K,
                         bool pdl, int max_ctas, int split, cudaStream_t stream) {
  using L = UmmaSmem<TN>;
  EncodeTiledFn enc = get_encoder();
  if (!enc) return B200AWQ_ERR_DRIVER;
  CUtensorMap tm_x, tm_w;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)TN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm_x, TypeTraits<T>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return B200AWQ_ERR_DRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)(N / 4)};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(kBM / 4)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(qw), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return B200AWQ_ERR_DRIVER;
  }
  auto kern = w4a16_umma_kernel<T, TN>;
  static bool attr_set[32] = {};  // per kernel instantiation and device
  if (cudaError_t ea = ensure_dyn_smem(kern, L::kTotalSplit, attr_set)) return (int)ea;
  const int tiles = ((M + TN - 1) / TN) * (N / kBM);
  int sms = sm_count();
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  // split-K over a cluster when the tiles alone cannot fill the machine (small token counts): S ranks share a tile,
  // each walks K / S and the partial tiles are reduced through distributed shared memory in the epilogue
  int S = 1;
  if (TN <= 64 && split > 1 && split <= 8 && (split & (split - 1)) == 0 && (K / kGroup) % split == 0) S = split;
  const int CS = S;
  int clusters = sms / CS;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * CS));
  cfg.blockDim = dim3(kUmmaThreads);
  cfg.dynamicSmemBytes = S > 1 ? L::kTotalSplit : L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (CS > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = CS;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_x, tm_w, (const T*)sc, (const T*)sz, (T*)y, M, N, K, S);
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_umma(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                bool pdl, const UmmaTuning& tune, cudaStream_t stream) {
  if (M < 1 || N % kBM || K % kGroup) return B200AWQ_ERR_SHAPE;
  int tn = tune.tn;
  int split = tune.split;
  if (tn != 32 && tn != 64 && tn != 128 && tn != 256) {
    // joint choice of the token tile and the k split by a small cost model (cycles per CTA, measured constants):
    // a 64-k stage costs max(dequantisation ~450, MMA 2*tn + 100) cycles, a tile's epilogue ~10*tn (+6000 when the
    // partial tiles of a split have to be reduced over the cluster), a launch ~12000; makespan = waves x that
    const int sms = sm_count();
    long best = -1;
    for (int c = 256; c >= 32; c >>= 1) {
      if (c > 32 && c / 2 >= M) continue;
      const int tiles = ((M + c - 1) / c) * (N / kBM);
      for (int sp = 1; sp <= 8; sp *= 2) {
        if (sp > 1 && (c > 64 || (K / kGroup) % sp || K / sp < 512 || tune.split == 1)) continue;
        if (tune.split > 1 && sp != tune.split && c <= 64) continue;
        const long per_tile = (long)(K / kBK / sp) * (2 * c + 100 > 450 ? 2 * c + 100 : 450) + 10 * c + (sp > 1 ? 6000 : 0);
        const long slots = sms / sp;  // clusters that run at once
        const long cost = 12000 + ((tiles + slots - 1) / slots) * per_tile;
        if (best < 0 || cost < best) best = cost, tn = c, split = sp;
      }
    }
  }
#define B200AWQ_UMMA_CASE(TT, TN_) \
  return launch_umma_t<TT, TN_>(x, qw, sc, sz, y, M, N, K, pdl, tune.max_ctas, split, stream)
  if (dtype == B200AWQ_DTYPE_F16) {
    switch (tn) {
      case 32: B200AWQ_UMMA_CASE(__half, 32);
      case 64: B200AWQ_UMMA_CASE(__half, 64);
      case 128: B200AWQ_UMMA_CASE(__half, 128);
      default: B200AWQ_UMMA_CASE(__half, 256);
    }
  } else {
    switch (tn) {
      case 32: B200AWQ_UMMA_CASE(__nv_bfloat16, 32);
      case 64: B200AWQ_UMMA_CASE(__nv_bfloat16, 64);
      case 128: B200AWQ_UMMA_CASE(__nv_bfloat16, 128);
      default: B200AWQ_UMMA_CASE(__nv_bfloat16, 256);
    }
  }
#undef B200AWQ_UMMA_CASE
}

}  // namespace b200awq
