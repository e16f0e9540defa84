// Large-batch (prefill) W4A16 kernel, second generation: 256 output channels x 128 tokens per CTA.
//
// Same building blocks as w4a16_umma.cu (TMA -> shared memory -> register dequant -> TENSOR MEMORY A operand ->
// tcgen05.mma TS form, fp32 accumulators in TMEM).  What changed, and why (profiles/README.md): at 2048 tokens the
// first-generation kernel (128 channels x 256 tokens per CTA) is bound by L2 -> SM bandwidth, not by the tensor
// pipe: every CTA streams its own copy of the 256-token activation tile, 36 KB per 64-k stage, 12.7 TB/s chip-wide
// against a ~12 TB/s L2 ceiling (stage time 800 cycles vs 512 for the MMAs).  Here one activation tile (128 tokens,
// 16 KB per stage) feeds TWO 128-channel MMAs, so a stage moves 24 KB for the same 2.1 M MACs:
//   TMEM (512 columns):  D0 [0,128) | D1 [128,256) | dequantised-A ring, 4 stages x 2 channel halves x 32 columns
//   smem per stage:      X 128 tok x 64 k (16 KB, 128-byte swizzle) + packed W 256 ch x 64 k (8 KB)
//   warps (20):          0 TMA | 1 MMA issuer | 2 TMEM alloc | 3 - | 4-19 four dequantisation groups of 4 warps:
//                        group (h, p) owns channel half h on the stages of parity p, and the epilogue chunks
//                        c = p, p + 2 of half h (tcgen05.ld -> cvt -> smem transpose -> 16-byte stores).
// Operands rn_T(q*s + z) are bit-identical to the reference's (gemm_cuda.cu:911-918); accumulation is fp32.
#include <cuda.h>

#include "w4_common.cuh"
#include "w4a16_kernels.h"

namespace b200awq {

constexpr int k2BM = 256;      // output channels per CTA (two UMMA M = 128 halves)
constexpr int k2TN = 128;      // tokens per CTA (UMMA N)
constexpr int k2BK = 64;       // k per stage
constexpr int k2AStages = 4;   // dequantised-A ring (per stage: 2 halves x 32 TMEM columns)
constexpr int k2Stages = 7;    // shared-memory ring
constexpr int k2Threads = 640;
constexpr int k2XBytes = k2TN * 128;            // 16384
constexpr int k2WBytes = k2BM * k2BK / 2;       // 8192
constexpr int k2StageBytes = k2XBytes + k2WBytes;
constexpr int k2OutRow = 256 + 16;              // 128 channels x 2 B (+ pad)
constexpr int k2OutBytes = 4 * 32 * k2OutRow;   // one 32-token staging buffer per group
constexpr int k2SmemTotal = k2Stages * k2StageBytes + k2OutBytes + 1024 + 512;

__device__ __forceinline__ void u2_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void u2_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void u2_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void u2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void u2_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint64_t u2_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

#define B200AWQ_U2_ST32(taddr, r)                                                                                       \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"  \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                              \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),      \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),        \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),        \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                     \
      : "memory")

#define B200AWQ_U2_LD32(taddr, r)                                                                                       \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"    \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
      : "r"(taddr)                                                                                                       \
      : "memory")

template <typename T>
__global__ void __launch_bounds__(k2Threads, 1)
w4a16_umma2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const T* __restrict__ scales, const T* __restrict__ szeros, T* __restrict__ y, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;                                   // [stages][16 KB]   (1024-aligned each)
  uint8_t* ws = xs + k2Stages * k2XBytes;               // [stages][8 KB]
  uint8_t* outb = ws + k2Stages * k2WBytes;             // [4 groups][32][k2OutRow]
  uint64_t* bars = reinterpret_cast<uint64_t*>(outb + k2OutBytes);
  uint64_t* full = bars;                   // [stages]   TMA -> dequant + MMA
  uint64_t* empty = full + k2Stages;       // [stages]   dequant (8 warps) + MMA commit (1) -> TMA
  uint64_t* afull = empty + k2Stages;      // [k2AStages] dequant (8 warps: both channel halves) -> MMA
  uint64_t* aempty = afull + k2AStages;    // [k2AStages] MMA commit -> dequant
  uint64_t* dfull = aempty + k2AStages;    // MMA commit -> epilogue
  uint64_t* dempty = dfull + 1;            // epilogue (16 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / k2BK;  // even (K % 128 == 0)
  const int tok_tiles = (M + k2TN - 1) / k2TN;
  const int num_tiles = tok_tiles * (N / k2BM);

  if (threadIdx.x == 0) {
    for (int i = 0; i < k2Stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 9);
    }
    for (int i = 0; i < k2AStages; ++i) {
      mbar_init(&afull[i], 8);
      mbar_init(&aempty[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(dempty, 16);
    mbar_fence_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  u2_fence_before();
  __syncthreads();
  u2_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_tmem = tmem_base;          // D0 columns [0,128), D1 [128,256)
  const uint32_t a_tmem = tmem_base + 256;    // ring: stage a, half h -> + (a * 2 + h) * 32

  pdl_launch_dependents();

  if (warp == 0) {
    // =========================================================== TMA producer
    if (lane == 0) {
      pdl_wait_prior_grid();
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int oc_tile = tile / tok_tiles, tok_tile = tile % tok_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], k2StageBytes);
          u2_tma_load_2d(xs + s * k2XBytes, &tm_x, &full[s], kb * k2BK, tok_tile * k2TN);
          u2_tma_load_2d(ws + s * k2WBytes, &tm_w, &full[s], kb * k2BK, oc_tile * (k2BM / 4));
          if (++s == k2Stages) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t kFmt = TypeTraits<T>::kIsBf16 ? 1u : 0u;
      constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(k2TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0, dph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(dempty, dph ^ 1);  // both accumulators drained by the epilogue of the previous tile
        u2_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          mbar_wait(&afull[as], aph);
          u2_fence_after();
          const uint64_t bdesc = u2_sw128_desc(smem_u32(xs + s * k2XBytes));
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < k2BK / 16; ++kk)
              u2_umma_ts(d_tmem + h * 128, a_tmem + (as * 2 + h) * 32 + kk * 8, bdesc + (uint64_t)(kk * 2), idesc,
                         (kb | kk) != 0 ? 1u : 0u);
          u2_commit(&empty[s]);
          u2_commit(&aempty[as]);
          if (++s == k2Stages) s = 0, ph ^= 1;
          if (++as == k2AStages) as = 0, aph ^= 1;
        }
        u2_commit(dfull);
        dph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // =========================================================== dequant + epilogue groups
    const int q = warp & 3;                    // TMEM lane quarter this warp may touch
    const int grp = (warp - 4) >> 2;           // 0..3
    const int h = grp & 1;                     // channel half (128 channels) this group owns
    const int par = grp >> 1;                  // stage parity / epilogue chunk parity this group owns
    const int row = q * 32 + lane;             // channel inside the half == TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int et = q * 32 + lane;              // 0..127 inside the group
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0, dph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int oc_tile = tile / tok_tiles, tok_tile = tile % tok_tiles;
      const int n = oc_tile * k2BM + h * 128 + row;
      const uint16_t* sp = reinterpret_cast<const uint16_t*>(scales) + n;
      const uint16_t* zp = reinterpret_cast<const uint16_t*>(szeros) + n;
      const int ngk = num_kb / 2;
      uint16_t sq[4], zq[4];  // scale / zero ring: fetched four 128-k groups ahead of their use
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sq[i] = (i < ngk) ? __ldg(sp + (size_t)i * N) : (uint16_t)0;
        zq[i] = (i < ngk) ? __ldg(zp + (size_t)i * N) : (uint16_t)0;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        if ((kb & 1) != par) {  // the other parity's stage (a 128-k group is one stage of each parity)
          if (++s == k2Stages) s = 0, ph ^= 1;
          if (++as == k2AStages) as = 0, aph ^= 1;
          continue;
        }
        const uint32_t s2 = splat16(sq[0]), z2 = splat16(zq[0]);
        sq[0] = sq[1], sq[1] = sq[2], sq[2] = sq[3];
        zq[0] = zq[1], zq[1] = zq[2], zq[2] = zq[3];
        if (kb / 2 + 4 < ngk) {
          sq[3] = __ldg(sp + (size_t)(kb / 2 + 4) * N);
          zq[3] = __ldg(zp + (size_t)(kb / 2 + 4) * N);
        }
        mbar_wait(&full[s], ph);
        const uint8_t* wp = ws + s * k2WBytes + h * 4096 + (row >> 2) * 128 + (row & 3) * 32;
        const uint4 w0 = *reinterpret_cast<const uint4*>(wp);
        const uint4 w1 = *reinterpret_cast<const uint4*>(wp + 16);
        uint32_t r[32];
        {
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
        mbar_wait(&aempty[as], aph ^ 1);
        u2_fence_after();
        B200AWQ_U2_ST32(a_tmem + lane_base + (as * 2 + h) * 32, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        u2_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&afull[as]);
          mbar_arrive(&empty[s]);
        }
        if (++s == k2Stages) s = 0, ph ^= 1;
        if (++as == k2AStages) as = 0, aph ^= 1;
      }
      // ------------------------------------------------ epilogue: chunks c = par, par + 2 of channel half h
      mbar_wait(dfull, dph);
      dph ^= 1;
      u2_fence_after();
      uint8_t* ob = outb + grp * 32 * k2OutRow;
#pragma unroll 1
      for (int c = par; c < k2TN / 32; c += 2) {
        const int tok0 = tok_tile * k2TN + c * 32;
        if (tok0 >= M) break;  // uniform across the group
        uint32_t v[32];
        B200AWQ_U2_LD32(d_tmem + lane_base + h * 128 + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j)
          *reinterpret_cast<T*>(ob + j * k2OutRow + row * 2) = from_float<T>(__uint_as_float(v[j]));
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = (et >> 4) + 8 * i, cc = et & 15;
          if (tok0 + rr < M) {
            const uint4 val = *reinterpret_cast<const uint4*>(ob + rr * k2OutRow + cc * 16);
            *reinterpret_cast<uint4*>(y + (size_t)(tok0 + rr) * N + oc_tile * k2BM + h * 128 + cc * 8) = val;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");  // buffer free for this group's next chunk
      }
      u2_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dempty);
    }
  }

  u2_fence_before();
  __syncthreads();
  if (warp == 2) {
    u2_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*Enc2Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static Enc2Fn enc2() {
  static Enc2Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<Enc2Fn>(p);
  }
  return fn;
}

template <typename T>
static int launch_umma2_t(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K,
                          bool pdl, int max_ctas, cudaStream_t stream) {
  Enc2Fn enc = enc2();
  if (!enc) return B200AWQ_ERR_DRIVER;
  CUtensorMap tm_x, tm_w;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)k2BK, (cuuint32_t)k2TN};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tm_x, TypeTraits<T>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
            const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200AWQ_ERR_DRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)(N / 4)};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)k2BK, (cuuint32_t)(k2BM / 4)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(qw), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200AWQ_ERR_DRIVER;
  }
  auto kern = w4a16_umma2_kernel<T>;
  static bool attr_set[32] = {};  // per kernel instantiation and device
  if (cudaError_t ea = ensure_dyn_smem(kern, k2SmemTotal, attr_set)) return (int)ea;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int tiles = ((M + k2TN - 1) / k2TN) * (N / k2BM);
  int ctas = sms;
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  if (tiles < ctas) ctas = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(k2Threads);
  cfg.dynamicSmemBytes = k2SmemTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  int na = 0;
  if (pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_x, tm_w, (const T*)sc, (const T*)sz, (T*)y, M, N, K);
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_umma2(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                 bool pdl, const UmmaTuning& tune, cudaStream_t stream) {
  if (M < 1 || N % k2BM || K % kGroup) return B200AWQ_ERR_SHAPE;
  if (dtype == B200AWQ_DTYPE_F16) return launch_umma2_t<__half>(x, qw, sc, sz, y, M, N, K, pdl, tune.max_ctas, stream);
  return launch_umma2_t<__nv_bfloat16>(x, qw, sc, sz, y, M, N, K, pdl, tune.max_ctas, stream);
}

}  // namespace b200awq
