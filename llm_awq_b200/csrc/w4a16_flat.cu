// Skinny-batch W4A16 kernel on the 5th-gen tensor cores (1 <= M <= 64 tokens), HBM-bound.
//
// Replaces the reference's gemv_kernel (quantization_new/gemv/gemv_cuda.cu:74-229) and its
// small-M split-K GEMM tiles (gemm_cuda.cu:1155-1206) for shapes with N % 128 == 0.
//
// Same operand orientation as the prefill kernel (w4a16_umma.cu): D[oc, tok] = W[oc, k] . X[tok, k]^T,
// the dequantised weight is the A operand in TENSOR MEMORY (lane = output channel, column = one
// k-adjacent pair = one LOP3 result), the activation tile is the K-major B operand in shared memory
// (TMA, 128-byte swizzle).  What is different, because this kernel is bound by HBM and by the CUDA-core
// dequantisation rate rather than by the tensor pipe:
//   * one CTA = 128 output channels x Kc input channels of ONE launch-wide split of k over a thread-block
//     cluster (S <= 8 CTAs, reduced through distributed shared memory: no atomics, no workspace);
//     small footprint (256 threads, 128 TMEM columns, ~50 KB smem) so that 4 CTAs co-reside per SM and
//     the next launch's CTAs can prefetch under programmatic dependent launch;
//   * the packed-weight ring is filled BEFORE griddepcontrol.wait (weights do not depend on the previous
//     kernel) and the dequantisation into the TMEM ring starts as soon as bytes land; only the activation
//     TMA, the MMAs and the drain are ordered after the wait;
//   * fp16 ("group-factored", MODE 2): the A operand is the RAW LOP3 output (1024+q for low nibbles,
//     64+q for high nibbles): 5 ALU ops per 8 weights and nothing else per weight.  Each 128-k group
//     accumulates into its own TMEM accumulator; scale, zero and the bias are applied when a pass of
//     groups is drained:  y += s_g * (D_g - C_g) + z_g * X_g  with X_g = sum_k x_k and
//     C_g = 1024 * X_lo + 64 * X_hi computed once per CTA by a spare warp.  Exact in the integer q.
//   * bf16 (MODE 0): operands rn_bf16(q*s + z) bit-identical to the reference's, one accumulator.
// Warp roles (8 warps): 0 weight TMA producer | 1 activation TMA producer | 2 TMEM alloc + MMA issuer |
// 3 group sums of x | 4-7 dequantise -> TMEM, drain, cluster reduce, store.
#include <cuda.h>

#include "w4_common.cuh"
#include "w4a16_kernels.h"

namespace b200awq {

constexpr int kFlatThreads = 256;
constexpr int kFlatBM = 128;   // output channels per CTA (TMEM lanes)
constexpr int kFlatBK = 64;    // k per stage
constexpr int kFlatNA = 2;     // dequantised-A ring in TMEM (32 columns each)
constexpr int kFlatDCols = 64; // accumulator columns (one pass of groups)
constexpr int kFlatWBytes = kFlatBM * kFlatBK / 2;  // 4096

#ifdef B200AWQ_TRACE
extern __device__ unsigned long long g_trace_buf[1024 * 2 * 8];  // defined in w4a16_stream.cu
__device__ __forceinline__ unsigned long long flat_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FLAT_STAMP(cond, ev)                                         \
  if ((cond) && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))   \
  g_trace_buf[(((unsigned)P.seq & 1023u) * 2 + (blockIdx.x != 0)) * 8 + (ev)] = flat_timer_ns()
#else
#define FLAT_STAMP(cond, ev)
#endif

struct FlatParams {
  int seq;     // launch sequence number (debug trace only)
  int nst;     // 64-k stages of this CTA's k range
  int nw;      // weight ring slots (4 KB each)
  int nx;      // activation ring slots (TN * 128 B each)
  int ng;      // 128-k groups of this CTA's k range
  int off_x, off_w, off_sum, off_rbuf, off_bar, total;  // shared-memory carve-up (off_x is 1024-aligned at run time)
};

// ---------------------------------------------------------------- tcgen05 / TMA PTX (same encodings as w4a16_umma.cu)
__device__ __forceinline__ void f_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void f_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void f_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint64_t f_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

#define B200AWQ_FLAT_ST32(taddr, r)                                                                                     \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"  \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                              \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),      \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),        \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),        \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                     \
      : "memory")

// tcgen05.ld of MT consecutive 32-bit columns of this thread's TMEM lane
template <int MT>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[MT]) {
  if constexpr (MT == 1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr) : "memory");
  } else if constexpr (MT == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
  } else {
    static_assert(MT % 16 == 0, "MT in {1, 4, 16, 32, 64}");
#pragma unroll
    for (int c = 0; c < MT; c += 16) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]), "=r"(v[c + 6]),
            "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]), "=r"(v[c + 12]),
            "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15])
          : "r"(taddr + c)
          : "memory");
    }
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One packed word -> the 4 pair-registers of this word in TMEM column order (MODE 2: raw biased, MODE 0: exact w~)
template <typename T, int MODE>
__device__ __forceinline__ void flat_word(uint32_t w, uint32_t s2, uint32_t z2, uint32_t (&o)[4]) {
  if constexpr (MODE == 2) unpack_word_biased<T>(w, o);
  else dequant_word<T>(w, s2, z2, o);
}

template <typename T, int TN, int MT, int MODE>
__global__ void __launch_bounds__(kFlatThreads, MT <= 4 ? 4 : (MT <= 16 ? 3 : 1))
w4a16_flat_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const T* __restrict__ scales, const T* __restrict__ szeros, const T* __restrict__ x, T* __restrict__ y,
                  int M, int N, int K, int Kc, int S, const __grid_constant__ FlatParams P) {
  constexpr bool kBf16 = TypeTraits<T>::kIsBf16;
  constexpr int GP = (MODE == 2) ? kFlatDCols / TN : 1;  // groups per accumulator pass (MODE 0: one accumulator)
  constexpr int kXBytes = TN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem + P.off_x;                                   // [nx][TN * 128]  (1024-aligned each)
  uint8_t* ws = smem + P.off_w;                                   // [nw][4096]
  float2* xsum = reinterpret_cast<float2*>(smem + P.off_sum);     // [token][group] {X, C}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.off_bar);
  uint64_t* wfull = bars;
  uint64_t* wempty = wfull + P.nw;
  uint64_t* xfull = wempty + P.nw;
  uint64_t* xempty = xfull + P.nx;
  uint64_t* afull = xempty + P.nx;
  uint64_t* aempty = afull + kFlatNA;
  uint64_t* dfull = aempty + kFlatNA;
  uint64_t* dempty = dfull + 1;
  uint64_t* sfull = dempty + 1;
  uint64_t* rbar = sfull + 1;  // split k: counts the bytes the cluster pushes into rbuf
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 1);
  float* rbuf = reinterpret_cast<float*>(smem + P.off_rbuf);  // [source rank][slice] fp32 partial sums pushed by the cluster

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (S > 1) ? (int)cluster_ctarank() : 0;
  const int rt = (S > 1) ? (int)(blockIdx.x / S) : (int)blockIdx.x;  // 128-row tile
  const int k0 = rank * Kc;
  const int nst = P.nst, ng = P.ng;
  FLAT_STAMP(threadIdx.x == 128, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nw; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 4);
    }
    for (int i = 0; i < P.nx; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < kFlatNA; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(dempty, 4);
    mbar_init(sfull, 1);
    mbar_init(rbar, 1);
    mbar_fence_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  f_fence_before();
  __syncthreads();
  if (S > 1) cluster_sync_all();  // barrier inits visible cluster-wide before any push (before the PDL wait: overlapped)
  f_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_tmem = tmem_base;                 // columns [0, 64): accumulators of one pass
  const uint32_t a_tmem = tmem_base + kFlatDCols;    // columns [64, 128): dequantised-A ring
  pdl_launch_dependents();
  FLAT_STAMP(threadIdx.x == 128, 1);

  if (warp == 0) {
    // =========================================================== packed-weight producer (not ordered after the previous grid)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nst; ++i) {
        mbar_wait(&wempty[s], ph ^ 1);
        mbar_expect_tx(&wfull[s], kFlatWBytes);
        f_tma_load_2d(ws + s * kFlatWBytes, &tm_w, &wfull[s], k0 + i * kFlatBK, rt * (kFlatBM / 4));
        if (++s == P.nw) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // =========================================================== activation producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
      pdl_wait_prior_grid();
      FLAT_STAMP(true, 3);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nst; ++i) {
        mbar_wait(&xempty[s], ph ^ 1);
        mbar_expect_tx(&xfull[s], kXBytes);
        f_tma_load_2d(xs + s * kXBytes, &tm_x, &xfull[s], k0 + i * kFlatBK, 0);
        if (++s == P.nx) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 2) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t kFmt = kBf16 ? 1u : 0u;
      constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kFlatBM >> 4) << 24);
      int sx = 0, sa = 0;
      uint32_t phx = 0, pha = 0;
      const int spp = 2 * GP;  // stages per pass (MODE 2)
      for (int i = 0; i < nst; ++i) {
        int gi = 0;
        bool first_of_acc, last_of_pass;
        if (MODE == 2) {
          const int in_pass = i % spp;
          gi = in_pass >> 1;
          first_of_acc = (in_pass & 1) == 0;
          last_of_pass = (in_pass == spp - 1) || (i == nst - 1);
          if (in_pass == 0 && i > 0) {
            mbar_wait(dempty, (uint32_t)((i / spp - 1) & 1));  // previous pass drained
            f_fence_after();
          }
        } else {
          first_of_acc = (i == 0);
          last_of_pass = (i == nst - 1);
        }
        mbar_wait(&xfull[sx], phx);
        mbar_wait(&afull[sa], pha);
        f_fence_after();
        const uint64_t bdesc = f_sw128_desc(smem_u32(xs + sx * kXBytes));
#pragma unroll
        for (int kk = 0; kk < kFlatBK / 16; ++kk)
          f_umma_ts(d_tmem + gi * TN, a_tmem + sa * 32 + kk * 8, bdesc + (uint64_t)(kk * 2), idesc,
                    (first_of_acc && kk == 0) ? 0u : 1u);
        f_commit(&xempty[sx]);
        f_commit(&aempty[sa]);
        if (last_of_pass) f_commit(dfull);
        if (++sx == P.nx) sx = 0, phx ^= 1;
        if (++sa == kFlatNA) sa = 0, pha ^= 1;
      }
    }
  } else if (warp == 3) {
    // =========================================================== group sums of the activations (MODE 2)
    if (MODE == 2) {
      pdl_wait_prior_grid();
      // 4 (token, group) pairs per pass: 8 lanes x 16 elements; the first 16 B of a lane's chunk sit on
      // low-nibble positions (bias 1024), the second on high-nibble positions (bias 64)
      const int sub = lane >> 3, c = lane & 7;
      const int pairs = M * ng;
      constexpr int UNR = 4;  // independent (token, group) quads in flight per lane (L2 latency)
      for (int p0 = 0; p0 < pairs; p0 += 4 * UNR) {
        uint4 va[UNR], vb[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int p = p0 + 4 * u + sub;
          va[u] = vb[u] = make_uint4(0, 0, 0, 0);
          if (p < pairs) {
            const int t = p / ng, g = p - t * ng;
            const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)t * K + k0 + g * kGroup + c * 16);
            va[u] = __ldg(src), vb[u] = __ldg(src + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int p = p0 + 4 * u + sub;
          float lo = 0.f, hi = 0.f;
          const uint32_t av[4] = {va[u].x, va[u].y, va[u].z, va[u].w}, bv[4] = {vb[u].x, vb[u].y, vb[u].z, vb[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 fa, fb;
            if constexpr (kBf16) {
              fa = __bfloat1622float2(u32_as_b2(av[e]));
              fb = __bfloat1622float2(u32_as_b2(bv[e]));
            } else {
              fa = __half22float2(u32_as_h2(av[e]));
              fb = __half22float2(u32_as_h2(bv[e]));
            }
            lo += fa.x + fa.y;
            hi += fb.x + fb.y;
          }
#pragma unroll
          for (int d = 1; d < 8; d <<= 1) {
            lo += __shfl_xor_sync(0xffffffffu, lo, d);
            hi += __shfl_xor_sync(0xffffffffu, hi, d);
          }
          if (p < pairs && c == 0) {
            const float X = lo + hi;
            xsum[p] = make_float2(X, kBf16 ? 128.f * X : 1024.f * lo + 64.f * hi);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sfull);
    }
  } else {
    // =========================================================== dequantise -> TMEM, drain, reduce, store
    const int q = warp & 3;
    const int row = q * 32 + lane;  // output channel inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int n = rt * kFlatBM + row;
    const uint16_t* sp = reinterpret_cast<const uint16_t*>(scales) + (size_t)(k0 / kGroup) * N + n;
    const uint16_t* zp = reinterpret_cast<const uint16_t*>(szeros) + (size_t)(k0 / kGroup) * N + n;
    float acc[MT];
#pragma unroll
    for (int t = 0; t < MT; ++t) acc[t] = 0.f;

    // scale / zero of this thread's channel for the groups of one accumulator pass, fetched one pass ahead
    float sv[GP], zv[GP];
    auto load_scales = [&](int pass) {
#pragma unroll
      for (int gi = 0; gi < GP; ++gi) {
        const int g = pass * GP + gi;
        sv[gi] = zv[gi] = 0.f;
        if (MODE == 2 && g < ng) {
          sv[gi] = bits16_to_float(__ldg(sp + (size_t)g * N), kBf16);
          zv[gi] = bits16_to_float(__ldg(zp + (size_t)g * N), kBf16);
        }
      }
    };
    load_scales(0);
    bool sums_ready = false;

    // drain of one pass of accumulators (MODE 2: groups [g_first, g_first + count) of this CTA)
    auto drain = [&](int pass, int g_first, int count) {
      if (MODE == 2 && !sums_ready) {
        mbar_wait(sfull, 0);
        sums_ready = true;
      }
      mbar_wait(dfull, (uint32_t)(pass & 1));
      f_fence_after();
      if (MODE == 2) {
#pragma unroll
        for (int gi = 0; gi < GP; ++gi) {
          if (gi < count) {
            uint32_t v[MT];
            tmem_ld_cols<MT>(d_tmem + lane_base + gi * TN, v);
            const int g = g_first + gi;
#pragma unroll
            for (int t = 0; t < MT; ++t) {
              if (t < M) {
                const float2 xc = xsum[t * ng + g];
                acc[t] = fmaf(zv[gi], xc.x, fmaf(sv[gi], __uint_as_float(v[t]) - xc.y, acc[t]));
              }
            }
          }
        }
      } else {
        uint32_t v[MT];
        tmem_ld_cols<MT>(d_tmem + lane_base, v);
#pragma unroll
        for (int t = 0; t < MT; ++t) acc[t] = __uint_as_float(v[t]);
      }
      f_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dempty);
      load_scales(pass + 1);
    };

    int sw = 0, sa = 0;
    uint32_t phw = 0, pha = 0;
    uint32_t s2 = 0, z2 = 0;
    int pending = -1;  // pass whose accumulators still have to be drained
    constexpr int spp = 2 * GP;
    for (int i = 0; i < nst; ++i) {
      if (MODE == 0 && (i & 1) == 0) {
        s2 = splat16(__ldg(sp + (size_t)(i >> 1) * N));
        z2 = splat16(__ldg(zp + (size_t)(i >> 1) * N));
      }
      if (MODE == 2 && pending >= 0 && (i % spp) == 1) {  // one stage of the next pass is already queued
        drain(pending, pending * GP, GP);
        pending = -1;
      }
      mbar_wait(&wfull[sw], phw);
      if (i == 0) FLAT_STAMP(threadIdx.x == 128, 2);
      const uint8_t* wp = ws + sw * kFlatWBytes + (row >> 2) * 128 + (row & 3) * 32;
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
            flat_word<T, MODE>(words[b * 4 + u], s2, z2, o);
#pragma unroll
            for (int j = 0; j < 4; ++j) r[b * 16 + 4 * j + u] = o[j];  // column = (k in block) / 2
          }
      }
      mbar_wait(&aempty[sa], pha ^ 1);
      f_fence_after();
      B200AWQ_FLAT_ST32(a_tmem + lane_base + sa * 32, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      f_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&afull[sa]);
        mbar_arrive(&wempty[sw]);
      }
      if (++sw == P.nw) sw = 0, phw ^= 1;
      if (++sa == kFlatNA) sa = 0, pha ^= 1;
      if (MODE == 2 && ((i % spp) == spp - 1) && i != nst - 1) pending = i / spp;
    }
    FLAT_STAMP(threadIdx.x == 128, 4);
    if (MODE == 2) {
      if (pending >= 0) drain(pending, pending * GP, GP);
      const int last = (nst - 1) / spp;
      drain(last, last * GP, ng - last * GP);
    } else {
      drain(0, 0, 1);
    }

    FLAT_STAMP(threadIdx.x == 128, 5);
    // ---- output: direct (S == 1) or through the cluster (split k)
    if (S == 1) {
#pragma unroll
      for (int t = 0; t < MT; ++t)
        if (t < M) y[(size_t)t * N + n] = from_float<T>(acc[t]);
    } else {
      // split k without cluster barriers / fences at the end (each is a GPU-scope MEMBAR): every rank PUSHES the
      // partial sums of slice j of the tile into rank j's shared memory (st.async, 4 bytes completed on rank j's
      // mbarrier per store), then reduces its own slice in fixed source order and stores it.
      const int total = M * kFlatBM;
      const int per = (total + S - 1) / S;
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        if (t < M) {
          const int e = t * kFlatBM + row;
          const int j = e / per;
          const uint32_t dst = map_to_rank(smem_u32(&rbuf[rank * per + (e - j * per)]), (uint32_t)j);
          const uint32_t dbar = map_to_rank(smem_u32(rbar), (uint32_t)j);
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst),
                       "r"(__float_as_uint(acc[t])), "r"(dbar)
                       : "memory");
        }
      }
      const int lo = rank * per, hi = min(total, lo + per);
      const int e0 = (int)threadIdx.x - 128;  // 0..127
      if (e0 == 0) mbar_expect_tx(rbar, (uint32_t)(hi > lo ? (hi - lo) * S * 4 : 0));
      mbar_wait(rbar, 0);
      for (int e = lo + e0; e < hi; e += 128) {
        float v = 0.f;
        for (int r = 0; r < S; ++r) v += rbuf[r * per + (e - lo)];
        const int t = e / kFlatBM, rr = e - t * kFlatBM;
        y[(size_t)t * N + rt * kFlatBM + rr] = from_float<T>(v);
      }
    }
  }

  FLAT_STAMP(threadIdx.x == 128, 6);
  f_fence_before();
  __syncthreads();
  if (warp == 2) {
    f_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*FlatEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static FlatEncodeFn flat_encoder() {
  static FlatEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<FlatEncodeFn>(p);
  }
  return fn;
}

static int flat_pick_splits(int N, int K, int kc_env) {
  const int groups = K / kGroup;
  if (kc_env > 0 && kc_env % kGroup == 0 && K % kc_env == 0 && K / kc_env <= 8) return K / kc_env;
  // smallest split (<= 8, dividing the group count) whose slice is <= 1024 k, but at least enough CTAs for ~2 per SM
  int best = 1;
  for (int s = 1; s <= 8; ++s) {
    if (groups % s) continue;
    best = s;
    const int kc = K / s;
    if (kc <= 512) break;
    if (kc <= 1024 && (N / kFlatBM) * s >= 2 * 148) break;
  }
  return best;
}

template <typename T, int TN, int MT, int MODE>
static int launch_flat_t(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int S,
                         bool pdl, cudaStream_t stream) {
  FlatEncodeFn enc = flat_encoder();
  if (!enc) return B200AWQ_ERR_DRIVER;
  const int Kc = K / S;
  FlatParams P{};
  static int seq_counter = 0;
  P.seq = seq_counter++;
  P.nst = Kc / kFlatBK;
  P.ng = Kc / kGroup;
  P.nw = P.nst < 12 ? P.nst : 12;
  P.nx = P.nst < 8 ? P.nst : 8;
  int off = 0;
  P.off_x = off, off += P.nx * TN * 128;
  P.off_w = off, off += P.nw * kFlatWBytes;
  P.off_sum = off, off += (MODE == 2 ? M * P.ng * 8 : 0);
  P.off_rbuf = off, off += (S > 1 ? (M * kFlatBM + 8 * S) * 4 : 0);
  off = (off + 15) & ~15;
  P.off_bar = off, off += (2 * P.nw + 2 * P.nx + 2 * kFlatNA + 4) * 8 + 16;
  P.total = off + 1024;  // run-time 1024-byte alignment of the base
  if (P.total > 200 * 1024) return B200AWQ_ERR_SHAPE;

  CUtensorMap tm_x, tm_w;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kFlatBK, (cuuint32_t)TN};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tm_x, TypeTraits<T>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
            const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200AWQ_ERR_DRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)(N / 4)};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kFlatBK, (cuuint32_t)(kFlatBM / 4)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(qw), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return B200AWQ_ERR_DRIVER;
  }
  auto kern = w4a16_flat_kernel<T, TN, MT, MODE>;
  static bool attr_set[32] = {};  // per kernel instantiation and device
  if (cudaError_t ea = ensure_dyn_smem(kern, 200 * 1024, attr_set)) return (int)ea;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(N / kFlatBM) * S);
  cfg.blockDim = dim3(kFlatThreads);
  cfg.dynamicSmemBytes = P.total;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (S > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = S;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_x, tm_w, (const T*)sc, (const T*)sz, (const T*)x, (T*)y, M, N, K, Kc, S, P);
  return e == cudaSuccess ? 0 : (int)e;
}

template <typename T, int MODE>
static int launch_flat_m(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int S,
                         bool pdl, cudaStream_t stream) {
  if (M == 1) return launch_flat_t<T, 16, 1, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
  if (M <= 4) return launch_flat_t<T, 16, 4, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
  if (M <= 16) return launch_flat_t<T, 16, 16, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
  return launch_flat_t<T, 64, 64, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
}

int launch_flat(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                bool pdl, const FlatTuning& tune, cudaStream_t stream) {
  if (M < 1 || M > 64 || N % kFlatBM || K % kGroup) return B200AWQ_ERR_SHAPE;
  const int S = flat_pick_splits(N, K, tune.kc);
  if (dtype == B200AWQ_DTYPE_F16) return launch_flat_m<__half, 2>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
  return launch_flat_m<__nv_bfloat16, 0>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
}

}  // namespace b200awq
