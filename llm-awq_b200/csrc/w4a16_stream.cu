// HBM-bound W4A16 streaming kernel for small token counts (decode GEMV, M <= 16).
//
// Replaces the reference's gemv_kernel (quantization_new/gemv/gemv_cuda.cu:74-229) and,
// for 8 <= M <= 16, its split-K GEMM tile (gemm_cuda.cu:1155-1168).  Design (B200-first):
//   * one CTA owns 8*RO output channels x Kc input channels.  Its slice of `qweight`
//     is 2*RO contiguous byte ranges (the packing interleaves 4 rows, so a 4-row "quad"
//     is contiguous over k); they are pulled into shared memory with bulk async copies
//     (cp.async.bulk -> SASS UBLKCP) issued by ONE thread at the very top of the kernel,
//     so all of the CTA's HBM traffic is in flight at once and no registers are pinned
//     by outstanding loads.  Several CTAs co-reside per SM -> 100+ KB in flight per SM.
//   * the weight prefetch is issued BEFORE griddepcontrol.wait: with programmatic
//     dependent launch the next linear layer streams its weights while the previous
//     one is still reducing / draining.  Activations are only touched after the wait.
//   * each lane dequantises its own 16-byte chunks (32 weights of ONE output channel,
//     for two channels 8 apart) in registers and feeds them as the A operand of
//     m16n8k16 tensor-core MACs (rows = 16 output channels, columns = 8 tokens).  The k
//     order inside a fragment is simply "whatever the lane holds": A and B use the same
//     lane-local k slots, so the reference's register-oriented packing needs no shuffles;
//     the activations are re-ordered once while they are staged into shared memory so
//     that every B fragment is one aligned register pair.  Accumulation is fp32.
//   * the 4 warps of a CTA split the k groups; partials are reduced through shared
//     memory, and across the CTAs of a thread-block cluster (split-K for large k)
//     through distributed shared memory.  No global atomics, no workspace.
//
// MODE 0 (default): operands w~ = rn_T(q*s + z) bit-identical to the reference's.
// MODE 1: exact integer q as operand; scale/zero applied per 128-k group in fp32:
//         y += s*(sum_k q_k x_k) + z*(sum_k x_k).
// MODE 2: raw biased operands (1024+q | 1024+16q for fp16, 128+q for bf16) with the
//         high-nibble activations pre-scaled by 1/16 (fp16); the bias is removed per group
//         in fp32:  y += s*(acc - C) + z*X,  C = 1024*X_lo + 64*X_hi  (bf16: 128*X).
#include "w4_common.cuh"
#include "w4a16_kernels.h"

namespace b200awq {

constexpr int kStreamThreads = 128;
constexpr int kStreamWarps = 4;

template <typename T, int RO, int TT, int MODE>
__global__ void __launch_bounds__(kStreamThreads)
w4a16_stream_kernel(const T* __restrict__ x, const uint16_t* __restrict__ qw, const T* __restrict__ scales,
                    const T* __restrict__ szeros, T* __restrict__ y, int M, int N, int K, int Kc, int S) {
  constexpr bool kBf16 = TypeTraits<T>::kIsBf16;
  constexpr int R = 8 * RO;  // output channels per CTA
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int rank = (S > 1) ? (int)cluster_ctarank() : 0;
  const int rb = blockIdx.x / S;
  const int n0 = rb * R;
  const int kbase = rank * Kc;
  const int ngroups = Kc / kGroup;
  const int wrow = Kc * 2;       // bytes of one 4-row quad over this CTA's k range
  const int xrow = Kc * 2 + 16;  // padded activation row (bank spread between tokens)

  uint8_t* wbuf = smem;
  uint8_t* xbuf = wbuf + 2 * RO * wrow;
  float* red = reinterpret_cast<float*>(xbuf + M * xrow);  // [warp][tt][16 rows][8 tok]
  float* cpart = red + kStreamWarps * TT * 128;            // [tt][16][8]
  float2* xsum = reinterpret_cast<float2*>(cpart + TT * 128);  // [8 TT tok][group] {X, C}   (MODE >= 1)
  uint8_t* zblk = reinterpret_cast<uint8_t*>(xsum + (MODE ? 8 * TT * ngroups : 0));  // 64 zero bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(zblk + 64);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_fence_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bars[0], 2 * RO * wrow);
#pragma unroll
    for (int qd = 0; qd < 2 * RO; ++qd)
      bulk_g2s(wbuf + qd * wrow, qw + (size_t)(rb * 2 * RO + qd) * K + kbase, wrow, &bars[0]);
  }
  {
    float* xz = reinterpret_cast<float*>(xsum);  // xsum (if any) and the zero block are contiguous
    for (int i = tid; i < (MODE ? 16 * TT * ngroups : 0) + 16; i += kStreamThreads) xz[i] = 0.f;
  }

  // scale / zero of this lane's output channels (row g of each octet), one 16-bit value each
  const uint16_t* sp = reinterpret_cast<const uint16_t*>(scales) + (size_t)(kbase / kGroup) * N + n0 + g;
  const uint16_t* zp = reinterpret_cast<const uint16_t*>(szeros) + (size_t)(kbase / kGroup) * N + n0 + g;
  uint32_t s_cur[RO], z_cur[RO];
#pragma unroll
  for (int ro = 0; ro < RO; ++ro) {
    s_cur[ro] = 0;
    z_cur[ro] = 0;
  }
  sp += (size_t)warp * N;
  zp += (size_t)warp * N;
  const size_t sstep = (size_t)kStreamWarps * N;
  if (warp < ngroups) {
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {
      s_cur[ro] = __ldg(sp + 8 * ro);
      z_cur[ro] = __ldg(zp + 8 * ro);
    }
  }

  pdl_launch_dependents();
  __syncthreads();        // barrier init + xsum zeroing visible
  pdl_wait_prior_grid();  // activations (and y) belong to the stream order from here on

  // ---- stage activations: x[m, kbase + ...] -> xbuf, re-ordered per 16-element chunk
  //      (pairs p0..p7) -> (p0,p4,p1,p5 | p2,p6,p3,p7) so that the pair of B registers of one
  //      MMA is one aligned 8-byte piece; MODE 2/fp16 also scales the upper half by 1/16.
  {
    const int chunks = Kc / 16;
    for (int idx = tid; idx < M * chunks; idx += kStreamThreads) {
      const int m = idx / chunks, c = idx - m * chunks;
      const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)m * K + kbase + c * 16);
      uint4 a = __ldg(src), b = __ldg(src + 1);
      if (MODE) {
        float lo = 0.f, hi = 0.f;
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 fa, fb;
          if constexpr (kBf16) {
            fa = __bfloat1622float2(u32_as_b2(av[i]));
            fb = __bfloat1622float2(u32_as_b2(bv[i]));
          } else {
            fa = __half22float2(u32_as_h2(av[i]));
            fb = __half22float2(u32_as_h2(bv[i]));
          }
          lo += fa.x + fa.y;
          hi += fb.x + fb.y;
        }
        float X = lo + hi, C = 0.f;
        if (MODE == 2) C = kBf16 ? 128.f * X : 1024.f * lo + 64.f * hi;
        float* dst = reinterpret_cast<float*>(&xsum[m * ngroups + (c >> 3)]);
        atomicAdd(dst, X);
        if (MODE == 2) atomicAdd(dst + 1, C);
        if (MODE == 2 && !kBf16) {
          const __half2 k16th = u32_as_h2(0x2c002c00u);
          b.x = h2_as_u32(__hmul2(u32_as_h2(b.x), k16th));
          b.y = h2_as_u32(__hmul2(u32_as_h2(b.y), k16th));
          b.z = h2_as_u32(__hmul2(u32_as_h2(b.z), k16th));
          b.w = h2_as_u32(__hmul2(u32_as_h2(b.w), k16th));
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(xbuf + m * xrow + c * 32);
      dst[0] = make_uint4(a.x, b.x, a.y, b.y);
      dst[1] = make_uint4(a.z, b.z, a.w, b.w);
    }
  }
  __syncthreads();
  mbar_wait(&bars[0], 0);

  float acc[TT][4];
#pragma unroll
  for (int t = 0; t < TT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};

  // per-lane walking pointers (lanes whose token does not exist read the zero block forever)
  const uint8_t* xp[TT];
  int xstep[TT];
  const float2* sump[TT][2];
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const int tok = g + 8 * t;
    const bool live = tok < M;
    xp[t] = live ? xbuf + tok * xrow + (warp * kGroup + tig * 32) * 2 : zblk;
    xstep[t] = live ? kStreamWarps * kGroup * 2 : 0;
    sump[t][0] = xsum + (8 * t + 2 * tig) * ngroups + warp;
    sump[t][1] = sump[t][0] + ngroups;
  }
  const uint8_t* wp = wbuf + (g >> 2) * wrow + (warp * 2 + (tig >> 1)) * 128 + (g & 3) * 32 + (tig & 1) * 16;

  for (int G = warp; G < ngroups; G += kStreamWarps) {
    uint32_t s_nxt[RO], z_nxt[RO];
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {
      s_nxt[ro] = 0;
      z_nxt[ro] = 0;
    }
    sp += sstep;
    zp += sstep;
    if (G + kStreamWarps < ngroups) {
#pragma unroll
      for (int ro = 0; ro < RO; ++ro) {
        s_nxt[ro] = __ldg(sp + 8 * ro);
        z_nxt[ro] = __ldg(zp + 8 * ro);
      }
    }
    // B fragments: token g (+ 8 t) of this lane, the 32 k of block (G, tig), staged order
    uint4 xq[TT][4];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) xq[t][i] = reinterpret_cast<const uint4*>(xp[t])[i];
      xp[t] += xstep[t];
    }
    // A: the 16-byte chunk of (row n0 + 8 ro + g, 32-k block 4G + tig) for ro = 0 (rows 0-7 of
    // the MMA) and ro = 1 (rows 8-15)
    uint32_t wa[4], wb[4] = {0, 0, 0, 0};
    {
      const uint4 v = *reinterpret_cast<const uint4*>(wp);
      wa[0] = v.x, wa[1] = v.y, wa[2] = v.z, wa[3] = v.w;
      if (RO == 2) {
        const uint4 v2 = *reinterpret_cast<const uint4*>(wp + 2 * wrow);
        wb[0] = v2.x, wb[1] = v2.y, wb[2] = v2.z, wb[3] = v2.w;
      }
      wp += kStreamWarps * 256;
    }
    const uint32_t sa2 = s_cur[0] * 0x00010001u, za2 = z_cur[0] * 0x00010001u;
    const uint32_t sb2 = s_cur[RO - 1] * 0x00010001u, zb2 = z_cur[RO - 1] * 0x00010001u;

    float part[TT][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t oa[4], ob[4] = {0, 0, 0, 0};
      if (MODE == 0) {
        dequant_word<T>(wa[u], sa2, za2, oa);
        if (RO == 2) dequant_word<T>(wb[u], sb2, zb2, ob);
      } else if (MODE == 1) {
        unpack_word<T>(wa[u], oa);
        if (RO == 2) unpack_word<T>(wb[u], ob);
      } else {
        unpack_word_biased<T>(wa[u], oa);
        if (RO == 2) unpack_word_biased<T>(wb[u], ob);
      }
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        // staged order: xq[0] = (p0,p4,p1,p5), xq[1] = (p2,p6,p3,p7), xq[2..3] the same + 8
        const uint4 lo = xq[t][u >> 1], hi = xq[t][2 + (u >> 1)];
        const uint32_t b00 = (u & 1) ? lo.z : lo.x, b01 = (u & 1) ? lo.w : lo.y;
        const uint32_t b10 = (u & 1) ? hi.z : hi.x, b11 = (u & 1) ? hi.w : hi.y;
        if (MODE == 0) {
          mma_16816<T>(acc[t], oa[0], ob[0], oa[1], ob[1], b00, b01, acc[t]);
          mma_16816<T>(acc[t], oa[2], ob[2], oa[3], ob[3], b10, b11, acc[t]);
        } else {
          if (u == 0)
            mma_16816<T>(part[t], oa[0], ob[0], oa[1], ob[1], b00, b01, zero4);
          else
            mma_16816<T>(part[t], oa[0], ob[0], oa[1], ob[1], b00, b01, part[t]);
          mma_16816<T>(part[t], oa[2], ob[2], oa[3], ob[3], b10, b11, part[t]);
        }
      }
    }
    if (MODE) {
      // part[t][0..1]: (row g of octet 0, tokens 2tig, 2tig+1); part[t][2..3]: row g of octet 1
      const float sa = bits16_to_float((uint16_t)s_cur[0], kBf16), za = bits16_to_float((uint16_t)z_cur[0], kBf16);
      const float sb = bits16_to_float((uint16_t)s_cur[RO - 1], kBf16), zb = bits16_to_float((uint16_t)z_cur[RO - 1], kBf16);
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        const float2 x0 = *sump[t][0], x1 = *sump[t][1];  // {X, C} of tokens 8t + 2tig, + 1 (zeros if absent)
        sump[t][0] += kStreamWarps;
        sump[t][1] += kStreamWarps;
        if (MODE == 2) {
          acc[t][0] += sa * (part[t][0] - x0.y) + za * x0.x;
          acc[t][1] += sa * (part[t][1] - x1.y) + za * x1.x;
          acc[t][2] += sb * (part[t][2] - x0.y) + zb * x0.x;
          acc[t][3] += sb * (part[t][3] - x1.y) + zb * x1.x;
        } else {
          acc[t][0] += sa * part[t][0] + za * x0.x;
          acc[t][1] += sa * part[t][1] + za * x1.x;
          acc[t][2] += sb * part[t][2] + zb * x0.x;
          acc[t][3] += sb * part[t][3] + zb * x1.x;
        }
      }
    }
#pragma unroll
    for (int ro = 0; ro < RO; ++ro) {
      s_cur[ro] = s_nxt[ro];
      z_cur[ro] = z_nxt[ro];
    }
  }

  // acc[t][0..1]: (channel g, tokens 8t + 2tig, +1); acc[t][2..3]: channel 8 + g
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    float* r = red + (warp * TT + t) * 128;
    *reinterpret_cast<float2*>(r + g * 8 + 2 * tig) = make_float2(acc[t][0], acc[t][1]);
    *reinterpret_cast<float2*>(r + (g + 8) * 8 + 2 * tig) = make_float2(acc[t][2], acc[t][3]);
  }
  __syncthreads();
  for (int e = tid; e < TT * 128; e += kStreamThreads) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kStreamWarps; ++w) v += red[w * TT * 128 + e];
    if (S == 1) {
      const int t = e >> 7, row = (e >> 3) & 15, tok = 8 * t + (e & 7);
      if (tok < M && row < R) y[(size_t)tok * N + n0 + row] = from_float<T>(v);
    } else {
      cpart[e] = v;
    }
  }
  if (S > 1) {
    cluster_sync_all();
    if (rank == 0) {
      for (int e = tid; e < TT * 128; e += kStreamThreads) {
        const int t = e >> 7, row = (e >> 3) & 15, tok = 8 * t + (e & 7);
        if (tok < M && row < R) {
          float v = 0.f;
          const uint32_t a = smem_u32(&cpart[e]);
          for (int r = 0; r < S; ++r) v += ld_cluster_f32(map_to_rank(a, (uint32_t)r));
          y[(size_t)tok * N + n0 + row] = from_float<T>(v);
        }
      }
    }
    cluster_sync_all();  // keep every CTA's shared memory alive until the leader has read it
  }
}

// ------------------------------------------------------------------------------------ host
static int pick_splits(int K, int kc_target, int kc_env) {
  const int groups = K / kGroup;
  if (kc_env > 0 && kc_env % kGroup == 0 && K % kc_env == 0) {
    const int s = K / kc_env;
    if (s == 1 || s == 2 || s == 4 || s == 8) return s;
  }
  int best = 1;
  for (int s = 1; s <= 8; s *= 2) {  // cluster sizes: powers of two, <= 8 (portable limit)
    if (groups % s) continue;
    best = s;
    if (K / s <= kc_target) break;
  }
  return best;
}

template <typename T, int RO, int TT, int MODE>
static int launch_stream_t(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K,
                           int S, bool pdl, cudaStream_t stream) {
  const int Kc = K / S;
  const size_t smem = (size_t)2 * RO * Kc * 2 + (size_t)M * (Kc * 2 + 16) + (size_t)(kStreamWarps + 1) * TT * 128 * 4 +
                      (MODE ? (size_t)8 * TT * (Kc / kGroup) * 8 : 0) + 64 + 16;
  if (smem > 200 * 1024) return B200AWQ_ERR_SHAPE;
  auto kern = w4a16_stream_kernel<T, RO, TT, MODE>;
  static bool attr_set = false;  // per instantiation
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(N / (8 * RO)) * S);
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = smem;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const T*)x, (const uint16_t*)qw, (const T*)sc, (const T*)sz, (T*)y, M, N,
                                     K, Kc, S);
  return e == cudaSuccess ? 0 : (int)e;
}

template <typename T, int MODE>
static int launch_stream_m(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K,
                           int S, int ro, bool pdl, cudaStream_t stream) {
  if (M <= 8) {
    if (ro == 2) return launch_stream_t<T, 2, 1, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
    return launch_stream_t<T, 1, 1, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
  }
  if (ro == 2) return launch_stream_t<T, 2, 2, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
  return launch_stream_t<T, 1, 2, MODE>(x, qw, sc, sz, y, M, N, K, S, pdl, stream);
}

int launch_stream(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                  bool pdl, const StreamTuning& tune, cudaStream_t stream) {
  if (M < 1 || M > 16 || N % 8 || K % kGroup) return B200AWQ_ERR_SHAPE;
  const int ro = (N % 16 == 0) ? 2 : 1;
  const int kc_target = M <= 2 ? 4096 : (M <= 8 ? 2048 : 1024);
  const int S = pick_splits(K, kc_target, tune.kc);
  const int mode = (tune.mode >= 0 && tune.mode <= 2) ? tune.mode : 0;
  if (dtype == B200AWQ_DTYPE_F16) {
    if (mode == 0) return launch_stream_m<__half, 0>(x, qw, sc, sz, y, M, N, K, S, ro, pdl, stream);
    if (mode == 1) return launch_stream_m<__half, 1>(x, qw, sc, sz, y, M, N, K, S, ro, pdl, stream);
    return launch_stream_m<__half, 2>(x, qw, sc, sz, y, M, N, K, S, ro, pdl, stream);
  }
  if (mode == 0) return launch_stream_m<__nv_bfloat16, 0>(x, qw, sc, sz, y, M, N, K, S, ro, pdl, stream);
  if (mode == 1) return launch_stream_m<__nv_bfloat16, 1>(x, qw, sc, sz, y, M, N, K, S, ro, pdl, stream);
  return launch_stream_m<__nv_bfloat16, 2>(x, qw, sc, sz, y, M, N, K, S, ro, pdl, stream);
}

}  // namespace b200awq
