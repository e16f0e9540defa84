// In-place NeoX-style RoPE from a cos|sin table: the per-work-item body, shared by the kernel (rope_neox.cu) and by
// the host harness of tests/test_rope_neox_host.py (which compiles THIS header with g++ and walks the same items on
// the CPU to check the indexing; the product has no CPU path).
//
// Replaces rotary_embedding_neox (awq/kernels/csrc/position_embedding/pos_encoding_kernels.cu:13-88, pybind.cpp:24,
// called from tinychat/modules/fused_attn.py:61-79).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#if defined(__CUDACC__)
#define B200AWQ_HD __host__ __device__ __forceinline__
#else
#define B200AWQ_HD inline
#endif

namespace b200awq {

struct RopeNeoxArgs {
  const long long* positions;  // [tokens]
  void* query;                 // [tokens, q_heads, head_size], token stride q_stride (elements)
  void* key;                   // [tokens, k_heads, head_size], token stride k_stride; may be null with k_heads == 0
  const void* cos_sin;         // [max_position, rot_dim]: cos of the rot_dim / 2 pair angles, then their sin
  int tokens, q_heads, k_heads, head_size, rot_dim;
  long long q_stride, k_stride;
};

B200AWQ_HD float neox_to_float(__half v) { return __half2float(v); }
B200AWQ_HD float neox_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
B200AWQ_HD void neox_from_float(float f, __half& o) { o = __float2half_rn(f); }
B200AWQ_HD void neox_from_float(float f, __nv_bfloat16& o) { o = __float2bfloat16_rn(f); }

template <typename T, int V>
struct alignas(sizeof(T) * V) NeoxVec {
  T v[V];
};

// Work items of one token: (head over q_heads + k_heads, V adjacent pairs of the rot_dim / 2).  Pair r of a head is
// (x, y) = (head[r], head[rot_dim / 2 + r]) with angle row[r] (cos) / row[rot_dim / 2 + r] (sin), the reference's
// :33-52; products in fp32, ONE rounding per output (the reference multiplies and adds in T).
template <int V>
B200AWQ_HD int rope_neox_items(const RopeNeoxArgs& a) {
  return (a.q_heads + a.k_heads) * (a.rot_dim / 2 / V);
}

template <typename T, int V>
B200AWQ_HD void rope_neox_item(const RopeNeoxArgs& a, int token, int item) {
  using Vec = NeoxVec<T, V>;
  const int embed = a.rot_dim / 2, per_head = embed / V;
  const int head = item / per_head, r = (item % per_head) * V;
  T* base = head < a.q_heads
                ? static_cast<T*>(a.query) + (long long)token * a.q_stride + (long long)head * a.head_size
                : static_cast<T*>(a.key) + (long long)token * a.k_stride + (long long)(head - a.q_heads) * a.head_size;
  const T* row = static_cast<const T*>(a.cos_sin) + a.positions[token] * (long long)a.rot_dim;
  const Vec c = *reinterpret_cast<const Vec*>(row + r), s = *reinterpret_cast<const Vec*>(row + embed + r);
  Vec x = *reinterpret_cast<const Vec*>(base + r), y = *reinterpret_cast<const Vec*>(base + embed + r);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float xf = neox_to_float(x.v[i]), yf = neox_to_float(y.v[i]);
    const float cf = neox_to_float(c.v[i]), sf = neox_to_float(s.v[i]);
    neox_from_float(xf * cf - yf * sf, x.v[i]);
    neox_from_float(yf * cf + xf * sf, y.v[i]);
  }
  *reinterpret_cast<Vec*>(base + r) = x;
  *reinterpret_cast<Vec*>(base + embed + r) = y;
}

// 8 pairs per item (16-byte accesses) when every address involved is 16-byte aligned, else one pair per item.
inline bool rope_neox_can_vectorise(const RopeNeoxArgs& a) {
  const auto al = [](const void* p) { return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0; };
  return (a.rot_dim / 2) % 8 == 0 && a.head_size % 8 == 0 && a.q_stride % 8 == 0 && (a.k_heads == 0 || a.k_stride % 8 == 0) &&
         al(a.query) && (a.k_heads == 0 || al(a.key)) && al(a.cos_sin);
}

}  // namespace b200awq
