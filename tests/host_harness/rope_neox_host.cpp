// TEST INFRASTRUCTURE (tests/test_rope_neox_host.py): walks the work items of llm_awq_b200/csrc/rope_neox.cuh on the
// CPU exactly as the kernel's grid does (token = CTA, item = thread index + k * block size), so that the indexing and
// the arithmetic of the kernel body can be checked against the oracle where there is no GPU.  Not part of the product;
// nothing under llm_awq_b200/ builds, loads or calls this file.
#include "rope_neox.cuh"

using namespace b200awq;

template <typename T, int V>
static void walk(const RopeNeoxArgs& a, int block) {
  const int items = rope_neox_items<V>(a);
  for (int token = 0; token < a.tokens; ++token)
    for (int tid = 0; tid < block; ++tid)
      for (int i = tid; i < items; i += block) rope_neox_item<T, V>(a, token, i);
}

// dtype: 0 = fp16, 1 = bf16 (include/b200awq.h).  Returns 8 or 1 = pairs per work item that were used, < 0 on error.
extern "C" int rope_neox_host(const long long* positions, void* query, void* key, const void* cos_sin, int tokens, int q_heads,
                              int k_heads, int head_size, int rot_dim, long long q_stride, long long k_stride, int dtype,
                              int force_scalar, int block) {
  const RopeNeoxArgs a{positions, query, key, cos_sin, tokens, q_heads, k_heads, head_size, rot_dim, q_stride, k_stride};
  const bool vec = !force_scalar && rope_neox_can_vectorise(a);
  if (dtype == 0) {
    if (vec) walk<__half, 8>(a, block); else walk<__half, 1>(a, block);
  } else if (dtype == 1) {
    if (vec) walk<__nv_bfloat16, 8>(a, block); else walk<__nv_bfloat16, 1>(a, block);
  } else {
    return -1;
  }
  return vec ? 8 : 1;
}
