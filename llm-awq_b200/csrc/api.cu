// C ABI of libb200awq.so (declared in include/b200awq.h): argument validation, dispatch
// between the HBM-bound streaming kernel and the tcgen05 tensor-core kernel, process-wide
// switches.  No torch types, no allocation, no synchronisation.
#include <atomic>
#include <cstdint>
#include <cstdlib>

#include "w4a16_kernels.h"

namespace {

std::atomic<int> g_pdl{-1};  // -1 = read B200AWQ_PDL on first use
std::atomic<unsigned long long> g_launches{0};

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    v = env_int("B200AWQ_PDL", 1) ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

bool aligned16(const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_device() {
  static signed char ok[32] = {};  // per device: 0 unknown, 1 ok, -1 wrong device
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B200AWQ_ERR_DEVICE;
  dev &= 31;
  if (ok[dev] == 0) {
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok[dev] = (major == 10) ? 1 : -1;
  }
  return ok[dev] == 1 ? 0 : B200AWQ_ERR_DEVICE;
}

int check_common(const void* x, const void* qw, const void* sc, const void* sz, void* y, int m, int n, int k,
                 int group_size, int dtype) {
  if (dtype != B200AWQ_DTYPE_F16 && dtype != B200AWQ_DTYPE_BF16) return B200AWQ_ERR_DTYPE;
  if (group_size != 128) return B200AWQ_ERR_GROUP;
  if (m < 1 || n < 8 || k < 128 || (n % 8) || (k % 128)) return B200AWQ_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(qw) || !aligned16(sc) || !aligned16(sz) || !aligned16(y)) return B200AWQ_ERR_ALIGN;
  return check_device();
}

b200awq::StreamTuning stream_tuning() {
  b200awq::StreamTuning t;
  t.mode = env_int("B200AWQ_STREAM_MODE", -1);
  t.kc = env_int("B200AWQ_STREAM_KC", 0);
  t.rpb = env_int("B200AWQ_STREAM_RPB", 0);
  t.pad = env_int("B200AWQ_STREAM_PAD", 0);
  t.rbs = env_int("B200AWQ_STREAM_RBS", 0);
  t.warps = env_int("B200AWQ_STREAM_WARPS", 0);
  return t;
}
b200awq::FlatTuning flat_tuning() {
  b200awq::FlatTuning t;
  t.kc = env_int("B200AWQ_FLAT_KC", 0);
  return t;
}
// which kernel serves small token counts: B200AWQ_SKINNY = "flat" (tcgen05, default) | "stream" (mma.sync)
bool use_flat(int m, int n, int k) {
  static int mode = -1;
  if (mode < 0) {
    const char* v = std::getenv("B200AWQ_SKINNY");
    mode = (v && v[0] == 's') ? 0 : 1;
  }
  // measured crossovers (profiles/): the mma.sync streaming kernel wins for m <= 4 (m <= 2 when k is long enough to
  // need a cluster split: 3 <= m <= 4 at k = 14336 is 10 % faster on the tcgen05 kernel), the tcgen05 kernels above
  const int min_m = env_int("B200AWQ_FLAT_MIN_M", k > 8192 ? 3 : 5);
  return mode == 1 && n % 128 == 0 && m >= min_m && m <= env_int("B200AWQ_FLAT_MAX_M", 16);
}
b200awq::UmmaTuning umma_tuning() {
  b200awq::UmmaTuning t;
  t.tn = env_int("B200AWQ_UMMA_TN", 0);
  t.max_ctas = env_int("B200AWQ_UMMA_CTAS", 0);
  t.split = env_int("B200AWQ_UMMA_SPLIT", 0);
  return t;
}

}  // namespace

extern "C" {

int b200awq_w4a16_gemv(const void* x, const void* qweight, const void* scales, const void* szeros, void* y, int m, int n,
                       int k, int group_size, int dtype, void* stream) {
  if (int e = check_common(x, qweight, scales, szeros, y, m, n, k, group_size, dtype)) return e;
  if (m > 7) return B200AWQ_ERR_BATCH;  // reference envelope: gemv_cuda.cu:291-330
  int r = B200AWQ_ERR_SHAPE;
  if (use_flat(m, n, k))
    r = b200awq::launch_flat(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), flat_tuning(),
                             static_cast<cudaStream_t>(stream));
  if (r == B200AWQ_ERR_SHAPE)
    r = b200awq::launch_stream(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), stream_tuning(),
                               static_cast<cudaStream_t>(stream));
  if (r == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
  return r;
}

int b200awq_w4a16_gemv_allreduce(const void* x, const void* qweight, const void* scales, const void* szeros, void* y, int m,
                                 int n, int k, int group_size, int dtype, const b200awq_peers* peers, void* stream) {
  if (int e = check_common(x, qweight, scales, szeros, y, m, n, k, group_size, dtype)) return e;
  if (m > 8) return B200AWQ_ERR_BATCH;
  if (!peers || peers->world < 1 || peers->world > 8 || peers->rank < 0 || peers->rank >= peers->world || !peers->epoch ||
      (long long)peers->cap_words < (long long)m * n)
    return B200AWQ_ERR_PEERS;
  b200awq::PeerArgs pa{};
  for (int r = 0; r < peers->world; ++r) {
    if (!peers->data[r]) return B200AWQ_ERR_PEERS;
    pa.data[r] = static_cast<unsigned long long*>(peers->data[r]);
  }
  pa.epoch = static_cast<unsigned int*>(peers->epoch);
  pa.rank = peers->rank;
  pa.world = peers->world;
  pa.cap = peers->cap_words;
  int r = b200awq::launch_stream(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), stream_tuning(),
                                 static_cast<cudaStream_t>(stream), &pa);
  if (r == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
  return r;
}

size_t b200awq_w4a16_gemm_workspace_bytes(int, int, int) { return 0; }

int b200awq_w4a16_gemm(const void* x, const void* qweight, const void* scales, const void* szeros, void* y, int m, int n,
                       int k, int group_size, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  if (int e = check_common(x, qweight, scales, szeros, y, m, n, k, group_size, dtype)) return e;
  if (n % 128) return B200AWQ_ERR_SHAPE;  // reference: N / CTA_N with CTA_N = 128, gemm_cuda.cu:38,1225
  const int stream_max_m = env_int("B200AWQ_STREAM_MAX_M", 16);
  int r = B200AWQ_ERR_SHAPE;
  if (use_flat(m, n, k))
    r = b200awq::launch_flat(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), flat_tuning(),
                             static_cast<cudaStream_t>(stream));
  if (r == B200AWQ_ERR_SHAPE && m <= stream_max_m && m <= 16)
    r = b200awq::launch_stream(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), stream_tuning(),
                               static_cast<cudaStream_t>(stream));
  if (r == B200AWQ_ERR_SHAPE && env_int("B200AWQ_UMMA2", 0) == 1 && n % 256 == 0) {
    // opt-in: second-generation prefill kernel (256 channels x 128 tokens per CTA, w4a16_umma2.cu).  It moves a third
    // less L2 -> SM traffic per MAC but measured ~10% SLOWER than the 128 x 256 kernel (profiles/README.md).
    r = b200awq::launch_umma2(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), umma_tuning(),
                              static_cast<cudaStream_t>(stream));
  }
  if (r == B200AWQ_ERR_SHAPE)  // 128-channel tiles; split-k over a cluster for small token counts
    r = b200awq::launch_umma(x, qweight, scales, szeros, y, m, n, k, dtype, pdl_enabled(), umma_tuning(),
                             static_cast<cudaStream_t>(stream));
  if (r == 0) g_launches.fetch_add(1, std::memory_order_relaxed);
  return r;
}

int gemv_forward_4bit(const void* x, const void* qweight, const void* scales, const void* szeros, void* y, int m, int n,
                      int k, int group_size, int dtype, void* stream) {
  return b200awq_w4a16_gemv(x, qweight, scales, szeros, y, m, n, k, group_size, dtype, stream);
}
int gemm_forward_4bit(const void* x, const void* qweight, const void* scales, const void* szeros, void* y, int m, int n,
                      int k, int group_size, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  return b200awq_w4a16_gemm(x, qweight, scales, szeros, y, m, n, k, group_size, dtype, workspace, workspace_bytes, stream);
}

int b200awq_set_pdl(int enable) {
  int prev = pdl_enabled() ? 1 : 0;
  g_pdl.store(enable ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

unsigned long long b200awq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* b200awq_strerror(int code) {
  switch (code) {
    case B200AWQ_OK: return "ok";
    case B200AWQ_ERR_SHAPE: return "unsupported shape (need n % 8 == 0 [gemm: n % 128 == 0], k % 128 == 0, m >= 1)";
    case B200AWQ_ERR_GROUP: return "unsupported group size (only 128)";
    case B200AWQ_ERR_ALIGN: return "null or misaligned pointer (16-byte alignment required)";
    case B200AWQ_ERR_DTYPE: return "unsupported dtype (fp16 / bf16 only)";
    case B200AWQ_ERR_BATCH: return "unsupported batch size for gemv kernel (1..7)";
    case B200AWQ_ERR_WORKSPACE: return "workspace too small";
    case B200AWQ_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case B200AWQ_ERR_DEVICE: return "device is not compute capability 10.x (sm_100a kernels only)";
    case B200AWQ_ERR_PEERS: return "bad peer description for the fused all-reduce";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

const char* b200awq_version(void) { return "b200awq 0.1 sm_100a"; }

}  // extern "C"
