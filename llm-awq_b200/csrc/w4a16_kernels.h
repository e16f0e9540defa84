// Internal launcher interface between the C ABI (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/b200awq.h"

namespace b200awq {

struct StreamTuning {
  int mode = -1;  // -1 = auto (fp16: 2, bf16: 0); 0 operand-exact, 2 group-factored    [env B200AWQ_STREAM_MODE]
  int kc = 0;     // 0 = auto, else input channels per CTA (K / kc in {1,2,4,8})       [env B200AWQ_STREAM_KC]
  int rpb = 0;    // 0 = auto, else 1024-k rounds per prefetch barrier (copy = 2 rpb KB) [env B200AWQ_STREAM_RPB]
  int pad = 0;    // extra dynamic smem bytes per CTA (co-residency limiter, tuning)     [env B200AWQ_STREAM_PAD]
};

struct UmmaTuning {
  int tn = 0;      // 0 = auto, else token tile (32 / 64 / 128 / 256)   [env B200AWQ_UMMA_TN]
  int max_ctas = 0;  // 0 = one per SM                                   [env B200AWQ_UMMA_CTAS]
};

struct FlatTuning {
  int kc = 0;  // 0 = auto, else input channels per CTA (K / kc <= 8)   [env B200AWQ_FLAT_KC]
};

// HBM-bound streaming kernel, 1 <= M <= 16 (w4a16_stream.cu)
int launch_stream(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                  bool pdl, const StreamTuning& tune, cudaStream_t stream);

// tcgen05 / TMA / TMEM tensor-core kernel, any M (w4a16_umma.cu)
int launch_umma(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                bool pdl, const UmmaTuning& tune, cudaStream_t stream);

// tcgen05 skinny-batch kernel, 1 <= M <= 64, N % 128 == 0 (w4a16_flat.cu)
int launch_flat(const void* x, const void* qw, const void* sc, const void* sz, void* y, int M, int N, int K, int dtype,
                bool pdl, const FlatTuning& tune, cudaStream_t stream);

}  // namespace b200awq
