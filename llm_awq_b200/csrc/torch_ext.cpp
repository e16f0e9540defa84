// Python module `awq_inference_engine`: the reference's extension name and the two entry
// points WQLinear / tinychat call (awq/kernels/csrc/pybind.cpp:22-23), as a thin shim over
// the C ABI in include/b200awq.h.  Same positional signatures, same return tensors, same
// exception types as the reference host functions (gemv_cuda.cu:245-338,
// gemm_cuda.cu:1126-1236); additionally: launches on the CURRENT stream of the input's
// device (the reference uses the legacy default stream and no device guard), checks
// contiguity / shapes, and surfaces launch errors.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <stdexcept>

#include "../../include/b200awq.h"

namespace {

int dtype_code(const torch::Tensor& t, const char* fn) {
  if (t.scalar_type() == at::ScalarType::Half) return B200AWQ_DTYPE_F16;
  if (t.scalar_type() == at::ScalarType::BFloat16) return B200AWQ_DTYPE_BF16;
  // reference: dispatch_utils.cuh:13-17
  TORCH_CHECK(false, fn, " failed to dispatch data type ", t.scalar_type());
}

void check_inputs(const torch::Tensor& in, const torch::Tensor& kernel, const torch::Tensor& scales,
                  const torch::Tensor& zeros) {
  TORCH_CHECK(in.is_cuda() && kernel.is_cuda() && scales.is_cuda() && zeros.is_cuda(), "all tensors must be CUDA tensors");
  TORCH_CHECK(scales.scalar_type() == in.scalar_type());  // reference: gemv_cuda.cu:260, gemm_cuda.cu:1145
  TORCH_CHECK(zeros.scalar_type() == in.scalar_type());   // reference: gemv_cuda.cu:261, gemm_cuda.cu:1146
  TORCH_CHECK(kernel.scalar_type() == at::ScalarType::Short, "qweight must be int16 [N/4, K]");
  TORCH_CHECK(in.is_contiguous() && kernel.is_contiguous() && scales.is_contiguous() && zeros.is_contiguous(),
              "inputs must be contiguous");
  TORCH_CHECK(kernel.dim() == 2 && scales.dim() == 2 && zeros.dim() == 2 && in.dim() >= 1, "bad tensor ranks");
}

void raise(int code, bool is_gemv) {
  if (code == 0) return;
  // the reference throws std::runtime_error for unsupported batch / group (gemv_cuda.cu:329,334)
  if (is_gemv && code == B200AWQ_ERR_BATCH) throw std::runtime_error("Unsupported batch size for gemv kernel.\n");
  if (is_gemv && code == B200AWQ_ERR_GROUP) throw std::runtime_error("Unsupported group size for gemv kernel.\n");
  TORCH_CHECK(false, "b200awq: ", b200awq_strerror(code), " (code ", code, ")");
}

}  // namespace

torch::Tensor gemv_forward_cuda_new(torch::Tensor _in_feats, torch::Tensor _kernel, torch::Tensor _scaling_factors,
                                    torch::Tensor _zeros, int m, int n, int k, int group_size) {
  check_inputs(_in_feats, _kernel, _scaling_factors, _zeros);
  const int dt = dtype_code(_in_feats, "gemv_forward_cuda_new");
  TORCH_CHECK(_in_feats.size(-1) == k && _in_feats.numel() == (int64_t)m * k, "in_feats does not match (m, k)");
  TORCH_CHECK(_kernel.size(0) * 4 == n && _kernel.size(1) == k, "qweight does not match (n, k)");
  TORCH_CHECK(_scaling_factors.size(1) == n && _zeros.size(1) == n && _scaling_factors.size(0) * group_size >= k &&
                  _zeros.size(0) * group_size >= k,
              "scales / zeros do not match (n, k, group)");
  const c10::cuda::CUDAGuard guard(_in_feats.device());
  std::vector<int64_t> output_shape = _in_feats.sizes().vec();
  output_shape.back() = n;
  at::Tensor out = torch::empty(output_shape, _in_feats.options());
  raise(b200awq_w4a16_gemv(_in_feats.data_ptr(), _kernel.data_ptr(), _scaling_factors.data_ptr(), _zeros.data_ptr(),
                           out.data_ptr(), m, n, k, group_size, dt, at::cuda::getCurrentCUDAStream().stream()),
        true);
  return out;
}

torch::Tensor gemm_forward_cuda_new(torch::Tensor _in_feats, torch::Tensor _kernel, torch::Tensor _scales,
                                    torch::Tensor _zeros) {
  check_inputs(_in_feats, _kernel, _scales, _zeros);
  const int dt = dtype_code(_in_feats, "gemm_forward_cuda_new");
  const int64_t k = _in_feats.size(-1);
  const int64_t n = _kernel.size(0) * 4;  // reference: gemm_cuda.cu:1133
  const int64_t m = _in_feats.numel() / k;
  TORCH_CHECK(_kernel.size(1) == k, "qweight does not match in_feats");
  TORCH_CHECK(_scales.size(1) == n && _zeros.size(1) == n && _scales.size(0) * 128 >= k && _zeros.size(0) * 128 >= k,
              "scales / zeros do not match (n, k)");
  const c10::cuda::CUDAGuard guard(_in_feats.device());
  std::vector<int64_t> output_shape = _in_feats.sizes().vec();
  output_shape.back() = n;
  at::Tensor out = torch::empty(output_shape, _in_feats.options());
  if (m == 0) return out;
  const size_t ws_bytes = b200awq_w4a16_gemm_workspace_bytes((int)m, (int)n, (int)k);
  at::Tensor ws;
  if (ws_bytes) ws = torch::empty({(int64_t)ws_bytes}, _in_feats.options().dtype(torch::kUInt8));
  raise(b200awq_w4a16_gemm(_in_feats.data_ptr(), _kernel.data_ptr(), _scales.data_ptr(), _zeros.data_ptr(), out.data_ptr(),
                           (int)m, (int)n, (int)k, 128, dt, ws_bytes ? ws.data_ptr() : nullptr, ws_bytes,
                           at::cuda::getCurrentCUDAStream().stream()),
        false);
  return out;
}

// New entry (no counterpart in the reference's extension): the whole of QuantLlamaMLP.our_llama_mlp
// (tinychat/modules/fused_mlp.py:36-83) in one call.  The zeros tensors are used as given.
torch::Tensor mlp_front_forward_cuda(torch::Tensor _in_feats, torch::Tensor _gate_kernel, torch::Tensor _gate_scales,
                                     torch::Tensor _gate_zeros, torch::Tensor _up_kernel, torch::Tensor _up_scales,
                                     torch::Tensor _up_zeros) {
  check_inputs(_in_feats, _gate_kernel, _gate_scales, _gate_zeros);
  check_inputs(_in_feats, _up_kernel, _up_scales, _up_zeros);
  const int dt = dtype_code(_in_feats, "mlp_front_forward_cuda");
  const int64_t k = _in_feats.size(-1);
  const int64_t n = _gate_kernel.size(0) * 4;
  const int64_t m = _in_feats.numel() / k;
  TORCH_CHECK(_gate_kernel.size(1) == k && _up_kernel.size(0) * 4 == n && _up_kernel.size(1) == k, "qweights do not match");
  for (const torch::Tensor* t : {&_gate_scales, &_gate_zeros, &_up_scales, &_up_zeros})
    TORCH_CHECK(t->size(1) == n && t->size(0) * 128 >= k, "scales / zeros do not match (n, k)");
  const c10::cuda::CUDAGuard guard(_in_feats.device());
  std::vector<int64_t> output_shape = _in_feats.sizes().vec();
  output_shape.back() = n;
  at::Tensor out = torch::empty(output_shape, _in_feats.options());
  if (m == 0) return out;
  at::Tensor ws;
  size_t ws_bytes = 0;
  if (!(m == 1 && dt == B200AWQ_DTYPE_F16)) {
    ws_bytes = b200awq_w4a16_mlp_front_workspace_bytes((int)m, (int)n, (int)k);
    ws = torch::empty({(int64_t)ws_bytes}, _in_feats.options().dtype(torch::kUInt8));
  }
  int rc = b200awq_w4a16_mlp_front(_in_feats.data_ptr(), _gate_kernel.data_ptr(), _gate_scales.data_ptr(), _gate_zeros.data_ptr(),
                                   _up_kernel.data_ptr(), _up_scales.data_ptr(), _up_zeros.data_ptr(), out.data_ptr(), (int)m,
                                   (int)n, (int)k, 128, dt, ws_bytes ? ws.data_ptr() : nullptr, ws_bytes,
                                   at::cuda::getCurrentCUDAStream().stream());
  if (rc == B200AWQ_ERR_WORKSPACE && ws_bytes == 0) {  // the one-kernel path declined the shape: composed path
    ws_bytes = b200awq_w4a16_mlp_front_workspace_bytes((int)m, (int)n, (int)k);
    ws = torch::empty({(int64_t)ws_bytes}, _in_feats.options().dtype(torch::kUInt8));
    rc = b200awq_w4a16_mlp_front(_in_feats.data_ptr(), _gate_kernel.data_ptr(), _gate_scales.data_ptr(), _gate_zeros.data_ptr(),
                                 _up_kernel.data_ptr(), _up_scales.data_ptr(), _up_zeros.data_ptr(), out.data_ptr(), (int)m, (int)n,
                                 (int)k, 128, dt, ws.data_ptr(), ws_bytes, at::cuda::getCurrentCUDAStream().stream());
  }
  raise(rc, false);
  return out;
}

// reference: awq/kernels/csrc/layernorm/layernorm.cu:111-131 (input [b, s, c], gamma [c], out like input; returns
// nothing).  Same dtype checks; additionally any rank >= 1 is accepted (rows = numel / last dim), the launch goes to
// the current stream, and contiguity is checked instead of assumed.
void layernorm_forward_cuda(torch::Tensor _input, torch::Tensor _gamma, torch::Tensor _out, float eps) {
  TORCH_CHECK(_input.is_cuda() && _gamma.is_cuda() && _out.is_cuda(), "all tensors must be CUDA tensors");
  const int dt = dtype_code(_input, "layernorm_forward_cuda");
  TORCH_CHECK(_gamma.scalar_type() == _input.scalar_type());  // layernorm.cu:122
  TORCH_CHECK(_out.scalar_type() == _input.scalar_type());    // layernorm.cu:123
  TORCH_CHECK(_input.dim() >= 1 && _input.is_contiguous() && _out.is_contiguous() && _gamma.is_contiguous(),
              "inputs must be contiguous");
  const int64_t n = _input.size(-1);
  TORCH_CHECK(_gamma.numel() == n && _out.numel() == _input.numel(), "gamma / out do not match input");
  if (_input.numel() == 0) return;
  const c10::cuda::CUDAGuard guard(_input.device());
  raise(b200awq_rmsnorm(_input.data_ptr(), _gamma.data_ptr(), _out.data_ptr(), (int)(_input.numel() / n), (int)n, eps, dt,
                        at::cuda::getCurrentCUDAStream().stream()),
        false);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "B200-native drop-in for llm-awq's awq_inference_engine (W4A16 path only)";
  m.def("gemm_forward_cuda_new", &gemm_forward_cuda_new, "New quantized GEMM kernel.");
  m.def("gemv_forward_cuda_new", &gemv_forward_cuda_new, "New quantized GEMV kernel.");
  m.def("layernorm_forward_cuda", &layernorm_forward_cuda, "FasterTransformer layernorm kernel");
  m.def("mlp_front_forward_cuda", &mlp_front_forward_cuda, "silu(x Wgate^T) * (x Wup^T) for W4A16 weights, one call");
  m.def("set_pdl", [](bool on) { return b200awq_set_pdl(on ? 1 : 0) != 0; }, "programmatic dependent launch on/off");
  m.def("launch_count", []() { return b200awq_launch_count(); }, "kernels launched by libb200awq so far");
  m.def("version", []() { return std::string(b200awq_version()); });
}
