// Optimizer step of the reference's training loop (train.py:71,154: torch.optim.Adam(params, lr) with PyTorch's defaults betas =
// (0.9, 0.999), eps = 1e-8, no weight decay, no amsgrad) for ALL parameter tensors in ONE launch: a table of
// (param, grad, exp_avg, exp_avg_sq, numel) records and a list of 64 Ki-element chunks (tensor, offset), so that the 217 tensors of
// KGnet (74 M parameters) cost one kernel instead of ~1 500 elementwise launches.  Per element, in PyTorch's operation order
// (torch/optim/adam.py, _single_tensor_adam, fp32):
//     m  = m + (1 - b1) * (g - m)                 (lerp)
//     v  = v * b2 + (1 - b2) * g * g              (mul_, addcmul_)
//     p  = p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
#include "common.cuh"

#include <cmath>

namespace kg {

struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };
struct AdamChunk { int tensor; int pad; long long start; };
constexpr int ADAM_CHUNK = 65536;

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamTensor* __restrict__ tensors, const AdamChunk* __restrict__ chunks,
                                                        float one_minus_b1, float b2, float one_minus_b2, float step_size,
                                                        float inv_bc2_sqrt, float eps) {
  const AdamChunk ck = chunks[blockIdx.x];
  const AdamTensor t = tensors[ck.tensor];
  const long long end = min(t.n, ck.start + (long long)ADAM_CHUNK);
  for (long long i = ck.start + threadIdx.x; i < end; i += 256) {
    const float g = __ldg(t.g + i);
    float m = t.m[i], v = t.v[i];
    m = __fadd_rn(m, __fmul_rn(one_minus_b1, __fsub_rn(g, m)));                           // lerp_(grad, 1 - beta1): m + w * (g - m)
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(one_minus_b2, g), g));           // mul_(beta2).addcmul_(g, g, value = 1 - beta2)
    const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), inv_bc2_sqrt), eps);           // (sqrt(v) / sqrt(bc2)).add_(eps); ATen divides by a
                                                                                          // scalar through its fp32 reciprocal
    t.m[i] = m; t.v[i] = v;
    t.p[i] = __fadd_rn(t.p[i], __fdiv_rn(__fmul_rn(-step_size, m), denom));               // addcdiv_(m, denom, value = -step_size)
  }
}

}  // namespace kg

using namespace kg;

extern "C" int kg_adam_step(const void* d_tensors, const void* d_chunks, int n_chunks, double lr, double beta1, double beta2, double eps,
                            int step, void* stream) {
  KG_REQUIRE(n_chunks >= 0 && step >= 1 && lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.,
             "kg_adam_step: bad arguments (step=%d lr=%g betas=(%g, %g) eps=%g)", step, lr, beta1, beta2, eps);
  if (n_chunks == 0) return KG_OK;
  KG_REQUIRE(d_tensors && d_chunks, "kg_adam_step: null table");
  // the hyper-parameters are Python floats (doubles) in torch/optim/adam.py: 1 - beta and the bias corrections are formed in double
  // and rounded to fp32 only where they meet the tensors (1 - 0.999f differs from float(1 - 0.999) by 1.3e-5)
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_bc2_sqrt = 1.f / (float)std::sqrt(bc2);
  adam_step_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamTensor*>(d_tensors),
                                                                         reinterpret_cast<const AdamChunk*>(d_chunks), (float)(1.0 - beta1),
                                                                         (float)beta2, (float)(1.0 - beta2), step_size, inv_bc2_sqrt, (float)eps);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}
