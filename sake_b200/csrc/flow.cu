// Coupling glue of the augmented normalising flow (sake/flows.py:12-27,118-142): everything around the
// DenseSAKEModel call of one AugmentedFlowLayer, as three small kernels instead of ~10 tensor ops per layer
// (a flow pass makes 2*depth such calls).
//   k_flow_pre     h <- [h | |pos|^2], append the zero dummy atom to h and pos               flows.py:118-122
//   k_flow_post    translation = (x_out - pos)[:N] - mean; scale = mean_i tanh(scale_mlp(y_i));
//                  other <- exp(scale) other + translation  (f_forward, flows.py:131-135)
//                  other <- exp(-scale) (other - translation) (f_backward, flows.py:137-142); log_det += scale*N*D
//   k_flow_logprob -log p(x) - log p(v) + sum_log_det with the centred Gaussian prior         flows.py:13-21,
//                                                                                scripts/lj13_aug/run.py:39-43
// Coordinates are 3 wide (2-D systems carry z = 0, scripts/dw4/run.py:17-19); D only enters the log-det / prior.
#include "common.cuh"

namespace sake {

__global__ void __launch_bounds__(256) k_flow_pre(int B, int N, int Fh, const float* __restrict__ h,
                                                  const float* __restrict__ pos, float* __restrict__ h_aug,
                                                  float* __restrict__ x_aug) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * (N + 1)) return;
  const int b = (int)(t / (N + 1)), i = (int)(t - (long long)b * (N + 1));
  float* ho = h_aug + t * (Fh + 1);
  float* xo = x_aug + t * 3;
  if (i == N) {                                           // the dummy origin atom
    for (int f = 0; f <= Fh; ++f) ho[f] = 0.f;
    xo[0] = 0.f; xo[1] = 0.f; xo[2] = 0.f;
    return;
  }
  const float* p = pos + ((size_t)b * N + i) * 3;
  const float p0 = p[0], p1 = p[1], p2 = p[2];
  for (int f = 0; f < Fh; ++f) ho[f] = h ? h[((size_t)b * N + i) * Fh + f] : 0.f;
  ho[Fh] = p0 * p0 + p1 * p1 + p2 * p2;
  xo[0] = p0; xo[1] = p1; xo[2] = p2;
}

// one CTA (64 threads = hidden units of scale_mlp) per molecule
__global__ void __launch_bounds__(64) k_flow_post(int N, int D, int Hs, int direction, const float* __restrict__ x_out,
                                                  const float* __restrict__ pos0, const float* __restrict__ y,
                                                  const float* __restrict__ W0, const float* __restrict__ b0,
                                                  const float* __restrict__ W2, float* __restrict__ other,
                                                  float* __restrict__ logdet) {
  __shared__ float red[2];
  __shared__ float sh_scale, sh_mean[3];
  const int b = blockIdx.x, f = threadIdx.x, lane = f & 31, warp = f >> 5;
  // scale = mean_i tanh( sum_f silu(y_i W0[f] + b0[f]) W2[f] )      (y has one feature: out_features = 1)
  float w0 = 0.f, bb = 0.f, w2 = 0.f;
  if (f < Hs) { w0 = W0[f]; bb = b0 ? b0[f] : 0.f; w2 = W2[f]; }
  float ssum = 0.f;
  for (int i = 0; i < N; ++i) {
    const float yi = y[(size_t)b * (N + 1) + i];
    float term = f < Hs ? siluf_(fmaf(yi, w0, bb)) * w2 : 0.f;
    for (int o = 16; o; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if (lane == 0) red[warp] = term;
    __syncthreads();
    if (f == 0) ssum += tanhf(red[0] + red[1]);
    __syncthreads();
  }
  if (f < 3) {
    float m = 0.f;
    for (int i = 0; i < N; ++i) m += x_out[((size_t)b * (N + 1) + i) * 3 + f] - pos0[((size_t)b * N + i) * 3 + f];
    sh_mean[f] = m / (float)N;
  }
  if (f == 0) {
    const float sc = ssum / (float)N;
    sh_scale = sc;
    logdet[b] += sc * (float)(N * D);
  }
  __syncthreads();
  const float sc = sh_scale;
  const float es = expf(direction > 0 ? sc : -sc);
  for (int t = f; t < N * 3; t += 64) {
    const int i = t / 3, d = t - 3 * i;
    const float tr = x_out[((size_t)b * (N + 1) + i) * 3 + d] - pos0[((size_t)b * N + i) * 3 + d] - sh_mean[d];
    float* o = other + (size_t)b * N * 3 + t;
    *o = direction > 0 ? fmaf(es, *o, tr) : es * (*o - tr);
  }
}

__global__ void __launch_bounds__(128) k_flow_logprob(int B, int N, int D, const float* __restrict__ x,
                                                      const float* __restrict__ v, const float* __restrict__ logdet,
                                                      float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float r2 = 0.f;
  for (int t = 0; t < N * 3; ++t) {
    const float a = x[(size_t)b * N * 3 + t], c = v[(size_t)b * N * 3 + t];
    r2 = fmaf(a, a, fmaf(c, c, r2));
  }
  const float cst = 0.5f * (float)((N - 1) * D) * 1.8378770664093453f;     // log(2 pi)
  out[b] = 0.5f * r2 + 2.0f * cst + (logdet ? logdet[b] : 0.f);            // -log p(x) - log p(v) + sum_log_det
}

}  // namespace sake

using namespace sake;

extern "C" {

int sake_flow_pre(int32_t B, int32_t N, int32_t h_features, const float* h, const float* pos, float* h_aug, float* x_aug,
                  sake_stream_t stream) {
  if (B < 0 || N <= 0 || h_features < 0 || !pos || !h_aug || !x_aug) { set_error("sake_flow_pre: bad argument"); return SAKE_EINVAL; }
  if (B == 0) return 0;
  const long long n = (long long)B * (N + 1);
  k_flow_pre<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, N, h_features, h, pos, h_aug, x_aug);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sake_flow_post(int32_t B, int32_t N, int32_t D, int32_t scale_hidden, int32_t direction, const float* x_aug_out,
                   const float* pos0, const float* y, const float* scale0_kernel, const float* scale0_bias,
                   const float* scale2_kernel, float* other, float* logdet, sake_stream_t stream) {
  if (B < 0 || N <= 0 || D <= 0 || D > 3 || scale_hidden <= 0 || scale_hidden > 64 || (direction != 1 && direction != -1) ||
      !x_aug_out || !pos0 || !y || !scale0_kernel || !scale2_kernel || !other || !logdet) {
    set_error("sake_flow_post: bad argument (scale_mlp hidden width must be <= 64)");
    return SAKE_EINVAL;
  }
  if (B == 0) return 0;
  k_flow_post<<<B, 64, 0, (cudaStream_t)stream>>>(N, D, scale_hidden, direction, x_aug_out, pos0, y, scale0_kernel,
                                                  scale0_bias, scale2_kernel, other, logdet);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sake_flow_logprob(int32_t B, int32_t N, int32_t D, const float* x, const float* v, const float* logdet, float* out,
                      sake_stream_t stream) {
  if (B < 0 || N <= 0 || D <= 0 || !x || !v || !out) { set_error("sake_flow_logprob: bad argument"); return SAKE_EINVAL; }
  if (B == 0) return 0;
  k_flow_logprob<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, N, D, x, v, logdet, out);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
