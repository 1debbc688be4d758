// nn.Dense (+ silu) forward / backward for the per-node embedding / readout layers of
// DenseSAKEModel (sake/models.py:24-31,57,60).  O(N) work; CUDA-core kernels.
#include "common.cuh"

namespace sake {

static constexpr int DR = 16;   // rows per CTA, forward
static constexpr int DRB = 16;  // rows per CTA, backward (64 measured slower: the kernel is bound by the serial work of a CTA, not by its atomics)

__global__ void __launch_bounds__(256) k_dense_fwd(long long rows, int in, int out, int act,
                                                   const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, float* __restrict__ y,
                                                   const RaggedHdr* hdr) {
  extern __shared__ float sm[];
  if (hdr) rows = hdr->R;                       // ragged batches: real rows (the grid covers the worst case)
  const long long r0 = (long long)blockIdx.x * DR;
  if (r0 >= rows) return;
  const int nn = (int)min((long long)DR, rows - r0);
  for (int t = threadIdx.x; t < DR * in; t += blockDim.x) sm[t] = (t / in) < nn ? x[r0 * in + t] : 0.f;
  __syncthreads();
  for (int o = threadIdx.x; o < out; o += blockDim.x) {
    float acc[DR];
    const float bias = b ? b[o] : 0.f;
#pragma unroll
    for (int n = 0; n < DR; ++n) acc[n] = bias;
    for (int f = 0; f < in; ++f) {
      const float wv = w[(size_t)f * out + o];
#pragma unroll
      for (int n = 0; n < DR; ++n) acc[n] = fmaf(sm[n * in + f], wv, acc[n]);
    }
    for (int n = 0; n < nn; ++n) y[(r0 + n) * out + o] = act == 1 ? siluf_(acc[n]) : acc[n];
  }
}

__global__ void __launch_bounds__(256) k_dense_bwd(long long rows, int in, int out, int act,
                                                   const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, const float* __restrict__ dy,
                                                   float* __restrict__ dx, float* __restrict__ dw,
                                                   float* __restrict__ db, const RaggedHdr* hdr) {
  extern __shared__ float sm[];
  if (hdr) rows = hdr->R;
  if ((long long)blockIdx.x * DRB >= rows) return;
  float* xs = sm;             // [DRB][in]
  float* gs = xs + DRB * in;   // [DRB][out]  cotangent of the pre-activation
  float* ws = gs + DRB * out;  // [in][out + 1]  weights, row stride padded: both access patterns are conflict-free
  const int ldw = out + 1;
  const long long r0 = (long long)blockIdx.x * DRB;
  const int nn = (int)min((long long)DRB, rows - r0);
  for (int t = threadIdx.x; t < DRB * in; t += blockDim.x) xs[t] = (t / in) < nn ? x[r0 * in + t] : 0.f;
  for (int t = threadIdx.x; t < in * out; t += blockDim.x) ws[(t / out) * ldw + (t % out)] = w[t];
  __syncthreads();
  for (int t = threadIdx.x; t < DRB * out; t += blockDim.x) {
    const int n = t / out, o = t % out;
    float gv = 0.f;
    if (n < nn) {
      gv = dy[r0 * out + t];
      if (act == 1) {
        float z = b ? b[o] : 0.f;
        for (int f = 0; f < in; ++f) z = fmaf(xs[n * in + f], ws[f * ldw + o], z);
        gv *= dsiluf_(z);
      }
    }
    gs[t] = gv;
  }
  __syncthreads();
  if (dx) {
    for (int t = threadIdx.x; t < nn * in; t += blockDim.x) {
      const int n = t / in, f = t % in;
      float acc = 0.f;
      const float* wr = ws + f * ldw;
      for (int o = 0; o < out; ++o) acc = fmaf(wr[o], gs[n * out + o], acc);
      dx[r0 * in + t] = acc;
    }
  }
  if (dw) {
    for (int t = threadIdx.x; t < in * out; t += blockDim.x) {
      const int f = t / out, o = t % out;
      float acc = 0.f;
      for (int n = 0; n < nn; ++n) acc = fmaf(xs[n * in + f], gs[n * out + o], acc);
      atomicAdd(dw + t, acc);
    }
  }
  if (db) {
    for (int o = threadIdx.x; o < out; o += blockDim.x) {
      float acc = 0.f;
      for (int n = 0; n < nn; ++n) acc += gs[n * out + o];
      atomicAdd(db + o, acc);
    }
  }
}

int dense_fwd(long long rows, int in, int out, int act, const float* x, const float* w, const float* b, float* y,
              const RaggedHdr* hdr, cudaStream_t st) {
  if (rows == 0) return 0;
  size_t smem = sizeof(float) * DR * in;
  if (smem > 227 * 1024) { set_error("dense: in_features %d too large", in); return SAKE_EUNSUPPORTED; }
  static unsigned long long optin = 0;
  if (smem > 48 * 1024) { const int rc = smem_optin(k_dense_fwd, 227 * 1024, optin); if (rc) return rc; }
  ProfScope prof(14, rows, st);
  k_dense_fwd<<<(unsigned)((rows + DR - 1) / DR), 256, smem, st>>>(rows, in, out, act, x, w, b, y, hdr);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int dense_bwd(long long rows, int in, int out, int act, const float* x, const float* w, const float* b,
              const float* dy, float* dx, float* dw, float* db, const RaggedHdr* hdr, cudaStream_t st) {
  if (rows == 0) return 0;
  size_t smem = sizeof(float) * (DRB * (in + out) + (size_t)in * (out + 1));
  if (smem > 227 * 1024) { set_error("dense: features %d/%d too large", in, out); return SAKE_EUNSUPPORTED; }
  static unsigned long long optin = 0;   // H = 128 stages 66 KB of weights: opt in beyond the 48 KB default
  if (smem > 48 * 1024) { const int rc = smem_optin(k_dense_bwd, 227 * 1024, optin); if (rc) return rc; }
  ProfScope prof(14, rows, st);
  k_dense_bwd<<<(unsigned)((rows + DRB - 1) / DRB), 256, smem, st>>>(rows, in, out, act, x, w, b, dy, dx, dw, db, hdr);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sake
