// Energy head (sum over atoms, L1 loss cotangent), the optimiser chain of the training drivers,
// and the per-launch event profiler.  scripts/md17/run.py:46-58, scripts/qm9/run.py:79-89,134-138.
#include <mutex>
#include <vector>
#include "common.cuh"

namespace sake {

__global__ void __launch_bounds__(128) k_energy_head(int B, int N, int out, int mode, const float* __restrict__ y,
                                                     const float* __restrict__ am, const float* __restrict__ target,
                                                     float mean, float std, float* __restrict__ energy,
                                                     float* __restrict__ loss, float* __restrict__ dy,
                                                     const int2* __restrict__ molinfo) {
  __shared__ float red[4];
  __shared__ float gsh;
  const int b = blockIdx.x;
  // ragged batches: y / dy are compact, molecule b owns rows [molinfo[b].x, + molinfo[b].y) and there is no mask
  const size_t base = molinfo ? (size_t)molinfo[b].x : (size_t)b * N;
  if (molinfo) N = molinfo[b].y;
  float s = 0.f;
  for (int t = threadIdx.x; t < N * out; t += blockDim.x) {
    float m = am ? am[(size_t)b * N + t / out] : 1.0f;
    s = fmaf(y[base * out + t], m, s);
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float e = red[0] + red[1] + red[2] + red[3];
    if (energy) energy[b] = e;
    float g = 1.0f;
    if (mode == 1) {
      float diff = fmaf(std, e, mean) - target[b];
      if (loss) atomicAdd(loss, fabsf(diff) / (float)B);
      g = (diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.0f)) * std / (float)B;
    }
    gsh = g;
  }
  __syncthreads();
  if (dy) {
    const float g = gsh;
    for (int t = threadIdx.x; t < N * out; t += blockDim.x) {
      float m = am ? am[(size_t)b * N + t / out] : 1.0f;
      dy[base * out + t] = g * m;
    }
  }
}

__global__ void __launch_bounds__(256) k_adam(long long n, float* __restrict__ p, const float* __restrict__ g,
                                              float* __restrict__ m, float* __restrict__ v, float lr, float b1,
                                              float b2, float eps, float wd, float max_delta, float gscale,
                                              float c1, float c2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = fmaf(wd, p[i], g[i] * gscale);                     // additive_weight_decay
    if (max_delta > 0.f) gi = fminf(fmaxf(gi, -max_delta), max_delta);   // optax.clip (element-wise)
    float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr * (mi * c1) / (sqrtf(vi * c2) + eps);
  }
}

// ---- event profiler ------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int kind; long long pairs; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static int g_prof_cap = 0;

bool prof_begin_launch(int kind, long long pairs, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if ((int)g_prof.size() >= g_prof_cap) return false;
  ProfRec r;
  r.kind = kind;
  r.pairs = pairs;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return false;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  return true;
}
void prof_end_launch(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st);
}

}  // namespace sake

using namespace sake;

extern "C" {

int sake_energy_head(int32_t B, int32_t N, int32_t out_features, int32_t mode, const float* y,
                     const float* atom_mask, const float* target, float mean, float std, float* energy,
                     float* loss, float* dy, const void* ragged, sake_stream_t stream) {
  const int2* molinfo = nullptr;
  if (ragged) {
    if (atom_mask) { set_error("sake_energy_head: ragged batches carry no atom mask"); return SAKE_EINVAL; }
    Dims d;
    d.B = B; d.N = N;
    ragged_attach(d, ragged);
    molinfo = d.molinfo;
  }
  if (B < 0 || N <= 0 || out_features <= 0 || !y || (mode == 1 && !target) || (mode != 0 && mode != 1)) {
    set_error("sake_energy_head: bad argument");
    return SAKE_EINVAL;
  }
  if (B == 0) return 0;
  k_energy_head<<<B, 128, 0, (cudaStream_t)stream>>>(B, N, out_features, mode, y, atom_mask, target, mean, std,
                                                     energy, loss, dy, molinfo);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sake_adam_step(int64_t n, float* params, const float* grads, float* m, float* v, int32_t step, float lr,
                   float b1, float b2, float eps, float weight_decay, float max_delta, float grad_scale,
                   sake_stream_t stream) {
  if (n < 0 || !params || !grads || !m || !v || step < 1) { set_error("sake_adam_step: bad argument"); return SAKE_EINVAL; }
  if (n == 0) return 0;
  const float c1 = 1.0f / (1.0f - powf(b1, (float)step));
  const float c2 = 1.0f / (1.0f - powf(b2, (float)step));
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(n, params, grads, m, v, lr, b1, b2, eps, weight_decay, max_delta,
                                                 grad_scale, c1, c2);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sake_profile_begin(int32_t capacity) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_cap = capacity > 0 ? capacity : 0;
  return 0;
}

int sake_profile_collect(float* ms, int32_t* kind, int64_t* pairs, int32_t capacity) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& r : g_prof) {
    if (n < capacity && ms) {
      float t = 0.f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
        ms[n] = t;
        if (kind) kind[n] = r.kind;
        if (pairs) pairs[n] = r.pairs;
        ++n;
      }
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_cap = 0;
  return n;
}

}  // extern "C"
