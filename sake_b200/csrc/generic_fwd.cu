// Generic fp32 CUDA-core forward of DenseSAKELayer (any H / A / K).
// Correctness-first engine and the permanent path for shapes the tcgen05 engine does not take
// (H = 4, 7, 16, 32 of the reference's tests and scripts).  Reference: sake/layers.py:188-235.
#include "common.cuh"

namespace sake {

static constexpr int NODES = SAKE_NODES;    // nodes per CTA in the per-node kernels
static constexpr int PJ = 16;      // pairs per chunk in the edge kernel
static constexpr int MJ = 8;       // pairs per chunk in the mix kernel

// ------------------------------------------------------------------------------------------
// node_pre: separable halves of mlp_in and mlp_out[0] applied per node (layers.py:30,33-38)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_node_pre(Dims d, const float* __restrict__ h,
                                                  const float* __restrict__ Win, const float* __restrict__ bin,
                                                  const float* __restrict__ W1, const float* __restrict__ b1,
                                                  float* __restrict__ proj) {
  extern __shared__ float sm[];
  float* hs = sm;  // [NODES][H]
  float* ps = sm + NODES * d.H;   // [NODES][NP] output staging (G8 layout only)
  const int r0 = blockIdx.x * NODES;
  if (r0 >= dims_rows(d)) return;                  // ragged: the grid covers the padded worst case
  const int nn = min(NODES, dims_rows(d) - r0);
  for (int t = threadIdx.x; t < nn * d.H; t += blockDim.x) hs[t] = h[(size_t)r0 * d.H + t];
  __syncthreads();
  const int H = d.H, K = d.K;
  const int Kp = d.Kp;
  for (int o = threadIdx.x; o < d.NP; o += blockDim.x) {
    const float* w = nullptr;
    int ld = 0;
    float bias = 0.f;
    if (o < Kp) { if (o < K) { w = Win + o; ld = K; } }
    else if (o < 2 * Kp) { const int k = o - Kp; if (k < K) { w = Win + (size_t)H * K + k; ld = K; bias = bin[k]; } }
    else if (o < 2 * Kp + H) { w = W1 + (o - 2 * Kp); ld = H; }
    else { w = W1 + (size_t)H * H + (o - 2 * Kp - H); ld = H; bias = b1[o - 2 * Kp - H]; }
    if (w == nullptr) {      // padding slot
      for (int n = 0; n < NODES; ++n) {
        if (d.g8) ps[n * d.NP + o] = 0.f;
        else if (n < nn) proj[(size_t)(r0 + n) * d.NP + o] = 0.f;
      }
      continue;
    }
    float acc[NODES];
#pragma unroll
    for (int n = 0; n < NODES; ++n) acc[n] = bias;
    for (int f = 0; f < H; ++f) {
      float wv = w[(size_t)f * ld];
#pragma unroll
      for (int n = 0; n < NODES; ++n) acc[n] = fmaf(hs[n * H + f], wv, acc[n]);
    }
    if (d.g8) {
#pragma unroll
      for (int n = 0; n < NODES; ++n) ps[n * d.NP + o] = acc[n];
    } else {
      for (int n = 0; n < nn; ++n) proj[(size_t)(r0 + n) * d.NP + o] = acc[n];
    }
  }
  if (d.g8) {
    // G8 layout (tcgen05 engines): the CTA's rows are two groups of 8, each one contiguous block of NP/4 units x 8 rows
    __syncthreads();
    const int U = d.NP >> 2;
    float4* p4 = reinterpret_cast<float4*>(proj) + g8_row(r0, U);           // r0 is a multiple of 8
    for (int t = threadIdx.x; t < (NODES / 8) * U * 8; t += blockDim.x) {
      const int grp = t / (U * 8), rem = t - grp * U * 8, u = rem >> 3, rr = rem & 7, n = grp * 8 + rr;
      if (n < nn) p4[t] = *reinterpret_cast<const float4*>(ps + n * d.NP + 4 * u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// edge_fwd: one CTA per receiving atom i; pair geometry, RBF, edge MLP, attention logits
// (functional.py:7-19, utils.py:61-65, layers.py:28-40,155-165)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edge_fwd(Dims d, const float* __restrict__ x,
                                                  const float* __restrict__ mask, const SakeLayerParams p,
                                                  const float* __restrict__ proj, float* __restrict__ e_out,
                                                  float* __restrict__ logit_out) {
  // logit_out holds celu(q) + masks; the pre-activation q is recovered in the backward pass from e
  extern __shared__ float sm[];
  const int H = d.H, K = d.K, A = d.A, N = d.N;
  float* pri = sm;                 // [NP] projections of node i
  float* gs = pri + d.NP;          // [PJ][K]
  float* a1 = gs + PJ * K;         // [PJ][H]
  float* es = a1 + PJ * H;         // [PJ][H]
  float* ns = es + PJ * H;         // [PJ]
  const int row = blockIdx.x;
  const int b = row / N, i = row % N;
  for (int t = threadIdx.x; t < d.NP; t += blockDim.x) pri[t] = proj[(size_t)row * d.NP + t];
  const float xi0 = x[(size_t)row * 3 + 0], xi1 = x[(size_t)row * 3 + 1], xi2 = x[(size_t)row * 3 + 2];
  const float* W1g = p.mlp_out0_kernel + (size_t)2 * H * H;   // rows [2H, 2H+K)
  const float* w1n = W1g + (size_t)K * H;                      // row 2H+K
  __syncthreads();
  for (int j0 = 0; j0 < N; j0 += PJ) {
    const int np = min(PJ, N - j0);
    if (threadIdx.x < np) {
      const int j = j0 + threadIdx.x;
      const float* xj = x + (size_t)(b * N + j) * 3;
      float r0 = xj[0] - xi0, r1 = xj[1] - xi1, r2 = xj[2] - xi2;
      float n2 = r0 * r0 + r1 * r1 + r2 * r2;
      ns[threadIdx.x] = sqrtf(fmaxf(n2, 0.f) + 1e-5f);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < np * K; t += blockDim.x) {
      const int pj = t / K, k = t % K;
      const int j = j0 + pj;
      float tt = expf(-ns[pj]);
      float dm = tt - p.rbf_means[k];
      float rho = expf(-p.rbf_betas[k] * dm * dm);
      float u = proj[(size_t)(b * N + j) * d.NP + k] + pri[d.Kp + k];
      if (d.pair_u) u += d.pair_u[((size_t)row * N + j) * d.Kp + k];       // edge features (SakePairTerms)
      gs[pj * K + k] = rho * u;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < np * H; t += blockDim.x) {
      const int pj = t / H, f = t % H;
      const int j = j0 + pj;
      float z = proj[(size_t)(b * N + j) * d.NP + 2 * d.Kp + f] + pri[2 * d.Kp + H + f];
      if (d.pair_p) z += d.pair_p[((size_t)row * N + j) * H + f];
      z = fmaf(ns[pj], w1n[f], z);
      for (int k = 0; k < K; ++k) z = fmaf(gs[pj * K + k], W1g[(size_t)k * H + f], z);
      a1[pj * H + f] = siluf_(z);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < np * H; t += blockDim.x) {
      const int pj = t / H, f = t % H;
      float acc = p.mlp_out2_bias[f];
      for (int g = 0; g < H; ++g) acc = fmaf(a1[pj * H + g], p.mlp_out2_kernel[(size_t)g * H + f], acc);
      es[pj * H + f] = acc;
      e_out[((size_t)row * N + j0 + pj) * H + f] = acc;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < np * A; t += blockDim.x) {
      const int pj = t / A, a = t % A;
      const int j = j0 + pj;
      float q = p.sem_bias[a];
      for (int f = 0; f < H; ++f) q = fmaf(es[pj * H + f], p.sem_kernel[(size_t)f * A + a], q);
      float s = celu2f_(q);
      if (j == i) s -= 1e5f;
      if (mask) s -= 1e5f * (1.0f - mask[(size_t)row * N + j]);
      logit_out[((size_t)row * N + j) * A + a] = s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// attn_fwd: softmax over senders j, mask, renormalise (layers.py:167,172-180) + aggregate
// (layers.py:135-140).  One warp per receiving atom, several rows per CTA; att is normalised in place.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attn_fwd(Dims d, const float* __restrict__ x, const float* __restrict__ mask,
                                                  const float* __restrict__ e, const float* lg, float* att,
                                                  float* __restrict__ he) {
  extern __shared__ float sm[];
  const int A = d.A, H = d.H, C = d.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x * nw + warp;
  if (row >= dims_rows(d)) return;
  const RowInfo ri = row_info(d, row);                 // ragged batches: n senders, pair slots from ri.pair0
  const int N = ri.n;
  float* as = sm + (size_t)warp * d.N * A;  // [N][A]
  float* arow = att + (size_t)ri.pair0 * A;
  const float* mrow = mask ? mask + (size_t)ri.pair0 : nullptr;
  const float* lrow = lg + (size_t)ri.pair0 * A;       // logits (may alias att: read completely before any write)
  if (A == 4 && N <= 32) {
    // short rows (every ragged workload of the scripts): lane = sender, the four heads side by side in registers —
    // three warp reductions of a float4 instead of twelve scalar ones with a shared-memory pass between them.
    // Same operations in the same order as the general path below, so the results are bit-identical.
    const bool on = lane < N;
    float4 s = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (on) s = *reinterpret_cast<const float4*>(lrow + lane * 4);
    float4 mx = s;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, o)); mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, o));
      mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, o)); mx.w = fmaxf(mx.w, __shfl_xor_sync(0xffffffffu, mx.w, o));
    }
    float4 ex = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) ex = make_float4(expf(s.x - mx.x), expf(s.y - mx.y), expf(s.z - mx.z), expf(s.w - mx.w));
    float4 sum = ex;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o); sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
      sum.z += __shfl_xor_sync(0xffffffffu, sum.z, o); sum.w += __shfl_xor_sync(0xffffffffu, sum.w, o);
    }
    float4 c = make_float4(ex.x * (1.0f / sum.x), ex.y * (1.0f / sum.y), ex.z * (1.0f / sum.z), ex.w * (1.0f / sum.w));
    if (on && d.cutoff) {
      const float eps = cosine_cutoff_(pair_dist_(x, row, ri.mol0 + lane), d.cut_lo, d.cut_hi);   // euclidean attention
      c.x *= eps; c.y *= eps; c.z *= eps; c.w *= eps;
    }
    if (on && mrow) { const float m = mrow[lane]; c.x *= m; c.y *= m; c.z *= m; c.w *= m; }
    float4 cs = c;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    c.x *= cs.x > 0.f ? 1.0f / cs.x : 0.f; c.y *= cs.y > 0.f ? 1.0f / cs.y : 0.f;      // guarded: fully masked row -> att = 0
    c.z *= cs.z > 0.f ? 1.0f / cs.z : 0.f; c.w *= cs.w > 0.f ? 1.0f / cs.w : 0.f;
    if (on) { *reinterpret_cast<float4*>(as + lane * 4) = c; *reinterpret_cast<float4*>(arow + lane * 4) = c; }
    __syncwarp();
  } else {
  for (int t = lane; t < N * A; t += 32) as[t] = lrow[t];
  __syncwarp();
  for (int a = 0; a < A; ++a) {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, as[j * A + a]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      float ex = expf(as[j * A + a] - mx);
      as[j * A + a] = ex;
      sum += ex;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float csum = 0.f;
    for (int j = lane; j < N; j += 32) {
      float c = as[j * A + a] * inv;          // semantic attention (softmax)
      if (d.cutoff) c *= cosine_cutoff_(pair_dist_(x, row, ri.mol0 + j), d.cut_lo, d.cut_hi);   // euclidean attention
      if (mrow) c *= mrow[j];                 // combined = euclidean * semantic * mask
      as[j * A + a] = c;
      csum += c;
    }
    for (int o = 16; o; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    const float rinv = csum > 0.f ? 1.0f / csum : 0.f;   // guarded: fully masked row -> att = 0
    for (int j = lane; j < N; j += 32) as[j * A + a] *= rinv;
  }
  __syncwarp();
  for (int t = lane; t < N * A; t += 32) arow[t] = as[t];
  }
  // aggregate: he[c = f*A+a] = sum_j e[j,f] * att[j,a] * m_j
  if (A == 4 && H == 64 && d.g8) {
    // e and he in the G8 layout (common.cuh): a group of 8 consecutive pair slots x 16 units is one contiguous 2 KB
    // block.  The warp pulls two groups at a time with fully coalesced 128-bit loads (lane l: float4 l + 32 k of the
    // group = pair l & 7, unit (l >> 3) + 4 k) into a padded [16 pairs][64 + 4] shared-memory tile, then every lane
    // owns f = 2*lane, 2*lane+1 and all four heads as before.  (Reading e per pair straight from the G8 rows costs
    // 16 L1 wavefronts per pair: it was half of this kernel's time at cfg5.)
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* es = sm + (size_t)nw * d.N * A + (size_t)warp * (16 * 68);
    const float4* e4 = reinterpret_cast<const float4*>(e);
    const long long q0 = ri.pair0, q1 = ri.pair0 + N;
    const long long glast = (q1 - 1) >> 3;
    for (long long g0 = q0 >> 3; g0 <= glast; g0 += 2) {
      float4 ev[8];
#pragma unroll
      for (int gg = 0; gg < 2; ++gg)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ev[gg * 4 + k] = g0 + gg <= glast ? __ldg(e4 + (g0 + gg) * (16 * G8S) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int gg = 0; gg < 2; ++gg)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float4*>(es + (gg * 8 + (lane & 7)) * 68 + ((lane >> 3) + 4 * k) * 4) = ev[gg * 4 + k];
      __syncwarp();
#pragma unroll 4
      for (int pp = 0; pp < 16; ++pp) {
        const long long q = g0 * 8 + pp;
        if (q >= q0 && q < q1) {                          // warp-uniform
          const int j = (int)(q - q0);
          const float2 ef = *reinterpret_cast<const float2*>(es + pp * 68 + 2 * lane);
          float4 w = *reinterpret_cast<const float4*>(as + j * 4);
          if (mrow) { const float m = mrow[j]; w.x *= m; w.y *= m; w.z *= m; w.w *= m; }
          acc[0] = fmaf(ef.x, w.x, acc[0]); acc[1] = fmaf(ef.x, w.y, acc[1]); acc[2] = fmaf(ef.x, w.z, acc[2]); acc[3] = fmaf(ef.x, w.w, acc[3]);
          acc[4] = fmaf(ef.y, w.x, acc[4]); acc[5] = fmaf(ef.y, w.y, acc[5]); acc[6] = fmaf(ef.y, w.z, acc[6]); acc[7] = fmaf(ef.y, w.w, acc[7]);
        }
      }
      __syncwarp();
    }
    float4* ho = reinterpret_cast<float4*>(he) + g8_row(row, 64) + 2 * lane * G8S;     // he columns 8*lane .. 8*lane+7
    ho[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    ho[G8S] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else if (A == 4 && H == 64) {
    // lane owns f = 2*lane, 2*lane+1 (coalesced float2 loads of e) and all four heads
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* ep = e + (size_t)ri.pair0 * 64 + 2 * lane;
    for (int j0 = 0; j0 < N; j0 += 8) {
      float2 ev[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)                        // 8 independent 256-byte row loads in flight per warp
        ev[u] = j0 + u < N ? __ldg(reinterpret_cast<const float2*>(ep + (size_t)(j0 + u) * 64)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + u < N) {
          float4 w = *reinterpret_cast<const float4*>(as + (j0 + u) * 4);
          if (mrow) { const float m = mrow[j0 + u]; w.x *= m; w.y *= m; w.z *= m; w.w *= m; }
          acc[0] = fmaf(ev[u].x, w.x, acc[0]); acc[1] = fmaf(ev[u].x, w.y, acc[1]); acc[2] = fmaf(ev[u].x, w.z, acc[2]); acc[3] = fmaf(ev[u].x, w.w, acc[3]);
          acc[4] = fmaf(ev[u].y, w.x, acc[4]); acc[5] = fmaf(ev[u].y, w.y, acc[5]); acc[6] = fmaf(ev[u].y, w.z, acc[6]); acc[7] = fmaf(ev[u].y, w.w, acc[7]);
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(he + (size_t)row * 256 + 8 * lane);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
    for (int c = lane; c < C; c += 32) {
      const int f = c / A, a = c % A;
      float acc = 0.f;
      for (int j = 0; j < N; ++j) {
        float w = as[j * A + a];
        if (mrow) w *= mrow[j];
        acc = fmaf(e[((size_t)ri.pair0 + j) * H + f], w, acc);
      }
      he[(size_t)row * C + c] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------
// mix_fwd (generic): coef = tanh(E @ Wx), ssum[c,d] = sum_j dir_d * coef_c * m  (layers.py:111-123)
// One CTA per receiving atom, one thread per output coefficient c'.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mix_fwd(Dims d, const float* __restrict__ x,
                                                 const float* __restrict__ mask, const float* __restrict__ Wx,
                                                 const float* __restrict__ e, const float* __restrict__ att,
                                                 float* __restrict__ ssum) {
  extern __shared__ float sm[];
  const int N = d.N, A = d.A, H = d.H, C = d.C;
  float* ET = sm;               // [C][MJ]
  float* dirm = ET + C * MJ;    // [MJ][4]
  const int row = blockIdx.x;
  const int b = row / N;
  const float xi0 = x[(size_t)row * 3 + 0], xi1 = x[(size_t)row * 3 + 1], xi2 = x[(size_t)row * 3 + 2];
  for (int cp0 = 0; cp0 < C; cp0 += blockDim.x) {
    const int cp = cp0 + threadIdx.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int j0 = 0; j0 < N; j0 += MJ) {
      const int np = min(MJ, N - j0);
      __syncthreads();
      for (int t = threadIdx.x; t < C * MJ; t += blockDim.x) {
        const int c = t / MJ, pj = t % MJ;
        float val = 0.f;
        if (pj < np) {
          size_t pr = (size_t)row * N + j0 + pj;
          val = e[pr * H + c / A] * att[pr * A + c % A];
        }
        ET[t] = val;
      }
      if (threadIdx.x < MJ) {
        const int pj = threadIdx.x;
        float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pj < np) {
          const int j = j0 + pj;
          const float* xj = x + (size_t)(b * N + j) * 3;
          float r0 = xj[0] - xi0, r1 = xj[1] - xi1, r2 = xj[2] - xi2;
          float n = sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 0.f) + 1e-5f);
          float inv = 1.0f / (n + 1e-5f);
          float m = mask ? mask[(size_t)row * N + j] : 1.0f;
          dm = make_float4(r0 * inv * m, r1 * inv * m, r2 * inv * m, m);
        }
        reinterpret_cast<float4*>(dirm)[pj] = dm;
      }
      __syncthreads();
      if (cp < C) {
        float acc[MJ];
#pragma unroll
        for (int q = 0; q < MJ; ++q) acc[q] = 0.f;
        for (int c = 0; c < C; ++c) {
          float w = Wx[(size_t)c * C + cp];
          const float4 ea = reinterpret_cast<const float4*>(ET + c * MJ)[0];
          const float4 eb = reinterpret_cast<const float4*>(ET + c * MJ)[1];
          acc[0] = fmaf(ea.x, w, acc[0]); acc[1] = fmaf(ea.y, w, acc[1]);
          acc[2] = fmaf(ea.z, w, acc[2]); acc[3] = fmaf(ea.w, w, acc[3]);
          acc[4] = fmaf(eb.x, w, acc[4]); acc[5] = fmaf(eb.y, w, acc[5]);
          acc[6] = fmaf(eb.z, w, acc[6]); acc[7] = fmaf(eb.w, w, acc[7]);
        }
#pragma unroll
        for (int q = 0; q < MJ; ++q) {
          float co = tanhf(acc[q]);
          const float4 dm = reinterpret_cast<const float4*>(dirm)[q];
          s0 = fmaf(dm.x, co, s0); s1 = fmaf(dm.y, co, s1); s2 = fmaf(dm.z, co, s2);
        }
      }
    }
    if (cp < C) {
      float* o = ssum + ((size_t)row * C + cp) * 3;
      o[0] = s0; o[1] = s1; o[2] = s2;
    }
  }
}

// ------------------------------------------------------------------------------------------
// node_post: norms, post_norm_mlp, node_mlp + residual, velocity / position update
// (layers.py:123-131,142-151,218-232).  NODES nodes per CTA.
// ------------------------------------------------------------------------------------------
constexpr int POST_WROWS = 32;    // weight rows per staged chunk (2 x 8 KB)
struct NodePostSmem {
  float *nrm, *he, *hin, *hp1, *hcomb, *n1, *hout, *den, *den2, *wbuf;
};
__device__ inline NodePostSmem node_post_carve(float* sm, const Dims& d) {
  NodePostSmem s;
  s.nrm = sm;                       // [NODES][C]
  s.he = s.nrm + NODES * d.C;       // [NODES][C]
  s.hin = s.he + NODES * d.C;       // [NODES][H]
  s.hp1 = s.hin + NODES * d.H;
  s.hcomb = s.hp1 + NODES * d.H;
  s.n1 = s.hcomb + NODES * d.H;
  s.hout = s.n1 + NODES * d.H;
  s.den = s.hout + NODES * d.H;     // [NODES]
  s.den2 = s.den + NODES;           // [NODES]
  s.wbuf = s.den2 + NODES + 32;     // [2][POST_WROWS][64] staged weight rows (16-byte aligned)
  return s;
}
size_t node_post_smem_bytes(const Dims& d) { return sizeof(float) * (NODES * (2 * d.C + 5 * d.H) + 2 * NODES + 64 + 2 * POST_WROWS * 64); }

__global__ void __launch_bounds__(256) k_node_post(Dims d, const SakeLayerParams p, const float* __restrict__ h,
                                                   const float* __restrict__ x, const float* __restrict__ v,
                                                   const float* __restrict__ mask, const float* __restrict__ ssum,
                                                   const float* __restrict__ he_in, float* __restrict__ h_out,
                                                   float* __restrict__ x_out, float* __restrict__ v_out) {
  extern __shared__ float sm[];
  NodePostSmem s = node_post_carve(sm, d);
  const int H = d.H, C = d.C, N = d.N;
  const int r0 = blockIdx.x * NODES;
  const int nn = min(NODES, d.R - r0);
  if (threadIdx.x < NODES) {
    float dn = (float)N, dn2 = (float)N;
    if (mask && threadIdx.x < nn) {
      float ms = 0.f;
      for (int j = 0; j < N; ++j) ms += mask[(size_t)(r0 + threadIdx.x) * N + j];
      dn = ms + 1e-8f;     // layers.py:123
      dn2 = ms + 1e-10f;   // layers.py:221
    }
    s.den[threadIdx.x] = dn;
    s.den2[threadIdx.x] = dn2;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NODES * C; t += blockDim.x) {
    const int n = t / C, c = t % C;
    float nr = 0.f, hv = 0.f;
    if (n < nn) {
      const float* sp = ssum + ((size_t)(r0 + n) * C + c) * 3;
      float inv = 1.0f / s.den[n];
      float a0 = sp[0] * inv, a1 = sp[1] * inv, a2 = sp[2] * inv;
      nr = a0 * a0 + a1 * a1 + a2 * a2;
      hv = he_in[(size_t)(r0 + n) * C + c];
    }
    s.nrm[t] = nr;
    s.he[t] = hv;
  }
  for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) {
    const int n = t / H;
    s.hin[t] = n < nn ? h[(size_t)r0 * H + t] : 0.f;
  }
  __syncthreads();
  // post_norm_mlp (layers.py:85-92)
  node_dense(s.hp1, s.nrm, C, C, p.post0_kernel, p.post0_bias, H, false, s.wbuf, POST_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) s.hp1[t] = siluf_(s.hp1[t]);
  __syncthreads();
  node_dense(s.hcomb, s.hp1, H, H, p.post2_kernel, p.post2_bias, H, false, s.wbuf, POST_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) s.hcomb[t] = d.spatial ? siluf_(s.hcomb[t]) : 0.f;
  __syncthreads();
  // node_mlp over [h | he | hcomb] + residual (layers.py:142-151)
  node_dense(s.n1, s.hin, H, H, p.node0_kernel, p.node0_bias, H, false, s.wbuf, POST_WROWS);
  __syncthreads();
  node_dense(s.n1, s.he, C, C, p.node0_kernel + (size_t)H * H, nullptr, H, true, s.wbuf, POST_WROWS);
  __syncthreads();
  node_dense(s.n1, s.hcomb, H, H, p.node0_kernel + (size_t)(H + C) * H, nullptr, H, true, s.wbuf, POST_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) s.n1[t] = siluf_(s.n1[t]);
  __syncthreads();
  node_dense(s.hout, s.n1, H, H, p.node2_kernel, p.node2_bias, H, false, s.wbuf, POST_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) {
    const float ho = s.hin[t] + siluf_(s.hout[t]);
    s.hout[t] = ho;
    if (t / H < nn) h_out[(size_t)r0 * H + t] = ho;
  }
  __syncthreads();
  // velocity / position update: one warp per node
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < nn; n += (blockDim.x >> 5)) {
    const size_t row = (size_t)(r0 + n);
    if (!d.update) {
      if (lane < 3) {
        x_out[row * 3 + lane] = x[row * 3 + lane];
        if (v && v_out) v_out[row * 3 + lane] = v[row * 3 + lane];
      }
      continue;
    }
    float dv0 = 0.f, dv1 = 0.f, dv2 = 0.f;
    if (d.spatial) {
      for (int c = lane; c < C; c += 32) {
        float wv = p.v_mixing_kernel[c];
        const float* sp = ssum + (row * C + c) * 3;
        dv0 = fmaf(wv, sp[0], dv0); dv1 = fmaf(wv, sp[1], dv1); dv2 = fmaf(wv, sp[2], dv2);
      }
    }
    float y = 0.f;
    if (d.has_v) {
      for (int f = lane; f < H; f += 32) {
        float acc = p.vel0_bias[f];
        for (int g = 0; g < H; ++g) acc = fmaf(s.hout[n * H + g], p.vel0_kernel[(size_t)g * H + f], acc);
        y = fmaf(siluf_(acc), p.vel2_kernel[f], y);
      }
    }
    for (int o = 16; o; o >>= 1) {
      dv0 += __shfl_xor_sync(0xffffffffu, dv0, o);
      dv1 += __shfl_xor_sync(0xffffffffu, dv1, o);
      dv2 += __shfl_xor_sync(0xffffffffu, dv2, o);
      y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if (lane < 3) {
      float dvv = (lane == 0 ? dv0 : lane == 1 ? dv1 : dv2) / s.den2[n];
      float vn = dvv;
      if (d.has_v) vn += 2.0f * sigmoidf_(y) * v[row * 3 + lane];
      v_out[row * 3 + lane] = vn;
      x_out[row * 3 + lane] = x[row * 3 + lane] + vn;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
template <typename Kern>
static int ensure_smem(Kern kern, size_t smem) {
  if (smem > 48 * 1024) {
    if (smem > 227 * 1024) {
      set_error("generic engine: %zu bytes of shared memory needed (H/A too large)", smem);
      return SAKE_EUNSUPPORTED;
    }
    SAKE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  return 0;
}

// per-node projections: fills sv.nodeproj
int gen_node_pre(const Dims& d, const SakeLayerParams& p, const float* h, const Saved& sv, cudaStream_t st) {
  int rc;
  size_t smem = sizeof(float) * (NODES * d.H + (d.g8 ? NODES * d.NP : 0));
  if ((rc = ensure_smem(k_node_pre, smem))) return rc;
  ProfScope prof(9, d.R, st);
  k_node_pre<<<(d.R + NODES - 1) / NODES, 256, smem, st>>>(d, h, p.mlp_in_kernel, p.mlp_in_bias, p.mlp_out0_kernel,
                                                            p.mlp_out0_bias, sv.nodeproj);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// edge model + attention logits on CUDA cores: fills sv.e and the logits in sv.att
int gen_edge_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                 cudaStream_t st) {
  int rc;
  size_t smem = sizeof(float) * (d.NP + PJ * d.K + 2 * PJ * d.H + PJ);
  if ((rc = ensure_smem(k_edge_fwd, smem))) return rc;
  k_edge_fwd<<<d.R, 256, smem, st>>>(d, x, mask, p, sv.nodeproj, sv.e, sv.att);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// softmax over senders + aggregate: normalises sv.att in place, fills sv.he
int gen_attn_fwd(const Dims& d, const float* x, const float* mask, const Saved& sv, cudaStream_t st) {
  int rc;
  int nw = 8;
  while (nw > 1 && sizeof(float) * d.N * d.A * nw > 160 * 1024) nw >>= 1;
  size_t smem = sizeof(float) * (d.N * d.A * nw + (d.g8 ? nw * 16 * 68 : 0));
  if ((rc = ensure_smem(k_attn_fwd, smem))) return rc;
  ProfScope prof(10, d.P, st);
  k_attn_fwd<<<(d.R + nw - 1) / nw, nw * 32, smem, st>>>(d, x, mask, sv.e, sv.logit, sv.att, sv.he);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// x_mixing + tanh + sum over senders on CUDA cores: fills sv.ssum
int gen_mix_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                cudaStream_t st) {
  size_t smem = sizeof(float) * (d.C * MJ + MJ * 4);
  int rc;
  if ((rc = ensure_smem(k_mix_fwd, smem))) return rc;
  ProfScope prof(1, d.P, st);
  k_mix_fwd<<<d.R, 256, smem, st>>>(d, x, mask, p.x_mixing_kernel, sv.e, sv.att, sv.ssum);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int gen_node_post(const Dims& d, const SakeLayerParams& p, const float* h, const float* x, const float* v,
                  const float* mask, float* h_out, float* x_out, float* v_out, const Saved& sv,
                  cudaStream_t st) {
  size_t smem = node_post_smem_bytes(d);
  int rc;
  if ((rc = ensure_smem(k_node_post, smem))) return rc;
  k_node_post<<<(d.R + NODES - 1) / NODES, 256, smem, st>>>(d, p, h, x, v, mask, sv.ssum, sv.he, h_out, x_out,
                                                            v_out);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sake
