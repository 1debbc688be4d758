// Generic fp32 CUDA-core backward (VJP) of DenseSAKELayer.  Mirrors generic_fwd.cu in reverse:
//   node_post_bwd -> mix_bwd -> attn_bwd -> edge_bwd -> node_pre_bwd   (+ x_mixing dW GEMM)
// Derivation: SURVEY Appendix B; every forward formula cites sake/layers.py in generic_fwd.cu.
#include "common.cuh"

namespace sake {

static constexpr int NODES = SAKE_NODES;    // nodes per CTA in the per-node kernels
static constexpr int PJ = 16;
static constexpr int MJ = 8;

// gW[r][f] += sum_n X[n][r] * G[n][f]   (X: smem [NODES][ldx], G: smem [NODES][out])
__device__ __forceinline__ void accum_outer(float* gW, const float* X, int ldx, int in, const float* G, int out,
                                            int nn) {
  for (int t = threadIdx.x; t < in * out; t += blockDim.x) {
    const int r = t / out, f = t % out;
    float s = 0.f;
    for (int n = 0; n < nn; ++n) s = fmaf(X[n * ldx + r], G[n * out + f], s);
    atomicAdd(gW + t, s);
  }
}
__device__ __forceinline__ void accum_bias(float* gb, const float* G, int out, int nn) {
  for (int f = threadIdx.x; f < out; f += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < nn; ++n) s += G[n * out + f];
    atomicAdd(gb + f, s);
  }
}

// ------------------------------------------------------------------------------------------
// node_post_bwd
// ------------------------------------------------------------------------------------------
// column offsets of the per-node record (NB_LD floats) consumed by tc_node_dw
enum { NB_N1 = 0, NB_GT2 = 64, NB_CAT = 128, NB_GT1 = 512, NB_HP1 = 576, NB_GTP2 = 640, NB_GTP1 = 704, NB_NRM = 768,
       NB_HOUT = 1024, NB_GTV = 1088, NB_AV = 1152, NB_GY = 1216, NB_LD = 1232 };
__device__ __forceinline__ void nb_store(float* nbuf, int r0, int nn, int col, const float* src, int w, bool silu) {
  for (int t = threadIdx.x; t < nn * w; t += blockDim.x) {
    const int n = t / w, f = t % w;
    const float v = src[n * w + f];
    nbuf[(size_t)(r0 + n) * NB_LD + col + f] = silu ? siluf_(v) : v;
  }
}
// transposed copies of the node-level weight matrices, so that the W^T g products of the backward pass
// read memory coalesced (thread index = input feature)
struct NodeWT {
  const float *node0T;   // [H][H+C+H]
  const float *node2T;   // [H][H]
  const float *post0T;   // [H][C]
  const float *post2T;   // [H][H]
  const float *vel0T;    // [H][H]
  const float *mlp_inT;  // [K][2H]
  const float *w1hT;     // [H][2H]   rows [0,2H) of mlp_out[0]
};
__host__ __device__ inline size_t node_wt_floats(const Dims& d) {
  return (size_t)d.H * (2 * d.H + d.C) + 3 * (size_t)d.H * d.H + (size_t)d.H * d.C + (size_t)d.K * 2 * d.H + (size_t)d.H * 2 * d.H;
}
static NodeWT carve_node_wt(const Dims& d, float* base) {
  NodeWT w;
  float* q = base;
  w.node0T = q; q += (size_t)d.H * (2 * d.H + d.C);
  w.node2T = q; q += (size_t)d.H * d.H;
  w.post0T = q; q += (size_t)d.H * d.C;
  w.post2T = q; q += (size_t)d.H * d.H;
  w.vel0T = q; q += (size_t)d.H * d.H;
  w.mlp_inT = q; q += (size_t)d.K * 2 * d.H;
  w.w1hT = q;
  return w;
}
__global__ void k_node_wt(Dims d, const SakeLayerParams p, float* __restrict__ base) {
  const int H = d.H, C = d.C, K = d.K;
  const size_t n0 = (size_t)H * (2 * H + C), n1 = (size_t)H * H, n2 = (size_t)H * C, n5 = (size_t)K * 2 * H, n6 = (size_t)H * 2 * H;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // out[f][r] = in[r][f]
  if (t < n0) { const int NC = 2 * H + C; const int f = t / NC, r = t % NC; base[t] = p.node0_kernel[(size_t)r * H + f]; return; }
  t -= n0; base += n0;
  if (t < n1) { const int f = t / H, r = t % H; base[t] = p.node2_kernel[(size_t)r * H + f]; return; }
  t -= n1; base += n1;
  if (t < n2) { const int f = t / C, r = t % C; base[t] = p.post0_kernel[(size_t)r * H + f]; return; }
  t -= n2; base += n2;
  if (t < n1) { const int f = t / H, r = t % H; base[t] = p.post2_kernel[(size_t)r * H + f]; return; }
  t -= n1; base += n1;
  if (t < n1) { const int f = t / H, r = t % H; base[t] = p.vel0_kernel ? p.vel0_kernel[(size_t)r * H + f] : 0.f; return; }
  t -= n1; base += n1;
  if (t < n5) { const int k = t / (2 * H), r = t % (2 * H); base[t] = p.mlp_in_kernel[(size_t)r * K + k]; return; }
  t -= n5; base += n5;
  if (t < n6) { const int q = t / (2 * H), r = t % (2 * H); base[t] = p.mlp_out0_kernel[(size_t)r * H + q]; return; }
}

constexpr int PRE_WROWS = 16;     // k_node_pre_bwd: weight rows per staged chunk
constexpr int BWD_WROWS = 16;     // weight rows per staged chunk (2 x 4 KB: two CTAs per SM still fit)
size_t node_post_bwd_smem_bytes(const Dims& d) {
  return sizeof(float) * (NODES * (2 * d.C + 17 * d.H) + 8 * NODES + 64 + 2 * BWD_WROWS * 64);
}

__global__ void __launch_bounds__(256) k_node_post_bwd(
    Dims d, const SakeLayerParams p, NodeWT wt, const float* __restrict__ h, const float* __restrict__ v,
    const float* __restrict__ mask, const float* __restrict__ ssum, const float* __restrict__ he_in,
    const float* __restrict__ dh_out, const float* __restrict__ dx_out, const float* __restrict__ dv_out,
    float* __restrict__ dh, float* __restrict__ dx, float* __restrict__ dv, float* __restrict__ T,
    float* __restrict__ ghe, SakeLayerGrads g, int want_grads, float* __restrict__ nbuf, float* __restrict__ tmax) {
  extern __shared__ float sm[];
  __shared__ int tmx[NODES];
  if (threadIdx.x < NODES) tmx[threadIdx.x] = 0;
  const int H = d.H, C = d.C, N = d.N, NH = NODES * d.H;
  // nbuf != NULL (tcgen05 engines, H = 64): the per-node operands of the weight-gradient contractions are
  // written out (NB_LD floats per node) and the contractions run on the tensor cores (tc_node_dw).
  const bool to_buf = want_grads && nbuf != nullptr;
  const bool atom = want_grads && nbuf == nullptr;
  float* nrm = sm;                   // [NODES][C]
  float* hes = nrm + NODES * C;      // [NODES][C]
  float* hin = hes + NODES * C;      // [NODES][H] each below
  float* hp1 = hin + NH;  float* dhp1 = hp1 + NH;      // silu(tp1), silu'(tp1)
  float* hcb = dhp1 + NH; float* dhcb = hcb + NH;      // silu(tp2) (h_combinations), silu'(tp2)
  float* n1 = dhcb + NH;  float* dn1 = n1 + NH;        // silu(t1), silu'(t1)
  float* dn2 = dn1 + NH;                               // silu'(t2)
  float* hout = dn2 + NH;
  float* av = hout + NH;  float* dav = av + NH;        // silu(tv), silu'(tv)
  float* ghout = dav + NH;
  float* gt2 = ghout + NH;
  float* gt1 = gt2 + NH;
  float* gtp2 = gt1 + NH;
  float* gtp1 = gtp2 + NH;
  float* gtv = gtp1 + NH;
  float* den = gtv + NH;             // [NODES]
  float* den2 = den + NODES;
  float* gy = den2 + NODES;
  float* gate = gy + NODES;
  float* gdv = gate + NODES;         // [NODES][3] (+pad)
  float* wbuf = gdv + 4 * NODES + 32;  // [2][BWD_WROWS][64] staged weight rows (16-byte aligned)
  const int r0 = blockIdx.x * NODES;
  const int nn = min(NODES, d.R - r0);
  const bool upd = d.update != 0, hv = d.has_v != 0, spatial = d.spatial != 0;

  // ---------------- recompute forward (layers.py:123-131,142-151,226-229) ----------------
  if (threadIdx.x < NODES) {
    float dn = (float)N, dnb = (float)N;
    if (mask && threadIdx.x < nn) {
      float ms = 0.f;
      for (int j = 0; j < N; ++j) ms += mask[(size_t)(r0 + threadIdx.x) * N + j];
      dn = ms + 1e-8f;
      dnb = ms + 1e-10f;
    }
    den[threadIdx.x] = dn;
    den2[threadIdx.x] = dnb;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < NODES * C; t += blockDim.x) {
    const int n = t / C, c = t % C;
    float nr = 0.f, hvv = 0.f;
    if (n < nn) {
      const float* sp = ssum + ((size_t)(r0 + n) * C + c) * 3;
      const float inv = 1.0f / den[n];
      const float a0 = sp[0] * inv, a1 = sp[1] * inv, a2 = sp[2] * inv;
      nr = a0 * a0 + a1 * a1 + a2 * a2;
      hvv = he_in[(size_t)(r0 + n) * C + c];
    }
    nrm[t] = nr;
    hes[t] = hvv;
  }
  for (int t = threadIdx.x; t < NH; t += blockDim.x) {
    const int n = t / H;
    hin[t] = n < nn ? h[(size_t)r0 * H + t] : 0.f;
    ghout[t] = n < nn ? dh_out[(size_t)r0 * H + t] : 0.f;
  }
  __syncthreads();
  node_dense(hp1, nrm, C, C, p.post0_kernel, p.post0_bias, H, false, wbuf, BWD_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NH; t += blockDim.x) { const float z = hp1[t]; hp1[t] = siluf_(z); dhp1[t] = dsiluf_(z); }
  __syncthreads();
  node_dense(hcb, hp1, H, H, p.post2_kernel, p.post2_bias, H, false, wbuf, BWD_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NH; t += blockDim.x) {
    const float z = hcb[t];
    hcb[t] = spatial ? siluf_(z) : 0.f;
    dhcb[t] = spatial ? dsiluf_(z) : 0.f;
  }
  __syncthreads();
  node_dense(n1, hin, H, H, p.node0_kernel, p.node0_bias, H, false, wbuf, BWD_WROWS);
  __syncthreads();
  node_dense(n1, hes, C, C, p.node0_kernel + (size_t)H * H, nullptr, H, true, wbuf, BWD_WROWS);
  __syncthreads();
  node_dense(n1, hcb, H, H, p.node0_kernel + (size_t)(H + C) * H, nullptr, H, true, wbuf, BWD_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NH; t += blockDim.x) { const float z = n1[t]; n1[t] = siluf_(z); dn1[t] = dsiluf_(z); }
  __syncthreads();
  node_dense(hout, n1, H, H, p.node2_kernel, p.node2_bias, H, false, wbuf, BWD_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NH; t += blockDim.x) { const float z = hout[t]; dn2[t] = dsiluf_(z); hout[t] = hin[t] + siluf_(z); }
  __syncthreads();

  // ---------------- velocity / position update backward (layers.py:226-232) ----------------
  if (upd && hv) {
    node_dense(av, hout, H, H, p.vel0_kernel, p.vel0_bias, H, false, wbuf, BWD_WROWS);
    __syncthreads();
    for (int t = threadIdx.x; t < NH; t += blockDim.x) { const float z = av[t]; av[t] = siluf_(z); dav[t] = dsiluf_(z); }
  }
  __syncthreads();
  if (threadIdx.x < NODES) {
    const int n = threadIdx.x;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gyv = 0.f, gt = 0.f;
    if (n < nn) {
      const size_t row = (size_t)(r0 + n);
      const float dxo0 = dx_out ? dx_out[row * 3 + 0] : 0.f, dxo1 = dx_out ? dx_out[row * 3 + 1] : 0.f,
                  dxo2 = dx_out ? dx_out[row * 3 + 2] : 0.f;
      const float dvo0 = dv_out ? dv_out[row * 3 + 0] : 0.f, dvo1 = dv_out ? dv_out[row * 3 + 1] : 0.f,
                  dvo2 = dv_out ? dv_out[row * 3 + 2] : 0.f;
      dx[row * 3 + 0] = dxo0; dx[row * 3 + 1] = dxo1; dx[row * 3 + 2] = dxo2;   // x' = x + v'
      if (upd) {
        g0 = dvo0 + dxo0; g1 = dvo1 + dxo1; g2 = dvo2 + dxo2;                   // cotangent of v'
        if (hv) {
          float y = 0.f;
          for (int f = 0; f < H; ++f) y = fmaf(av[n * H + f], p.vel2_kernel[f], y);
          gt = 2.0f * sigmoidf_(y);
          const float* vv = v + row * 3;
          const float ggate = g0 * vv[0] + g1 * vv[1] + g2 * vv[2];
          gyv = ggate * gt * (1.0f - 0.5f * gt);
          if (dv) { dv[row * 3 + 0] = gt * g0; dv[row * 3 + 1] = gt * g1; dv[row * 3 + 2] = gt * g2; }
        }
      } else if (dv && hv) {
        dv[row * 3 + 0] = dvo0; dv[row * 3 + 1] = dvo1; dv[row * 3 + 2] = dvo2;   // v passes through
      }
    }
    gdv[n * 3 + 0] = g0; gdv[n * 3 + 1] = g1; gdv[n * 3 + 2] = g2;
    gy[n] = gyv;
    gate[n] = gt;
  }
  __syncthreads();
  if (upd && hv) {
    for (int t = threadIdx.x; t < NH; t += blockDim.x) gtv[t] = p.vel2_kernel[t % H] * gy[t / H] * dav[t];
    __syncthreads();
    if (atom) {
      for (int f = threadIdx.x; f < H; f += blockDim.x) {
        float s = 0.f;
        for (int n = 0; n < nn; ++n) s = fmaf(av[n * H + f], gy[n], s);
        atomicAdd(g.vel2_kernel + f, s);
      }
      accum_outer(g.vel0_kernel, hout, H, H, gtv, H, nn);
      accum_bias(g.vel0_bias, gtv, H, nn);
    }
    if (to_buf) {
      nb_store(nbuf, r0, nn, NB_HOUT, hout, H, false);
      nb_store(nbuf, r0, nn, NB_GTV, gtv, H, false);
      nb_store(nbuf, r0, nn, NB_AV, av, H, false);
      if (threadIdx.x < nn) nbuf[(size_t)(r0 + threadIdx.x) * NB_LD + NB_GY] = gy[threadIdx.x];
    }
    node_dense(ghout, gtv, H, H, wt.vel0T, nullptr, H, true, wbuf, BWD_WROWS);      // g_h' += Wv1 g_tv
    __syncthreads();
  }

  // ---------------- node_mlp backward (layers.py:142-151) ----------------
  for (int t = threadIdx.x; t < NH; t += blockDim.x) gt2[t] = ghout[t] * dn2[t];
  __syncthreads();
  node_dense(gt1, gt2, H, H, wt.node2T, nullptr, H, false, wbuf, BWD_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NH; t += blockDim.x) gt1[t] *= dn1[t];
  __syncthreads();
  if (to_buf) {
    nb_store(nbuf, r0, nn, NB_N1, n1, H, false);
    nb_store(nbuf, r0, nn, NB_GT2, gt2, H, false);
    nb_store(nbuf, r0, nn, NB_GT1, gt1, H, false);
    nb_store(nbuf, r0, nn, NB_CAT, hin, H, false);
    nb_store(nbuf, r0, nn, NB_CAT + H, hes, C, false);
    nb_store(nbuf, r0, nn, NB_CAT + H + C, hcb, H, false);
  }
  if (atom) {
    accum_outer(g.node2_kernel, n1, H, H, gt2, H, nn);
    accum_bias(g.node2_bias, gt2, H, nn);
    accum_outer(g.node0_kernel, hin, H, H, gt1, H, nn);
    accum_outer(g.node0_kernel + (size_t)H * H, hes, C, C, gt1, H, nn);
    if (spatial) accum_outer(g.node0_kernel + (size_t)(H + C) * H, hcb, H, H, gt1, H, nn);
    accum_bias(g.node0_bias, gt1, H, nn);
  }
  // g_cat = Wn1 g_t1 : [dh | ghe | g_hcomb]   (coalesced through the transposed copy)
  {
    const int NC = 2 * H + C;
    auto put = [&](int n, int o, float a) {
      if (o < H) {
        if (n < nn) dh[(size_t)(r0 + n) * H + o] = ghout[n * H + o] + a;
      } else if (o < H + C) {
        if (n < nn) ghe[(size_t)(r0 + n) * C + (o - H)] = a;
      } else {
        const int q = o - H - C;
        gtp2[n * H + q] = a * dhcb[n * H + q];
      }
    };
    if (blockDim.x == 256 && (NC & 63) == 0 && (reinterpret_cast<uintptr_t>(wt.node0T) & 15) == 0) {
      for (int ob = 0; ob < NC; ob += 64)
        node_gemm64(gt1, H, H, wt.node0T + ob, NC, wbuf, BWD_WROWS, [&](int n, int o4, const float* a) {
#pragma unroll
          for (int i = 0; i < 4; ++i) put(n, ob + o4 + i, a[i]);
        });
    } else {
      for (int idx = threadIdx.x; idx < NC * (NODES / 2); idx += blockDim.x) {
        const int o = idx % NC, n0 = (idx / NC) * 2;
        float a0 = 0.f, a1 = 0.f;
        const float* g0p = gt1 + n0 * H;
        const float* g1p = g0p + H;
#pragma unroll 16
        for (int f = 0; f < H; ++f) {
          const float w = wt.node0T[(size_t)f * NC + o];
          a0 = fmaf(g0p[f], w, a0);
          a1 = fmaf(g1p[f], w, a1);
        }
        put(n0, o, a0);
        put(n0 + 1, o, a1);
      }
    }
  }
  __syncthreads();
  // ---------------- post_norm_mlp backward (layers.py:85-92,129-131) ----------------
  node_dense(gtp1, gtp2, H, H, wt.post2T, nullptr, H, false, wbuf, BWD_WROWS);
  __syncthreads();
  for (int t = threadIdx.x; t < NH; t += blockDim.x) gtp1[t] *= dhp1[t];
  __syncthreads();
  if (to_buf && spatial) {
    nb_store(nbuf, r0, nn, NB_HP1, hp1, H, false);
    nb_store(nbuf, r0, nn, NB_GTP2, gtp2, H, false);
    nb_store(nbuf, r0, nn, NB_GTP1, gtp1, H, false);
    nb_store(nbuf, r0, nn, NB_NRM, nrm, C, false);
  }
  if (atom && spatial) {
    accum_outer(g.post2_kernel, hp1, H, H, gtp2, H, nn);
    accum_bias(g.post2_bias, gtp2, H, nn);
    accum_outer(g.post0_kernel, nrm, C, C, gtp1, H, nn);
    accum_bias(g.post0_bias, gtp1, H, nn);
  }
  __syncthreads();
  // g_nrm = Wp1 g_tp1 (overwrites nrm), then
  // T[c][d] = 2*ssum[c][d]*g_nrm[c]/den^2 + Wv[c]*g_dv[d]/den2 ;  gWv[c] += sum_d ssum[c][d]*g_dv[d]/den2
  if (spatial) node_dense(nrm, gtp1, H, H, wt.post0T, nullptr, C, false, wbuf, BWD_WROWS);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float wv = (spatial && upd) ? p.v_mixing_kernel[c] : 0.f;
    float gwv = 0.f;
    for (int n = 0; n < nn; ++n) {
      const size_t row = (size_t)(r0 + n);
      float t0 = 0.f, t1v = 0.f, t2v = 0.f;
      if (spatial) {
        const float* sp = ssum + (row * C + c) * 3;
        const float k2 = 2.0f * nrm[n * C + c] / (den[n] * den[n]);
        t0 = k2 * sp[0]; t1v = k2 * sp[1]; t2v = k2 * sp[2];
        if (upd) {
          const float wq = wv / den2[n];
          t0 = fmaf(wq, gdv[n * 3 + 0], t0); t1v = fmaf(wq, gdv[n * 3 + 1], t1v); t2v = fmaf(wq, gdv[n * 3 + 2], t2v);
          gwv += (sp[0] * gdv[n * 3 + 0] + sp[1] * gdv[n * 3 + 1] + sp[2] * gdv[n * 3 + 2]) / den2[n];
        }
      }
      *reinterpret_cast<float4*>(T + (row * C + c) * 4) = make_float4(t0, t1v, t2v, 0.f);   // [R][C] float4
      atomicMax(&tmx[n], __float_as_int(fmaxf(fabsf(t0), fmaxf(fabsf(t1v), fabsf(t2v)))));      // non-negative floats order as ints
    }
    if (want_grads && spatial && upd) atomicAdd(g.v_mixing_kernel + c, gwv);
  }
  __syncthreads();
  if (threadIdx.x < nn) tmax[r0 + threadIdx.x] = __int_as_float(tmx[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// mix_bwd (generic): recompute coef, then g_e, g_att, g_dir (and gZ for the dW GEMM)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mix_bwd(Dims d, const float* __restrict__ x, const float* __restrict__ mask,
                                                 const float* __restrict__ Wx, const float* __restrict__ WxT,
                                                 const float* __restrict__ e, const float* __restrict__ att,
                                                 const float* __restrict__ T, const float* __restrict__ ghe,
                                                 float* __restrict__ ge, float* __restrict__ gatt,
                                                 float* __restrict__ gdir, float* __restrict__ gZ_out) {
  extern __shared__ float sm[];
  const int N = d.N, A = d.A, H = d.H, C = d.C;
  float* ET = sm;                  // [C][MJ]
  float* coefT = ET + C * MJ;      // [C][MJ]
  float* gZT = coefT + C * MJ;     // [C][MJ]  (later reused as gE[C][MJ])
  float* Ts = gZT + C * MJ;        // [C][3]
  float* ghes = Ts + C * 3;        // [C]
  float* dirm = ghes + C;          // [MJ][4]
  const int row = blockIdx.x;
  const int b = row / N;
  const float xi0 = x[(size_t)row * 3 + 0], xi1 = x[(size_t)row * 3 + 1], xi2 = x[(size_t)row * 3 + 2];
  for (int t = threadIdx.x; t < C * 3; t += blockDim.x) Ts[t] = T[((size_t)row * C + t / 3) * 4 + t % 3];
  for (int t = threadIdx.x; t < C; t += blockDim.x) ghes[t] = ghe[(size_t)row * C + t];
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
    for (int cp = threadIdx.x; cp < C; cp += blockDim.x) {
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
      const float tt0 = Ts[cp * 3 + 0], tt1 = Ts[cp * 3 + 1], tt2 = Ts[cp * 3 + 2];
#pragma unroll
      for (int q = 0; q < MJ; ++q) {
        float co = tanhf(acc[q]);
        const float4 dm = reinterpret_cast<const float4*>(dirm)[q];
        float gco = dm.x * tt0 + dm.y * tt1 + dm.z * tt2;     // includes the mask
        float gz = gco * (1.0f - co * co);
        coefT[cp * MJ + q] = co;
        gZT[cp * MJ + q] = gz;
        if (gZ_out && q < np) gZ_out[((size_t)row * N + j0 + q) * C + cp] = gz;
      }
    }
    __syncthreads();
    // g_dir[pj][dd] = m * sum_c coef[c] * T[c][dd]
    if (threadIdx.x < np * 3) {
      const int pj = threadIdx.x / 3, dd = threadIdx.x % 3;
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(coefT[c * MJ + pj], Ts[c * 3 + dd], acc);
      gdir[((size_t)row * N + j0 + pj) * 3 + dd] = acc * dirm[pj * 4 + 3];
    }
    // g_E[c] = sum_c' gZ[c'] * Wx[c][c'] + m * ghe[c]   -> stored over ET (ET no longer needed)
    float gEv[MJ];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
#pragma unroll
      for (int q = 0; q < MJ; ++q) gEv[q] = 0.f;
      for (int cp = 0; cp < C; ++cp) {
        float w = WxT[(size_t)cp * C + c];
        const float4 ga = reinterpret_cast<const float4*>(gZT + cp * MJ)[0];
        const float4 gb = reinterpret_cast<const float4*>(gZT + cp * MJ)[1];
        gEv[0] = fmaf(ga.x, w, gEv[0]); gEv[1] = fmaf(ga.y, w, gEv[1]);
        gEv[2] = fmaf(ga.z, w, gEv[2]); gEv[3] = fmaf(ga.w, w, gEv[3]);
        gEv[4] = fmaf(gb.x, w, gEv[4]); gEv[5] = fmaf(gb.y, w, gEv[5]);
        gEv[6] = fmaf(gb.z, w, gEv[6]); gEv[7] = fmaf(gb.w, w, gEv[7]);
      }
      const float gh = ghes[c];
#pragma unroll
      for (int q = 0; q < MJ; ++q) ET[c * MJ + q] = fmaf(dirm[q * 4 + 3], gh, gEv[q]);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < np * H; t += blockDim.x) {
      const int pj = t / H, f = t % H;
      const size_t pr = (size_t)row * N + j0 + pj;
      float acc = 0.f;
      for (int a = 0; a < A; ++a) acc = fmaf(ET[(f * A + a) * MJ + pj], att[pr * A + a], acc);
      ge[pr * H + f] = acc;
    }
    for (int t = threadIdx.x; t < np * A; t += blockDim.x) {
      const int pj = t / A, a = t % A;
      const size_t pr = (size_t)row * N + j0 + pj;
      float acc = 0.f;
      for (int f = 0; f < H; ++f) acc = fmaf(ET[(f * A + a) * MJ + pj], e[pr * H + f], acc);
      gatt[pr * A + a] = acc;
    }
  }
}

// no spatial attention: only the aggregate path feeds e / att
__global__ void __launch_bounds__(256) k_mix_bwd_nospatial(Dims d, const float* __restrict__ mask,
                                                           const float* __restrict__ e, const float* __restrict__ att,
                                                           const float* __restrict__ ghe, float* __restrict__ ge,
                                                           float* __restrict__ gatt, float* __restrict__ gdir) {
  const int A = d.A, H = d.H, C = d.C;
  const size_t row = blockIdx.x;                       // one CTA per receiving atom, all its senders
  if ((int)row >= dims_rows(d)) return;
  const RowInfo ri = row_info(d, (int)row);
  for (int j = 0; j < ri.n; ++j) {
    const size_t pr = (size_t)ri.pair0 + j;
    const float m = mask ? mask[pr] : 1.0f;
    for (int f = threadIdx.x; f < H; f += blockDim.x) {
      float acc = 0.f;
      for (int a = 0; a < A; ++a) acc = fmaf(m * ghe[row * C + f * A + a], att[pr * A + a], acc);
      ge[d.g8 ? g8_elem((long long)pr, H >> 2, f) : pr * H + f] = acc;
    }
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
      float acc = 0.f;
      for (int f = 0; f < H; ++f) acc = fmaf(m * ghe[row * C + f * A + a], e[d.g8 ? g8_elem((long long)pr, H >> 2, f) : pr * H + f], acc);
      gatt[pr * A + a] = acc;
    }
    if (threadIdx.x < 3) gdir[pr * 3 + threadIdx.x] = 0.f;
  }
}

// dWx[c][c'] += sum_p E[p][c] * gZ[p][c']
__global__ void __launch_bounds__(256) k_mix_dw(Dims d, const float* __restrict__ e, const float* __restrict__ att,
                                                const float* __restrict__ gZ, float* __restrict__ gWx,
                                                long long pairs_per_split) {
  __shared__ float Es[16][64 + 1];
  __shared__ float Gs[16][64 + 1];
  const int C = d.C, A = d.A, H = d.H;
  const int c0 = blockIdx.x * 64, cp0 = blockIdx.y * 64;
  const long long pbeg = (long long)blockIdx.z * pairs_per_split;
  const long long pend = min(d.P, pbeg + pairs_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (long long p0 = pbeg; p0 < pend; p0 += 16) {
    for (int t = threadIdx.x; t < 16 * 64; t += blockDim.x) {
      const int pp = t / 64, cc = t % 64;
      const long long pr = p0 + pp;
      float ev = 0.f, gv = 0.f;
      if (pr < pend) {
        if (c0 + cc < C) ev = e[pr * H + (c0 + cc) / A] * att[pr * A + (c0 + cc) % A];
        if (cp0 + cc < C) gv = gZ[pr * C + cp0 + cc];
      }
      Es[pp][cc] = ev;
      Gs[pp][cc] = gv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < 16; ++pp) {
      float ev[4], gv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { ev[q] = Es[pp][ty * 4 + q]; gv[q] = Gs[pp][tx * 4 + q]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = fmaf(ev[a], gv[b2], acc[a][b2]);
    }
    __syncthreads();
  }
  for (int a = 0; a < 4; ++a)
    for (int b2 = 0; b2 < 4; ++b2) {
      const int c = c0 + ty * 4 + a, cp = cp0 + tx * 4 + b2;
      if (c < C && cp < C) atomicAdd(gWx + (size_t)c * C + cp, acc[a][b2]);
    }
}

__global__ void k_transpose(int n, const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
    int xx = threadIdx.x;
    if (x0 + xx < n && y0 + yy < n) tile[yy][xx] = in[(size_t)(y0 + yy) * n + x0 + xx];
  }
  __syncthreads();
  for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
    int xx = threadIdx.x;
    if (y0 + xx < n && x0 + yy < n) out[(size_t)(x0 + yy) * n + y0 + xx] = tile[xx][yy];
  }
}

// ------------------------------------------------------------------------------------------
// attn_bwd: att (softmax over unmasked senders) backward, celu', and the W_s transpose.
// One warp per receiving atom.  g_s = att*(g_att - sum_j g_att*att) (the renormalisation of
// layers.py:180 is the identity on the gradient); g_q = g_s*celu'(q); g_e += Ws g_q.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attn_bwd(Dims d, const SakeLayerParams p, const float* __restrict__ x,
                                                  const float* __restrict__ e,
                                                  const float* __restrict__ att, float* __restrict__ gatt,
                                                  float* __restrict__ ge, float* __restrict__ gcut) {
  extern __shared__ float sm[];
  const int N = d.N, A = d.A, H = d.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = blockIdx.x * nw + warp;
  if (row >= d.R) return;
  float* as = sm + (size_t)warp * 2 * N * A;   // [N][A]
  float* gs = as + N * A;                      // [N][A]
  const size_t base = (size_t)row * N;
  for (int t = lane; t < N * A; t += 32) {
    as[t] = att[base * A + t];
    gs[t] = gatt[base * A + t];
  }
  __syncwarp();
  for (int a = 0; a < A; ++a) {
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s = fmaf(gs[j * A + a], as[j * A + a], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int j = lane; j < N; j += 32) gs[j * A + a] = as[j * A + a] * (gs[j * A + a] - s);    // g_s
  }
  __syncwarp();
  if (gcut != nullptr) {
    // cosine cutoff: cotangent of eps = sum over heads of g_s / eps (att is proportional to eps); see k_attn_bwd_tc
    const int mol0 = (row / N) * N;
    for (int j = lane; j < N; j += 32) {
      float t = 0.f;
      for (int a = 0; a < A; ++a) t += gs[j * A + a];
      gcut[base + j] = t * cosine_cutoff_dlog_(pair_dist_(x, row, mol0 + j), d.cut_lo, d.cut_hi);
    }
  }
  if (A == 4 && H == 64) {
    // lane owns f = 2*lane, 2*lane+1: q via a warp reduction of the per-lane partial dot products
    // (parameter pointers are only guaranteed 4-byte aligned: scalar loads)
    const float* wp = p.sem_kernel + (size_t)(2 * lane) * 4;
    const float4 w0 = make_float4(wp[0], wp[1], wp[2], wp[3]);
    const float4 w1 = make_float4(wp[4], wp[5], wp[6], wp[7]);
    const float b0 = p.sem_bias[0], b1 = p.sem_bias[1], b2 = p.sem_bias[2], b3 = p.sem_bias[3];
    for (int j = 0; j < N; ++j) {
      const float4 g4 = *reinterpret_cast<const float4*>(gs + j * 4);
      if (g4.x == 0.f && g4.y == 0.f && g4.z == 0.f && g4.w == 0.f) continue;   // masked / self pair: exact zero
      float2* gep = reinterpret_cast<float2*>(ge + (base + j) * 64 + 2 * lane);
      const float2 ev = *reinterpret_cast<const float2*>(e + (base + j) * 64 + 2 * lane);
      float q0 = ev.x * w0.x + ev.y * w1.x, q1 = ev.x * w0.y + ev.y * w1.y, q2 = ev.x * w0.z + ev.y * w1.z,
            q3 = ev.x * w0.w + ev.y * w1.w;
      for (int o = 16; o; o >>= 1) {
        q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        q2 += __shfl_xor_sync(0xffffffffu, q2, o); q3 += __shfl_xor_sync(0xffffffffu, q3, o);
      }
      const float gq0 = g4.x * dcelu2f_(q0 + b0), gq1 = g4.y * dcelu2f_(q1 + b1), gq2 = g4.z * dcelu2f_(q2 + b2),
                  gq3 = g4.w * dcelu2f_(q3 + b3);
      if (lane == 0) *reinterpret_cast<float4*>(gs + j * 4) = make_float4(gq0, gq1, gq2, gq3);
      float2 gv = *gep;
      gv.x += w0.x * gq0 + w0.y * gq1 + w0.z * gq2 + w0.w * gq3;
      gv.y += w1.x * gq0 + w1.y * gq1 + w1.z * gq2 + w1.w * gq3;
      *gep = gv;
    }
    __syncwarp();
  } else {
    for (int t = lane; t < N * A; t += 32) {
      const int j = t / A, a = t % A;
      float q = p.sem_bias[a];
      const float* er = e + (base + j) * H;
      for (int f = 0; f < H; ++f) q = fmaf(er[f], p.sem_kernel[(size_t)f * A + a], q);
      gs[t] *= dcelu2f_(q);
    }
    __syncwarp();
    for (int t = lane; t < N * H; t += 32) {
      const int j = t / H, f = t % H;
      float acc = 0.f;
      for (int a = 0; a < A; ++a) acc = fmaf(p.sem_kernel[(size_t)f * A + a], gs[j * A + a], acc);
      ge[(base + j) * H + f] += acc;
    }
  }
  for (int t = lane; t < N * A; t += 32) gatt[base * A + t] = gs[t];
}

// ------------------------------------------------------------------------------------------
// edge_bwd: persistent over receiving atoms; recomputes the edge MLP, back-propagates to the
// per-node projections, coordinates and the pair-level parameters.
// ------------------------------------------------------------------------------------------
struct EdgeBwdSmem {
  float *aW2, *aW1g, *aw1n, *ab2, *aWs, *abs_, *amu, *abeta;   // CTA-lifetime accumulators
  float *pri, *gui, *gpi, *dxi;                                // per receiving atom
  float *g, *rho, *u, *z1, *a1, *ges, *gz1, *gg, *rs, *ns, *ts; // per chunk
};
__host__ __device__ inline size_t edge_bwd_floats(const Dims& d) {
  return (size_t)d.H * d.H + (size_t)d.K * d.H + 2 * d.H + (size_t)d.H * d.A + d.A + 2 * d.K  // accumulators
         + d.NP + d.K + d.H + 4                                                                  // per row
         + (size_t)PJ * (4 * d.K + 4 * d.H + 3 + 2) + 16;
}
__device__ inline EdgeBwdSmem edge_bwd_carve(float* sm, const Dims& d) {
  EdgeBwdSmem s;
  const int H = d.H, K = d.K, A = d.A;
  s.aW2 = sm; s.aW1g = s.aW2 + H * H; s.aw1n = s.aW1g + K * H; s.ab2 = s.aw1n + H; s.aWs = s.ab2 + H;
  s.abs_ = s.aWs + H * A; s.amu = s.abs_ + A; s.abeta = s.amu + K;
  s.pri = s.abeta + K; s.gui = s.pri + d.NP; s.gpi = s.gui + K; s.dxi = s.gpi + H;
  s.g = s.dxi + 4; s.rho = s.g + PJ * K; s.u = s.rho + PJ * K; s.gg = s.u + PJ * K;
  s.z1 = s.gg + PJ * K; s.a1 = s.z1 + PJ * H; s.ges = s.a1 + PJ * H; s.gz1 = s.ges + PJ * H;
  s.rs = s.gz1 + PJ * H; s.ns = s.rs + PJ * 3; s.ts = s.ns + PJ;
  return s;
}

__global__ void __launch_bounds__(256) k_edge_bwd(Dims d, const float* __restrict__ x, const SakeLayerParams p,
                                                  const float* __restrict__ proj, const float* __restrict__ e,
                                                  const float* __restrict__ ge, const float* __restrict__ gq,
                                                  const float* __restrict__ gdir, float* __restrict__ gproj,
                                                  float* __restrict__ dx, SakeLayerGrads g, int want_grads,
                                                  const float* __restrict__ gcut, float* __restrict__ g_pair_u,
                                                  float* __restrict__ g_pair_p) {
  extern __shared__ float sm[];
  EdgeBwdSmem s = edge_bwd_carve(sm, d);
  const int H = d.H, K = d.K, A = d.A, N = d.N;
  const float* W1g = p.mlp_out0_kernel + (size_t)2 * H * H;
  const float* w1n = W1g + (size_t)K * H;
  const int nacc = H * H + K * H + 2 * H + H * A + A + 2 * K;
  for (int t = threadIdx.x; t < nacc; t += blockDim.x) s.aW2[t] = 0.f;
  __syncthreads();
  for (int row = blockIdx.x; row < d.R; row += gridDim.x) {
    const int b = row / N;
    for (int t = threadIdx.x; t < d.NP; t += blockDim.x) s.pri[t] = proj[(size_t)row * d.NP + t];
    for (int t = threadIdx.x; t < K + H + 4; t += blockDim.x) s.gui[t] = 0.f;   // gui, gpi, dxi contiguous
    const float xi0 = x[(size_t)row * 3 + 0], xi1 = x[(size_t)row * 3 + 1], xi2 = x[(size_t)row * 3 + 2];
    __syncthreads();
    for (int j0 = 0; j0 < N; j0 += PJ) {
      const int np = min(PJ, N - j0);
      if (threadIdx.x < np) {
        const int j = j0 + threadIdx.x;
        const float* xj = x + (size_t)(b * N + j) * 3;
        float r0 = xj[0] - xi0, r1 = xj[1] - xi1, r2 = xj[2] - xi2;
        s.rs[threadIdx.x * 3 + 0] = r0; s.rs[threadIdx.x * 3 + 1] = r1; s.rs[threadIdx.x * 3 + 2] = r2;
        float n = sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 0.f) + 1e-5f);
        s.ns[threadIdx.x] = n;
        s.ts[threadIdx.x] = expf(-n);
      }
      __syncthreads();
      for (int t = threadIdx.x; t < np * K; t += blockDim.x) {
        const int pj = t / K, k = t % K;
        const int j = j0 + pj;
        float dm = s.ts[pj] - p.rbf_means[k];
        float rho = expf(-p.rbf_betas[k] * dm * dm);
        float u = proj[(size_t)(b * N + j) * d.NP + k] + s.pri[d.Kp + k];
        if (d.pair_u) u += d.pair_u[((size_t)row * N + j) * d.Kp + k];     // edge features (SakePairTerms)
        s.rho[pj * K + k] = rho;
        s.u[pj * K + k] = u;
        s.g[pj * K + k] = rho * u;
      }
      __syncthreads();
      for (int t = threadIdx.x; t < np * H; t += blockDim.x) {
        const int pj = t / H, f = t % H;
        const int j = j0 + pj;
        float z = proj[(size_t)(b * N + j) * d.NP + 2 * d.Kp + f] + s.pri[2 * d.Kp + H + f];
        if (d.pair_p) z += d.pair_p[((size_t)row * N + j) * H + f];
        z = fmaf(s.ns[pj], w1n[f], z);
        for (int k = 0; k < K; ++k) z = fmaf(s.g[pj * K + k], W1g[(size_t)k * H + f], z);
        s.z1[pj * H + f] = z;
        s.a1[pj * H + f] = siluf_(z);
        s.ges[pj * H + f] = ge[((size_t)row * N + j) * H + f];
      }
      __syncthreads();
      for (int t = threadIdx.x; t < np * H; t += blockDim.x) {
        const int pj = t / H, q = t % H;
        float acc = 0.f;
        const float* w = p.mlp_out2_kernel + (size_t)q * H;
        for (int f = 0; f < H; ++f) acc = fmaf(w[f], s.ges[pj * H + f], acc);
        s.gz1[pj * H + q] = acc * dsiluf_(s.z1[pj * H + q]);
      }
      __syncthreads();
      if (want_grads) {
        for (int t = threadIdx.x; t < H * H; t += blockDim.x) {
          const int q = t / H, f = t % H;
          float acc = 0.f;
          for (int pj = 0; pj < np; ++pj) acc = fmaf(s.a1[pj * H + q], s.ges[pj * H + f], acc);
          s.aW2[t] += acc;
        }
        for (int t = threadIdx.x; t < K * H; t += blockDim.x) {
          const int k = t / H, f = t % H;
          float acc = 0.f;
          for (int pj = 0; pj < np; ++pj) acc = fmaf(s.g[pj * K + k], s.gz1[pj * H + f], acc);
          s.aW1g[t] += acc;
        }
        for (int f = threadIdx.x; f < H; f += blockDim.x) {
          float acc = 0.f, accb = 0.f;
          for (int pj = 0; pj < np; ++pj) {
            acc = fmaf(s.ns[pj], s.gz1[pj * H + f], acc);
            accb += s.ges[pj * H + f];
          }
          s.aw1n[f] += acc;
          s.ab2[f] += accb;
        }
        for (int t = threadIdx.x; t < H * A; t += blockDim.x) {
          const int f = t / A, a = t % A;
          float acc = 0.f;
          for (int pj = 0; pj < np; ++pj) {
            const size_t pr = (size_t)row * N + j0 + pj;
            acc = fmaf(e[pr * H + f], gq[pr * A + a], acc);
          }
          s.aWs[t] += acc;
        }
        for (int a = threadIdx.x; a < A; a += blockDim.x) {
          float acc = 0.f;
          for (int pj = 0; pj < np; ++pj) acc += gq[((size_t)row * N + j0 + pj) * A + a];
          s.abs_[a] += acc;
        }
      }
      // g_g = W1g @ g_z1
      for (int t = threadIdx.x; t < np * K; t += blockDim.x) {
        const int pj = t / K, k = t % K;
        float acc = 0.f;
        const float* w = W1g + (size_t)k * H;
        for (int f = 0; f < H; ++f) acc = fmaf(w[f], s.gz1[pj * H + f], acc);
        s.gg[pj * K + k] = acc;
      }
      __syncthreads();
      // per-k and per-f reductions / scatters
      for (int k = threadIdx.x; k < K + H; k += blockDim.x) {
        if (k < K) {
          const float mu = p.rbf_means[k], beta = p.rbf_betas[k];
          float gui = 0.f, gmu = 0.f, gbeta = 0.f;
          for (int pj = 0; pj < np; ++pj) {
            const int j = j0 + pj;
            const float ggv = s.gg[pj * K + k], rho = s.rho[pj * K + k];
            const float gu = ggv * rho;
            const float grho = ggv * s.u[pj * K + k];
            const float dm = s.ts[pj] - mu;
            atomicAdd(gproj + (size_t)(b * N + j) * d.NP + k, gu);
            if (g_pair_u) g_pair_u[((size_t)row * N + j) * d.Kp + k] = gu;
            gui += gu;
            const float w = dm * rho * grho;
            gmu = fmaf(2.0f * beta, w, gmu);
            gbeta = fmaf(-dm, w, gbeta);
            s.gg[pj * K + k] = -2.0f * beta * w;        // d/dt contribution
          }
          s.gui[k] += gui;
          if (want_grads) { s.amu[k] += gmu; s.abeta[k] += gbeta; }
        } else {
          const int f = k - K;
          float gpi = 0.f;
          for (int pj = 0; pj < np; ++pj) {
            const int j = j0 + pj;
            const float gz = s.gz1[pj * H + f];
            if (g_pair_p) g_pair_p[((size_t)row * N + j) * H + f] = gz;
            atomicAdd(gproj + (size_t)(b * N + j) * d.NP + 2 * d.Kp + f, gz);
            gpi += gz;
          }
          s.gpi[f] += gpi;
        }
      }
      __syncthreads();
      if (threadIdx.x < np) {
        const int pj = threadIdx.x, j = j0 + pj;
        float gt = 0.f;
        for (int k = 0; k < K; ++k) gt += s.gg[pj * K + k];
        float gn = -s.ts[pj] * gt;
        for (int f = 0; f < H; ++f) gn = fmaf(w1n[f], s.gz1[pj * H + f], gn);
        if (gcut) gn += gcut[(size_t)row * N + j];               // euclidean attention eps(n) (layers.py:172-176)
        const float n = s.ns[pj];
        const float r0 = s.rs[pj * 3 + 0], r1 = s.rs[pj * 3 + 1], r2 = s.rs[pj * 3 + 2];
        const float* gd = gdir + ((size_t)row * N + j) * 3;
        const float inv = 1.0f / (n + 1e-5f);
        float gr0 = gd[0] * inv, gr1 = gd[1] * inv, gr2 = gd[2] * inv;
        gn -= (gd[0] * r0 + gd[1] * r1 + gd[2] * r2) * inv * inv;
        const float n2 = r0 * r0 + r1 * r1 + r2 * r2;
        const float gn2 = n2 > 0.f ? gn / (2.0f * n) : 0.f;     // relu'(0) = 0 (functional.py:15)
        gr0 = fmaf(2.0f * r0, gn2, gr0); gr1 = fmaf(2.0f * r1, gn2, gr1); gr2 = fmaf(2.0f * r2, gn2, gr2);
        if (j == row % N) { gr0 = 0.f; gr1 = 0.f; gr2 = 0.f; }   // r_ii = x_i - x_i: +g and -g cancel exactly
        float* dxj = dx + (size_t)(b * N + j) * 3;
        atomicAdd(dxj + 0, gr0); atomicAdd(dxj + 1, gr1); atomicAdd(dxj + 2, gr2);
        atomicAdd(s.dxi + 0, -gr0); atomicAdd(s.dxi + 1, -gr1); atomicAdd(s.dxi + 2, -gr2);
      }
      __syncthreads();
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) gproj[(size_t)row * d.NP + d.Kp + k] = s.gui[k];
    for (int f = threadIdx.x; f < H; f += blockDim.x) gproj[(size_t)row * d.NP + 2 * d.Kp + H + f] = s.gpi[f];
    if (threadIdx.x < 3) atomicAdd(dx + (size_t)row * 3 + threadIdx.x, s.dxi[threadIdx.x]);
    __syncthreads();
  }
  if (want_grads) {
    for (int t = threadIdx.x; t < H * H; t += blockDim.x) atomicAdd(g.mlp_out2_kernel + t, s.aW2[t]);
    for (int t = threadIdx.x; t < K * H; t += blockDim.x)
      atomicAdd(g.mlp_out0_kernel + (size_t)2 * H * H + t, s.aW1g[t]);
    for (int f = threadIdx.x; f < H; f += blockDim.x) {
      atomicAdd(g.mlp_out0_kernel + (size_t)(2 * H + K) * H + f, s.aw1n[f]);
      atomicAdd(g.mlp_out2_bias + f, s.ab2[f]);
    }
    for (int t = threadIdx.x; t < H * A; t += blockDim.x) atomicAdd(g.sem_kernel + t, s.aWs[t]);
    for (int a = threadIdx.x; a < A; a += blockDim.x) atomicAdd(g.sem_bias + a, s.abs_[a]);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      atomicAdd(g.rbf_means + k, s.amu[k]);
      atomicAdd(g.rbf_betas + k, s.abeta[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// node_pre_bwd: cotangent of the per-node projections -> dh and W_in / W_1[0:2H] / biases
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_node_pre_bwd(Dims d, const SakeLayerParams p, NodeWT wt, const float* __restrict__ h,
                                                      const float* __restrict__ gproj, float* __restrict__ dh,
                                                      SakeLayerGrads g, int want_grads) {
  extern __shared__ float sm[];
  const int H = d.H, K = d.K, NP = d.NP, Kp = d.Kp;
  float* gp = sm;               // [NODES][NP]
  float* hs = gp + NODES * NP;  // [NODES][H]
  const int r0 = blockIdx.x * NODES;
  if (r0 >= dims_rows(d)) return;                  // ragged: the grid covers the padded worst case
  const int nn = min(NODES, dims_rows(d) - r0);
  for (int t = threadIdx.x; t < NODES * NP; t += blockDim.x) {
    const int nl = t / NP, o = t - nl * NP;
    gp[t] = nl < nn ? gproj[d.g8 ? g8_elem(r0 + nl, NP >> 2, o) : (size_t)r0 * NP + t] : 0.f;
  }
  for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) hs[t] = (t / H) < nn ? h[(size_t)r0 * H + t] : 0.f;
  __syncthreads();
  // dh[n][f] += sum_k Win[f][k] g_uj[k] + Win[H+f][k] g_ui[k] + sum_q W1[f][q] g_pj[q] + W1[H+f][q] g_pi[q]
  if (blockDim.x == 256 && H == 64 && (reinterpret_cast<uintptr_t>(wt.mlp_inT) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(wt.w1hT) & 15) == 0) {
    // four staged GEMMs over the blocks of the projection cotangent (the transposed copies are [rows][2H]:
    // sender half in columns [0,H), receiver half in [H,2H))
    float* wbuf = hs + NODES * H;                       // [2][PRE_WROWS][64]
    float* acc = wbuf + 2 * PRE_WROWS * 64;             // [NODES][H]
    for (int t = threadIdx.x; t < NODES * H; t += blockDim.x) acc[t] = 0.f;
    __syncthreads();
    auto add = [&](int n, int o4, const float* a) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[n * H + o4 + i] += a[i];
    };
    node_gemm64(gp, NP, K, wt.mlp_inT, 2 * H, wbuf, PRE_WROWS, add);
    node_gemm64(gp + Kp, NP, K, wt.mlp_inT + H, 2 * H, wbuf, PRE_WROWS, add);
    node_gemm64(gp + 2 * Kp, NP, H, wt.w1hT, 2 * H, wbuf, PRE_WROWS, add);
    node_gemm64(gp + 2 * Kp + H, NP, H, wt.w1hT + H, 2 * H, wbuf, PRE_WROWS, add);
    for (int t = threadIdx.x; t < nn * H; t += blockDim.x) dh[(size_t)r0 * H + t] += acc[t];
  } else {
    for (int idx = threadIdx.x; idx < H * (NODES / 2); idx += blockDim.x) {
      const int f = idx % H, n0 = (idx / H) * 2;
      const float* ga = gp + n0 * NP;
      const float* gb = ga + NP;
      float a0 = 0.f, a1 = 0.f;
      for (int k = 0; k < K; ++k) {
        const float wj = wt.mlp_inT[(size_t)k * 2 * H + f], wi = wt.mlp_inT[(size_t)k * 2 * H + H + f];
        a0 = fmaf(wj, ga[k], fmaf(wi, ga[Kp + k], a0));
        a1 = fmaf(wj, gb[k], fmaf(wi, gb[Kp + k], a1));
      }
      for (int q = 0; q < H; ++q) {
        const float vj = wt.w1hT[(size_t)q * 2 * H + f], vi = wt.w1hT[(size_t)q * 2 * H + H + f];
        a0 = fmaf(vj, ga[2 * Kp + q], fmaf(vi, ga[2 * Kp + H + q], a0));
        a1 = fmaf(vj, gb[2 * Kp + q], fmaf(vi, gb[2 * Kp + H + q], a1));
      }
      if (n0 < nn) dh[(size_t)(r0 + n0) * H + f] += a0;
      if (n0 + 1 < nn) dh[(size_t)(r0 + n0 + 1) * H + f] += a1;
    }
  }
  if (want_grads) {
    for (int t = threadIdx.x; t < H * K; t += blockDim.x) {
      const int f = t / K, k = t % K;
      float sj = 0.f, si = 0.f;
      for (int n = 0; n < nn; ++n) {
        sj = fmaf(hs[n * H + f], gp[n * NP + k], sj);
        si = fmaf(hs[n * H + f], gp[n * NP + Kp + k], si);
      }
      atomicAdd(g.mlp_in_kernel + t, sj);
      atomicAdd(g.mlp_in_kernel + (size_t)H * K + t, si);
    }
    for (int t = threadIdx.x; t < H * H; t += blockDim.x) {
      const int f = t / H, q = t % H;
      float sj = 0.f, si = 0.f;
      for (int n = 0; n < nn; ++n) {
        sj = fmaf(hs[n * H + f], gp[n * NP + 2 * Kp + q], sj);
        si = fmaf(hs[n * H + f], gp[n * NP + 2 * Kp + H + q], si);
      }
      atomicAdd(g.mlp_out0_kernel + t, sj);
      atomicAdd(g.mlp_out0_kernel + (size_t)H * H + t, si);
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float sb = 0.f;
      for (int n = 0; n < nn; ++n) sb += gp[n * NP + Kp + k];
      atomicAdd(g.mlp_in_bias + k, sb);
    }
    for (int q = threadIdx.x; q < H; q += blockDim.x) {
      float sb = 0.f;
      for (int n = 0; n < nn; ++n) sb += gp[n * NP + 2 * Kp + H + q];
      atomicAdd(g.mlp_out0_bias + q, sb);
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

static SakeLayerGrads null_grads() {
  SakeLayerGrads g;
  memset(&g, 0, sizeof(g));
  return g;
}

// transposed copies of the node-level weights (k_node_pre_bwd reads mlp_inT / w1hT from them)
size_t node_wt_bytes(const Dims& d) { return sizeof(float) * node_wt_floats(d); }
int gen_node_wt(const Dims& d, const SakeLayerParams& p, float* nodeWT, cudaStream_t st) {
  const size_t nwt = node_wt_floats(d);
  k_node_wt<<<(unsigned)((nwt + 255) / 256), 256, 0, st>>>(d, p, nodeWT);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int gen_node_post_bwd(const Dims& d, const SakeLayerParams& p, const float* h, const float* x, const float* v,
                      const float* mask, const Saved& sv, const float* dh_out, const float* dx_out,
                      const float* dv_out, float* dh, float* dx, float* dv, const SakeLayerGrads* g,
                      const BwdScratch& sc, cudaStream_t st) {
  (void)x;
  const size_t nwt = node_wt_floats(d);
  k_node_wt<<<(unsigned)((nwt + 255) / 256), 256, 0, st>>>(d, p, sc.nodeWT);
  size_t smem = node_post_bwd_smem_bytes(d);
  int rc;
  if ((rc = ensure_smem(k_node_post_bwd, smem))) return rc;
  k_node_post_bwd<<<(d.R + NODES - 1) / NODES, 256, smem, st>>>(d, p, carve_node_wt(d, sc.nodeWT), h, v, mask, sv.ssum,
                                                                sv.he, dh_out, dx_out, dv_out, dh, dx, dv, sc.T, sc.ghe,
                                                                g ? *g : null_grads(), g != nullptr, sc.nbuf, sc.tmax);
  note_launches(2);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// weight gradients of node_mlp / post_norm_mlp / velocity_mlp from the per-node record (K = nodes)
size_t tc_node_dw_scratch_bytes(const Dims& d) { return align_up(sizeof(float) * rows_pad128(d.R) * NB_LD) + 1024; }

// direct (the tcgen05 node kernels ran): the X operands are read where they live — h, saved.he and the forward
// kernel's stash (saved.nstash, row stride NS_LD) — and the record holds the cotangents and nrm only; otherwise
// (k_node_post_bwd, SAKE_NODE_TC=0) the record holds copies of everything.
int tc_node_dw(const Dims& d, const float* h, const Saved& sv, bool direct, const SakeLayerGrads& g, const BwdScratch& sc,
               XtgList& L, cudaStream_t st) {
  const float* nb = sc.nbuf;
  // a source: row-major (ld floats per row, tt = 0) or a field of a G8 buffer (tt = units per row; common.cuh)
  struct Src { const float* p; int ld, tt; };
  auto rowmajor = [](const float* p, int ld) { return Src{p, ld, 0}; };
  auto field = [](const float* buf, int col, int ld) { return Src{buf + (size_t)(col / 4) * G8S * 4, ld, ld / 4}; };
  auto call = [&](Src X, int xw, int ones, int mxpad, Src G, int gw, int ng, float* out, int ldo,
                  int out_rows, int out_cols, float* extra, int extra_ld) {
    XtgArgs q;
    memset(&q, 0, sizeof(q));
    q.X = X.p; q.ldx = X.ld; q.x_tt = X.tt; q.xw = xw; q.ones_col = ones;
    q.G = G.p; q.ldg = G.ld; q.g_tt = G.tt; q.gw = gw; q.MXpad = mxpad; q.NG = ng;
    q.P = d.R; q.Pdev = d.hdr ? &d.hdr->R64 : nullptr;
    q.out = out; q.ldo = ldo; q.out_rows = out_rows; q.out_cols = out_cols;
    q.extra = extra; q.extra_rows = extra ? 1 : 0; q.extra_ld = extra_ld;
    return L.push(q);
  };
  // record fields: G8 when the tcgen05 node kernel wrote them, rows of NB_LD floats otherwise
  auto rec = [&](int col) { return direct ? field(nb, col, NB_LD) : rowmajor(nb + col, NB_LD); };
  auto stash = [&](int col) { return field(sv.nstash, col, NS_LD); };
  int rc = 0;
  // node_mlp (layers.py:58-66)
  if (direct) {
    rc |= call(stash(NS_N1), 64, 64, 128, rec(NB_GT2), 64, 64, g.node2_kernel, 64, 64, 64, g.node2_bias, 64);
    rc |= call(rowmajor(h, 64), 64, -1, 128, rec(NB_GT1), 64, 64, g.node0_kernel, 64, 64, 64, nullptr, 64);
    rc |= call(field(sv.he, 0, 256), 256, -1, 256, rec(NB_GT1), 64, 64, g.node0_kernel + 64 * 64, 64, 256, 64, nullptr, 64);
    rc |= call(stash(NS_HCOMB), 64, 64, 128, rec(NB_GT1), 64, 64, g.node0_kernel + 320 * 64, 64, 64, 64, g.node0_bias, 64);
  } else {
    rc |= call(rec(NB_N1), 64, 64, 128, rec(NB_GT2), 64, 64, g.node2_kernel, 64, 64, 64, g.node2_bias, 64);
    rc |= call(rec(NB_CAT), 256, -1, 256, rec(NB_GT1), 64, 64, g.node0_kernel, 64, 256, 64, nullptr, 64);
    rc |= call(rec(NB_CAT + 256), 128, 128, 256, rec(NB_GT1), 64, 64, g.node0_kernel + 256 * 64, 64, 128, 64, g.node0_bias, 64);
  }
  if (d.spatial) {
    // post_norm_mlp (layers.py:85-92); the ones-row of the first call carries both bias gradients
    rc |= call(direct ? stash(NS_HP1) : rec(NB_HP1), 64, 64, 128, rec(NB_GTP2), 128, 128, g.post2_kernel, 64, 64, 64, g.post2_bias, 128);
    if (rc == 0) { L.a[L.n - 1].extra2 = g.post0_bias; L.a[L.n - 1].extra2_col0 = 64; }   // columns 64.. = sums of g_tp1
    rc |= call(rec(NB_NRM), 256, -1, 256, rec(NB_GTP1), 64, 64, g.post0_kernel, 64, 256, 64, nullptr, 64);
  }
  if (d.update && d.has_v) {
    // velocity_mlp (layers.py:69-76)
    rc |= call(direct ? stash(NS_HOUT) : rec(NB_HOUT), 64, 64, 128, rec(NB_GTV), 64, 64, g.vel0_kernel, 64, 64, 64, g.vel0_bias, 64);
    rc |= call(direct ? stash(NS_AV) : rec(NB_AV), 64, -1, 128, rec(NB_GY), 1, 16, g.vel2_kernel, 1, 64, 1, nullptr, 16);
  }
  if (rc) { set_error("xtg list full"); return SAKE_EINVAL; }
  return 0;
}

// Weight gradients of the separable per-node projections (mlp_in, rows [0, 2H) of mlp_out[0] and their biases:
// layers.py:19,22,30,33-38) as four more problems of the layer's batched X^T G launch: X = h [R, H], G = a column
// block of the projection cotangent gproj [R, NP] (layout: common.cuh).  k_node_pre_bwd used to accumulate them
// with one atomicAdd per (CTA, weight): 4.3 M atomics per layer at cfg2, 70 % of that kernel's time.
int tc_node_pre_dw(const Dims& d, const float* h, const SakeLayerGrads& g, const BwdScratch& sc, XtgList& L) {
  const int H = d.H, K = d.K, Kp = d.Kp;
  auto call = [&](int col0, int gw, float* out, int ldo, float* bias) {
    XtgArgs q;
    memset(&q, 0, sizeof(q));
    q.X = h; q.ldx = H; q.xw = H; q.ones_col = bias ? H : -1; q.MXpad = 128;
    q.G = sc.gproj + (size_t)(col0 / 4) * G8S * 4; q.ldg = d.NP; q.g_tt = d.NP / 4; q.gw = gw; q.NG = 64;   // G8 field
    q.P = d.R; q.Pdev = d.hdr ? &d.hdr->R64 : nullptr;
    q.out = out; q.ldo = ldo; q.out_rows = H; q.out_cols = gw;
    q.extra = bias; q.extra_rows = bias ? 1 : 0; q.extra_ld = gw;
    return L.push(q);
  };
  int rc = 0;
  rc |= call(0, K, g.mlp_in_kernel, K, nullptr);                               // sender half of mlp_in
  rc |= call(Kp, K, g.mlp_in_kernel + (size_t)H * K, K, g.mlp_in_bias);        // receiver half + bias
  rc |= call(2 * Kp, H, g.mlp_out0_kernel, H, nullptr);                        // sender half of mlp_out[0]
  rc |= call(2 * Kp + H, H, g.mlp_out0_kernel + (size_t)H * H, H, g.mlp_out0_bias);
  if (rc) { set_error("xtg list full"); return SAKE_EINVAL; }
  return 0;
}

int gen_mix_dw_from_gz(const Dims& d, const Saved& sv, const float* gZ, float* gWx, cudaStream_t st) {
  ProfScope prof(3, d.P, st);
  int splits = (int)min((long long)64, (d.P + 255) / 256);
  if (splits < 1) splits = 1;
  long long pps = (d.P + splits - 1) / splits;
  pps = (pps + 15) / 16 * 16;
  dim3 grid((d.C + 63) / 64, (d.C + 63) / 64, splits);
  k_mix_dw<<<grid, 256, 0, st>>>(d, sv.e, sv.att, gZ, gWx, pps);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int gen_mix_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                const BwdScratch& sc, float* gWx, cudaStream_t st) {
  int rc;
  if (!d.spatial) {
    k_mix_bwd_nospatial<<<(unsigned)d.R, 64, 0, st>>>(d, mask, sv.e, sv.att, sc.ghe, sc.ge, sc.gatt, sc.gdir);
    note_launches(1);
    SAKE_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  dim3 tb(32, 8), tg((d.C + 31) / 32, (d.C + 31) / 32);
  k_transpose<<<tg, tb, 0, st>>>(d.C, p.x_mixing_kernel, sc.wxT);
  size_t smem = sizeof(float) * (3 * d.C * MJ + 4 * d.C + MJ * 4);
  if ((rc = ensure_smem(k_mix_bwd, smem))) return rc;
  {
    ProfScope prof(2, d.P, st);
    k_mix_bwd<<<d.R, 256, smem, st>>>(d, x, mask, p.x_mixing_kernel, sc.wxT, sv.e, sv.att, sc.T, sc.ghe, sc.ge,
                                      sc.gatt, sc.gdir, gWx ? sc.gZ : nullptr);
  }
  note_launches(2);
  SAKE_CUDA_CHECK(cudaGetLastError());
  if (gWx) return gen_mix_dw_from_gz(d, sv, sc.gZ, gWx, st);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int gen_attn_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const Saved& sv, const BwdScratch& sc, cudaStream_t st) {
  int rc;
  int nw = 8;
  while (nw > 1 && sizeof(float) * 2 * d.N * d.A * nw > 160 * 1024) nw >>= 1;
  size_t smem = sizeof(float) * 2 * d.N * d.A * nw;
  if ((rc = ensure_smem(k_attn_bwd, smem))) return rc;
  k_attn_bwd<<<(d.R + nw - 1) / nw, nw * 32, smem, st>>>(d, p, x, sv.e, sv.att, sc.gatt, sc.ge, d.cutoff ? sc.gcut : nullptr);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int gen_edge_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const Saved& sv, float* dx,
                 const SakeLayerGrads* g, const BwdScratch& sc, float* g_pair_u, float* g_pair_p, cudaStream_t st) {
  int rc;
  SAKE_CUDA_CHECK(cudaMemsetAsync(sc.gproj, 0, sizeof(float) * (size_t)d.R * d.NP, st));
  if (g_pair_u) SAKE_CUDA_CHECK(cudaMemsetAsync(g_pair_u, 0, sizeof(float) * (size_t)d.P * d.Kp, st));   // padding columns
  size_t smem = sizeof(float) * edge_bwd_floats(d);
  if ((rc = ensure_smem(k_edge_bwd, smem))) return rc;
  int grid = d.R < 148 * 2 ? d.R : 148 * 2;
  k_edge_bwd<<<grid, 256, smem, st>>>(d, x, p, sv.nodeproj, sv.e, sc.ge, sc.gatt, sc.gdir, sc.gproj, dx,
                                      g ? *g : null_grads(), g != nullptr, d.cutoff ? sc.gcut : nullptr, g_pair_u, g_pair_p);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// g == NULL: input cotangent dh only (the tcgen05 engines get the parameter gradients from tc_node_pre_dw)
int gen_node_pre_bwd(const Dims& d, const SakeLayerParams& p, const float* h, float* dh, const SakeLayerGrads* g,
                     const BwdScratch& sc, cudaStream_t st) {
  int rc;
  size_t smem = sizeof(float) * (NODES * (d.NP + 2 * d.H) + 2 * PRE_WROWS * 64 + 16);
  if ((rc = ensure_smem(k_node_pre_bwd, smem))) return rc;
  ProfScope prof(12, d.R, st);
  k_node_pre_bwd<<<(d.R + NODES - 1) / NODES, 256, smem, st>>>(d, p, carve_node_wt(d, sc.nodeWT), h, sc.gproj, dh,
                                                               g ? *g : null_grads(), g != nullptr);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sake
