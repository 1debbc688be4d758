// Shared definitions for libsake_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sake_b200.h"

namespace sake {

// ---- ragged batches (real atoms only) ---------------------------------------------------------------
// A padded batch (QM9-style, scripts/qm9/run.py:23-24,35: molecule b has n_real[b] <= N real atoms stored
// first, mask = outer(m, m)) is processed in a COMPACT layout that holds the real atoms only: molecules are
// ordered by n_real (stable), molecule b owns n_b consecutive rows and row r owns n consecutive pair slots.
// The tables below are built on the device from n_real (ragged.cu: sake_ragged_prepare) and every kernel
// reads its loop bounds from the header, so the host never needs the counts (no sync, CUDA-graph replays
// stay valid when the next batch has different n_real); grids are sized for the padded worst case.
struct RaggedHdr {
  int R;            // real rows (atoms)
  int num_tiles;    // 128-pair tiles
  int B;            // molecules
  int reserved;
  long long P;      // real pairs = sum n_b^2
  long long R64;    // = R (64-bit copy: the K extent of the node-level weight-gradient contractions)
};
// rowinfo[r]  = {first row of r's molecule, n of that molecule, first pair slot of row r, padded row b*N+i}
// tileinfo[t] = {first row, rows in the tile, n (uniform inside a tile), first pair slot}
// molinfo[b]  = {first row of molecule b, n_b}   (b = index in the padded batch)

// Resolved problem description handed to every kernel by value.
struct Dims {
  int B, N, H, A, K, C;   // C = A*H
  int R;                  // rows = B*N (one row per receiving atom i); ragged: the padded worst case
  long long P;            // pairs = R*N                                  ; ragged: the padded worst case
  int Kp;                 // K rounded up to a multiple of 4 (16-byte aligned projection blocks)
  int NP;                 // per-node projection width = 2Kp + 2H
  int update, has_v, has_mask, spatial;
  int g8;                 // 1 (tcgen05 engines): the internal per-pair / per-atom buffers e, he, nodeproj, ge, gproj, gZ and the
                          // backward records use the G8 layout below; 0 (generic engine): row-major rows
  int prepared;           // 1: weight operand images in `saved` are current (SAKE_WEIGHTS_PREPARED)
  int cutoff;             // 1: cosine cutoff on the attention (layers.py:172-176), parameters below
  float cut_lo, cut_hi;
  const float *pair_u, *pair_p;   // `he` edge features as per-pair additive terms [P,Kp], [P,H] (SakePairTerms) or NULL
  const RaggedHdr* hdr;   // NULL: uniform batch (every molecule has N atoms, optional float mask)
  const int4* rowinfo;
  const int4* tileinfo;
  const int2* molinfo;
};
// rows / pairs actually present (device side)
__device__ __forceinline__ int dims_rows(const Dims& d) { return d.hdr ? d.hdr->R : d.R; }
__device__ __forceinline__ long long dims_pairs(const Dims& d) { return d.hdr ? d.hdr->P : d.P; }
struct RowInfo { int mol0, n; long long pair0; };
__device__ __forceinline__ RowInfo row_info(const Dims& d, int row) {
  RowInfo r;
  if (d.rowinfo) { const int4 q = __ldg(d.rowinfo + row); r.mol0 = q.x; r.n = q.y; r.pair0 = q.z; }
  else { r.mol0 = (row / d.N) * d.N; r.n = d.N; r.pair0 = (long long)row * d.N; }
  return r;
}

// Layout of the per-node projection buffer nodeproj[R][NP]:
//   [0,K)            uj = h @ W_in[0:H]            (sender part of mlp_in,   layers.py:30)
//   [Kp,Kp+K)        ui = h @ W_in[H:2H] + b_in    (receiver part)
//   [2Kp,2Kp+H)      pj = h @ W_1[0:H]             (sender part of mlp_out[0], layers.py:33-38)
//   [2Kp+H,2Kp+2H)   pi = h @ W_1[H:2H] + b_1      (receiver part)
// (Kp = K rounded up to 4; the padding slots are zero)
// (get_h_cat_ht, functional.py:33-44, is never materialised: Dense on [h_j | h_i] is separable.)

// `saved` buffer (fwd -> bwd), all fp32:
struct Saved {
  float* e;         // [P,H]   edge features h_e_mtx (layers.py:204)
  float* att;       // [P,A]   combined attention (layers.py:205)
  float* logit;     // [P,A]   attention logits s (layers.py:155-165); aliases att on the generic engine (normalised in place)
  float* ssum;      // [R,C,3] sum_j dir*coef*mask (numerator of combinations_sum, layers.py:123,127)
  int ssum_tt;      // ssum is in the G8 layout, unit (c/4)*3 + d = component d of four coefficients (tcgen05 engines)
  float* he;        // [R,C]   aggregate (layers.py:135-140)
  float* nodeproj;  // [R,NP]
  float* nstash;    // [R, NS_LD] in the G8 layout per-atom activations of the node tail (tc_node.cu forward -> backward: no recompute)
  // tcgen05 engines: operand images of the layer's weights, built ONCE by the forward call and reused by the
  // backward call of the same step (round 1 rebuilt them in both: 5 + 5 small launches per layer)
  void* wmix;       // x_mixing images for GEMM1 / GEMM2 + {scale, 1/scale}   (tc_mix.cu)
  void* wedge;      // edge-model images WA..WD + vectors                      (tc_edge.cu)
  void* wnode;      // node-tail chunk images, forward and transposed blocks   (tc_node.cu)
  float* nodeWT;    // transposed mlp_in / mlp_out[0] halves for k_node_pre_bwd (generic_bwd.cu)
};
// nstash row: silu' of the five hidden layers (tp1 tp2 t1 t2 tv: 5 x 64), then the activations the weight-gradient
// record needs (hp1, hcomb, n1, av, h': 5 x 64) and the velocity-gate logit y
enum { NS_D = 0, NS_HP1 = 320, NS_HCOMB = 384, NS_N1 = 448, NS_AV = 512, NS_HOUT = 576, NS_Y = 640, NS_LD = 648 };

// `scratch` buffer for the backward pass
struct BwdScratch {
  float* T;         // [R,C,4] cotangent of ssum (float4 per coefficient, w unused)
  float* tmax;      // [R]     max |T| per receiver (row scale of the fp16-split backward operand)
  float* ghe;       // [R,C]   cotangent of he
  float* ge;        // [P,H]   cotangent of e
  float* gatt;      // [P,A]   cotangent of att, then of the pre-celu logits q
  float* gdir;      // [P,3]   cotangent of the unit direction
  float* gproj;     // [R,NP]  cotangent of nodeproj
  float* wxT;       // [C,C]   x_mixing kernel transposed
  float* gZ;        // [P,C]   cotangent of pre-tanh coefficients (training only, feeds the dW GEMM)
  float* gcut;      // [P]     cotangent of the pair distance through the cosine cutoff (NULL without cutoff)
  float* nodeWT;       // transposed copies of the node-level weight matrices (k_node_wt)
  float* nbuf;         // per-node record for the node-level weight-gradient contractions (tcgen05 engines, training)
  float* qv;           // [R,4] g_dv / den2 per atom (tcgen05 node kernel -> v_mixing gradient in k_pair_reduce), or NULL
  float* xtg_partial;  // per-CTA partial sums of the weight-gradient contractions (tcgen05 engines, training)
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
// ---- "G8" layout of per-row buffers (rows = atoms or atom pairs, U 16-byte units per row) ----------------------
// The tcgen05 kernels are thread-per-row: a thread walks the units of ITS row, so with row-major rows the 32 lanes
// of a load touch 32 different 128-byte lines (32 L1 wavefronts per instruction; ncu: the edge and node kernels sat
// at the L1 wavefront ceiling).  G8 interleaves groups of 8 consecutive rows unit by unit:
//     float4 index of (row r, unit u) = ((r >> 3) * U + u) * 8 + (r & 7)
// so that 8 consecutive lanes read one full line (4 wavefronts per 128-bit load, the minimum), a whole 8-row group
// is still one contiguous block (warp-per-row kernels read groups with fully coalesced loads), and the X^T G builder
// (tc_xtg.cu: 8 lanes = the 8 units of a row) touches as many lines as with row-major rows.
constexpr int G8S = 8;                                       // float4 stride between consecutive units of a row
__host__ __device__ inline size_t g8_row(long long r, int U) { return (size_t)(r >> 3) * U * 8 + (size_t)(r & 7); }
__host__ __device__ inline size_t g8_elem(long long r, int U, int col) {   // float index of column col of row r
  return (g8_row(r, U) + (size_t)(col >> 2) * G8S) * 4 + (size_t)(col & 3);
}
inline size_t rows_pad8(long long R) { return (size_t)((R + 7) / 8 * 8); }
inline size_t rows_pad128(long long R) { return (size_t)((R + 127) / 128 * 128); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d silu / dx
__device__ __forceinline__ float dsiluf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}
// sake.utils.cosine_cutoff (utils.py:10-26; the range masks computed there are discarded, so this is the whole
// function): eps(d) = 0.5 (cos(pi (2 (d - lo) / (hi - lo) + 1)) + 1) = 0.5 (1 - cos(theta)), theta = 2 pi (d - lo) / (hi - lo)
__device__ __forceinline__ float cosine_cutoff_(float d, float lo, float hi) {
  return 0.5f * (cosf(3.14159265358979323846f * (2.0f * (d - lo) / (hi - lo) + 1.0f)) + 1.0f);
}
// d eps / dd divided by eps: (2 pi / (hi - lo)) * sin(theta) / (1 - cos(theta)); 0 where eps = 0
__device__ __forceinline__ float cosine_cutoff_dlog_(float d, float lo, float hi) {
  const float th = 6.28318530717958647692f * (d - lo) / (hi - lo);
  const float den = 1.0f - cosf(th);
  return den > 0.f ? (6.28318530717958647692f / (hi - lo)) * sinf(th) / den : 0.f;
}
// pair distance as the layer defines it (functional.py:14-17)
__device__ __forceinline__ float pair_dist_(const float* __restrict__ x, int i, int j) {
  const float r0 = x[(size_t)j * 3] - x[(size_t)i * 3], r1 = x[(size_t)j * 3 + 1] - x[(size_t)i * 3 + 1],
              r2 = x[(size_t)j * 3 + 2] - x[(size_t)i * 3 + 2];
  return sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 0.f) + 1e-5f);
}
// nn.celu(alpha=2): max(x,0) + 2*expm1(min(x,0)/2)   (layers.py:81)
__device__ __forceinline__ float celu2f_(float x) { return x > 0.f ? x : 2.0f * expm1f(0.5f * x); }
__device__ __forceinline__ float dcelu2f_(float x) { return x > 0.f ? 1.0f : expf(0.5f * x); }

constexpr int SAKE_NODES = 16;  // nodes per CTA in the per-node kernels

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// acc[n][o] = sum_i x[n][i] * W[i][o] for SAKE_NODES nodes and ONE 64-column block of W (row stride ldw):
// the weight rows go through shared memory in chunks of `wrows` rows (cp.async, double-buffered), so every
// CTA reads a weight matrix from L2 once instead of once per warp (the node kernels were L2-bound on that).
// 256 threads: warp = node pair, lane = (4 output columns og, K half ks); the halves meet by one shuffle.
// epi(node, first column within the block, acc[4]) runs on the ks = 0 lanes.  Ends with __syncthreads().
template <class Epi>
__device__ __forceinline__ void node_gemm64(const float* x, int ldx, int in, const float* __restrict__ W, int ldw,
                                            float* wbuf, int wrows, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int og = lane & 15, ks = lane >> 4;
  const float* x0 = x + (2 * warp) * ldx;
  const float* x1 = x0 + ldx;
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  const int nchunk = (in + wrows - 1) / wrows, half = wrows / 2;
  auto prefetch = [&](int c) {
    float* dst = wbuf + (c & 1) * wrows * 64;
    const int r0 = c * wrows;
    for (int t = threadIdx.x; t < wrows * 16; t += blockDim.x) {
      const int r = t >> 4, q = t & 15;
      if (r0 + r < in) cp_async16(dst + r * 64 + q * 4, W + (size_t)(r0 + r) * ldw + q * 4);
    }
    cp_async_commit();
  };
  prefetch(0);
  for (int c = 0; c < nchunk; ++c) {
    cp_async_wait_all();
    __syncthreads();                       // chunk c landed; everyone is done with the buffer chunk c+1 goes into
    if (c + 1 < nchunk) prefetch(c + 1);
    const float* wb = wbuf + (c & 1) * wrows * 64 + 4 * og;
    const int rows = min(wrows, in - c * wrows);
    const int kb = ks * half, ke = min(rows, kb + half);
    const float* xa = x0 + c * wrows;
    const float* xb = x1 + c * wrows;
#pragma unroll 8
    for (int k = kb; k < ke; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(wb + k * 64);
      const float v0 = xa[k], v1 = xb[k];
      a0[0] = fmaf(v0, w.x, a0[0]); a0[1] = fmaf(v0, w.y, a0[1]); a0[2] = fmaf(v0, w.z, a0[2]); a0[3] = fmaf(v0, w.w, a0[3]);
      a1[0] = fmaf(v1, w.x, a1[0]); a1[1] = fmaf(v1, w.y, a1[1]); a1[2] = fmaf(v1, w.z, a1[2]); a1[3] = fmaf(v1, w.w, a1[3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], 16);
    a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], 16);
  }
  if (ks == 0) { epi(2 * warp, 4 * og, a0); epi(2 * warp + 1, 4 * og, a1); }
  __syncthreads();                         // the weight buffers are free again
}

// y[n][o] = bias[o] + sum_i x[n][i] * W[i][o]      (W row-major [in][out])
// Fast path (wbuf != NULL, 256 threads, out a multiple of 64, W 16-byte aligned): node_gemm64 per 64-column block.
// Generic path: work item = (output o, pair of nodes), weights read from global (coalesced over o).
__device__ __forceinline__ void node_dense(float* y, const float* x, int ldx, int in, const float* __restrict__ W,
                                           const float* __restrict__ bias, int out, bool accumulate,
                                           float* wbuf = nullptr, int wrows = 0) {
  if (wbuf != nullptr && blockDim.x == 256 && (out & 63) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    for (int ob = 0; ob < out; ob += 64)
      node_gemm64(x, ldx, in, W + ob, out, wbuf, wrows, [&](int n, int o4, const float* a) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = ob + o4 + i;
          y[n * out + o] = a[i] + (accumulate ? y[n * out + o] : (bias ? bias[o] : 0.f));
        }
      });
    return;
  }
  for (int idx = threadIdx.x; idx < out * (SAKE_NODES / 2); idx += blockDim.x) {
    const int o = idx % out, n0 = (idx / out) * 2;
    float a0 = accumulate ? y[n0 * out + o] : (bias ? bias[o] : 0.f);
    float a1 = accumulate ? y[(n0 + 1) * out + o] : (bias ? bias[o] : 0.f);
    const float* x0 = x + n0 * ldx;
    const float* x1 = x0 + ldx;
#pragma unroll 16
    for (int i = 0; i < in; ++i) {      // 16 independent weight loads in flight (L2-latency bound otherwise)
      const float w = W[(size_t)i * out + o];
      a0 = fmaf(x0[i], w, a0);
      a1 = fmaf(x1[i], w, a1);
    }
    y[n0 * out + o] = a0;
    y[(n0 + 1) * out + o] = a1;
  }
  __syncthreads();
}

// ---- fast fp32-class transcendentals for the tensor-core kernels --------------------------------
// One MUFU each (ex2.approx / rcp.approx, <= 2 ulp) instead of the ~25-instruction libdevice paths.
// Errors stay at the 1e-7 .. 1e-6 level (absolute for tanh/sigmoid, relative ~|x|*2^-24 for exp), well
// inside the 1e-5 energy / 1e-4 force tolerance — and the epilogues are issue-bound, not MUFU-bound.
// (raw ex2.approx / rcp.approx with flush-to-zero: none of __expf's / __fdividef's range fix-ups; e^x -> 0 or
// inf outside the fp32 range, 1/inf -> 0, which is what the callers need)
__device__ __forceinline__ float fex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float frcpa_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fexp_(float x) { return fex2_(x * 1.4426950408889634f); }
__device__ __forceinline__ float frcp_(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float fsigmoid_(float x) { return frcpa_(1.0f + fex2_(x * -1.4426950408889634f)); }
__device__ __forceinline__ float fsilu_(float x) { return x * fsigmoid_(x); }
__device__ __forceinline__ float fdsilu_(float x) {
  const float s = fsigmoid_(x);
  return s * (1.0f + x * (1.0f - s));
}
// tanh(x) = sign(x) * (1 - 2/(exp(2|x|)+1)); the exponent is >= 0, so the raw ex2.approx needs none of
// __expf's denormal-range handling; overflow -> rcp(inf) = 0 -> +-1.  6 instructions, 2 MUFU.
__device__ __forceinline__ float ftanh_(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return copysignf(fmaf(-2.0f, r, 1.0f), x);
}
// single-MUFU tanh (abs error ~5e-4): only for the bf16 engine, whose operands are no more precise
__device__ __forceinline__ float ftanh_mufu_(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

// tanh and its derivative sech^2 = 4t/(t+1)^2 (t = e^{2|x|}) from ONE exponential; computing the
// derivative directly avoids the 1 - tanh^2 cancellation in saturated coefficients.
__device__ __forceinline__ void ftanh_sech2_(float x, float& th, float& s2) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fminf(fabsf(x), 40.0f) * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  th = copysignf(fmaf(-2.0f, r, 1.0f), x);
  s2 = 4.0f * t * r * r;
}

void set_error(const char* fmt, ...);
void note_launches(int n);
void ragged_attach(Dims& d, const void* blob);   // ragged.cu: point d.hdr / rowinfo / tileinfo / molinfo into a table blob

#define SAKE_CUDA_CHECK(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      sake::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SAKE_ECUDA;                                                             \
    }                                                                                \
  } while (0)

// The > 48 KB dynamic shared-memory opt-in is a per-DEVICE function attribute: every launcher sets it once
// per (kernel, device) — `done` is the launcher's own static bit mask, one bit per device ordinal.  A process
// that drives several GPUs (one host thread per device, as XLA's pmap does) therefore opts in on each of them.
template <typename Kern>
static inline int smem_optin(Kern kern, size_t smem, unsigned long long& done) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  const unsigned long long bit = 1ull << (dev & 63);
  if ((__atomic_load_n(&done, __ATOMIC_RELAXED) & bit) == 0) {
    SAKE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    __atomic_fetch_or(&done, bit, __ATOMIC_RELAXED);
  }
  return 0;
}

// event profiler hooks around the dominant kernels (optim.cu)
bool prof_begin_launch(int kind, long long pairs, cudaStream_t st);
void prof_end_launch(cudaStream_t st);
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(int kind, long long pairs, cudaStream_t s) : st(s), on(kind > 0 && prof_begin_launch(kind, pairs, s)) {}
  ~ProfScope() { if (on) prof_end_launch(st); }
};

// D[MX x NG] += sum_p X[p][:]^T G[p][:]  (tc_xtg.cu) — every weight-gradient contraction (K = pairs or nodes)
struct XtgArgs {
  const float* X; int ldx; int xw;     // X source [P, xw] row-major fp32 (ignored when e != nullptr)
  const float* e; const float* att;    // X = e (x) att  (xw = 256, feature c = f*4 + head) when e != nullptr
  int ones_col;                        // X feature index forced to 1.0 (column sums of G for free), or -1
  const float* G; int ldg; int gw;     // G source [P, gw]
  int x_tt, g_tt;                      // > 0: the source is in the G8 layout with this many 16-byte units per row
  int MXpad, NG;                       // operand image sizes: MXpad in {128,256}; NG multiple of 16, <= 256
  long long P, pairs_per_cta;          // ragged batches: P is the padded worst case, the real extent is *Pdev
  const long long* Pdev;               // device-resident K extent (RaggedHdr::P or ::R64), or NULL
  float* out; int ldo, out_rows, out_cols;   // out[r][c] += D[r][c], r < out_rows, c < out_cols
  float* extra; int extra_rows, extra_ld;    // extra[r - out_rows][c] += D[r][c] for the next extra_rows rows
  float* extra2; int extra2_col0;            // single extra row split over two vectors: columns >= extra2_col0 go to extra2
  // RBF mean / width gradients folded into the reduction (tc_edge.cu): the two rows behind out_rows hold, for column
  // 128 + k, S1 = sum_p w_pk and S2 = sum_p t_p w_pk:  dmu_k += 2 beta_k S1,  dbeta_k -= S2 - mu_k S1
  const float *mb_mu, *mb_beta; float *mb_gmu, *mb_gbeta; int mb_K;
  float* partial;                            // scratch for per-CTA partial sums (tc_xtg_partial_bytes()); NULL = atomics
  int gx;                                    // CTAs assigned to this problem (set by the launcher)
};
// All weight-gradient contractions of one layer backward are collected and run as ONE batched launch
// (+ one reduction launch): blockIdx.y selects the problem.
struct XtgList {
  static constexpr int MAXP = 20;
  XtgArgs a[MAXP];
  int n = 0;
  int push(const XtgArgs& q) { if (n >= MAXP) return -1; a[n++] = q; return 0; }
};
// red_st != st: the reduction runs on red_st, forked from st with `fork` (the caller joins red_st later)
int tc_xtg_flush(XtgList& L, float* partial, int engine, int prof_kind, cudaStream_t st, cudaStream_t red_st = nullptr,
                 cudaEvent_t fork = nullptr);
int tc_xtg(const XtgArgs& a, int engine, int prof_kind, cudaStream_t st);
size_t tc_xtg_partial_bytes();

// ---- generic fp32 engine ------------------------------------------------------------------
int gen_node_pre(const Dims& d, const SakeLayerParams& p, const float* h, const Saved& sv, cudaStream_t st);
int gen_edge_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                 cudaStream_t st);
int gen_attn_fwd(const Dims& d, const float* x, const float* mask, const Saved& sv, cudaStream_t st);
// tc_node.cu: per-node tail of the layer on the tensor cores (H = 64, A = 4)
bool tc_node_supported(const Dims& d);
int gen_node_wt(const Dims& d, const SakeLayerParams& p, float* nodeWT, cudaStream_t st);
size_t node_wt_bytes(const Dims& d);
size_t tc_node_w_bytes();
size_t tc_node_bwd_scratch_bytes(const Dims& d);
int tc_node_post_bwd(const Dims& d, const SakeLayerParams& p, const float* h, const float* v, const float* mask,
                     const Saved& sv, const float* dh_out, const float* dx_out, const float* dv_out, float* dh, float* dx,
                     float* dv, const SakeLayerGrads* g, const BwdScratch& sc, void* wscratch, void* nscratch,
                     cudaStream_t st);
int tc_node_post(const Dims& d, const SakeLayerParams& p, const float* h, const float* x, const float* v,
                 const float* mask, float* h_out, float* x_out, float* v_out, const Saved& sv, void* wscratch,
                 cudaStream_t st);
int tc_attn_bwd(const Dims& d, const float* x, const Saved& sv, const BwdScratch& sc, cudaStream_t st);
int gen_mix_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                cudaStream_t st);
int gen_node_post(const Dims& d, const SakeLayerParams& p, const float* h, const float* x, const float* v,
                  const float* mask, float* h_out, float* x_out, float* v_out, const Saved& sv,
                  cudaStream_t st);
// backward, in execution order
int gen_node_post_bwd(const Dims& d, const SakeLayerParams& p, const float* h, const float* x, const float* v,
                      const float* mask, const Saved& sv, const float* dh_out, const float* dx_out,
                      const float* dv_out, float* dh, float* dx, float* dv, const SakeLayerGrads* g,
                      const BwdScratch& sc, cudaStream_t st);
size_t tc_node_dw_scratch_bytes(const Dims& d);
int tc_node_dw(const Dims& d, const float* h, const Saved& sv, bool direct, const SakeLayerGrads& g, const BwdScratch& sc,
               XtgList& L, cudaStream_t st);
int tc_node_pre_dw(const Dims& d, const float* h, const SakeLayerGrads& g, const BwdScratch& sc, XtgList& L);
int gen_mix_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                const BwdScratch& sc, float* gWx, cudaStream_t st);
int gen_attn_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const Saved& sv, const BwdScratch& sc, cudaStream_t st);
int gen_edge_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const Saved& sv, float* dx,
                 const SakeLayerGrads* g, const BwdScratch& sc, float* g_pair_u, float* g_pair_p, cudaStream_t st);
int gen_node_pre_bwd(const Dims& d, const SakeLayerParams& p, const float* h, float* dh, const SakeLayerGrads* g,
                     const BwdScratch& sc, cudaStream_t st);

// ---- tcgen05 engine (x_mixing GEMM family) -------------------------------------------------
// mix forward: ssum[R,C,3] from e, att, x (replaces the generic k_mix_fwd)
int tc_mix_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask,
               const Saved& sv, void* tc_scratch, int engine, cudaStream_t st);
// weight operand images (built by sake_layer_prepare, or by the forward call when d.prepared == 0)
int tc_mix_prepare(const SakeLayerParams& p, void* wmix, int engine, cudaStream_t st);
int tc_edge_prepare(const Dims& d, const SakeLayerParams& p, void* wedge, cudaStream_t st);
int tc_node_prepare(const Dims& d, const SakeLayerParams& p, void* wnode, cudaStream_t st);
int tc_node_pre(const Dims& d, const SakeLayerParams& p, const float* h, const Saved& sv, cudaStream_t st);
int tc_node_pre_bwd(const Dims& d, const Saved& sv, const BwdScratch& sc, float* dh, cudaStream_t st);
// mix backward: ge, gatt, gdir (and dWx when gWx != nullptr)
int tc_mix_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask,
               const Saved& sv, const BwdScratch& sc, float* gWx, void* tc_scratch, int engine, XtgList& L,
               cudaStream_t st);      // tc_scratch: the images tc_mix_fwd built (saved.wmix)
size_t tc_scratch_bytes(const Dims& d, int engine, int for_backward, int with_param_grads);


bool tc_supported(const Dims& d);

// ---- tcgen05 engine (edge model), tc_edge.cu ---------------------------------------------------
bool tc_edge_supported(const Dims& d);
size_t edge_w_bytes();
size_t tc_edge_bwd_scratch_bytes(const Dims& d, int with_grads);
int tc_edge_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                void* wscratch, cudaStream_t st);
int tc_edge_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                const BwdScratch& sc, float* dx, const SakeLayerGrads* g, void* wscratch, void* escratch, XtgList& L,
                float* g_pair_u, float* g_pair_p, cudaStream_t st);

}  // namespace sake
