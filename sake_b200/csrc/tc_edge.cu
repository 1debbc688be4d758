// tcgen05 engine for the edge model of DenseSAKELayer (sake/layers.py:12-40,153-165): pair geometry,
// ExpNormalSmearing RBFs (sake/utils.py:61-65), the two edge-MLP GEMMs and the attention logits,
// forward and backward, on 128-pair tiles.  One thread owns one pair (= one TMEM lane), so every
// per-pair chain (RBF, silu, celu, the scalar geometry backward) is thread-local; the GEMMs
//   Z1 = [rho*u | n] W1[2H:]          (K = 64: 50 RBF channels + distance, zero padded)
//   E' = silu(z1) [W2 | W2 Ws]        (edge features and pre-activation attention logits in one MMA)
//   GA1 = g_e W2^T ,  GG = g_z1 W1[2H:]^T          (backward)
// run on the tensor cores in 3xTF32 (fp32-class accuracy) with the small weight images resident in
// shared memory.  Two 128-thread groups per CTA work on alternate tiles so that one group's MMA /
// TMEM traffic overlaps the other group's transcendental work.  h_cat_ht (functional.py:33-44) is
// never formed: Dense on [h_j | h_i] is separable and arrives as per-node projections.
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_tile.cuh"

namespace sake {
using namespace tc;

constexpr int EP_IMG = TILE * 128;          // one chunk image, one split (16 KB)
constexpr int EG_IMG = 2 * EP_IMG;          // pair-side image of one group: ONE K-chunk x {hi, lo}
constexpr int WA_BYTES = 2 * 2 * 64 * 128;  // rows f  (64), K = k
constexpr int WB_BYTES = 2 * 2 * 80 * 128;  // rows [e(64) | q(4) | 0(12)], K = f'
constexpr int WC_BYTES = 2 * 2 * 64 * 128;  // rows f' (64), K = f
constexpr int WD_BYTES = 2 * 2 * 64 * 128;  // rows k  (64), K = f
constexpr int EVEC = 512;                   // floats: mu[64] beta[64] b2[64] bq[4] ... | Ws[64][4] at 256
constexpr int PB_LD = 192;                  // per-pair backward record: gz1[64] | gu[<=60] | g_r[124..126] | w[128..]
constexpr int EDGE_GROUPS = 3;              // independent 128-thread groups (tiles in flight) per CTA
constexpr int EDGE_THREADS = 128 * EDGE_GROUPS;

struct EdgeW {
  uint8_t *WA, *WB, *WC, *WD;
  float* vec;
};
size_t edge_w_bytes() { return WA_BYTES + WB_BYTES + WC_BYTES + WD_BYTES + EVEC * sizeof(float) + 1024; }
static EdgeW carve_edge_w(void* p) {
  EdgeW w;
  uint8_t* b = (uint8_t*)p;
  w.WA = b; w.WB = w.WA + WA_BYTES; w.WC = w.WB + WB_BYTES; w.WD = w.WC + WC_BYTES;
  w.vec = (float*)(w.WD + WD_BYTES);
  return w;
}

// ---- weight images --------------------------------------------------------------------------
__global__ void k_edge_prep(int H, int K, int A, const float* __restrict__ W1, const float* __restrict__ W2,
                            const float* __restrict__ b2, const float* __restrict__ Ws, const float* __restrict__ bs,
                            const float* __restrict__ mu, const float* __restrict__ beta, EdgeW w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  // units: WA 2*64*8, WB 2*80*8, WC 2*64*8, WD 2*64*8
  const int nA = 2 * 64 * 8, nB = 2 * 80 * 8, nC = nA, nD = nA;
  if (t < nA + nB + nC + nD) {
    int which, r = t, rows;
    if (r < nA) { which = 0; rows = 64; }
    else if ((r -= nA) < nB) { which = 1; rows = 80; }
    else if ((r -= nB) < nC) { which = 2; rows = 64; }
    else { r -= nC; which = 3; rows = 64; }
    const int chunk = r / (rows * 8);
    r %= rows * 8;
    const int row = r / 8, u = r % 8;
    float vals[4];
    for (int i = 0; i < 4; ++i) {
      const int kk = chunk * 32 + u * 4 + i;
      float v = 0.f;
      if (which == 0) { if (kk <= K) v = W1[(size_t)(2 * H + kk) * H + row]; }
      else if (which == 1) {
        if (row < 64) v = W2[(size_t)kk * H + row];
        else if (row < 64 + A) {
          float s = 0.f;
          for (int f = 0; f < H; ++f) s = fmaf(W2[(size_t)kk * H + f], Ws[(size_t)f * A + (row - 64)], s);
          v = s;
        }
      }
      else if (which == 2) v = W2[(size_t)row * H + kk];
      else { if (row <= K) v = W1[(size_t)(2 * H + row) * H + kk]; }
      vals[i] = v;
    }
    uint8_t* base = (which == 0 ? w.WA : which == 1 ? w.WB : which == 2 ? w.WC : w.WD) + (size_t)chunk * 2 * rows * 128;
    const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)u);
    float4 hi, lo;
    split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
    split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
    *reinterpret_cast<float4*>(base + off) = hi;
    *reinterpret_cast<float4*>(base + (size_t)rows * 128 + off) = lo;
  }
  if (t < EVEC) {
    float v = 0.f;
    if (t < 64) v = t < K ? mu[t] : 0.f;
    else if (t < 128) v = (t - 64) < K ? beta[t - 64] : 0.f;
    else if (t < 192) v = b2[t - 128];
    else if (t < 192 + A) {
      const int a = t - 192;
      float s = bs[a];
      for (int f = 0; f < H; ++f) s = fmaf(b2[f], Ws[(size_t)f * A + a], s);
      v = s;
    } else if (t >= 256) {
      const int f = (t - 256) / 4, a = (t - 256) % 4;
      if (f < H && a < A) v = Ws[(size_t)f * A + a];
    }
    w.vec[t] = v;
  }
}

struct EdgeArgs {
  TileGeom g;
  int K, Kp, NP;
  const float *x, *mask, *proj;
  EdgeW w;
  float *e_out, *logit_out;            // forward outputs  [P,64], [P,4]
  float* ge;                           // backward input   [P,64] cotangent of e through x_mixing / aggregate; the
                                       // attention-logit term W_s g_q is added here (written back when training)
  const float *gdir, *gq;              // backward inputs  [P,3], [P,4] (cotangent of the pre-celu logits)
  const float* gcut;                   // backward input   [P] cotangent of the distance through the cutoff, or NULL
  const float *pair_u, *pair_p;        // `he` edge features as per-pair additive terms of u [P,Kp] and z1 [P,64], or NULL
  float *PB, *a1buf, *gbuf;            // backward outputs [P,192], [P,64], [P,64] (a1buf/gbuf: training only)
  int train;
};

__device__ __forceinline__ void store_unit_tf32(uint8_t* chunk_img, int row, int u, const float* vals) {
  const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)u);
  float4 hi, lo;
  split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
  split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
  *reinterpret_cast<float4*>(chunk_img + off) = hi;
  *reinterpret_cast<float4*>(chunk_img + EP_IMG + off) = lo;
}

// one K-chunk (32 of the 64 K values) of a 3xTF32 GEMM: D[128 x N] (+)= A[128 x 32] * B[N x 32]^T
//   a_img: {hi, lo} images of the chunk; b_img: weight image [2 chunks][{hi, lo}][b_rows x 128 B]
__device__ __forceinline__ void edge_gemm_chunk(uint32_t d_tmem, uint32_t a_img, uint32_t b_img, int chunk, int b_rows,
                                                uint32_t idesc) {
  const int pp[3] = {0, 1, 0}, pw[3] = {0, 0, 1};
#pragma unroll
  for (int pr = 0; pr < 3; ++pr)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a0 = a_img + pp[pr] * EP_IMG + ks * 32;
      const uint32_t b0 = b_img + (chunk * 2 + pw[pr]) * (b_rows * 128) + ks * 32;
      umma<true>(d_tmem, umma_desc_k_sw128(a0), umma_desc_k_sw128(b0), idesc, (chunk | pr | ks) != 0);
    }
}

// EDGE_GROUPS independent 128-thread groups per CTA (one tile each, round robin): the kernel is bound by the
// latency of its per-pair chains (projection loads -> RBF / silu -> operand image -> MMA -> TMEM), so the
// lever is the number of tiles in flight per SM (3 groups of 168 registers measured best; 4 x 128 spills).  A group owns ONE 32 KB chunk image
// ({hi, lo} of 32 K values) and 128 TMEM columns: every GEMM is issued as two K-chunks through the same
// image, and accumulators are recycled (forward: E' overwrites Z1; backward: GG overwrites Z1) — the half
// of Z1 that is still needed is pulled into registers before the overwriting MMA is issued.
template <bool BWD>
__global__ void __launch_bounds__(EDGE_THREADS, 1) k_tc_edge(EdgeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  if (base != smem_raw) __trap();                          // the budget has no alignment slack
  constexpr int W_BYTES = BWD ? (WA_BYTES + WC_BYTES + WD_BYTES) : (WA_BYTES + WB_BYTES);
  uint8_t* sWA = base;
  uint8_t* sW2 = base + WA_BYTES;                          // fwd: WB ; bwd: WC
  uint8_t* sWD = base + WA_BYTES + WC_BYTES;               // bwd only
  uint8_t* imgs = base + W_BYTES;                          // [EDGE_GROUPS][EG_IMG]
  float* svec = reinterpret_cast<float*>(imgs + EDGE_GROUPS * EG_IMG);
  uint64_t* wbar = reinterpret_cast<uint64_t*>(svec + EVEC);
  uint64_t* mbar = wbar + 1;                               // [EDGE_GROUPS]
  uint32_t* tptr = reinterpret_cast<uint32_t*>(mbar + EDGE_GROUPS);
  const int warp = threadIdx.x >> 5;
  const int grp = threadIdx.x >> 7, pl = threadIdx.x & 127;
  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    for (int g = 0; g < EDGE_GROUPS; ++g) mbar_init(mbar + g, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(wbar, W_BYTES);
    bulk_g2s(sWA, a.w.WA, WA_BYTES, wbar);
    if (BWD) { bulk_g2s(sW2, a.w.WC, WC_BYTES, wbar); bulk_g2s(sWD, a.w.WD, WD_BYTES, wbar); }
    else bulk_g2s(sW2, a.w.WB, WB_BYTES, wbar);
  }
  if (warp == 0) tmem_alloc<512>(tptr);
  for (int t = threadIdx.x; t < EVEC; t += EDGE_THREADS) svec[t] = a.w.vec[t];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  mbar_wait_warp(wbar, 0);
  const uint32_t tmem_base = *tptr;
  const uint32_t tcol = tmem_base + grp * 128;                         // this group's 128 TMEM columns
  const uint32_t lane_addr = tcol + ((uint32_t)((warp & 3) * 32) << 16);
  uint8_t* img = imgs + grp * EG_IMG;
  const uint32_t img_u32 = smem_u32(img);
  const float* s_mu = svec;
  const float* s_beta = svec + 64;
  const float* s_b2 = svec + 128;
  const float* s_bq = svec + 192;
  const int K = a.K, Kp = a.Kp, NP = a.NP;
  const int ntl = (geom_tiles(a.g) - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  uint32_t ph = 0;
  constexpr uint32_t idesc64 = umma_idesc(2, 128, 64);
  constexpr uint32_t idesc80 = umma_idesc(2, 128, 80);
  const int bar_id = 1 + grp;
  // publish the chunk image, run one K-chunk of a GEMM on it, wait until the MMAs (and their reads of the
  // image) are complete
  auto run_chunk = [&](uint32_t dcol, const uint8_t* wimg, int chunk, int b_rows, uint32_t idesc) {
    fence_proxy_async();
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    if (pl == 0) {
      tc_fence_after();
      edge_gemm_chunk(dcol, img_u32, smem_u32(wimg), chunk, b_rows, idesc);
      umma_commit(mbar + grp);
    }
    mbar_wait_warp(mbar + grp, ph);
    ph ^= 1;
    tc_fence_after();
  };

  for (int it = grp; it < ntl; it += EDGE_GROUPS) {
    const int tile = blockIdx.x + it * gridDim.x;
    bool valid;
    int row, j;
    long long prx;
    tile_pair(tile_desc(a.g, tile), pl, valid, row, j, prx);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, nrm = 0.f, tt = 0.f, m = 1.0f;
    // per-node projections (G8 layout, NP / 4 units per atom): sender j varies with the lane -> full lines,
    // receiver i is the same for all lanes of a row segment -> broadcast
    const float4* nj4 = reinterpret_cast<const float4*>(a.proj);
    const float4* ni4 = nj4;
    bool diag = false;
    if (!valid) prx = 0;
    if (valid) {
      const int mol0 = geom_mol0(a.g, row);
      diag = (row - mol0) == j;
      nj4 += g8_row(mol0 + j, NP >> 2);
      ni4 += g8_row(row, NP >> 2);
      const float* xi = a.x + (size_t)row * 3;
      const float* xj = a.x + (size_t)(mol0 + j) * 3;
      r0 = xj[0] - xi[0]; r1 = xj[1] - xi[1]; r2 = xj[2] - xi[2];
      nrm = sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 0.f) + 1e-5f);     // functional.py:14-17
      tt = fexp_(-nrm);                                                    // utils.py:62-64 (alpha = 1, lower = 0)
      if (a.mask) m = a.mask[prx];
    }
    // ---------------- (a) G = [rho*u | n | 1 | t | 0...]  ->  A operand of GEMM A, one K-chunk at a time
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      // issue the 16 projection loads of this chunk before any of the exp chains (latency overlap)
      float4 uj4[8], ui4[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int u = hb * 8 + q;
        uj4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        ui4[q] = uj4[q];
        if (4 * u < Kp) {                      // idle lanes read node 0 (valid memory); their rows are never stored
          uj4[q] = __ldg(&nj4[((4 * u) >> 2) * G8S]);
          ui4[q] = __ldg(&ni4[((Kp + 4 * u) >> 2) * G8S]);
          if (a.pair_u) {                      // edge features: + he @ W_in[2F:2F+E]  (SakePairTerms)
            const float4 ue = __ldg(reinterpret_cast<const float4*>(a.pair_u + prx * Kp + 4 * u));
            ui4[q].x += ue.x; ui4[q].y += ue.y; ui4[q].z += ue.z; ui4[q].w += ue.w;
          }
        }
      }
      // 32 independent exp chains, no control flow in between: mu / beta are zero-padded beyond K (rho = 1) and
      // so are the projections (u = 0), hence rho*u = 0 there; the three extra columns n, 1, t are patched in
      // afterwards under ONE warp-uniform test per chunk (per-element or per-unit branches serialise the chains)
      float vals[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int u = hb * 8 + q;
        const float us[4] = {uj4[q].x + ui4[q].x, uj4[q].y + ui4[q].y, uj4[q].z + ui4[q].z, uj4[q].w + ui4[q].w};
        const float4 mu4 = *reinterpret_cast<const float4*>(s_mu + 4 * u);
        const float4 be4 = *reinterpret_cast<const float4*>(s_beta + 4 * u);
        const float mus[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, bes[4] = {be4.x, be4.y, be4.z, be4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dm = tt - mus[i];
          vals[4 * q + i] = fexp_(-bes[i] * dm * dm) * us[i];
        }
      }
      if (hb * 32 + 31 >= K && hb * 32 <= K + 2) {
#pragma unroll
        for (int idx = 0; idx < 32; ++idx) {
          const int k = hb * 32 + idx;
          vals[idx] = k == K ? nrm : (k == K + 1 ? 1.0f : (k == K + 2 ? tt : vals[idx]));   // 1: column sums for free in the dW contraction
        }
      }
      if (BWD && a.train && valid) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(a.gbuf)[g8_row(prx, 16) + (hb * 8 + q) * G8S] = make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) store_unit_tf32(img, pl, q, vals + 4 * q);
      run_chunk(tcol + 0, sWA, hb, 64, idesc64);                           // Z1 -> cols [0,64)
    }

    if (!BWD) {
      // ---------------- (c) a1 = silu(Z1 + pj[j] + pi[i])  (layers.py:33-38, 23) -> A operand of GEMM B.
      // GEMM B writes E' over Z1, so the second half of Z1 is read before its first chunk is issued.
      float z1b[32];
      {
        float4 pj4[8], pi4[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          pj4[u] = __ldg(&nj4[((2 * Kp + 4 * u) >> 2) * G8S]);
          pi4[u] = __ldg(&ni4[((2 * Kp + 64 + 4 * u) >> 2) * G8S]);
          if (a.pair_p) {                      // edge features: + he @ W_1[2F:2F+E]
            const float4 pe = __ldg(reinterpret_cast<const float4*>(a.pair_p + prx * 64 + 4 * u));
            pi4[u].x += pe.x; pi4[u].y += pe.y; pi4[u].z += pe.z; pi4[u].w += pe.w;
          }
        }
        float v[32];
        tmem_ld32(lane_addr, v);
        tmem_ld32(lane_addr + 32, z1b);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float vals[4] = {fsilu_(v[4 * u] + pj4[u].x + pi4[u].x), fsilu_(v[4 * u + 1] + pj4[u].y + pi4[u].y),
                                 fsilu_(v[4 * u + 2] + pj4[u].z + pi4[u].z), fsilu_(v[4 * u + 3] + pj4[u].w + pi4[u].w)};
          store_unit_tf32(img, pl, u, vals);
        }
      }
      // second-half projections: requested before the hand-off so that their latency hides behind the MMA
      float4 pj4[8], pi4[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        pj4[u] = __ldg(&nj4[((2 * Kp + 32 + 4 * u) >> 2) * G8S]);
        pi4[u] = __ldg(&ni4[((2 * Kp + 64 + 32 + 4 * u) >> 2) * G8S]);
        if (a.pair_p) {
          const float4 pe = __ldg(reinterpret_cast<const float4*>(a.pair_p + prx * 64 + 32 + 4 * u));
          pi4[u].x += pe.x; pi4[u].y += pe.y; pi4[u].z += pe.z; pi4[u].w += pe.w;
        }
      }
      run_chunk(tcol + 0, sW2, 0, 80, idesc80);                            // E' (chunk 0) -> cols [0,80)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float vals[4] = {fsilu_(z1b[4 * u] + pj4[u].x + pi4[u].x), fsilu_(z1b[4 * u + 1] + pj4[u].y + pi4[u].y),
                               fsilu_(z1b[4 * u + 2] + pj4[u].z + pi4[u].z), fsilu_(z1b[4 * u + 3] + pj4[u].w + pi4[u].w)};
        store_unit_tf32(img, pl, u, vals);
      }
      run_chunk(tcol + 0, sW2, 1, 80, idesc80);                            // E' (chunk 1)
      // ---------------- (e) e = E' + b2 ; logits = celu(q) - 1e5*diag - 1e5*(1-m)  (layers.py:24,155-165)
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(lane_addr + half * 32, v);
        tmem_ld_wait();
        if (valid) {
          float4* o = reinterpret_cast<float4*>(a.e_out) + g8_row(prx, 16) + half * 8 * G8S;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int f0 = half * 32 + 4 * u;
            o[u * G8S] = make_float4(v[4 * u] + s_b2[f0], v[4 * u + 1] + s_b2[f0 + 1], v[4 * u + 2] + s_b2[f0 + 2],
                               v[4 * u + 3] + s_b2[f0 + 3]);
          }
        }
      }
      {
        float v[16];
        tmem_ld16(lane_addr + 64, v);
        tmem_ld_wait();
        if (valid) {
          float s[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            s[q] = celu2f_(v[q] + s_bq[q]);
            if (diag) s[q] -= 1e5f;
            if (a.mask) s[q] -= 1e5f * (1.0f - m);
          }
          *reinterpret_cast<float4*>(a.logit_out + prx * 4) = make_float4(s[0], s[1], s[2], s[3]);
        }
      }
      tc_fence_before();
    } else {
      // ---------------- (c') a1 for the dW2 contraction (training), then GE -> A operand of GEMM C
      if (a.train) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tmem_ld32(lane_addr + half * 32, v);
          tmem_ld_wait();
          if (valid) {
            float4* o = reinterpret_cast<float4*>(a.a1buf) + g8_row(prx, 16) + half * 8 * G8S;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int f0 = half * 32 + 4 * u;
              const float4 pj = __ldg(&nj4[((2 * Kp + f0) >> 2) * G8S]);
              float4 pi = __ldg(&ni4[((2 * Kp + 64 + f0) >> 2) * G8S]);
              if (a.pair_p) {
                const float4 pe = __ldg(reinterpret_cast<const float4*>(a.pair_p + prx * 64 + f0));
                pi.x += pe.x; pi.y += pe.y; pi.z += pe.z; pi.w += pe.w;
              }
              o[u * G8S] = make_float4(fsilu_(v[4 * u] + pj.x + pi.x), fsilu_(v[4 * u + 1] + pj.y + pi.y),
                                 fsilu_(v[4 * u + 2] + pj.z + pi.z), fsilu_(v[4 * u + 3] + pj.w + pi.w));
            }
          }
        }
      }
      {
        // g_e = (cotangent through x_mixing / aggregate) + W_s g_q   (layers.py:155: logits = e W_s + b_s)
        float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
        gq = __ldg(reinterpret_cast<const float4*>(a.gq + prx * 4));
        const float4* s_ws = reinterpret_cast<const float4*>(svec + 256);
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
          float4 g4[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            g4[u] = reinterpret_cast<const float4*>(a.ge)[g8_row(prx, 16) + (hb * 8 + u) * G8S];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float vals[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 ws = s_ws[hb * 32 + 4 * u + i];
              vals[i] += ws.x * gq.x + ws.y * gq.y + ws.z * gq.z + ws.w * gq.w;
            }
            if (a.train && valid) reinterpret_cast<float4*>(a.ge)[g8_row(prx, 16) + (hb * 8 + u) * G8S] = make_float4(vals[0], vals[1], vals[2], vals[3]);
            store_unit_tf32(img, pl, u, vals);
          }
          run_chunk(tcol + 64, sW2, hb, 64, idesc64);                      // GA1 = GE W2^T -> cols [64,128)
        }
      }
      float4* const pb4 = reinterpret_cast<float4*>(a.PB) + g8_row(prx, PB_LD / 4);       // this pair's record (G8 layout)
      // ---------------- (e') g_z1 = GA1 * silu'(z1)  -> A operand of GEMM D, and the per-pair record.
      // GEMM D writes GG over Z1, so the second half of Z1 is read before its first chunk is issued.
      float z1b[32];
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float4 pj4[8], pi4[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          pj4[u] = __ldg(&nj4[((2 * Kp + half * 32 + 4 * u) >> 2) * G8S]);
          pi4[u] = __ldg(&ni4[((2 * Kp + 64 + half * 32 + 4 * u) >> 2) * G8S]);
          if (a.pair_p) {
            const float4 pe = __ldg(reinterpret_cast<const float4*>(a.pair_p + prx * 64 + half * 32 + 4 * u));
            pi4[u].x += pe.x; pi4[u].y += pe.y; pi4[u].z += pe.z; pi4[u].w += pe.w;
          }
        }
        float z[32], ga[32];
        if (half == 0) {
          tmem_ld32(lane_addr, z);
          tmem_ld32(lane_addr + 32, z1b);
        }
        tmem_ld32(lane_addr + 64 + half * 32, ga);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float vals[4];
          {
            const int f0 = half * 32 + 4 * u;
            const float4 pj = pj4[u];
            const float4 pi = pi4[u];
            const float zz0 = half == 0 ? z[4 * u] : z1b[4 * u], zz1 = half == 0 ? z[4 * u + 1] : z1b[4 * u + 1],
                        zz2 = half == 0 ? z[4 * u + 2] : z1b[4 * u + 2], zz3 = half == 0 ? z[4 * u + 3] : z1b[4 * u + 3];
            vals[0] = ga[4 * u] * fdsilu_(zz0 + pj.x + pi.x);
            vals[1] = ga[4 * u + 1] * fdsilu_(zz1 + pj.y + pi.y);
            vals[2] = ga[4 * u + 2] * fdsilu_(zz2 + pj.z + pi.z);
            vals[3] = ga[4 * u + 3] * fdsilu_(zz3 + pj.w + pi.w);
            if (valid) pb4[((f0) >> 2) * G8S] = make_float4(vals[0], vals[1], vals[2], vals[3]);
          }
          store_unit_tf32(img, pl, u, vals);
        }
        run_chunk(tcol + 0, sWD, half, 64, idesc64);                       // GG = GZ1 W1[2H:]^T -> cols [0,64)
      }
      // ---------------- (g) RBF / geometry backward (utils.py:61-65, functional.py:7-19, layers.py:115)
      float gt = 0.f, gn = 0.f;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float4 uj4[8], ui4[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uj4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          ui4[u] = uj4[u];
          if (half * 32 + 4 * u < Kp) {               // warp-uniform; idle lanes read node 0
            uj4[u] = __ldg(&nj4[((half * 32 + 4 * u) >> 2) * G8S]);
            ui4[u] = __ldg(&ni4[((Kp + half * 32 + 4 * u) >> 2) * G8S]);
            if (a.pair_u) {
              const float4 ue = __ldg(reinterpret_cast<const float4*>(a.pair_u + prx * Kp + half * 32 + 4 * u));
              ui4[u].x += ue.x; ui4[u].y += ue.y; ui4[u].z += ue.z; ui4[u].w += ue.w;
            }
          }
        }
        float gg[32];
        tmem_ld32(lane_addr + half * 32, gg);
        tmem_ld_wait();
        // 32 straight-line chains (zero-padded mu / beta / projections beyond K: w = 0, beta = 0); the columns
        // >= K of gu are never read (k_pair_reduce writes zeros there); stores and the distance column afterwards
        float gu[32], wv[32];
        float gta = 0.f, gtb = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k0 = half * 32 + 4 * u;
          const float us[4] = {uj4[u].x + ui4[u].x, uj4[u].y + ui4[u].y, uj4[u].z + ui4[u].z, uj4[u].w + ui4[u].w};
          const float4 mu4 = *reinterpret_cast<const float4*>(s_mu + k0);
          const float4 be4 = *reinterpret_cast<const float4*>(s_beta + k0);
          const float mus[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, bes[4] = {be4.x, be4.y, be4.z, be4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float dm = tt - mus[i];
            const float rho = fexp_(-bes[i] * dm * dm);
            gu[4 * u + i] = gg[4 * u + i] * rho;                         // d/du
            const float w = dm * rho * gg[4 * u + i] * us[i];            // dm * rho * d/drho
            wv[4 * u + i] = w;
            if (u & 1) gtb = fmaf(-2.0f * bes[i], w, gtb); else gta = fmaf(-2.0f * bes[i], w, gta);
          }
        }
        gt += gta + gtb;
        if (K >= half * 32 && K < half * 32 + 32) {
#pragma unroll
          for (int idx = 0; idx < 32; ++idx)
            if (half * 32 + idx == K) gn = gg[idx];                      // through the distance input of mlp_out[0]
        }
        if (valid) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k0 = half * 32 + 4 * u;
            if (k0 < 60) pb4[((64 + k0) >> 2) * G8S] = make_float4(gu[4 * u], gu[4 * u + 1], gu[4 * u + 2], gu[4 * u + 3]);
          }
          if (a.train) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              pb4[((128 + half * 32 + 4 * u) >> 2) * G8S] = make_float4(wv[4 * u], wv[4 * u + 1], wv[4 * u + 2], wv[4 * u + 3]);
          }
        }
      }
      if (valid) {
        gn = fmaf(-tt, gt, gn);                                  // t = exp(-n)
        if (a.gcut) gn += a.gcut[prx];                           // euclidean attention eps(n) (layers.py:172-176)
        const float* gd = a.gdir + prx * 3;
        const float inv = 1.0f / (nrm + 1e-5f);
        float g0 = gd[0] * inv, g1 = gd[1] * inv, g2 = gd[2] * inv;
        gn -= (gd[0] * r0 + gd[1] * r1 + gd[2] * r2) * inv * inv;
        const float n2 = r0 * r0 + r1 * r1 + r2 * r2;
        const float gn2 = n2 > 0.f ? gn / (2.0f * nrm) : 0.f;   // relu'(0) = 0 (functional.py:15)
        g0 = fmaf(2.0f * r0, gn2, g0); g1 = fmaf(2.0f * r1, gn2, g1); g2 = fmaf(2.0f * r2, gn2, g2);
        if (diag) { g0 = 0.f; g1 = 0.f; g2 = 0.f; }
        pb4[((124) >> 2) * G8S] = make_float4(g0, g1, g2, 0.f);
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ---- attention backward (tcgen05 edge path) --------------------------------------------------------
// One warp per receiving atom: g_s = att * (g_att - sum_j g_att att) (the renormalisation of layers.py:180
// is the identity on the gradient), g_q = g_s * celu_2'(q) with celu_2'(q) = 1 (q > 0) or e^{q/2} =
// celu_2(q)/2 + 1 read off the saved logits (the -1e5 offsets of self / masked pairs meet att = 0).
// Lane l owns the elements t = l, l+32, ... of the [N,4] row, i.e. always head a = l & 3.
__global__ void __launch_bounds__(256) k_attn_bwd_tc(Dims d, const float* __restrict__ x, const float* __restrict__ att,
                                                     const float* __restrict__ logit, float* __restrict__ gatt,
                                                     float* __restrict__ gcut) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= dims_rows(d)) return;
  const RowInfo ri = row_info(d, row);
  const size_t base = (size_t)ri.pair0 * 4;
  const int n4 = ri.n * 4;
  constexpr int MAXI = 8;                                  // register-resident up to N = 64; longer rows re-read
  float av[MAXI], gv[MAXI];
  float s = 0.f;
  for (int i = 0; i * 32 + lane < n4; ++i) {
    const int t = i * 32 + lane;
    const float a_ = att[base + t], g_ = gatt[base + t];
    if (i < MAXI) { av[i] = a_; gv[i] = g_; }
    s = fmaf(g_, a_, s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  if (gcut != nullptr) {
    // cosine cutoff (layers.py:172-180): att = eps sigma m / sum_k(eps sigma m).  The cotangent of the softmax
    // logits is still g_s = att (g_att - s); the one of eps is g_s / eps per head (att is proportional to eps),
    // summed over the heads that share the pair's eps, times d eps / d d.  eps = 0 exactly: att = 0 and the
    // term is dropped (a set of measure zero: d = lower + k (upper - lower)).
    for (int t0 = 0; t0 < n4; t0 += 32) {                 // warp-uniform trip count: the shuffles need every lane
      const int t = t0 + lane;
      const bool on = t < n4;
      float gs = on ? att[base + t] * (gatt[base + t] - s) : 0.f;
      gs += __shfl_xor_sync(0xffffffffu, gs, 1);          // heads of one pair sit in 4 adjacent lanes (n4 % 4 == 0)
      gs += __shfl_xor_sync(0xffffffffu, gs, 2);
      if (on && (t & 3) == 0) {
        const int j = t >> 2;
        gcut[(size_t)ri.pair0 + j] = gs * cosine_cutoff_dlog_(pair_dist_(x, row, ri.mol0 + j), d.cut_lo, d.cut_hi);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int t = i * 32 + lane;
    if (t < n4) {
      const float lgv = logit[base + t];
      gatt[base + t] = av[i] * (gv[i] - s) * (lgv > 0.f ? 1.0f : fmaf(0.5f, lgv, 1.0f));
    }
  }
  for (int i = MAXI; i * 32 + lane < n4; ++i) {
    const int t = i * 32 + lane;
    const float lgv = logit[base + t];
    gatt[base + t] = att[base + t] * (gatt[base + t] - s) * (lgv > 0.f ? 1.0f : fmaf(0.5f, lgv, 1.0f));
  }
}

int tc_attn_bwd(const Dims& d, const float* x, const Saved& sv, const BwdScratch& sc, cudaStream_t st) {
  ProfScope prof(13, d.P, st);
  k_attn_bwd_tc<<<(d.R + 7) / 8, 256, 0, st>>>(d, x, sv.att, sv.logit, sc.gatt, d.cutoff ? sc.gcut : nullptr);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---- reductions of the per-pair record over senders / receivers -----------------------------------
// gproj[n] = [ sum_i gu[(i,n)] | sum_j gu[(n,j)] | sum_i gz1[(i,n)] | sum_j gz1[(n,j)] ],  dx[n] += sum_i g_r[(i,n)] - sum_j g_r[(n,j)]
// The record is in the G8 layout (8 pairs x one 16-byte unit = one line).  One CTA = 8 consecutive atoms, thread =
// (atom a' = t & 7, unit u = t >> 3 of the first 128 record columns):
//   column sums  sum_i rec[(i, n0 + a')]: the 8 threads of a unit read 8 consecutive pair slots of row i -> full lines
//   row sums     sum_j rec[(n, j)]: for each of the 8 atoms in turn, the 8 threads of a unit read 8 consecutive j and
//                the partial sums meet in three shuffles; thread a' keeps the sum of atom a'.
// Every thread then owns both sums of (its atom, its unit) and writes the G8 rows of gproj itself.
// Training with update = True also finishes the v_mixing gradient here (layers.py:94,220-223):
//   gWv[c] += sum_n sum_d ssum[n][c][d] * qv[n][d],  qv = g_dv / den2 left by k_tc_node_post_bwd; thread = coefficient c.
__global__ void __launch_bounds__(256) k_pair_reduce(Dims d, const float* __restrict__ PB,
                                                     float* __restrict__ gproj, float* __restrict__ dx,
                                                     const float* __restrict__ ssum, const float* __restrict__ qv,
                                                     float* __restrict__ gWv) {
  const int n0 = blockIdx.x * 8, R = dims_rows(d);
  if (n0 >= R) return;
  if (gWv != nullptr) {
    const int c = threadIdx.x;
    const float* sp = ssum + (g8_row(n0, 192) + (size_t)(c >> 2) * 3 * G8S) * 4 + (c & 3);   // G8: the 8 atoms of this CTA are one group
    float acc = 0.f;
    for (int a2 = 0; a2 < 8 && n0 + a2 < R; ++a2) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(qv) + n0 + a2);
      acc += sp[4 * a2] * q.x + sp[4 * G8S + 4 * a2] * q.y + sp[8 * G8S + 4 * a2] * q.z;
    }
    atomicAdd(gWv + c, acc);
  }
  const int K = d.K, Kp = d.Kp, NP = d.NP;
  const int ap = threadIdx.x & 7, u = threadIdx.x >> 3;            // u = 0 .. 31
  const float4* pb = reinterpret_cast<const float4*>(PB) + u * G8S;
  constexpr int U = PB_LD / 4;
  float4 sj = make_float4(0.f, 0.f, 0.f, 0.f), si = sj;
  const int n = n0 + ap;
  if (n < R) {
    const RowInfo ri = row_info(d, n);
    const long long c0 = ri.pair0 - (long long)(n - ri.mol0) * ri.n + (n - ri.mol0);   // pair (0, n) of the molecule
    // four loads in flight per thread (the sums keep the order i = 0, 1, 2, ...: one accumulator)
    int i = 0;
    for (; i + 4 <= ri.n; i += 4) {
      const float4 v0 = __ldg(pb + g8_row(c0 + (long long)i * ri.n, U));
      const float4 v1 = __ldg(pb + g8_row(c0 + (long long)(i + 1) * ri.n, U));
      const float4 v2 = __ldg(pb + g8_row(c0 + (long long)(i + 2) * ri.n, U));
      const float4 v3 = __ldg(pb + g8_row(c0 + (long long)(i + 3) * ri.n, U));
      sj.x += v0.x; sj.y += v0.y; sj.z += v0.z; sj.w += v0.w;
      sj.x += v1.x; sj.y += v1.y; sj.z += v1.z; sj.w += v1.w;
      sj.x += v2.x; sj.y += v2.y; sj.z += v2.z; sj.w += v2.w;
      sj.x += v3.x; sj.y += v3.y; sj.z += v3.z; sj.w += v3.w;
    }
    for (; i < ri.n; ++i) {
      const float4 v = __ldg(pb + g8_row(c0 + (long long)i * ri.n, U));
      sj.x += v.x; sj.y += v.y; sj.z += v.z; sj.w += v.w;
    }
  }
  // row sums: the loads of all eight atoms are issued before the first shuffle (they are independent)
  float4 racc[8];
#pragma unroll
  for (int a2 = 0; a2 < 8; ++a2) {
    racc[a2] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int m = n0 + a2;
    if (m < R) {                                                   // block-uniform
      const RowInfo ri = row_info(d, m);
      for (int j = ap; j < ri.n; j += 8) {
        const float4 v = __ldg(pb + g8_row(ri.pair0 + j, U));
        racc[a2].x += v.x; racc[a2].y += v.y; racc[a2].z += v.z; racc[a2].w += v.w;
      }
    }
  }
#pragma unroll
  for (int a2 = 0; a2 < 8; ++a2) {
    float4 acc = racc[a2];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (ap == a2) si = acc;
  }
  if (n >= R) return;
  float4* gp = reinterpret_cast<float4*>(gproj) + g8_row(n, NP >> 2);
  if (u < 16) {                                                    // gz1: record columns [0, 64)
    gp[((2 * Kp) / 4 + u) * G8S] = sj;
    gp[((2 * Kp) / 4 + 16 + u) * G8S] = si;
  } else if (u < 16 + Kp / 4) {                                    // gu: record columns [64, 64 + Kp), zero beyond K
    const int k0 = 4 * (u - 16);
    float4 a = sj, b = si;
    if (k0 + 0 >= K) { a.x = 0.f; b.x = 0.f; }
    if (k0 + 1 >= K) { a.y = 0.f; b.y = 0.f; }
    if (k0 + 2 >= K) { a.z = 0.f; b.z = 0.f; }
    if (k0 + 3 >= K) { a.w = 0.f; b.w = 0.f; }
    gp[(u - 16) * G8S] = a;
    gp[(Kp / 4 + u - 16) * G8S] = b;
  } else if (u == 31) {                                            // g_r: record columns 124 .. 126
    dx[(size_t)n * 3] += sj.x - si.x; dx[(size_t)n * 3 + 1] += sj.y - si.y; dx[(size_t)n * 3 + 2] += sj.z - si.z;
  }
}

// cotangents of SakePairTerms out of the per-pair record (columns >= K of g_u are zero)
__global__ void k_pair_terms_out(long long P, int K, int Kp, const float* __restrict__ PB, float* __restrict__ g_u,
                                 float* __restrict__ g_p) {
  const int w = 64 + Kp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * w) return;
  const long long pr = t / w;
  const int c = (int)(t - pr * w);
  const float v = PB[g8_elem(pr, PB_LD / 4, c)];
  if (c < 64) { if (g_p) g_p[pr * 64 + c] = v; }
  else if (g_u) { const int k = c - 64; g_u[pr * Kp + k] = k < K ? v : 0.f; }
}

// ---- host ------------------------------------------------------------------------------------------
bool tc_edge_supported(const Dims& d) { return d.H == 64 && d.A == 4 && d.K <= 58; }

static int edge_prep(const Dims& d, const SakeLayerParams& p, const EdgeW& w, cudaStream_t st) {
  const int units = 2 * 64 * 8 * 3 + 2 * 80 * 8;
  k_edge_prep<<<(units + 255) / 256, 256, 0, st>>>(d.H, d.K, d.A, p.mlp_out0_kernel, p.mlp_out2_kernel, p.mlp_out2_bias,
                                                   p.sem_kernel, p.sem_bias, p.rbf_means, p.rbf_betas, w);
  note_launches(1);
  return 0;
}

static int edge_grid(const TileGeom& g) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (g.num_tiles + EDGE_GROUPS - 1) / EDGE_GROUPS;
  return want < sms ? (want < 1 ? 1 : want) : sms;
}

int tc_edge_prepare(const Dims& d, const SakeLayerParams& p, void* wedge, cudaStream_t st) {
  edge_prep(d, p, carve_edge_w(wedge), st);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tc_edge_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                void* wscratch, cudaStream_t st) {
  EdgeW w = carve_edge_w(wscratch);
  if (!d.prepared) edge_prep(d, p, w, st);
  EdgeArgs a;
  memset(&a, 0, sizeof(a));
  a.g = make_geom(d);
  a.K = d.K; a.Kp = d.Kp; a.NP = d.NP;
  a.x = x; a.mask = mask; a.proj = sv.nodeproj; a.w = w;
  a.pair_u = d.pair_u; a.pair_p = d.pair_p;
  a.e_out = sv.e; a.logit_out = sv.logit;
  const size_t smem = WA_BYTES + WB_BYTES + EDGE_GROUPS * EG_IMG + EVEC * 4 + 64;
  static unsigned long long optin = 0;
  { const int rc = smem_optin(k_tc_edge<false>, smem, optin); if (rc) return rc; }
  {
    ProfScope prof(4, d.P, st);
    k_tc_edge<false><<<edge_grid(a.g), EDGE_THREADS, smem, st>>>(a);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// PB [P,192] | a1buf [P,64] | gbuf [P,64] | extra [2,192]   (a1buf / gbuf only when training)
size_t tc_edge_bwd_scratch_bytes(const Dims& d, int with_grads) {
  size_t n = align_up(sizeof(float) * rows_pad8(d.P) * PB_LD) + align_up(sizeof(float) * 2 * PB_LD);
  if (with_grads) n += 2 * align_up(sizeof(float) * rows_pad8(d.P) * 64);
  return n;
}

int tc_edge_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                const BwdScratch& sc, float* dx, const SakeLayerGrads* g, void* wscratch, void* escratch, XtgList& L,
                float* g_pair_u, float* g_pair_p, cudaStream_t st) {
  EdgeW w = carve_edge_w(wscratch);                      // built by tc_edge_fwd of the same step (saved.wedge)
  char* eb = (char*)escratch;
  float* PB = (float*)eb; eb += align_up(sizeof(float) * rows_pad8(d.P) * PB_LD);
  float* extra = (float*)eb; eb += align_up(sizeof(float) * 2 * PB_LD);
  float* a1buf = nullptr; float* gbuf = nullptr;
  if (g) { a1buf = (float*)eb; eb += align_up(sizeof(float) * rows_pad8(d.P) * 64); gbuf = (float*)eb; }
  EdgeArgs a;
  memset(&a, 0, sizeof(a));
  a.g = make_geom(d);
  a.K = d.K; a.Kp = d.Kp; a.NP = d.NP;
  a.x = x; a.mask = mask; a.proj = sv.nodeproj; a.w = w;
  a.gcut = d.cutoff ? sc.gcut : nullptr;
  a.pair_u = d.pair_u; a.pair_p = d.pair_p;
  a.ge = sc.ge; a.gdir = sc.gdir; a.gq = sc.gatt; a.PB = PB; a.a1buf = a1buf; a.gbuf = gbuf; a.train = g != nullptr;
  const size_t smem = WA_BYTES + WC_BYTES + WD_BYTES + EDGE_GROUPS * EG_IMG + EVEC * 4 + 64;   // 226.1 KB: no alignment slack
  static unsigned long long optin = 0;
  { const int rc = smem_optin(k_tc_edge<true>, smem, optin); if (rc) return rc; }
  {
    ProfScope prof(5, d.P, st);
    k_tc_edge<true><<<edge_grid(a.g), EDGE_THREADS, smem, st>>>(a);
  }
  const bool wvg = g != nullptr && d.update && d.spatial && sc.qv != nullptr;
  {
    ProfScope prof(11, d.P, st);
    k_pair_reduce<<<(d.R + 7) / 8, 256, 0, st>>>(d, PB, sc.gproj, dx, sv.ssum, sc.qv, wvg ? g->v_mixing_kernel : nullptr);
  }
  note_launches(2);
  if (g_pair_u || g_pair_p) {
    // cotangents of the `he` terms are columns of the per-pair record: g_z1 = PB[:, 0:64], g_u = PB[:, 64:64+Kp)
    const long long total = d.P * (long long)(64 + d.Kp);
    k_pair_terms_out<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d.P, d.K, d.Kp, PB, g_pair_u, g_pair_p);
    note_launches(1);
  }
  SAKE_CUDA_CHECK(cudaGetLastError());
  if (g) {
    XtgArgs q;
    // dW2, db2 (layers.py:24):  a1^T g_e
    memset(&q, 0, sizeof(q));
    q.X = a1buf; q.ldx = 64; q.x_tt = 16; q.xw = 64; q.ones_col = 64; q.G = sc.ge; q.ldg = 64; q.g_tt = 16; q.gw = 64; q.MXpad = 128; q.NG = 64;
    q.P = d.P; q.Pdev = d.hdr ? &d.hdr->P : nullptr; q.out = g->mlp_out2_kernel; q.ldo = 64; q.out_rows = 64; q.out_cols = 64;
    q.extra = g->mlp_out2_bias; q.extra_rows = 1; q.extra_ld = 64;
    L.push(q);
    // dW1[2H : 2H+K+1] (RBF channels + distance row, layers.py:22) and the RBF mean / width sums
    memset(&q, 0, sizeof(q));
    q.X = gbuf; q.ldx = 64; q.x_tt = 16; q.xw = 64; q.ones_col = -1; q.G = PB; q.ldg = PB_LD; q.g_tt = PB_LD / 4; q.gw = PB_LD; q.MXpad = 128; q.NG = PB_LD;
    q.P = d.P; q.Pdev = d.hdr ? &d.hdr->P : nullptr; q.out = g->mlp_out0_kernel + (size_t)2 * d.H * d.H; q.ldo = 64; q.out_rows = d.K + 1; q.out_cols = 64;
    // the two rows behind the K+1 outputs (the ones column and the t column of X) are the sums S1, S2 of the RBF mean /
    // width gradients: finished inside the reduction (XtgArgs::mb_*), no scratch rows and no follow-up kernel
    q.extra = nullptr; q.extra_rows = 2; q.extra_ld = PB_LD;
    q.mb_mu = p.rbf_means; q.mb_beta = p.rbf_betas; q.mb_gmu = g->rbf_means; q.mb_gbeta = g->rbf_betas; q.mb_K = d.K;
    L.push(q);
    // dWs, dbs (layers.py:80):  e^T g_q
    memset(&q, 0, sizeof(q));
    q.X = sv.e; q.ldx = 64; q.x_tt = 16; q.xw = 64; q.ones_col = 64; q.G = sc.gatt; q.ldg = 4; q.g_tt = 1; q.gw = 4; q.MXpad = 128; q.NG = 16;   // [P,4] rows = G8 with one unit
    q.P = d.P; q.Pdev = d.hdr ? &d.hdr->P : nullptr; q.out = g->sem_kernel; q.ldo = 4; q.out_rows = 64; q.out_cols = 4;
    q.extra = g->sem_bias; q.extra_rows = 1; q.extra_ld = 4;
    if (L.push(q)) { set_error("xtg list full"); return SAKE_EINVAL; }
  }
  return 0;
}

}  // namespace sake
