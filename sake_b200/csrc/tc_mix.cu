// tcgen05 engine for the x_mixing GEMM family of DenseSAKELayer (sake/layers.py:95,111) — 89 % of
// the layer's MACs — with everything that hangs off it fused around the tensor-core tiles:
//   forward : E = e (x) att built on the fly -> coef = tanh(E Wx) -> ssum[c,d] = sum_j dir_d*m*coef_c
//   backward: recompute coef, dZ = (m dir.T)(1-coef^2), dE = dZ Wx^T + m*ghe -> g_e, g_att, g_dir
// Tiles are 128 atom pairs.  Operands are 128-byte-swizzled K-major images in shared memory:
// the pair-side image is written by builder / epilogue warps, the weight-side image is streamed
// with bulk async copies (TMA) from a pre-swizzled copy of Wx.  Accumulators live in TMEM.
// Precision: SAKE_ENGINE_TF32X3 = kind::tf32 with an exact hi/lo split of both operands (3 MMAs:
// hi*hi + lo*hi + hi*lo) -> fp32-class accuracy; SAKE_ENGINE_BF16 = kind::f16, bf16 operands.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_tile.cuh"

namespace sake {
using namespace tc;

constexpr int CC = 256;              // C = A*H (the engine is specialised to H=64, A=4)
constexpr int P_IMG = TILE * 128;    // bytes of one pair-side chunk image   (128 rows x 128 B)
constexpr int W_IMG = CC * 128;      // bytes of one weight-side chunk image (256 rows x 128 B)
constexpr int NTHREADS = 448;        // forward: 14 warps: 0 TMA producer, 1 MMA issuer, 2-5 builders, 6-13 epilogue
constexpr int BWD_THREADS = 512;     // backward: 4 warpgroups: {TMA, MMA, 2 idle}, builders, 2 x epilogue (setmaxnreg needs whole groups)

template <int ENGINE> struct Cfg;
template <> struct Cfg<SAKE_ENGINE_TF32X3> {
  static constexpr bool TF32 = true, F16 = false;
  static constexpr int KCH = 32, NSPLIT = 2, NPROD = 3, NCHUNK = 8, FMT = 2, NSTAGE = 2, EPU = 4;
};
template <> struct Cfg<SAKE_ENGINE_BF16> {
  static constexpr bool TF32 = false, F16 = false;
  static constexpr int KCH = 64, NSPLIT = 1, NPROD = 1, NCHUNK = 4, FMT = 1, NSTAGE = 4, EPU = 8;
};
// fp16 hi/lo split (x = hi + lo, 22 mantissa bits), 3 MMAs at the 16-bit rate: half the tensor cycles
// and half the operand bytes of 3xTF32.  fp16's narrow exponent range is handled by exact per-row
// power-of-two scaling of the pair-side operand (E rows by max|e|, dZ rows by max|T|).
template <> struct Cfg<SAKE_ENGINE_F16X2> {
  static constexpr bool TF32 = false, F16 = true;
  static constexpr int KCH = 64, NSPLIT = 2, NPROD = 3, NCHUNK = 4, FMT = 0, NSTAGE = 2, EPU = 8;
};
constexpr int TSM_ROWS = 6;                       // receiver rows whose T = d(loss)/d(ssum) fits the staging buffer
// auxiliary shared memory behind the stage ring.  There is no alignment slack: the dynamic window must
// start 1024-byte aligned (it does: it follows the 1 KB the runtime reserves); the kernels trap otherwise.
constexpr int AUX_BYTES = 2 * TILE * 16 /*fwd: dirm[2][TILE]; bwd: g_dir / g_att exchange*/ + TSM_ROWS * CC * 16 /*T rows*/ +
                          TSM_ROWS * CC * 4 /*ghe rows*/ + TILE * 4 /*escale*/;
template <class CF> __host__ __device__ constexpr int stage_bytes() { return CF::NSPLIT * (P_IMG + W_IMG); }
template <class CF> __host__ __device__ constexpr size_t smem_bytes() {
  return (size_t)CF::NSTAGE * stage_bytes<CF>() + AUX_BYTES + 256 /*barriers*/;
}
static_assert(2 * 2 * (P_IMG + W_IMG) + AUX_BYTES + 256 <= 232448, "shared-memory budget (227 KB per CTA)");
// products (pair-side split, weight-side split), small terms last
__device__ __constant__ int c_prod_p[3] = {0, 1, 0};
__device__ __constant__ int c_prod_w[3] = {0, 0, 1};

// L2 prefetch of the 128-byte line holding p (the next tile's rows: the demand loads then hit L2)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- operand image writers -------------------------------------------------------------------
template <class CF>
__device__ __forceinline__ void store_unit(uint8_t* img, int row, int u, const float* vals) {
  // img: base of the NSPLIT pair-side images of one stage; vals: EPU consecutive K elements
  const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)u);
  if constexpr (CF::TF32) {
    float4 hi, lo;
    split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
    split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
    *reinterpret_cast<float4*>(img + off) = hi;
    *reinterpret_cast<float4*>(img + P_IMG + off) = lo;
  } else if constexpr (CF::F16) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h2 = __floats2half2_rn(vals[2 * i], vals[2 * i + 1]);
      const float2 hf = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(vals[2 * i] - hf.x, vals[2 * i + 1] - hf.y);
      hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + P_IMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  } else {
    __nv_bfloat162 b0 = __floats2bfloat162_rn(vals[0], vals[1]);
    __nv_bfloat162 b1 = __floats2bfloat162_rn(vals[2], vals[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(vals[4], vals[5]);
    __nv_bfloat162 b3 = __floats2bfloat162_rn(vals[6], vals[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&b0); pk.y = *reinterpret_cast<uint32_t*>(&b1);
    pk.z = *reinterpret_cast<uint32_t*>(&b2); pk.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(img + off) = pk;
  }
}

// E chunk kc of one pair row: E[c = f*4+a] = e[f]*att[a]
template <class CF>
__device__ __forceinline__ void build_E_chunk(uint8_t* img, int row, int kc, const float* e, const float4& at) {
  if constexpr (CF::TF32) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float ef = e[kc * 8 + u];
      float vals[4] = {ef * at.x, ef * at.y, ef * at.z, ef * at.w};
      store_unit<CF>(img, row, u, vals);
    }
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float e0 = e[kc * 16 + 2 * u], e1 = e[kc * 16 + 2 * u + 1];
      float vals[8] = {e0 * at.x, e0 * at.y, e0 * at.z, e0 * at.w, e1 * at.x, e1 * at.y, e1 * at.z, e1 * at.w};
      store_unit<CF>(img, row, u, vals);
    }
  }
}

// fp16-split engine: exact power-of-two pair (sc, inv = 1/sc) that brings `bound` into [2^11, 2^12).
// With the operand's largest magnitude there, hi = fp16(x) and lo = fp16(x - hi) together carry 22 bits of
// the row / tensor maximum and nothing can overflow (fp16 max 65504) or drown in fp16 subnormals, however
// large or small the fp32 values are.  Identity for 0, inf and nan.
__device__ __forceinline__ void pow2_norm(float bound, float& sc, float& inv) {
  const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu);   // biased exponent: bound in [2^(e-127), 2^(e-126))
  int k = 138 - e;
  if (e == 0 || e == 255) k = 0;
  k = max(-100, min(100, k));
  sc = __uint_as_float((uint32_t)(k + 127) << 23);
  inv = __uint_as_float((uint32_t)(127 - k) << 23);
}

// ---- weight image preparation ------------------------------------------------------------------
// w1img[kc][split][row=c'][128B]: K index = c      (GEMM1: Z = E Wx)
// w2img[q ][split][row=c ][128B]: K index = c', ring order q -> chunk (q%2)*(NCHUNK/2) + q/2 (GEMM2: dE = dZ Wx^T)
// (fp16-split engine: every CTA first reduces max |Wx| itself — 256 coalesced loads per thread from L2 — instead of
// waiting for a separate one-CTA launch; block 0 publishes {scale, 1/scale} for the mix kernels)
template <int ENGINE>
__global__ void __launch_bounds__(256) k_tc_prep(const float* __restrict__ Wx, uint8_t* __restrict__ w1img,
                                                 uint8_t* __restrict__ w2img, float* __restrict__ wscale) {
  using CF = Cfg<ENGINE>;
  float wsc = 1.0f;
  if constexpr (CF::F16) {
    __shared__ float red[8];
    float mx = 0.f;
    for (int q = threadIdx.x; q < CC * CC / 4; q += 256) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(Wx) + q);
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(w4.x), fabsf(w4.y))), fmaxf(fabsf(w4.z), fabsf(w4.w)));
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) mx = fmaxf(mx, red[q]);
    float inv;
    pow2_norm(mx, wsc, inv);
    if (blockIdx.x == 0 && threadIdx.x == 0) { wscale[0] = wsc; wscale[1] = inv; }
  } else {
    if (blockIdx.x == 0 && threadIdx.x == 0) { wscale[0] = 1.0f; wscale[1] = 1.0f; }
  }
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_img = CF::NCHUNK * CC * 8;
  if (t >= 2 * per_img) return;
  const int which = t / per_img;
  int r = t % per_img;
  const int chunk = r / (CC * 8);
  r %= CC * 8;
  const int row = r / 8, u = r % 8;
  float vals[CF::EPU];
#pragma unroll
  for (int i = 0; i < CF::EPU; ++i) {
    const int k = u * CF::EPU + i;
    if (which == 0) {
      vals[i] = wsc * Wx[(size_t)(chunk * CF::KCH + k) * CC + row];
    } else {
      const int kc2 = (chunk % 2) * (CF::NCHUNK / 2) + chunk / 2;
      vals[i] = wsc * Wx[(size_t)row * CC + kc2 * CF::KCH + k];
    }
  }
  uint8_t* base = (which == 0 ? w1img : w2img) + (size_t)chunk * CF::NSPLIT * W_IMG;
  const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)u);
  if constexpr (CF::TF32) {
    float4 hi, lo;
    split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
    split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
    *reinterpret_cast<float4*>(base + off) = hi;
    *reinterpret_cast<float4*>(base + W_IMG + off) = lo;
  } else if constexpr (CF::F16) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h2 = __floats2half2_rn(vals[2 * i], vals[2 * i + 1]);
      const float2 hf = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(vals[2 * i] - hf.x, vals[2 * i + 1] - hf.y);
      hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + W_IMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  } else {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(vals[2 * i], vals[2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(base + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// max |Wx| -> {scale, 1/scale} of the weight images (fp16-split engine; {1, 1} otherwise)
__global__ void __launch_bounds__(1024) k_tc_wscale(const float* __restrict__ Wx, float* __restrict__ ws, int f16) {
  __shared__ float red[32];
  float mx = 0.f;
  if (f16)
    for (int t = threadIdx.x; t < CC * CC; t += 1024) mx = fmaxf(mx, fabsf(Wx[t]));
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = red[threadIdx.x];
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) {
      float sc = 1.0f, inv = 1.0f;
      if (f16) pow2_norm(mx, sc, inv);
      ws[0] = sc; ws[1] = inv;
    }
  }
}

// ---- shared-memory carve-up -----------------------------------------------------------------
template <class CF>
struct Smem {
  uint8_t* stages;
  float4* dirm;   // [2][TILE]  (dir*m xyz, 1/scale of the E row)           forward
  float4* gdX;    // [TILE]  g_dir partial of column half 1 -> half 0        backward (aliases dirm)
  float4* gaX;    // [TILE]  g_att partial of column half 1 -> half 0        backward
  float4* Tsm;    // [TSM_ROWS][CC] cotangent rows of the tile's receivers   backward (TMA-staged)
  float* ghS;     // [TSM_ROWS][CC] cotangent of the aggregate h_e           backward (TMA-staged)
  float* escale;  // [TILE] 1/scale of the E rows (fp16-split engine)        backward
  uint64_t *full_w, *full_e, *empty, *acc_full, *acc_empty, *t_full, *t_empty, *g_full, *g_empty;
  uint32_t* tmem_ptr;
  __device__ Smem(uint8_t* raw) {
    uint8_t* b = align1024_shared(raw);
    if (b != raw) __trap();                            // no slack is budgeted (see AUX_BYTES)
    stages = b;
    dirm = reinterpret_cast<float4*>(b + (size_t)CF::NSTAGE * stage_bytes<CF>());
    gdX = dirm;
    gaX = gdX + TILE;
    Tsm = dirm + 2 * TILE;
    ghS = reinterpret_cast<float*>(Tsm + TSM_ROWS * CC);
    escale = ghS + TSM_ROWS * CC;
    full_w = reinterpret_cast<uint64_t*>(escale + TILE);
    full_e = full_w + CF::NSTAGE;
    empty = full_e + CF::NSTAGE;
    acc_full = empty + CF::NSTAGE;      // [2]
    acc_empty = acc_full + 2;           // [2]
    t_full = acc_empty + 2;
    t_empty = t_full + 1;
    g_full = t_empty + 1;
    g_empty = g_full + 1;
    tmem_ptr = reinterpret_cast<uint32_t*>(g_empty + 1);
  }
  __device__ uint8_t* p_img(int s) const { return stages + (size_t)s * stage_bytes<CF>(); }
  __device__ uint8_t* w_img(int s) const { return stages + (size_t)s * stage_bytes<CF>() + CF::NSPLIT * P_IMG; }
};

// optional wait-time accounting of the MMA issuer (SAKE_DEBUG_WSPLITS=7): cycles spent waiting for
// [0] accumulator free, [1] weight chunk (TMA), [2] pair-side chunk (builders / epilogue), [3] total loop
__device__ unsigned long long g_mma_wait[8];
// backward kernel (same switch): MMA issuer [0] d2_empty [1] G1 weights [2] G1 pair chunks [3] G2 weights [4] G2 dZ chunks
// [5] total [6] tiles; epilogue warp 6 lane 0: [7] d1_full wait [8] ring-slot wait [9] bar1 [10] d2_full wait [11] E1 total
// [12] E2 total; builder warp 2 lane 0: [13] ring waits [14] load+build
__device__ unsigned long long g_bwd_wait[16];

__device__ __noinline__ void emit_ssum(float* o, int ds, float s0, float s1, float s2, bool accumulate) {
  if (accumulate) { atomicAdd(o, s0); atomicAdd(o + ds, s1); atomicAdd(o + 2 * ds, s2); }
  else { o[0] = s0; o[ds] = s1; o[2 * ds] = s2; }
}

// 32 TMEM columns (= 32 pairs) of one coefficient: coef = tanh(z), s += dir*m*coef.  Straight-line code:
// 32 independent MUFU chains, two partial sums per component.  dm.w = 1/scale of the pair's E row.
template <class CF, bool MASKED>
__device__ __forceinline__ void epi_fwd_chunk(const float (&v)[32], const float4* __restrict__ dmp, uint32_t mask,
                                              float& s0, float& s1, float& s2) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float4 dm = dmp[k];
    float z = v[k];
    if constexpr (CF::F16) z *= dm.w;
    float co = CF::TF32 || CF::F16 ? ftanh_(z) : ftanh_mufu_(z);
    if constexpr (MASKED) co = ((mask >> k) & 1u) ? co : 0.f;
    if (k & 1) { b0 = fmaf(dm.x, co, b0); b1 = fmaf(dm.y, co, b1); b2 = fmaf(dm.z, co, b2); }
    else { a0 = fmaf(dm.x, co, a0); a1 = fmaf(dm.y, co, a1); a2 = fmaf(dm.z, co, a2); }
  }
  s0 += a0 + b0; s1 += a1 + b1; s2 += a2 + b2;
}

// The same block for TWO short rows at once (2 * seglen <= 32): every tanh of the 32-column window is evaluated once and
// feeds the masked sums of both rows.  One window per row evaluated 32 tanh per row: 288 per tile for the 14-atom
// molecules of the flows (9 rows of 14 columns) against 126 pair columns.
template <class CF>
__device__ __forceinline__ void epi_fwd_chunk2(const float (&v)[32], const float4* __restrict__ dmp, uint32_t maskA,
                                               uint32_t maskB, float& sa0, float& sa1, float& sa2, float& sb0, float& sb1,
                                               float& sb2) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float4 dm = dmp[k];
    float z = v[k];
    if constexpr (CF::F16) z *= dm.w;
    const float co = CF::TF32 || CF::F16 ? ftanh_(z) : ftanh_mufu_(z);
    const float ca = ((maskA >> k) & 1u) ? co : 0.f, cb = ((maskB >> k) & 1u) ? co : 0.f;
    a0 = fmaf(dm.x, ca, a0); a1 = fmaf(dm.y, ca, a1); a2 = fmaf(dm.z, ca, a2);
    b0 = fmaf(dm.x, cb, b0); b1 = fmaf(dm.y, cb, b1); b2 = fmaf(dm.z, cb, b2);
  }
  sa0 = a0; sa1 = a1; sa2 = a2; sb0 = b0; sb1 = b1; sb2 = b2;
}

// =================================================================================================
// forward:  D^T[c' (lane), pair (column)] = sum_c Wx[c][c'] * E[pair][c]
//   A = weight image (M = 128 of the 256 c' per MMA, two halves), B = E image (N = 128 pairs)
//   -> every thread of the epilogue owns one coefficient c' and walks the 128 pairs of the tile, so
//      the sum over senders j (layers.py:123,127) is a thread-local accumulation.
// =================================================================================================
template <int ENGINE>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tc_mix_fwd(TileGeom g, const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ e,
             const float* __restrict__ att, const uint8_t* __restrict__ w1img, const float* __restrict__ wscale,
             float* __restrict__ ssum, int ssum_tt, int dbg_wsplits) {
  using CF = Cfg<ENGINE>;
  extern __shared__ uint8_t smem_raw[];
  Smem<CF> sm(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CF::NSTAGE; ++s) { mbar_init(sm.full_w + s, 1); mbar_init(sm.full_e + s, 128); mbar_init(sm.empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(sm.acc_full + b, 1); mbar_init(sm.acc_empty + b, 256); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(sm.tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;
  const int ntl = (geom_tiles(g) - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer (TMA)
    if (lane == 0) {
      int pos = 0;
      for (int it = 0; it < ntl; ++it)
        for (int kc = 0; kc < CF::NCHUNK; ++kc, ++pos) {
          const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
          mbar_wait(sm.empty + s, (n & 1) ^ 1);
          constexpr int nsp = CF::NSPLIT;
          mbar_arrive_expect_tx(sm.full_w + s, nsp * W_IMG);
          for (int sp = 0; sp < nsp; ++sp)
            bulk_g2s(sm.w_img(s) + sp * W_IMG, w1img + ((size_t)kc * CF::NSPLIT + sp) * W_IMG, W_IMG, sm.full_w + s);
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(CF::FMT, 128, TILE);
      int pos = 0;
      long long w_acc = 0, w_w = 0, w_e = 0;
      const long long t_begin = clock64();
      for (int it = 0; it < ntl; ++it) {
        const int buf = it & 1, use = it >> 1;
        long long t0 = clock64();
        mbar_wait(sm.acc_empty + buf, (use & 1) ^ 1);
        w_acc += clock64() - t0;
        tc_fence_after();
        for (int kc = 0; kc < CF::NCHUNK; ++kc, ++pos) {
          const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
          t0 = clock64();
          mbar_wait(sm.full_w + s, n & 1);
          long long t1 = clock64();
          mbar_wait(sm.full_e + s, n & 1);
          w_w += t1 - t0;
          w_e += clock64() - t1;
          tc_fence_after();
          const uint32_t wbase = smem_u32(sm.w_img(s)), pbase = smem_u32(sm.p_img(s));
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint32_t d = tmem_base + buf * 256 + mh * 128;
#pragma unroll
            for (int pr = 0; pr < CF::NPROD; ++pr) {
              const uint32_t a0 = wbase + c_prod_w[pr] * W_IMG + mh * (128 * 128);
              const uint32_t b0 = pbase + c_prod_p[pr] * P_IMG;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma<CF::TF32>(d, umma_desc_k_sw128(a0 + ks * 32), umma_desc_k_sw128(b0 + ks * 32), idesc,
                               (kc | pr | ks) != 0);
            }
          }
          umma_commit(sm.empty + s);
        }
        umma_commit(sm.acc_full + buf);
      }
      if (dbg_wsplits == 7) {
        atomicAdd(&g_mma_wait[0], (unsigned long long)w_acc);
        atomicAdd(&g_mma_wait[1], (unsigned long long)w_w);
        atomicAdd(&g_mma_wait[2], (unsigned long long)w_e);
        atomicAdd(&g_mma_wait[3], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_mma_wait[4], (unsigned long long)ntl);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ builders: E image + pair geometry
    const int p = (warp - 2) * 32 + lane;
    int pos = 0;
    for (int it = 0; it < ntl; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int buf = it & 1, use = it >> 1;
      bool valid;
      int row, j;
      long long prx;
      tile_pair(tile_desc(g, tile), p, valid, row, j, prx);
      if (it + 1 < ntl) {                           // pull the next tile's e / att rows into L2
        bool v2;
        int row2, j2;
        long long px2;
        tile_pair(tile_desc(g, tile + gridDim.x), p, v2, row2, j2, px2);
        if (v2) {                                   // G8: a line = one unit of 8 pairs; each pair of a group pulls two of the 16
          const float4* e8 = reinterpret_cast<const float4*>(e) + g8_row(px2 & ~7LL, 16) + 2 * (px2 & 7) * G8S;
          prefetch_l2(e8); prefetch_l2(e8 + G8S); prefetch_l2(att + px2 * 4);
        }
      }
      float ev[64];
      float4 at = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        at = *reinterpret_cast<const float4*>(att + prx * 4);
        const float4* ep = reinterpret_cast<const float4*>(e) + g8_row(prx, 16);     // G8 layout: unit q at ep[q * 8]
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float4 t4 = __ldg(ep + q * G8S);
          ev[4 * q] = t4.x; ev[4 * q + 1] = t4.y; ev[4 * q + 2] = t4.z; ev[4 * q + 3] = t4.w;
        }
        const float* xi = x + (size_t)row * 3;
        const float* xj = x + (size_t)(geom_mol0(g, row) + j) * 3;
        const float r0 = xj[0] - xi[0], r1 = xj[1] - xi[1], r2 = xj[2] - xi[2];
        const float nrm = sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 0.f) + 1e-5f);   // functional.py:14-17
        const float inv = 1.0f / (nrm + 1e-5f);                                      // layers.py:115
        const float m = mask ? mask[prx] : 1.0f;
        dm = make_float4(r0 * inv * m, r1 * inv * m, r2 * inv * m, 1.0f);
      } else {
#pragma unroll
        for (int q = 0; q < 64; ++q) ev[q] = 0.f;
      }
      float inv_s = 1.0f;
      if constexpr (CF::F16) {                      // normalise the row E = e (x) att (exact 2^k, applied through att)
        float mx = 0.f;
#pragma unroll
        for (int q = 0; q < 64; ++q) mx = fmaxf(mx, fabsf(ev[q]));
        float sc;
        pow2_norm(mx * fmaxf(fmaxf(fabsf(at.x), fabsf(at.y)), fmaxf(fabsf(at.z), fabsf(at.w))), sc, inv_s);
        at.x *= sc; at.y *= sc; at.z *= sc; at.w *= sc;
        inv_s *= wscale[1];                         // ... and undo the weight-image scale with the same factor
      }
      mbar_wait_warp(sm.acc_empty + buf, (use & 1) ^ 1);     // epilogue of the tile that last used dirm[buf] is done
      dm.w = inv_s;                                          // fp16-split engine: undoes the E row scale
      sm.dirm[buf * TILE + p] = dm;
#pragma unroll
      for (int kc = 0; kc < CF::NCHUNK; ++kc, ++pos) {
        const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
        mbar_wait_warp(sm.empty + s, (n & 1) ^ 1);
        build_E_chunk<CF>(sm.p_img(s), p, kc, ev, at);
        fence_proxy_async();
        mbar_arrive(sm.full_e + s);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: tanh + sum over senders
    // Thread = coefficient c'.  The tile's receiver rows are walked as segments of TMEM columns; every
    // 32-column load is one branch-free block of 32 independent tanh chains (the segment boundaries are
    // handled by a warp-uniform bit mask on a clamped, possibly overlapping load), and the row sums are
    // written once per segment.
    const int q = warp & 3, mh = (warp - 6) >> 2;
    const int cp = mh * 128 + q * 32 + lane;
    const bool accumulate = g.nseg > 1;
    for (int it = 0; it < ntl; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int buf = it & 1, use = it >> 1;
      const TileDesc td = tile_desc(g, tile);
      const int row0 = td.row0, nsegs = td.nrows, seglen = td.n;
      mbar_wait_warp(sm.acc_full + buf, use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + mh * 128;
      const float4* dmt = sm.dirm + buf * TILE;
      if (!accumulate && 2 * seglen <= 32) {
        // short rows: two per 32-column window (epi_fwd_chunk2)
#pragma unroll 1
        for (int sg = 0; sg < nsegs; sg += 2) {
          const int c0 = sg * seglen;
          const int ls = min(c0, TILE - 32);           // clamped window start; both rows end at or before column 128
          const bool two = sg + 1 < nsegs;
          const uint32_t m1 = ((1u << seglen) - 1u) << (c0 - ls);
          const uint32_t m2 = two ? m1 << seglen : 0u;
          float v[32];
          tmem_ld32(taddr + ls, v);
          tmem_ld_wait();
          float sa0, sa1, sa2, sb0, sb1, sb2;
          epi_fwd_chunk2<CF>(v, dmt + ls, m1, m2, sa0, sa1, sa2, sb0, sb1, sb2);
          for (int h2 = 0; h2 < (two ? 2 : 1); ++h2) {
            const size_t orow = (size_t)(row0 + sg + h2);
            float* o = ssum_tt ? ssum + (g8_row((long long)orow, 192) + (size_t)(cp >> 2) * 3 * G8S) * 4 + (cp & 3)
                               : ssum + (orow * CC + cp) * 3;
            emit_ssum(o, ssum_tt ? 4 * G8S : 1, h2 ? sb0 : sa0, h2 ? sb1 : sa1, h2 ? sb2 : sa2, false);
          }
        }
      } else
#pragma unroll 1
      for (int sg = 0; sg < nsegs; ++sg) {
        const int c0 = sg * seglen, c1 = c0 + seglen;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = c0; c < c1; c += 32) {
          const int ls = min(c, TILE - 32);            // clamped load start: never reads past the tile's 128 columns
          const int n = min(32, c1 - c);
          float v[32];
          tmem_ld32(taddr + ls, v);
          tmem_ld_wait();
          if (n == 32 && ls == c) epi_fwd_chunk<CF, false>(v, dmt + ls, 0u, s0, s1, s2);
          else epi_fwd_chunk<CF, true>(v, dmt + ls, (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) << (c - ls), s0, s1, s2);
        }
        // ssum_tt: the G8 layout the tcgen05 node kernels read (tc_node.cu), unit (c'/4)*3 + d = component d of 4 coefficients
        const size_t orow = (size_t)(row0 + sg);
        float* o = ssum_tt ? ssum + (g8_row((long long)orow, 192) + (size_t)(cp >> 2) * 3 * G8S) * 4 + (cp & 3)
                           : ssum + (orow * CC + cp) * 3;
        emit_ssum(o, ssum_tt ? 4 * G8S : 1, s0, s1, s2, accumulate);
      }
      tc_fence_before();
      mbar_arrive(sm.acc_empty + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// backward epilogue 1 on 32 TMEM columns (= 32 coefficients c') of one pair: coef = tanh z,
// dZ = (dir*m . T[c']) sech^2 z, g_dir += coef * T[c'].  Straight-line: the loads of T carry no predicate
// (idle lanes read a valid row with dir = 0), so the 32 chains interleave freely.  dZ goes straight into
// the GEMM2 operand image, EPU values at a time (unit0 = first 16-byte unit of this block in the row).
//   q? = 4 * dzs * dir * m: the derivative is computed as sech^2 z / 4 = r - r^2 with r = 1/(e^{2|z|}+1),
//   which has no cancellation in saturated coefficients and needs no clamp (e^{2|z|} = inf -> r = 0 ->
//   coef = +-1, dZ = 0); dzs is the fp16-split engine's power-of-two row scale (1 otherwise), undone
//   (idz) for the fp32 copy of dZ that the weight-gradient contraction reads.
template <class CF>
__device__ __forceinline__ void epi1_block(const float (&v)[32], const float4* __restrict__ Tp, float zs, float q0,
                                           float q1, float q2, float& g0, float& g1, float& g2, uint8_t* img, int p,
                                           int unit0, float4* __restrict__ gz, float idz) {
  float dz[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {                       // 32 independent chains, no control flow in between
    const float4 t4 = Tp[k];
    const float z = CF::F16 ? v[k] * zs : v[k];
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(z) * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
    const float co = copysignf(fmaf(-2.0f, r, 1.0f), z);
    const float s2q = fmaf(-r, r, r);
    const float gco = fmaf(q2, t4.z, fmaf(q1, t4.y, q0 * t4.x));
    g0 = fmaf(co, t4.x, g0); g1 = fmaf(co, t4.y, g1); g2 = fmaf(co, t4.z, g2);
    dz[k] = gco * s2q;
  }
  if (gz != nullptr) {                                 // training: fp32 copy of dZ for the weight-gradient contraction
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      gz[(i / 4) * G8S] =
          CF::F16 ? make_float4(dz[i] * idz, dz[i + 1] * idz, dz[i + 2] * idz, dz[i + 3] * idz)
                  : make_float4(dz[i], dz[i + 1], dz[i + 2], dz[i + 3]);
  }
#pragma unroll
  for (int u = 0; u < 32 / CF::EPU; ++u) store_unit<CF>(img, p, unit0 + u, dz + u * CF::EPU);
}

// backward epilogue 2 on 32 TMEM columns (= 8 features f x 4 heads) of one pair:
// dE += m*ghe;  g_e[f] = sum_a dE att[a];  g_att[a] += sum_f dE e[f]
__device__ __forceinline__ void epi2_part(const float (&v)[32], const float4* __restrict__ gh4, const float4& ea,
                                          const float4& eb, const float4& at, float m, float idz, float* gev,
                                          float& ga0, float& ga1, float& ga2, float& ga3) {
  const float ef[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
  for (int k8 = 0; k8 < 8; ++k8) {
    const float4 gh = gh4[k8];
    const float x0 = fmaf(m, gh.x, v[4 * k8] * idz), x1 = fmaf(m, gh.y, v[4 * k8 + 1] * idz),
                x2 = fmaf(m, gh.z, v[4 * k8 + 2] * idz), x3 = fmaf(m, gh.w, v[4 * k8 + 3] * idz);
    gev[k8] = x0 * at.x + x1 * at.y + x2 * at.z + x3 * at.w;
    ga0 = fmaf(x0, ef[k8], ga0); ga1 = fmaf(x1, ef[k8], ga1);
    ga2 = fmaf(x2, ef[k8], ga2); ga3 = fmaf(x3, ef[k8], ga3);
  }
}

// =================================================================================================
// backward (dX): per tile  GEMM1  Z[pair (lane), c'] = E Wx          (recompute)
//                          epi-1  coef = tanh Z; dZ = (m dir.T[c'])(1-coef^2); g_dir += coef*T
//                          GEMM2  dE[pair, c] = dZ Wx^T
//                          epi-2  dE += m*ghe;  g_e[f] = sum_a dE*att;  g_att[a] = sum_f dE*e
//   every epilogue thread owns one pair (TMEM lane), so all reductions over c / c' are thread-local.
// =================================================================================================
template <int ENGINE>
__global__ void __launch_bounds__(BWD_THREADS, 1)
k_tc_mix_bwd(TileGeom g, const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ e,
             const float* __restrict__ att, const uint8_t* __restrict__ w1img, const uint8_t* __restrict__ w2img,
             const float* __restrict__ wscale, const float4* __restrict__ T4, const float* __restrict__ tmax,
             const float* __restrict__ ghe,
             float* __restrict__ ge, float* __restrict__ gatt, float* __restrict__ gdir, float* __restrict__ gZ_out, int dbg) {
  using CF = Cfg<ENGINE>;
  constexpr int NCH = CF::NCHUNK;
  extern __shared__ uint8_t smem_raw[];
  Smem<CF> sm(smem_raw);
  uint64_t* d1_full = sm.acc_full;        // GEMM1 accumulator ready
  uint64_t* d2_full = sm.acc_full + 1;    // GEMM2 accumulator ready
  uint64_t* d2_empty = sm.acc_empty;      // GEMM2 accumulator drained
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CF::NSTAGE; ++s) { mbar_init(sm.full_w + s, 1); mbar_init(sm.full_e + s, 128); mbar_init(sm.empty + s, 1); }
    mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d2_empty, 256); mbar_init(sm.acc_empty + 1, 1);
    mbar_init(sm.t_full, 1); mbar_init(sm.t_empty, 256); mbar_init(sm.g_full, 1); mbar_init(sm.g_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(sm.tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;
  const int ntl = (geom_tiles(g) - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // T / ghe rows of a tile are staged in smem by TMA when the tile has few enough receiver rows.  Ragged
  // batches decide per tile, so every role counts the staged tiles itself (`tn`: phase of the t_* / g_* barriers).
  // register re-balancing between the warpgroups (the kernel starts with 128 per thread): the TMA / MMA
  // group keeps 40, the epilogue groups grow to 168 so that a 32-column block keeps all its chains in flight
  // (each setmaxnreg sits inside its role's branch: ptxas budgets registers per branch only then)
  if (warp < 4) {
   setmaxnreg_dec<40>();
   if (warp == 0) {
    if (lane == 0) {
      int pos = 0, tn = 0;
      for (int it = 0; it < ntl; ++it) {
        const TileDesc td = tile_desc(g, blockIdx.x + it * gridDim.x);
        const bool use_tsm = td.nrows <= TSM_ROWS;
        if (use_tsm) {
          mbar_wait(sm.t_empty, (tn & 1) ^ 1);
          mbar_arrive_expect_tx(sm.t_full, (uint32_t)td.nrows * CC * 16);
          bulk_g2s(sm.Tsm, T4 + (size_t)td.row0 * CC, (uint32_t)td.nrows * CC * 16, sm.t_full);
        }
        for (int c2 = 0; c2 < 2 * NCH; ++c2, ++pos) {
          if (c2 == NCH && use_tsm) {
            // ghe rows for epilogue 2: requested only now, so that waiting for epilogue 2 of the previous tile
            // (g_empty) cannot hold back the GEMM1 weight stream; GEMM2 needs that epilogue finished anyway
            mbar_wait(sm.g_empty, (tn & 1) ^ 1);
            mbar_arrive_expect_tx(sm.g_full, (uint32_t)td.nrows * CC * 4);
            bulk_g2s(sm.ghS, ghe + (size_t)td.row0 * CC, (uint32_t)td.nrows * CC * 4, sm.g_full);
          }
          const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
          const uint8_t* src = (c2 < NCH ? w1img + (size_t)c2 * CF::NSPLIT * W_IMG
                                         : w2img + (size_t)(c2 - NCH) * CF::NSPLIT * W_IMG);
          mbar_wait(sm.empty + s, (n & 1) ^ 1);
          mbar_arrive_expect_tx(sm.full_w + s, CF::NSPLIT * W_IMG);
          for (int sp = 0; sp < CF::NSPLIT; ++sp)
            bulk_g2s(sm.w_img(s) + sp * W_IMG, src + (size_t)sp * W_IMG, W_IMG, sm.full_w + s);
        }
        tn += use_tsm ? 1 : 0;
      }
    }
   } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(CF::FMT, TILE, CC);
      int pos = 0;
      long long w[5] = {0, 0, 0, 0, 0};
      const long long t_begin = clock64();
      for (int it = 0; it < ntl; ++it) {
        for (int c2 = 0; c2 < 2 * NCH; ++c2, ++pos) {
          const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
          long long t0 = clock64();
          if (c2 == NCH) {                              // before GEMM2 overwrites D2: previous tile drained
            mbar_wait(d2_empty, (it & 1) ^ 1);
            tc_fence_after();
          }
          long long t1 = clock64();
          mbar_wait(sm.full_w + s, n & 1);
          long long t2 = clock64();
          mbar_wait(sm.full_e + s, n & 1);
          long long t3 = clock64();
          w[0] += t1 - t0; w[c2 < NCH ? 1 : 3] += t2 - t1; w[c2 < NCH ? 2 : 4] += t3 - t2;
          tc_fence_after();
          const uint32_t wbase = smem_u32(sm.w_img(s)), pbase = smem_u32(sm.p_img(s));
          const uint32_t d = tmem_base + (c2 < NCH ? 0 : 256);
          const int kc = c2 < NCH ? c2 : c2 - NCH;
#pragma unroll
          for (int pr = 0; pr < CF::NPROD; ++pr) {
            const uint32_t a0 = pbase + c_prod_p[pr] * P_IMG;
            const uint32_t b0 = wbase + c_prod_w[pr] * W_IMG;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma<CF::TF32>(d, umma_desc_k_sw128(a0 + ks * 32), umma_desc_k_sw128(b0 + ks * 32), idesc,
                             (kc | pr | ks) != 0);
          }
          umma_commit(sm.empty + s);
          if (c2 == NCH - 1) umma_commit(d1_full);
          if (c2 == 2 * NCH - 1) umma_commit(d2_full);
        }
      }
      if (dbg == 7) {
        for (int i = 0; i < 5; ++i) atomicAdd(&g_bwd_wait[i], (unsigned long long)w[i]);
        atomicAdd(&g_bwd_wait[5], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_bwd_wait[6], (unsigned long long)ntl);
      }
    }
   }   // warps 2, 3: idle, they only lend their registers
  } else if (warp < 8) {
    // ------------------------------------------------------------ builders: E image for GEMM1
    const int p = (warp - 4) * 32 + lane;
    long long bw_ring = 0;
    const long long bw_begin = clock64();
    for (int it = 0; it < ntl; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      bool valid;
      int row, j;
      long long prx;
      tile_pair(tile_desc(g, tile), p, valid, row, j, prx);
      if (it + 1 < ntl) {                           // pull the next tile's e / att rows into L2
        bool v2;
        int row2, j2;
        long long px2;
        tile_pair(tile_desc(g, tile + gridDim.x), p, v2, row2, j2, px2);
        if (v2) {                                   // G8: a line = one unit of 8 pairs; each pair of a group pulls two of the 16
          const float4* e8 = reinterpret_cast<const float4*>(e) + g8_row(px2 & ~7LL, 16) + 2 * (px2 & 7) * G8S;
          prefetch_l2(e8); prefetch_l2(e8 + G8S); prefetch_l2(att + px2 * 4);
        }
      }
      float ev[64];
      float4 at = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        at = *reinterpret_cast<const float4*>(att + prx * 4);
        const float4* ep = reinterpret_cast<const float4*>(e) + g8_row(prx, 16);     // G8 layout: unit q at ep[q * 8]
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float4 t4 = __ldg(ep + q * G8S);
          ev[4 * q] = t4.x; ev[4 * q + 1] = t4.y; ev[4 * q + 2] = t4.z; ev[4 * q + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 64; ++q) ev[q] = 0.f;
      }
      if constexpr (CF::F16) {                      // normalise the row E = e (x) att (exact 2^k), undone in epilogue 1
        float mx = 0.f;
#pragma unroll
        for (int q = 0; q < 64; ++q) mx = fmaxf(mx, fabsf(ev[q]));
        float sc, inv;
        pow2_norm(mx * fmaxf(fmaxf(fabsf(at.x), fabsf(at.y)), fmaxf(fabsf(at.z), fabsf(at.w))), sc, inv);
        at.x *= sc; at.y *= sc; at.z *= sc; at.w *= sc;
        sm.escale[p] = inv * wscale[1];   // single buffer: the builders only get here after epilogue 1 of the previous tile started
      }
      int pos = it * 2 * NCH;
#pragma unroll
      for (int kc = 0; kc < NCH; ++kc, ++pos) {
        const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
        const long long t0 = dbg == 7 ? clock64() : 0;
        mbar_wait(sm.empty + s, (n & 1) ^ 1);
        if (dbg == 7) bw_ring += clock64() - t0;
        build_E_chunk<CF>(sm.p_img(s), p, kc, ev, at);
        fence_proxy_async();
        mbar_arrive(sm.full_e + s);
      }
      // The GEMM2 ring slots of this tile are filled by the epilogue warps.  A parity wait can only tell
      // adjacent phases apart, so walk through those slots too (no work) to stay phase-synchronised with
      // the `empty` barriers before building the next tile's chunks.
      const long long t1 = dbg == 7 ? clock64() : 0;
      for (int kc = 0; kc < NCH; ++kc, ++pos) {
        const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
        mbar_wait(sm.empty + s, (n & 1) ^ 1);
      }
      if (dbg == 7) bw_ring += clock64() - t1;
    }
    if (dbg == 7 && threadIdx.x == 128) {
      atomicAdd(&g_bwd_wait[13], (unsigned long long)bw_ring);
      atomicAdd(&g_bwd_wait[14], (unsigned long long)(clock64() - bw_begin - bw_ring));
    }
  } else {
    setmaxnreg_inc<168>();
    // ------------------------------------------------------------ epilogue warps (thread = pair)
    // Every thread walks 4 blocks of 32 TMEM columns in each epilogue; the load of the next block is in
    // flight while the current one is processed (two register buffers).
    const int q = warp & 3, hh = (warp - 8) >> 2;
    const int p = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int PPC = CF::KCH / 32;                   // 32-column blocks per ring chunk (1: tf32, 2: 16-bit formats)
    long long ew[6] = {0, 0, 0, 0, 0, 0};
    int tn = 0;
    for (int it = 0; it < ntl; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const TileDesc td = tile_desc(g, tile);
      const bool use_tsm = td.nrows <= TSM_ROWS;
      bool valid;
      int row, j;
      long long prx;
      tile_pair(td, p, valid, row, j, prx);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, m = 0.f;
      int trow = 0;                                   // row of the T operand (any valid row for idle lanes: d = 0)
      if (!valid) prx = 0;
      if (valid) {
        const float* xi = x + (size_t)row * 3;
        const float* xj = x + (size_t)(geom_mol0(g, row) + j) * 3;
        const float r0 = xj[0] - xi[0], r1 = xj[1] - xi[1], r2 = xj[2] - xi[2];
        const float nrm = sqrtf(fmaxf(r0 * r0 + r1 * r1 + r2 * r2, 0.f) + 1e-5f);
        const float inv = 1.0f / (nrm + 1e-5f);
        m = mask ? mask[prx] : 1.0f;
        d0 = r0 * inv * m; d1 = r1 * inv * m; d2 = r2 * inv * m;
        trow = use_tsm ? row - td.row0 : row;
      }
      const float4* Tg = T4 + (size_t)trow * CC;      // global copy of the row
      const float4* Ts = sm.Tsm + trow * CC;          // TMA-staged copy (use_tsm)
      // ---------------- epilogue 1: dZ chunks for GEMM2 (this half owns ring slots of parity hh)
      const long long e_t0 = dbg == 7 ? clock64() : 0;
      if (use_tsm) mbar_wait(sm.t_full, tn & 1);
      mbar_wait(d1_full, it & 1);
      tc_fence_after();
      const long long e_t1 = dbg == 7 ? clock64() : 0;
      ew[0] += e_t1 - e_t0;
      float zs = 1.0f, dzs = 1.0f, idz = 1.0f;      // fp16-split engine: 1/scale of the E row, scale of the dZ row
      if constexpr (CF::F16) {
        zs = sm.escale[p];
        if (valid) pow2_norm(2.0f * tmax[row], dzs, idz);       // |dZ| <= |dir|_1 * max|T| * max sech^2 < 2 max|T|
      }
      const float idzw = CF::F16 ? idz * wscale[1] : 1.0f;      // GEMM2 output: undo the dZ row and weight-image scales
      const float q0 = 4.0f * dzs * d0, q1 = 4.0f * dzs * d1, q2 = 4.0f * dzs * d2;    // see epi1_block
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      // training: fp32 copy of dZ for the weight-gradient contraction, G8 layout (64 units per pair)
      float4* gzrow = (gZ_out != nullptr && valid) ? reinterpret_cast<float4*>(gZ_out) + g8_row(prx, CC / 4) : nullptr;
      // block jb (0..3) of this thread: ring chunk kc2, first column cb, ring position pos
      auto blk_col = [&](int jb) { return (hh * (NCH / 2) + jb / PPC) * CF::KCH + (jb % PPC) * 32; };
      auto run_block1 = [&](int jb, const float (&v)[32]) {
        const int cb = blk_col(jb);
        const int pos = it * 2 * NCH + NCH + 2 * (jb / PPC) + hh;
        const int s = pos % CF::NSTAGE, n = pos / CF::NSTAGE;
        if (jb % PPC == 0) {
          const long long e_s0 = dbg == 7 ? clock64() : 0;
          mbar_wait(sm.empty + s, (n & 1) ^ 1);
          if (dbg == 7) ew[1] += clock64() - e_s0;
        }
        float4* gz = gzrow ? gzrow + (cb / 4) * G8S : nullptr;
        if (use_tsm) epi1_block<CF>(v, Ts + cb, zs, q0, q1, q2, g0, g1, g2, sm.p_img(s), p, (jb % PPC) * (32 / CF::EPU), gz, idz);
        else epi1_block<CF>(v, Tg + cb, zs, q0, q1, q2, g0, g1, g2, sm.p_img(s), p, (jb % PPC) * (32 / CF::EPU), gz, idz);
        if (jb % PPC == PPC - 1) {
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(sm.full_e + s);
        }
      };
      {
        float va[32], vb[32];
        tmem_ld32(lane_addr + blk_col(0), va);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          tmem_ld_wait_dep(va);
          tmem_ld32(lane_addr + blk_col(2 * jj + 1), vb);
          run_block1(2 * jj, va);
          tmem_ld_wait_dep(vb);
          if (jj == 0) tmem_ld32(lane_addr + blk_col(2), va);
          run_block1(2 * jj + 1, vb);
        }
      }
      if (use_tsm) mbar_arrive(sm.t_empty);          // T rows of this tile are consumed
      // operands of epilogue 2 that do not depend on GEMM2: fetched now, their latency hides behind the hand-off
      float4 at = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 ef4[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) ef4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        at = *reinterpret_cast<const float4*>(att + prx * 4);
        const float4* e4 = reinterpret_cast<const float4*>(e) + g8_row(prx, 16) + hh * 8 * G8S;
#pragma unroll
        for (int k = 0; k < 8; ++k) ef4[k] = __ldg(e4 + k * G8S);
      }
      // ---------------- epilogue 2: dE -> g_e, g_att   (this half owns f in [32 hh, 32 hh + 32))
      const long long e_t3 = dbg == 7 ? clock64() : 0;
      ew[4] += e_t3 - e_t1;
      if (use_tsm) mbar_wait(sm.g_full, tn & 1);
      mbar_wait(d2_full, it & 1);
      tc_fence_after();
      const long long e_t4 = dbg == 7 ? clock64() : 0;
      ew[3] += e_t4 - e_t3;
      float ga0 = 0.f, ga1 = 0.f, ga2 = 0.f, ga3 = 0.f;
      const float4* ghs4 = reinterpret_cast<const float4*>(sm.ghS + trow * CC) + hh * 32;      // staged row (use_tsm)
      const float4* ghg4 = reinterpret_cast<const float4*>(ghe + (size_t)trow * CC) + hh * 32;  // global row
      auto run_block2 = [&](int cc, const float (&v)[32]) {
        float gev[8];
        if (use_tsm) epi2_part(v, ghs4 + cc * 8, ef4[2 * cc], ef4[2 * cc + 1], at, m, idzw, gev, ga0, ga1, ga2, ga3);
        else epi2_part(v, ghg4 + cc * 8, ef4[2 * cc], ef4[2 * cc + 1], at, m, idzw, gev, ga0, ga1, ga2, ga3);
        if (valid) {
          float4* o = reinterpret_cast<float4*>(ge) + g8_row(prx, 16) + (hh * 8 + cc * 2) * G8S;   // G8 layout
          o[0] = make_float4(gev[0], gev[1], gev[2], gev[3]);
          o[G8S] = make_float4(gev[4], gev[5], gev[6], gev[7]);
        }
      };
      {
        const uint32_t a2 = lane_addr + 256 + hh * 128;
        float va[32], vb[32];
        tmem_ld32(a2, va);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          tmem_ld_wait_dep(va);
          tmem_ld32(a2 + (2 * jj + 1) * 32, vb);
          run_block2(2 * jj, va);
          tmem_ld_wait_dep(vb);
          if (jj == 0) tmem_ld32(a2 + 64, va);
          run_block2(2 * jj + 1, vb);
        }
      }
      tc_fence_before();
      mbar_arrive(d2_empty);
      if (use_tsm) { mbar_arrive(sm.g_empty); ++tn; }   // ghe rows of this tile are consumed
      // the two column halves of a pair meet once per tile: half 1 hands its partial g_dir / g_att to half 0
      if (hh == 1) { sm.gdX[p] = make_float4(g0, g1, g2, 0.f); sm.gaX[p] = make_float4(ga0, ga1, ga2, ga3); }
      const long long e_t5 = dbg == 7 ? clock64() : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (dbg == 7) ew[2] += clock64() - e_t5;
      if (hh == 0 && valid) {
        const float4 bd = sm.gdX[p], ba = sm.gaX[p];
        float* o = gdir + prx * 3;
        o[0] = (g0 + bd.x) * m; o[1] = (g1 + bd.y) * m; o[2] = (g2 + bd.z) * m;
        *reinterpret_cast<float4*>(gatt + prx * 4) = make_float4(ga0 + ba.x, ga1 + ba.y, ga2 + ba.z, ga3 + ba.w);
      }
      if (dbg == 7) ew[5] += clock64() - e_t4;
    }
    if (dbg == 7 && threadIdx.x == 256) {
      atomicAdd(&g_bwd_wait[7], (unsigned long long)ew[0]); atomicAdd(&g_bwd_wait[8], (unsigned long long)ew[1]);
      atomicAdd(&g_bwd_wait[9], (unsigned long long)ew[2]); atomicAdd(&g_bwd_wait[10], (unsigned long long)ew[3]);
      atomicAdd(&g_bwd_wait[11], (unsigned long long)ew[4]); atomicAdd(&g_bwd_wait[12], (unsigned long long)ew[5]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =================================================================================================
// self-test: D[128 x 256] = A[128 x K] * B[256 x K]^T through the same descriptors / images / TMEM loads
// =================================================================================================
template <int ENGINE>
__global__ void __launch_bounds__(128, 1) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ D, int nchunk) {
  using CF = Cfg<ENGINE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  uint8_t* pimg = base;                                  // [nchunk][NSPLIT][P_IMG]
  uint8_t* wimg = base + (size_t)nchunk * CF::NSPLIT * P_IMG;    // [nchunk][NSPLIT][W_IMG]
  uint64_t* bar = reinterpret_cast<uint64_t*>(wimg + (size_t)nchunk * CF::NSPLIT * W_IMG);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int Kt = nchunk * CF::KCH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(tptr);
  for (int kc = 0; kc < nchunk; ++kc)
    for (int u = 0; u < 8; ++u) {
      float vals[CF::EPU];
      const int r = threadIdx.x;
      for (int i = 0; i < CF::EPU; ++i) vals[i] = A[(size_t)r * Kt + kc * CF::KCH + u * CF::EPU + i];
      store_unit<CF>(pimg + (size_t)kc * CF::NSPLIT * P_IMG, r, u, vals);
      for (int rr = r; rr < CC; rr += 128) {
        for (int i = 0; i < CF::EPU; ++i) vals[i] = B[(size_t)rr * Kt + kc * CF::KCH + u * CF::EPU + i];
        uint8_t* wb = wimg + (size_t)kc * CF::NSPLIT * W_IMG;
        const uint32_t off = sw128_offset((uint32_t)rr, (uint32_t)u);
        if constexpr (CF::TF32) {
          float4 hi, lo;
          split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
          split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
          *reinterpret_cast<float4*>(wb + off) = hi;
          *reinterpret_cast<float4*>(wb + W_IMG + off) = lo;
        } else if constexpr (CF::F16) {
          uint32_t hi[4], lo[4];
          for (int i = 0; i < 4; ++i) {
            const __half2 h2 = __floats2half2_rn(vals[2 * i], vals[2 * i + 1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(vals[2 * i] - hf.x, vals[2 * i + 1] - hf.y);
            hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          *reinterpret_cast<uint4*>(wb + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(wb + W_IMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        } else {
          uint32_t pk[4];
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(vals[2 * i], vals[2 * i + 1]);
            pk[i] = *reinterpret_cast<uint32_t*>(&b2);
          }
          *reinterpret_cast<uint4*>(wb + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc(CF::FMT, TILE, CC);
    for (int kc = 0; kc < nchunk; ++kc)
      for (int pr = 0; pr < CF::NPROD; ++pr)
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a0 = smem_u32(pimg + (size_t)kc * CF::NSPLIT * P_IMG + c_prod_p[pr] * P_IMG) + ks * 32;
          const uint32_t b0 = smem_u32(wimg + (size_t)kc * CF::NSPLIT * W_IMG + c_prod_w[pr] * W_IMG) + ks * 32;
          umma<CF::TF32>(tmem_base, umma_desc_k_sw128(a0), umma_desc_k_sw128(b0), idesc, (kc | pr | ks) != 0);
        }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int cc = 0; cc < 8; ++cc) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
    tmem_ld_wait();
    for (int k = 0; k < 32; ++k) D[(size_t)(warp * 32 + lane) * CC + cc * 32 + k] = v[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// =================================================================================================
// host side
// =================================================================================================
// K <= 58: the tcgen05 edge kernel's K = 64 operand holds the RBF channels plus three extra columns; the tcgen05
// kernels share G8-layout buffers, so an engine is tcgen05 for all of them or for none
bool tc_supported(const Dims& d) { return d.H == 64 && d.A == 4 && d.K <= 58; }

int tc_debug_counters(unsigned long long* out8) {
  SAKE_CUDA_CHECK(cudaMemcpyFromSymbol(out8, g_mma_wait, sizeof(unsigned long long) * 8));
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  SAKE_CUDA_CHECK(cudaMemcpyToSymbol(g_mma_wait, z, sizeof(z)));
  return 0;
}
int tc_debug_counters_bwd(unsigned long long* out16) {
  SAKE_CUDA_CHECK(cudaMemcpyFromSymbol(out16, g_bwd_wait, sizeof(unsigned long long) * 16));
  unsigned long long z[16];
  memset(z, 0, sizeof(z));
  SAKE_CUDA_CHECK(cudaMemcpyToSymbol(g_bwd_wait, z, sizeof(z)));
  return 0;
}

template <class CF> static size_t wimg_bytes() { return (size_t)CF::NCHUNK * CF::NSPLIT * W_IMG; }

size_t tc_scratch_bytes(const Dims& d, int engine, int for_backward, int with_param_grads) {
  (void)d; (void)for_backward; (void)with_param_grads;
  size_t w = engine == SAKE_ENGINE_BF16 ? wimg_bytes<Cfg<SAKE_ENGINE_BF16>>()
             : engine == SAKE_ENGINE_F16X2 ? wimg_bytes<Cfg<SAKE_ENGINE_F16X2>>() : wimg_bytes<Cfg<SAKE_ENGINE_TF32X3>>();
  return 2 * w + 1024;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int ENGINE>
static int tc_fwd_impl(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                       void* scratch, cudaStream_t st) {
  using CF = Cfg<ENGINE>;
  uint8_t* w1 = (uint8_t*)scratch;
  uint8_t* w2 = w1 + wimg_bytes<CF>();
  TileGeom g = make_geom(d);
  float* wsc = reinterpret_cast<float*>(w2 + wimg_bytes<CF>());       // {scale, 1/scale} of the weight images
  if (!d.prepared) { const int rc = tc_mix_prepare(p, scratch, ENGINE, st); if (rc) return rc; }
  if (g.nseg > 1) SAKE_CUDA_CHECK(cudaMemsetAsync(sv.ssum, 0, sizeof(float) * rows_pad128(d.R) * d.C * 3, st));
  static unsigned long long optin = 0;
  { const int rc = smem_optin(k_tc_mix_fwd<ENGINE>, smem_bytes<CF>(), optin); if (rc) return rc; }
  const int grid = g.num_tiles < num_sms() ? g.num_tiles : num_sms();
  {
    ProfScope prof(1, d.P, st);
    static int dbg = -1;
    if (dbg < 0) { const char* s = getenv("SAKE_DEBUG_WSPLITS"); dbg = s ? atoi(s) : 0; }
    k_tc_mix_fwd<ENGINE><<<grid, NTHREADS, smem_bytes<CF>(), st>>>(g, x, mask, sv.e, sv.att, w1, wsc, sv.ssum, sv.ssum_tt, dbg);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int ENGINE>
static int tc_prepare_impl(const SakeLayerParams& p, void* wmix, cudaStream_t st) {
  using CF = Cfg<ENGINE>;
  uint8_t* w1 = (uint8_t*)wmix;
  uint8_t* w2 = w1 + wimg_bytes<CF>();
  float* wsc = reinterpret_cast<float*>(w2 + wimg_bytes<CF>());
  const int prep_threads = 2 * CF::NCHUNK * CC * 8;
  if ((reinterpret_cast<uintptr_t>(p.x_mixing_kernel) & 15) != 0) { set_error("x_mixing kernel must be 16-byte aligned"); return SAKE_EINVAL; }
  k_tc_prep<ENGINE><<<(prep_threads + 255) / 256, 256, 0, st>>>(p.x_mixing_kernel, w1, w2, wsc);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}
int tc_mix_prepare(const SakeLayerParams& p, void* wmix, int engine, cudaStream_t st) {
  if (engine == SAKE_ENGINE_BF16) return tc_prepare_impl<SAKE_ENGINE_BF16>(p, wmix, st);
  if (engine == SAKE_ENGINE_F16X2) return tc_prepare_impl<SAKE_ENGINE_F16X2>(p, wmix, st);
  return tc_prepare_impl<SAKE_ENGINE_TF32X3>(p, wmix, st);
}

int gen_mix_dw_from_gz(const Dims& d, const Saved& sv, const float* gZ, float* gWx, cudaStream_t st);

template <int ENGINE>
static int tc_bwd_impl(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
                       const BwdScratch& sc, float* gWx, void* scratch, XtgList& L, cudaStream_t st) {
  using CF = Cfg<ENGINE>;
  uint8_t* w1 = (uint8_t*)scratch;
  uint8_t* w2 = w1 + wimg_bytes<CF>();
  TileGeom g = make_geom(d);
  float* wsc = reinterpret_cast<float*>(w2 + wimg_bytes<CF>());       // {scale, 1/scale} of the weight images
  (void)p;                                                            // the images were built by the forward call
  static unsigned long long optin = 0;
  { const int rc = smem_optin(k_tc_mix_bwd<ENGINE>, smem_bytes<CF>(), optin); if (rc) return rc; }
  const int grid = g.num_tiles < num_sms() ? g.num_tiles : num_sms();
  {
    ProfScope prof(2, d.P, st);
    static int dbg = -1;
    if (dbg < 0) { const char* s = getenv("SAKE_DEBUG_WSPLITS"); dbg = s ? atoi(s) : 0; }
    k_tc_mix_bwd<ENGINE><<<grid, BWD_THREADS, smem_bytes<CF>(), st>>>(g, x, mask, sv.e, sv.att, w1, w2, wsc,
                                                                  reinterpret_cast<const float4*>(sc.T), sc.tmax, sc.ghe,
                                                                  sc.ge, sc.gatt, sc.gdir, gWx ? sc.gZ : nullptr, dbg);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  if (gWx) {
    // dWx[c][c'] += sum_pairs E[pair][c] * dZ[pair][c']  (layers.py:95) on the tensor cores
    XtgArgs a;
    memset(&a, 0, sizeof(a));
    a.e = sv.e; a.att = sv.att; a.xw = CC; a.ones_col = -1;
    a.G = sc.gZ; a.ldg = CC; a.g_tt = CC / 4; a.gw = CC;
    a.MXpad = CC; a.NG = CC; a.P = d.P; a.Pdev = d.hdr ? &d.hdr->P : nullptr;
    a.out = gWx; a.ldo = CC; a.out_rows = CC; a.out_cols = CC;
    if (L.push(a)) { set_error("xtg list full"); return SAKE_EINVAL; }
  }
  return 0;
}

int tc_mix_fwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
               void* tc_scratch, int engine, cudaStream_t st) {
  if (engine == SAKE_ENGINE_BF16) return tc_fwd_impl<SAKE_ENGINE_BF16>(d, p, x, mask, sv, tc_scratch, st);
  if (engine == SAKE_ENGINE_F16X2) return tc_fwd_impl<SAKE_ENGINE_F16X2>(d, p, x, mask, sv, tc_scratch, st);
  return tc_fwd_impl<SAKE_ENGINE_TF32X3>(d, p, x, mask, sv, tc_scratch, st);
}

int tc_mix_bwd(const Dims& d, const SakeLayerParams& p, const float* x, const float* mask, const Saved& sv,
               const BwdScratch& sc, float* gWx, void* tc_scratch, int engine, XtgList& L, cudaStream_t st) {
  if (engine == SAKE_ENGINE_BF16) return tc_bwd_impl<SAKE_ENGINE_BF16>(d, p, x, mask, sv, sc, gWx, tc_scratch, L, st);
  if (engine == SAKE_ENGINE_F16X2) return tc_bwd_impl<SAKE_ENGINE_F16X2>(d, p, x, mask, sv, sc, gWx, tc_scratch, L, st);
  return tc_bwd_impl<SAKE_ENGINE_TF32X3>(d, p, x, mask, sv, sc, gWx, tc_scratch, L, st);
}

// ---- self-test ------------------------------------------------------------------------------------
template <int ENGINE>
static int selftest_one(float* max_err, cudaStream_t st) {
  using CF = Cfg<ENGINE>;
  const int nchunk = 2, Kt = nchunk * CF::KCH;
  std::vector<float> A((size_t)TILE * Kt), B((size_t)CC * Kt), D((size_t)TILE * CC);
  uint32_t seed = 12345u;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  float *dA, *dB, *dD;
  SAKE_CUDA_CHECK(cudaMalloc(&dA, A.size() * 4));
  SAKE_CUDA_CHECK(cudaMalloc(&dB, B.size() * 4));
  SAKE_CUDA_CHECK(cudaMalloc(&dD, D.size() * 4));
  SAKE_CUDA_CHECK(cudaMemcpyAsync(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice, st));
  SAKE_CUDA_CHECK(cudaMemcpyAsync(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice, st));
  const size_t smem = (size_t)nchunk * CF::NSPLIT * (P_IMG + W_IMG) + 1024 + 64;
  SAKE_CUDA_CHECK(cudaFuncSetAttribute(k_tc_selftest<ENGINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_selftest<ENGINE><<<1, 128, smem, st>>>(dA, dB, dD, nchunk);
  SAKE_CUDA_CHECK(cudaGetLastError());
  SAKE_CUDA_CHECK(cudaMemcpyAsync(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost, st));
  SAKE_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  double worst = 0.0;
  for (int r = 0; r < TILE; ++r)
    for (int c = 0; c < CC; ++c) {
      double ref = 0.0;
      for (int k = 0; k < Kt; ++k) ref += (double)A[(size_t)r * Kt + k] * (double)B[(size_t)c * Kt + k];
      double err = fabs(ref - (double)D[(size_t)r * CC + c]);
      if (err > worst) worst = err;
    }
  *max_err = (float)worst;
  return 0;
}

int tc_selftest(float* max_abs_err, cudaStream_t st) {
  float e1 = 0.f, e2 = 0.f, e3 = 0.f;
  int rc = selftest_one<SAKE_ENGINE_TF32X3>(&e1, st);
  if (rc) return rc;
  rc = selftest_one<SAKE_ENGINE_BF16>(&e2, st);
  if (rc) return rc;
  rc = selftest_one<SAKE_ENGINE_F16X2>(&e3, st);
  if (rc) return rc;
  if (max_abs_err) { max_abs_err[0] = e1; max_abs_err[1] = e2; }
  if (!(e3 < 1e-5f)) { set_error("tcgen05 selftest: f16x2 max abs err %g", e3); return SAKE_ECUDA; }
  if (!(e1 < 1e-5f)) { set_error("tcgen05 selftest: tf32x3 max abs err %g", e1); return SAKE_ECUDA; }
  if (!(e2 < 5e-2f)) { set_error("tcgen05 selftest: bf16 max abs err %g", e2); return SAKE_ECUDA; }
  return 0;
}

}  // namespace sake
