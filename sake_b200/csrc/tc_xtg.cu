// tcgen05 "X^T G" kernel: D[MX x NG] += sum_p X[p][:]^T G[p][:]  with p running over atom pairs (or
// nodes).  Every weight-gradient contraction of the layer has this shape (K = pairs):
//   dWx = E^T dZ (sake/layers.py:95), dW2 = a1^T g_e, dW1[2H:] = g^T g_z1 (layers.py:20-26), dWs = e^T g_q, ...
// Operands are MN-major 128B-swizzled images: rows = pairs (the K dimension), 128 contiguous bytes of
// features per row and MN block — exactly what a thread-per-pair builder writes with 16-byte stores.
// fp32 inputs are converted on the fly (exact 3-way bf16 split, 6 MMAs; or plain bf16).  The accumulator
// lives in TMEM for the whole CTA and is flushed once with atomics.
#include <cuda_bf16.h>
#include <stdint.h>
#include <type_traits>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace sake {
using namespace tc;

// Operand precision.  kind::tf32 has no MN-major (transposing) operand path on sm_100a (measured: the
// MMA is a silent no-op), and K-major images with K = pairs need a transposing builder that is L1-bound.
// kind::f16 does support MN-major, so the fp32-parity engine splits every operand into two bf16 terms by
// round-to-nearest (x = hi + mid + O(2^-17 |x|)) and issues 3 MMAs (hh, hm, mh; the mid*mid product is below the
// split's own rounding error).  Unbiased rounding errors average out over K = pairs: CPU experiment
// scripts/xtg_split_error.py 4.4e-6 rms; measured on B200 against the fp64 oracle: worst parameter-gradient error
// 7.5e-5 of max|g| per tensor at the multi-tile test size (tests/test_gpu_round2.py), energies and forces never pass
// through this kernel.  make XTG_TRUNC4=1 builds the round-1 scheme (split by truncation, 4 products: 8.7e-5, 10 %
// slower).  The bf16 engine uses one bf16 image and one MMA.
#ifndef SAKE_XTG_TRUNC4
#define SAKE_XTG_RN3 1
#endif
template <int ENGINE> struct XCfg;
#ifdef SAKE_XTG_RN3
template <> struct XCfg<SAKE_ENGINE_TF32X3> { static constexpr int NSPLIT = 2, NPROD = 3; };
#else
template <> struct XCfg<SAKE_ENGINE_TF32X3> { static constexpr int NSPLIT = 2, NPROD = 4; };
#endif
template <> struct XCfg<SAKE_ENGINE_BF16> { static constexpr int NSPLIT = 1, NPROD = 1; };
constexpr int XEPU = 8;      // features per 16-byte unit (bf16)
constexpr int XBLK = 64;     // features per 128-byte MN block
constexpr int XKP = 32;      // pairs per stage
constexpr int XKSTEP = 16;   // pairs per MMA (K of kind::f16)
__device__ __constant__ int x_prod_x[4] = {0, 0, 1, 1};
__device__ __constant__ int x_prod_g[4] = {0, 1, 0, 1};

constexpr int XTG_THREADS = 288;   // warp 0: MMA issuer / TMEM owner; warps 1-8: builders + epilogue
constexpr int XTG_NSTAGE = 3;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// upper halves of two fp32 words -> one bf16x2 word (truncation), one PRMT
__device__ __forceinline__ uint32_t pack_hi16(float a, float b) {
  return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
}
__device__ __forceinline__ float trunc_bf16(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFF0000u); }
// 8 consecutive features of one pair -> one 16-byte unit in each split image.
// 2-way split by truncation: x = hi + mid up to 2^-16 |x| (each term keeps 8 mantissa bits); one PRMT per
// packed pair plus one LOP + FSUB per element instead of round-to-nearest conversions.
template <class CF>
__device__ __forceinline__ void xtg_store_unit(uint8_t* img, size_t split_stride, uint32_t off, const float* v) {
  if constexpr (CF::NSPLIT == 1) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  } else {
    uint32_t pk[4];
    float r[8];
#ifdef SAKE_XTG_RN3
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[2 * i] = v[2 * i] - __uint_as_float(pk[i] << 16);
      r[2 * i + 1] = v[2 * i + 1] - __uint_as_float(pk[i] & 0xFFFF0000u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack_bf16(r[2 * i], r[2 * i + 1]);
#else
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack_hi16(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i] - trunc_bf16(v[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[i] = pack_hi16(r[2 * i], r[2 * i + 1]);
#endif
    *reinterpret_cast<uint4*>(img + split_stride + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// load 8 consecutive floats of a row-major fp32 source (zero beyond `width`)
__device__ __forceinline__ void load8(const float* __restrict__ src, int c0, int width, bool vec_ok, float* v) {
  if (vec_ok && c0 + 8 <= width) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + c0 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (c0 + i < width) ? __ldg(src + c0 + i) : 0.f;
  }
}

struct XtgBatch {
  XtgArgs a[XtgList::MAXP];
  int colbase[XtgList::MAXP + 1];   // reduction grid: first block of every problem (one block per output column)
  int nprob;
};

// K extent, slice per CTA and number of working CTAs of one problem.  Uniform batches: the host's values.
// Ragged batches: the extent lives in device memory (a.Pdev), the host sized the grid and the partial-sum
// buffer for a.gx CTAs, and the split is recomputed here so that the REAL pairs are spread over all of them.
struct XtgSplit { long long P, per_cta; int gx; };
__device__ __forceinline__ XtgSplit xtg_split(const XtgArgs& a) {
  XtgSplit s;
  s.P = a.P; s.per_cta = a.pairs_per_cta; s.gx = a.gx;
  if (a.Pdev != nullptr) {
    s.P = *a.Pdev;
    const long long stages = (s.P + 31) / 32;             // XKP pairs per stage
    long long per = (stages + a.gx - 1) / a.gx;
    if (per < 4) per = 4;
    s.per_cta = per * 32;
    s.gx = (int)((s.P + s.per_cta - 1) / s.per_cta);
  }
  return s;
}
__device__ __align__(16) float g_xtg_zeros[2304];  // a source row of zeros (pair rows past the end of a contraction), long
                                                   // enough for the unit strides of a G8 source

// TCOLS: TMEM columns of the CTA (512: the 256 x 256 x_mixing gradient; 256: everything else, so that two CTAs of
//        the small contractions share an SM); nstage: operand ring depth (<= XTG_NSTAGE).
// MINB:  CTAs per SM the register budget is cut for (2 for the lean 256-column kernel: 96 registers).
// LEAN:  every problem of the batch has one of the layer's regular shapes (xtg_is_lean).  The builder then maps
//        thread -> (pair row, 16-byte unit position) once, so that a stage costs one pointer bump per operand plus
//        two LDG.128, the bf16 split and two STS.128 per unit, and nothing else.  The generic builder (any width
//        and alignment) executes ~2.5x the instructions per stage, and instruction issue is what bounds this kernel
//        (DESIGN.md section 5).
template <int ENGINE, int TCOLS, bool LEAN, int MINB>
__global__ void __launch_bounds__(XTG_THREADS, MINB) k_tc_xtg(const __grid_constant__ XtgBatch batch, int nstage) {
  using CF = XCfg<ENGINE>;
  const XtgArgs& a = batch.a[blockIdx.y];
  const XtgSplit split = xtg_split(a);
  if ((int)blockIdx.x >= split.gx) return;                    // CTA beyond this problem's range
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  const int xblocks = a.MXpad / XBLK;
  const int gblocks = (a.NG + XBLK - 1) / XBLK;
  constexpr uint32_t LBO = XKP * 128;                         // bytes between 64-feature MN blocks
  const size_t ximg = (size_t)xblocks * LBO, gimg = (size_t)gblocks * LBO;
  const size_t stage = CF::NSPLIT * (ximg + gimg);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + nstage * stage);
  uint64_t* empty = full + XTG_NSTAGE;
  uint64_t* done = empty + XTG_NSTAGE;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(full + s, 256); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TCOLS>(tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  const long long p_beg = (long long)blockIdx.x * split.per_cta;
  const long long p_end = min(split.P, p_beg + split.per_cta);
  const int nst = p_end > p_beg ? (int)((p_end - p_beg + XKP - 1) / XKP) : 0;
  const int MH = a.MXpad / 128;

  if (warp == 0) {
    if (lane == 0 && nst > 0) {
      const uint32_t idesc = umma_idesc(1 /*bf16*/, 128, a.NG, 1, 1);      // both operands MN-major
      for (int it = 0; it < nst; ++it) {
        const int s = it % nstage, n = it / nstage;
        mbar_wait(full + s, n & 1);
        tc_fence_after();
        const uint32_t xb = smem_u32(base + s * stage), gb = xb + (uint32_t)(CF::NSPLIT * ximg);
        for (int mh = 0; mh < MH; ++mh)
#pragma unroll
          for (int pr = 0; pr < CF::NPROD; ++pr)
#pragma unroll
            for (int ks = 0; ks < XKP / XKSTEP; ++ks) {
              const uint32_t aaddr = xb + x_prod_x[pr] * (uint32_t)ximg + mh * (128 / XBLK) * LBO + ks * (XKSTEP * 128);
              const uint32_t baddr = gb + x_prod_g[pr] * (uint32_t)gimg + ks * (XKSTEP * 128);
              umma<false>(tmem_base + mh * a.NG, umma_desc_mn_sw128(aaddr, LBO, 1024), umma_desc_mn_sw128(baddr, LBO, 1024),
                          idesc, (it | pr | ks) != 0);
            }
        umma_commit(empty + s);
      }
      umma_commit(done);
    }
  } else if constexpr (LEAN) {
    // ------------------------------------------------------------ lean builders: thread = (pair row r, unit position uu)
    // of every 64-feature block.  Block j of X is either all data (j < nxf), the block whose first feature is the
    // ones column (j == nxf when ones_col >= 0), or zero padding up to MXpad — written once, below.
    // The 512-column instantiation is the x_mixing gradient (X = e (x) att, 4 + 4 blocks), the 256-column one the rest.
    constexpr bool EMODE = TCOLS == 512;
    const int bt = threadIdx.x - 32;                 // 0..255
    // lanes: 8 rows x 4 unit positions, so that a warp's load covers whole lines of a G8 source (8 rows of one unit =
    // 128 bytes) and, for row-major sources, 4 x 32 contiguous bytes of each of 8 rows — the same lines either way
    const int r = (bt & 7) | ((bt >> 6) << 3), uu = (bt >> 3) & 7;
    const uint32_t uoff = sw128_offset((uint32_t)r, (uint32_t)uu);
    const int nxf = EMODE ? 4 : (a.xw >> 6);
    const bool has_ones = !EMODE && a.ones_col >= 0;
    const int ngf = EMODE ? 4 : (a.gw >> 6);
    const bool gnarrow = !EMODE && (a.gw & 63) != 0; // gw < 64: one block, scalar loads
    {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (int s = 0; s < nstage; ++s)
        for (int sp = 0; sp < CF::NSPLIT; ++sp) {
          uint8_t* xs = base + s * stage + sp * ximg + uoff;
          for (int j = nxf; j < xblocks; ++j) *reinterpret_cast<uint4*>(xs + j * LBO) = z;
          uint8_t* gs = base + s * stage + CF::NSPLIT * ximg + sp * gimg + uoff;
          for (int j = ngf + (gnarrow ? 1 : 0); j < gblocks; ++j) *reinterpret_cast<uint4*>(gs + j * LBO) = z;
        }
    }
    // one stage of raw operands in registers.  EMODE: R.xa[j].xy = the two e features of unit j; a narrow G block
    // (gw < 64, no full block) lives in R.ga[0], R.gb[0]
    struct StageRegs { float4 xa[4], xb[4], ga[4], gb[4], at; bool ones_on; };
    const bool ones_here = has_ones && uu == 0;      // ones unit: bf16(1.0) in element 0 (valid rows only)
    // One definition path for the loop-carried registers (no per-row branches, which cost a register shuffle at
    // every merge): rows past the end (last stage of the grid only) read a page of zeros instead.
    // this thread's source rows for the next load; every stage bumps them by XKP rows
    // G8 sources (x_tt / g_tt = 16-byte units per row; common.cuh): unit u of row p sits at float4 index
    // g8_row(p, U) + 8 u.  The two float4 of a thread are then 8 float4 apart and consecutive 64-feature blocks 128;
    // a stage step (32 rows = 4 groups) is the same number of floats as with row-major rows.  Eight consecutive rows
    // of one unit are one line, so a warp touches as many lines per load as with row-major rows.
    const bool xtt = !EMODE && a.x_tt > 0, gtt = a.g_tt > 0;
    const long long pl = p_beg + r;
    int left = (int)min((long long)(1 << 30), p_end - pl);          // rows of this thread's slice still to load (<= 0: none)
    const float* xrow = EMODE ? a.e + (g8_row(pl, 16) + (uu >> 1) * G8S) * 4 + 2 * (uu & 1)   // e features 16 j + 2 uu, +1 (G8)
                        : xtt ? a.X + (g8_row(pl, a.x_tt) + 2 * uu * G8S) * 4
                              : a.X + pl * a.ldx + 8 * uu;
    const float* grow = gtt ? a.G + (g8_row(pl, a.g_tt) + 2 * uu * G8S) * 4 : a.G + pl * a.ldg + 8 * uu;
    const float* arow = EMODE ? a.att + pl * 4 : nullptr;
    const int xstep = EMODE ? XKP * 64 : xtt ? XKP * a.x_tt * 4 : XKP * a.ldx;      // floats per stage
    const int gstep = gtt ? XKP * a.g_tt * 4 : XKP * a.ldg;
    const int xs1 = xtt ? G8S : 1, xs16 = xtt ? 16 * G8S : 16;     // float4 strides: unit pair / 64-feature block
    const int gs1 = gtt ? G8S : 1, gs16 = gtt ? 16 * G8S : 16;
    auto load_stage = [&](StageRegs& R) {
      const bool rv = left > 0;
      R.ones_on = rv && ones_here;
      const float* zp = g_xtg_zeros + 2 * uu * G8S * 4;             // page of zeros: either set of strides stays inside
      const float* gp = rv ? grow : zp;
      if constexpr (EMODE) {                         // raw operands of E = e (x) att; the product is formed at store time
        R.at = __ldg(reinterpret_cast<const float4*>(rv ? arow : g_xtg_zeros));
        const float2* ep = reinterpret_cast<const float2*>(rv ? xrow : g_xtg_zeros + 2 * uu);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 ef = __ldg(ep + 4 * G8S * 2 * j); R.xa[j].x = ef.x; R.xa[j].y = ef.y; }   // 4 units on
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          R.ga[j] = __ldg(reinterpret_cast<const float4*>(gp) + gs16 * j);
          R.gb[j] = __ldg(reinterpret_cast<const float4*>(gp) + gs16 * j + gs1);
        }
        arow += XKP * 4;
      } else {
        const float4* xp = reinterpret_cast<const float4*>(rv ? xrow : zp);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nxf) { R.xa[j] = __ldg(xp + xs16 * j); R.xb[j] = __ldg(xp + xs16 * j + xs1); }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < ngf) {
            R.ga[j] = __ldg(reinterpret_cast<const float4*>(gp) + gs16 * j);
            R.gb[j] = __ldg(reinterpret_cast<const float4*>(gp) + gs16 * j + gs1);
          }
        if (gnarrow) {
          float gn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) gn[i] = (8 * uu + i < a.gw) ? __ldg(gp + (i >> 2) * 4 * gs1 + (i & 3)) : 0.f;
          R.ga[0] = make_float4(gn[0], gn[1], gn[2], gn[3]);
          R.gb[0] = make_float4(gn[4], gn[5], gn[6], gn[7]);
        }
      }
      left -= XKP; xrow += xstep; grow += gstep;
    };
    auto build_stage = [&](const StageRegs& R, int it) {
      const int s = it % nstage, n = it / nstage;
      mbar_wait_warp(empty + s, (n & 1) ^ 1);
      uint8_t* xs = base + s * stage + uoff;
      uint8_t* gs = xs + CF::NSPLIT * ximg;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nxf) {
          if constexpr (EMODE) {                     // X = e (x) att, feature c = f*4 + head  (layers.py:206-207)
            const float vals[8] = {R.xa[j].x * R.at.x, R.xa[j].x * R.at.y, R.xa[j].x * R.at.z, R.xa[j].x * R.at.w,
                                   R.xa[j].y * R.at.x, R.xa[j].y * R.at.y, R.xa[j].y * R.at.z, R.xa[j].y * R.at.w};
            xtg_store_unit<CF>(xs + j * LBO, ximg, 0u, vals);
          } else {
            const float vals[8] = {R.xa[j].x, R.xa[j].y, R.xa[j].z, R.xa[j].w, R.xb[j].x, R.xb[j].y, R.xb[j].z, R.xb[j].w};
            xtg_store_unit<CF>(xs + j * LBO, ximg, 0u, vals);
          }
        }
      if (has_ones) *reinterpret_cast<uint4*>(xs + nxf * LBO) = make_uint4(R.ones_on ? 0x00003F80u : 0u, 0u, 0u, 0u);   // (the residual image stays zero)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < ngf || (j == 0 && gnarrow)) {
          const float vals[8] = {R.ga[j].x, R.ga[j].y, R.ga[j].z, R.ga[j].w, R.gb[j].x, R.gb[j].y, R.gb[j].z, R.gb[j].w};
          xtg_store_unit<CF>(gs + j * LBO, gimg, 0u, vals);
        }
      fence_proxy_async();
      mbar_arrive(full + s);
    };
    if constexpr (EMODE) {
      // the 256 x 256 contraction runs one CTA per SM with registers to spare: two stages of operands are kept in
      // flight (set B is requested before set A is consumed and vice versa), so a load has a whole build + hand-off
      // of lead time instead of only the wait for a free ring slot — the kernel was bound by that latency
      StageRegs A, B;
      if (nst > 0) load_stage(A);
      for (int it = 0; it < nst; it += 2) {
        if (it + 1 < nst) load_stage(B);
        build_stage(A, it);
        if (it + 2 < nst) load_stage(A);
        if (it + 1 < nst) build_stage(B, it + 1);
      }
    } else {
      StageRegs A;
      if (nst > 0) load_stage(A);
      for (int it = 0; it < nst; ++it) {
        build_stage(A, it);
        if (it + 1 < nst) load_stage(A);
      }
    }
  } else {
    // ------------------------------------------------------------ builders (thread = pair x 8-feature unit)
    const int bt = threadIdx.x - 32;                 // 0..255
    const int xu = a.MXpad / XEPU;                   // 16-byte units per pair row, X side
    const int gu = gblocks * (XBLK / XEPU);          // G side
    const bool xvec = a.X != nullptr && (a.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.X) & 15) == 0);
    const bool gvec = (a.ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.G) & 15) == 0);
    // Per-thread work items are loop invariant: up to 4 X units and 4 G units (pair row r, features c0..c0+7).
    // Everything that needs an integer division is computed once; per stage only pointers advance.
    int xr[4], xc[4], gr[4], gc[4];
    uint32_t xo[4], go[4];
    bool xok[4], gok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = bt + 256 * k;
      xok[k] = idx < XKP * xu;
      gok[k] = idx < XKP * gu;
      xr[k] = xok[k] ? idx / xu : 0;
      xc[k] = xok[k] ? (idx - xr[k] * xu) * XEPU : 0;
      gr[k] = gok[k] ? idx / gu : 0;
      gc[k] = gok[k] ? (idx - gr[k] * gu) * XEPU : 0;
      xo[k] = (uint32_t)(xc[k] / XBLK) * LBO + sw128_offset((uint32_t)xr[k], (uint32_t)((xc[k] % XBLK) / XEPU));
      go[k] = (uint32_t)(gc[k] / XBLK) * LBO + sw128_offset((uint32_t)gr[k], (uint32_t)((gc[k] % XBLK) / XEPU));
    }
    const bool emode = a.e != nullptr;
    // Software pipeline: the global loads of stage it+1 are issued right after the stores of stage it, so
    // their latency overlaps the wait for the next free slot instead of sitting in front of every store.
    float xv[4][8], gv[4][8];
    auto load_stage = [&](int it) {
      const long long p0 = p_beg + (long long)it * XKP;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { xv[k][i] = 0.f; gv[k][i] = 0.f; }
        if (xok[k] && p0 + xr[k] < p_end) {
          const long long p = p0 + xr[k];
          if (emode) {                      // raw operands of E = e (x) att; the product is formed at store time
            const float4 at = __ldg(reinterpret_cast<const float4*>(a.att + p * 4));
            const float2 ef = __ldg(reinterpret_cast<const float2*>(a.e + p * 64 + (xc[k] >> 2)));
            xv[k][0] = at.x; xv[k][1] = at.y; xv[k][2] = at.z; xv[k][3] = at.w; xv[k][4] = ef.x; xv[k][5] = ef.y;
          } else {
            load8(a.X + p * a.ldx, xc[k], a.xw, xvec, xv[k]);
          }
        }
        if (gok[k] && p0 + gr[k] < p_end) load8(a.G + (p0 + gr[k]) * a.ldg, gc[k], a.gw, gvec, gv[k]);
      }
    };
    if (nst > 0) load_stage(0);
    for (int it = 0; it < nst; ++it) {
      const int s = it % nstage, n = it / nstage;
      mbar_wait_warp(empty + s, (n & 1) ^ 1);
      uint8_t* ximgp = base + s * stage;
      uint8_t* gimgp = ximgp + CF::NSPLIT * ximg;
      const long long p0 = p_beg + (long long)it * XKP;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (xok[k]) {
          float vals[8];
          if (emode) {                      // X = e (x) att, feature c = f*4 + head  (layers.py:206-207)
            vals[0] = xv[k][4] * xv[k][0]; vals[1] = xv[k][4] * xv[k][1]; vals[2] = xv[k][4] * xv[k][2]; vals[3] = xv[k][4] * xv[k][3];
            vals[4] = xv[k][5] * xv[k][0]; vals[5] = xv[k][5] * xv[k][1]; vals[6] = xv[k][5] * xv[k][2]; vals[7] = xv[k][5] * xv[k][3];
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) vals[i] = xv[k][i];
          }
          if (a.ones_col >= xc[k] && a.ones_col < xc[k] + 8 && p0 + xr[k] < p_end) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (xc[k] + i == a.ones_col) vals[i] = 1.0f;
          }
          xtg_store_unit<CF>(ximgp, ximg, xo[k], vals);
        }
        if (gok[k]) xtg_store_unit<CF>(gimgp, gimg, go[k], gv[k]);
      }
      fence_proxy_async();
      mbar_arrive(full + s);
      if (it + 1 < nst) load_stage(it + 1);
    }
  }
  if (warp != 0) {
    // ------------------------------------------------------------ epilogue: flush the accumulator
    if (nst > 0) {
      mbar_wait_warp(done, 0);
      tc_fence_after();
      const int q = warp & 3, half = (warp - 1) >> 2;      // two warps per lane quarter
      const int nchunks = (a.NG + 31) / 32;
      for (int mh = 0; mh < MH; ++mh) {
        const int row = mh * 128 + q * 32 + lane;
        for (int cc = half; cc < nchunks; cc += 2) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mh * a.NG + cc * 32, v);
          tmem_ld_wait();
          if (a.partial != nullptr) {
            // per-CTA partial, stored transposed [col][row] so that a warp writes 128 contiguous bytes
            float* pp = a.partial + (size_t)blockIdx.x * a.MXpad * a.NG + row;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int col = cc * 32 + k;
              if (col < a.NG) pp[(size_t)col * a.MXpad] = v[k];
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int col = cc * 32 + k;
              if (col >= a.NG) continue;
              if (row < a.out_rows) {
                if (col < a.out_cols) atomicAdd(a.out + (size_t)row * a.ldo + col, v[k]);
              } else if (row < a.out_rows + a.extra_rows) {
                if (col < a.extra_ld) atomicAdd(a.extra + (size_t)(row - a.out_rows) * a.extra_ld + col, v[k]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TCOLS>(tmem_base);
}

// out[row][col] += sum_cta partial[cta][col][row]   (deterministic second stage of the flush)
// Thread = four consecutive rows of one column (one float4 per partial: a warp reads 512 contiguous bytes), a block =
// 128 / (MXpad / 4) columns x XRED_SLICES slices of the CTA range; each thread keeps 8 independent loads in flight (the
// partials are L2-resident; one load per thread at a time left this kernel latency-bound at ~1 TB/s), the slices are
// combined in a fixed order through shared memory.
constexpr int XRED_SLICES = 4;
__host__ __device__ inline int xred_cols_per_block(int MXpad) { return 128 / (MXpad / 4); }   // 4 (MXpad 128) or 2 (256)
__global__ void __launch_bounds__(128 * XRED_SLICES) k_xtg_reduce(const __grid_constant__ XtgBatch batch) {
  __shared__ float4 red[XRED_SLICES][128];
  int pi = 0;
  while (pi + 1 < batch.nprob && (int)blockIdx.x >= batch.colbase[pi + 1]) ++pi;   // <= 20 problems
  const XtgArgs& a = batch.a[pi];
  if (a.partial == nullptr) return;                          // block-uniform
  const int tx = threadIdx.x & 127, sl = threadIdx.x >> 7;
  const int rq = a.MXpad >> 2;                               // threads per column
  const int col = ((int)blockIdx.x - batch.colbase[pi]) * (128 / rq) + tx / rq;
  const int r4 = tx % rq, row0 = 4 * r4;
  const int ncta = xtg_split(a).gx;
  const int per = (ncta + XRED_SLICES - 1) / XRED_SLICES;
  const int c0 = min(ncta, sl * per), c1 = min(ncta, c0 + per);
  const size_t cstride = ((size_t)a.MXpad * a.NG) >> 2;      // float4 per partial
  const int rows = a.out_rows + a.extra_rows;
  const bool on = col < a.NG && row0 < rows;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) {
    const float4* pp = reinterpret_cast<const float4*>(a.partial) + (((size_t)col * a.MXpad) >> 2) + r4 + (size_t)c0 * cstride;
    int c = c0;
    float4 t0 = s, t1 = s;
    for (; c + 8 <= c1; c += 8, pp += 8 * cstride) {
      const float4 v0 = pp[0], v1 = pp[cstride], v2 = pp[2 * cstride], v3 = pp[3 * cstride];
      const float4 v4 = pp[4 * cstride], v5 = pp[5 * cstride], v6 = pp[6 * cstride], v7 = pp[7 * cstride];
      t0.x += v0.x; t0.y += v0.y; t0.z += v0.z; t0.w += v0.w;   t1.x += v1.x; t1.y += v1.y; t1.z += v1.z; t1.w += v1.w;
      t0.x += v2.x; t0.y += v2.y; t0.z += v2.z; t0.w += v2.w;   t1.x += v3.x; t1.y += v3.y; t1.z += v3.z; t1.w += v3.w;
      t0.x += v4.x; t0.y += v4.y; t0.z += v4.z; t0.w += v4.w;   t1.x += v5.x; t1.y += v5.y; t1.z += v5.z; t1.w += v5.w;
      t0.x += v6.x; t0.y += v6.y; t0.z += v6.z; t0.w += v6.w;   t1.x += v7.x; t1.y += v7.y; t1.z += v7.z; t1.w += v7.w;
    }
    for (; c < c1; ++c, pp += cstride) { const float4 v = pp[0]; t0.x += v.x; t0.y += v.y; t0.z += v.z; t0.w += v.w; }
    s = make_float4(t0.x + t1.x, t0.y + t1.y, t0.z + t1.z, t0.w + t1.w);
  }
  red[sl][tx] = s;
  __syncthreads();
  if (sl != 0 || !on) return;
  auto total = [&](int t) {
    float4 q = red[0][t];
#pragma unroll
    for (int k = 1; k < XRED_SLICES; ++k) { const float4 w = red[k][t]; q.x += w.x; q.y += w.y; q.z += w.z; q.w += w.w; }
    return q;
  };
  const float4 tq = total(tx);
  const float tv[4] = {tq.x, tq.y, tq.z, tq.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + i;
    const float tot = tv[i];
    if (row >= rows) break;
    if (row < a.out_rows) {
      if (col < a.out_cols) a.out[(size_t)row * a.ldo + col] += tot;
    } else if (a.mb_gmu != nullptr) {
      // rows out_rows (S1) and out_rows + 1 (S2) of column 128 + k: the thread that owns S2 fetches S1 (same column,
      // possibly the neighbouring row quad) and finishes the RBF mean / width gradients (utils.py:61-65)
      const int k = col - 128;
      if (row == a.out_rows + 1 && k >= 0 && k < a.mb_K) {
        const int r1 = a.out_rows;                           // row of S1
        const float4 q1 = total(tx - r4 + (r1 >> 2));
        const float s1 = (r1 & 3) == 0 ? q1.x : (r1 & 3) == 1 ? q1.y : (r1 & 3) == 2 ? q1.z : q1.w;
        a.mb_gmu[k] += 2.0f * a.mb_beta[k] * s1;
        a.mb_gbeta[k] += -(tot - a.mb_mu[k] * s1);
      }
    } else if (a.extra2 != nullptr && col >= a.extra2_col0) {
      a.extra2[col - a.extra2_col0] += tot;
    } else if (col < a.extra_ld) {
      a.extra[(size_t)(row - a.out_rows) * a.extra_ld + col] += tot;
    }
  }
}

size_t tc_xtg_partial_bytes() { return (size_t)192 << 20; }

static int xtg_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// Launch the collected contractions as two grids (blockIdx.y = problem): the 256 x 256 x_mixing gradient, which
// needs the whole TMEM and a 192 KB operand ring (this launch is the one the profiler times as "mix_dw"), and
// all the small ones (<= 256 TMEM columns, two-stage ring: two CTAs per SM) — each followed by its reduction grid.
// the regular shapes of the layer (see the LEAN builder): 64-multiples of 16-byte-aligned rows, the ones column
// right behind the data, and G either whole 64-feature blocks or one narrow block
static bool xtg_is_lean(const XtgArgs& a, bool is_big) {
  if (is_big != (a.e != nullptr)) return false;        // the 512-column lean kernel is the E = e (x) att one
  if (is_big && (a.gw != 256 || a.NG != 256)) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a.e != nullptr) {
    if (a.att == nullptr || !al16(a.e) || !al16(a.att) || a.xw != 256 || a.MXpad != 256 || a.ones_col >= 0) return false;
  } else {
    if (a.X == nullptr || !al16(a.X) || (a.x_tt == 0 && a.ldx % 4 != 0) || a.xw <= 0 || a.xw % 64 != 0) return false;
    if (a.ones_col >= 0 && a.ones_col != a.xw) return false;
    if (a.xw + (a.ones_col >= 0 ? 1 : 0) > a.MXpad) return false;
  }
  if (a.gw <= 0 || a.gw > a.NG) return false;
  if (a.gw % 64 == 0) return al16(a.G) && (a.g_tt > 0 || a.ldg % 4 == 0);
  return a.gw < 64 && (a.g_tt == 0 || al16(a.G));
}

template <int TCOLS, bool LEAN, int MINB>
static int xtg_launch(const XtgBatch& batch, int nb, int gx_max, size_t smem, int nstage, bool bf, cudaStream_t st) {
  if (nb == 0) return 0;
  if (smem > 200 * 1024) { set_error("tc_xtg: smem %zu", smem); return SAKE_EUNSUPPORTED; }
  static unsigned long long optin_tf = 0, optin_bf = 0;
  const int orc = bf ? smem_optin(k_tc_xtg<SAKE_ENGINE_BF16, TCOLS, LEAN, MINB>, 200 * 1024, optin_bf)
                     : smem_optin(k_tc_xtg<SAKE_ENGINE_TF32X3, TCOLS, LEAN, MINB>, 200 * 1024, optin_tf);
  if (orc) return orc;
  dim3 grid(gx_max, nb);
  if (bf) k_tc_xtg<SAKE_ENGINE_BF16, TCOLS, LEAN, MINB><<<grid, XTG_THREADS, smem, st>>>(batch, nstage);
  else k_tc_xtg<SAKE_ENGINE_TF32X3, TCOLS, LEAN, MINB><<<grid, XTG_THREADS, smem, st>>>(batch, nstage);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}
// ONE reduction grid for every problem of both launches: a block per 4 (or 2) output columns, none for columns that do not exist
static int xtg_reduce_all(const XtgBatch& all, cudaStream_t st) {
  const int ncols = all.colbase[all.nprob];
  if (ncols == 0) return 0;
  k_xtg_reduce<<<ncols, 128 * XRED_SLICES, 0, st>>>(all);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tc_xtg_flush(XtgList& L, float* partial, int engine, int prof_kind, cudaStream_t st, cudaStream_t red_st, cudaEvent_t fork) {
  if (L.n == 0) return 0;
  const bool bf = engine == SAKE_ENGINE_BF16;
  const int nsplit = bf ? 1 : 2;
  const int sms = xtg_num_sms();
  XtgBatch big, small;
  memset(&big, 0, sizeof(big));
  memset(&small, 0, sizeof(small));
  size_t smem_b = 0, smem_s = 0, poff = 0;
  int gx_b = 0, gx_s = 0, ng_b = 0, ng_s = 0, nb_b = 0, nb_s = 0;
  long long prof_pairs = 0;
  bool lean_b = true, lean_s = true;
  constexpr int NST_BIG = XTG_NSTAGE, NST_SMALL = 2;
  // Small problems with a long K (the pair-level ones) are split so that TOGETHER they fill the machine exactly once at
  // two CTAs per SM: with every problem split over all SMs the long CTAs ran as one and a half waves, the second half
  // on half the SMs.  The short (node-level) problems fill the slots as the long CTAs retire.
  int n_long = 0;
  for (int i = 0; i < L.n; ++i) {
    const XtgArgs& a = L.a[i];
    if (a.P > 0 && (a.MXpad / 128) * a.NG <= 256 && (a.P + XKP - 1) / XKP >= 8LL * sms) ++n_long;
  }
  const int gx_long = n_long > 0 ? (2 * sms / n_long > 2 * sms ? 2 * sms : (2 * sms / n_long < sms / 2 ? sms / 2 : 2 * sms / n_long)) : sms;
  for (int i = 0; i < L.n; ++i) {
    XtgArgs a = L.a[i];
    if (a.P <= 0) continue;
    if (a.extra_ld == 0) a.extra_ld = a.NG;
    if (a.MXpad % 128 != 0 || a.MXpad > 256 || a.NG % 16 != 0 || a.NG > 256 || (a.MXpad / 128) * a.NG > 512) {
      set_error("tc_xtg: unsupported shape MXpad=%d NG=%d", a.MXpad, a.NG);
      return SAKE_EUNSUPPORTED;
    }
    const bool is_big = (a.MXpad / 128) * a.NG > 256;
    const size_t lbo = (size_t)XKP * 128;
    const size_t stage = nsplit * ((size_t)(a.MXpad / XBLK) * lbo + (size_t)((a.NG + XBLK - 1) / XBLK) * lbo);
    const size_t smem = (is_big ? NST_BIG : NST_SMALL) * stage + 256 + 1024;
    long long stages_total = (a.P + XKP - 1) / XKP;
    const int ctas = (!is_big && stages_total >= 8LL * sms) ? gx_long : sms;
    long long per = (stages_total + ctas - 1) / ctas;
    if (per < 4) per = 4;                                  // keep the flush amortised
    a.pairs_per_cta = per * XKP;
    a.gx = (int)((a.P + a.pairs_per_cta - 1) / a.pairs_per_cta);
    if (a.Pdev != nullptr) a.gx = ctas;                    // ragged: the kernel splits the real extent over `gx` CTAs
    const size_t need = (size_t)a.gx * a.MXpad * a.NG;
    if (partial != nullptr && (poff + need) * sizeof(float) <= tc_xtg_partial_bytes()) {
      a.partial = partial + poff;
      poff += need;
    } else {
      a.partial = nullptr;                                 // falls back to atomics
    }
    if (a.partial == nullptr && (a.extra2 != nullptr || a.mb_gmu != nullptr)) {
      set_error("tc_xtg: split / folded extra rows need the partial-sum buffer");
      return SAKE_EINVAL;
    }
    if (is_big) {
      if (smem > smem_b) smem_b = smem;
      if (a.gx > gx_b) gx_b = a.gx;
      if (a.NG > ng_b) ng_b = a.NG;
      if (a.P > prof_pairs) prof_pairs = a.P;
      lean_b = lean_b && xtg_is_lean(a, true);
      big.a[nb_b++] = a;
    } else {
      if (smem > smem_s) smem_s = smem;
      if (a.gx > gx_s) gx_s = a.gx;
      if (a.NG > ng_s) ng_s = a.NG;
      lean_s = lean_s && xtg_is_lean(a, false);
      small.a[nb_s++] = a;
    }
  }
  L.n = 0;
  static const bool no_lean = [] { const char* e = getenv("SAKE_XTG_GENERIC"); return e && atoi(e) != 0; }();   // A/B switch
  bool any_tt = false;
  for (int i = 0; i < nb_s; ++i) any_tt = any_tt || small.a[i].x_tt > 0 || small.a[i].g_tt > 0;
  for (int i = 0; i < nb_b; ++i) any_tt = any_tt || big.a[i].x_tt > 0 || big.a[i].g_tt > 0;
  if (any_tt && (no_lean || !lean_s || !lean_b)) {
    set_error("tc_xtg: G8 sources need the lean builder (SAKE_XTG_GENERIC must be off)");
    return SAKE_EUNSUPPORTED;
  }
  // the small contractions: longest K first, so that the CTAs of the pair-level problems start in the first wave and the
  // short node-level ones fill the tail
  for (int i = 1; i < nb_s; ++i)
    for (int j = i; j > 0 && small.a[j].P > small.a[j - 1].P; --j) { const XtgArgs t = small.a[j]; small.a[j] = small.a[j - 1]; small.a[j - 1] = t; }
  XtgBatch all;
  memset(&all, 0, sizeof(all));
  int ncols = 0;
  auto add_red = [&](const XtgArgs& a) {
    if (a.partial == nullptr) return;
    const int cpb = xred_cols_per_block(a.MXpad);           // reduction blocks of this problem
    all.a[all.nprob] = a; all.colbase[all.nprob] = ncols; ncols += (a.NG + cpb - 1) / cpb; ++all.nprob;
  };
  for (int i = 0; i < nb_b; ++i) add_red(big.a[i]);
  for (int i = 0; i < nb_s; ++i) add_red(small.a[i]);
  all.colbase[all.nprob] = ncols;
  (void)ng_b; (void)ng_s;
  int rc = 0;
  {
    ProfScope prof(prof_kind, prof_pairs, st);
    rc = lean_b && !no_lean ? xtg_launch<512, true, 1>(big, nb_b, gx_b, smem_b, NST_BIG, bf, st)
                            : xtg_launch<512, false, 1>(big, nb_b, gx_b, smem_b, NST_BIG, bf, st);
    if (rc == 0 && nb_s == 0) rc = xtg_reduce_all(all, st);
  }
  if (rc || nb_s == 0) return rc;
  // the lean small kernel fits two CTAs per SM (96 registers, 256 TMEM columns, <= 82 KB): 4 builder warps per scheduler
  const int pk_small = prof_kind == 3 ? 8 : 0;   // profiler kind 8: the batched small contractions of a layer + the reduction
  ProfScope prof(pk_small, prof_pairs, st);
  rc = lean_s && !no_lean ? xtg_launch<256, true, 2>(small, nb_s, gx_s, smem_s, NST_SMALL, bf, st)
                          : xtg_launch<256, false, 1>(small, nb_s, gx_s, smem_s, NST_SMALL, bf, st);
  if (rc) return rc;
  if (red_st != nullptr && red_st != st && fork != nullptr) {
    SAKE_CUDA_CHECK(cudaEventRecord(fork, st));
    SAKE_CUDA_CHECK(cudaStreamWaitEvent(red_st, fork, 0));
    return xtg_reduce_all(all, red_st);
  }
  return xtg_reduce_all(all, st);
}

int tc_xtg(const XtgArgs& a0, int engine, int prof_kind, cudaStream_t st) {
  XtgList L;
  L.push(a0);
  return tc_xtg_flush(L, a0.partial, engine, prof_kind, st);
}

}  // namespace sake
