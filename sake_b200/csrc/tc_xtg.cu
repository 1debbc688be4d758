// tcgen05 "X^T G" kernel: D[MX x NG] += sum_p X[p][:]^T G[p][:]  with p running over atom pairs (or
// nodes).  Every weight-gradient contraction of the layer has this shape (K = pairs):
//   dWx = E^T dZ (sake/layers.py:95), dW2 = a1^T g_e, dW1[2H:] = g^T g_z1 (layers.py:20-26), dWs = e^T g_q, ...
// Operands are MN-major 128B-swizzled images: rows = pairs (the K dimension), 128 contiguous bytes of
// features per row and MN block — exactly what a thread-per-pair builder writes with 16-byte stores.
// fp32 inputs are converted on the fly (tf32 hi/lo split, 3 MMAs; or bf16).  The accumulator lives in
// TMEM for the whole CTA and is flushed once with atomics.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace sake {
using namespace tc;

template <int ENGINE> struct XCfg;
template <> struct XCfg<SAKE_ENGINE_TF32X3> {
  static constexpr bool TF32 = true;
  // kind::tf32 has no MN-major (transposing) operand path on sm_100a — measured: the MMA is a silent
  // no-op — so the tf32 images are K-major: rows = features, 128 contiguous bytes = 32 pairs.
  static constexpr int NSPLIT = 2, NPROD = 3, FMT = 2, EPU = 4, BLK = 32, KP = 32, KSTEP = 8;
};
template <> struct XCfg<SAKE_ENGINE_BF16> {
  static constexpr bool TF32 = false;
  // kind::f16 supports MN-major operands: rows = pairs (K), 128 contiguous bytes = 64 features per block.
  static constexpr int NSPLIT = 1, NPROD = 1, FMT = 1, EPU = 8, BLK = 64 /*features per MN block*/, KP = 32, KSTEP = 16;
};
__device__ __constant__ int x_prod_x[3] = {0, 1, 0};
__device__ __constant__ int x_prod_g[3] = {0, 0, 1};

constexpr int XTG_THREADS = 288;   // warp 0: MMA issuer / TMEM owner; warps 1-8: builders + epilogue
constexpr int XTG_NSTAGE = 2;

template <class CF>
__device__ __forceinline__ void xtg_store_unit(uint8_t* img, size_t split_stride, uint32_t off, const float* vals) {
  if constexpr (CF::TF32) {
    float4 hi, lo;
    split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
    split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
    *reinterpret_cast<float4*>(img + off) = hi;
    *reinterpret_cast<float4*>(img + split_stride + off) = lo;
  } else {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(vals[2 * i], vals[2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(img + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

template <int ENGINE>
__global__ void __launch_bounds__(XTG_THREADS, 1) k_tc_xtg(XtgArgs a) {
  using CF = XCfg<ENGINE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncol0 = blockIdx.y * a.NG;                        // first G column of this CTA
  const int xblocks = a.MXpad / CF::BLK;
  const int gblocks = (a.NG + CF::BLK - 1) / CF::BLK;
  const uint32_t LBO = CF::KP * 128;                          // bf16: bytes between MN blocks
  const size_t ximg = CF::TF32 ? (size_t)a.MXpad * 128 : (size_t)xblocks * LBO;
  const size_t gimg = CF::TF32 ? (size_t)((a.NG + 7) / 8 * 8) * 128 : (size_t)gblocks * LBO;
  const size_t stage = CF::NSPLIT * (ximg + gimg);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + XTG_NSTAGE * stage);
  uint64_t* empty = full + XTG_NSTAGE;
  uint64_t* done = empty + XTG_NSTAGE;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < XTG_NSTAGE; ++s) { mbar_init(full + s, 256); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  const long long p_beg = (long long)blockIdx.x * a.pairs_per_cta;
  const long long p_end = min(a.P, p_beg + a.pairs_per_cta);
  const int nst = p_end > p_beg ? (int)((p_end - p_beg + CF::KP - 1) / CF::KP) : 0;
  const int MH = a.MXpad / 128;

  if (warp == 0) {
    if (lane == 0 && nst > 0) {
      const uint32_t idesc = CF::TF32 ? umma_idesc(CF::FMT, 128, a.NG, 0, 0) : umma_idesc(CF::FMT, 128, a.NG, 1, 1);
      for (int it = 0; it < nst; ++it) {
        const int s = it % XTG_NSTAGE, n = it / XTG_NSTAGE;
        mbar_wait(full + s, n & 1);
        tc_fence_after();
        const uint32_t xb = smem_u32(base + s * stage), gb = xb + (uint32_t)(CF::NSPLIT * ximg);
        for (int mh = 0; mh < MH; ++mh)
          for (int pr = 0; pr < CF::NPROD; ++pr)
#pragma unroll
            for (int ks = 0; ks < CF::KP / CF::KSTEP; ++ks) {
              if constexpr (CF::TF32) {
                const uint32_t aaddr = xb + x_prod_x[pr] * (uint32_t)ximg + mh * (128 * 128) + ks * 32;
                const uint32_t baddr = gb + x_prod_g[pr] * (uint32_t)gimg + ks * 32;
                umma<true>(tmem_base + mh * a.NG, umma_desc_k_sw128(aaddr), umma_desc_k_sw128(baddr), idesc,
                           (it | pr | ks) != 0);
              } else {
                const uint32_t aaddr = xb + x_prod_x[pr] * (uint32_t)ximg + mh * (128 / CF::BLK) * LBO + ks * (CF::KSTEP * 128);
                const uint32_t baddr = gb + x_prod_g[pr] * (uint32_t)gimg + ks * (CF::KSTEP * 128);
                umma<false>(tmem_base + mh * a.NG, umma_desc_mn_sw128(aaddr, LBO, 1024),
                            umma_desc_mn_sw128(baddr, LBO, 1024), idesc, (it | pr | ks) != 0);
              }
            }
        umma_commit(empty + s);
      }
      umma_commit(done);
    }
  } else {
    // ------------------------------------------------------------ builders
    const int bt = threadIdx.x - 32;                 // 0..255
    const int xu = a.MXpad / CF::EPU;                // 16-byte units per pair row, X side
    const int gu = gblocks * (CF::BLK / CF::EPU);    // G side
    for (int it = 0; it < nst; ++it) {
      const int s = it % XTG_NSTAGE, n = it / XTG_NSTAGE;
      mbar_wait(empty + s, (n & 1) ^ 1);
      uint8_t* ximgp = base + s * stage;
      uint8_t* gimgp = ximgp + CF::NSPLIT * ximg;
      const long long p0 = p_beg + (long long)it * CF::KP;
      if constexpr (CF::TF32) {
        // K-major: lane = pair of the 32-pair chunk, each warp walks features; 4-byte conflict-free stores
        const int bw = bt >> 5, r = bt & 31;
        const long long p = p0 + r;
        const bool ok = p < p_end;
        const uint32_t kof = (uint32_t)(r & 3) * 4;
        float4 at = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && a.e != nullptr) at = __ldg(reinterpret_cast<const float4*>(a.att + p * 4));
        for (int c = bw; c < a.MXpad; c += 8) {
          float val = 0.f;
          if (ok) {
            if (a.e != nullptr) {
              if (c < 256) {
                const float ef = __ldg(a.e + p * 64 + (c >> 2));
                const int hd = c & 3;
                val = ef * (hd == 0 ? at.x : hd == 1 ? at.y : hd == 2 ? at.z : at.w);
              }
            } else if (c < a.xw) {
              val = __ldg(a.X + p * a.ldx + c);
            }
            if (c == a.ones_col) val = 1.0f;
          }
          float hi, lo;
          split_tf32(val, hi, lo);
          const uint32_t off = sw128_offset((uint32_t)c, (uint32_t)(r >> 2)) + kof;
          *reinterpret_cast<float*>(ximgp + off) = hi;
          *reinterpret_cast<float*>(ximgp + ximg + off) = lo;
        }
        const int grows = (a.NG + 7) / 8 * 8;
        for (int c = bw; c < grows; c += 8) {
          float val = 0.f;
          if (ok && c < a.NG && ncol0 + c < a.gw) val = __ldg(a.G + p * a.ldg + ncol0 + c);
          float hi, lo;
          split_tf32(val, hi, lo);
          const uint32_t off = sw128_offset((uint32_t)c, (uint32_t)(r >> 2)) + kof;
          *reinterpret_cast<float*>(gimgp + off) = hi;
          *reinterpret_cast<float*>(gimgp + gimg + off) = lo;
        }
      } else {
      for (int idx = bt; idx < CF::KP * xu; idx += 256) {
        const int r = idx / xu, ug = idx - r * xu;
        const long long p = p0 + r;
        const int c0 = ug * CF::EPU;
        float vals[CF::EPU];
#pragma unroll
        for (int i = 0; i < CF::EPU; ++i) vals[i] = 0.f;
        if (p < p_end) {
          if (a.e != nullptr) {             // X = e (x) att, feature c = f*4 + head
            if (c0 < 256) {
              const float4 at = __ldg(reinterpret_cast<const float4*>(a.att + p * 4));
#pragma unroll
              for (int q = 0; q < CF::EPU / 4; ++q) {
                const float ef = __ldg(a.e + p * 64 + c0 / 4 + q);
                vals[4 * q] = ef * at.x; vals[4 * q + 1] = ef * at.y; vals[4 * q + 2] = ef * at.z; vals[4 * q + 3] = ef * at.w;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < CF::EPU; ++i)
              if (c0 + i < a.xw) vals[i] = __ldg(a.X + p * a.ldx + c0 + i);
          }
#pragma unroll
          for (int i = 0; i < CF::EPU; ++i)
            if (c0 + i == a.ones_col) vals[i] = 1.0f;
        }
        const int mb = c0 / CF::BLK, u = (c0 % CF::BLK) / CF::EPU;
        xtg_store_unit<CF>(ximgp, ximg, (uint32_t)mb * LBO + sw128_offset((uint32_t)r, (uint32_t)u), vals);
      }
      for (int idx = bt; idx < CF::KP * gu; idx += 256) {
        const int r = idx / gu, ug = idx - r * gu;
        const long long p = p0 + r;
        const int c0 = ug * CF::EPU;
        float vals[CF::EPU];
#pragma unroll
        for (int i = 0; i < CF::EPU; ++i) vals[i] = (p < p_end && ncol0 + c0 + i < a.gw) ? __ldg(a.G + p * a.ldg + ncol0 + c0 + i) : 0.f;
        const int mb = c0 / CF::BLK, u = (c0 % CF::BLK) / CF::EPU;
        xtg_store_unit<CF>(gimgp, gimg, (uint32_t)mb * LBO + sw128_offset((uint32_t)r, (uint32_t)u), vals);
      }
      }
      fence_proxy_async();
      mbar_arrive(full + s);
    }
    // ------------------------------------------------------------ epilogue: flush the accumulator
    if (nst > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int q = warp & 3, half = (warp - 1) >> 2;      // two warps per lane quarter
      const int nchunks = (a.NG + 31) / 32;
      for (int mh = 0; mh < MH; ++mh) {
        const int row = mh * 128 + q * 32 + lane;
        for (int cc = half; cc < nchunks; cc += 2) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mh * a.NG + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int col = cc * 32 + k;
            if (col >= a.NG) continue;
            if (row < a.out_rows) {
              if (ncol0 + col < a.out_cols) atomicAdd(a.out + (size_t)row * a.ldo + ncol0 + col, v[k]);
            } else if (row < a.out_rows + a.extra_rows) {
              if (ncol0 + col < a.extra_ld) atomicAdd(a.extra + (size_t)(row - a.out_rows) * a.extra_ld + ncol0 + col, v[k]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int tc_xtg(const XtgArgs& a0, int engine, int prof_kind, cudaStream_t st) {
  XtgArgs a = a0;
  if (a.P <= 0) return 0;
  const bool bf = engine == SAKE_ENGINE_BF16;
  const int kp = 32;
  int ny = 1;
  if (!bf && a.NG > 128) { ny = (a.NG + 127) / 128; a.NG = 128; }      // tf32: <= 128 G columns per CTA (smem)
  if (a.extra_ld == 0) a.extra_ld = a0.NG;
  if (a.MXpad % 128 != 0 || a.MXpad > 256 || a.NG % 16 != 0 || a.NG > 256 || (a.MXpad / 128) * a.NG > 512) {
    set_error("tc_xtg: unsupported shape MXpad=%d NG=%d", a.MXpad, a.NG);
    return SAKE_EUNSUPPORTED;
  }
  size_t stage;
  if (bf) stage = (size_t)(a.MXpad / 64) * kp * 128 + (size_t)((a.NG + 63) / 64) * kp * 128;
  else stage = 2 * ((size_t)a.MXpad * 128 + (size_t)((a.NG + 7) / 8 * 8) * 128);
  const size_t smem = XTG_NSTAGE * stage + 256 + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long stages_total = (a.P + kp - 1) / kp;
  long long per = (stages_total * ny + sms - 1) / sms;
  if (per < 4) per = 4;                                  // keep the atomic flush amortised
  a.pairs_per_cta = per * kp;
  const int gx = (int)((a.P + a.pairs_per_cta - 1) / a.pairs_per_cta);
  static bool attr_tf = false, attr_bf = false;
  if (bf) {
    if (!attr_bf) { SAKE_CUDA_CHECK(cudaFuncSetAttribute(k_tc_xtg<SAKE_ENGINE_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_bf = true; }
  } else {
    if (!attr_tf) { SAKE_CUDA_CHECK(cudaFuncSetAttribute(k_tc_xtg<SAKE_ENGINE_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_tf = true; }
  }
  if (smem > 200 * 1024) { set_error("tc_xtg: smem %zu", smem); return SAKE_EUNSUPPORTED; }
  {
    ProfScope prof(prof_kind, a.P, st);
    dim3 grid(gx, ny);
    if (bf) k_tc_xtg<SAKE_ENGINE_BF16><<<grid, XTG_THREADS, smem, st>>>(a);
    else k_tc_xtg<SAKE_ENGINE_TF32X3><<<grid, XTG_THREADS, smem, st>>>(a);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sake
