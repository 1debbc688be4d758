// XLA FFI handlers for libsake_b200 — thin adapters from XLA custom-call buffers to the C ABI.
// Compiled only when the jaxlib headers are available (make XLA_FFI_INCLUDE=...): this image has no
// jax / jaxlib, so this file is UNTESTED here; INTEGRATION.md shows the Python side.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define SAKE_HAVE_XLA_FFI 1
#endif
#endif

#ifdef SAKE_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include <cstring>
#include "../../include/sake_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static SakeDims make_dims(const ffi::AnyBuffer& h, int32_t heads, int32_t n_rbf, int32_t flags, int32_t engine = SAKE_ENGINE_AUTO,
                          float cut_lo = 0.f, float cut_hi = 5.f) {
  auto dims = h.dimensions();            // [B, N, H] (leading dims flattened on the Python side)
  SakeDims d;
  std::memset(&d, 0, sizeof(d));
  d.B = (int32_t)dims[0]; d.N = (int32_t)dims[1]; d.H = (int32_t)dims[2];
  d.A = heads; d.K = n_rbf; d.flags = flags; d.engine = engine;   // SAKE_ENGINE_* (AUTO = fp32-parity split when H = 64, A = 4)
  d.cutoff_lower = cut_lo; d.cutoff_upper = cut_hi;      // read only with SAKE_COSINE_CUTOFF in flags
  return d;
}
static const float* opt(const ffi::AnyBuffer& b) { return b.element_count() ? (const float*)b.untyped_data() : nullptr; }

static ffi::Error FwdImpl(cudaStream_t stream, ffi::AnyBuffer h, ffi::AnyBuffer x, ffi::AnyBuffer v, ffi::AnyBuffer mask,
                          ffi::RemainingArgs leaves, ffi::Result<ffi::AnyBuffer> h2, ffi::Result<ffi::AnyBuffer> x2,
                          ffi::Result<ffi::AnyBuffer> v2, ffi::Result<ffi::AnyBuffer> saved,
                          ffi::Result<ffi::AnyBuffer> scratch, int32_t n_heads, int32_t n_rbf, int32_t flags, int32_t engine) {
  SakeDims d = make_dims(h, n_heads, n_rbf, flags, engine);
  SakeLayerParams p;
  const float** pp = reinterpret_cast<const float**>(&p);
  for (size_t i = 0; i < sizeof(p) / sizeof(float*); ++i) {
    auto b = leaves.get<ffi::AnyBuffer>(i);
    pp[i] = (b.has_value() && b->element_count()) ? (const float*)b->untyped_data() : nullptr;
  }
  int rc = sake_layer_fwd(&d, &p, (const float*)h.untyped_data(), (const float*)x.untyped_data(), opt(v), opt(mask), /*ragged=*/nullptr, /*pair=*/nullptr,
                          (float*)h2->untyped_data(), (float*)x2->untyped_data(), (float*)v2->untyped_data(),
                          saved->untyped_data(), saved->size_bytes(), scratch->untyped_data(), scratch->size_bytes(),
                          stream);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::InvalidArgument(sake_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(SakeLayerFwd, FwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .RemainingArgs()
                                  .Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("n_heads").Attr<int32_t>("n_rbf").Attr<int32_t>("flags").Attr<int32_t>("engine"));

static ffi::Error BwdImpl(cudaStream_t stream, ffi::AnyBuffer h, ffi::AnyBuffer x, ffi::AnyBuffer v, ffi::AnyBuffer mask,
                          ffi::AnyBuffer saved, ffi::AnyBuffer dh2, ffi::AnyBuffer dx2, ffi::AnyBuffer dv2,
                          ffi::RemainingArgs leaves, ffi::RemainingRets rets, int32_t n_heads, int32_t n_rbf,
                          int32_t flags, int32_t engine) {
  SakeDims d = make_dims(h, n_heads, n_rbf, flags, engine);
  constexpr size_t NL = sizeof(SakeLayerParams) / sizeof(float*);
  SakeLayerParams p;
  SakeLayerGrads g;
  const float** pp = reinterpret_cast<const float**>(&p);
  float** gp = reinterpret_cast<float**>(&g);
  // rets: dh, dx, dv, NL gradient leaves, scratch
  for (size_t i = 0; i < NL; ++i) {
    auto b = leaves.get<ffi::AnyBuffer>(i);
    auto r = rets.get<ffi::AnyBuffer>(3 + i);
    pp[i] = (b.has_value() && b->element_count()) ? (const float*)b->untyped_data() : nullptr;
    gp[i] = (r.has_value() && (*r)->element_count()) ? (float*)(*r)->untyped_data() : nullptr;
    if (gp[i] && cudaMemsetAsync(gp[i], 0, (*r)->size_bytes(), stream) != cudaSuccess)      // the C ABI accumulates
      return ffi::Error::Internal("cudaMemsetAsync of a gradient buffer failed");
  }
  auto dh = *rets.get<ffi::AnyBuffer>(0);
  auto dx = *rets.get<ffi::AnyBuffer>(1);
  auto dv = *rets.get<ffi::AnyBuffer>(2);
  auto scratch = *rets.get<ffi::AnyBuffer>(3 + NL);
  int rc = sake_layer_bwd(&d, &p, (const float*)h.untyped_data(), (const float*)x.untyped_data(), opt(v), opt(mask), /*ragged=*/nullptr, /*pair=*/nullptr,
                          saved.untyped_data(), saved.size_bytes(), (const float*)dh2.untyped_data(), opt(dx2), opt(dv2),
                          (float*)dh->untyped_data(), (float*)dx->untyped_data(),
                          dv->element_count() ? (float*)dv->untyped_data() : nullptr, &g, scratch->untyped_data(),
                          scratch->size_bytes(), stream);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::InvalidArgument(sake_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(SakeLayerBwd, BwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int32_t>("n_heads").Attr<int32_t>("n_rbf").Attr<int32_t>("flags").Attr<int32_t>("engine"));
#endif  // SAKE_HAVE_XLA_FFI
