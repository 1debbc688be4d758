// C ABI of libsake_b200.so (see include/sake_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace sake {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;   // kernels launched by this library (diagnostic counter)
void note_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int dense_fwd(long long rows, int in, int out, int act, const float* x, const float* w, const float* b, float* y,
              const RaggedHdr* hdr, cudaStream_t st);
int dense_bwd(long long rows, int in, int out, int act, const float* x, const float* w, const float* b,
              const float* dy, float* dx, float* dw, float* db, const RaggedHdr* hdr, cudaStream_t st);
int tc_selftest(float* max_abs_err, cudaStream_t st);
int tc_debug_counters(unsigned long long* out8);
int tc_debug_counters_bwd(unsigned long long* out16);

static int make_dims(const SakeDims* s, Dims* d) {
  if (!s) { set_error("dims is NULL"); return SAKE_EINVAL; }
  if (s->B < 0 || s->N <= 0 || s->H <= 0 || s->A <= 0 || s->K <= 0) {
    set_error("bad dims B=%d N=%d H=%d A=%d K=%d", s->B, s->N, s->H, s->A, s->K);
    return SAKE_EINVAL;
  }
  if ((long long)s->B * s->N > 0x7fffffffLL / 4) { set_error("B*N too large"); return SAKE_EUNSUPPORTED; }
  d->B = s->B; d->N = s->N; d->H = s->H; d->A = s->A; d->K = s->K;
  d->C = s->A * s->H;
  d->R = s->B * s->N;
  d->P = (long long)d->R * s->N;
  d->Kp = (s->K + 3) / 4 * 4;
  d->NP = 2 * d->Kp + 2 * s->H;
  d->update = (s->flags & SAKE_UPDATE) != 0;
  d->has_v = (s->flags & SAKE_HAS_V) != 0;
  d->has_mask = (s->flags & SAKE_HAS_MASK) != 0;
  d->spatial = (s->flags & SAKE_NO_SPATIAL) == 0;
  d->prepared = (s->flags & SAKE_WEIGHTS_PREPARED) != 0;
  d->g8 = 0;                                  // set with the engine (sake_layer_fwd / bwd)
  d->cutoff = (s->flags & SAKE_COSINE_CUTOFF) != 0;
  d->cut_lo = s->cutoff_lower; d->cut_hi = s->cutoff_upper;
  if (d->cutoff && !(d->cut_hi > d->cut_lo)) { set_error("cosine cutoff needs upper > lower (got %g, %g)", d->cut_lo, d->cut_hi); return SAKE_EINVAL; }
  d->hdr = nullptr; d->rowinfo = nullptr; d->tileinfo = nullptr; d->molinfo = nullptr;
  d->pair_u = nullptr; d->pair_p = nullptr;
  return 0;
}
// ragged batches (n_real-packed, see ragged.cu): compact tensors, no float mask, tcgen05 engines only
static int attach_pair(Dims* d, const SakePairTerms* pair, const void* ragged) {
  if (!pair) return 0;
  if (!pair->u || !pair->p) { set_error("SakePairTerms needs both u and p"); return SAKE_EINVAL; }
  if (ragged) { set_error("edge features (SakePairTerms) are not available for ragged batches"); return SAKE_EUNSUPPORTED; }
  if (((reinterpret_cast<uintptr_t>(pair->u) | reinterpret_cast<uintptr_t>(pair->p)) & 15) != 0) { set_error("SakePairTerms buffers must be 16-byte aligned"); return SAKE_EINVAL; }
  d->pair_u = pair->u; d->pair_p = pair->p;
  return 0;
}
static int attach_ragged(Dims* d, const void* ragged, const float* mask, int engine) {
  if (!ragged) return 0;
  if (mask || d->has_mask) { set_error("a ragged batch carries no float mask (padding atoms are not stored)"); return SAKE_EINVAL; }
  if (d->N > 128) { set_error("ragged batches need N <= 128 (got %d)", d->N); return SAKE_EUNSUPPORTED; }
  if (engine == SAKE_ENGINE_FP32) { set_error("ragged batches run on the tcgen05 engines only (H=64, A=4)"); return SAKE_EUNSUPPORTED; }
  ragged_attach(*d, ragged);
  return 0;
}

// Every leaf the configuration reads must be present: a flax tree initialised with v=None has no
// velocity_mlp (layers.py:226-229), update=False has no v_mixing (layers.py:94,217) — applying such a tree
// with v / update is a missing-parameter error in the reference, and an error (not a NULL dereference or a
// silent zero weight) here.  `what` is "params" or "grads".
template <class PS>
static int check_leaves(const Dims& d, const PS& p, const char* what) {
#define SAKE_NEED(field)                                                                                  \
  if (!p.field) { set_error("%s: required leaf `%s` is NULL for this configuration (update=%d has_v=%d spatial=%d)", \
                            what, #field, d.update, d.has_v, d.spatial); return SAKE_EINVAL; }
  SAKE_NEED(rbf_means) SAKE_NEED(rbf_betas) SAKE_NEED(mlp_in_kernel) SAKE_NEED(mlp_in_bias)
  SAKE_NEED(mlp_out0_kernel) SAKE_NEED(mlp_out0_bias) SAKE_NEED(mlp_out2_kernel) SAKE_NEED(mlp_out2_bias)
  SAKE_NEED(sem_kernel) SAKE_NEED(sem_bias)
  SAKE_NEED(node0_kernel) SAKE_NEED(node0_bias) SAKE_NEED(node2_kernel) SAKE_NEED(node2_bias)
  // flax creates x_mixing / post_norm_mlp even with use_spatial_attention=False (layers.py:208-212 calls
  // spatial_attention and then zeroes its outputs), and the per-node kernels read them unconditionally
  SAKE_NEED(x_mixing_kernel) SAKE_NEED(post0_kernel) SAKE_NEED(post0_bias) SAKE_NEED(post2_kernel) SAKE_NEED(post2_bias)
  if (d.update && d.spatial) { SAKE_NEED(v_mixing_kernel) }
  if (d.update && d.has_v) { SAKE_NEED(vel0_kernel) SAKE_NEED(vel0_bias) SAKE_NEED(vel2_kernel) }
#undef SAKE_NEED
  return 0;
}

// ---- deferred weight-gradient work (SAKE_DEFER_DW): one side stream per device -----------------------------
struct SideCtx {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, done[2] = {nullptr, nullptr}, all = nullptr;
  bool pending[2] = {false, false}, any = false;
};
static SideCtx g_side[64];
static SideCtx* side_ctx() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  SideCtx* c = &g_side[dev & 63];
  if (!c->s) {
    if (cudaStreamCreateWithFlags(&c->s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->done[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->done[1], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->all, cudaEventDisableTiming);
  }
  return c;
}

static inline bool H_is_64(const Dims& d) { return d.H == 64; }

static int resolve_engine(const SakeDims* s, const Dims& d) {
  int e = s->engine;
  if (e == SAKE_ENGINE_AUTO) return tc_supported(d) ? SAKE_ENGINE_F16X2 : SAKE_ENGINE_FP32;
  if (e == SAKE_ENGINE_FP32) return e;
  if (e == SAKE_ENGINE_TF32X3 || e == SAKE_ENGINE_BF16 || e == SAKE_ENGINE_F16X2) {
    if (!tc_supported(d)) {
      set_error("tcgen05 engine needs H=64, A=4 (C=256), n_rbf <= 58; got H=%d A=%d K=%d", d.H, d.A, d.K);
      return SAKE_EUNSUPPORTED;
    }
    return e;
  }
  set_error("unknown engine %d", e);
  return SAKE_EINVAL;
}

struct SavedLayout { size_t e, att, logit, ssum, he, nodeproj, nstash, wmix, wedge, wnode, nodeWT, total; };
static SavedLayout saved_layout(const Dims& d, int engine) {
  SavedLayout L;
  size_t o = 0;
  L.e = o; o += align_up(sizeof(float) * rows_pad8(d.P) * d.H);
  L.att = o; o += align_up(sizeof(float) * (size_t)d.P * d.A);
  L.logit = o; o += align_up(sizeof(float) * (size_t)d.P * d.A);
  L.ssum = o; o += align_up(sizeof(float) * rows_pad128(d.R) * d.C * 3);
  L.he = o; o += align_up(sizeof(float) * rows_pad8(d.R) * d.C);
  L.nodeproj = o; o += align_up(sizeof(float) * rows_pad8(d.R) * d.NP);
  L.nstash = o; o += align_up(sizeof(float) * rows_pad128(d.R) * NS_LD);
  L.wmix = L.wedge = L.wnode = L.nodeWT = o;
  if (engine != SAKE_ENGINE_FP32) {
    // 1024-byte aligned: the images are sources of bulk (TMA) copies
    o = align_up(o, 1024);
    L.wmix = o; o += align_up(tc_scratch_bytes(d, engine, 1, 1), 1024);
    L.wedge = o; if (tc_edge_supported(d)) o += align_up(edge_w_bytes(), 1024);
    L.wnode = o; if (tc_node_supported(d)) o += align_up(tc_node_w_bytes(), 1024);
    L.nodeWT = o; o += align_up(node_wt_bytes(d));
  }
  L.total = o;
  return L;
}
static Saved carve_saved(const Dims& d, void* base, bool tc_edge, int engine) {
  SavedLayout L = saved_layout(d, engine);
  char* b = (char*)base;
  Saved s;
  s.e = (float*)(b + L.e); s.att = (float*)(b + L.att); s.ssum = (float*)(b + L.ssum);
  s.logit = tc_edge ? (float*)(b + L.logit) : s.att;   // tcgen05 edge path keeps the logits for the backward pass
  s.he = (float*)(b + L.he); s.nodeproj = (float*)(b + L.nodeproj); s.nstash = (float*)(b + L.nstash);
  s.wmix = b + L.wmix; s.wedge = b + L.wedge; s.wnode = b + L.wnode; s.nodeWT = (float*)(b + L.nodeWT);
  // the tcgen05 node kernels read ssum tile-transposed; k_tc_mix_fwd writes it that way when they run
  s.ssum_tt = engine != SAKE_ENGINE_FP32 && tc_node_supported(d) && true;
  return s;
}

struct ScratchLayout { size_t T, tmax, ghe, ge, gatt, gdir, gcut, gproj, wxT, nodeWT, gZ, edgeb, xtgp, nbuf, noded, total; };
static ScratchLayout scratch_layout(const Dims& d, int engine, int for_backward, int with_grads) {
  ScratchLayout L;
  memset(&L, 0, sizeof(L));
  size_t o = 0;
  if (for_backward) {
    L.T = o; o += align_up(sizeof(float) * (size_t)d.R * d.C * 4);
    L.tmax = o; o += align_up(sizeof(float) * (size_t)d.R);
    L.ghe = o; o += align_up(sizeof(float) * (size_t)d.R * d.C);
    L.ge = o; o += align_up(sizeof(float) * rows_pad8(d.P) * d.H);
    L.gatt = o; o += align_up(sizeof(float) * (size_t)d.P * d.A);
    L.gdir = o; o += align_up(sizeof(float) * (size_t)d.P * 3);
    L.gcut = o; if (d.cutoff) o += align_up(sizeof(float) * (size_t)d.P);
    L.gproj = o; o += align_up(sizeof(float) * rows_pad8(d.R) * d.NP);
    L.wxT = o; o += align_up(sizeof(float) * (size_t)d.C * d.C);
    L.nodeWT = o; o += align_up(sizeof(float) * ((size_t)d.H * (2 * d.H + d.C) + 3 * (size_t)d.H * d.H + (size_t)d.H * d.C +
                                                  (size_t)d.K * 2 * d.H + (size_t)d.H * 2 * d.H));
    L.gZ = o;
    if (with_grads) o += align_up(sizeof(float) * rows_pad8(d.P) * d.C);
  }
  L.edgeb = o;
  if (engine != SAKE_ENGINE_FP32 && tc_edge_supported(d) && for_backward) o += align_up(tc_edge_bwd_scratch_bytes(d, with_grads));
  L.xtgp = o;
  if (engine != SAKE_ENGINE_FP32 && for_backward && with_grads) o += 2 * align_up(tc_xtg_partial_bytes());   // two regions: SAKE_DEFER_REDUCE
  L.nbuf = o;
  if (engine != SAKE_ENGINE_FP32 && for_backward && with_grads) o += align_up(tc_node_dw_scratch_bytes(d));
  L.noded = o;
  if (engine != SAKE_ENGINE_FP32 && tc_node_supported(d) && for_backward) o += align_up(tc_node_bwd_scratch_bytes(d));
  L.total = o + 256;
  return L;
}

}  // namespace sake

using namespace sake;

extern "C" {

const char* sake_version(void) { return "sake_b200 0.1 (sm_100a)"; }
const char* sake_last_error(void) { return g_err; }

int sake_resolve_engine(const SakeDims* dims) {
  Dims d;
  int rc = make_dims(dims, &d);
  if (rc) return rc;
  return resolve_engine(dims, d);
}

size_t sake_layer_saved_bytes(const SakeDims* dims) {
  Dims d;
  if (make_dims(dims, &d)) return 0;
  int e = resolve_engine(dims, d);
  if (e < 0) return 0;
  return saved_layout(d, e).total;
}

size_t sake_layer_scratch_bytes(const SakeDims* dims, int for_backward, int with_param_grads) {
  Dims d;
  if (make_dims(dims, &d)) return 0;
  int e = resolve_engine(dims, d);
  if (e < 0) return 0;
  return scratch_layout(d, e, for_backward, with_param_grads).total;
}

int sake_layer_fwd(const SakeDims* dims, const SakeLayerParams* params, const float* h, const float* x,
                   const float* v, const float* mask, const void* ragged, const SakePairTerms* pair, float* h_out,
                   float* x_out, float* v_out,
                   void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, sake_stream_t stream) {
  Dims d;
  int rc = make_dims(dims, &d);
  if (rc) return rc;
  if (!params || !h || !x || !h_out || !x_out || !saved) { set_error("sake_layer_fwd: NULL argument"); return SAKE_EINVAL; }
  if (d.has_v != (v != nullptr)) { set_error("SAKE_HAS_V flag does not match v pointer"); return SAKE_EINVAL; }
  if (d.has_mask != (mask != nullptr)) { set_error("SAKE_HAS_MASK flag does not match mask pointer"); return SAKE_EINVAL; }
  if (d.update && !v_out) { set_error("update=True needs v_out"); return SAKE_EINVAL; }
  if ((rc = check_leaves(d, *params, "params"))) return rc;
  int engine = resolve_engine(dims, d);
  if (engine < 0) return engine;
  d.g8 = engine != SAKE_ENGINE_FP32;
  if (saved_bytes < saved_layout(d, engine).total) { set_error("saved buffer too small: %zu < %zu", saved_bytes, saved_layout(d, engine).total); return SAKE_EINVAL; }
  if ((reinterpret_cast<uintptr_t>(saved) & 255) != 0) { set_error("saved buffer must be 256-byte aligned"); return SAKE_EINVAL; }
  ScratchLayout SL = scratch_layout(d, engine, 0, 0);
  if (SL.total > 256 && (!scratch || scratch_bytes < SL.total)) { set_error("scratch buffer too small: %zu < %zu", scratch_bytes, SL.total); return SAKE_EINVAL; }
  if (d.R == 0) return 0;
  if ((rc = attach_ragged(&d, ragged, mask, engine))) return rc;
  if ((rc = attach_pair(&d, pair, ragged))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_edge = engine != SAKE_ENGINE_FP32 && tc_edge_supported(d);
  Saved sv = carve_saved(d, saved, tc_edge, engine);
  const bool tc_node_f = engine != SAKE_ENGINE_FP32 && tc_node_supported(d);
  if ((rc = tc_node_f ? tc_node_pre(d, *params, h, sv, st) : gen_node_pre(d, *params, h, sv, st))) return rc;
  if (tc_edge) rc = tc_edge_fwd(d, *params, x, mask, sv, sv.wedge, st);
  else rc = gen_edge_fwd(d, *params, x, mask, sv, st);
  if (rc) return rc;
  if ((rc = gen_attn_fwd(d, x, mask, sv, st))) return rc;
  if (!d.spatial) {
    SAKE_CUDA_CHECK(cudaMemsetAsync(sv.ssum, 0, sizeof(float) * rows_pad128(d.R) * d.C * 3, st));
  } else if (engine == SAKE_ENGINE_FP32) {
    if ((rc = gen_mix_fwd(d, *params, x, mask, sv, st))) return rc;
  } else {
    if ((rc = tc_mix_fwd(d, *params, x, mask, sv, sv.wmix, engine, st))) return rc;
  }
  if (engine != SAKE_ENGINE_FP32 && tc_node_supported(d) && true) {
    if (!d.prepared && (rc = gen_node_wt(d, *params, sv.nodeWT, st))) return rc;   // for the backward call's k_node_pre_bwd
    return tc_node_post(d, *params, h, x, v, mask, h_out, x_out, v_out, sv, sv.wnode, st);
  }
  return gen_node_post(d, *params, h, x, v, mask, h_out, x_out, v_out, sv, st);
}

int sake_layer_prepare(const SakeDims* dims, const SakeLayerParams* params, void* saved, size_t saved_bytes,
                       sake_stream_t stream) {
  Dims d;
  int rc = make_dims(dims, &d);
  if (rc) return rc;
  if (!params || !saved) { set_error("sake_layer_prepare: NULL argument"); return SAKE_EINVAL; }
  if ((rc = check_leaves(d, *params, "params"))) return rc;
  int engine = resolve_engine(dims, d);
  if (engine < 0) return engine;
  d.g8 = engine != SAKE_ENGINE_FP32;
  if (engine == SAKE_ENGINE_FP32) return 0;
  if (saved_bytes < saved_layout(d, engine).total) { set_error("saved buffer too small"); return SAKE_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  Saved sv = carve_saved(d, saved, tc_edge_supported(d), engine);
  if (d.spatial && (rc = tc_mix_prepare(*params, sv.wmix, engine, st))) return rc;
  if (tc_edge_supported(d) && (rc = tc_edge_prepare(d, *params, sv.wedge, st))) return rc;
  if (tc_node_supported(d)) {
    if ((rc = tc_node_prepare(d, *params, sv.wnode, st))) return rc;
    if ((rc = gen_node_wt(d, *params, sv.nodeWT, st))) return rc;
  }
  return 0;
}

int sake_layer_bwd(const SakeDims* dims, const SakeLayerParams* params, const float* h, const float* x,
                   const float* v, const float* mask, const void* ragged, const SakePairTerms* pair, const void* saved,
                   size_t saved_bytes, const float* dh_out,
                   const float* dx_out, const float* dv_out, float* dh, float* dx, float* dv,
                   const SakeLayerGrads* grads, void* scratch, size_t scratch_bytes, sake_stream_t stream) {
  Dims d;
  int rc = make_dims(dims, &d);
  if (rc) return rc;
  if (!params || !h || !x || !saved || !dh_out || !dh || !dx || !scratch) { set_error("sake_layer_bwd: NULL argument"); return SAKE_EINVAL; }
  if (d.has_v != (v != nullptr)) { set_error("SAKE_HAS_V flag does not match v pointer"); return SAKE_EINVAL; }
  if (d.has_mask != (mask != nullptr)) { set_error("SAKE_HAS_MASK flag does not match mask pointer"); return SAKE_EINVAL; }
  if (d.has_v && !dv) { set_error("v given but dv is NULL"); return SAKE_EINVAL; }
  if ((rc = check_leaves(d, *params, "params"))) return rc;
  if (grads && (rc = check_leaves(d, *grads, "grads"))) return rc;
  int engine = resolve_engine(dims, d);
  if (engine < 0) return engine;
  d.g8 = engine != SAKE_ENGINE_FP32;
  if (saved_bytes < saved_layout(d, engine).total) { set_error("saved buffer too small"); return SAKE_EINVAL; }
  ScratchLayout SL = scratch_layout(d, engine, 1, grads != nullptr);
  if (scratch_bytes < SL.total) { set_error("scratch buffer too small: %zu < %zu", scratch_bytes, SL.total); return SAKE_EINVAL; }
  if (d.R == 0) return 0;
  if ((rc = attach_ragged(&d, ragged, mask, engine))) return rc;
  if ((rc = attach_pair(&d, pair, ragged))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // deferred weight gradients: wait for the previous user of this scratch slot before touching the scratch
  SideCtx* side = nullptr;
  const int slot = dims->reserved & 1;
  const bool defer_all = (dims->flags & SAKE_DEFER_DW) != 0, defer_red = !defer_all && (dims->flags & SAKE_DEFER_REDUCE) != 0;
  if ((defer_all || defer_red) && grads && engine != SAKE_ENGINE_FP32) {
    side = side_ctx();
    if (!side) { set_error("SAKE_DEFER_DW / SAKE_DEFER_REDUCE: cannot create the side stream"); return SAKE_ECUDA; }
    if (side->pending[slot]) { SAKE_CUDA_CHECK(cudaStreamWaitEvent(st, side->done[slot], 0)); side->pending[slot] = false; }
  }
  const bool tc_edge = engine != SAKE_ENGINE_FP32 && tc_edge_supported(d);
  Saved sv = carve_saved(d, const_cast<void*>(saved), tc_edge, engine);
  char* b = (char*)scratch;
  BwdScratch sc;
  sc.T = (float*)(b + SL.T); sc.tmax = (float*)(b + SL.tmax); sc.ghe = (float*)(b + SL.ghe); sc.ge = (float*)(b + SL.ge);
  sc.gatt = (float*)(b + SL.gatt); sc.gdir = (float*)(b + SL.gdir); sc.gproj = (float*)(b + SL.gproj);
  sc.gcut = d.cutoff ? (float*)(b + SL.gcut) : nullptr;
  sc.wxT = (float*)(b + SL.wxT); sc.gZ = (float*)(b + SL.gZ); sc.nodeWT = (float*)(b + SL.nodeWT);
  const bool tc_node = engine != SAKE_ENGINE_FP32 && tc_node_supported(d) && true;
  if (tc_node) sc.nodeWT = sv.nodeWT;                 // transposed copies left by the forward call
  sc.xtg_partial = (engine != SAKE_ENGINE_FP32 && grads)
                       ? (float*)(b + SL.xtgp + ((dims->flags & SAKE_DEFER_REDUCE) ? (size_t)slot * align_up(tc_xtg_partial_bytes()) : 0))
                       : nullptr;
  sc.nbuf = (engine != SAKE_ENGINE_FP32 && grads) ? (float*)(b + SL.nbuf) : nullptr;
  sc.qv = (tc_node && grads && d.update && d.spatial) ? (float*)(b + SL.noded) : nullptr;
  if (tc_node)
    rc = tc_node_post_bwd(d, *params, h, v, mask, sv, dh_out, dx_out, dv_out, dh, dx, dv, grads, sc, sv.wnode,
                          b + SL.noded, st);
  else
    rc = gen_node_post_bwd(d, *params, h, x, v, mask, sv, dh_out, dx_out, dv_out, dh, dx, dv, grads, sc, st);
  if (rc) return rc;
  XtgList xl;
  if (sc.nbuf && (rc = tc_node_dw(d, h, sv, tc_node, *grads, sc, xl, st))) return rc;
  float* gWx = grads ? grads->x_mixing_kernel : nullptr;
  if (engine == SAKE_ENGINE_FP32 || !d.spatial) {
    if ((rc = gen_mix_bwd(d, *params, x, mask, sv, sc, gWx, st))) return rc;
  } else {
    if ((rc = tc_mix_bwd(d, *params, x, mask, sv, sc, gWx, sv.wmix, engine, xl, st))) return rc;
  }
  // tcgen05 edge path: softmax backward only (celu' comes from the saved logits, the W_s g_q term of g_e is
  // added inside the edge kernel); generic path: the original kernel that recomputes q and updates g_e
  if ((rc = tc_edge ? tc_attn_bwd(d, x, sv, sc, st) : gen_attn_bwd(d, *params, x, sv, sc, st))) return rc;
  float* gpu_ = pair ? pair->g_u : nullptr;
  float* gpp_ = pair ? pair->g_p : nullptr;
  if (tc_edge)
    rc = tc_edge_bwd(d, *params, x, mask, sv, sc, dx, grads, sv.wedge, b + SL.edgeb, xl, gpu_, gpp_, st);
  else
    rc = gen_edge_bwd(d, *params, x, sv, dx, grads, sc, gpu_, gpp_, st);
  if (rc) return rc;
  const bool pre_dw_tc = tc_edge && grads && H_is_64(d);
  if (pre_dw_tc && (rc = tc_node_pre_dw(d, h, *grads, sc, xl))) return rc;
  if (xl.n > 0) {
    // every weight-gradient contraction of this layer in one batched tensor-core launch (+ its reductions);
    // nothing downstream of the layer reads dW, so with SAKE_DEFER_DW it runs on the side stream, forked here
    cudaStream_t ws = st;
    if (side && defer_all) {
      SAKE_CUDA_CHECK(cudaEventRecord(side->fork, st));
      SAKE_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->fork, 0));
      ws = side->s;
    }
    // SAKE_DEFER_REDUCE: the contractions stay on the caller's stream, only the partial-sum reduction is forked
    if ((rc = tc_xtg_flush(xl, sc.xtg_partial, engine, 3, ws, side && defer_red ? side->s : nullptr, side ? side->fork : nullptr))) return rc;
    if (side) {
      SAKE_CUDA_CHECK(cudaEventRecord(side->done[slot], side->s));
      side->pending[slot] = true;
      side->any = true;
    }
  }
  if (tc_node && (pre_dw_tc || !grads) && d.NP <= 256) return tc_node_pre_bwd(d, sv, sc, dh, st);
  return gen_node_pre_bwd(d, *params, h, dh, pre_dw_tc ? nullptr : grads, sc, st);
}

int sake_dw_sync(sake_stream_t stream) {
  SideCtx* side = side_ctx();
  if (!side) { set_error("sake_dw_sync: cannot create the side stream"); return SAKE_ECUDA; }
  if (side->any) {
    SAKE_CUDA_CHECK(cudaEventRecord(side->all, side->s));
    SAKE_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, side->all, 0));
    side->any = false;
    side->pending[0] = side->pending[1] = false;
  }
  return 0;
}

int sake_dense_fwd(int64_t rows, int32_t in_features, int32_t out_features, int32_t act, const float* x,
                   const float* kernel, const float* bias, float* y, const void* ragged, sake_stream_t stream) {
  if (rows < 0 || in_features <= 0 || out_features <= 0 || !x || !kernel || !y) { set_error("sake_dense_fwd: bad argument"); return SAKE_EINVAL; }
  note_launches(1);
  return dense_fwd(rows, in_features, out_features, act, x, kernel, bias, y, (const RaggedHdr*)ragged, (cudaStream_t)stream);
}

int sake_dense_bwd(int64_t rows, int32_t in_features, int32_t out_features, int32_t act, const float* x,
                   const float* kernel, const float* bias, const float* dy, float* dx, float* dkernel,
                   float* dbias, const void* ragged, sake_stream_t stream) {
  if (rows < 0 || in_features <= 0 || out_features <= 0 || !x || !kernel || !dy) { set_error("sake_dense_bwd: bad argument"); return SAKE_EINVAL; }
  note_launches(1);
  return dense_bwd(rows, in_features, out_features, act, x, kernel, bias, dy, dx, dkernel, dbias, (const RaggedHdr*)ragged,
                   (cudaStream_t)stream);
}

int sake_debug_counters(unsigned long long* out8) { return tc_debug_counters(out8); }
int sake_debug_counters_bwd(unsigned long long* out16) { return tc_debug_counters_bwd(out16); }

unsigned long long sake_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sake_selftest_xtg(int32_t engine, int64_t P, int32_t xw, int32_t gw, const float* X, const float* G, float* out,
                      sake_stream_t stream) {
  XtgArgs a;
  memset(&a, 0, sizeof(a));
  a.X = X; a.ldx = xw; a.xw = xw; a.ones_col = -1;
  a.G = G; a.ldg = gw; a.gw = gw;
  a.MXpad = (xw + 127) / 128 * 128; a.NG = (gw + 15) / 16 * 16; a.P = P;
  a.out = out; a.ldo = gw; a.out_rows = xw; a.out_cols = gw;
  return tc_xtg(a, engine, 0, (cudaStream_t)stream);
}

int sake_selftest_tcgen05(float* max_abs_err, sake_stream_t stream) {
  return tc_selftest(max_abs_err, (cudaStream_t)stream);
}

}  // extern "C"
