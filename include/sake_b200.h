/*
 * sake_b200.h — C ABI of libsake_b200.so: the B200 (sm_100a) implementation of SAKE's dense
 * spatial-attention message-passing layer (forward + backward).
 *
 * The reference (ArnNag/sake) is pure Python/JAX/flax and has NO plugin / FFI layer of its own;
 * the only interface the hot path sits behind is the flax module API
 *     sake/layers.py:42-52,188-235   DenseSAKELayer.__call__(h, x, v=None, mask=None, he=None)
 *     sake/models.py:11-61           DenseSAKEModel.__call__
 * so this header is the boundary a JAX custom call (XLA FFI) or any other host binds.  Each entry
 * point cites the reference code it replaces.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; all tensors are dense row-major fp32 DEVICE buffers.
 *   - The callee never allocates, frees or retains caller memory; outputs, `saved` and `scratch`
 *     are caller-allocated (sizes from the *_bytes queries).  Weights are read every call.
 *   - Enqueue-only on the caller's stream: no host synchronisation, CUDA-graph capturable.
 *   - Return 0 on success, a negative SAKE_E* code on error; sake_last_error() gives the message
 *     (thread-local).  Unsupported configurations are errors — there is no CPU fallback.
 *   - Parameter gradients are ACCUMULATED (+=) into the SakeLayerGrads buffers.
 */
#ifndef SAKE_B200_H_
#define SAKE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sake_stream_t; /* cudaStream_t */

enum {
  SAKE_OK = 0,
  SAKE_EINVAL = -1,       /* bad dims / null pointer / too-small buffer */
  SAKE_EUNSUPPORTED = -2, /* configuration outside what the kernels implement */
  SAKE_ECUDA = -3         /* a CUDA runtime call failed (message has the CUDA error string) */
};

/* flags (SakeDims.flags) — mirror the module fields / call arguments of DenseSAKELayer */
enum {
  SAKE_UPDATE = 1,       /* DenseSAKELayer.update (sake/layers.py:46,217)                      */
  SAKE_HAS_V = 2,        /* v argument is not None (sake/layers.py:226-229)                    */
  SAKE_HAS_MASK = 4,     /* mask argument is not None, float [B,N,N]                           */
  SAKE_NO_SPATIAL = 8,   /* use_spatial_attention=False (sake/layers.py:210-212)              */
  SAKE_COSINE_CUTOFF = 32, /* DenseSAKELayer.cutoff = sake.utils.cosine_cutoff(lower, upper) (sake/utils.py:10-26,
                            sake/layers.py:172-176): euclidean_attention = 0.5 (cos(pi (2 (d - lower) / (upper - lower) + 1)) + 1)
                            with SakeDims.cutoff_lower / cutoff_upper; as in the reference the range masks are
                            NOT applied (utils.py:24-25 discards them), so the factor is periodic in d       */
  SAKE_DEFER_REDUCE = 128, /* sake_layer_bwd only (training, tcgen05 engines): the partial-sum reduction that finishes the
                            layer's weight gradients runs on the library's side stream, under the per-node kernels of this
                            layer and of the next one (which occupy a quarter of the SMs); SakeDims.reserved & 1 selects one
                            of two partial-sum regions (alternate it between consecutive layers); sake_dw_sync() joins */
  SAKE_WEIGHTS_PREPARED = 64, /* sake_layer_fwd: the operand images of this layer's weights in `saved` are current
                            (sake_layer_prepare ran on the same parameters since they last changed): skip
                            rebuilding them.  Inference loops prepare once; a training step prepares after
                            every optimiser step (or leaves the flag clear and lets the forward call do it).  */
  SAKE_DEFER_DW = 16     /* sake_layer_bwd only: enqueue the weight-gradient contractions (dW = X^T G over all
                            pairs / atoms; nothing downstream of the layer reads them) on the library's side
                            stream so that they overlap the next layer's backward.  SakeDims.reserved = scratch
                            slot (0 or 1): the caller alternates two scratch buffers between consecutive layers,
                            and a call first makes `stream` wait for the deferred work of the previous call that
                            used the same slot.  Gradients are complete after sake_dw_sync(stream).           */
};

/* precision / engine (SakeDims.engine) */
enum {
  SAKE_ENGINE_AUTO = 0,   /* tcgen05 fp16-split (F16X2) when the shape allows, else the generic fp32 path */
  SAKE_ENGINE_FP32 = 1,   /* generic CUDA-core fp32 kernels, any H / A / K                      */
  SAKE_ENGINE_TF32X3 = 2, /* tcgen05.mma kind::tf32, hi/lo split (3 MMAs) — fp32-parity mode     */
  SAKE_ENGINE_BF16 = 3,   /* tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate) — fast mode  */
  SAKE_ENGINE_F16X2 = 4   /* tcgen05.mma kind::f16, fp16 hi/lo split with per-row power-of-two scaling
                             (3 MMAs at the 16-bit rate, 22 mantissa bits) — fp32-parity mode at half
                             the tensor and shared-memory cost of TF32X3                                */
};

typedef struct SakeDims {
  int32_t B;      /* number of molecules (all leading batch dims flattened)                      */
  int32_t N;      /* atoms per molecule (padded width)                                           */
  int32_t H;      /* hidden_features == in_features == out_features (residual, layers.py:150)   */
  int32_t A;      /* n_heads (layers.py:45)                                                      */
  int32_t K;      /* number of RBFs, kernel_features (layers.py:14) — 50 in every script        */
  int32_t flags;  /* SAKE_UPDATE | SAKE_HAS_V | SAKE_HAS_MASK | SAKE_NO_SPATIAL                  */
  int32_t engine; /* SAKE_ENGINE_*                                                               */
  int32_t reserved;
  float cutoff_lower; /* only read with SAKE_COSINE_CUTOFF                                         */
  float cutoff_upper;
} SakeDims;

/* Edge features `he` (DenseSAKELayer.__call__(..., he), sake/layers.py:201-202: h_cat_ht = concat(h_cat_ht, he)).
 * The concatenated block only ever meets two Dense layers — rows [2F, 2F+E) of edge_model/mlp_in and of
 * edge_model/mlp_out/layers_0 (sake/layers.py:30,33-38) — so it enters the layer as two per-pair additive terms
 * that the caller computes with two small GEMMs over he [B,N,N,E]:
 *     u = he @ mlp_in_kernel[2F:2F+E]        [B,N,N,Kp]  (K columns, zero-padded to Kp = K rounded up to 4)
 *     p = he @ mlp_out0_kernel[2F:2F+E]      [B,N,N,H]
 * and the kernels are handed the two kernels WITHOUT those rows ([2F, K] and [2F+K+1, H]).  The backward call
 * returns the cotangents g_u / g_p of the two terms (same shapes), from which the caller's autodiff obtains d he and
 * the gradients of the two row blocks.  Not available for ragged batches. */
typedef struct SakePairTerms {
  const float* u;
  const float* p;
  float* g_u; /* backward only; may be NULL */
  float* g_p;
} SakePairTerms;

/* Parameters of one DenseSAKELayer, flax layouts (kernel = [in, out]); SURVEY Appendix C.
 * C = A*H.  Pointers that a configuration does not use may be NULL. */
typedef struct SakeLayerParams {
  const float* rbf_means;       /* edge_model/kernel/means            [K]            utils.py:37 */
  const float* rbf_betas;       /* edge_model/kernel/betas            [K]            utils.py:43 */
  const float* mlp_in_kernel;   /* edge_model/mlp_in/kernel           [2H, K]       layers.py:19 */
  const float* mlp_in_bias;     /*                                    [K]                       */
  const float* mlp_out0_kernel; /* edge_model/mlp_out/layers_0/kernel [2H+K+1, H]   layers.py:22 */
  const float* mlp_out0_bias;   /*                                    [H]                       */
  const float* mlp_out2_kernel; /* edge_model/mlp_out/layers_2/kernel [H, H]        layers.py:24 */
  const float* mlp_out2_bias;   /*                                    [H]                       */
  const float* sem_kernel;      /* semantic_attention_mlp/layers_0    [H, A]        layers.py:80 */
  const float* sem_bias;        /*                                    [A]                       */
  const float* x_mixing_kernel; /* x_mixing/layers_0/kernel           [C, C]        layers.py:95 */
  const float* post0_kernel;    /* post_norm_mlp/layers_0             [C, H]        layers.py:87 */
  const float* post0_bias;      /*                                    [H]                       */
  const float* post2_kernel;    /* post_norm_mlp/layers_2             [H, H]        layers.py:89 */
  const float* post2_bias;      /*                                    [H]                       */
  const float* node0_kernel;    /* node_mlp/layers_0                  [H+C+H, H]    layers.py:61 */
  const float* node0_bias;      /*                                    [H]                       */
  const float* node2_kernel;    /* node_mlp/layers_2                  [H, H]        layers.py:63 */
  const float* node2_bias;      /*                                    [H]                       */
  const float* v_mixing_kernel; /* v_mixing/kernel                    [C, 1]        layers.py:94 */
  const float* vel0_kernel;     /* velocity_mlp/layers_0              [H, H]        layers.py:71 */
  const float* vel0_bias;       /*                                    [H]                       */
  const float* vel2_kernel;     /* velocity_mlp/layers_2/kernel       [H, 1]        layers.py:73 */
} SakeLayerParams;

/* Same field order as SakeLayerParams; gradient buffers (accumulated).  log_gamma has no
 * gradient in the dense layer (it is never read: layers.py:97-105 vs :107-235). */
typedef struct SakeLayerGrads {
  float* rbf_means;
  float* rbf_betas;
  float* mlp_in_kernel;
  float* mlp_in_bias;
  float* mlp_out0_kernel;
  float* mlp_out0_bias;
  float* mlp_out2_kernel;
  float* mlp_out2_bias;
  float* sem_kernel;
  float* sem_bias;
  float* x_mixing_kernel;
  float* post0_kernel;
  float* post0_bias;
  float* post2_kernel;
  float* post2_bias;
  float* node0_kernel;
  float* node0_bias;
  float* node2_kernel;
  float* node2_bias;
  float* v_mixing_kernel;
  float* vel0_kernel;
  float* vel0_bias;
  float* vel2_kernel;
} SakeLayerGrads;

/* Library / build information. */
const char* sake_version(void);
const char* sake_last_error(void);

/* Which engine SAKE_ENGINE_AUTO resolves to for these dims (one of SAKE_ENGINE_*; <0 on error). */
int sake_resolve_engine(const SakeDims* dims);

/* Bytes of the `saved` buffer: what sake_layer_fwd leaves for sake_layer_bwd (edge features
 * e[B,N,N,H], attention att[B,N,N,A], per-node projections and reductions, the weight operand images).
 * Nothing O(N^2*C).  The buffer is OPAQUE: its internal layout depends on the engine (the tcgen05 engines interleave
 * groups of 8 rows unit by unit) and may change between builds. */
size_t sake_layer_saved_bytes(const SakeDims* dims);
/* Bytes of the `scratch` buffer (temporaries; may be shared by all layers on one stream).
 * for_backward != 0 sizes it for sake_layer_bwd (with_param_grads selects the training variant). */
size_t sake_layer_scratch_bytes(const SakeDims* dims, int for_backward, int with_param_grads);

/* ---- ragged batches: compute the real atoms of a padded batch only -------------------------------------
 * The reference pads every molecule of a batch to N atoms and multiplies every O(N^2) tensor by
 * mask = outer(m, m) (scripts/qm9/run.py:23-24,35; sake/layers.py:120-123,137-138,164-165,178-179,219-221).
 * With `ragged` tables the layer instead works on a COMPACT layout that stores the n_real[b] real atoms of
 * every molecule only (molecules ordered by n_real, R = sum n_real rows, P = sum n_real^2 pairs): same
 * function on the real atoms (sake/tests/test_mask.py:202-240), no work on padding.
 *   sake_ragged_bytes      size of the table buffer for a [B, N] batch (N <= 128)
 *   sake_ragged_prepare    builds the tables on the device from n_real[B] (int32 DEVICE buffer, clamped to
 *                          [0, N]; real atoms are the first n_real[b] of each molecule); no host sync
 *   sake_ragged_gather     compact[r, :width] = padded[b, i, :width] for every real atom
 *   sake_ragged_scatter    padded[b, i, :width] = alpha * compact[r, :width]; padding rows are left untouched
 * Every entry point that takes `ragged` then expects compact tensors in buffers sized for the worst case
 * (B*N rows), mask = NULL, and SakeDims {B, N} of the padded batch.  The real counts stay on the device:
 * launch grids cover the worst case and kernels read their bounds from the tables. */
size_t sake_ragged_bytes(int32_t B, int32_t N);
int sake_ragged_prepare(int32_t B, int32_t N, const int32_t* n_real, void* ragged, size_t ragged_bytes,
                        sake_stream_t stream);
int sake_ragged_gather(const void* ragged, int32_t B, int32_t N, int32_t width, const float* padded,
                       float* compact, sake_stream_t stream);
int sake_ragged_scatter(const void* ragged, int32_t B, int32_t N, int32_t width, float alpha,
                        const float* compact, float* padded, sake_stream_t stream);

/* DenseSAKELayer.__call__ (sake/layers.py:188-235) with he=None, cutoff=None.
 * `ragged`: NULL, or the tables of sake_ragged_prepare (then mask must be NULL and h / x / v and all
 * outputs are compact; tcgen05 engines only).
 *   h [B,N,H], x [B,N,3], v [B,N,3] or NULL, mask [B,N,N] float or NULL
 *   -> h_out [B,N,H], x_out [B,N,3], v_out [B,N,3]
 * Coordinates are always 3 wide; 2-D systems (scripts/dw4) pad z = 0 on the host side.
 * With SAKE_UPDATE clear, x_out / v_out are plain copies of x / v (v_out untouched if v NULL).
 * Guarded masking: a row whose attention normaliser is 0 gets att = 0 (the reference yields
 * 0/0 = NaN there, layers.py:178-180); every other value follows the reference formulas. */
int sake_layer_fwd(const SakeDims* dims, const SakeLayerParams* params,
                   const float* h, const float* x, const float* v, const float* mask, const void* ragged,
                   const SakePairTerms* pair /* he terms, or NULL */,
                   float* h_out, float* x_out, float* v_out,
                   void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes,
                   sake_stream_t stream);

/* Builds the tcgen05 operand images of one layer's weights (swizzled, split-precision copies of x_mixing, the edge
 * MLP, the node-tail and per-node projection MLPs and their transposes; ~2.0 MB) into `saved`, where sake_layer_fwd (with
 * SAKE_WEIGHTS_PREPARED) and sake_layer_bwd read them.  No-op on the generic fp32 engine.  The images depend on the
 * parameters only: the reference's XLA program has no counterpart (it re-reads the flax kernels every call). */
int sake_layer_prepare(const SakeDims* dims, const SakeLayerParams* params, void* saved, size_t saved_bytes,
                       sake_stream_t stream);

/* Vector-Jacobian product of sake_layer_fwd (what jax.grad / jax.vjp of the layer computes:
 * scripts/md17/run.py:58 forces, scripts/qm9/run.py:84-89 parameter gradients).
 *   cotangents dh_out [B,N,H], dx_out [B,N,3] (NULL = 0), dv_out [B,N,3] (NULL = 0)
 *   -> dh [B,N,H], dx [B,N,3], dv [B,N,3] (dv may be NULL when v is NULL)
 *   grads: NULL for the forces-only (inference) variant, else accumulated parameter gradients.
 * `saved` must be the buffer written by sake_layer_fwd for the same inputs. */
int sake_layer_bwd(const SakeDims* dims, const SakeLayerParams* params,
                   const float* h, const float* x, const float* v, const float* mask, const void* ragged,
                   const SakePairTerms* pair /* he terms, or NULL */,
                   const void* saved, size_t saved_bytes,
                   const float* dh_out, const float* dx_out, const float* dv_out,
                   float* dh, float* dx, float* dv, const SakeLayerGrads* grads,
                   void* scratch, size_t scratch_bytes, sake_stream_t stream);

/* Joins the deferred weight-gradient work (SAKE_DEFER_DW) of this device into `stream`: everything enqueued after
 * this call sees the complete parameter gradients, and the scratch / saved buffers may be reused.  Capturable. */
int sake_dw_sync(sake_stream_t stream);

/* nn.Dense (+ optional silu) over the last axis — embedding_in / embedding_out of
 * DenseSAKEModel (sake/models.py:24-31,57,60).  y[rows,out] = act(x[rows,in] @ kernel + bias).
 * act: 0 = identity, 1 = silu.  bias may be NULL.  `ragged` (nullable): x / y are compact, `rows` is the
 * worst case B*N and the real row count is read from the tables on the device. */
int sake_dense_fwd(int64_t rows, int32_t in_features, int32_t out_features, int32_t act,
                   const float* x, const float* kernel, const float* bias, float* y,
                   const void* ragged, sake_stream_t stream);
/* VJP of sake_dense_fwd: dx (may be NULL), and accumulated dkernel / dbias (may be NULL). */
int sake_dense_bwd(int64_t rows, int32_t in_features, int32_t out_features, int32_t act,
                   const float* x, const float* kernel, const float* bias, const float* dy,
                   float* dx, float* dkernel, float* dbias, const void* ragged, sake_stream_t stream);

/* Energy head of the drivers: E[b] = sum_i sum_o y[b,i,o] * atom_mask[b,i]
 * (scripts/md17/run.py:46-52; masked sum of scripts/qm9/run.py:58-60), plus the cotangent dy that
 * starts the backward pass:
 *   mode 0 (forces, scripts/md17/run.py:54-58): dy = atom_mask  (d sum_b E / dy)
 *   mode 1 (L1 loss, scripts/qm9/run.py:79-82, coloring utils.py:7-8):
 *           loss += mean_b |std*E[b] + mean - target[b]| ;  dy = sign(.)*std/B * atom_mask
 * atom_mask [B,N] may be NULL (all ones); target NULL in mode 0; loss is a device scalar that is
 * accumulated (zero it first).  `ragged` (nullable): y / dy are compact (atom_mask must be NULL), energies
 * are still ordered by the molecule's index in the padded batch. */
int sake_energy_head(int32_t B, int32_t N, int32_t out_features, int32_t mode, const float* y,
                     const float* atom_mask, const float* target, float mean, float std,
                     float* energy, float* loss, float* dy, const void* ragged, sake_stream_t stream);

/* Coupling glue of the augmented flow (sake/flows.py:97-142; scripts/lj13_aug/run.py:32-43), one launch each:
 *   sake_flow_pre      AugmentedFlowLayer.mp, flows.py:118-122: h_aug[B,N+1,F+1] = [h | sum pos^2] with a zero
 *                      dummy atom appended, x_aug[B,N+1,3] = [pos ; 0]   (h may be NULL = zeros)
 *   sake_flow_post     flows.py:123-142 after the DenseSAKEModel call: translation = (x_aug_out - pos0)[:N] minus
 *                      its mean over atoms; scale = mean_i tanh(Dense(silu(Dense(y_i)))) with the scale_mlp leaves
 *                      (y = the model's h output, out_features = 1; hidden width <= 64);
 *                      direction +1 (f_forward):  other = exp(scale) * other + translation
 *                      direction -1 (f_backward): other = exp(-scale) * (other - translation)
 *                      logdet[b] += scale * N * D   (D = coordinate dimension of the system, 2 or 3)
 *   sake_flow_logprob  out[b] = -log p(x) - log p(v) + logdet[b], p = CenteredGaussian (flows.py:13-21)
 * All coordinate tensors are [B,N,3] (2-D systems carry z = 0). */
int sake_flow_pre(int32_t B, int32_t N, int32_t h_features, const float* h, const float* pos, float* h_aug,
                  float* x_aug, sake_stream_t stream);
int sake_flow_post(int32_t B, int32_t N, int32_t D, int32_t scale_hidden, int32_t direction, const float* x_aug_out,
                   const float* pos0, const float* y, const float* scale0_kernel, const float* scale0_bias,
                   const float* scale2_kernel, float* other, float* logdet, sake_stream_t stream);
int sake_flow_logprob(int32_t B, int32_t N, int32_t D, const float* x, const float* v, const float* logdet,
                      float* out, sake_stream_t stream);

/* One optimiser step over a flat fp32 parameter vector, the chain every training driver uses
 * (scripts/qm9/run.py:134-138): additive_weight_decay(wd) -> clip(max_delta, element-wise)
 * -> adam(lr, b1, b2, eps) with bias correction for `step` (1-based).  grad_scale multiplies the
 * gradients first (1/world_size after an all-reduce sum = lax.pmean, scripts/ani/run_gpu.py:130). */
int sake_adam_step(int64_t n, float* params, const float* grads, float* m, float* v, int32_t step,
                   float lr, float b1, float b2, float eps, float weight_decay, float max_delta,
                   float grad_scale, sake_stream_t stream);

/* Per-launch device timing of the dominant (x_mixing GEMM) kernels with CUDA events recorded on
 * the launch stream.  begin(capacity) arms it, collect() synchronises the recorded events and
 * returns how many records were written: ms[i] = duration, kind[i] = 1 mix-forward, 2 mix-backward (dX),
 * 3 mix-dW, 4 edge-forward, 5 edge-backward, 6 node-tail forward, 7 node-tail backward, 8 small weight-gradient
 * contractions + the partial-sum reduction of the layer, 9 per-node projections, 10 softmax + aggregate, 11 pair-record
 * reductions, 12 per-node projection backward, 13 softmax backward, 14 dense embedding / readout;
 * pairs[i] = atom pairs (kinds 6, 7, 9, 12, 14: atoms / rows) of the padded batch that launch covers. */
int sake_profile_begin(int32_t capacity);
int sake_profile_collect(float* ms, int32_t* kind, int64_t* pairs, int32_t capacity);

/* Diagnostic entry for the weight-gradient contraction kernel: out[xw, gw] += X[P, xw]^T G[P, gw]
 * (device pointers, fp32 row-major) on the tcgen05 engine `engine` (SAKE_ENGINE_TF32X3 / _BF16). */
int sake_selftest_xtg(int32_t engine, int64_t P, int32_t xw, int32_t gw, const float* X, const float* G,
                      float* out, sake_stream_t stream);

/* Diagnostic: wait-cycle counters of the forward mix kernel's MMA issuer (armed with the environment
 * variable SAKE_DEBUG_WSPLITS=7): [0] accumulator, [1] weight chunk, [2] pair chunk, [3] total, [4] tiles.
 * Reads and clears 8 counters (host memory). */
int sake_debug_counters(unsigned long long* out8);
/* Same switch, backward mix kernel: 16 counters (MMA issuer waits, epilogue phases, builder waits; see tc_mix.cu). */
int sake_debug_counters_bwd(unsigned long long* out16);

/* Number of CUDA kernels this library has launched in this process (diagnostic). */
unsigned long long sake_launch_count(void);

/* Self-test of the tcgen05 building blocks (descriptor encodings, swizzled operand images,
 * TMEM load layout) on the current device; returns 0 when every check passes.
 * max_abs_err[0] = 3xTF32 GEMM error, max_abs_err[1] = bf16 GEMM error (two floats, host memory). */
int sake_selftest_tcgen05(float* max_abs_err, sake_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAKE_B200_H_ */
