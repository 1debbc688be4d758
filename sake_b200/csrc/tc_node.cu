// tcgen05 engine for the per-node tail of DenseSAKELayer (sake/layers.py:85-92,123-131,142-151,184-186,
// 218-232): post_norm_mlp on the squared norms of the spatial sums, node_mlp on [h | h_e | h_comb] with the
// residual, and the velocity / position update.  Tile = 128 atoms, thread = atom (= TMEM lane), so every
// activation, bias add and reduction over features is thread-local; the Dense layers
//   post0 (256 -> 64), post2 (64 -> 64), node0 (384 -> 64), node2 (64 -> 64), vel0 (64 -> 64)
// run as 3xTF32 GEMMs D[128 x 64] += A[128 x 32] B[64 x 32]^T, one 32-wide K-chunk at a time: the thread
// writes its 32 input values into the (128-byte-swizzled, hi/lo split) A image, the weight chunk arrives by
// TMA from a pre-swizzled image of all five matrices (26 chunks x 16 KB, L2 resident), one thread issues the
// 12 MMAs.  Two accumulators of 64 TMEM columns alternate between consecutive layers.
// Replaces the CUDA-core kernel k_node_post (generic_fwd.cu) when H = 64, A = 4.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace sake {
using namespace tc;

constexpr int NT_TILE = 128;                 // atoms per CTA = builder threads (warps 0-3)
constexpr int NT_IMG = NT_TILE * 128;        // one split of the A chunk image (16 KB)
constexpr int NT_WCH = 2 * 64 * 128;         // one weight chunk: {hi, lo} x 64 output rows x 128 B (16 KB)
// chunk index of every matrix inside the weight image
constexpr int NTW_POST0 = 0, NTW_POST2 = 8, NTW_NODE0 = 10, NTW_NODE2 = 22, NTW_VEL0 = 24, NTW_CHUNKS = 26;
// backward (W^T g products): B rows = INPUT features of the layer, K = its outputs; 64-row blocks x 2 K-chunks
constexpr int NTB_VEL0T = 26, NTB_NODE2T = 28, NTB_NODE0T = 30, NTB_POST2T = 42, NTB_POST0T = 44, NTB_CHUNKS = 52;
constexpr int NT_VEC = 64 * 6 + 256;         // b_p1 b_p2 b_n1 b_n2 b_v1 vel2 | v_mixing

// weight image of k_tc_node_pre behind the chunks of the node tail: [K chunk (2)][{hi, lo}][256 output rows x 128 B] + 256 biases
constexpr size_t NTP_OFF = (size_t)NTB_CHUNKS * NT_WCH + 1024;
constexpr int NTP_WSPLIT = 256 * 128;                    // one split of one K chunk (32 KB)
constexpr int NTP_WBYTES = 4 * NTP_WSPLIT;               // 128 KB
// ... and of k_tc_node_pre_bwd: [K chunk of projection columns (8)][{hi, lo}][64 input-feature rows x 128 B]
constexpr size_t NTQ_OFF = NTP_OFF + NTP_WBYTES + 1024;
constexpr int NTQ_WSPLIT = 64 * 128;                     // one split of one K chunk (8 KB)
constexpr int NTQ_WBYTES = 8 * 2 * NTQ_WSPLIT;           // 128 KB
size_t tc_node_w_bytes() { return NTQ_OFF + NTQ_WBYTES; }

// weight image: chunk c = K rows [32c', 32c'+32) of its matrix W[in][64]; B operand rows = outputs
__global__ void k_node_w_prep(const SakeLayerParams p, uint8_t* __restrict__ img, int nchunks) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks * 64 * 8) return;
  const int chunk = t / (64 * 8), o = (t / 8) % 64, u = t % 8;
  float vals[4];
  if (chunk < NTW_CHUNKS) {                                // forward: element (o, k) = W[k][o]
    const float* W;
    int c0;
    if (chunk < NTW_POST2) { W = p.post0_kernel; c0 = chunk; }
    else if (chunk < NTW_NODE0) { W = p.post2_kernel; c0 = chunk - NTW_POST2; }
    else if (chunk < NTW_NODE2) { W = p.node0_kernel; c0 = chunk - NTW_NODE0; }
    else if (chunk < NTW_VEL0) { W = p.node2_kernel; c0 = chunk - NTW_NODE2; }
    else { W = p.vel0_kernel; c0 = chunk - NTW_VEL0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) vals[i] = W ? W[(size_t)(c0 * 32 + u * 4 + i) * 64 + o] : 0.f;
  } else {                                                 // backward: row r = input feature, element (r, k) = W[r][k]
    const float* W;
    int c0;
    if (chunk < NTB_NODE2T) { W = p.vel0_kernel; c0 = chunk - NTB_VEL0T; }
    else if (chunk < NTB_NODE0T) { W = p.node2_kernel; c0 = chunk - NTB_NODE2T; }
    else if (chunk < NTB_POST2T) { W = p.node0_kernel; c0 = chunk - NTB_NODE0T; }
    else if (chunk < NTB_POST0T) { W = p.post2_kernel; c0 = chunk - NTB_POST2T; }
    else { W = p.post0_kernel; c0 = chunk - NTB_POST0T; }
    const int blk = c0 >> 1, kc = c0 & 1;                  // 64-row block, K chunk
#pragma unroll
    for (int i = 0; i < 4; ++i) vals[i] = W ? W[(size_t)(blk * 64 + o) * 64 + kc * 32 + u * 4 + i] : 0.f;
  }
  uint8_t* base = img + (size_t)chunk * NT_WCH;
  const uint32_t off = sw128_offset((uint32_t)o, (uint32_t)u);
  float4 hi, lo;
  split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
  split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
  *reinterpret_cast<float4*>(base + off) = hi;
  *reinterpret_cast<float4*>(base + 64 * 128 + off) = lo;
}

__device__ __forceinline__ void nt_store_unit(uint8_t* img, int row, int u, const float* vals) {
  const uint32_t off = sw128_offset((uint32_t)row, (uint32_t)u);
  float4 hi, lo;
  split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
  split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
  *reinterpret_cast<float4*>(img + off) = hi;
  *reinterpret_cast<float4*>(img + NT_IMG + off) = lo;
}

// Per-atom buffers of the node kernels (stash, record, ssum) use the G8 layout (common.cuh): thread = atom, and the
// 16-byte unit u of row r sits at float4 index g8_row(r, U) + 8 u, so that a warp's access to one unit is four full
// lines (4 L1 wavefronts instead of the 32 of a thread-per-row walk over row-major rows: the node kernels were bound
// by exactly those wavefronts).  tc_xtg.cu reads the same layout (XtgArgs::x_tt / g_tt).
__device__ __forceinline__ float4* tt_base(float* buf, int units, int tile, int t) {
  return reinterpret_cast<float4*>(buf) + g8_row((long long)tile * 128 + t, units);
}
__device__ __forceinline__ const float4* tt_base(const float* buf, int units, int tile, int t) {
  return reinterpret_cast<const float4*>(buf) + g8_row((long long)tile * 128 + t, units);
}
#define TT(q, col) (q)[((col) >> 2) * G8S]            /* the float4 holding columns col .. col+3 (col % 4 == 0) */

// ---- per-node projections on the tensor cores ---------------------------------------------------------------
// proj[n] = h[n] @ [W_in[0:H] | W_in[H:2H] | W_1[0:H] | W_1[H:2H]] + [0 | b_in | 0 | b_1]   (layers.py:30,33-38; layout of
// nodeproj in common.cuh): one 3xTF32 GEMM [128 atoms x 64] x [64 x 256] per tile, thread = atom = TMEM lane, G8 rows out.
// Replaces the CUDA-core k_node_pre for the tcgen05 engines (cfg5: 109 -> ~20 us per layer call).
__global__ void k_node_pre_w_prep(const SakeLayerParams p, int H, int K, int Kp, uint8_t* __restrict__ img,
                                  float* __restrict__ bias, uint8_t* __restrict__ imgT) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 256) {
    float b = 0.f;
    if (t >= Kp && t < Kp + K) b = p.mlp_in_bias[t - Kp];
    else if (t >= 2 * Kp + H && t < 2 * Kp + 2 * H) b = p.mlp_out0_bias[t - 2 * Kp - H];
    bias[t] = b;
  }
  if (t >= 2 * 256 * 8) return;
  const int chunk = t / (256 * 8), o = (t / 8) % 256, u = t % 8;
  float vals[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = chunk * 32 + u * 4 + i;                // input feature
    float v = 0.f;
    if (o < Kp) { if (o < K) v = p.mlp_in_kernel[(size_t)k * K + o]; }
    else if (o < 2 * Kp) { if (o - Kp < K) v = p.mlp_in_kernel[(size_t)(H + k) * K + (o - Kp)]; }
    else if (o < 2 * Kp + H) v = p.mlp_out0_kernel[(size_t)k * H + (o - 2 * Kp)];
    else if (o < 2 * Kp + 2 * H) v = p.mlp_out0_kernel[(size_t)(H + k) * H + (o - 2 * Kp - H)];
    vals[i] = v;
  }
  uint8_t* base = img + (size_t)chunk * 2 * NTP_WSPLIT;
  const uint32_t off = sw128_offset((uint32_t)o, (uint32_t)u);
  float4 hi, lo;
  split_tf32(vals[0], hi.x, lo.x); split_tf32(vals[1], hi.y, lo.y);
  split_tf32(vals[2], hi.z, lo.z); split_tf32(vals[3], hi.w, lo.w);
  *reinterpret_cast<float4*>(base + off) = hi;
  *reinterpret_cast<float4*>(base + NTP_WSPLIT + off) = lo;
  // the same numbers for the backward GEMM dh = gproj W^T: rows = input features, K = projection columns.
  // This thread holds W[k = chunk*32 + 4u + i][o]; there they are element (row k, K index o), one scalar each.
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = chunk * 32 + u * 4 + i;
    float h1, l1;
    split_tf32(vals[i], h1, l1);
    uint8_t* b2 = imgT + (size_t)(o >> 5) * 2 * NTQ_WSPLIT + sw128_offset((uint32_t)k, (uint32_t)((o & 31) >> 2)) + (o & 3) * 4;
    *reinterpret_cast<float*>(b2) = h1;
    *reinterpret_cast<float*>(b2 + NTQ_WSPLIT) = l1;
  }
}

struct NodePreArgs { int R, NP; const RaggedHdr* hdr; const float* h; const uint8_t* wimg; const float* bias; float* proj; };

__global__ void __launch_bounds__(NT_TILE, 1) k_tc_node_pre(NodePreArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  uint8_t* sw = base;                                    // weight image, 128 KB
  uint8_t* imgs = base + NTP_WBYTES;                     // two A chunk images {hi, lo}: 2 x 32 KB
  float* sbias = reinterpret_cast<float*>(imgs + 4 * NT_IMG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 256);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int R = a.hdr ? a.hdr->R : a.R;
  const int ntiles = (R + NT_TILE - 1) / NT_TILE;
  if ((int)blockIdx.x >= ntiles) return;                 // ragged: the grid covers the padded worst case
  if (tid == 0) {
    mbar_init(bars, 1); mbar_init(bars + 1, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bars, NTP_WBYTES);
    bulk_g2s(sw, a.wimg, NTP_WBYTES / 2, bars);
    bulk_g2s(sw + NTP_WBYTES / 2, a.wimg + NTP_WBYTES / 2, NTP_WBYTES / 2, bars);
  }
  if (warp == 0) tmem_alloc<256>(tptr);
  for (int t = tid; t < 256; t += NT_TILE) sbias[t] = a.bias[t];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t img_u32 = smem_u32(imgs), sw_u32 = smem_u32(sw);
  constexpr uint32_t idesc = umma_idesc(2, 128, 256);
  const int units = a.NP >> 2;
  uint32_t ph = 0;
  bool w_ready = false;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile * NT_TILE + tid;
    const bool valid = n < R;
    const float4* hp = reinterpret_cast<const float4*>(a.h + (size_t)(valid ? n : 0) * 64);
    float4 hv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) hv[u] = __ldg(hp + u);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float vals[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w};
      nt_store_unit(imgs + (u >> 3) * 2 * NT_IMG, tid, u & 7, vals);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      if (!w_ready) mbar_wait(bars, 0);
      tc_fence_after();
      const int pp[3] = {0, 1, 0}, pw[3] = {0, 0, 1};
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int pr = 0; pr < 3; ++pr)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma<true>(tmem_base, umma_desc_k_sw128(img_u32 + (c * 2 + pp[pr]) * NT_IMG + ks * 32),
                       umma_desc_k_sw128(sw_u32 + (c * 2 + pw[pr]) * NTP_WSPLIT + ks * 32), idesc, (c | pr | ks) != 0);
      umma_commit(bars + 1);
    }
    w_ready = true;
    mbar_wait_warp(bars + 1, ph);
    ph ^= 1;
    tc_fence_after();
    float4* p4 = reinterpret_cast<float4*>(a.proj) + g8_row(valid ? n : 0, units);
#pragma unroll 1
    for (int cc = 0; cc * 8 < units; ++cc) {
      float v[32];
      tmem_ld32(lane_addr + cc * 32, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int col = cc * 32 + 4 * u;
          if (cc * 8 + u < units)
            p4[(cc * 8 + u) * G8S] = make_float4(v[4 * u] + sbias[col], v[4 * u + 1] + sbias[col + 1],
                                                 v[4 * u + 2] + sbias[col + 2], v[4 * u + 3] + sbias[col + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// dh[n] += gproj[n] @ W^T (cotangent of the per-node projections back to h; layers.py:30,33-38): [128 atoms x 256] x
// [256 x 64] in 3xTF32, K = the projection columns in eight chunks of 32 through two image buffers.  Replaces the
// CUDA-core k_node_pre_bwd for the tcgen05 engines (its weight gradients are X^T G problems: tc_node_pre_dw).
struct NodePreBwdArgs { int R, NP; const RaggedHdr* hdr; const float* gproj; const uint8_t* wimg; float* dh; };

__global__ void __launch_bounds__(NT_TILE, 1) k_tc_node_pre_bwd(NodePreBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  uint8_t* sw = base;                                    // weight image, 128 KB
  uint8_t* imgs = base + NTQ_WBYTES;                     // two A chunk images {hi, lo}: 2 x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(imgs + 4 * NT_IMG);   // [0] weights, [1], [2] image buffers
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int R = a.hdr ? a.hdr->R : a.R;
  const int ntiles = (R + NT_TILE - 1) / NT_TILE;
  if ((int)blockIdx.x >= ntiles) return;                 // ragged: the grid covers the padded worst case
  if (tid == 0) {
    mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bars, NTQ_WBYTES);
    bulk_g2s(sw, a.wimg, NTQ_WBYTES / 2, bars);
    bulk_g2s(sw + NTQ_WBYTES / 2, a.wimg + NTQ_WBYTES / 2, NTQ_WBYTES / 2, bars);
  }
  if (warp == 0) tmem_alloc<64>(tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t img_u32 = smem_u32(imgs), sw_u32 = smem_u32(sw);
  constexpr uint32_t idesc = umma_idesc(2, 128, 64);
  const int units = a.NP >> 2;
  uint32_t ph = 0;                                       // bit b: parity the next wait on image barrier b looks for
  bool w_ready = false;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile * NT_TILE + tid;
    const bool valid = n < R;
    const float4* gp = reinterpret_cast<const float4*>(a.gproj) + g8_row(valid ? n : 0, units);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      const int b = c & 1;
      float4 gv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) gv[u] = c * 8 + u < units ? __ldg(gp + (c * 8 + u) * G8S) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (c >= 2) { mbar_wait_warp(bars + 1 + b, (ph >> b) & 1u); ph ^= 1u << b; }     // the MMAs of chunk c - 2 have read image b
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float vals[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
        nt_store_unit(imgs + b * 2 * NT_IMG, tid, u, vals);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        if (!w_ready) mbar_wait(bars, 0);
        tc_fence_after();
        const int pp[3] = {0, 1, 0}, pw[3] = {0, 0, 1};
#pragma unroll
        for (int pr = 0; pr < 3; ++pr)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma<true>(tmem_base, umma_desc_k_sw128(img_u32 + (b * 2 + pp[pr]) * NT_IMG + ks * 32),
                       umma_desc_k_sw128(sw_u32 + (c * 2 + pw[pr]) * NTQ_WSPLIT + ks * 32), idesc, (c | pr | ks) != 0);
        umma_commit(bars + 1 + b);
      }
      w_ready = true;
    }
    mbar_wait_warp(bars + 1, ph & 1u); ph ^= 1u;         // chunks 6 and 7: every MMA of the tile is complete
    mbar_wait_warp(bars + 2, (ph >> 1) & 1u); ph ^= 2u;
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      float v[32];
      tmem_ld32(lane_addr + cc * 32, v);
      tmem_ld_wait();
      if (valid) {
        float4* o = reinterpret_cast<float4*>(a.dh + (size_t)n * 64 + cc * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 t4 = o[u];
          t4.x += v[4 * u]; t4.y += v[4 * u + 1]; t4.z += v[4 * u + 2]; t4.w += v[4 * u + 3];
          o[u] = t4;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<64>(tmem_base);
}

struct NodeFwdArgs {
  int R, N, update, has_v, spatial;
  const RaggedHdr* hdr; const int4* rowinfo;             // ragged batches (no mask): rows from hdr, n per row
  const float *h, *x, *v, *mask, *ssum, *he;
  const uint8_t* wimg;
  const float *b_p1, *b_p2, *b_n1, *b_n2, *b_v1, *vel2, *wv;
  float *h_out, *x_out, *v_out;
  float* stash;                                          // [R, NS_LD] activations kept for the backward kernel (or NULL)
};

// ---- pipelined chunk-GEMM rounds ------------------------------------------------------------------------------
// Both node kernels are chains of small GEMM rounds (one 32-wide K chunk: 12 MMAs), and at every size they are bound
// by the LATENCY of a round, not by throughput (cfg2: 37 CTAs of 128 threads on 148 SMs).  Round 1 ran a round as
// build image -> barrier -> MMA -> wait; here issue() does not wait: a round commits to a caller-chosen barrier and
// the kernel waits only where a result or an image buffer is actually needed, so the image build and the global
// loads of round c+1 run under the MMAs of round c (two image buffers, used alternately).  Weight chunks stream
// through a four-slot ring two rounds ahead; a slot is re-filled once the MMAs that read it have completed
// (tcgen05.commit on the slot's own barrier), which only the issuing thread waits for.
// The issuing thread is lane 0 of a fifth warp: the 128 builder threads (one per atom) hand an image over with
// a plain mbarrier arrive and go on to the next round — no block-wide barrier, and no builder thread spends its
// time on descriptors and the weight ring.  The rounds of both kernels form a fixed sequence, which the issuing
// warp walks on its own (mma_round), gated by the builders' arrives.
// SLOTS = 4 (look-ahead 2: consecutive rounds' MMAs overlap; 130 KB of shared memory, one CTA per SM) is the
// low-latency configuration for small batches; SLOTS = 2 (look-ahead 1; 99 KB) lets two CTAs share an SM and is
// used when there are more than two tiles per SM, where throughput matters and the second CTA hides the latency.
constexpr int NT_UBARS = 4;
constexpr int NT_MAXSLOTS = 4;
template <int SLOTS>
struct NodePipe {
  static constexpr int LA = SLOTS / 2;
  uint8_t* wring;
  uint64_t *wfull, *wfree, *ubar, *full;   // full[k]: image buffer k published (128 builder arrives)
  uint64_t* taken;                         // taken[k]: the issuing thread has seen that publication
  uint32_t npub;                           // builders: bit k = parity of the number of publications of buffer k so far
  uint32_t anypub;                         // builders: bit k = buffer k has been published at least once
  const uint8_t* wimg;
  int wpos;                      // rounds issued so far
  int c0, c1, c2;                // weight chunk of round wpos and wpos+1 (requested), wpos+2 (requested by the next issue); -1: none
  uint32_t uph;                  // bit i: phase parity the next wait on user barrier i looks for

  __device__ __forceinline__ void init_barriers(int builders) {      // one thread
    for (int i = 0; i < SLOTS; ++i) { mbar_init(wfull + i, 1); mbar_init(wfree + i, 1); }
    for (int i = 0; i < NT_UBARS; ++i) mbar_init(ubar + i, 1);
    mbar_init(full, builders); mbar_init(full + 1, builders);
    mbar_init(taken, 1); mbar_init(taken + 1, 1);
    fence_barrier_init();
  }
  __device__ __forceinline__ void request(int chunk, int slot) {   // one thread
    mbar_arrive_expect_tx(wfull + slot, NT_WCH);
    bulk_g2s(wring + slot * NT_WCH, wimg + (size_t)chunk * NT_WCH, NT_WCH, wfull + slot);
  }
  // builders: publish image buffer k (this thread's rows are written).  Non-blocking, except that a buffer is not
  // published a second time before the issuing thread has SEEN the first publication: an mbarrier wait can only
  // tell adjacent phases apart, and the block loops of the backward kernel publish the same (read-only) images
  // for consecutive GEMMs without waiting for MMA completion in between.
  __device__ __forceinline__ void publish(int k) {
    if ((anypub >> k) & 1u) mbar_wait_warp(taken + k, ((npub >> k) & 1u) ^ 1u);   // previous publication of k was taken
    anypub |= 1u << k;
    npub ^= 1u << k;
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(full + k);
  }
  // issuing thread: round r = the r-th image handed over (buffer r & 1); 12 MMAs against the round's weight chunk.
  // ub >= 0: the MMAs issued so far arrive on user barrier ub when complete (at most one un-waited commit per barrier).
  template <class Next>
  __device__ __forceinline__ void mma_round(uint32_t img_u32, uint32_t dcol, bool first, int ub, Next next) {
    const int slot = wpos & (SLOTS - 1);
    const int creq = LA == 2 ? c2 : c1;                // chunk of round wpos + LA
    if (creq >= 0) {
      const int s2 = (wpos + LA) & (SLOTS - 1);        // last read by round wpos + LA - SLOTS
      if (wpos + LA >= SLOTS) mbar_wait(wfree + s2, ((wpos + LA - SLOTS) / SLOTS) & 1);
      request(creq, s2);
    }
    mbar_wait(wfull + slot, (wpos / SLOTS) & 1);
    mbar_wait(full + (wpos & 1), (wpos >> 1) & 1);
    mbar_arrive(taken + (wpos & 1));
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc(2, 128, 64);
    const uint32_t wb = smem_u32(wring + slot * NT_WCH);
    const int pp[3] = {0, 1, 0}, pw[3] = {0, 0, 1};
#pragma unroll
    for (int pr = 0; pr < 3; ++pr)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma<true>(dcol, umma_desc_k_sw128(img_u32 + pp[pr] * NT_IMG + ks * 32),
                   umma_desc_k_sw128(wb + pw[pr] * (64 * 128) + ks * 32), idesc, !(first && pr == 0 && ks == 0));
    umma_commit(wfree + slot);
    if (ub >= 0) umma_commit(ubar + ub);
    ++wpos;
    c0 = c1; c1 = c2; c2 = c2 >= 0 ? next(c2) : -1;
  }
  __device__ __forceinline__ void wait(int ub) {
    mbar_wait_warp(ubar + ub, (uph >> ub) & 1u);
    uph ^= 1u << ub;
    tc_fence_after();
  }
};

// HALVES = 2: two threads per atom.  Threads t and t + 128 share TMEM lane t (warps w and w + 4 address the same lane
// quarter) and split every 32-wide chunk by 16-byte units: half hf owns units 4 hf .. 4 hf + 3, i.e. 16 of the 32
// columns — of the images it writes, of the accumulator columns it reads back, of the global rows it loads and stores.
// Everything in the chain is per column, so the halves never exchange data except the three reductions over features
// (velocity gate logit, v_mixing sums), which meet in shared memory at the end.  A round is bound by the thread-side
// work of building an image with one warp per scheduler; two threads per atom halve it.  HALVES = 1 (thread = atom)
// is kept for the two-CTAs-per-SM configuration (SLOTS = 2), whose register budget has no room for 288 threads.
template <int SLOTS, int HALVES>
__global__ void __launch_bounds__(NT_TILE * HALVES + 32, SLOTS == 2 ? 2 : 1) k_tc_node_post(NodeFwdArgs a) {
  constexpr int NTH = NT_TILE * HALVES + 32;             // builders + the issuing warp
  constexpr int UH = 8 / HALVES, CH = 32 / HALVES;       // units / columns of a chunk owned by one thread
  const int nrows_real = a.hdr ? a.hdr->R : a.R;
  if ((int)blockIdx.x * NT_TILE >= nrows_real) return;   // ragged: the grid covers the padded worst case
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  uint8_t* imgs = base;                                  // two A chunk images {hi, lo}: 2 x 32 KB
  uint8_t* wring = base + 4 * NT_IMG;                    // SLOTS weight chunk slots of 16 KB
  float* svec = reinterpret_cast<float*>(wring + SLOTS * NT_WCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(svec + NT_VEC);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2 * NT_MAXSLOTS + NT_UBARS + 4);
  float4* xch = reinterpret_cast<float4*>(tptr + 4);     // [NT_TILE] partial sums of half 1 -> half 0 (HALVES = 2)
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool upd = a.update != 0, hv = a.has_v != 0, spatial = a.spatial != 0, uv = upd && hv;
  const float* vsrc[7] = {a.b_p1, a.b_p2, a.b_n1, a.b_n2, a.b_v1, a.vel2, a.wv};
  for (int t = tid; t < NT_VEC; t += NTH) {
    const int k = t < 384 ? t / 64 : 6, i = t < 384 ? t % 64 : t - 384;
    svec[t] = vsrc[k] ? vsrc[k][i] : 0.f;
  }
  const int n_chunks = uv ? NTW_CHUNKS : NTW_VEL0;       // the forward consumes chunks 0 .. n_chunks-1 in order
  auto next_w = [&](int k) { return k + 1 < n_chunks ? k + 1 : -1; };
  NodePipe<SLOTS> pp;
  pp.wring = wring; pp.wfull = bars; pp.wfree = bars + NT_MAXSLOTS; pp.ubar = bars + 2 * NT_MAXSLOTS; pp.wimg = a.wimg;
  pp.full = bars + 2 * NT_MAXSLOTS + NT_UBARS; pp.taken = pp.full + 2; pp.npub = 0; pp.anypub = 0;
  pp.wpos = 0; pp.c0 = 0; pp.c1 = 1; pp.c2 = 2; pp.uph = 0;
  if (tid == NT_TILE * HALVES) {
    pp.init_barriers(NT_TILE * HALVES);
    pp.request(0, 0);
    if (SLOTS == 4) pp.request(1, 1);
  }
  if (warp == 0) tmem_alloc<128>(tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t img_u32[2] = {smem_u32(imgs), smem_u32(imgs + 2 * NT_IMG)};
  uint8_t* const img_p[2] = {imgs, imgs + 2 * NT_IMG};
  const float *s_bp1 = svec, *s_bp2 = svec + 64, *s_bn1 = svec + 128, *s_bn2 = svec + 192, *s_bv1 = svec + 256,
              *s_vel2 = svec + 320, *s_wv = svec + 384;
  const uint32_t D0 = tmem_base, D1 = tmem_base + 64;

  if (warp == 4 * HALVES) {
    // ---------------- issuing warp: the fixed sequence of rounds — post0 (8, -> D0), post2 (2, -> D1), node0 (12, -> D0),
    // node2 (2, -> D1), vel0 (2, -> D0, with a velocity gate); image buffer and user barrier = round parity
    if ((tid & 31) == 0) {
      for (int r = 0; r < n_chunks; ++r) {
        const bool to_d1 = (r >= 8 && r < 10) || (r >= 22 && r < 24);
        const bool first = r == 0 || r == 8 || r == 10 || r == 22 || r == 24;
        pp.mma_round(img_u32[r & 1], to_d1 ? D1 : D0, first, r & 1, next_w);
      }
    }
  } else {
  const int t = tid & (NT_TILE - 1), hf = HALVES == 2 ? tid >> 7 : 0;      // atom (= TMEM lane) and column half
  const int u0 = hf * UH, c0h = hf * CH;                                   // first unit / column of a chunk owned here
  const int n = blockIdx.x * NT_TILE + t;
  const bool valid = n < nrows_real;
  const size_t row = valid ? (size_t)n : 0;        // idle lanes read atom 0 (valid memory) and never store
  float den = (float)a.N, den2 = (float)a.N;
  if (a.rowinfo) den = den2 = (float)__ldg(a.rowinfo + row).y;   // unpadded molecule: mean over its own n atoms
  if (a.mask && valid) {
    float ms = 0.f;
    const float* mr = a.mask + row * a.N;
    for (int j = 0; j < a.N; ++j) ms += mr[j];
    den = ms + 1e-8f;      // layers.py:123
    den2 = ms + 1e-10f;    // layers.py:221
  }
  const float inv_den = 1.0f / den;
  float4* ns = (a.stash && valid) ? tt_base(a.stash, NS_LD / 4, blockIdx.x, t) : nullptr;   // this atom's fwd -> bwd stash
  // Image buffer k = round parity; user barrier k tracks the last round that read image k.
  // busy: bit k set while a round on image k is un-waited.
  uint32_t busy = 0;
  auto acquire = [&](int k) { if (busy & (1u << k)) { pp.wait(k); busy &= ~(1u << k); } };
  auto launch = [&](int k, uint32_t, bool) { pp.publish(k); busy |= 1u << k; };   // the issuing warp knows the rest
  auto drain = [&]() { acquire(0); acquire(1); };       // every MMA issued so far is complete (results readable)
  auto ld_chunk = [&](uint32_t col, float* v) {         // this thread's CH columns of a 32-column accumulator chunk
    if constexpr (HALVES == 2) tmem_ld16(lane_addr + col + c0h, v); else tmem_ld32(lane_addr + col, v);
    tmem_ld_wait();
  };

  // ---------------- post0: nrm[c] = sum_d (ssum[c][d] / den)^2  (layers.py:123-129), K = 256 -> D0
  // the row loads of chunk c+1 are issued before the hand-off of chunk c: their latency hides under its MMAs
  // ssum is in the G8 layout (written so by k_tc_mix_fwd): unit (c'/4)*3 + d holds component d of four coefficients
  float dv0 = 0.f, dv1 = 0.f, dv2 = 0.f;
  float4 sreg[3 * UH];
  const float4* ssq = tt_base(a.ssum, 192, blockIdx.x, t) + 3 * u0 * G8S;
  {
#pragma unroll
    for (int q = 0; q < 3 * UH; ++q) sreg[q] = __ldg(ssq + q * G8S);
  }
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    const int k = c & 1;
    acquire(k);
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      const float4 t0 = sreg[u * 3], t1 = sreg[u * 3 + 1], t2 = sreg[u * 3 + 2];
      const float s[12] = {t0.x, t1.x, t2.x, t0.y, t1.y, t2.y, t0.z, t1.z, t2.z, t0.w, t1.w, t2.w};
      float vals[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a0 = s[3 * i] * inv_den, a1 = s[3 * i + 1] * inv_den, a2 = s[3 * i + 2] * inv_den;
        vals[i] = a0 * a0 + a1 * a1 + a2 * a2;
        const float w = s_wv[c * 32 + (u0 + u) * 4 + i];
        dv0 = fmaf(w, s[3 * i], dv0); dv1 = fmaf(w, s[3 * i + 1], dv1); dv2 = fmaf(w, s[3 * i + 2], dv2);
      }
      nt_store_unit(img_p[k], t, u0 + u, vals);
    }
    if (c + 1 < 8) {
#pragma unroll
      for (int q = 0; q < 3 * UH; ++q) sreg[q] = __ldg(ssq + ((c + 1) * 24 + q) * G8S);
    }
    launch(k, D0, c == 0);
  }
  // node0's first inputs (h, he: 10 chunks) start loading now
  const float4* hp = reinterpret_cast<const float4*>(a.h + row * 64) + u0;
  const float4* hep = reinterpret_cast<const float4*>(a.he) + g8_row((long long)row, 64) + u0 * G8S;   // G8: unit q at hep[q * 8]
  float4 creg[UH];
#pragma unroll
  for (int u = 0; u < UH; ++u) creg[u] = __ldg(hp + u);
  drain();
  // ---------------- post2: h_p1 = silu(tp1 + b)  -> D1
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    float v[CH];
    ld_chunk(c * 32, v);
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      float vals[4], dvs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float z = v[4 * u + i] + s_bp1[c * 32 + c0h + 4 * u + i]; vals[i] = fsilu_(z); dvs[i] = fdsilu_(z); }
      if (ns) {
        TT(ns, NS_D + c * 32 + c0h + 4 * u) = make_float4(dvs[0], dvs[1], dvs[2], dvs[3]);
        TT(ns, NS_HP1 + c * 32 + c0h + 4 * u) = make_float4(vals[0], vals[1], vals[2], vals[3]);
      }
      nt_store_unit(img_p[c], t, u0 + u, vals);
    }
    launch(c, D1, c == 0);
  }
  // ---------------- node0 over [h | h_e | h_comb] (layers.py:142-151) -> D0 (free: post2 only read it through registers)
#pragma unroll 1
  for (int c = 0; c < 12; ++c) {
    const int k = c & 1;
    if (c == 10) drain();                                // h_comb needs the post2 result in D1
    acquire(k);
    if (c < 10) {
#pragma unroll
      for (int u = 0; u < UH; ++u) {
        const float vals[4] = {creg[u].x, creg[u].y, creg[u].z, creg[u].w};
        nt_store_unit(img_p[k], t, u0 + u, vals);
      }
      if (c + 1 < 10) {
        const float4* src = c + 1 < 2 ? hp + (c + 1) * 8 : hep + (c + 1 - 2) * 8 * G8S;
        const int us = c + 1 < 2 ? 1 : G8S;
#pragma unroll
        for (int u = 0; u < UH; ++u) creg[u] = __ldg(src + u * us);
      }
    } else {
      float v[CH];
      ld_chunk(64 + (c - 10) * 32, v);
#pragma unroll
      for (int u = 0; u < UH; ++u) {
        float vals[4], dvs[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float z = v[4 * u + i] + s_bp2[(c - 10) * 32 + c0h + 4 * u + i];
          vals[i] = spatial ? fsilu_(z) : 0.f;
          dvs[i] = spatial ? fdsilu_(z) : 0.f;
        }
        if (ns) {
          TT(ns, NS_D + 64 + (c - 10) * 32 + c0h + 4 * u) = make_float4(dvs[0], dvs[1], dvs[2], dvs[3]);
          TT(ns, NS_HCOMB + (c - 10) * 32 + c0h + 4 * u) = make_float4(vals[0], vals[1], vals[2], vals[3]);
        }
        nt_store_unit(img_p[k], t, u0 + u, vals);
      }
    }
    // D0 was last written by post0 and read (into registers) by post2: both long complete
    launch(k, D0, c == 0);
  }
  drain();
  // ---------------- node2: n1 = silu(t1 + b) -> D1
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    float v[CH];
    ld_chunk(c * 32, v);
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      float vals[4], dvs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float z = v[4 * u + i] + s_bn1[c * 32 + c0h + 4 * u + i]; vals[i] = fsilu_(z); dvs[i] = fdsilu_(z); }
      if (ns) {
        TT(ns, NS_D + 128 + c * 32 + c0h + 4 * u) = make_float4(dvs[0], dvs[1], dvs[2], dvs[3]);
        TT(ns, NS_N1 + c * 32 + c0h + 4 * u) = make_float4(vals[0], vals[1], vals[2], vals[3]);
      }
      nt_store_unit(img_p[c], t, u0 + u, vals);
    }
    launch(c, D1, c == 0);
  }
  drain();
  // ---------------- h' = h + silu(t2 + b)  (layers.py:150); velocity gate MLP on h' (layers.py:184-186) -> D0
  float y = 0.f;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    float v[CH];
    ld_chunk(64 + c * 32, v);
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      const float4 h4 = __ldg(hp + c * 8 + u);
      const float hin[4] = {h4.x, h4.y, h4.z, h4.w};
      float vals[4], dvs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float z = v[4 * u + i] + s_bn2[c * 32 + c0h + 4 * u + i]; vals[i] = hin[i] + fsilu_(z); dvs[i] = fdsilu_(z); }
      if (ns) {
        TT(ns, NS_D + 192 + c * 32 + c0h + 4 * u) = make_float4(dvs[0], dvs[1], dvs[2], dvs[3]);
        TT(ns, NS_HOUT + c * 32 + c0h + 4 * u) = make_float4(vals[0], vals[1], vals[2], vals[3]);
      }
      if (valid) *reinterpret_cast<float4*>(a.h_out + row * 64 + c * 32 + c0h + 4 * u) = make_float4(vals[0], vals[1], vals[2], vals[3]);
      if (uv) nt_store_unit(img_p[c], t, u0 + u, vals);
    }
    if (uv) launch(c, D0, c == 0);
  }
  if (uv) {
    drain();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[CH];
      ld_chunk(c * 32, v);
#pragma unroll
      for (int u = 0; u < UH; ++u) {
        float av[4], dvs[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = 4 * u + i;
          const float z = v[k] + s_bv1[c * 32 + c0h + k];
          av[i] = fsilu_(z); dvs[i] = fdsilu_(z);
          y = fmaf(av[i], s_vel2[c * 32 + c0h + k], y);
        }
        if (ns) {
          TT(ns, NS_D + 256 + c * 32 + c0h + 4 * u) = make_float4(dvs[0], dvs[1], dvs[2], dvs[3]);
          TT(ns, NS_AV + c * 32 + c0h + 4 * u) = make_float4(av[0], av[1], av[2], av[3]);
        }
      }
    }
  }
  if constexpr (HALVES == 2) {                           // the halves' sums over features meet: half 1 -> half 0
    if (hf == 1) xch[t] = make_float4(dv0, dv1, dv2, y);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) { const float4 o = xch[t]; dv0 += o.x; dv1 += o.y; dv2 += o.z; y += o.w; }
  }
  if (hf == 0) {
  if (ns) reinterpret_cast<float*>(&TT(ns, NS_Y))[0] = y;
  // ---------------- velocity / position update (layers.py:218-232)
  if (valid) {
    const float* xr = a.x + row * 3;
    if (!upd) {
      a.x_out[row * 3] = xr[0]; a.x_out[row * 3 + 1] = xr[1]; a.x_out[row * 3 + 2] = xr[2];
      if (a.v && a.v_out) { a.v_out[row * 3] = a.v[row * 3]; a.v_out[row * 3 + 1] = a.v[row * 3 + 1]; a.v_out[row * 3 + 2] = a.v[row * 3 + 2]; }
    } else {
      float vn0 = spatial ? dv0 / den2 : 0.f, vn1 = spatial ? dv1 / den2 : 0.f, vn2 = spatial ? dv2 / den2 : 0.f;
      if (hv) {
        const float gate = 2.0f * fsigmoid_(y);
        vn0 = fmaf(gate, a.v[row * 3], vn0); vn1 = fmaf(gate, a.v[row * 3 + 1], vn1); vn2 = fmaf(gate, a.v[row * 3 + 2], vn2);
      }
      a.v_out[row * 3] = vn0; a.v_out[row * 3 + 1] = vn1; a.v_out[row * 3 + 2] = vn2;
      a.x_out[row * 3] = xr[0] + vn0; a.x_out[row * 3 + 1] = xr[1] + vn1; a.x_out[row * 3 + 2] = xr[2] + vn2;
    }
  }
  }
  }   // builders
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

// =================================================================================================
// backward of the per-node tail (VJP of everything k_tc_node_post computes; replaces k_node_post_bwd):
//   recompute the forward (same chunk GEMMs), stash the four/five silu' vectors of the atom in a global
//   scratch row, then walk the chain back: velocity gate -> node_mlp -> [dh | g_he | g_hcomb] -> post_norm_mlp
//   -> g_nrm -> T = d(loss)/d(ssum) (+ the v_mixing path), row maximum of |T| for the fp16-split mix kernel.
//   The W^T g products use the row-block images NTB_*: D[128 x 64] = G[128 x 64] * W[64-row block][64]^T.
//   Training additionally writes the per-atom record (nbuf, layout NB_* of generic_bwd.cu) that the batched
//   weight-gradient contractions read.
// =================================================================================================
enum { NB_N1 = 0, NB_GT2 = 64, NB_CAT = 128, NB_GT1 = 512, NB_HP1 = 576, NB_GTP2 = 640, NB_GTP1 = 704, NB_NRM = 768,
       NB_HOUT = 1024, NB_GTV = 1088, NB_AV = 1152, NB_GY = 1216, NB_LD = 1232 };

struct NodeBwdArgs {
  int R, N, update, has_v, spatial;
  const RaggedHdr* hdr; const int4* rowinfo;
  const float *h, *v, *mask, *ssum, *he, *dh_out, *dx_out, *dv_out;
  const uint8_t* wimg;
  const float *b_p1, *b_p2, *b_n1, *b_n2, *b_v1, *vel2, *wv;
  float *dh, *dx, *dv, *T, *ghe, *tmax;
  const float* stash;                        // saved.nstash [R, NS_LD] written by k_tc_node_post
  float *qv, *nbuf;                          // g_dv / den2 [R,4] (training, v_mixing grad); record or NULL
};

template <int SLOTS, int HALVES>
__global__ void __launch_bounds__(NT_TILE * HALVES + 32, SLOTS == 2 ? 2 : 1) k_tc_node_post_bwd(NodeBwdArgs a) {
  constexpr int NTH = NT_TILE * HALVES + 32;
  constexpr int UH = 8 / HALVES, CH = 32 / HALVES;       // units / columns of a 32-wide chunk owned by one thread (see the forward kernel)
  const int nrows_real = a.hdr ? a.hdr->R : a.R;
  if ((int)blockIdx.x * NT_TILE >= nrows_real) return;   // ragged: the grid covers the padded worst case
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = align1024_shared(smem_raw);
  uint8_t* imgA = base;                                  // A image, K chunk 0: 32 KB
  uint8_t* imgB = base + 2 * NT_IMG;                     // A image, K chunk 1: 32 KB
  uint8_t* wring = imgB + 2 * NT_IMG;                    // SLOTS weight chunk slots of 16 KB
  float* svec = reinterpret_cast<float*>(wring + SLOTS * NT_WCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(svec + NT_VEC);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2 * NT_MAXSLOTS + NT_UBARS + 4);
  float* xch = reinterpret_cast<float*>(tptr + 4);       // [NT_TILE] row maximum of half 1 -> half 0 (HALVES = 2)
  const int tid = threadIdx.x, warp = tid >> 5;
  const float* vsrc[7] = {a.b_p1, a.b_p2, a.b_n1, a.b_n2, a.b_v1, a.vel2, a.wv};
  for (int t = tid; t < NT_VEC; t += NTH) {
    const int k = t < 384 ? t / 64 : 6, i = t < 384 ? t % 64 : t - 384;
    svec[t] = vsrc[k] ? vsrc[k][i] : 0.f;
  }
  const bool upd = a.update != 0, hv = a.has_v != 0, spatial = a.spatial != 0, uv = upd && hv;
  // order in which the weight chunks are consumed (the image holds them in this order, with optional parts):
  // vel0^T (with a velocity gate), node2^T, node0^T, then post2^T and post0^T (with spatial attention)
  auto next_w = [&](int k) {
    int nx = k + 1;
    if (nx == NTB_POST2T && !spatial) nx = -1;
    if (nx >= NTB_CHUNKS) nx = -1;
    return nx;
  };
  NodePipe<SLOTS> pp;
  pp.wring = wring; pp.wfull = bars; pp.wfree = bars + NT_MAXSLOTS; pp.ubar = bars + 2 * NT_MAXSLOTS; pp.wimg = a.wimg;
  pp.full = bars + 2 * NT_MAXSLOTS + NT_UBARS; pp.taken = pp.full + 2; pp.npub = 0; pp.anypub = 0;
  pp.wpos = 0; pp.uph = 0;
  pp.c0 = uv ? NTB_VEL0T : NTB_NODE2T; pp.c1 = next_w(pp.c0); pp.c2 = next_w(pp.c1);
  if (tid == NT_TILE * HALVES) {
    pp.init_barriers(NT_TILE * HALVES);
    pp.request(pp.c0, 0);
    if (SLOTS == 4) pp.request(pp.c1, 1);
  }
  if (warp == 0) tmem_alloc<128>(tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t imgA_u32 = smem_u32(imgA), imgB_u32 = smem_u32(imgB);
  const float *s_vel2 = svec + 320, *s_wv = svec + 384;
  const uint32_t D0 = tmem_base, D1 = tmem_base + 64;
  // One GEMM of the chain = the pair (imgA, imgB) = K 64 against two weight chunks; it commits to user barrier ub.
  if (warp == 4 * HALVES) {
    // ---------------- issuing warp: vel0^T (with a gate), node2^T, six blocks of node0^T, then post2^T and four blocks
    // of post0^T (with spatial attention); blocks alternate D0 / D1 and user barriers 0 / 1
    if ((tid & 31) == 0) {
      auto pair = [&](uint32_t dcol, int ub) {
        pp.mma_round(imgA_u32, dcol, true, -1, next_w);
        pp.mma_round(imgB_u32, dcol, false, ub, next_w);
      };
      if (uv) pair(D0, 0);
      pair(D1, 0);
      for (int b = 0; b < 6; ++b) pair((b & 1) ? D1 : D0, b & 1);
      if (spatial) {
        pair(D0, 0);
        for (int b = 0; b < 4; ++b) pair((b & 1) ? D1 : D0, b & 1);
      }
    }
  } else {
  auto run_pair = [&](uint32_t, int) { pp.publish(0); pp.publish(1); };   // the issuing warp knows destination and barrier

  const int t = tid & (NT_TILE - 1), hf = HALVES == 2 ? tid >> 7 : 0;      // atom (= TMEM lane) and column half
  const int u0 = hf * UH, c0h = hf * CH;                                   // first unit / column of a chunk owned here
  const int n = blockIdx.x * NT_TILE + t;
  const bool valid = n < nrows_real;
  const size_t row = valid ? (size_t)n : 0;        // idle lanes read atom 0 (valid memory) and never store
  float den = (float)a.N, den2 = (float)a.N;
  if (a.rowinfo) den = den2 = (float)__ldg(a.rowinfo + row).y;   // unpadded molecule: mean over its own n atoms
  if (a.mask && valid) {
    float ms = 0.f;
    const float* mr = a.mask + row * a.N;
    for (int j = 0; j < a.N; ++j) ms += mr[j];
    den = ms + 1e-8f;
    den2 = ms + 1e-10f;
  }
  const float inv_den = 1.0f / den;
  float4* nb = a.nbuf ? tt_base(a.nbuf, NB_LD / 4, blockIdx.x, t) : nullptr;   // this atom's record (G8 layout)
  const bool rec = nb != nullptr && valid;
  auto ld_chunk = [&](uint32_t col, float* v) {         // this thread's CH columns of a 32-column accumulator chunk
    if constexpr (HALVES == 2) tmem_ld16(lane_addr + col + c0h, v); else tmem_ld32(lane_addr + col, v);
    tmem_ld_wait();
  };
  uint8_t* const img_c[2] = {imgA, imgB};

  // =============================== forward activations: kept by k_tc_node_post ===============================
  // (round 1 recomputed the forward here: 26 of the kernel's 52 serial chunk-GEMM rounds, on a kernel that is
  // latency-bound at every size; the forward kernel now leaves silu' of the five hidden layers, the inputs of the
  // Dense layers and the gate logit in saved.nstash, 2.6 KB per atom)
  const float4* nd = tt_base(a.stash, NS_LD / 4, blockIdx.x, valid ? t : 0);
  // The X operands of the batched weight-gradient contractions (the inputs of every Dense layer) are read by
  // tc_node_dw where they already live: h, saved.he and the stash fields.  Only the cotangents (and nrm, which
  // falls out of the T loop below) are written to the record: 2.3 KB per atom instead of 4.9 KB + 5.3 KB of
  // thread-per-row copies in front of the chain.
  const float y = reinterpret_cast<const float*>(&TT(nd, NS_Y))[0];
  // =============================== velocity / position update backward (layers.py:226-232) ===============================
  // (scalars: both halves compute them, half 0 stores)
  float gdv0 = 0.f, gdv1 = 0.f, gdv2 = 0.f, gy = 0.f;
  if (valid) {
    const float dxo0 = a.dx_out ? a.dx_out[row * 3] : 0.f, dxo1 = a.dx_out ? a.dx_out[row * 3 + 1] : 0.f,
                dxo2 = a.dx_out ? a.dx_out[row * 3 + 2] : 0.f;
    const float dvo0 = a.dv_out ? a.dv_out[row * 3] : 0.f, dvo1 = a.dv_out ? a.dv_out[row * 3 + 1] : 0.f,
                dvo2 = a.dv_out ? a.dv_out[row * 3 + 2] : 0.f;
    if (hf == 0) { a.dx[row * 3] = dxo0; a.dx[row * 3 + 1] = dxo1; a.dx[row * 3 + 2] = dxo2; }          // x' = x + v'
    if (upd) {
      gdv0 = dvo0 + dxo0; gdv1 = dvo1 + dxo1; gdv2 = dvo2 + dxo2;                        // cotangent of v'
      if (hv) {
        const float gt = 2.0f * fsigmoid_(y);
        const float* vv = a.v + row * 3;
        const float ggate = gdv0 * vv[0] + gdv1 * vv[1] + gdv2 * vv[2];
        gy = ggate * gt * (1.0f - 0.5f * gt);
        if (a.dv && hf == 0) { a.dv[row * 3] = gt * gdv0; a.dv[row * 3 + 1] = gt * gdv1; a.dv[row * 3 + 2] = gt * gdv2; }
      }
    } else if (a.dv && hv && hf == 0) {
      a.dv[row * 3] = dvo0; a.dv[row * 3 + 1] = dvo1; a.dv[row * 3 + 2] = dvo2;          // v passes through
    }
    if (rec && uv && hf == 0) reinterpret_cast<float*>(&TT(nb, NB_GY))[0] = gy;
    if (a.qv && hf == 0) *reinterpret_cast<float4*>(a.qv + row * 4) = make_float4(gdv0 / den2, gdv1 / den2, gdv2 / den2, 0.f);
  }
  // =============================== backward chain ===============================
  // every 64-wide vector lives as two chunks of 32 columns; this thread holds columns c*32 + c0h .. + CH - 1 of chunk c
  float gho[2][CH];                                      // cotangent of h' (dh_out + velocity-gate path)
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.dh_out + row * 64) + c * 8 + u0 + u);
      gho[c][4 * u] = t4.x; gho[c][4 * u + 1] = t4.y; gho[c][4 * u + 2] = t4.z; gho[c][4 * u + 3] = t4.w;
    }
  // write CH floats (this thread's part of chunk c) of a record field / a row-major row
  auto rec_store = [&](int col0, int c, const float* v) {
#pragma unroll
    for (int u = 0; u < UH; ++u) TT(nb, col0 + c * 32 + c0h + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
  };
  auto row_store = [&](float* dst, int c, const float* v) {
#pragma unroll
    for (int u = 0; u < UH; ++u)
      *reinterpret_cast<float4*>(dst + c * 32 + c0h + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
  };
  auto img_store = [&](int c, const float* v) {
#pragma unroll
    for (int u = 0; u < UH; ++u) nt_store_unit(img_c[c], t, u0 + u, v + 4 * u);
  };
  if (uv) {
    // g_tv = vel2 * g_y * silu'(tv);  g_h' += g_tv W_v1^T
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float gtv[CH];
#pragma unroll
      for (int u = 0; u < UH; ++u) {
        const float4 d4 = TT(nd, 256 + c * 32 + c0h + 4 * u);
        const int k0 = c * 32 + c0h + 4 * u;
        gtv[4 * u] = s_vel2[k0] * gy * d4.x; gtv[4 * u + 1] = s_vel2[k0 + 1] * gy * d4.y;
        gtv[4 * u + 2] = s_vel2[k0 + 2] * gy * d4.z; gtv[4 * u + 3] = s_vel2[k0 + 3] * gy * d4.w;
      }
      if (rec) rec_store(NB_GTV, c, gtv);
      img_store(c, gtv);
    }
    run_pair(D0, 0);
    pp.wait(0);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[CH];
      ld_chunk(c * 32, v);
#pragma unroll
      for (int k = 0; k < CH; ++k) gho[c][k] += v[k];
    }
  }
  // g_t2 = g_h' * silu'(t2);  g_t1 = (g_t2 W_n2^T) * silu'(t1)
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float g2[CH];
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      const float4 d4 = TT(nd, 192 + c * 32 + c0h + 4 * u);
      g2[4 * u] = gho[c][4 * u] * d4.x; g2[4 * u + 1] = gho[c][4 * u + 1] * d4.y;
      g2[4 * u + 2] = gho[c][4 * u + 2] * d4.z; g2[4 * u + 3] = gho[c][4 * u + 3] * d4.w;
    }
    if (rec) rec_store(NB_GT2, c, g2);
    img_store(c, g2);
  }
  run_pair(D1, 0);
  pp.wait(0);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    float v[CH];
    ld_chunk(64 + c * 32, v);
#pragma unroll
    for (int u = 0; u < UH; ++u) {
      const float4 d4 = TT(nd, 128 + c * 32 + c0h + 4 * u);
      v[4 * u] *= d4.x; v[4 * u + 1] *= d4.y; v[4 * u + 2] *= d4.z; v[4 * u + 3] *= d4.w;
    }
    if (rec) rec_store(NB_GT1, c, v);
    img_store(c, v);
  }
  // g_cat = g_t1 W_n1^T : six 64-row blocks [dh | g_he (4 blocks) | g_hcomb]; the g_t1 image stays in place
  float gp2[2][CH];                                      // g_tp2 = g_hcomb * silu'(tp2)
  // block b+1 is issued before block b is read back: its MMAs run under the read-back (accumulators D0 / D1 and
  // user barriers 0 / 1 alternate; the g_t1 images are read-only during the loop)
  run_pair(D0, 0);
#pragma unroll 1
  for (int b = 0; b < 6; ++b) {
    if (b + 1 < 6) run_pair(((b + 1) & 1) ? D1 : D0, (b + 1) & 1);
    pp.wait(b & 1);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[CH];
      ld_chunk((b & 1) * 64 + c * 32, v);
      if (b == 0) {
        if (valid) {
#pragma unroll
          for (int k = 0; k < CH; ++k) v[k] += gho[c][k];
          row_store(a.dh + row * 64, c, v);
        }
      } else if (b < 5) {
        if (valid) row_store(a.ghe + row * 256 + (b - 1) * 64, c, v);
      } else {
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          const float4 d4 = TT(nd, 64 + c * 32 + c0h + 4 * u);
          gp2[c][4 * u] = v[4 * u] * d4.x; gp2[c][4 * u + 1] = v[4 * u + 1] * d4.y;
          gp2[c][4 * u + 2] = v[4 * u + 2] * d4.z; gp2[c][4 * u + 3] = v[4 * u + 3] * d4.w;
        }
      }
    }
  }
  float tmx = 0.f;
  float4* Trow = reinterpret_cast<float4*>(a.T) + row * 256;
  if (spatial) {
    // g_tp1 = (g_tp2 W_p2^T) * silu'(tp1)
    if (rec) { rec_store(NB_GTP2, 0, gp2[0]); rec_store(NB_GTP2, 1, gp2[1]); }
    img_store(0, gp2[0]); img_store(1, gp2[1]);
    run_pair(D0, 0);
    pp.wait(0);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float v[CH];
      ld_chunk(c * 32, v);
#pragma unroll
      for (int u = 0; u < UH; ++u) {
        const float4 d4 = TT(nd, c * 32 + c0h + 4 * u);
        v[4 * u] *= d4.x; v[4 * u + 1] *= d4.y; v[4 * u + 2] *= d4.z; v[4 * u + 3] *= d4.w;
      }
      if (rec) rec_store(NB_GTP1, c, v);
      img_store(c, v);
    }
    // g_nrm = g_tp1 W_p1^T (four 64-row blocks);  T[c][d] = 2 ssum[c][d] g_nrm[c] / den^2 + Wv[c] g_dv[d] / den2
    const float k2 = 2.0f * inv_den * inv_den, q0 = gdv0 / den2, q1 = gdv1 / den2, q2 = gdv2 / den2;
    // same look-ahead over the four blocks; the row loads of ssum for the next 32 coefficients are issued
    // before the current ones are consumed, so they are in flight during the barrier wait and the TMEM load
    // (SLOTS == 2, the two-CTAs-per-SM variant, has 200 registers per thread: it loads the rows where they are used)
    float4 sreg[3 * UH];
    const float4* ssq = tt_base(a.ssum, 192, blockIdx.x, t) + 3 * u0 * G8S;     // G8 layout, see the forward kernel
    if constexpr (SLOTS == 4) {
#pragma unroll
      for (int q = 0; q < 3 * UH; ++q) sreg[q] = __ldg(ssq + q * G8S);
    }
    run_pair(D0, 0);
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      if (b + 1 < 4) run_pair(((b + 1) & 1) ? D1 : D0, (b + 1) & 1);
      pp.wait(b & 1);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[CH];
        ld_chunk((b & 1) * 64 + c * 32, v);
        const int cb = b * 64 + c * 32;                  // first coefficient of the chunk; this thread: cb + c0h ..
        float4 scur[3 * UH];
        if constexpr (SLOTS == 4) {
#pragma unroll
          for (int q = 0; q < 3 * UH; ++q) scur[q] = sreg[q];
          if (cb + 32 < 256) {
#pragma unroll
            for (int q = 0; q < 3 * UH; ++q) sreg[q] = __ldg(ssq + ((cb / 32 + 1) * 24 + q) * G8S);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 3 * UH; ++q) scur[q] = __ldg(ssq + ((cb / 32) * 24 + q) * G8S);
        }
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          const float4 t0 = scur[u * 3], t1 = scur[u * 3 + 1], t2 = scur[u * 3 + 2];
          const float s[12] = {t0.x, t1.x, t2.x, t0.y, t1.y, t2.y, t0.z, t1.z, t2.z, t0.w, t1.w, t2.w};
          const int cc = cb + c0h + 4 * u;
          float nrm[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float gk = k2 * v[4 * u + i], w = upd ? s_wv[cc + i] : 0.f;
            const float t0_ = fmaf(w, q0, gk * s[3 * i]), t1_ = fmaf(w, q1, gk * s[3 * i + 1]), t2_ = fmaf(w, q2, gk * s[3 * i + 2]);
            tmx = fmaxf(tmx, fmaxf(fabsf(t0_), fmaxf(fabsf(t1_), fabsf(t2_))));
            if (valid) Trow[cc + i] = make_float4(t0_, t1_, t2_, 0.f);
            const float a0 = s[3 * i] * inv_den, a1 = s[3 * i + 1] * inv_den, a2 = s[3 * i + 2] * inv_den;
            nrm[i] = a0 * a0 + a1 * a1 + a2 * a2;            // layers.py:123-129: the X operand of post0's weight gradient
          }
          if (rec) TT(nb, NB_NRM + cc) = make_float4(nrm[0], nrm[1], nrm[2], nrm[3]);
        }
      }
    }
  } else if (valid) {
    for (int c = hf * (256 / HALVES); c < (hf + 1) * (256 / HALVES); ++c) Trow[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if constexpr (HALVES == 2) {                           // the row maximum of |T| over both column halves
    if (hf == 1) xch[t] = tmx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (hf == 0) tmx = fmaxf(tmx, xch[t]);
  }
  if (valid && hf == 0) a.tmax[row] = tmx;
  }   // builders
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

// tiles per SM up to which the one-CTA-per-SM configuration (4 weight slots, two threads per atom) is used: 2 for the
// forward kernel, 4 for the backward kernel (measured: cfg3, 2.4 tiles per SM: backward 207 -> 182 us, forward 116 -> 119;
// cfg5, 3 tiles per SM: forward 139 -> 150).  SAKE_NODE_DEEP overrides both (A/B switch).
static int node_deep_factor(bool backward) {
  static int f = -1;
  if (f < 0) { const char* e = getenv("SAKE_NODE_DEEP"); f = e ? atoi(e) : 0; if (f < 0) f = 0; }
  return f > 0 ? f : (backward ? 4 : 2);
}
static int node_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

size_t tc_node_bwd_scratch_bytes(const Dims& d) {
  return align_up(sizeof(float) * (size_t)d.R * 4);
}

int tc_node_post_bwd(const Dims& d, const SakeLayerParams& p, const float* h, const float* v, const float* mask,
                     const Saved& sv, const float* dh_out, const float* dx_out, const float* dv_out, float* dh, float* dx,
                     float* dv, const SakeLayerGrads* g, const BwdScratch& sc, void* wscratch, void* nscratch,
                     cudaStream_t st) {
  uint8_t* wimg = (uint8_t*)wscratch;                     // built by tc_node_post of the same step (saved.wnode)
  NodeBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.R = d.R; a.N = d.N; a.update = d.update; a.has_v = d.has_v; a.spatial = d.spatial;
  a.hdr = d.hdr; a.rowinfo = d.rowinfo;
  a.h = h; a.v = v; a.mask = mask; a.ssum = sv.ssum; a.he = sv.he;
  a.dh_out = dh_out; a.dx_out = dx_out; a.dv_out = dv_out; a.wimg = wimg;
  a.b_p1 = p.post0_bias; a.b_p2 = p.post2_bias; a.b_n1 = p.node0_bias; a.b_n2 = p.node2_bias;
  a.b_v1 = p.vel0_bias; a.vel2 = p.vel2_kernel; a.wv = (d.update && d.spatial) ? p.v_mixing_kernel : nullptr;
  a.dh = dh; a.dx = dx; a.dv = dv; a.T = sc.T; a.ghe = sc.ghe; a.tmax = sc.tmax;
  a.stash = sv.nstash;
  a.qv = sc.qv;                                           // v_mixing gradient: finished by k_pair_reduce (tc_edge.cu)
  (void)nscratch;
  a.nbuf = g ? sc.nbuf : nullptr;
  const int tiles = (d.R + NT_TILE - 1) / NT_TILE;
  const bool deep = tiles <= node_deep_factor(true) * node_num_sms();        // few tiles: latency configuration (see NodePipe)
  // deep: 4 weight slots, two threads per atom (288 threads, one CTA per SM); else 2 slots, thread = atom, two CTAs per SM
  const size_t smem = 4 * NT_IMG + (deep ? 4 : 2) * NT_WCH + NT_VEC * sizeof(float) + 192 + (deep ? 2048 : 0) + 1024;
  static unsigned long long optin4 = 0, optin2 = 0;
  { const int rc = deep ? smem_optin(k_tc_node_post_bwd<4, 2>, smem, optin4) : smem_optin(k_tc_node_post_bwd<2, 1>, smem, optin2); if (rc) return rc; }
  {
    ProfScope prof(7, d.R, st);
    if (deep) k_tc_node_post_bwd<4, 2><<<tiles, 2 * NT_TILE + 32, smem, st>>>(a);
    else k_tc_node_post_bwd<2, 1><<<tiles, NT_TILE + 32, smem, st>>>(a);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

bool tc_node_supported(const Dims& d) { return d.H == 64 && d.A == 4; }

// per-node projections (fills sv.nodeproj, G8 layout); builds the node weight images first unless they are prepared
int tc_node_pre(const Dims& d, const SakeLayerParams& p, const float* h, const Saved& sv, cudaStream_t st) {
  if (d.NP > 256) { set_error("tc_node_pre: projection width %d > 256", d.NP); return SAKE_EUNSUPPORTED; }
  if (!d.prepared) { const int rc = tc_node_prepare(d, p, sv.wnode, st); if (rc) return rc; }
  NodePreArgs a;
  a.R = d.R; a.NP = d.NP; a.hdr = d.hdr; a.h = h;
  a.wimg = (const uint8_t*)sv.wnode + NTP_OFF;
  a.bias = reinterpret_cast<const float*>((const uint8_t*)sv.wnode + NTP_OFF + NTP_WBYTES);
  a.proj = sv.nodeproj;
  const int tiles = (d.R + NT_TILE - 1) / NT_TILE;
  const int sms = node_num_sms();
  const size_t smem = NTP_WBYTES + 4 * NT_IMG + 256 * sizeof(float) + 64 + 1024;
  static unsigned long long optin = 0;
  { const int rc = smem_optin(k_tc_node_pre, smem, optin); if (rc) return rc; }
  {
    ProfScope prof(9, d.R, st);
    k_tc_node_pre<<<tiles < sms ? tiles : sms, NT_TILE, smem, st>>>(a);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// all chunks, forward and transposed: the backward call of the step reuses the image
int tc_node_prepare(const Dims& d, const SakeLayerParams& p, void* wnode, cudaStream_t st) {
  k_node_w_prep<<<(NTB_CHUNKS * 64 * 8 + 255) / 256, 256, 0, st>>>(p, (uint8_t*)wnode, NTB_CHUNKS);
  k_node_pre_w_prep<<<(2 * 256 * 8 + 255) / 256, 256, 0, st>>>(p, d.H, d.K, d.Kp, (uint8_t*)wnode + NTP_OFF,
                                                              reinterpret_cast<float*>((uint8_t*)wnode + NTP_OFF + NTP_WBYTES),
                                                              (uint8_t*)wnode + NTQ_OFF);
  note_launches(2);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// dh += gproj W^T (the images were built by the forward call of the step)
int tc_node_pre_bwd(const Dims& d, const Saved& sv, const BwdScratch& sc, float* dh, cudaStream_t st) {
  NodePreBwdArgs a;
  a.R = d.R; a.NP = d.NP; a.hdr = d.hdr; a.gproj = sc.gproj; a.wimg = (const uint8_t*)sv.wnode + NTQ_OFF; a.dh = dh;
  const int tiles = (d.R + NT_TILE - 1) / NT_TILE;
  const int sms = node_num_sms();
  const size_t smem = NTQ_WBYTES + 4 * NT_IMG + 64 + 1024;
  static unsigned long long optin = 0;
  { const int rc = smem_optin(k_tc_node_pre_bwd, smem, optin); if (rc) return rc; }
  {
    ProfScope prof(12, d.R, st);
    k_tc_node_pre_bwd<<<tiles < sms ? tiles : sms, NT_TILE, smem, st>>>(a);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tc_node_post(const Dims& d, const SakeLayerParams& p, const float* h, const float* x, const float* v,
                 const float* mask, float* h_out, float* x_out, float* v_out, const Saved& sv, void* wscratch,
                 cudaStream_t st) {
  uint8_t* wimg = (uint8_t*)wscratch;
  (void)p;                                                // the images were built by tc_node_pre of the same call
  NodeFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.R = d.R; a.N = d.N; a.update = d.update; a.has_v = d.has_v; a.spatial = d.spatial;
  a.hdr = d.hdr; a.rowinfo = d.rowinfo;
  a.h = h; a.x = x; a.v = v; a.mask = mask; a.ssum = sv.ssum; a.he = sv.he; a.wimg = wimg;
  a.b_p1 = p.post0_bias; a.b_p2 = p.post2_bias; a.b_n1 = p.node0_bias; a.b_n2 = p.node2_bias;
  a.b_v1 = p.vel0_bias; a.vel2 = p.vel2_kernel; a.wv = (d.update && d.spatial) ? p.v_mixing_kernel : nullptr;
  a.h_out = h_out; a.x_out = x_out; a.v_out = v_out; a.stash = sv.nstash;
  const int tiles = (d.R + NT_TILE - 1) / NT_TILE;
  const bool deep = tiles <= node_deep_factor(false) * node_num_sms();
  const size_t smem = 4 * NT_IMG + (deep ? 4 : 2) * NT_WCH + NT_VEC * sizeof(float) + 192 + (deep ? 2048 : 0) + 1024;
  static unsigned long long optin4 = 0, optin2 = 0;
  { const int rc = deep ? smem_optin(k_tc_node_post<4, 2>, smem, optin4) : smem_optin(k_tc_node_post<2, 1>, smem, optin2); if (rc) return rc; }
  {
    ProfScope prof(6, d.R, st);
    if (deep) k_tc_node_post<4, 2><<<tiles, 2 * NT_TILE + 32, smem, st>>>(a);
    else k_tc_node_post<2, 1><<<tiles, NT_TILE + 32, smem, st>>>(a);
  }
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sake
