// Ragged batches: the tables that let every kernel work on the REAL atoms of a padded batch only
// (common.cuh: RaggedHdr / rowinfo / tileinfo / molinfo), built on the device from n_real[B], plus the
// gather / scatter between the caller's padded [B, N, W] tensors and the compact [R, W] layout.
//
// What this replaces in the reference: QM9-style padding carries a float mask = outer(m, m) through every
// O(N^2) tensor (scripts/qm9/run.py:23-24,35; sake/layers.py:120-123,137-138,164-165,178-179,219-221) and
// computes all N^2 pairs of the padded width.  With guarded masking a padded atom contributes exactly zero
// to every real atom, so running each molecule unpadded is the same function on the real atoms — the
// property sake/tests/test_mask.py:202-240 asserts — at sum(n_b^2) instead of B*N^2 pairs.
//
// Compact order: molecules sorted by n_real (stable in the batch index), so that consecutive rows of one
// 128-pair tile share the segment length n (tc_tile.cuh).  Everything is computed on the device: no host
// copy of n_real is needed, the launch sequence has no data-dependent host decision and stays valid inside
// a CUDA graph when the next batch has other sizes.
#include "common.cuh"

namespace sake {

constexpr int RG_MAXN = 128;                 // ragged tiles hold whole receiver rows: N <= 128
constexpr int RG_CLS = RG_MAXN + 1;          // classes n = 0 .. 128

struct RaggedBlob {
  RaggedHdr* hdr;
  int* cls;            // [4][RG_CLS]: molecules / first row / first pair / first tile of every class (exclusive prefix)
  int2* molinfo;       // [B]
  int4* rowinfo;       // [B*N]
  int4* tileinfo;      // [B*N]
};
static size_t ragged_bytes(int B, int N) {
  return 256 + align_up(sizeof(int) * 4 * RG_CLS) + align_up(sizeof(int2) * (size_t)B) +
         2 * align_up(sizeof(int4) * (size_t)B * N);
}
static RaggedBlob carve_ragged(const void* blob, int B, int N) {
  RaggedBlob r;
  char* b = (char*)blob;
  r.hdr = (RaggedHdr*)b; b += 256;
  r.cls = (int*)b; b += align_up(sizeof(int) * 4 * RG_CLS);
  r.molinfo = (int2*)b; b += align_up(sizeof(int2) * (size_t)B);
  r.rowinfo = (int4*)b; b += align_up(sizeof(int4) * (size_t)B * N);
  r.tileinfo = (int4*)b;
  return r;
}
void ragged_attach(Dims& d, const void* blob) {
  RaggedBlob r = carve_ragged(blob, d.B, d.N);
  d.hdr = r.hdr; d.rowinfo = r.rowinfo; d.tileinfo = r.tileinfo; d.molinfo = r.molinfo;
}

__device__ __forceinline__ int clamp_n(int n, int N) { return n < 0 ? 0 : (n > N ? N : n); }

// One CTA: class histogram, exclusive prefixes, header and the tile table.
__global__ void __launch_bounds__(256) k_ragged_classes(int B, int N, const int* __restrict__ n_real, RaggedBlob r) {
  __shared__ int hist[RG_CLS], cmol[RG_CLS + 1], crow[RG_CLS + 1], cpair[RG_CLS + 1], ctile[RG_CLS + 1];
  for (int t = threadIdx.x; t < RG_CLS; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) atomicAdd(&hist[clamp_n(n_real[b], N)], 1);
  __syncthreads();
  // exclusive prefixes of molecules / rows / pairs / tiles over the classes: four Hillis-Steele scans side by side
  // (a serial loop over the 129 classes in one thread was most of this kernel's 26 us)
  __shared__ int4 scan[2][256];
  {
    const int n = threadIdx.x;
    int4 v = make_int4(0, 0, 0, 0);
    if (n < RG_CLS) {
      const int rows = hist[n] * n;
      v = make_int4(hist[n], rows, rows * n, n > 0 ? (rows + 128 / n - 1) / (128 / n) : 0);
    }
    scan[0][n] = v;
    __syncthreads();
    int cur = 0;
    for (int o = 1; o < 256; o <<= 1) {
      int4 w = scan[cur][n];
      if (n >= o) { const int4 q = scan[cur][n - o]; w.x += q.x; w.y += q.y; w.z += q.z; w.w += q.w; }
      scan[cur ^ 1][n] = w;
      cur ^= 1;
      __syncthreads();
    }
    const int4 inc = scan[cur][n];                       // inclusive prefix; exclusive = inclusive - own
    if (n < RG_CLS) { cmol[n] = inc.x - v.x; crow[n] = inc.y - v.y; cpair[n] = inc.z - v.z; ctile[n] = inc.w - v.w; }
    if (n == RG_CLS - 1) {
      cmol[RG_CLS] = inc.x; crow[RG_CLS] = inc.y; cpair[RG_CLS] = inc.z; ctile[RG_CLS] = inc.w;
      r.hdr->R = inc.y; r.hdr->num_tiles = inc.w; r.hdr->B = B; r.hdr->reserved = 0;
      r.hdr->P = inc.z; r.hdr->R64 = inc.y;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < RG_CLS; t += blockDim.x) {
    r.cls[t] = cmol[t]; r.cls[RG_CLS + t] = crow[t]; r.cls[2 * RG_CLS + t] = cpair[t]; r.cls[3 * RG_CLS + t] = ctile[t];
  }
  // tiles of class n: rpt = 128 / n whole rows each (the last one of the class may hold fewer)
  for (int n = 1; n < RG_CLS; ++n) {
    const int nt = ctile[n + 1] - ctile[n];
    if (nt == 0) continue;                                  // block-uniform
    const int rpt = 128 / n, rows = hist[n] * n;
    for (int k = threadIdx.x; k < nt; k += blockDim.x) {
      const int lr0 = k * rpt;
      r.tileinfo[ctile[n] + k] = make_int4(crow[n] + lr0, min(rpt, rows - lr0), n, cpair[n] + lr0 * n);
    }
  }
}

// One thread per molecule: stable rank inside its class -> position in the compact order; writes its rows.
__global__ void __launch_bounds__(128) k_ragged_rows(int B, int N, const int* __restrict__ n_real, RaggedBlob r) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = clamp_n(n_real[b], N);
  int rank = 0;
  for (int q = 0; q < b; ++q) rank += clamp_n(__ldg(n_real + q), N) == n ? 1 : 0;   // same address across the warp: broadcast
  const int mol0 = r.cls[RG_CLS + n] + rank * n;
  const int pair0 = r.cls[2 * RG_CLS + n] + rank * n * n;
  r.molinfo[b] = make_int2(mol0, n);
  for (int i = 0; i < n; ++i) r.rowinfo[mol0 + i] = make_int4(mol0, n, pair0 + i * n, b * N + i);
}

// compact[r][:] = padded[prow(r)][:]
__global__ void __launch_bounds__(256) k_ragged_gather(const RaggedHdr* hdr, const int4* __restrict__ rowinfo, int width,
                                                       const float* __restrict__ padded, float* __restrict__ compact) {
  const long long total = (long long)hdr->R * width;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / width), w = (int)(t - (long long)r * width);
    compact[t] = padded[(size_t)__ldg(rowinfo + r).w * width + w];
  }
}
// padded[prow(r)][:] = alpha * compact[r][:]   (rows of padding atoms are not touched)
__global__ void __launch_bounds__(256) k_ragged_scatter(const RaggedHdr* hdr, const int4* __restrict__ rowinfo, int width,
                                                        float alpha, const float* __restrict__ compact,
                                                        float* __restrict__ padded) {
  const long long total = (long long)hdr->R * width;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / width), w = (int)(t - (long long)r * width);
    padded[(size_t)__ldg(rowinfo + r).w * width + w] = alpha * compact[t];
  }
}

}  // namespace sake

using namespace sake;

extern "C" {

size_t sake_ragged_bytes(int32_t B, int32_t N) {
  if (B < 0 || N <= 0) return 0;
  return ragged_bytes(B, N);
}

int sake_ragged_prepare(int32_t B, int32_t N, const int32_t* n_real, void* ragged, size_t ragged_bytes_,
                        sake_stream_t stream) {
  if (B < 0 || N <= 0 || !n_real || !ragged) { set_error("sake_ragged_prepare: bad argument"); return SAKE_EINVAL; }
  if (N > RG_MAXN) { set_error("ragged batches need N <= %d (got %d): rows longer than one 128-pair tile use the float mask", RG_MAXN, N); return SAKE_EUNSUPPORTED; }
  if ((long long)B * N * N > 0x7fffffffLL) { set_error("ragged batch too large: B*N*N = %lld pairs", (long long)B * N * N); return SAKE_EUNSUPPORTED; }
  if (ragged_bytes_ < ragged_bytes(B, N)) { set_error("ragged buffer too small: %zu < %zu", ragged_bytes_, ragged_bytes(B, N)); return SAKE_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  RaggedBlob r = carve_ragged(ragged, B, N);
  k_ragged_classes<<<1, 256, 0, st>>>(B, N, n_real, r);
  if (B > 0) k_ragged_rows<<<(B + 127) / 128, 128, 0, st>>>(B, N, n_real, r);
  note_launches(B > 0 ? 2 : 1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int ragged_move(const void* ragged, int32_t B, int32_t N, int32_t width, bool gather, float alpha, const float* src,
                       float* dst, cudaStream_t st) {
  if (B < 0 || N <= 0 || width <= 0 || !ragged || !src || !dst) { set_error("sake_ragged_gather/scatter: bad argument"); return SAKE_EINVAL; }
  if (B == 0) return 0;
  RaggedBlob r = carve_ragged(ragged, B, N);
  long long total = (long long)B * N * width;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (gather) k_ragged_gather<<<grid, 256, 0, st>>>(r.hdr, r.rowinfo, width, src, dst);
  else k_ragged_scatter<<<grid, 256, 0, st>>>(r.hdr, r.rowinfo, width, alpha, src, dst);
  note_launches(1);
  SAKE_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sake_ragged_gather(const void* ragged, int32_t B, int32_t N, int32_t width, const float* padded, float* compact,
                       sake_stream_t stream) {
  return ragged_move(ragged, B, N, width, true, 1.0f, padded, compact, (cudaStream_t)stream);
}

int sake_ragged_scatter(const void* ragged, int32_t B, int32_t N, int32_t width, float alpha, const float* compact,
                        float* padded, sake_stream_t stream) {
  return ragged_move(ragged, B, N, width, false, alpha, compact, padded, (cudaStream_t)stream);
}

}  // extern "C"
