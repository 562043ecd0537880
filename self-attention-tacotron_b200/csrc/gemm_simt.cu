// fp32 SIMT GEMM tile (engine 1 of satk_gemm).  General strides / transposes / batching /
// shifted-row convolution taps / split-K.  This is the exact-fp32 path: it carries the small and
// oddly shaped products (batched attention, conv weight gradients, N<64 layers) and is the
// numerical cross-check for the tcgen05 3xTF32 tile in gemm_tc.cu.
#include "common.cuh"

namespace satk {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

struct GemmArgs {
  satk_gemm_desc d;
};

__device__ __forceinline__ bool shifted_row(int row, int shift, int seq_len, int limit, int& out) {
  // row -> row+shift within its sequence (seq_len rows per sequence); false when it leaves the sequence
  if (seq_len > 0) {
    int t = row % seq_len + shift;
    if (t < 0 || t >= seq_len) return false;
  }
  out = row + shift;
  return out >= 0 && out < limit;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs args) {
  const satk_gemm_desc& d = args.d;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM;
  int n0, ks, nks;
  if (d.split_k > 1) {
    int ntn = (d.N + BN - 1) / BN;
    n0 = (blockIdx.y % ntn) * BN;
    ks = blockIdx.y / ntn;
    nks = d.split_k;
  } else {
    n0 = blockIdx.y * BN;
    ks = 0;
    nks = 1;
  }
  const int z = blockIdx.z;
  const int z1 = z / d.batch2, z2 = z % d.batch2;
  const float* __restrict__ A = d.A + z1 * d.sA1 + z2 * d.sA2;
  const float* __restrict__ B = d.B + z1 * d.sB1 + z2 * d.sB2;
  float* __restrict__ C = d.C + z1 * d.sC1 + z2 * d.sC2;

  if (d.causal_skip == 1 && n0 > m0 + BM - 1) {
    // QK^T under a causal mask: tile entirely above the diagonal is never read by the softmax
    return;
  }

  // K range of this split
  int kchunk = ((d.K + nks - 1) / nks + BK - 1) / BK * BK;
  int kbeg = ks * kchunk;
  int kend = min(d.K, kbeg + kchunk);
  if (d.causal_skip == 2) kend = min(kend, m0 + BM);  // P.V under a causal mask: P[m,k]=0 for k>m
  const int ktiles = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  const int total = ktiles * d.taps;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float ra[4], rb[4];

  auto load_tiles = [&](int it) {
    const int tap = it / ktiles;
    const int k0 = kbeg + (it % ktiles) * BK;
    const int shift = d.shift0 + tap * d.tap_dir + z1 * d.shift_per_batch1;
    const float* __restrict__ Bt = B + (long long)tap * d.sBtap;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int r, c;  // r: m index in tile, c: k index in tile
      if (!d.transA) { r = idx / BK; c = idx % BK; } else { c = idx / BM; r = idx % BM; }
      int m = m0 + r, k = k0 + c;
      float v = 0.0f;
      if (m < d.M && k < kend) {
        if (!d.transA) {
          int row;
          if (shifted_row(m, shift, d.seq_len, d.M, row)) v = __ldg(A + (long long)row * d.lda + k);
        } else {
          int row;
          if (shifted_row(k, shift, d.seq_len, d.K, row)) v = __ldg(A + (long long)row * d.lda + m);
        }
      }
      ra[e] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int r, c;  // r: n index in tile, c: k index
      if (!d.transB) { c = idx / BN; r = idx % BN; } else { r = idx / BK; c = idx % BK; }
      int n = n0 + r, k = k0 + c;
      float v = 0.0f;
      if (n < d.N && k < kend) v = d.transB ? __ldg(Bt + (long long)n * d.ldb + k) : __ldg(Bt + (long long)k * d.ldb + n);
      rb[e] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int r, c;
      if (!d.transA) { r = idx / BK; c = idx % BK; } else { c = idx / BM; r = idx % BM; }
      As[buf][c][r] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;
      int r, c;
      if (!d.transB) { c = idx / BN; r = idx % BN; } else { r = idx / BK; c = idx % BK; }
      Bs[buf][c][r] = rb[e];
    }
  };

  if (total > 0) {
    load_tiles(0);
    store_tiles(0);
  }
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) load_tiles(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < total) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= d.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= d.N) continue;
      float v = d.alpha * acc[i][j];
      float* cp = C + (long long)m * d.ldc + n;
      if (d.split_k > 1) {
        atomicAdd(cp, v);
        continue;
      }
      if (d.bias) v += __ldg(d.bias + n);
      v = apply_act(v, d.act);
      if (d.keep_mask) v = d.keep_mask[(long long)m * d.N + n] ? v * d.keep_scale : 0.0f;
      if (d.residual) v += __ldg(d.residual + (long long)m * d.ldres + n);
      if (d.beta != 0.0f) v += d.beta * (*cp);
      *cp = v;
    }
  }
}

int gemm_simt_launch(const satk_gemm_desc* d, cudaStream_t st) {
  SATK_CHECK_ARG(d->kshift0 == 0 && d->kshift_per_batch1 == 0 && d->bank_widths == 0,
                 "satk_gemm: K-shifted batches and conv banks are served by the tcgen05 tile only");
  GemmArgs a;
  a.d = *d;
  if (a.d.batch1 < 1) a.d.batch1 = 1;
  if (a.d.batch2 < 1) a.d.batch2 = 1;
  if (a.d.taps < 1) a.d.taps = 1;
  if (a.d.split_k < 1) a.d.split_k = 1;
  SATK_CHECK_ARG(a.d.M > 0 && a.d.N > 0 && a.d.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.d.M, a.d.N, a.d.K);
  SATK_CHECK_ARG(!(a.d.split_k > 1 && (a.d.bias || a.d.act || a.d.residual || a.d.keep_mask)),
                 "gemm: split_k excludes bias/act/residual/mask epilogues");
  SATK_CHECK_ARG(!(a.d.keep_mask && (a.d.batch1 * a.d.batch2 > 1)), "gemm: keep_mask with batching unsupported");
  int gm = ceil_div(a.d.M, BM), gn = ceil_div(a.d.N, BN);
  dim3 grid(gm, gn * a.d.split_k, a.d.batch1 * a.d.batch2);
  SATK_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm: grid too large");
  gemm_simt_kernel<<<grid, 256, 0, st>>>(a);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

}  // namespace satk
