// Persistent zoneout-LSTM recurrence (sm_100a): one thread-block cluster of H/16 CTAs owns 4 batch
// rows for the whole sequence.  Each CTA owns 16 hidden units (= 64 gate columns); its slice of the
// recurrent kernel Wh stays in REGISTERS for all T steps; the 4x H hidden state is exchanged through
// distributed shared memory with one hardware cluster barrier per step.  No global synchronisation,
// no re-reading of weights: per step the only HBM/L2 traffic is the pre-computed input projection
// xg[t] (prefetched one step ahead) and the saved activations.
//
// Math: TF LSTMCell (i,j,f,o; forget_bias) + tacotron2 ZoneoutLSTMCell, SURVEY.md A.5/A.6;
// sequence-length semantics of tf.nn.bidirectional_dynamic_rnn (module.py:93-108).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace satk {

constexpr int LBG = 4;    // batch rows per cluster
constexpr int LUH = 16;   // hidden units per CTA

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2/(1+e^{2x}); absolute error ~2e-7, saturates correctly for |x| large
  float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

template <int H>
__global__ void __launch_bounds__(256, 1) lstm_fwd_kernel(const satk_lstm_fwd_desc d) {
  constexpr int CS = H / LUH;
  constexpr int KPT = H / 4;  // k's per thread
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * LBG;
  const int tid = threadIdx.x;

  __shared__ __align__(16) float hbuf[2][H][LBG];
  __shared__ float gsm[LBG][64];

  // --- gate-GEMM role: thread = (col 0..63, kq 0..3)
  const int col = tid >> 2, kq = tid & 3;
  const int gate = col >> 4, unit = col & 15;
  const int gcol = gate * H + rank * LUH + unit;  // column in the [.,4H] kernel
  float w[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) w[i] = __ldg(d.Wh + (long long)(kq + 4 * i) * (4 * H) + gcol);

  const int myb = b0 + kq;  // batch row whose xg this thread adds
  const bool myb_ok = myb < d.B;
  const int mylen = myb_ok ? (d.lengths ? (int)d.lengths[myb] : d.T) : 0;

  // --- pointwise role: tid < 64 -> (pb, pu)
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < d.B;
  const int plen = prow_ok ? (d.lengths ? (int)d.lengths[prow] : d.T) : 0;
  const int pidx = rank * LUH + pu;
  float c_st = 0.f, h_st = 0.f;

  for (int i = tid; i < 2 * H * LBG; i += 256) (&hbuf[0][0][0])[i] = 0.f;
  cluster.sync();

  auto xg_at = [&](int s) -> float {
    if (!(s < mylen)) return 0.f;
    int p = d.reverse ? (mylen - 1 - s) : s;
    return __ldg(d.xg + ((long long)p * d.B + myb) * (4 * H) + gcol);
  };
  float xg_next = xg_at(0);

  for (int s = 0; s < d.T; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    const float xg_cur = xg_next;
    if (s + 1 < d.T) xg_next = xg_at(s + 1);

    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const float4 hv = *reinterpret_cast<const float4*>(&hbuf[cur][kq + 4 * i][0]);
      acc0 = fmaf(w[i], hv.x, acc0);
      acc1 = fmaf(w[i], hv.y, acc1);
      acc2 = fmaf(w[i], hv.z, acc2);
      acc3 = fmaf(w[i], hv.w, acc3);
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
      acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
      acc3 += __shfl_xor_sync(0xffffffffu, acc3, o);
    }
    float mine = (kq == 0) ? acc0 : (kq == 1) ? acc1 : (kq == 2) ? acc2 : acc3;
    gsm[kq][col] = mine + xg_cur;
    __syncthreads();

    if (tid < 64) {
      const bool valid = prow_ok && (s < plen);
      float h_new_state = h_st;
      if (valid) {
        const int p = d.reverse ? (plen - 1 - s) : s;
        float gi = fast_sigmoid(gsm[pb][0 * 16 + pu]);
        float gj = fast_tanh(gsm[pb][1 * 16 + pu]);
        float gf = fast_sigmoid(gsm[pb][2 * 16 + pu] + d.forget_bias);
        float go = fast_sigmoid(gsm[pb][3 * 16 + pu]);
        float c_new = gf * c_st + gi * gj;
        float h_new = go * fast_tanh(c_new);
        const long long o1 = ((long long)p * d.B + prow) * H + pidx;
        d.out[((long long)p * d.B + prow) * d.ld_out + pidx] = h_new;
        if (d.gates) {
          const long long o4 = ((long long)p * d.B + prow) * (4 * H) + pidx;
          d.gates[o4] = gi;
          d.gates[o4 + H] = gj;
          d.gates[o4 + 2 * H] = gf;
          d.gates[o4 + 3 * H] = go;
          d.c_prev[o1] = c_st;
          d.h_prev[o1] = h_st;
        }
        const long long om = ((long long)s * d.B + prow) * H + pidx;  // masks are indexed by processing step
        float mc = d.mask_c ? (float)d.mask_c[om] : (1.f - d.zc);
        float mh = d.mask_h ? (float)d.mask_h[om] : (1.f - d.zh);
        c_st = c_st + mc * (c_new - c_st);
        h_new_state = h_st + mh * (h_new - h_st);
        h_st = h_new_state;
      } else if (prow_ok) {
        // position s is past this row's length: zero output (dynamic_rnn), state frozen
        const long long o1 = ((long long)s * d.B + prow) * H + pidx;
        d.out[((long long)s * d.B + prow) * d.ld_out + pidx] = 0.f;
        if (d.gates) {
          const long long o4 = ((long long)s * d.B + prow) * (4 * H) + pidx;
          d.gates[o4] = 0.f; d.gates[o4 + H] = 0.f; d.gates[o4 + 2 * H] = 0.f; d.gates[o4 + 3 * H] = 0.f;
          d.c_prev[o1] = 0.f;
          d.h_prev[o1] = 0.f;
        }
      }
      // publish the new hidden state to every CTA of the cluster
#pragma unroll 4
      for (int r = 0; r < CS; ++r) {
        float* remote = cluster.map_shared_rank(&hbuf[nxt][pidx][pb], r);
        *remote = h_new_state;
      }
    }
    cluster.sync();
  }
}

template <int H>
__global__ void __launch_bounds__(256, 1) lstm_bwd_kernel(const satk_lstm_bwd_desc d) {
  constexpr int CS = H / LUH;
  constexpr int K4 = 4 * H;
  constexpr int KPT = K4 / 16;  // 64 (H=256) or 32 (H=128)
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * LBG;
  const int tid = threadIdx.x;

  __shared__ __align__(16) float dgbuf[2][K4][LBG];
  __shared__ float dhsm[LBG][LUH];

  // --- GEMM role (dh_prev = dg . Wh^T restricted to my 16 units): thread = (unit 0..15, kq 0..15)
  const int gu = tid >> 4, kq = tid & 15;
  float w[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) w[i] = __ldg(d.Wh + (long long)(rank * LUH + gu) * K4 + kq + 16 * i);

  // --- pointwise role
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < d.B;
  const int plen = prow_ok ? (d.lengths ? (int)d.lengths[prow] : d.T) : 0;
  const int pidx = rank * LUH + pu;
  float dc = 0.f, dh = 0.f;

  for (int s = d.T - 1; s >= 0; --s) {
    const int cur = s & 1;
    float dh_part = dh;
    if (tid < 64) {
      float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
      const bool valid = prow_ok && (s < plen);
      if (valid) {
        const int p = d.reverse ? (plen - 1 - s) : s;
        const long long o1 = ((long long)p * d.B + prow) * H + pidx;
        const long long o4 = ((long long)p * d.B + prow) * K4 + pidx;
        const float gi = d.gates[o4], gj = d.gates[o4 + H], gf = d.gates[o4 + 2 * H], go = d.gates[o4 + 3 * H];
        const float cp = d.c_prev[o1];
        const long long om = ((long long)s * d.B + prow) * H + pidx;
        const float mc = d.mask_c ? (float)d.mask_c[om] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)d.mask_h[om] : (1.f - d.zh);
        const float c_new = gf * cp + gi * gj;
        const float tc = fast_tanh(c_new);
        const float dh_new = d.dout[((long long)p * d.B + prow) * d.ld_dout + pidx] + mh * dh;
        dh_part = (1.f - mh) * dh;
        const float dcn = mc * dc + dh_new * go * (1.f - tc * tc);
        dgo = dh_new * tc * go * (1.f - go);
        dgi = dcn * gj * gi * (1.f - gi);
        dgj = dcn * gi * (1.f - gj * gj);
        dgf = dcn * cp * gf * (1.f - gf);
        dc = (1.f - mc) * dc + dcn * gf;
        d.dgates[o4] = dgi; d.dgates[o4 + H] = dgj; d.dgates[o4 + 2 * H] = dgf; d.dgates[o4 + 3 * H] = dgo;
      } else if (prow_ok) {
        const long long o4 = ((long long)s * d.B + prow) * K4 + pidx;
        d.dgates[o4] = 0.f; d.dgates[o4 + H] = 0.f; d.dgates[o4 + 2 * H] = 0.f; d.dgates[o4 + 3 * H] = 0.f;
      }
#pragma unroll 4
      for (int r = 0; r < CS; ++r) {
        float* base = cluster.map_shared_rank(&dgbuf[cur][0][0], r);
        base[(0 * H + pidx) * LBG + pb] = dgi;
        base[(1 * H + pidx) * LBG + pb] = dgj;
        base[(2 * H + pidx) * LBG + pb] = dgf;
        base[(3 * H + pidx) * LBG + pb] = dgo;
      }
    }
    cluster.sync();
    if (s == 0) break;  // no earlier step consumes dh_prev
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const float4 g4 = *reinterpret_cast<const float4*>(&dgbuf[cur][kq + 16 * i][0]);
      acc0 = fmaf(w[i], g4.x, acc0);
      acc1 = fmaf(w[i], g4.y, acc1);
      acc2 = fmaf(w[i], g4.z, acc2);
      acc3 = fmaf(w[i], g4.w, acc3);
    }
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) {
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
      acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
      acc3 += __shfl_xor_sync(0xffffffffu, acc3, o);
    }
    if (kq < 4) dhsm[kq][gu] = (kq == 0) ? acc0 : (kq == 1) ? acc1 : (kq == 2) ? acc2 : acc3;
    __syncthreads();
    if (tid < 64) dh = dh_part + dhsm[pb][pu];
    __syncthreads();
  }
}

template <typename Kern, typename Desc>
static int launch_cluster(Kern kern, const Desc& d, int H, int B, cudaStream_t st) {
  const int CS = H / LUH;
  const int groups = (B + LBG - 1) / LBG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CS);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CS > 8) SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  SATK_CUDA(cudaLaunchKernelEx(&cfg, kern, d));
  return SATK_OK;
}

int lstm_max_clusters_h256() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16 * 64);
  cfg.blockDim = dim3(256);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 16;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaFuncSetAttribute(lstm_fwd_kernel<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_fwd_kernel<256>, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

}  // namespace satk

using namespace satk;

extern "C" {

int satk_lstm_seq_fwd(const satk_lstm_fwd_desc* d, void* stream) {
  SATK_CHECK_ARG(d->H == 128 || d->H == 256, "lstm_seq_fwd: H=%d unsupported (128 or 256)", d->H);
  SATK_CHECK_ARG(d->T > 0 && d->B > 0, "lstm_seq_fwd: empty T=%d B=%d", d->T, d->B);
  SATK_CHECK_ARG(d->ld_out >= d->H, "lstm_seq_fwd: ld_out=%lld < H", d->ld_out);
  SATK_CHECK_ARG((d->gates == nullptr) == (d->c_prev == nullptr) && (d->gates == nullptr) == (d->h_prev == nullptr),
                 "lstm_seq_fwd: gates/c_prev/h_prev must be all set or all NULL");
  if (d->H == 256) return launch_cluster(lstm_fwd_kernel<256>, *d, 256, d->B, (cudaStream_t)stream);
  return launch_cluster(lstm_fwd_kernel<128>, *d, 128, d->B, (cudaStream_t)stream);
}

int satk_lstm_seq_bwd(const satk_lstm_bwd_desc* d, void* stream) {
  SATK_CHECK_ARG(d->H == 128 || d->H == 256, "lstm_seq_bwd: H=%d unsupported (128 or 256)", d->H);
  SATK_CHECK_ARG(d->T > 0 && d->B > 0, "lstm_seq_bwd: empty T=%d B=%d", d->T, d->B);
  SATK_CHECK_ARG(d->ld_dout >= d->H, "lstm_seq_bwd: ld_dout=%lld < H", d->ld_dout);
  if (d->H == 256) return launch_cluster(lstm_bwd_kernel<256>, *d, 256, d->B, (cudaStream_t)stream);
  return launch_cluster(lstm_bwd_kernel<128>, *d, 128, d->B, (cudaStream_t)stream);
}

}  // extern "C"
