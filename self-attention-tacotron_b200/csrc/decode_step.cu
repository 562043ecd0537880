// Free-running decoder step (sm_100a): the kernels behind PREDICT mode (predict_mel.py path).
// One decoder step = a fixed sequence of these launches; every kernel reads the step index t from DEVICE memory, so the
// host captures the sequence once in a CUDA graph and replays it — no per-step host work, no re-attention over the
// history (the reference's TransformerWrapper recomputes self-attention over all past outputs each step, O(T^3) in total,
// rnn_wrappers.py:111-124; here keys/values of past steps are cached and only row t is computed: O(T^2)).
//   rowgemm      skinny dense layers  C[M<=batch, N] = act(A.W + b) (+res): pre-net, LSTM gate rows, query / K / V / Q / O
//                projections, transform, mel + stop projections.  Weights stream from L2 once per launch.
//                With lstm_H > 0 the epilogue is the ZoneoutLSTMCell pointwise update (inference interpolation, A.5/A.6).
//   attn_step    ForwardAttention / LocationSensitive / Bahdanau step, one cluster of 8 CTAs per utterance
//                (forward_attention.py:88-122, :13-26; A.8) incl. the transition agent (:111-114)
//   sa_step      causal scaled-dot-product attention of the newest query over the cached history (self_attention.py:45-65)
//   tick         t += 1 and stop-token bookkeeping (StopTokenBasedInferenceHelper: sigmoid(stop) > 0.5 for the whole batch
//                after min_iters)
#include <cooperative_groups.h>
#include "common.cuh"

namespace satk {
namespace dstep {

namespace cg = cooperative_groups;

constexpr int MR = 16;     // rows per pass
constexpr int KC = 128;    // reduction chunk staged in shared memory
constexpr int NC = 32;     // columns per CTA (one per lane)
constexpr int TCS = 8;     // CTAs per cluster of the one-utterance-per-cluster kernels

// Programmatic dependent launch: the step's kernels are launched with programmaticStreamSerialization, so kernel N+1 is scheduled
// while kernel N drains; `pdl_wait` blocks until kernel N has completed and its writes are visible, `pdl_launch` lets N+1 start early.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float ftanh_(float x) {
  x = fminf(fmaxf(x, -15.f), 15.f);
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

// Skinny GEMM.  A cluster of KS CTAs shares one block of 32 output columns and splits the reduction; the partial
// [16 x 32] tiles are summed by rank 0 through distributed shared memory, which then runs the epilogue: bias / activation /
// residual, or (lstm_H > 0) the whole ZoneoutLSTMCell pointwise update — in that mode the 32 columns are the 4 gates of 8
// hidden units, so the gate pre-activations never leave the chip.
template <int KS>
__global__ void __launch_bounds__(256) rowgemm_k(const satk_rowgemm_desc d) {
  __shared__ __align__(16) float As[KC][MR];
  __shared__ float red[8][MR][NC + 1];
  __shared__ float part[MR * NC];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = KS > 1 ? (int)cluster.block_rank() : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_enter();
  int i = 0, cb = blockIdx.x / KS;
  const bool lstm = d.lstm_H > 0;
  if (!lstm)
    while (i < d.nmat - 1 && cb >= (d.N[i] + NC - 1) / NC) { cb -= (d.N[i] + NC - 1) / NC; ++i; }
  const int N = d.N[i];
  const int tt = d.t_ptr ? *d.t_ptr : 0;
  const long long t = tt;
  const float* __restrict__ A = d.A + t * d.a_tstride + (long long)(tt & 1) * d.a_pstride;
  const float* __restrict__ W = d.W[i];
  const int col = lstm ? ((lane >> 3) * d.lstm_H + cb * 8 + (lane & 7)) : (cb * NC + lane);
  const bool cok = col < N;
  const int kper = (((d.K + KS - 1) / KS) + 15) / 16 * 16;
  const int kr0 = rank * kper, kr1 = min(d.K, kr0 + kper);
  for (int m0 = 0; m0 < d.M; m0 += MR) {
    float acc[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m] = 0.f;
    for (int k0 = kr0; k0 < kr1; k0 += KC) {
      __syncthreads();
      for (int idx = tid; idx < MR * KC; idx += 256) {
        const int m = idx % MR, k = idx / MR;
        As[k][m] = (m0 + m < d.M && k0 + k < kr1) ? A[(long long)(m0 + m) * d.lda + k0 + k] : 0.f;
      }
      __syncthreads();
      const int kb = k0 + warp * 16;
      if (kb < kr1) {
        // all 16 weight loads of this warp are in flight before the first FMA
        float wv[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) wv[q] = (cok && kb + q < kr1) ? __ldg(W + (long long)(kb + q) * N + col) : 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4* a4 = reinterpret_cast<const float4*>(&As[warp * 16 + q][0]);
#pragma unroll
          for (int r = 0; r < MR / 4; ++r) {
            const float4 a = a4[r];
            acc[4 * r + 0] = fmaf(a.x, wv[q], acc[4 * r + 0]);
            acc[4 * r + 1] = fmaf(a.y, wv[q], acc[4 * r + 1]);
            acc[4 * r + 2] = fmaf(a.z, wv[q], acc[4 * r + 2]);
            acc[4 * r + 3] = fmaf(a.w, wv[q], acc[4 * r + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) red[warp][m][lane] = acc[m];
    __syncthreads();
    for (int o = tid; o < MR * NC; o += 256) {
      const int m = o / NC, c = o % NC;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][m][c];
      part[o] = s;
    }
    if (KS > 1) cluster.sync(); else __syncthreads();
    if (rank == 0) {
      if (KS > 1) {
        float sv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int o = tid + 256 * h;
          float s = part[o];
#pragma unroll
          for (int r = 1; r < KS; ++r) s += cluster.map_shared_rank(part, r)[o];
          sv[h] = s;
        }
        __syncthreads();
        part[tid] = sv[0];
        part[tid + 256] = sv[1];
        __syncthreads();
      }
      if (!lstm) {
        for (int o = tid; o < MR * NC; o += 256) {
          const int m = o / NC, c = o % NC;
          const int cc = cb * NC + c;
          if (m0 + m < d.M && cc < N) {
            float s = part[o];
            if (d.bias[i]) s += __ldg(d.bias[i] + cc);
            s = apply_act(s, d.act[i]);
            if (d.residual[i]) s += d.residual[i][t * d.res_tstride[i] + (long long)(m0 + m) * d.ldres[i] + cc];
            d.C[i][t * d.c_tstride[i] + (long long)(tt & 1) * d.c_pstride[i] + (long long)(m0 + m) * d.ldc[i] + cc] = s;
          }
        }
      } else if (tid < MR * 8) {
        // ZoneoutLSTMCell pointwise (TF LSTMCell gate order i,j,f,o, A.5; eval-mode zoneout = expectation of the mask, A.6)
        const int m = tid >> 3, uu = tid & 7, H = d.lstm_H;
        const int row = m0 + m, unit = cb * 8 + uu;
        if (row < d.M && unit < H) {
          const float* bs = d.bias[0];
          const float* pr = part + m * NC;
          const float gi = sigmoidf_(pr[uu] + (bs ? __ldg(bs + unit) : 0.f));
          const float gj = tanhf_(pr[8 + uu] + (bs ? __ldg(bs + H + unit) : 0.f));
          const float gf = sigmoidf_(pr[16 + uu] + (bs ? __ldg(bs + 2 * H + unit) : 0.f) + d.forget_bias);
          const float go = sigmoidf_(pr[24 + uu] + (bs ? __ldg(bs + 3 * H + unit) : 0.f));
          const long long si = (long long)row * H + unit;
          const float c_old = d.lstm_c[si], h_old = d.lstm_h[si];
          const float c_new = gf * c_old + gi * gj;
          const float h_new = go * tanhf_(c_new);
          const float h_st = (1.f - d.zh) * h_new + d.zh * h_old;
          d.lstm_c[si] = (1.f - d.zc) * c_new + d.zc * c_old;
          d.lstm_h[si] = h_st;
          // cell output is the un-zoned h (as in the training kernels); the zoned state feeds the next step's input row
          if (d.lstm_out) d.lstm_out[(long long)(tt & 1) * d.out_pstride + (long long)row * d.ld_out + unit] = h_new;
          if (d.lstm_hdst) d.lstm_hdst[(long long)((tt + 1) & 1) * d.hdst_pstride + (long long)row * d.ld_hdst + unit] = h_st;
        }
      }
    }
    if (KS > 1) cluster.sync();     // peers keep their partial tiles alive until rank 0 has read them
    else __syncthreads();
  }
}

// y[col] (+)= sum_k x[k] W[k, col] for this CTA's columns [c0, c0+nc) (nc <= 32); thread = (column lane, 1/8 of the reduction);
// returns the full sums for tid < nc through shared memory `part` [8][32]
__device__ __forceinline__ float tail_matvec(const float* __restrict__ W, int ldw, int K, const float* xs, int c0, int nc, float (*part)[33],
                                             int tid) {
  const int cl = tid & 31, ks = tid >> 5;
  const int kper = (K + 7) / 8, k0 = ks * kper, k1 = min(K, k0 + kper);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  if (cl < nc) {
    const float* wp = W + (long long)k0 * ldw + c0 + cl;
    int k = k0;
    for (; k + 3 < k1; k += 4) {
      const float w0 = __ldg(wp), w1 = __ldg(wp + ldw), w2 = __ldg(wp + 2 * ldw), w3 = __ldg(wp + 3 * ldw);
      acc0 = fmaf(xs[k], w0, acc0); acc1 = fmaf(xs[k + 1], w1, acc1);
      acc2 = fmaf(xs[k + 2], w2, acc2); acc3 = fmaf(xs[k + 3], w3, acc3);
      wp += 4 * ldw;
    }
    for (; k < k1; ++k) { acc0 = fmaf(xs[k], __ldg(wp), acc0); wp += ldw; }
  }
  __syncthreads();                       // previous use of `part` is over
  part[ks][cl] = (acc0 + acc1) + (acc2 + acc3);
  __syncthreads();
  float v = 0.f;
  if (tid < nc) {
#pragma unroll
    for (int q = 0; q < 8; ++q) v += part[q][tid];
  }
  return v;
}

// Attention step: a cluster of 8 CTAs per utterance.  Each CTA computes the energies of 1/8 of the source positions and
// scatters them into every peer's shared memory (DSMEM); after one cluster barrier each CTA holds all energies, repeats the
// (tiny) softmax / forward recursion locally and produces 1/8 of the context columns.
constexpr int ACS = 8;
__global__ void __cluster_dims__(ACS, 1, 1) __launch_bounds__(256) attn_step_k(const satk_attn_step_desc d) {
  extern __shared__ float sm[];
  __shared__ float redbuf[32];
  __shared__ float upart[ACS];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / ACS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_enter();
  const int Tt = d.Tt, B = d.B, A1 = d.A1, A2 = d.A2, M1 = d.M1, M2 = d.M2, AF = d.att_filters;
  const int tt = d.t_ptr ? *d.t_ptr : 0;
  const int len = (int)d.lengths[b];
  const int AFp = AF > 0 ? AF : 1;
  const int PP = (Tt + ACS - 1) / ACS;            // positions per CTA
  const int j0 = min(Tt, rank * PP), j1 = min(Tt, j0 + PP);
  float* qs = sm;                         // [A1+A2]
  float* fS = qs + A1 + A2;               // [PP][AF]
  float* WfS = fS + PP * AFp;             // [AF][A1]
  float* e1 = WfS + AFp * A1;             // [Tt]  (filled by all CTAs of the cluster)
  float* e2 = e1 + Tt;                    // [Tt]
  float* w1 = e2 + Tt;                    // [Tt] weights that build context 1
  float* apv = w1 + Tt;                   // [Tt] previous alignments (location input)
  float* alo = apv + Tt;                  // [Tt] previous alpha
  float* cpart = alo + Tt;                // [groups][cols per CTA]
  const bool loc = d.att_kernel > 0;
  __shared__ float part[8][33];
  const bool forced = d.forced1 != nullptr;     // teacher_forcing_attention.py:28-35: alignments = teacher_alignments[:, index]
  if (forced) {
    for (int j = tid; j < Tt; j += 256) {
      w1[j] = __ldg(d.forced1 + ((long long)tt * B + b) * Tt + j);
      e2[j] = (A2 > 0 && d.forced2) ? __ldg(d.forced2 + ((long long)tt * B + b) * Tt + j) : 0.f;
      if (rank == 0) {
        if (d.align1) d.align1[((long long)tt * B + b) * Tt + j] = w1[j];
        if (A2 > 0 && d.align2) d.align2[((long long)tt * B + b) * Tt + j] = e2[j];
      }
    }
    __syncthreads();
  } else {
  if (d.Wq1) {
    // processed queries q = out1 . Wq (query_layer of both mechanisms): each CTA computes 1/8 of the A1+A2 columns and stores its
    // slice into every peer's shared memory; the cluster barrier below (state reads) also publishes it
    float* xq = cpart + (256 / ((M1 + M2 + ACS - 1) / ACS)) * ((M1 + M2 + ACS - 1) / ACS);   // staging for out1 [q_in], behind cpart
    const float* xr = d.q_x + (long long)(tt & 1) * d.q_x_pstride + (long long)b * d.q_x_ld;
    for (int i = tid; i < d.q_in; i += 256) xq[i] = xr[i];
    __syncthreads();
    const int QP = (A1 + A2 + ACS - 1) / ACS;           // host guarantees QP <= 32 and A1 % QP == 0: a block never straddles the layers
    const int c0 = rank * QP, nc = max(0, min(QP, A1 + A2 - c0));
    const bool second = c0 >= A1;
    const float v = tail_matvec(second ? d.Wq2 : d.Wq1, second ? A2 : A1, d.q_in, xq, second ? (c0 - A1) : c0, nc, part, tid);
    cluster.sync();                         // every peer is running: its shared memory may be written
    if (tid < nc) {
#pragma unroll
      for (int r = 0; r < ACS; ++r) cluster.map_shared_rank(qs, r)[c0 + tid] = v;
    }
  } else {
    for (int i = tid; i < A1 + A2; i += 256) qs[i] = d.q[(long long)b * d.ldq + i];
  }
  for (int i = tid; i < Tt; i += 256) {
    apv[i] = loc ? d.aprev[(long long)b * Tt + i] : 0.f;
    alo[i] = (d.mode == 2) ? d.alpha[(long long)b * Tt + i] : 0.f;
  }
  if (loc)
    for (int i = tid; i < AF * A1; i += 256) WfS[i] = __ldg(d.loc_layer_w + i);
  const float u = (d.mode == 2) ? d.u[b] : 0.f;
  // every CTA has read the recurrent state (aprev / alpha / u) before rank 0 may overwrite it below
  cluster.sync();
  if (loc) {
    // location features f = conv1d(prev alignments) + bias, SAME padding: left pad (k-1)/2 (forward_attention.py:98-100)
    const int pl = (d.att_kernel - 1) / 2;
    for (int i = tid; i < (j1 - j0) * AF; i += 256) {
      const int j = j0 + i / AF, f = i % AF;
      float acc = __ldg(d.loc_conv_b + f);
      for (int k = 0; k < d.att_kernel; ++k) {
        const int jj = j - pl + k;
        if (jj >= 0 && jj < Tt) acc = fmaf(apv[jj], __ldg(d.loc_conv_w + k * AF + f), acc);
      }
      fS[i] = acc;
    }
  }
  __syncthreads();
  // energies of my positions: warp per position, lanes over score channels; result -> every CTA of the cluster
  for (int j = j0 + warp; j < j1; j += 8) {
    const float* kr = d.keys1 + ((long long)j * B + b) * A1;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < A1; c += 32) {
      float s = __ldg(kr + c) + qs[c];
      if (d.b1) s += __ldg(d.b1 + c);
      if (loc)
        for (int f = 0; f < AF; ++f) s = fmaf(fS[(j - j0) * AF + f], WfS[f * A1 + c], s);
      s1 = fmaf(__ldg(d.v1 + c), ftanh_(s), s1);
    }
    s1 = warp_sum(s1);
    if (A2 > 0) {
      const float* k2 = d.keys2 + ((long long)j * B + b) * A2;
      for (int c = lane; c < A2; c += 32) s2 = fmaf(__ldg(d.v2 + c), ftanh_(__ldg(k2 + c) + qs[A1 + c]), s2);
      s2 = warp_sum(s2);
    }
    if (lane < ACS) {
      cluster.map_shared_rank(e1, lane)[j] = s1;
      if (A2 > 0) cluster.map_shared_rank(e2, lane)[j] = s2;
    }
  }
  cluster.sync();
  // masked softmax (scores past the source length are -inf, A.8)
  float mx1 = -INFINITY, mx2 = -INFINITY;
  for (int j = tid; j < len; j += 256) { mx1 = fmaxf(mx1, e1[j]); if (A2 > 0) mx2 = fmaxf(mx2, e2[j]); }
  mx1 = block_max(mx1, redbuf);
  if (A2 > 0) mx2 = block_max(mx2, redbuf);
  float s1 = 0.f, s2 = 0.f;
  for (int j = tid; j < Tt; j += 256) {
    const float p1 = (j < len) ? __expf(e1[j] - mx1) : 0.f;
    e1[j] = p1; s1 += p1;
    if (A2 > 0) { const float p2 = (j < len) ? __expf(e2[j] - mx2) : 0.f; e2[j] = p2; s2 += p2; }
  }
  s1 = block_sum(s1, redbuf);
  if (A2 > 0) s2 = block_sum(s2, redbuf);
  __syncthreads();
  float sa = 0.f;
  for (int j = tid; j < Tt; j += 256) {
    const float a = e1[j] / s1;
    e1[j] = a;                                                     // softmax alignment a_t
    if (A2 > 0) e2[j] = e2[j] / s2;
    float wv = a;
    if (d.mode == 2) {
      const float am1 = (j > 0) ? alo[j - 1] : 0.f;
      wv = ((1.f - u) * alo[j] + u * am1 + 1e-7f) * a;             // forward_attention.py:108-109
      sa += wv;
    }
    w1[j] = wv;
  }
  if (d.mode == 2) {
    sa = block_sum(sa, redbuf);
    __syncthreads();
    for (int j = tid; j < Tt; j += 256) w1[j] = w1[j] / sa;        // :110
  }
  __syncthreads();
  if (rank == 0) {
    // state + history
    for (int j = tid; j < Tt; j += 256) {
      if (loc) d.aprev[(long long)b * Tt + j] = d.cumulative ? (apv[j] + e1[j]) : e1[j];
      if (d.mode == 2) d.alpha[(long long)b * Tt + j] = w1[j];
      if (d.align1) d.align1[((long long)tt * B + b) * Tt + j] = w1[j];
      if (A2 > 0 && d.align2) d.align2[((long long)tt * B + b) * Tt + j] = e2[j];
    }
  }
  }   // !forced
  // my share of the context columns: thread = (position group, column)
  const int MC = M1 + M2;
  const int CP = (MC + ACS - 1) / ACS;           // columns per CTA
  const int groups = 256 / CP;
  const int cl = tid % CP, gq = tid / CP;
  const int c = rank * CP + cl;
  float agent_part = 0.f;
  if (gq < groups && c < MC) {
    const float* wS = (c < M1) ? w1 : e2;
    const float* vb = (c < M1) ? (d.values1 + (long long)b * M1 + c) : (d.values2 + (long long)b * M2 + (c - M1));
    const long long vs = (long long)B * ((c < M1) ? M1 : M2);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = gq;
    for (; j + 3 * groups < len; j += 4 * groups) {
      const float x0 = __ldg(vb + (long long)j * vs), x1 = __ldg(vb + (long long)(j + groups) * vs);
      const float x2 = __ldg(vb + (long long)(j + 2 * groups) * vs), x3 = __ldg(vb + (long long)(j + 3 * groups) * vs);
      a0 = fmaf(wS[j], x0, a0); a1 = fmaf(wS[j + groups], x1, a1);
      a2 = fmaf(wS[j + 2 * groups], x2, a2); a3 = fmaf(wS[j + 3 * groups], x3, a3);
    }
    for (; j < len; j += groups) a0 = fmaf(wS[j], __ldg(vb + (long long)j * vs), a0);
    cpart[gq * CP + cl] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (tid < CP && rank * CP + tid < MC) {
    const int cc = rank * CP + tid;
    float cx = 0.f;
    for (int g2 = 0; g2 < groups; ++g2) cx += cpart[g2 * CP + tid];
    if (d.ctx_dst0) d.ctx_dst0[(long long)((tt + 1) & 1) * d.pstride0 + (long long)b * d.ld0 + cc] = cx;
    if (d.ctx_dst1) d.ctx_dst1[(long long)(tt & 1) * d.pstride1 + (long long)b * d.ld1 + cc] = cx;
    if (d.use_agent && cc < M1) agent_part = cx * __ldg(d.agent_w + cc);
  }
  if (d.mode == 2 && d.use_agent && !forced) {
    // transition agent: u = sigmoid([context1, processed_query1] . W + b)   (forward_attention.py:111-114)
    if (rank == 0)
      for (int cq = tid; cq < A1; cq += 256) agent_part = fmaf(qs[cq], __ldg(d.agent_w + M1 + cq), agent_part);
    agent_part = block_sum(agent_part, redbuf);
    if (tid == 0) cluster.map_shared_rank(upart, 0)[rank] = agent_part;
    cluster.sync();
    if (rank == 0 && tid == 0) {
      float s = 0.f;
      for (int r = 0; r < ACS; ++r) s += upart[r];
      d.u[b] = sigmoidf_(s + __ldg(d.agent_b));
    }
  }
}

// one CTA (256 threads) per (utterance, head): newest query against cached keys/values of steps 0..t
__global__ void __launch_bounds__(256) sa_step_k(const satk_sa_step_desc d) {
  extern __shared__ float ps[];            // [Tmax] scores / probabilities, then [2][dh] partial outputs
  __shared__ float redbuf[32];
  const int b = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_enter();
  const int D = d.D, dh = D / d.heads, B = d.B;
  const int t = *d.t_ptr;
  const int n = t + 1;
  const float scale = rsqrtf((float)dh);
  const float* q = d.q + (long long)b * d.ldq + hd * dh;
  float* opart = ps + d.Tmax;
  for (int j = warp; j < n; j += 8) {
    const float* kr = d.Kc + ((long long)j * B + b) * D + hd * dh;
    float s = 0.f;
    for (int c = lane; c < dh; c += 32) s = fmaf(q[c], kr[c], s);
    s = warp_sum(s);
    if (lane == 0) ps[j] = s * scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 256) mx = fmaxf(mx, ps[j]);
  mx = block_max(mx, redbuf);
  float sum = 0.f;
  for (int j = tid; j < n; j += 256) { const float p = expf(ps[j] - mx); ps[j] = p; sum += p; }
  sum = block_sum(sum, redbuf);
  __syncthreads();
  for (int j = tid; j < n; j += 256) {
    const float p = ps[j] / sum;
    ps[j] = p;
    if (d.probs) d.probs[(((long long)b * d.heads + hd) * d.Tmax + t) * d.Tmax + j] = p;
  }
  __syncthreads();
  // PV: thread = (position group, column of the head)
  const int groups = 256 / dh, c = tid % dh, gq = tid / dh;
  float acc = 0.f;
  for (int j = gq; j < n; j += groups) acc = fmaf(ps[j], d.Vc[((long long)j * B + b) * D + hd * dh + c], acc);
  opart[gq * dh + c] = acc;
  __syncthreads();
  if (tid < dh) {
    float s = 0.f;
    for (int g2 = 0; g2 < groups; ++g2) s += opart[g2 * dh + tid];
    d.out[(long long)b * d.ldo + hd * dh + tid] = s;
  }
}

// Chain of up to three small dense layers for ONE row per cluster of 8 CTAs (decoder pre-net, module.py:1509-1511; speaker variant
// multi_speaker_modules.py:27-32): each layer gives every CTA <= 32 output columns, the layer output is re-assembled in every CTA's
// shared memory through DSMEM, the last layer goes to global memory.  Replaces 2-3 launches of the step.
constexpr int MLP_MAXW = 256;
__global__ void __cluster_dims__(TCS, 1, 1) __launch_bounds__(256) mlp_chain_k(const satk_mlp_chain_desc d) {
  __shared__ float bufs[2][MLP_MAXW];
  __shared__ float part[8][33];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / TCS, tid = threadIdx.x;
  pdl_enter();
  const int tt = d.t_ptr ? *d.t_ptr : 0;
  const float* xr = d.x + (long long)tt * d.x_tstride + (long long)b * d.x_ld;
  for (int i = tid; i < d.K0; i += 256) bufs[0][i] = xr[i];
  __syncthreads();
  int K = d.K0, cur = 0;
  for (int l = 0; l < d.nlayers; ++l) {
    const int N = d.N[l];
    const int CPn = (N + TCS - 1) / TCS;
    const int c0 = rank * CPn, nc = max(0, min(CPn, N - c0));
    float v = tail_matvec(d.W[l], N, K, bufs[cur], c0, nc, part, tid);
    if (l == 0) cluster.sync();               // every peer is running before the first remote store
    if (tid < nc) {
      if (d.bias[l]) v += __ldg(d.bias[l] + c0 + tid);
      v = apply_act(v, d.act[l]);
      if (d.residual[l]) v += d.residual[l][(long long)b * d.ldres[l] + c0 + tid];
      if (l + 1 < d.nlayers) {
#pragma unroll
        for (int r = 0; r < TCS; ++r) cluster.map_shared_rank(&bufs[0][0], r)[(cur ^ 1) * MLP_MAXW + c0 + tid] = v;
      } else {
        d.out[(long long)(tt & 1) * d.out_pstride + (long long)b * d.out_ld + c0 + tid] = v;
      }
    }
    cluster.sync();                           // layer output complete everywhere (and nobody still reads the buffer written next)
    K = N;
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused tail of a decoder step for ONE utterance per cluster of 8 CTAs: everything after LSTM-3 only couples the vectors of one
// utterance (TransformerWrapper over the cached history, rnn_wrappers.py:111-124; OutputAndStopTokenTransparentWrapper, :188-214):
//   hop h:  K/V/Q projections of the newest row (K, V appended to the caches) -> causal attention of the newest query over rows
//           0..t (rows split over the 64 warps of the cluster, online softmax, flash-style combine through DSMEM) -> output
//           projection -> tanh transform + residual;   then the mel and stop projections.
// Every dense layer gives each CTA 32 output columns x 8 reduction slices; the full vector is re-assembled in every CTA's shared
// memory by DSMEM stores + one cluster barrier per layer.  Replaces 5 launches (13 -> 9 per step) and their global round trips.
constexpr int TMAXD = 256;

__global__ void __cluster_dims__(TCS, 1, 1) __launch_bounds__(256) sa_tail_k(const satk_sa_tail_desc d) {
  extern __shared__ float dyn[];           // scores of this CTA's rows: [heads][rows per CTA]
  __shared__ float xin[TMAXD], qv[TMAXD], att[TMAXD], ao[TMAXD], ybuf[TMAXD];
  __shared__ float part[8][33];
  __shared__ float comb[TCS][TMAXD + 16];  // per source CTA: partial attention output [D] + (m, l) per head
  __shared__ float wcomb[8][TMAXD + 16];   // per warp of this CTA
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / TCS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_enter();
  const int D = d.D, heads = d.heads, dh = D / heads, B = d.B;
  const int t = *d.t_ptr;
  const int n = t + 1;
  const int CP = D / TCS;                   // columns of a D-wide layer per CTA (32 for D = 256)
  const int c0 = rank * CP;
  const int rows_cap = (d.Tmax + TCS - 1) / TCS;
  for (int i = tid; i < D; i += 256) xin[i] = d.x[(long long)b * d.ldx + i];
  __syncthreads();
  for (int h = 0; h < d.hops; ++h) {
    // ---- K / V / Q projections of the newest row
    float* Kc = d.Kc[h];
    float* Vc = d.Vc[h];
    const long long rowoff = ((long long)t * B + b) * D;
    float v = tail_matvec(d.Wk[h], D, D, xin, c0, CP, part, tid);
    if (tid < CP) Kc[rowoff + c0 + tid] = v + __ldg(d.bk[h] + c0 + tid);
    v = tail_matvec(d.Wv[h], D, D, xin, c0, CP, part, tid);
    if (tid < CP) Vc[rowoff + c0 + tid] = v + __ldg(d.bv[h] + c0 + tid);
    v = tail_matvec(d.Wq[h], D, D, xin, c0, CP, part, tid);
    if (tid < CP) {
      v += __ldg(d.bq[h] + c0 + tid);
#pragma unroll
      for (int r = 0; r < TCS; ++r) cluster.map_shared_rank(qv, r)[c0 + tid] = v;
    }
    __threadfence();                        // the newest K / V row is read by the other CTAs of the cluster below
    cluster.sync();
    // ---- causal attention of the newest query over rows 0..t: row j belongs to CTA j % 8, warp (j / 8) % 8
    {
      const float scale = rsqrtf((float)dh);
      const int EPL = D / 32;               // elements per lane (8 for D = 256); a lane never straddles two heads (dh % EPL == 0)
      const int hd = (lane * EPL) / dh;
      const int lanes_per_head = dh / EPL;
      float m = -INFINITY, l = 0.f, acc[TMAXD / 32];
#pragma unroll
      for (int e = 0; e < TMAXD / 32; ++e) acc[e] = 0.f;
      float qreg[TMAXD / 32];
#pragma unroll
      for (int e = 0; e < TMAXD / 32; ++e) qreg[e] = (e < EPL) ? qv[lane * EPL + e] : 0.f;
      for (int j = rank + TCS * warp; j < n; j += TCS * 8) {
        const float* kr = Kc + ((long long)j * B + b) * D + lane * EPL;
        const float* vr = Vc + ((long long)j * B + b) * D + lane * EPL;
        float s = 0.f, vv[TMAXD / 32];
#pragma unroll
        for (int e = 0; e < TMAXD / 32; ++e)
          if (e < EPL) { s = fmaf(qreg[e], kr[e], s); vv[e] = vr[e]; }
        for (int o = 1; o < lanes_per_head; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= scale;
        if ((lane % lanes_per_head) == 0) dyn[hd * rows_cap + j / TCS] = s;       // kept for the alignment output
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), pj = __expf(s - mn);
        l = l * corr + pj;
#pragma unroll
        for (int e = 0; e < TMAXD / 32; ++e)
          if (e < EPL) acc[e] = fmaf(pj, vv[e], acc[e] * corr);
        m = mn;
      }
      // combine the 8 warps of this CTA, then the 8 CTAs of the cluster (each lane carries the (m, l) of its head)
#pragma unroll
      for (int e = 0; e < TMAXD / 32; ++e)
        if (e < EPL) wcomb[warp][lane * EPL + e] = acc[e];
      if ((lane % lanes_per_head) == 0) { wcomb[warp][TMAXD + 2 * hd] = m; wcomb[warp][TMAXD + 2 * hd + 1] = l; }
      __syncthreads();
      for (int i = tid; i < D; i += 256) {
        const int hh = i / dh;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, wcomb[w][TMAXD + 2 * hh]);
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float mw = wcomb[w][TMAXD + 2 * hh];
          if (mw > -INFINITY) o = fmaf(__expf(mw - M), wcomb[w][i], o);
        }
#pragma unroll
        for (int r = 0; r < TCS; ++r) cluster.map_shared_rank(&comb[0][0], r)[rank * (TMAXD + 16) + i] = o;
      }
      if (tid < heads) {
        float M = -INFINITY, L = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, wcomb[w][TMAXD + 2 * tid]);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float mw = wcomb[w][TMAXD + 2 * tid];
          if (mw > -INFINITY) L = fmaf(__expf(mw - M), wcomb[w][TMAXD + 2 * tid + 1], L);
        }
#pragma unroll
        for (int r = 0; r < TCS; ++r) {
          float* cr = cluster.map_shared_rank(&comb[0][0], r) + rank * (TMAXD + 16) + TMAXD;
          cr[2 * tid] = M;
          cr[2 * tid + 1] = L;
        }
      }
      cluster.sync();
      __shared__ float ML[2 * 16];          // final (max, sum) per head
      if (tid < heads) {
        float M = -INFINITY, L = 0.f;
#pragma unroll
        for (int r = 0; r < TCS; ++r) M = fmaxf(M, comb[r][TMAXD + 2 * tid]);
#pragma unroll
        for (int r = 0; r < TCS; ++r) {
          const float mr = comb[r][TMAXD + 2 * tid];
          if (mr > -INFINITY) L = fmaf(__expf(mr - M), comb[r][TMAXD + 2 * tid + 1], L);
        }
        ML[2 * tid] = M;
        ML[2 * tid + 1] = L;
      }
      __syncthreads();
      for (int i = tid; i < D; i += 256) {
        const int hh = i / dh;
        const float M = ML[2 * hh], L = ML[2 * hh + 1];
        float o = 0.f;
#pragma unroll
        for (int r = 0; r < TCS; ++r) {
          const float mr = comb[r][TMAXD + 2 * hh];
          if (mr > -INFINITY) o = fmaf(__expf(mr - M), comb[r][i], o);
        }
        att[i] = o / L;
      }
      if (d.probs[h]) {
        // alignment row t (self_attention.py:45-65): this CTA's rows j = rank + 8 i
        float* pr = d.probs[h];
        for (int idx = tid; idx < heads * rows_cap; idx += 256) {
          const int hh = idx / rows_cap, i = idx % rows_cap, j = rank + TCS * i;
          if (j < n) pr[(((long long)b * heads + hh) * d.Tmax + t) * d.Tmax + j] = __expf(dyn[hh * rows_cap + i] - ML[2 * hh]) / ML[2 * hh + 1];
        }
      }
      __syncthreads();
    }
    // ---- output projection, then tanh transform + residual (SelfAttentionTransformer.call, module.py:363-371)
    v = tail_matvec(d.Wo[h], D, D, att, c0, CP, part, tid);
    if (tid < CP) {
      v += __ldg(d.bo[h] + c0 + tid);
#pragma unroll
      for (int r = 0; r < TCS; ++r) cluster.map_shared_rank(ao, r)[c0 + tid] = v;
    }
    cluster.sync();
    v = tail_matvec(d.Wt[h], D, D, ao, c0, CP, part, tid);
    if (tid < CP) {
      v = xin[c0 + tid] + tanhf_(v + __ldg(d.bt[h] + c0 + tid));
#pragma unroll
      for (int r = 0; r < TCS; ++r) cluster.map_shared_rank(ybuf, r)[c0 + tid] = v;
    }
    cluster.sync();
    for (int i = tid; i < D; i += 256) xin[i] = ybuf[i];
    __syncthreads();
    // buffer reuse across hops: a peer writes my qv again before the next hop's first barrier (my last read of qv was before this hop's
    // second barrier), my ybuf only after three more barriers, my comb / ao after at least one — all later than my reads above
  }
  // ---- mel and stop projections (module.py:639-643): NO columns split over the cluster
  {
    const int NO = d.n_out + 1;
    const int CPo = (NO + TCS - 1) / TCS;
    const int o0 = rank * CPo, nc = max(0, min(CPo, NO - o0));
    // columns [0, n_out) come from W_out, column n_out from W_stop (a [D,1] matrix): handled as two matvecs
    const int nc_mel = max(0, min(nc, d.n_out - o0));
    float v = tail_matvec(d.W_out, d.n_out, D, xin, o0, nc_mel, part, tid);
    if (tid < nc_mel) d.mel_dst[(long long)(t + 1) * d.mel_tstride + (long long)b * d.n_out + o0 + tid] = v + __ldg(d.b_out + o0 + tid);
    const bool has_stop = (o0 + nc == NO) && nc > 0;       // the CTA that owns the last column
    v = tail_matvec(d.W_stop, 1, D, xin, 0, has_stop ? 1 : 0, part, tid);
    if (has_stop && tid == 0) d.stop_dst[(long long)t * B + b] = v + __ldg(d.b_stop);
  }
  cluster.sync();     // peers may still be writing into this CTA's shared memory until their last remote store has landed
  if (d.tick_counter && tid == 0) {
    // end-of-step bookkeeping by the LAST CTA of the grid to get here (every CTA read t at its start, so the increment cannot race):
    // StopTokenBasedInferenceHelper.is_finished (sigmoid(stop) > 0.5 for every utterance and t > min_iters), then t += 1
    __threadfence();
    const unsigned prev = atomicAdd(d.tick_counter, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence();
      bool all = d.use_stop != 0;
      if (all)
        for (int bb = 0; bb < B; ++bb) all = all && (((volatile float*)d.stop_dst)[(long long)t * B + bb] > 0.f);
      if (all && t > d.min_iters && *d.done_step < 0) *d.done_step = t;
      *d.tick_counter = 0u;
      *d.tick_t = t + 1;
    }
  }
}

__global__ void tick_k(int* t_ptr, const float* stop, int B, int min_iters, int* done_step) {
  // StopTokenBasedInferenceHelper.is_finished: sigmoid(stop) > 0.5 for EVERY utterance and time > min_iters (helpers.py:103-107)
  pdl_enter();
  const int t = *t_ptr;
  bool all = true;
  if (stop)
    for (int b = threadIdx.x; b < B; b += 32) all = all && (stop[(long long)t * B + b] > 0.f);   // sigmoid(x) > 0.5 <=> x > 0
  all = __all_sync(0xffffffffu, all);
  if (threadIdx.x == 0) {
    if (stop && all && t > min_iters && *done_step < 0) *done_step = t;
    *t_ptr = t + 1;
  }
}

}  // namespace dstep
}  // namespace satk

using namespace satk;

// every decode-step kernel is launched with programmatic stream serialization (see pdl_enter) and, where it uses one, its cluster shape
template <typename Kern, typename... Args>
static int launch_step(Kern kern, dim3 grid, int threads, size_t smem, cudaStream_t st, int cluster, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  SATK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  return SATK_OK;
}

template <int KS>
static int launch_rowgemm(const satk_rowgemm_desc* d, int blocks, cudaStream_t st) {
  return launch_step(dstep::rowgemm_k<KS>, dim3(blocks * KS), 256, 0, st, KS, *d);
}

extern "C" int satk_rowgemm(const satk_rowgemm_desc* d, void* stream) {
  SATK_CHECK_ARG(d->nmat >= 1 && d->nmat <= 3, "satk_rowgemm: nmat=%d out of range (1..3)", d->nmat);
  SATK_CHECK_ARG(d->M >= 1 && d->K >= 1, "satk_rowgemm: M=%d K=%d", d->M, d->K);
  int blocks = 0;
  if (d->lstm_H > 0) {
    SATK_CHECK_ARG(d->nmat == 1 && d->N[0] == 4 * d->lstm_H && d->lstm_H % 8 == 0 && d->lstm_c && d->lstm_h && d->W[0],
                   "satk_rowgemm: LSTM epilogue needs one [K,4H] matrix, H %% 8 == 0 and the c/h state (H=%d N=%d)", d->lstm_H, d->N[0]);
    blocks = d->lstm_H / 8;
  } else {
    for (int i = 0; i < d->nmat; ++i) {
      SATK_CHECK_ARG(d->N[i] >= 1 && d->W[i] && d->C[i], "satk_rowgemm: matrix %d incomplete", i);
      blocks += (d->N[i] + dstep::NC - 1) / dstep::NC;
    }
  }
  // split the reduction over a cluster so that ~128 CTAs stream the weights, each keeping >= 64 rows of K
  int ks = 1;
  while (ks < 8 && blocks * ks * 2 <= 160 && d->K / (ks * 2) >= 64) ks *= 2;
  cudaStream_t st = (cudaStream_t)stream;
  switch (ks) {
    case 1: return launch_rowgemm<1>(d, blocks, st);
    case 2: return launch_rowgemm<2>(d, blocks, st);
    case 4: return launch_rowgemm<4>(d, blocks, st);
    default: return launch_rowgemm<8>(d, blocks, st);
  }
}

extern "C" int satk_attn_step(const satk_attn_step_desc* d, void* stream) {
  SATK_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "satk_attn_step: unknown mode %d", d->mode);
  SATK_CHECK_ARG(d->mode == 0 || d->att_kernel > 0, "satk_attn_step: mode %d needs a location convolution", d->mode);
  SATK_CHECK_ARG(d->B > 0 && d->Tt > 0 && d->A1 > 0 && d->M1 > 0, "satk_attn_step: bad sizes");
  SATK_CHECK_ARG(d->mode != 2 || (d->alpha && d->u), "satk_attn_step: forward attention needs alpha / u state");
  SATK_CHECK_ARG(!d->use_agent || (d->agent_w && d->agent_b), "satk_attn_step: transition agent weights missing");
  const int AF = d->att_filters > 0 ? d->att_filters : 1;
  const int PP = (d->Tt + dstep::ACS - 1) / dstep::ACS;
  const int CP = (d->M1 + d->M2 + dstep::ACS - 1) / dstep::ACS;
  SATK_CHECK_ARG(CP <= 256, "satk_attn_step: memory depth %d too large", d->M1 + d->M2);
  if (d->Wq1) {
    const int QP = (d->A1 + d->A2 + dstep::ACS - 1) / dstep::ACS;
    SATK_CHECK_ARG(QP <= 32 && d->A1 % QP == 0 && (d->A2 == 0 || d->Wq2) && d->q_x && d->q_in > 0,
                   "satk_attn_step: fused query projection needs (A1+A2)/8 <= 32 columns per CTA and A1 a multiple of it (A1=%d A2=%d)", d->A1, d->A2);
  } else {
    SATK_CHECK_ARG(d->q != nullptr, "satk_attn_step: processed queries missing");
  }
  const size_t smem = sizeof(float) * ((size_t)d->A1 + d->A2 + (size_t)PP * AF + (size_t)AF * d->A1 + 5 * (size_t)d->Tt + (size_t)(256 / CP) * CP +
                                       (d->Wq1 ? (size_t)d->q_in : 0));
  SATK_CHECK_ARG(smem <= 200 * 1024, "satk_attn_step: Tt=%d needs %zu B of shared memory", d->Tt, smem);
  if (smem > 48 * 1024) SATK_CUDA(cudaFuncSetAttribute(dstep::attn_step_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_step(dstep::attn_step_k, dim3(d->B * dstep::ACS), 256, smem, (cudaStream_t)stream, 1, *d);   // cluster dims are compiled in
}

extern "C" int satk_sa_step(const satk_sa_step_desc* d, void* stream) {
  SATK_CHECK_ARG(d->heads > 0 && d->D % d->heads == 0, "satk_sa_step: D=%d not divisible by heads=%d", d->D, d->heads);
  const int dh = d->D / d->heads;
  SATK_CHECK_ARG(dh <= 256 && 256 % dh == 0, "satk_sa_step: head depth %d unsupported (must divide 256)", dh);
  SATK_CHECK_ARG(d->t_ptr && d->q && d->Kc && d->Vc && d->out, "satk_sa_step: null pointer");
  const int groups = 256 / dh;
  const size_t smem = sizeof(float) * ((size_t)d->Tmax + (size_t)groups * dh);
  SATK_CHECK_ARG(smem <= 200 * 1024, "satk_sa_step: Tmax=%d too large", d->Tmax);
  if (smem > 48 * 1024) SATK_CUDA(cudaFuncSetAttribute(dstep::sa_step_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_step(dstep::sa_step_k, dim3(d->B, d->heads), 256, smem, (cudaStream_t)stream, 1, *d);
}

extern "C" int satk_mlp_chain(const satk_mlp_chain_desc* d, void* stream) {
  SATK_CHECK_ARG(d->nlayers >= 1 && d->nlayers <= 3 && d->B > 0 && d->x && d->out, "satk_mlp_chain: bad arguments");
  SATK_CHECK_ARG(d->K0 > 0 && d->K0 <= dstep::MLP_MAXW, "satk_mlp_chain: input width %d unsupported (<= %d)", d->K0, dstep::MLP_MAXW);
  for (int l = 0; l < d->nlayers; ++l)
    SATK_CHECK_ARG(d->W[l] && d->N[l] > 0 && d->N[l] <= dstep::MLP_MAXW, "satk_mlp_chain: layer %d width %d unsupported (<= %d)", l, d->N[l],
                   dstep::MLP_MAXW);
  return launch_step(dstep::mlp_chain_k, dim3(d->B * dstep::TCS), 256, 0, (cudaStream_t)stream, 1, *d);
}

extern "C" int satk_sa_tail(const satk_sa_tail_desc* d, void* stream) {
  SATK_CHECK_ARG(d->hops >= 0 && d->hops <= 4, "satk_sa_tail: hops=%d out of range (0..4)", d->hops);
  SATK_CHECK_ARG(d->D == dstep::TMAXD, "satk_sa_tail: D=%d unsupported (the fused tail is built for %d units)", d->D, dstep::TMAXD);
  SATK_CHECK_ARG(d->n_out + 1 <= 32 * dstep::TCS, "satk_sa_tail: n_out=%d too wide", d->n_out);
  SATK_CHECK_ARG(d->heads > 0 && d->heads <= 16 && d->D % d->heads == 0 && (d->D / d->heads) % (d->D / 32) == 0,
                 "satk_sa_tail: heads=%d does not divide D=%d into lane-aligned heads", d->heads, d->D);
  SATK_CHECK_ARG(d->t_ptr && d->x && d->W_out && d->W_stop && d->mel_dst && d->stop_dst && d->n_out > 0, "satk_sa_tail: null pointer");
  const size_t smem = sizeof(float) * (size_t)d->heads * ((d->Tmax + dstep::TCS - 1) / dstep::TCS);
  SATK_CHECK_ARG(smem <= 64 * 1024, "satk_sa_tail: Tmax=%d too large", d->Tmax);
  if (smem > 8 * 1024) SATK_CUDA(cudaFuncSetAttribute(dstep::sa_tail_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
  return launch_step(dstep::sa_tail_k, dim3(d->B * dstep::TCS), 256, smem, (cudaStream_t)stream, 1, *d);
}

extern "C" int satk_decode_tick(int* t_ptr, const float* stop, int B, int min_iters, int* done_step, void* stream) {
  SATK_CHECK_ARG(t_ptr && done_step, "satk_decode_tick: null pointer");
  return launch_step(dstep::tick_k, dim3(1), 32, 0, (cudaStream_t)stream, 1, t_ptr, stop, B, min_iters, done_step);
}
