// Free-running decoder step (sm_100a): the kernels behind PREDICT mode (predict_mel.py path).
// One decoder step = a fixed sequence of these launches; every kernel reads the step index t from DEVICE memory, so the
// host captures the sequence once in a CUDA graph and replays it — no per-step host work, no re-attention over the
// history (the reference's TransformerWrapper recomputes self-attention over all past outputs each step, O(T^3) in total,
// rnn_wrappers.py:111-124; here keys/values of past steps are cached and only row t is computed: O(T^2)).
//   rowgemm      skinny dense layers  C[M<=batch, N] = act(A.W + b) (+res): pre-net, LSTM gate rows, query / K / V / Q / O
//                projections, transform, mel + stop projections.  Weights stream from L2 once per launch.
//   lstm_point   ZoneoutLSTMCell pointwise part, inference interpolation (tacotron2 ZoneoutLSTMCell, A.5/A.6)
//   attn_step    ForwardAttention / LocationSensitive / Bahdanau step for one utterance per CTA
//                (forward_attention.py:88-122, :13-26; A.8) incl. the transition agent (:111-114)
//   sa_step      causal scaled-dot-product attention of the newest query over the cached history (self_attention.py:45-65)
//   tick         t += 1 and stop-token bookkeeping (StopTokenBasedInferenceHelper: sigmoid(stop) > 0.5 for the whole batch
//                after min_iters)
#include "common.cuh"

namespace satk {
namespace dstep {

constexpr int MR = 16;     // rows per pass
constexpr int KC = 128;    // reduction chunk staged in shared memory
constexpr int NC = 32;     // columns per CTA (one per lane)

__global__ void __launch_bounds__(256) rowgemm_k(const satk_rowgemm_desc d) {
  __shared__ __align__(16) float As[KC][MR];
  __shared__ float red[8][MR][NC + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // which matrix / column block
  int i = 0, cb = blockIdx.x;
  while (i < d.nmat - 1 && cb >= (d.N[i] + NC - 1) / NC) { cb -= (d.N[i] + NC - 1) / NC; ++i; }
  const int N = d.N[i];
  const long long t = d.t_ptr ? (long long)(*d.t_ptr) : 0;
  const float* __restrict__ A = d.A + t * d.a_tstride;
  const float* __restrict__ W = d.W[i];
  const int col = cb * NC + lane;
  const bool cok = col < N;
  for (int m0 = 0; m0 < d.M; m0 += MR) {
    float acc[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m] = 0.f;
    for (int k0 = 0; k0 < d.K; k0 += KC) {
      __syncthreads();
      for (int idx = tid; idx < MR * KC; idx += 256) {
        const int m = idx % MR, k = idx / MR;
        As[k][m] = (m0 + m < d.M && k0 + k < d.K) ? A[(long long)(m0 + m) * d.lda + k0 + k] : 0.f;
      }
      __syncthreads();
      // warp w owns k = k0 + 16w .. +15: all 16 weight loads are issued before the first FMA
      float wv[16];
      const int kb = k0 + warp * 16;
#pragma unroll
      for (int q = 0; q < 16; ++q) wv[q] = (cok && kb + q < d.K) ? __ldg(W + (long long)(kb + q) * N + col) : 0.f;
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
#pragma unroll
    for (int m = 0; m < MR; ++m) red[warp][m][lane] = acc[m];
    __syncthreads();
    for (int o = tid; o < MR * NC; o += 256) {
      const int m = o / NC, c = o % NC;
      const int cc = cb * NC + c;
      if (m0 + m < d.M && cc < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][m][c];
        if (d.bias[i]) s += __ldg(d.bias[i] + cc);
        s = apply_act(s, d.act[i]);
        if (d.residual[i]) s += d.residual[i][t * d.res_tstride[i] + (long long)(m0 + m) * d.ldres[i] + cc];
        d.C[i][t * d.c_tstride[i] + (long long)(m0 + m) * d.ldc[i] + cc] = s;
      }
    }
  }
}

// gates [B,4H] pre-activation (bias included), order i,j,f,o (TF LSTMCell, A.5)
__global__ void lstm_point_k(const float* __restrict__ gates, float* c, float* h, int B, int H, float zc, float zh, float forget_bias,
                             float* out, long long ld_out, float* hdst, long long ld_h) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  const float* g = gates + (long long)b * 4 * H;
  const float gi = sigmoidf_(g[u]), gj = tanhf_(g[H + u]), gf = sigmoidf_(g[2 * H + u] + forget_bias), go = sigmoidf_(g[3 * H + u]);
  const float c_old = c[idx], h_old = h[idx];
  const float c_new = gf * c_old + gi * gj;
  const float h_new = go * tanhf_(c_new);
  const float c_st = (1.f - zc) * c_new + zc * c_old;       // eval-mode zoneout: expectation of the keep mask (A.6)
  const float h_st = (1.f - zh) * h_new + zh * h_old;
  c[idx] = c_st;
  h[idx] = h_st;
  if (out) out[(long long)b * ld_out + u] = h_new;           // cell output is the un-zoned h (matches the training kernels)
  if (hdst) hdst[(long long)b * ld_h + u] = h_st;
}

// one CTA (512 threads) per utterance
__global__ void __launch_bounds__(512) attn_step_k(const satk_attn_step_desc d) {
  extern __shared__ float sm[];
  __shared__ float redbuf[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, A1 = d.A1, A2 = d.A2, M1 = d.M1, M2 = d.M2, AF = d.att_filters;
  const int t = d.t_ptr ? *d.t_ptr : 0;
  const int len = (int)d.lengths[b];
  float* qs = sm;                         // [A1+A2]
  float* fS = qs + A1 + A2;               // [Tt][AF]
  float* WfS = fS + Tt * (AF > 0 ? AF : 1);   // [AF][A1]
  float* e1 = WfS + (AF > 0 ? AF : 1) * A1;   // [Tt]
  float* e2 = e1 + Tt;                    // [Tt]
  float* w1 = e2 + Tt;                    // [Tt] weights that build context 1
  float* apv = w1 + Tt;                   // [Tt] previous alignments (location input)
  float* alo = apv + Tt;                  // [Tt] previous alpha
  float* cpart = alo + Tt;                // [2][M1+M2]
  const bool loc = d.att_kernel > 0;
  for (int i = tid; i < A1 + A2; i += 512) qs[i] = d.q[(long long)b * d.ldq + i];
  for (int i = tid; i < Tt; i += 512) {
    apv[i] = loc ? d.aprev[(long long)b * Tt + i] : 0.f;
    alo[i] = (d.mode == 2) ? d.alpha[(long long)b * Tt + i] : 0.f;
  }
  if (loc)
    for (int i = tid; i < AF * A1; i += 512) WfS[i] = __ldg(d.loc_layer_w + i);
  __syncthreads();
  if (loc) {
    // location features f = conv1d(prev alignments) + bias, SAME padding: left pad (k-1)/2 (forward_attention.py:98-100)
    const int pl = (d.att_kernel - 1) / 2;
    for (int i = tid; i < Tt * AF; i += 512) {
      const int j = i / AF, f = i % AF;
      float acc = __ldg(d.loc_conv_b + f);
      for (int k = 0; k < d.att_kernel; ++k) {
        const int jj = j - pl + k;
        if (jj >= 0 && jj < Tt) acc = fmaf(apv[jj], __ldg(d.loc_conv_w + k * AF + f), acc);
      }
      fS[i] = acc;
    }
  }
  __syncthreads();
  // energies: warp per position, lanes over score channels
  for (int j = warp; j < Tt; j += 16) {
    const float* kr = d.keys1 + ((long long)j * B + b) * A1;
    float s1 = 0.f;
    for (int c = lane; c < A1; c += 32) {
      float s = __ldg(kr + c) + qs[c];
      if (d.b1) s += __ldg(d.b1 + c);
      if (loc)
        for (int f = 0; f < AF; ++f) s = fmaf(fS[j * AF + f], WfS[f * A1 + c], s);
      s1 = fmaf(__ldg(d.v1 + c), tanhf_(s), s1);
    }
    s1 = warp_sum(s1);
    if (lane == 0) e1[j] = s1;
    if (A2 > 0) {
      const float* k2 = d.keys2 + ((long long)j * B + b) * A2;
      float s2 = 0.f;
      for (int c = lane; c < A2; c += 32) s2 = fmaf(__ldg(d.v2 + c), tanhf_(__ldg(k2 + c) + qs[A1 + c]), s2);
      s2 = warp_sum(s2);
      if (lane == 0) e2[j] = s2;
    }
  }
  __syncthreads();
  // masked softmax (scores past the source length are -inf, A.8), positions strided over the block
  float mx1 = -INFINITY, mx2 = -INFINITY;
  for (int j = tid; j < len; j += 512) { mx1 = fmaxf(mx1, e1[j]); if (A2 > 0) mx2 = fmaxf(mx2, e2[j]); }
  mx1 = block_max(mx1, redbuf);
  if (A2 > 0) mx2 = block_max(mx2, redbuf);
  float s1 = 0.f, s2 = 0.f;
  for (int j = tid; j < Tt; j += 512) {
    const float p1 = (j < len) ? expf(e1[j] - mx1) : 0.f;
    e1[j] = p1; s1 += p1;
    if (A2 > 0) { const float p2 = (j < len) ? expf(e2[j] - mx2) : 0.f; e2[j] = p2; s2 += p2; }
  }
  s1 = block_sum(s1, redbuf);
  if (A2 > 0) s2 = block_sum(s2, redbuf);
  __syncthreads();
  const float u = (d.mode == 2) ? d.u[b] : 0.f;
  float sa = 0.f;
  for (int j = tid; j < Tt; j += 512) {
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
    for (int j = tid; j < Tt; j += 512) w1[j] = w1[j] / sa;        // :110
  }
  __syncthreads();
  // state + history
  for (int j = tid; j < Tt; j += 512) {
    if (loc) d.aprev[(long long)b * Tt + j] = d.cumulative ? (apv[j] + e1[j]) : e1[j];
    if (d.mode == 2) d.alpha[(long long)b * Tt + j] = w1[j];
    if (d.align1) d.align1[((long long)t * B + b) * Tt + j] = w1[j];
    if (A2 > 0 && d.align2) d.align2[((long long)t * B + b) * Tt + j] = e2[j];
  }
  // context vectors: thread = (half of the positions, column)
  const int MC = M1 + M2;
  for (int c0 = 0; c0 < MC; c0 += 256) {
    const int c = c0 + (tid & 255), hf = tid >> 8;
    if (c < MC) {
      const float* wS = (c < M1) ? w1 : e2;
      float acc = 0.f;
      const int jb = hf ? (len + 1) / 2 : 0, je = hf ? len : (len + 1) / 2;
      if (c < M1) for (int j = jb; j < je; ++j) acc = fmaf(wS[j], __ldg(d.values1 + ((long long)j * B + b) * M1 + c), acc);
      else for (int j = jb; j < je; ++j) acc = fmaf(wS[j], __ldg(d.values2 + ((long long)j * B + b) * M2 + (c - M1)), acc);
      cpart[hf * MC + c] = acc;
    }
  }
  __syncthreads();
  for (int c = tid; c < MC; c += 512) {
    const float cx = cpart[c] + cpart[MC + c];
    cpart[c] = cx;
    if (d.ctx_dst0) d.ctx_dst0[(long long)b * d.ld0 + c] = cx;
    if (d.ctx_dst1) d.ctx_dst1[(long long)b * d.ld1 + c] = cx;
  }
  if (d.mode == 2 && d.use_agent) {
    // transition agent: u = sigmoid([context1, processed_query1] . W + b)   (forward_attention.py:111-114)
    __syncthreads();
    float acc = 0.f;
    for (int c = tid; c < M1; c += 512) acc = fmaf(cpart[c], __ldg(d.agent_w + c), acc);
    for (int c = tid; c < A1; c += 512) acc = fmaf(qs[c], __ldg(d.agent_w + M1 + c), acc);
    acc = block_sum(acc, redbuf);
    if (tid == 0) d.u[b] = sigmoidf_(acc + __ldg(d.agent_b));
  }
}

// one CTA (256 threads) per (utterance, head): newest query against cached keys/values of steps 0..t
__global__ void __launch_bounds__(256) sa_step_k(const satk_sa_step_desc d) {
  extern __shared__ float ps[];            // [Tmax] scores / probabilities, then [2][dh] partial outputs
  __shared__ float redbuf[32];
  const int b = blockIdx.x, hd = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

__global__ void tick_k(int* t_ptr, const float* stop, int B, int min_iters, int* done_step) {
  // StopTokenBasedInferenceHelper.is_finished: sigmoid(stop) > 0.5 for EVERY utterance and time > min_iters (helpers.py:103-107)
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

extern "C" int satk_rowgemm(const satk_rowgemm_desc* d, void* stream) {
  SATK_CHECK_ARG(d->nmat >= 1 && d->nmat <= 3, "satk_rowgemm: nmat=%d out of range (1..3)", d->nmat);
  SATK_CHECK_ARG(d->M >= 1 && d->K >= 1, "satk_rowgemm: M=%d K=%d", d->M, d->K);
  int blocks = 0;
  for (int i = 0; i < d->nmat; ++i) {
    SATK_CHECK_ARG(d->N[i] >= 1 && d->W[i] && d->C[i], "satk_rowgemm: matrix %d incomplete", i);
    blocks += (d->N[i] + dstep::NC - 1) / dstep::NC;
  }
  dstep::rowgemm_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(*d);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

extern "C" int satk_lstm_point(const float* gates, float* c, float* h, int B, int H, float zc, float zh, float forget_bias,
                               float* out, long long ld_out, float* hdst, long long ld_h, void* stream) {
  SATK_CHECK_ARG(gates && c && h && B > 0 && H > 0, "satk_lstm_point: bad arguments");
  const int n = B * H;
  dstep::lstm_point_k<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gates, c, h, B, H, zc, zh, forget_bias, out, ld_out, hdst, ld_h);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

extern "C" int satk_attn_step(const satk_attn_step_desc* d, void* stream) {
  SATK_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "satk_attn_step: unknown mode %d", d->mode);
  SATK_CHECK_ARG(d->mode == 0 || d->att_kernel > 0, "satk_attn_step: mode %d needs a location convolution", d->mode);
  SATK_CHECK_ARG(d->B > 0 && d->Tt > 0 && d->A1 > 0 && d->M1 > 0, "satk_attn_step: bad sizes");
  SATK_CHECK_ARG(d->mode != 2 || (d->alpha && d->u), "satk_attn_step: forward attention needs alpha / u state");
  SATK_CHECK_ARG(!d->use_agent || (d->agent_w && d->agent_b), "satk_attn_step: transition agent weights missing");
  const int AF = d->att_filters > 0 ? d->att_filters : 1;
  const size_t smem = sizeof(float) * ((size_t)d->A1 + d->A2 + (size_t)d->Tt * AF + (size_t)AF * d->A1 + 5 * (size_t)d->Tt + 2 * ((size_t)d->M1 + d->M2));
  SATK_CHECK_ARG(smem <= 200 * 1024, "satk_attn_step: Tt=%d needs %zu B of shared memory", d->Tt, smem);
  if (smem > 48 * 1024) SATK_CUDA(cudaFuncSetAttribute(dstep::attn_step_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dstep::attn_step_k<<<d->B, 512, smem, (cudaStream_t)stream>>>(*d);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
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
  dstep::sa_step_k<<<dim3(d->B, d->heads), 256, smem, (cudaStream_t)stream>>>(*d);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

extern "C" int satk_decode_tick(int* t_ptr, const float* stop, int B, int min_iters, int* done_step, void* stream) {
  SATK_CHECK_ARG(t_ptr && done_step, "satk_decode_tick: null pointer");
  dstep::tick_k<<<1, 32, 0, (cudaStream_t)stream>>>(t_ptr, stop, B, min_iters, done_step);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}
