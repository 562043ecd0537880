// Attention-RNN forward (sm_100a): LSTM-1 + attention mechanism(s) for all Td decoder steps in ONE
// launch.  A cluster of 16 CTAs owns 4 utterances:
//   P1  gate GEMM  [4 x (ctx+h)] x [(ctx+h) x 64 cols]  — each CTA owns 16 hidden units, its slice of
//       the recurrent kernel stays in registers for all steps; LSTM pointwise + zoneout;
//       h -> every CTA (DSMEM), out1 -> the 4 CTAs of that utterance.
//   P2  each utterance is served by 4 CTAs that split the SCORE CHANNELS: query slice, location
//       features, partial energies sum_c v_c tanh(keys + q + f.Wf) over their 56(+8) channels for all
//       Tt positions; partial energies -> the 4 CTAs of the utterance (DSMEM).
//   P3  masked softmax, forward-attention recursion (one warp per mechanism), context slice
//       (64(+8) value columns) -> every CTA (DSMEM) for the next step's gate GEMM.
// Three hardware cluster barriers per step, no global synchronisation, keys/values/weights never
// re-read from HBM.  Reference semantics: forward_attention.py:88-136 (+ :13-26), TF
// BahdanauAttention / AttentionWrapper, ZoneoutLSTMCell (A.5, A.6).
#include "attn_rnn.cuh"

namespace satk {
namespace arnn {

template <bool HAS2>
struct FwdSmem {
  using D = Dims<HAS2>;
  int TtP;
  float *xrec, *gsm, *out1buf, *Wqs, *keyS, *valS, *fS, *Wfs, *wconv, *bconv, *vs, *qs, *qpart, *epart, *aprev, *alphaS,
      *w1S, *w2S, *cpart;
  __host__ __device__ size_t carve(float* base, int Tt) {
    TtP = tt_pad(Tt);
    float* p = base;
    xrec = p; p += 2 * D::KREC * BG;
    gsm = p; p += BG * 64;
    out1buf = p; p += H;
    Wqs = p; p += H * QC;
    keyS = p; p += (size_t)TtP * KS;
    valS = p; p += (size_t)TtP * KS;
    fS = p; p += (size_t)TtP * MAXF;
    Wfs = p; p += MAXF * QC;
    wconv = p; p += MAXK * MAXF;
    bconv = p; p += MAXF;
    vs = p; p += QC;
    qs = p; p += QC;
    qpart = p; p += 8 * QC;
    epart = p; p += 2 * 4 * (size_t)TtP;
    aprev = p; p += TtP + 2 * HALO;
    alphaS = p; p += TtP;
    w1S = p; p += TtP;
    w2S = p; p += TtP;
    cpart = p; p += 8 * VC;
    return (size_t)(p - base) * sizeof(float);
  }
};

template <bool HAS2, int AFT, int NP>
__global__ void __launch_bounds__(NT, 1) attn_rnn_fwd_kernel(const satk_attn_rnn_fwd_desc d) {
  using D = Dims<HAS2>;
  constexpr int KREC = D::KREC, KPT = D::KPT, A1Q = D::A1Q, NI1 = D::NI1;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * BG;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B;

  extern __shared__ __align__(16) float smem_raw[];
  FwdSmem<HAS2> S;
  S.carve(smem_raw, Tt);
  const int TtP = S.TtP;

  // attention role of this CTA
  const int ab = rank >> 2, cq = rank & 3;
  const int arow = b0 + ab;
  const bool arow_ok = arow < B;
  const int alen = arow_ok ? (int)d.lengths[arow] : 0;
  const int pl = d.att_kernel > 0 ? (d.att_kernel - 1) / 2 : 0;

  // ---------------- one-time loads
  for (int i = tid; i < H * QC; i += NT) {
    int u = i / QC, c = i % QC;
    float v = 0.f;
    if (c < A1Q) v = __ldg(d.Wq1 + (long long)u * d.A1 + cq * A1Q + c);
    else if (HAS2) v = __ldg(d.Wq2 + (long long)u * d.A2 + cq * 8 + (c - A1Q));
    S.Wqs[i] = v;
  }
  for (int i = tid; i < TtP * KS; i += NT) {
    int j = i / KS, c = i % KS;
    float kv = 0.f, vv = 0.f;
    if (arow_ok && j < Tt) {
      long long rowi = (long long)j * B + arow;  // time-major memory
      if (c < A1Q) kv = __ldg(d.keys1 + rowi * d.A1 + cq * A1Q + c) + (d.b1 ? __ldg(d.b1 + cq * A1Q + c) : 0.f);
      else if (HAS2 && c < QC) kv = __ldg(d.keys2 + rowi * d.A2 + cq * 8 + (c - A1Q));
      if (c < 64) vv = __ldg(d.values1 + rowi * M1 + cq * 64 + c);
      else if (HAS2) vv = __ldg(d.values2 + rowi * D::M2 + cq * 8 + (c - 64));
    }
    S.keyS[i] = kv;
    S.valS[i] = vv;
  }
  for (int i = tid; i < MAXF * QC; i += NT) {
    int f = i / QC, c = i % QC;
    S.Wfs[i] = (f < d.att_filters && c < A1Q && d.att_kernel > 0) ? __ldg(d.loc_layer_w + (long long)f * d.A1 + cq * A1Q + c) : 0.f;
  }
  for (int i = tid; i < MAXK * MAXF; i += NT) {
    int k = i / MAXF, f = i % MAXF;
    S.wconv[i] = (k < d.att_kernel && f < d.att_filters) ? __ldg(d.loc_conv_w + k * d.att_filters + f) : 0.f;
  }
  if (tid < MAXF) S.bconv[tid] = (tid < d.att_filters && d.att_kernel > 0) ? __ldg(d.loc_conv_b + tid) : 0.f;
  if (tid < QC) {
    float v = 0.f;
    if (tid < A1Q) v = __ldg(d.v1 + cq * A1Q + tid);
    else if (HAS2) v = __ldg(d.v2 + cq * 8 + (tid - A1Q));
    S.vs[tid] = v;
  }
  for (int i = tid; i < 2 * KREC * BG; i += NT) S.xrec[i] = 0.f;
  for (int i = tid; i < TtP + 2 * HALO; i += NT) S.aprev[i] = 0.f;
  for (int i = tid; i < TtP; i += NT) {
    S.alphaS[i] = (d.mode == 2 && i == 0) ? 1.f : 0.f;  // alpha_0 = one-hot(0), forward_attention.py:131-133
    S.w1S[i] = 0.f;
    S.w2S[i] = 0.f;
  }
  for (int i = tid; i < TtP * MAXF; i += NT) S.fS[i] = 0.f;

  // ---------------- P1 role: thread = (col 0..63, kq 0..7)
  const int col = tid >> 3, kq = tid & 7;
  const int gate = col >> 4, unit = col & 15;
  const int gcol = gate * H + rank * UH + unit;
  float w[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) w[i] = __ldg(d.Wrec + (long long)(kq + 8 * i) * (4 * H) + gcol);
  const int xb = b0 + (kq & 3);
  const bool xb_ok = (kq < 4) && xb < B;
  // pointwise role: tid < 64 -> (pb, pu)
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < B;
  const int pidx = rank * UH + pu;
  float c_st = 0.f, h_st = 0.f;

  // P2 role: position group / channel lane
  const int pg = warp * 4 + (lane >> 3), cl = lane & 7;

  cluster.sync();

  float xg_next = xb_ok ? __ldg(d.xg + ((long long)0 * B + xb) * (4 * H) + gcol) : 0.f;

  for (int t = 0; t < d.Td; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    // ======================= P1: gates + LSTM cell
    {
      const float xg_cur = xg_next;
      if (t + 1 < d.Td && xb_ok) xg_next = __ldg(d.xg + ((long long)(t + 1) * B + xb) * (4 * H) + gcol);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* xr = S.xrec + cur * KREC * BG;
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + (kq + 8 * i) * BG);
        a0 = fmaf(w[i], xv.x, a0);
        a1 = fmaf(w[i], xv.y, a1);
        a2 = fmaf(w[i], xv.z, a2);
        a3 = fmaf(w[i], xv.w, a3);
      }
#pragma unroll
      for (int o = 1; o <= 4; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        a3 += __shfl_xor_sync(0xffffffffu, a3, o);
      }
      if (kq < 4) S.gsm[kq * 64 + col] = ((kq == 0) ? a0 : (kq == 1) ? a1 : (kq == 2) ? a2 : a3) + xg_cur;
    }
    __syncthreads();
    if (tid < 64 && prow_ok) {
      float gi = fsigmoid(S.gsm[pb * 64 + 0 * 16 + pu]);
      float gj = ftanh(S.gsm[pb * 64 + 1 * 16 + pu]);
      float gf = fsigmoid(S.gsm[pb * 64 + 2 * 16 + pu] + d.forget_bias);
      float go = fsigmoid(S.gsm[pb * 64 + 3 * 16 + pu]);
      float c_new = gf * c_st + gi * gj;
      float h_new = go * ftanh(c_new);
      const long long o1 = ((long long)t * B + prow) * H + pidx;
      if (d.gates) {
        const long long o4 = ((long long)t * B + prow) * (4 * H) + pidx;
        d.gates[o4] = gi; d.gates[o4 + H] = gj; d.gates[o4 + 2 * H] = gf; d.gates[o4 + 3 * H] = go;
        d.c_prev[o1] = c_st;
        d.h_prev[o1] = h_st;
      }
      float mc = d.mask_c ? (float)d.mask_c[o1] : (1.f - d.zc);
      float mh = d.mask_h ? (float)d.mask_h[o1] : (1.f - d.zh);
      c_st = c_st + mc * (c_new - c_st);
      h_st = h_st + mh * (h_new - h_st);
      d.x2[((long long)t * B + prow) * D::X2W + pidx] = h_new;
#pragma unroll 4
      for (int r = 0; r < CS; ++r) {
        float* rx = cluster.map_shared_rank(S.xrec, r);
        rx[(nxt * KREC + (M1 + D::M2) + pidx) * BG + pb] = h_st;
      }
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        float* ro = cluster.map_shared_rank(S.out1buf, pb * 4 + r4);
        ro[pidx] = h_new;
      }
    }
    cluster.sync();  // ---- barrier A: out1 / h published

    // ======================= P2: query slice, location features, partial energies
    if (arow_ok) {
      // location features f[j][.] = conv1d(a_prev) (forward_attention.py:98-100)
      if (d.att_kernel > 0) {
        for (int idx = tid; idx < Tt * AFT; idx += NT) {
          int j = idx / AFT, f = idx % AFT;
          float acc = S.bconv[f];
          const float* ap = S.aprev + HALO + j - pl;
          for (int k = 0; k < d.att_kernel; ++k) acc = fmaf(ap[k], S.wconv[k * MAXF + f], acc);
          S.fS[j * MAXF + f] = acc;
        }
      }
      // query slice partials
      {
        const int c = tid & 63, uq = tid >> 6;
        float acc = 0.f;
#pragma unroll 8
        for (int u = uq * 32; u < uq * 32 + 32; ++u) acc = fmaf(S.out1buf[u], S.Wqs[u * QC + c], acc);
        S.qpart[uq * QC + c] = acc;
      }
    }
    __syncthreads();
    if (arow_ok && tid < QC) {
      float q = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) q += S.qpart[u * QC + tid];
      S.qs[tid] = q;
      if (d.q_save) {
        const int qcol = (tid < A1Q) ? (cq * A1Q + tid) : (d.A1 + cq * 8 + (tid - A1Q));
        d.q_save[((long long)t * B + arow) * (d.A1 + d.A2) + qcol] = q;
      }
    }
    __syncthreads();
    if (arow_ok) {
      float e1[NP], e2[NP], fv[NP][AFT];
      int jm[NP];
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        int j = pg + 64 * m;
        jm[m] = (j < Tt) ? j : (Tt - 1);
        e1[m] = 0.f;
        e2[m] = 0.f;
#pragma unroll
        for (int f = 0; f < AFT; ++f) fv[m][f] = S.fS[jm[m] * MAXF + f];
      }
#pragma unroll
      for (int i = 0; i < NI1; ++i) {
        const int c = cl + 8 * i;
        float wf[AFT];
#pragma unroll
        for (int f = 0; f < AFT; ++f) wf[f] = S.Wfs[f * QC + c];
        const float qc = S.qs[c], vc = S.vs[c];
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          float s = S.keyS[jm[m] * KS + c] + qc;
#pragma unroll
          for (int f = 0; f < AFT; ++f) s = fmaf(fv[m][f], wf[f], s);
          e1[m] = fmaf(vc, ftanh(s), e1[m]);
        }
      }
      if (HAS2) {
        const int c = A1Q + cl;
        const float qc = S.qs[c], vc = S.vs[c];
#pragma unroll
        for (int m = 0; m < NP; ++m) e2[m] = vc * ftanh(S.keyS[jm[m] * KS + c] + qc);
      }
#pragma unroll
      for (int m = 0; m < NP; ++m) {
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
          e1[m] += __shfl_xor_sync(0xffffffffu, e1[m], o);
          if (HAS2) e2[m] += __shfl_xor_sync(0xffffffffu, e2[m], o);
        }
      }
      if (cl < 4) {
        // lane cl sends to CTA (ab, cl) of this utterance
        float* re = cluster.map_shared_rank(S.epart, ab * 4 + cl);
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          int j = pg + 64 * m;
          if (j < Tt) {
            re[(0 * 4 + cq) * TtP + j] = e1[m];
            if (HAS2) re[(1 * 4 + cq) * TtP + j] = e2[m];
          }
        }
      }
    }
    cluster.sync();  // ---- barrier B: partial energies published

    // ======================= P3: softmax, forward recursion, context
    if (arow_ok && warp == 0) {
      constexpr int MAXM = 8;
      const int nm = TtP / 32;
      float e[MAXM], ap_[MAXM], apm1[MAXM];
      float mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if (m < nm) {
          int j = lane + 32 * m;
          float v = S.epart[(0 * 4 + 0) * TtP + j] + S.epart[(0 * 4 + 1) * TtP + j] + S.epart[(0 * 4 + 2) * TtP + j] +
                    S.epart[(0 * 4 + 3) * TtP + j];
          e[m] = (j < alen) ? v : -INFINITY;
          mx = fmaxf(mx, e[m]);
          ap_[m] = S.alphaS[j];
          apm1[m] = (j > 0) ? S.alphaS[j - 1] : 0.f;
        }
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          e[m] = (lane + 32 * m < alen) ? __expf(e[m] - mx) : 0.f;
          sum += e[m];
        }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      float asum = 0.f;
      const float u = 0.5f;  // transition factor stays at its initial value without the agent (forward_attention.py:116,135)
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          e[m] *= inv;  // a_t
          if (d.mode == 2) {
            apm1[m] = ((1.f - u) * ap_[m] + u * apm1[m] + 1e-7f) * e[m];  // forward_attention.py:109
            asum += apm1[m];
          }
        }
      __syncwarp();
      float ainv = 1.f;
      if (d.mode == 2) {
        asum = warp_sum(asum);
        ainv = 1.f / asum;
      }
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          int j = lane + 32 * m;
          float a = e[m];
          float wgt = (d.mode == 2) ? apm1[m] * ainv : a;
          if (d.mode == 2) S.alphaS[j] = wgt;
          S.w1S[j] = wgt;
          if (d.att_kernel > 0) S.aprev[HALO + j] = d.cumulative ? (S.aprev[HALO + j] + a) : a;
          if (cq == 0 && j < Tt) {
            const long long oa = ((long long)t * B + arow) * Tt + j;
            d.align1[oa] = wgt;
            if (d.soft1) d.soft1[oa] = a;
          }
        }
    }
    if (HAS2 && arow_ok && warp == 1) {
      constexpr int MAXM = 8;
      const int nm = TtP / 32;
      float e[MAXM];
      float mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          int j = lane + 32 * m;
          float v = S.epart[(1 * 4 + 0) * TtP + j] + S.epart[(1 * 4 + 1) * TtP + j] + S.epart[(1 * 4 + 2) * TtP + j] +
                    S.epart[(1 * 4 + 3) * TtP + j];
          e[m] = (j < alen) ? v : -INFINITY;
          mx = fmaxf(mx, e[m]);
        }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          e[m] = (lane + 32 * m < alen) ? __expf(e[m] - mx) : 0.f;
          sum += e[m];
        }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          int j = lane + 32 * m;
          float a = e[m] * inv;
          S.w2S[j] = a;
          if (cq == 0 && j < Tt) d.align2[((long long)t * B + arow) * Tt + j] = a;
        }
    }
    __syncthreads();
    if (arow_ok) {
      {
        const int c = tid & 63, jg = tid >> 6;
        float acc = 0.f;
        for (int j = jg; j < Tt; j += 8) acc = fmaf(S.w1S[j], S.valS[j * KS + c], acc);
        S.cpart[jg * VC + c] = acc;
      }
      if (HAS2 && tid < 64) {
        const int c2 = tid & 7, jg = tid >> 3;
        float acc = 0.f;
        for (int j = jg; j < Tt; j += 8) acc = fmaf(S.w2S[j], S.valS[j * KS + 64 + c2], acc);
        S.cpart[jg * VC + 64 + c2] = acc;
      }
    }
    __syncthreads();
    if (arow_ok && tid < (HAS2 ? VC : 64)) {
      float cx = 0.f;
#pragma unroll
      for (int jg = 0; jg < 8; ++jg) cx += S.cpart[jg * VC + tid];
      const int k = (tid < 64) ? (cq * 64 + tid) : (M1 + cq * 8 + (tid - 64));
      d.x2[((long long)t * B + arow) * D::X2W + H + k] = cx;
#pragma unroll 4
      for (int r = 0; r < CS; ++r) {
        float* rx = cluster.map_shared_rank(S.xrec, r);
        rx[(nxt * KREC + k) * BG + ab] = cx;
      }
    }
    cluster.sync();  // ---- barrier C: context published
  }
}

template <bool HAS2>
static size_t fwd_smem_bytes(int Tt) {
  FwdSmem<HAS2> S;
  return S.carve(nullptr, Tt);
}

template <typename Kern>
static int launch16(Kern kern, const satk_attn_rnn_fwd_desc& d, size_t smem, cudaStream_t st) {
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((d.B + BG - 1) / BG) * CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SATK_CUDA(cudaLaunchKernelEx(&cfg, kern, d));
  return SATK_OK;
}

int attn_rnn_check(const satk_attn_rnn_fwd_desc* d, bool& has2) {
  SATK_CHECK_ARG(d->H == H && d->M1 == M1, "attn_rnn: H=%d M1=%d unsupported (256/256)", d->H, d->M1);
  has2 = d->A2 > 0;
  if (has2) SATK_CHECK_ARG(d->A1 == 224 && d->A2 == 32 && d->M2 == 32, "attn_rnn: dual attention needs A1=224 A2=32 M2=32 (got %d %d %d)", d->A1, d->A2, d->M2);
  else SATK_CHECK_ARG(d->A1 == 256 && d->M2 == 0, "attn_rnn: single attention needs A1=256 (got %d)", d->A1);
  SATK_CHECK_ARG(d->att_kernel >= 0 && d->att_kernel <= MAXK && d->att_filters <= MAXF, "attn_rnn: location conv %dx%d exceeds %dx%d",
                 d->att_kernel, d->att_filters, MAXK, MAXF);
  SATK_CHECK_ARG(d->mode == 0 || d->att_kernel > 0, "attn_rnn: mode %d needs a location convolution", d->mode);
  SATK_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "attn_rnn: unknown mode %d", d->mode);
  SATK_CHECK_ARG(d->Tt >= 1 && d->Tt <= 256, "attn_rnn: Tt=%d out of range", d->Tt);
  SATK_CHECK_ARG(d->lengths != nullptr, "attn_rnn: lengths required");
  return SATK_OK;
}

int attn_rnn_max_clusters() {
  auto kern = attn_rnn_fwd_kernel<true, 5, 3>;
  size_t smem = fwd_smem_bytes<true>(148);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1; }
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(16 * 64);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

}  // namespace arnn
}  // namespace satk

using namespace satk;
using namespace satk::arnn;

extern "C" int satk_attn_rnn_fwd(const satk_attn_rnn_fwd_desc* d, void* stream) {
  bool has2;
  int rc = attn_rnn_check(d, has2);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int np = (d->Tt + 63) / 64;
  const bool af5 = d->att_filters == 5 || d->att_kernel == 0;
  size_t smem = has2 ? fwd_smem_bytes<true>(d->Tt) : fwd_smem_bytes<false>(d->Tt);
  SATK_CHECK_ARG(smem <= 227 * 1024, "attn_rnn_fwd: Tt=%d needs %zu B of shared memory (> 227 KB)", d->Tt, smem);
#define SATK_ARNN_DISPATCH(H2, AF, NPV) return launch16(attn_rnn_fwd_kernel<H2, AF, NPV>, *d, smem, st)
  if (has2) {
    if (af5) { if (np <= 3) SATK_ARNN_DISPATCH(true, 5, 3); else SATK_ARNN_DISPATCH(true, 5, 4); }
    else { if (np <= 3) SATK_ARNN_DISPATCH(true, 8, 3); else SATK_ARNN_DISPATCH(true, 8, 4); }
  } else {
    if (af5) { if (np <= 3) SATK_ARNN_DISPATCH(false, 5, 3); else SATK_ARNN_DISPATCH(false, 5, 4); }
    else { if (np <= 3) SATK_ARNN_DISPATCH(false, 8, 3); else SATK_ARNN_DISPATCH(false, 8, 4); }
  }
#undef SATK_ARNN_DISPATCH
  return SATK_ERR_INVALID;
}
