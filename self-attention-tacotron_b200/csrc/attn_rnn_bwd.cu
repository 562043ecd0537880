// Attention-RNN backward (BPTT through LSTM-1 + attention mechanisms), one launch for all Td steps.
// Same cluster geometry as the forward kernel (16 CTAs own 4 utterances).  Per step, descending t:
//   BA1  total d(ctx) = external + recurrent carry; partial d(weights) = dctx . values^T over this
//        CTA's value columns                                  -> 4 CTAs of the utterance   [barrier 1]
//   BA2  forward-attention recursion + softmax backward (one warp per mechanism); energy backward
//        over this CTA's score channels (recomputes tanh): dkeys (smem accumulators), dq slice,
//        d(location layer/conv), partial d(state a_{t-1})     -> dq to all CTAs, dstate to 4 [barrier 2]
//   BB   d(out1) = external + dq . Wq^T (unit owners); LSTM cell backward; d(gates)
//                                                              -> all CTAs                 [barrier 3]
//   BC   d(ctx, h)(t-1) = d(gates) . Wrec^T — each CTA owns 34 rows of Wrec (32 in registers, 2 in
//        shared memory)                                        -> owners of ctx columns / units [barrier 4]
// Weight gradients that are dense over time (dWrec, dWq, dW_memory, dvalues) are NOT computed here:
// the kernel saves d(gates), dq and the total d(ctx) and the caller runs plain GEMMs over all steps.
#include "attn_rnn.cuh"

namespace satk {
namespace arnn {

template <bool HAS2>
struct BwdSmem {
  using D = Dims<HAS2>;
  int TtP, Tt8;
  float *keyS, *valS, *dkeyS, *WqU, *dgbuf /* aliases stageW */, *WragS, *fS, *dfS, *Wfs, *wconv, *bconv, *vs, *qs, *aprev,
      *aS, *alphaPrevS, *alphaS, *a2S, *dwpart, *deS, *dstate_part, *dalpha_carry, *dmixS, *dctx_in, *dctxS, *dqB, *dh_in,
      *dout1S, *stageQ;
  __host__ __device__ size_t carve(float* base, int Tt, int stage_floats) {
    TtP = tt_pad(Tt);
    Tt8 = (Tt + 7) / 8 * 8;
    float* p = base;
    keyS = p; p += (size_t)Tt8 * KS;
    valS = p; p += (size_t)Tt8 * KS;
    dkeyS = p; p += (size_t)Tt8 * KS;
    WqU = p; p += UH * 256;
    dgbuf = p; p += (stage_floats > 4 * H * BG) ? stage_floats : 4 * H * BG;  // d(gates) buffer, aliased by the dWf staging area
    WragS = p; p += HAS2 ? 2 * 4 * H : 0;
    fS = p; p += (size_t)TtP * MAXF;
    dfS = p; p += (size_t)(TtP + 2 * HALO) * MAXF;
    Wfs = p; p += MAXF * QC;
    wconv = p; p += MAXK * MAXF;
    bconv = p; p += MAXF;
    vs = p; p += QC;
    qs = p; p += QC;
    aprev = p; p += TtP + 2 * HALO;
    aS = p; p += TtP;
    alphaPrevS = p; p += TtP + 8;
    alphaS = p; p += TtP;
    a2S = p; p += TtP;
    dwpart = p; p += 2 * 4 * (size_t)TtP;
    deS = p; p += 2 * (size_t)TtP;
    dstate_part = p; p += 2 * 4 * (size_t)TtP;
    dalpha_carry = p; p += TtP;
    dmixS = p; p += TtP + 8;
    dctx_in = p; p += VC + 8;
    dctxS = p; p += VC + 8;
    dqB = p; p += BG * 256;
    dh_in = p; p += BG * UH;
    dout1S = p; p += BG * UH;
    stageQ = p; p += 16 * QC;
    return (size_t)(p - base) * sizeof(float);
  }
};

template <bool HAS2, int AFT, int NP>
__global__ void __launch_bounds__(NT, 1) attn_rnn_bwd_kernel(const satk_attn_rnn_bwd_desc dd) {
  using D = Dims<HAS2>;
  constexpr int A1Q = D::A1Q, NI1 = D::NI1, X2W = D::X2W, M2 = D::M2;
  constexpr int NCH = NI1 + (HAS2 ? 1 : 0);  // channels per lane (att1 + att2)
  const satk_attn_rnn_fwd_desc& d = dd.f;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * BG;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, Td = d.Td;
  const int QT = d.A1 + d.A2;

  extern __shared__ __align__(16) float smem_raw[];
  BwdSmem<HAS2> S;
  constexpr int SW = NI1 * AFT;  // staging stride per (warp, channel lane)
  S.carve(smem_raw, Tt, 16 * 8 * SW);
  const int TtP = S.TtP;

  const int ab = rank >> 2, cq = rank & 3;
  const int arow = b0 + ab;
  const bool arow_ok = arow < B;
  const int alen = arow_ok ? (int)d.lengths[arow] : 0;
  const int pl = d.att_kernel > 0 ? (d.att_kernel - 1) / 2 : 0;
  const float u = 0.5f;

  // ---------------- one-time loads
  for (int i = tid; i < S.Tt8 * KS; i += NT) {
    int j = i / KS, c = i % KS;
    float kv = 0.f, vv = 0.f;
    if (arow_ok && j < Tt) {
      long long rowi = (long long)j * B + arow;  // time-major memory
      if (c < A1Q) kv = __ldg(d.keys1 + rowi * d.A1 + cq * A1Q + c) + (d.b1 ? __ldg(d.b1 + cq * A1Q + c) : 0.f);
      else if (HAS2 && c < QC) kv = __ldg(d.keys2 + rowi * d.A2 + cq * 8 + (c - A1Q));
      if (c < 64) vv = __ldg(d.values1 + rowi * M1 + cq * 64 + c);
      else if (HAS2) vv = __ldg(d.values2 + rowi * M2 + cq * 8 + (c - 64));
    }
    S.keyS[i] = kv;
    S.valS[i] = vv;
    S.dkeyS[i] = 0.f;
  }
  for (int i = tid; i < UH * 256; i += NT) {
    int uu = i / 256, c = i % 256;
    float v = 0.f;
    if (c < d.A1) v = __ldg(d.Wq1 + (long long)(rank * UH + uu) * d.A1 + c);
    else if (HAS2 && c < QT) v = __ldg(d.Wq2 + (long long)(rank * UH + uu) * d.A2 + (c - d.A1));
    S.WqU[i] = v;
  }
  if (HAS2)
    for (int i = tid; i < 2 * 4 * H; i += NT) {
      int rr = i / (4 * H), c = i % (4 * H);
      S.WragS[i] = __ldg(d.Wrec + (long long)(512 + 2 * rank + rr) * (4 * H) + c);
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
  for (int i = tid; i < TtP + 2 * HALO; i += NT) S.aprev[i] = 0.f;
  for (int i = tid; i < (TtP + 2 * HALO) * MAXF; i += NT) S.dfS[i] = 0.f;
  for (int i = tid; i < TtP * MAXF; i += NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * 4 * TtP; i += NT) { S.dstate_part[i] = 0.f; S.dwpart[i] = 0.f; }
  for (int i = tid; i < TtP; i += NT) { S.dalpha_carry[i] = 0.f; S.aS[i] = 0.f; S.alphaS[i] = 0.f; S.a2S[i] = 0.f; }
  for (int i = tid; i < TtP + 8; i += NT) { S.alphaPrevS[i] = 0.f; S.dmixS[i] = 0.f; }
  for (int i = tid; i < 2 * TtP; i += NT) S.deS[i] = 0.f;
  if (tid < VC + 8) { S.dctx_in[tid] = 0.f; S.dctxS[tid] = 0.f; }
  if (tid < BG * UH) { S.dh_in[tid] = 0.f; S.dout1S[tid] = 0.f; }
  for (int i = tid; i < BG * 256; i += NT) S.dqB[i] = 0.f;

  // ---------------- BC role: rows [32*rank, 32*rank+32) of Wrec in registers; thread = (kr, cs)
  const int kr = tid >> 4, cs = tid & 15;
  float wr[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) wr[i] = __ldg(d.Wrec + (long long)(32 * rank + kr) * (4 * H) + cs + 16 * i);

  // ---------------- pointwise role
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < B;
  const int pidx = rank * UH + pu;
  float dc = 0.f, dh = 0.f;

  // ---------------- energy role
  const int pg = warp * 4 + (lane >> 3), cl = lane & 7;
  float dv_acc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) dv_acc[i] = 0.f;
  float dWf_acc = 0.f;     // thread (c = tid % 64 .. ) see below: tid < NI1*8*AFT owns (cl_, i_, f_)
  float dwconv_acc = 0.f;  // thread 320 + (k*AFT + f)
  float dbconv_acc = 0.f;  // thread 480 + f

  cluster.sync();

  for (int t = Td - 1; t >= 0; --t) {
    const int par = t & 1;
    // ======================= BA1
    if (arow_ok) {
      if (tid < (HAS2 ? VC : 64)) {
        const int k = (tid < 64) ? (cq * 64 + tid) : (M1 + cq * 8 + (tid - 64));
        float* gp = dd.dx2 + ((long long)t * B + arow) * X2W + H + k;
        float v = S.dctx_in[tid] + *gp;
        *gp = v;  // total d(ctx) for the dense dvalues GEMM
        S.dctxS[tid] = v;
      }
      const long long oa = ((long long)t * B + arow) * Tt;
      for (int j = tid; j < TtP; j += NT) {
        const bool in = j < Tt;
        S.aS[j] = (in && d.soft1) ? __ldg(d.soft1 + oa + j) : 0.f;
        S.alphaS[j] = in ? __ldg(d.align1 + oa + j) : 0.f;
        if (HAS2) S.a2S[j] = in ? __ldg(d.align2 + oa + j) : 0.f;
        float ap = 0.f, alp = (j == 0 && d.mode == 2) ? 1.f : 0.f;
        if (t > 0 && in) {
          if (d.soft1) ap = __ldg(d.soft1 + oa - (long long)B * Tt + j);
          alp = __ldg(d.align1 + oa - (long long)B * Tt + j);
        }
        S.aprev[HALO + j] = ap;
        S.alphaPrevS[j] = alp;
      }
      if (tid < QC) {
        const int qcol = (tid < A1Q) ? (cq * A1Q + tid) : (d.A1 + cq * 8 + (tid - A1Q));
        S.qs[tid] = (tid < A1Q || HAS2) ? __ldg(d.q_save + ((long long)t * B + arow) * QT + qcol) : 0.f;
      }
    }
    __syncthreads();
    if (arow_ok) {
      // partial d(weights): thread = (j = tid>>2 (+128 per pass), part = tid&3)
      const int part = tid & 3;
      float* rw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) rw[q] = cluster.map_shared_rank(S.dwpart, ab * 4 + q);
      for (int j0 = 0; j0 < Tt; j0 += 128) {
        const int j = j0 + (tid >> 2);
        float a1 = 0.f, a2 = 0.f;
        if (j < Tt) {
          const float* vr = S.valS + j * KS;
#pragma unroll
          for (int i = 0; i < 16; ++i) a1 = fmaf(S.dctxS[part * 16 + i], vr[part * 16 + i], a1);
          if (HAS2) {
            a2 = S.dctxS[64 + part * 2] * vr[64 + part * 2] + S.dctxS[64 + part * 2 + 1] * vr[64 + part * 2 + 1];
          }
        }
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        if (HAS2) {
          a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
          a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
        }
        if (j < Tt) {
          // lane `part` sends to CTA (ab, part)
          rw[part][(0 * 4 + cq) * TtP + j] = a1;
          if (HAS2) rw[part][(1 * 4 + cq) * TtP + j] = a2;
        }
      }
      // location features of this step (input: a_{t-1})
      if (d.att_kernel > 0) {
        for (int idx = tid; idx < Tt * AFT; idx += NT) {
          int j = idx / AFT, f = idx % AFT;
          float acc = S.bconv[f];
          const float* ap = S.aprev + HALO + j - pl;
          for (int k = 0; k < d.att_kernel; ++k) acc = fmaf(ap[k], S.wconv[k * MAXF + f], acc);
          S.fS[j * MAXF + f] = acc;
        }
      }
    }
    cluster.sync();  // ---- barrier 1

    // ======================= BA2: recursion / softmax backward
    if (arow_ok && warp == 0) {
      constexpr int MAXM = 8;
      const int nm = TtP / 32;
      float a[MAXM], dal[MAXM], mix[MAXM], dst[MAXM];
      float S_ = 0.f, dot = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          const int j = lane + 32 * m;
          a[m] = S.aS[j];
          float dw = S.dwpart[(0 * 4 + 0) * TtP + j] + S.dwpart[(0 * 4 + 1) * TtP + j] + S.dwpart[(0 * 4 + 2) * TtP + j] +
                     S.dwpart[(0 * 4 + 3) * TtP + j];
          dst[m] = S.dstate_part[(par * 4 + 0) * TtP + j] + S.dstate_part[(par * 4 + 1) * TtP + j] +
                   S.dstate_part[(par * 4 + 2) * TtP + j] + S.dstate_part[(par * 4 + 3) * TtP + j];
          if (t == Td - 1) dst[m] = 0.f;
          if (d.mode == 2) {
            dal[m] = dw + S.dalpha_carry[j];
            const float apm1 = (j > 0) ? S.alphaPrevS[j - 1] : 0.f;
            mix[m] = (1.f - u) * S.alphaPrevS[j] + u * apm1 + 1e-7f;
            S_ += mix[m] * a[m];
            dot += dal[m] * S.alphaS[j];
          } else {
            dal[m] = dw;
          }
        }
      float da[MAXM];
      if (d.mode == 2) {
        S_ = warp_sum(S_);
        dot = warp_sum(dot);
        const float invS = 1.f / S_;
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
          if (m < nm) {
            const int j = lane + 32 * m;
            const float dau = (j < alen) ? (dal[m] - dot) * invS : 0.f;
            da[m] = dau * mix[m] + dst[m];
            S.dmixS[j] = dau * a[m];
          }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
          if (m < nm) {
            const int j = lane + 32 * m;
            const float nx = (j + 1 < TtP) ? S.dmixS[j + 1] : 0.f;
            S.dalpha_carry[j] = (1.f - u) * S.dmixS[j] + u * nx;  // adjoint of the shift (forward_attention.py:108-109)
          }
      } else {
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
          if (m < nm) da[m] = dal[m] + dst[m];
      }
      float dot2 = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) dot2 += da[m] * a[m];
      dot2 = warp_sum(dot2);
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) S.deS[0 * TtP + lane + 32 * m] = a[m] * (da[m] - dot2);
    }
    if (HAS2 && arow_ok && warp == 1) {
      constexpr int MAXM = 8;
      const int nm = TtP / 32;
      float a[MAXM], dw[MAXM];
      float dot = 0.f;
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) {
          const int j = lane + 32 * m;
          a[m] = S.a2S[j];
          dw[m] = S.dwpart[(1 * 4 + 0) * TtP + j] + S.dwpart[(1 * 4 + 1) * TtP + j] + S.dwpart[(1 * 4 + 2) * TtP + j] +
                  S.dwpart[(1 * 4 + 3) * TtP + j];
          dot += dw[m] * a[m];
        }
      dot = warp_sum(dot);
#pragma unroll
      for (int m = 0; m < MAXM; ++m)
        if (m < nm) S.deS[1 * TtP + lane + 32 * m] = a[m] * (dw[m] - dot);
    }
    __syncthreads();

    // ---- energy backward over this CTA's channels
    float* stageW = S.dgbuf;  // aliased: d(gates) buffer is idle during BA2
    if (arow_ok) {
      float fv[NP][AFT], dfp[NP][AFT], de1[NP], de2[NP];
      int jm[NP];
      bool jok[NP];
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const int j = pg + 64 * m;
        jok[m] = j < Tt;
        jm[m] = jok[m] ? j : (Tt - 1);
        de1[m] = jok[m] ? S.deS[0 * TtP + jm[m]] : 0.f;
        de2[m] = (HAS2 && jok[m]) ? S.deS[1 * TtP + jm[m]] : 0.f;
#pragma unroll
        for (int f = 0; f < AFT; ++f) {
          fv[m][f] = S.fS[jm[m] * MAXF + f];
          dfp[m][f] = 0.f;
        }
      }
      float dq_acc[NCH];
#pragma unroll
      for (int i = 0; i < NI1; ++i) {
        const int c = cl + 8 * i;
        float wf[AFT], P[AFT];
#pragma unroll
        for (int f = 0; f < AFT; ++f) { wf[f] = S.Wfs[f * QC + c]; P[f] = 0.f; }
        const float qc = S.qs[c], vc = S.vs[c];
        float dq = 0.f;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          float s = S.keyS[jm[m] * KS + c] + qc;
#pragma unroll
          for (int f = 0; f < AFT; ++f) s = fmaf(fv[m][f], wf[f], s);
          const float th = ftanh(s);
          const float dsv = de1[m] * vc * (1.f - th * th);
          dv_acc[i] = fmaf(de1[m], th, dv_acc[i]);
          if (jok[m]) S.dkeyS[jm[m] * KS + c] += dsv;
          dq += dsv;
#pragma unroll
          for (int f = 0; f < AFT; ++f) {
            P[f] = fmaf(fv[m][f], dsv, P[f]);
            dfp[m][f] = fmaf(dsv, wf[f], dfp[m][f]);
          }
        }
        dq_acc[i] = dq;
        // reduce P over the 4 position lanes of the warp, stage per warp
#pragma unroll
        for (int f = 0; f < AFT; ++f) {
          P[f] += __shfl_xor_sync(0xffffffffu, P[f], 8);
          P[f] += __shfl_xor_sync(0xffffffffu, P[f], 16);
        }
        if (lane < 8) {
#pragma unroll
          for (int f = 0; f < AFT; ++f) stageW[(warp * 8 + cl) * SW + i * AFT + f] = P[f];
        }
      }
      if (HAS2) {
        const int c = A1Q + cl;
        const float qc = S.qs[c], vc = S.vs[c];
        float dq = 0.f;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          const float th = ftanh(S.keyS[jm[m] * KS + c] + qc);
          const float dsv = de2[m] * vc * (1.f - th * th);
          dv_acc[NI1] = fmaf(de2[m], th, dv_acc[NI1]);
          if (jok[m]) S.dkeyS[jm[m] * KS + c] += dsv;
          dq += dsv;
        }
        dq_acc[NI1] = dq;
      }
      // dq: reduce over position lanes, stage per warp
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        dq_acc[i] += __shfl_xor_sync(0xffffffffu, dq_acc[i], 8);
        dq_acc[i] += __shfl_xor_sync(0xffffffffu, dq_acc[i], 16);
        if (lane < 8) S.stageQ[warp * QC + cl + 8 * i] = dq_acc[i];
      }
      // d(location features): reduce over the 8 channel lanes
#pragma unroll
      for (int m = 0; m < NP; ++m) {
#pragma unroll
        for (int f = 0; f < AFT; ++f) {
          float v = dfp[m][f];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          if (cl == 0 && jok[m]) S.dfS[(HALO + jm[m]) * MAXF + f] = v;
        }
      }
    }
    __syncthreads();
    if (arow_ok) {
      if (tid < QC) {
        float q = 0.f;
        if (tid < A1Q || HAS2) {
#pragma unroll
          for (int w_ = 0; w_ < 16; ++w_) q += S.stageQ[w_ * QC + tid];
          const int qcol = (tid < A1Q) ? (cq * A1Q + tid) : (d.A1 + cq * 8 + (tid - A1Q));
          dd.dq[((long long)t * B + arow) * QT + qcol] = q;
#pragma unroll 4
          for (int r = 0; r < CS; ++r) {
            float* rq = cluster.map_shared_rank(S.dqB, r);
            rq[ab * 256 + qcol] = q;
          }
        }
      }
      if (d.att_kernel > 0) {
        if (tid < NI1 * 8 * AFT) {
          // thread owns (cl_, i_, f_) of d(location_features_layer)
          const int cl_ = tid / SW, rem = tid % SW;
          float acc = 0.f;
#pragma unroll
          for (int w_ = 0; w_ < 16; ++w_) acc += stageW[(w_ * 8 + cl_) * SW + rem];
          dWf_acc += acc;
        }
        // partial d(state a_{t-1}) = conv-transpose of d(location features)
        if (t > 0) {
          float* rs[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) rs[q] = cluster.map_shared_rank(S.dstate_part, ab * 4 + q);
          for (int j = tid; j < Tt; j += NT) {
            float acc = 0.f;
            for (int k = 0; k < d.att_kernel; ++k) {
              const float* dfr = S.dfS + (HALO + j - k + pl) * MAXF;  // rows outside [0,Tt) are zero
#pragma unroll
              for (int f = 0; f < AFT; ++f) acc = fmaf(dfr[f], S.wconv[k * MAXF + f], acc);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) rs[q][((par ^ 1) * 4 + cq) * TtP + j] = acc;
          }
        }
        if (tid >= 256 && tid - 256 < d.att_kernel * AFT) {
          // thread owns (k, f) of d(location conv kernel)
          const int e = tid - 256, k = e / AFT, f = e % AFT;
          float acc = 0.f;
          for (int j = 0; j < Tt; ++j) acc = fmaf(S.aprev[HALO + j + k - pl], S.dfS[(HALO + j) * MAXF + f], acc);
          dwconv_acc += acc;
        }
        if (tid >= 480 && tid < 480 + AFT) {
          const int f = tid - 480;
          float acc = 0.f;
          for (int j = 0; j < Tt; ++j) acc += S.dfS[(HALO + j) * MAXF + f];
          dbconv_acc += acc;
        }
      }
    }
    cluster.sync();  // ---- barrier 2: dq, dstate parts published

    // ======================= BB: d(out1), LSTM cell backward
    {
      // thread = (b = tid>>7, u = (tid>>3)&15, part = tid&7): 32 columns of dq each
      const int bb = tid >> 7, uu = (tid >> 3) & 15, part = tid & 7;
      float acc = 0.f;
      const float* qrow = S.dqB + bb * 256 + part * 32;
      const float* wrow = S.WqU + uu * 256 + part * 32;
#pragma unroll 8
      for (int c = 0; c < 32; ++c) acc = fmaf(qrow[c], wrow[c], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (part == 0) S.dout1S[bb * UH + uu] = acc;
    }
    __syncthreads();
    if (tid < 64) {
      float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
      if (prow_ok) {
        dh += S.dh_in[pb * UH + pu];  // recurrent carry delivered by BC of step t+1
        const long long o1 = ((long long)t * B + prow) * H + pidx;
        const long long o4 = ((long long)t * B + prow) * (4 * H) + pidx;
        const float gi = d.gates[o4], gj = d.gates[o4 + H], gf = d.gates[o4 + 2 * H], go = d.gates[o4 + 3 * H];
        const float cp = d.c_prev[o1];
        const float mc = d.mask_c ? (float)d.mask_c[o1] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)d.mask_h[o1] : (1.f - d.zh);
        const float c_new = gf * cp + gi * gj;
        const float tc = ftanh(c_new);
        const float dout1 = dd.dx2[((long long)t * B + prow) * X2W + pidx] + S.dout1S[pb * UH + pu];
        const float dh_new = dout1 + mh * dh;
        dh = (1.f - mh) * dh;
        const float dcn = mc * dc + dh_new * go * (1.f - tc * tc);
        dgo = dh_new * tc * go * (1.f - go);
        dgi = dcn * gj * gi * (1.f - gi);
        dgj = dcn * gi * (1.f - gj * gj);
        dgf = dcn * cp * gf * (1.f - gf);
        dc = (1.f - mc) * dc + dcn * gf;
        dd.dgates[o4] = dgi; dd.dgates[o4 + H] = dgj; dd.dgates[o4 + 2 * H] = dgf; dd.dgates[o4 + 3 * H] = dgo;
      }
#pragma unroll 4
      for (int r = 0; r < CS; ++r) {
        float* base = cluster.map_shared_rank(S.dgbuf, r);
        base[(0 * H + pidx) * BG + pb] = dgi;
        base[(1 * H + pidx) * BG + pb] = dgj;
        base[(2 * H + pidx) * BG + pb] = dgf;
        base[(3 * H + pidx) * BG + pb] = dgo;
      }
    }
    cluster.sync();  // ---- barrier 3: d(gates) published

    // ======================= BC: d(ctx, h)(t-1) = d(gates) . Wrec^T
    if (t > 0) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float4 g4 = *reinterpret_cast<const float4*>(S.dgbuf + (cs + 16 * i) * BG);
        a0 = fmaf(wr[i], g4.x, a0);
        a1 = fmaf(wr[i], g4.y, a1);
        a2 = fmaf(wr[i], g4.z, a2);
        a3 = fmaf(wr[i], g4.w, a3);
      }
#pragma unroll
      for (int o = 1; o <= 8; o <<= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        a3 += __shfl_xor_sync(0xffffffffu, a3, o);
      }
      if (cs < 4) {
        const float v = (cs == 0) ? a0 : (cs == 1) ? a1 : (cs == 2) ? a2 : a3;
        const int k = 32 * rank + kr;  // row of Wrec, < 512
        const int bb = cs;
        if (k < M1) {
          float* dst = cluster.map_shared_rank(S.dctx_in, bb * 4 + (k >> 6));
          dst[k & 63] = v;
        } else if (k < M1 + M2) {
          float* dst = cluster.map_shared_rank(S.dctx_in, bb * 4 + ((k - M1) >> 3));
          dst[64 + ((k - M1) & 7)] = v;
        } else {
          const int un = k - (M1 + M2);
          float* dst = cluster.map_shared_rank(S.dh_in, un >> 4);
          dst[bb * UH + (un & 15)] = v;
        }
      }
      if (HAS2 && warp == 15) {
        // ragged rows 512 + 2*rank + {0,1} (hidden units 224..255), weights in shared memory
        const int rr = lane >> 4, c16 = lane & 15;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
        const float* wrow = S.WragS + rr * 4 * H;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) {
          const float wv = wrow[c16 + 16 * i];
          const float4 g4 = *reinterpret_cast<const float4*>(S.dgbuf + (c16 + 16 * i) * BG);
          r0 = fmaf(wv, g4.x, r0);
          r1 = fmaf(wv, g4.y, r1);
          r2 = fmaf(wv, g4.z, r2);
          r3 = fmaf(wv, g4.w, r3);
        }
#pragma unroll
        for (int o = 1; o <= 8; o <<= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, o);
          r1 += __shfl_xor_sync(0xffffffffu, r1, o);
          r2 += __shfl_xor_sync(0xffffffffu, r2, o);
          r3 += __shfl_xor_sync(0xffffffffu, r3, o);
        }
        if (c16 < 4) {
          const float v = (c16 == 0) ? r0 : (c16 == 1) ? r1 : (c16 == 2) ? r2 : r3;
          const int un = (512 + 2 * rank + rr) - (M1 + M2);
          float* dst = cluster.map_shared_rank(S.dh_in, un >> 4);
          dst[c16 * UH + (un & 15)] = v;
        }
      }
    }
    cluster.sync();  // ---- barrier 4: recurrent carries delivered
  }

  // ---------------- epilogue: flush accumulators
  if (arow_ok) {
    for (int i = tid; i < Tt * QC; i += NT) {
      const int j = i / QC, c = i % QC;
      const float v = S.dkeyS[j * KS + c];
      if (c < A1Q) dd.dkeys1[((long long)j * B + arow) * d.A1 + cq * A1Q + c] = v;
      else if (HAS2) dd.dkeys2[((long long)j * B + arow) * d.A2 + cq * 8 + (c - A1Q)] = v;
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float v = dv_acc[i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) {
        const int c = cl + 8 * i;
        if (i < NI1) atomicAdd(dd.dv1 + cq * A1Q + c, v);
        else atomicAdd(dd.dv2 + cq * 8 + cl, v);
      }
    }
    if (d.att_kernel > 0) {
      if (tid < NI1 * 8 * AFT) {
        const int cl_ = tid / SW, rem = tid % SW;
        const int i_ = rem / AFT, f_ = rem % AFT;
        if (f_ < d.att_filters) atomicAdd(dd.dloc_layer_w + (long long)f_ * d.A1 + cq * A1Q + cl_ + 8 * i_, dWf_acc);
      }
      if (tid >= 256 && tid - 256 < d.att_kernel * AFT) {
        const int e = tid - 256, k = e / AFT, f = e % AFT;
        if (f < d.att_filters) atomicAdd(dd.dloc_conv_w + k * d.att_filters + f, dwconv_acc);
      }
      if (tid >= 480 && tid < 480 + AFT && (tid - 480) < d.att_filters) atomicAdd(dd.dloc_conv_b + (tid - 480), dbconv_acc);
    }
  }
}

template <bool HAS2>
static size_t bwd_smem_bytes(int Tt, int aft) {
  BwdSmem<HAS2> S;
  return S.carve(nullptr, Tt, 16 * 8 * Dims<HAS2>::NI1 * aft);
}

template <typename Kern>
static int launch16b(Kern kern, const satk_attn_rnn_bwd_desc& d, size_t smem, cudaStream_t st) {
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((d.f.B + BG - 1) / BG) * CS);
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

int attn_rnn_check(const satk_attn_rnn_fwd_desc* d, bool& has2);

int attn_bwd_phase_cycles(long long* out16) {
#ifdef SATK_PHASE_TIMING
  SATK_CUDA(cudaMemcpyFromSymbol(out16, satk::g_phase, sizeof(long long) * 16));
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
#endif
  return 0;
}

}  // namespace arnn
}  // namespace satk

using namespace satk;
using namespace satk::arnn;

extern "C" int satk_attn_rnn_bwd(const satk_attn_rnn_bwd_desc* d, void* stream) {
  bool has2;
  int rc = attn_rnn_check(&d->f, has2);
  if (rc) return rc;
  SATK_CHECK_ARG(!d->f.cumulative, "attn_rnn_bwd: cumulative_weights=True is not supported in the backward pass");
  SATK_CHECK_ARG(d->f.gates && d->f.c_prev && d->f.q_save && d->f.soft1,
                 "attn_rnn_bwd: forward must have saved gates/c_prev/q_save/soft1");
  cudaStream_t st = (cudaStream_t)stream;
  const int np = (d->f.Tt + 63) / 64;
  const bool af5 = d->f.att_filters == 5 || d->f.att_kernel == 0;
  size_t smem = has2 ? bwd_smem_bytes<true>(d->f.Tt, af5 ? 5 : 8) : bwd_smem_bytes<false>(d->f.Tt, af5 ? 5 : 8);
  SATK_CHECK_ARG(smem <= 227 * 1024, "attn_rnn_bwd: Tt=%d needs %zu B of shared memory (> 227 KB)", d->f.Tt, smem);
#define SATK_ARNN_DISPATCH(H2, AF, NPV) return launch16b(attn_rnn_bwd_kernel<H2, AF, NPV>, *d, smem, st)
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
