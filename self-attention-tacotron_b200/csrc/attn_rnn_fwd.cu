// Attention-RNN forward (sm_100a): LSTM-1 + attention mechanism(s) for all Td decoder steps in ONE
// launch.  A cluster of 16 CTAs owns 4 utterances:
//   P1  gate GEMM  [4 x (ctx+h)] x [(ctx+h) x 64 cols]  — each CTA owns 16 hidden units, its slice of
//       the recurrent kernel stays in registers for all steps; LSTM pointwise + zoneout;
//       h -> every CTA, out1 -> the 4 CTAs of that utterance.
//   P2  each utterance is served by 4 CTAs that split the SCORE CHANNELS: query slice, location
//       features, partial energies sum_c v_c tanh(keys + q + f.Wf) over their 56(+8) channels for all
//       Tt positions; partial energies -> the 4 CTAs of the utterance.
//   P3  masked softmax, forward-attention recursion (one warp per mechanism), context slice
//       (64(+8) value columns) -> every CTA for the next step's gate GEMM.
// Exchange = st.async / cp.async.bulk into the peers' shared memory completing mbarrier transactions
// (cluster_sync.cuh): no cluster-wide barrier inside the loop, every CTA proceeds as soon as ITS inputs
// have landed.  The warps that issue the exchange touch no global memory (stores queued in front of a
// st.async delay every peer); saved activations and alignments are staged in shared memory and written
// out by other warps; x-projection and zoneout masks arrive through cp.async rings 6 steps ahead.
// Keys / values / weights are never re-read from HBM.
// Reference semantics: forward_attention.py:88-136 (+ :13-26), TF BahdanauAttention / AttentionWrapper
// (SURVEY.md A.7, A.8), ZoneoutLSTMCell (A.5, A.6).
#include "attn_rnn.cuh"
#include "cluster_sync.cuh"

namespace satk {
namespace arnn {

using cl::cp_async4;
using cl::cp_async_commit;
using cl::cp_async_wait;
using cl::st_async_v4;
using cl::RING;
using cl::PFD;

template <bool HAS2>
struct FwdSmem {
  using D = Dims<HAS2>;
  int TtP, Tt4;
  float *xrec, *gsm, *out1buf, *Wqs, *keyS, *valS, *fS, *Wfs, *wconv, *bconv, *vs, *qs, *qpart, *epart, *aprev, *alphaS,
      *w1S, *w2S, *softS, *cpart, *ctxS, *save1, *xg_ring, *red, *upart;
  uint8_t* mk_ring;
  uint64_t* bars;   // [0..1] X, [2..3] O, [4..5] E
  __host__ __device__ size_t carve(float* base, int Tt) {
    TtP = tt_pad(Tt);
    Tt4 = (Tt + 3) & ~3;
    float* p = base;
    xrec = p; p += 2 * BG * D::KREC;                 // [buf][row][k]
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
    epart = p; p += 2 * 4 * (size_t)TtP;             // [att][source CTA of the row][j]
    aprev = p; p += TtP + 2 * HALO;
    alphaS = p; p += TtP;
    w1S = p; p += TtP;
    w2S = p; p += TtP;
    softS = p; p += TtP;
    cpart = p; p += 8 * VC;
    ctxS = p; p += VC + 8;
    save1 = p; p += 7 * 64;
    red = p; p += 32;
    upart = p; p += 2 * 4 + 4;                       // transition agent: [parity][source CTA of the quad] partial sums, [8] = u
    xg_ring = p; p += RING * 256;
    mk_ring = reinterpret_cast<uint8_t*>(p); p += RING * 2 * BG * UH / 4;
    bars = reinterpret_cast<uint64_t*>(p); p += 2 * 6;
    return (size_t)(p - base) * sizeof(float);
  }
};

template <bool HAS2, int AFT, int NP, bool AGENT>
__global__ void __launch_bounds__(NT, 1) attn_rnn_fwd_kernel(const satk_attn_rnn_fwd_desc d) {
  using D = Dims<HAS2>;
  constexpr int KREC = D::KREC, A1Q = D::A1Q, NI1 = D::NI1, M2 = D::M2, X2W = D::X2W;
  constexpr int KPT = KREC / 32;                       // k's per lane in the gate GEMM (17 / 16)
  constexpr int VCW = HAS2 ? VC : 64;                  // context columns produced per CTA
  constexpr int NATT = HAS2 ? 2 : 1;
  // bytes landing on barX per step: h slices of the 15 peers + context slices of every other CTA (+ 3 agent partial sums)
  constexpr uint32_t RX_X = (uint32_t)((CS - 1) * UH * BG + (BG * (M1 + M2) - VCW) + (AGENT ? 3 : 0)) * 4u;
  constexpr uint32_t RX_O = (uint32_t)(CS - 1) * UH * 4u;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * BG;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B;

  extern __shared__ __align__(16) float smem_raw[];
  FwdSmem<HAS2> S;
  S.carve(smem_raw, Tt);
  const int TtP = S.TtP, Tt4 = S.Tt4;
  const uint32_t RX_E = 3u * NATT * (uint32_t)Tt4 * 4u;
  uint64_t* barX = S.bars;
  uint64_t* barO = S.bars + 2;
  uint64_t* barE = S.bars + 4;

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
      else if (HAS2) vv = __ldg(d.values2 + rowi * M2 + cq * 8 + (c - 64));
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
    S.qs[tid] = 0.f;
  }
  for (int i = tid; i < 2 * BG * KREC; i += NT) S.xrec[i] = 0.f;
  for (int i = tid; i < TtP + 2 * HALO; i += NT) S.aprev[i] = 0.f;
  for (int i = tid; i < TtP; i += NT) {
    S.alphaS[i] = (d.mode == 2 && i == 0) ? 1.f : 0.f;  // alpha_0 = one-hot(0), forward_attention.py:131-133
    S.w1S[i] = 0.f;
    S.w2S[i] = 0.f;
    S.softS[i] = 0.f;
  }
  for (int i = tid; i < TtP * MAXF; i += NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * 4 * TtP; i += NT) S.epart[i] = 0.f;
  for (int i = tid; i < RING * 2 * BG * UH; i += NT) S.mk_ring[i] = 0;
  if (tid < H) S.out1buf[tid] = 0.f;
  if (tid < VC + 8) S.ctxS[tid] = 0.f;
  if (tid < 12) S.upart[tid] = (tid == 8) ? 0.5f : 0.f;       // u_0 = 0.5 (forward_attention.py:135)
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) cl::mbar_init(&S.bars[i], 1);
    cl::fence_mbar_init();
  }

  // ---------------- P1 role: thread = (kq = lane, column group = warp): 4 gate columns x the k's congruent to
  // lane mod 32; x[row][k] scalars feed 16 FMAs per k; partial sums reduce-scattered over the lanes.
  float w[KPT][4];
#pragma unroll
  for (int i = 0; i < KPT; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int colc = warp * 4 + c;                                   // CTA-local gate column = gate*16 + unit
      w[i][c] = __ldg(d.Wrec + (long long)(lane + 32 * i) * (4 * H) + (colc >> 4) * H + rank * UH + (colc & 15));
    }
  // lanes < 16 finalise (column warp*4 + (lane>>2), row lane&3)
  const int fcol = warp * 4 + ((lane >> 2) & 3), frow = lane & 3;
  const int fgcol = (fcol >> 4) * H + rank * UH + (fcol & 15);
  const int xb = b0 + frow;
  const bool xb_ok = (lane < 16) && xb < B;
  const int xslot = warp * 16 + (lane & 15);
  // pointwise role: tid < 64 -> (pb, pu); these warps issue the exchange and touch no global memory
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < B;
  const int pidx = rank * UH + pu;
  float c_st = 0.f, h_st = 0.f;
  // saver role A (LSTM activations): tid in [128,192) -> (sb, su)
  const int sb = (tid - 128) >> 4, su = tid & 15;
  const int srow = b0 + sb;
  const bool srow_ok = (tid >= 128 && tid < 192) && srow < B;
  // P2 role: position group / channel lane
  const int pg = warp * 4 + (lane >> 3), cl_ = lane & 7;
  // passes m = 0..nact-1 of this warp touch a position inside the utterance (j = pg + 64 m < alen); positions past the source length
  // have zero attention weight, so the other passes are skipped (warp-uniform)
  const int nact = min(NP, max(0, (alen - 4 * warp + 63) / 64));

  float wa = 0.f;
  if (AGENT) {
    if (tid < 64) wa = __ldg(d.agent_w + cq * 64 + tid);
    else if (tid < 64 + A1Q) wa = __ldg(d.agent_w + M1 + cq * A1Q + (tid - 64));
  }
  const float agent_b = AGENT ? __ldg(d.agent_b) : 0.f;

  cluster.sync();

  auto prefetch = [&](int t) {
    if (t < d.Td) {
      if (xb_ok) cp_async4(&S.xg_ring[(t % RING) * 256 + xslot], d.xg + ((long long)t * B + xb) * (4 * H) + fgcol);
      if (srow_ok && (su & 3) == 0) {
        const long long om = ((long long)t * B + srow) * H + rank * UH + su;
        uint8_t* mr = S.mk_ring + (t % RING) * 2 * BG * UH;
        if (d.mask_c) cp_async4(mr + (0 * BG + sb) * UH + su, d.mask_c + om);
        if (d.mask_h) cp_async4(mr + (1 * BG + sb) * UH + su, d.mask_h + om);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int t = 0; t < PFD; ++t) prefetch(t);

  PT_DECL
#pragma unroll 1
  for (int t = 0; t < d.Td; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    PT(15)
    const uint32_t par = (uint32_t)(t >> 1) & 1u;
    const bool last = (t + 1 == d.Td);
    prefetch(t + PFD);
    cp_async_wait<PFD>();
    if (t > 0) cl::mbar_wait(&barX[cur], (uint32_t)((t - 1) >> 1) & 1u);    // h(t), ctx(t-1) of every peer have landed
    if (AGENT && t > 0 && tid == 32) {
      // transition factor of the previous step from the four partial sums of this utterance's CTAs (forward_attention.py:111-114)
      const float* up = S.upart + cur * 4;
      S.upart[8] = fsigmoid(((up[0] + up[1]) + (up[2] + up[3])) + agent_b);
    }
    if (tid == 0) {
      if (!last) cl::mbar_arrive_expect_tx(&barX[nxt], RX_X);
      cl::mbar_arrive_expect_tx(&barO[cur], RX_O);
      cl::mbar_arrive_expect_tx(&barE[cur], RX_E);
    }
    PT(0)
    // ======================= P1: gates + LSTM cell
    {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      const float* xr = S.xrec + cur * BG * KREC;
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const int k = lane + 32 * i;
        const float x0 = xr[0 * KREC + k], x1 = xr[1 * KREC + k], x2v = xr[2 * KREC + k], x3 = xr[3 * KREC + k];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[c * 4 + 0] = fmaf(w[i][c], x0, acc[c * 4 + 0]);
          acc[c * 4 + 1] = fmaf(w[i][c], x1, acc[c * 4 + 1]);
          acc[c * 4 + 2] = fmaf(w[i][c], x2v, acc[c * 4 + 2]);
          acc[c * 4 + 3] = fmaf(w[i][c], x3, acc[c * 4 + 3]);
        }
      }
      float v = cl::reduce_scatter16(acc, lane);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 16) S.gsm[frow * 64 + fcol] = v + (xb_ok ? S.xg_ring[(t % RING) * 256 + xslot] : 0.f);
    }
    PT(1)
    __syncthreads();
    PT(2)
    if (tid < 64) {
      float gi = 0.f, gj = 0.f, gf = 0.f, go = 0.f, h_new = 0.f;
      const float c_old = c_st, h_old = h_st;
      if (prow_ok) {
        const uint8_t* mr = S.mk_ring + (t % RING) * 2 * BG * UH;
        const float mc = d.mask_c ? (float)mr[(0 * BG + pb) * UH + pu] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mr[(1 * BG + pb) * UH + pu] : (1.f - d.zh);
        gi = fsigmoid(S.gsm[pb * 64 + 0 * 16 + pu]);
        gj = ftanh(S.gsm[pb * 64 + 1 * 16 + pu]);
        gf = fsigmoid(S.gsm[pb * 64 + 2 * 16 + pu] + d.forget_bias);
        go = fsigmoid(S.gsm[pb * 64 + 3 * 16 + pu]);
        const float c_new = gf * c_st + gi * gj;
        h_new = go * ftanh(c_new);
        c_st = c_st + mc * (c_new - c_st);
        h_st = h_st + mh * (h_new - h_st);
      }
      // ---- exchange: 4 consecutive units of a row travel as one 16-byte st.async
      const int l4 = lane & ~3;
      const float hs0 = __shfl_sync(0xffffffffu, h_st, l4), hs1 = __shfl_sync(0xffffffffu, h_st, l4 + 1);
      const float hs2 = __shfl_sync(0xffffffffu, h_st, l4 + 2), hs3 = __shfl_sync(0xffffffffu, h_st, l4 + 3);
      const float hn0 = __shfl_sync(0xffffffffu, h_new, l4), hn1 = __shfl_sync(0xffffffffu, h_new, l4 + 1);
      const float hn2 = __shfl_sync(0xffffffffu, h_new, l4 + 2), hn3 = __shfl_sync(0xffffffffu, h_new, l4 + 3);
      if (!last) S.xrec[(nxt * BG + pb) * KREC + (M1 + M2) + pidx] = h_st;
      S.save1[0 * 64 + tid] = gi; S.save1[1 * 64 + tid] = gj; S.save1[2 * 64 + tid] = gf; S.save1[3 * 64 + tid] = go;
      S.save1[4 * 64 + tid] = c_old; S.save1[5 * 64 + tid] = h_old; S.save1[6 * 64 + tid] = h_new;
      if (pb == ab) S.out1buf[pidx] = h_new;   // own utterance: local copy
      if ((lane & 3) == 0) {
        const int u0 = rank * UH + (pu & ~3);
        if (!last) {
          const uint32_t dsta = cl::smem_u32(&S.xrec[(nxt * BG + pb) * KREC + (M1 + M2) + u0]);
          const uint32_t bara = cl::smem_u32(&barX[nxt]);
#pragma unroll
          for (int r = 0; r < CS; ++r)
            if (r != rank) st_async_v4(cl::mapa(dsta, r), hs0, hs1, hs2, hs3, cl::mapa(bara, r));
        }
        const uint32_t dsto = cl::smem_u32(&S.out1buf[u0]);
        const uint32_t baro = cl::smem_u32(&barO[cur]);
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const int r = pb * 4 + r4;
          if (r != rank) st_async_v4(cl::mapa(dsto, r), hn0, hn1, hn2, hn3, cl::mapa(baro, r));
        }
      }
    }
    else if (d.att_kernel > 0) {
      // warps 2..15 are idle while the two pointwise warps run the cell update and the exchange: they compute the location
      // features f = conv1d(a_{t-1}) of this step now (a_{t-1} is final since the previous step's softmax)
      location_features<AFT>(S.fS, S.aprev, S.wconv, S.bconv, Tt, d.att_kernel, pl, tid - 64, NT - 64);
    }
    PT(3)
    __syncthreads();
    PT(4)
    if (srow_ok) {
      // saver A: LSTM activations of step t -> global (fire and forget, off the exchange warps)
      const int e = tid - 128;
      const int sidx = rank * UH + su;
      const long long o1 = ((long long)t * B + srow) * H + sidx;
      d.x2[((long long)t * B + srow) * X2W + sidx] = S.save1[6 * 64 + e];
      if (d.gates) {
        const long long o4 = ((long long)t * B + srow) * (4 * H) + sidx;
        d.gates[o4] = S.save1[0 * 64 + e];
        d.gates[o4 + H] = S.save1[1 * 64 + e];
        d.gates[o4 + 2 * H] = S.save1[2 * 64 + e];
        d.gates[o4 + 3 * H] = S.save1[3 * 64 + e];
        d.c_prev[o1] = S.save1[4 * 64 + e];
        d.h_prev[o1] = S.save1[5 * 64 + e];
      }
    }
    cl::mbar_wait(&barO[cur], par);   // out1 of my utterance complete
    PT(5)

    // ======================= P2: query slice, location features, partial energies
    {
      // (location features f[j][.] = conv1d(a_prev), forward_attention.py:98-100, were computed beside the pointwise phase)
      // query slice partials
      const int c = tid & 63, uq = tid >> 6;
      float acc = 0.f;
#pragma unroll 8
      for (int u = uq * 32; u < uq * 32 + 32; ++u) acc = fmaf(S.out1buf[u], S.Wqs[u * QC + c], acc);
      S.qpart[uq * QC + c] = acc;
    }
    PT(6)
    __syncthreads();
    if (tid < QC) {
      float q = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) q += S.qpart[u * QC + tid];
      S.qs[tid] = q;
    }
    __syncthreads();
    PT(7)
    {
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
        const int c = cl_ + 8 * i;
        float wf[AFT];
#pragma unroll
        for (int f = 0; f < AFT; ++f) wf[f] = S.Wfs[f * QC + c];
        const float qc = S.qs[c], vc = S.vs[c];
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          if (m < nact) {
            float s = S.keyS[jm[m] * KS + c] + qc;
#pragma unroll
            for (int f = 0; f < AFT; ++f) s = fmaf(fv[m][f], wf[f], s);
            e1[m] = fmaf(vc, ftanh(s), e1[m]);
          }
        }
      }
      if (HAS2) {
        const int c = A1Q + cl_;
        const float qc = S.qs[c], vc = S.vs[c];
#pragma unroll
        for (int m = 0; m < NP; ++m)
          if (m < nact) e2[m] = vc * ftanh(S.keyS[jm[m] * KS + c] + qc);
      }
#pragma unroll
      for (int m = 0; m < NP; ++m) {
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
          e1[m] += __shfl_xor_sync(0xffffffffu, e1[m], o);
          if (HAS2) e2[m] += __shfl_xor_sync(0xffffffffu, e2[m], o);
        }
      }
      if (cl_ == 0) {
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          int j = pg + 64 * m;
          if (j < Tt) {
            S.epart[(0 * 4 + cq) * TtP + j] = e1[m];
            if (HAS2) S.epart[(1 * 4 + cq) * TtP + j] = e2[m];
          }
        }
      }
    }
    PT(8)
    cl::fence_proxy_async();
    __syncthreads();
    PT(9)
    if (tid < 3 * NATT) {
      // partial energies -> the three other CTAs of this utterance: one bulk DSMEM copy per (mechanism, peer)
      const int att = tid / 3, q3 = tid % 3;
      const int r4 = q3 + (q3 >= cq ? 1 : 0);
      const uint32_t src = cl::smem_u32(&S.epart[(att * 4 + cq) * TtP]);
      cl::bulk_copy_to_cta(cl::mapa(src, ab * 4 + r4), src, (uint32_t)Tt4 * 4u, cl::mapa(cl::smem_u32(&barE[cur]), ab * 4 + r4));
    }
    if (tid >= 192 && tid < 192 + QC && arow_ok && d.q_save) {
      // saver: processed queries (for the backward pass)
      const int qi = tid - 192;
      if (qi < A1Q || HAS2) {
        const int qcol = (qi < A1Q) ? (cq * A1Q + qi) : (d.A1 + cq * 8 + (qi - A1Q));
        d.q_save[((long long)t * B + arow) * (d.A1 + d.A2) + qcol] = S.qs[qi];
      }
    }
    cl::mbar_wait(&barE[cur], par);   // partial energies of my utterance complete
    PT(10)

    // ======================= P3: softmax, forward recursion, context
    if (tid >= 256) {
      // attention 1: one position per thread (threads 256..511), reductions over the 8-warp group
      const int j = tid - 256;
      const bool in = j < TtP && arow_ok;
      float e = -INFINITY;
      if (in && j < alen)
        e = S.epart[(0 * 4 + 0) * TtP + j] + S.epart[(0 * 4 + 1) * TtP + j] + S.epart[(0 * 4 + 2) * TtP + j] + S.epart[(0 * 4 + 3) * TtP + j];
      const float mx = cl::group_max(e, S.red, warp & 7, lane, 2);
      const float pexp = (in && j < alen) ? __expf(e - mx) : 0.f;
      const float u = AGENT ? S.upart[8] : 0.5f;  // without the agent the factor stays at its initial value (forward_attention.py:116,135)
      float mixp = 0.f;
      if (d.mode == 2 && in) {
        const float apm1 = (j > 0) ? S.alphaS[j - 1] : 0.f;
        mixp = ((1.f - u) * S.alphaS[j] + u * apm1 + 1e-7f) * pexp;     // forward_attention.py:109 (up to the softmax normaliser)
      }
      float s1 = pexp, s2 = mixp;
      cl::group_sum2(s1, s2, S.red, warp & 7, lane, 2);                  // the barrier inside also orders the alphaS reads above
      if (in) {
        const float a = pexp / s1;                                      // softmax alignment a_t
        const float wgt = (d.mode == 2) ? mixp / s2 : a;                // alpha_t = mix*a / sum(mix*a): the 1/sum(p) cancels
        if (d.mode == 2) S.alphaS[j] = wgt;
        S.w1S[j] = wgt;
        S.softS[j] = a;
        if (d.att_kernel > 0) S.aprev[HALO + j] = d.cumulative ? (S.aprev[HALO + j] + a) : a;
      }
    } else if (HAS2) {
      // attention 2: threads 0..255
      const int j = tid;
      const bool in = j < TtP && arow_ok;
      float e = -INFINITY;
      if (in && j < alen)
        e = S.epart[(1 * 4 + 0) * TtP + j] + S.epart[(1 * 4 + 1) * TtP + j] + S.epart[(1 * 4 + 2) * TtP + j] + S.epart[(1 * 4 + 3) * TtP + j];
      const float mx = cl::group_max(e, S.red + 16, warp & 7, lane, 3);
      float pexp = (in && j < alen) ? __expf(e - mx) : 0.f, dummy = 0.f;
      float s1 = pexp;
      cl::group_sum2(s1, dummy, S.red + 16, warp & 7, lane, 3);
      if (in) S.w2S[j] = pexp / s1;
    }
    PT(11)
    __syncthreads();
    PT(12)
    {
      // context partial sums: thread = (column c of the 64(+8) value columns, position group jg of 7)
      if (tid < VCW * 7) {
        const int c = tid % VCW, jg = tid / VCW;
        const float* wS = (c < 64) ? S.w1S : S.w2S;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        int j = jg;
        for (; j + 21 < alen; j += 28) {        // weights past the source length are exactly zero
          acc0 = fmaf(wS[j], S.valS[j * KS + c], acc0);
          acc1 = fmaf(wS[j + 7], S.valS[(j + 7) * KS + c], acc1);
          acc2 = fmaf(wS[j + 14], S.valS[(j + 14) * KS + c], acc2);
          acc3 = fmaf(wS[j + 21], S.valS[(j + 21) * KS + c], acc3);
        }
        for (; j < alen; j += 7) acc0 = fmaf(wS[j], S.valS[j * KS + c], acc0);
        S.cpart[jg * VC + c] = (acc0 + acc1) + (acc2 + acc3);
      }
    }
    PT(13)
    __syncthreads();
    if (tid < 96) {
      // context slice of my utterance -> every CTA's x[row][k] for the next step (16-byte st.async per 4 columns)
      float cx = 0.f;
      if (tid < VCW) {
#pragma unroll
        for (int jg = 0; jg < 7; ++jg) cx += S.cpart[jg * VC + tid];
        S.ctxS[tid] = cx;
      }
      const int l4 = lane & ~3;
      const float c0 = __shfl_sync(0xffffffffu, cx, l4), c1 = __shfl_sync(0xffffffffu, cx, l4 + 1);
      const float c2 = __shfl_sync(0xffffffffu, cx, l4 + 2), c3 = __shfl_sync(0xffffffffu, cx, l4 + 3);
      if (tid < VCW && !last) {
        const int k = (tid < 64) ? (cq * 64 + tid) : (M1 + cq * 8 + (tid - 64));
        S.xrec[(nxt * BG + ab) * KREC + k] = cx;
        if ((lane & 3) == 0) {
          const uint32_t dsta = cl::smem_u32(&S.xrec[(nxt * BG + ab) * KREC + k]);
          const uint32_t bara = cl::smem_u32(&barX[nxt]);
#pragma unroll
          for (int r = 0; r < CS; ++r)
            if (r != rank) st_async_v4(cl::mapa(dsta, r), c0, c1, c2, c3, cl::mapa(bara, r));
        }
      }
    }
    PT(14)
    __syncthreads();
    if (AGENT && tid < 128 && !last) {
      // partial sum of [context1, processed_query1] . W over this CTA's context columns and score channels -> the quad
      float part = 0.f;
      if (tid < 64) part = S.ctxS[tid] * wa;
      else if (tid < 64 + A1Q) part = S.qs[tid - 64] * wa;
      part = warp_sum(part);
      if (lane == 0) S.red[16 + warp] = part;
      asm volatile("bar.sync 5, 128;" ::: "memory");
      if (tid == 0) {
        const float v = (S.red[16] + S.red[17]) + (S.red[18] + S.red[19]);
        S.upart[nxt * 4 + cq] = v;
        const uint32_t dsta = cl::smem_u32(&S.upart[nxt * 4 + cq]), bara = cl::smem_u32(&barX[nxt]);
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4)
          if (r4 != cq) cl::st_async_f32(cl::mapa(dsta, ab * 4 + r4), v, cl::mapa(bara, ab * 4 + r4));
      }
    }
    if (arow_ok && tid >= 256) {
      // saver B: context and alignments of step t -> global
      const int e = tid - 256;
      if (e < VCW) {
        const int k = (e < 64) ? (cq * 64 + e) : (M1 + cq * 8 + (e - 64));
        d.x2[((long long)t * B + arow) * X2W + H + k] = S.ctxS[e];
      }
      if (AGENT && cq == 0 && e == 0 && d.u_save) d.u_save[(long long)t * B + arow] = S.upart[8];
      if (cq == 0) {
        const long long oa = ((long long)t * B + arow) * Tt;
        for (int j = e; j < Tt; j += 256) {
          d.align1[oa + j] = S.w1S[j];
          if (d.soft1) d.soft1[oa + j] = S.softS[j];
          if (HAS2) d.align2[oa + j] = S.w2S[j];
        }
      }
    }
  }
  PT_FLUSH(d.Td)
  cp_async_wait<0>();
  if (d.state_final && d.att_kernel > 0 && cq == 0 && arow_ok)
    for (int j = tid; j < Tt; j += NT) d.state_final[(long long)arow * Tt + j] = S.aprev[HALO + j];
  cluster.sync();
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
  auto kern = attn_rnn_fwd_kernel<true, 5, 3, false>;
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

int attn_fwd_phase_cycles(long long* out16) {
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

namespace satk { namespace arnn2 {
bool v2_eligible(const satk_attn_rnn_fwd_desc* d);
int attn_rnn2_fwd_launch(const satk_attn_rnn_fwd_desc* d, cudaStream_t st);
} }

extern "C" int satk_attn_rnn_fwd(const satk_attn_rnn_fwd_desc* d, void* stream) {
  bool has2;
  int rc = attn_rnn_check(d, has2);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // dual-source decoder of the shipped configurations: second-generation kernel (one wave at B = 32), attn_rnn2_fwd.cu
  if (arnn2::v2_eligible(d)) return arnn2::attn_rnn2_fwd_launch(d, st);
  const int np = (d->Tt + 63) / 64;
  const bool af5 = d->att_filters == 5 || d->att_kernel == 0;
  const bool agent = d->agent_w != nullptr;
  SATK_CHECK_ARG(!agent || (d->mode == 2 && d->agent_b), "attn_rnn_fwd: the transition agent needs forward attention (mode 2) and its bias");
  size_t smem = has2 ? fwd_smem_bytes<true>(d->Tt) : fwd_smem_bytes<false>(d->Tt);
  SATK_CHECK_ARG(smem <= 227 * 1024, "attn_rnn_fwd: Tt=%d needs %zu B of shared memory (> 227 KB)", d->Tt, smem);
#define SATK_ARNN_DISPATCH(H2, AF, NPV) \
  do { if (agent) return launch16(attn_rnn_fwd_kernel<H2, AF, NPV, true>, *d, smem, st); \
       return launch16(attn_rnn_fwd_kernel<H2, AF, NPV, false>, *d, smem, st); } while (0)
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
