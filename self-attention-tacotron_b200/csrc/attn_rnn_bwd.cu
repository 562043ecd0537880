// Attention-RNN backward (BPTT through LSTM-1 + attention mechanisms), one launch for all Td steps.
// Same cluster geometry as the forward kernel (16 CTAs own 4 utterances).  Per step, descending t:
//   BA1  total d(ctx) = external + recurrent carry; partial d(weights) = dctx . values^T over this
//        CTA's value columns                                  -> 3 sibling CTAs of the utterance  [W]
//   BA2  forward-attention recursion + softmax backward (block-parallel); energy backward over this
//        CTA's score channels (recomputes tanh): dkeys (smem accumulators), dq slice,
//        d(location layer/conv), partial d(state a_{t-1})     -> dq to all CTAs [Q], dstate to siblings [W]
//   BB   d(out1) = external + dq . Wq^T (unit owners); LSTM cell backward; d(gates) -> all CTAs      [G]
//   BC   d(ctx, h)(t-1) = d(gates) . Wrec^T — each CTA owns 34 rows of Wrec (32 in registers, 2 in
//        shared memory)                                        -> owners of ctx columns / units    [C]
// [X] = mbarrier transaction barrier completed by st.async / bulk DSMEM copies (cluster_sync.cuh); there is no
// cluster-wide barrier inside the loop.  Global inputs arrive through cp.async rings two steps ahead; global
// outputs are staged in shared memory and written by warps that issue no exchange traffic.
// Weight gradients that are dense over time (dWrec, dWq, dW_memory, dvalues) are NOT computed here:
// the kernel saves d(gates), dq and the total d(ctx) and the caller runs plain GEMMs over all steps.
#include "attn_rnn.cuh"
#include "cluster_sync.cuh"

namespace satk {
namespace arnn {

using cl::cp_async4;
using cl::cp_async_commit;
using cl::cp_async_wait;
using cl::st_async_v4;

constexpr int RINGB = 4;   // ring slots (time-indexed)
constexpr int PFDB = 2;    // prefetch distance (steps)
constexpr int WQS = 264;   // padded row stride of the per-unit query weights

template <bool HAS2>
struct BwdSmem {
  using D = Dims<HAS2>;
  int TtP, Tt8, Tt4;
  float *keyS, *valS, *dkeyS, *WqU, *dgx /* aliases stageW */, *WragS, *fS, *dfS, *Wfs, *wconv, *bconv, *vs, *aprev, *alphaPrevS,
      *dwpart, *deS, *dstate_part, *dstate_own, *dalpha_carry, *dmixS, *dctx_in, *dctxS, *dqB, *dqS, *dh_in, *dout1S, *stageQ,
      *bcpart, *red, *ringA, *ringB, *save_dg, *dwconvS;
  uint8_t* mk_ring;
  uint64_t* bars;  // [0..1] W, [2..3] Q, [4..5] G, [6..7] C
  __host__ __device__ size_t carve(float* base, int Tt, int stage_floats) {
    TtP = tt_pad(Tt);
    Tt8 = (Tt + 7) / 8 * 8;
    Tt4 = (Tt + 3) & ~3;
    float* p = base;
    keyS = p; p += (size_t)Tt8 * KS;
    valS = p; p += (size_t)Tt8 * KS;
    dkeyS = p; p += (size_t)Tt8 * KS;
    WqU = p; p += UH * WQS;                 // [unit][264]: row stride padded so 4 units x 8 lanes hit 32 distinct banks
    dgx = p; p += (stage_floats > 4 * H * BG) ? stage_floats : 4 * H * BG;  // d(gates) [row][unit][gate]; aliased by the dWf staging
    WragS = p; p += HAS2 ? 2 * 4 * H : 0;
    fS = p; p += (size_t)TtP * MAXF;
    dfS = p; p += (size_t)(TtP + 2 * HALO) * MAXF;   // d(location features), filter-major [f][HALO + j] (conflict-free over j)
    Wfs = p; p += MAXF * QC;
    wconv = p; p += MAXK * MAXF;
    bconv = p; p += MAXF;
    vs = p; p += QC;
    aprev = p; p += TtP + 2 * HALO;
    alphaPrevS = p; p += TtP + 8;
    dwpart = p; p += 2 * 4 * (size_t)TtP;
    deS = p; p += 2 * (size_t)TtP;
    dstate_part = p; p += 2 * 4 * (size_t)TtP;
    dstate_own = p; p += TtP;
    dalpha_carry = p; p += TtP;
    dmixS = p; p += TtP + 8;
    dctx_in = p; p += VC + 8;
    dctxS = p; p += VC + 8;
    dqB = p; p += BG * 256;
    dqS = p; p += QC;
    dh_in = p; p += BG * UH;
    dout1S = p; p += BG * UH;
    stageQ = p; p += 16 * QC;              // also the BC partial-sum area (phase-disjoint)
    bcpart = stageQ;
    red = p; p += 2 * 2 * 8 + 16;          // block-group reduction scratch
    ringA = p; p += RINGB * 3 * (size_t)TtP;               // soft1 / align1 / align2 of time tau
    ringB = p; p += RINGB * (QC + VC + 8 + 6 * 64);        // q_save slice, external d(ctx) slice, pointwise inputs
    save_dg = p; p += 64 * 4;
    dwconvS = p; p += MAXK * MAXF + MAXF;   // d(location conv kernel) [k][f] + d(bias) accumulators
    mk_ring = reinterpret_cast<uint8_t*>(p); p += RINGB * 2 * BG * UH / 4;
    bars = reinterpret_cast<uint64_t*>(p); p += 2 * 8;
    return (size_t)(p - base) * sizeof(float);
  }
};

using cl::group_sum2;

template <bool HAS2, int AFT, int NP, bool AGENT>
__global__ void __launch_bounds__(NT, 1) attn_rnn_bwd_kernel(const satk_attn_rnn_bwd_desc dd) {
  using D = Dims<HAS2>;
  constexpr int A1Q = D::A1Q, NI1 = D::NI1, X2W = D::X2W, M2 = D::M2;
  constexpr int NCH = NI1 + (HAS2 ? 1 : 0);  // channels per lane (att1 + att2)
  constexpr int VCW = HAS2 ? VC : 64;
  constexpr int NATT = HAS2 ? 2 : 1;
  constexpr int RB = QC + VC + 8 + 6 * 64;   // floats per ringB slot
  constexpr uint32_t RX_Q = (uint32_t)(CS - 1) * (A1Q + (HAS2 ? 8 : 0)) * 4u;
  constexpr uint32_t RX_G = (uint32_t)(CS - 1) * UH * BG * 4 * 4u;
  constexpr uint32_t RX_C = (uint32_t)(VCW + UH * BG) * 4u;
  const satk_attn_rnn_fwd_desc& d = dd.f;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * BG;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, Td = d.Td;
  const int QT = d.A1 + d.A2;
  // steps [Te, Td) of this cluster's utterances carry exactly zero gradient (satk_attn_rnn_bwd_desc.step_end): the walk starts at
  // Te - 1.  With 8 clusters on 7 cluster slots a short cluster frees its slot early and the waiting cluster is short as well.
  int Te = Td;
  if (dd.step_end && !d.cumulative) {
    Te = 1;
    for (int r = 0; r < BG; ++r)
      if (b0 + r < B) Te = max(Te, min(Td, __ldg(dd.step_end + b0 + r)));
  }

  extern __shared__ __align__(16) float smem_raw[];
  BwdSmem<HAS2> S;
  constexpr int SW = NI1 * AFT;  // staging stride per (warp, channel lane)
  S.carve(smem_raw, Tt, 16 * 8 * SW);
  const int TtP = S.TtP, Tt4 = S.Tt4;
  const int DFW = TtP + 2 * HALO;   // row pitch of dfS
  uint64_t* barW = S.bars;
  uint64_t* barQ = S.bars + 2;
  uint64_t* barG = S.bars + 4;
  uint64_t* barC = S.bars + 6;

  const int ab = rank >> 2, cq = rank & 3;
  const int arow = b0 + ab;
  const bool arow_ok = arow < B;
  const int alen = arow_ok ? (int)d.lengths[arow] : 0;
  const int pl = d.att_kernel > 0 ? (d.att_kernel - 1) / 2 : 0;
  const bool loc = d.att_kernel > 0;

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
    S.WqU[uu * WQS + c] = v;
  }
  if (HAS2)
    for (int i = tid; i < 2 * 4 * H; i += NT) {
      int rr = i / (4 * H), c = i % (4 * H);
      S.WragS[i] = __ldg(d.Wrec + (long long)(512 + 2 * rank + rr) * (4 * H) + c);
    }
  for (int i = tid; i < MAXF * QC; i += NT) {
    int f = i / QC, c = i % QC;
    S.Wfs[i] = (f < d.att_filters && c < A1Q && loc) ? __ldg(d.loc_layer_w + (long long)f * d.A1 + cq * A1Q + c) : 0.f;
  }
  for (int i = tid; i < MAXK * MAXF; i += NT) {
    int k = i / MAXF, f = i % MAXF;
    S.wconv[i] = (k < d.att_kernel && f < d.att_filters) ? __ldg(d.loc_conv_w + k * d.att_filters + f) : 0.f;
  }
  if (tid < MAXF) S.bconv[tid] = (tid < d.att_filters && loc) ? __ldg(d.loc_conv_b + tid) : 0.f;
  if (tid < QC) {
    float v = 0.f;
    if (tid < A1Q) v = __ldg(d.v1 + cq * A1Q + tid);
    else if (HAS2) v = __ldg(d.v2 + cq * 8 + (tid - A1Q));
    S.vs[tid] = v;
    S.dqS[tid] = 0.f;
  }
  for (int i = tid; i < TtP + 2 * HALO; i += NT) {
    // cumulative_weights: the location input of step t is the sum of all earlier alignments; start from the final state
    const int j = i - HALO;
    S.aprev[i] = (d.cumulative && loc && arow_ok && j >= 0 && j < Tt) ? __ldg(d.state_final + (long long)arow * Tt + j) : 0.f;
  }
  for (int i = tid; i < (TtP + 2 * HALO) * MAXF; i += NT) S.dfS[i] = 0.f;
  for (int i = tid; i < TtP * MAXF; i += NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * 4 * TtP; i += NT) { S.dstate_part[i] = 0.f; S.dwpart[i] = 0.f; }
  for (int i = tid; i < TtP; i += NT) { S.dalpha_carry[i] = 0.f; S.dstate_own[i] = 0.f; }
  for (int i = tid; i < TtP + 8; i += NT) { S.alphaPrevS[i] = 0.f; S.dmixS[i] = 0.f; }
  for (int i = tid; i < 2 * TtP; i += NT) S.deS[i] = 0.f;
  for (int i = tid; i < RINGB * 3 * TtP; i += NT) S.ringA[i] = 0.f;
  for (int i = tid; i < RINGB * RB; i += NT) S.ringB[i] = 0.f;
  for (int i = tid; i < RINGB * 2 * BG * UH; i += NT) S.mk_ring[i] = 0;
  if (tid < VC + 8) { S.dctx_in[tid] = 0.f; S.dctxS[tid] = 0.f; }
  if (tid < BG * UH) { S.dh_in[tid] = 0.f; S.dout1S[tid] = 0.f; }
  for (int i = tid; i < BG * 256; i += NT) S.dqB[i] = 0.f;
  for (int i = tid; i < MAXK * MAXF + MAXF; i += NT) S.dwconvS[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) cl::mbar_init(&S.bars[i], 1);
    cl::fence_mbar_init();
  }

  // ---------------- BC role: rows [32*rank, 32*rank+32) of Wrec in registers.
  // thread = (kq = tid & 63, output group og = tid >> 6): 4 rows x the source units congruent to kq mod 64 x 4 gates
  const int kq = tid & 63, og = tid >> 6;
  float wr[4][4][4];  // [i][row c][gate]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4)
        wr[i][c][g4] = __ldg(d.Wrec + (long long)(32 * rank + og * 4 + c) * (4 * H) + g4 * H + kq + 64 * i);

  // ---------------- pointwise role (exchange-issuing warps: no global stores)
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < B;
  const int pidx = rank * UH + pu;
  float dc = 0.f, dh = 0.f;

  // ---------------- energy role
  const int pg = warp * 4 + (lane >> 3), cl_ = lane & 7;
  // position passes of this warp that touch the utterance (see attn_rnn_fwd.cu): gradients of positions past the source length are zero
  const int nact = min(NP, max(0, (alen - 4 * warp + 63) / 64));
  float dv_acc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) dv_acc[i] = 0.f;
  float dWf_acc = 0.f;     // tid < NI1*8*AFT owns one (channel, filter) entry of d(location_features_layer)
  float Gacc = 0.f;        // cumulative_weights: running sum of d(state) over the later steps (thread 256 + j owns position j)
  // transition agent (forward_attention.py:111-114): thread tid < 64 owns context column cq*64+tid and thread tid < A1Q score channel
  // cq*A1Q+tid of [ctx1, q1] . W; S.red[40] carries d(u_t) from the step that used it, S.red[41] = d(pre-sigmoid) of step t
  float wa = 0.f, waq = 0.f, dwa_acc = 0.f, dwaq_acc = 0.f, db_acc = 0.f;
  if (AGENT) {
    if (tid < 64) wa = __ldg(d.agent_w + cq * 64 + tid);
    if (tid < A1Q) waq = __ldg(d.agent_w + M1 + cq * A1Q + tid);
    if (tid == 0) { S.red[40] = 0.f; S.red[41] = 0.f; }
  }

  cluster.sync();

  // prefetch of the global inputs of time tau into ring slot tau % RINGB (always commits a group)
  auto prefetch = [&](int tau) {
    if (tau >= 0) {
      const int slot = tau % RINGB;
      if (arow_ok && tid < Tt) {
        const long long oa = ((long long)tau * B + arow) * Tt + tid;
        float* ra = S.ringA + (size_t)slot * 3 * TtP;
        cp_async4(ra + tid, d.soft1 + oa);
        cp_async4(ra + TtP + tid, d.align1 + oa);
        if (HAS2) cp_async4(ra + 2 * TtP + tid, d.align2 + oa);
      }
      float* rb = S.ringB + (size_t)slot * RB;
      if (arow_ok && tid >= 64 && tid < 64 + QC) {
        const int qi = tid - 64;
        if (qi < A1Q || HAS2) {
          const int qcol = (qi < A1Q) ? (cq * A1Q + qi) : (d.A1 + cq * 8 + (qi - A1Q));
          cp_async4(rb + qi, d.q_save + ((long long)tau * B + arow) * QT + qcol);
        }
      }
      if (arow_ok && tid >= 128 && tid < 128 + VCW) {
        const int e = tid - 128;
        const int k = (e < 64) ? (cq * 64 + e) : (M1 + cq * 8 + (e - 64));
        cp_async4(rb + QC + e, dd.dx2 + ((long long)tau * B + arow) * X2W + H + k);
      }
      if (prow_ok) {
        const long long o1 = ((long long)tau * B + prow) * H + pidx;
        const long long o4 = ((long long)tau * B + prow) * (4 * H) + pidx;
        float* rp = rb + QC + VC + 8 + tid;
        cp_async4(rp + 0 * 64, d.gates + o4);
        cp_async4(rp + 1 * 64, d.gates + o4 + H);
        cp_async4(rp + 2 * 64, d.gates + o4 + 2 * H);
        cp_async4(rp + 3 * 64, d.gates + o4 + 3 * H);
        cp_async4(rp + 4 * 64, d.c_prev + o1);
        cp_async4(rp + 5 * 64, dd.dx2 + ((long long)tau * B + prow) * X2W + pidx);
        if ((pu & 3) == 0) {
          uint8_t* mr = S.mk_ring + slot * 2 * BG * UH;
          if (d.mask_c) cp_async4(mr + (0 * BG + pb) * UH + pu, d.mask_c + o1);
          if (d.mask_h) cp_async4(mr + (1 * BG + pb) * UH + pu, d.mask_h + o1);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int i = 0; i <= PFDB; ++i) prefetch(Te - 1 - i);
  // rows of the skipped steps: zero gradients (round-robin over the CTAs of the cluster; plain stores, long before the first exchange)
  for (int r = rank; r < (Td - Te) * BG; r += CS) {
    const int tz = Te + r / BG, bz = b0 + r % BG;
    if (bz < B) {
      float* gz = dd.dgates + ((long long)tz * B + bz) * (4 * H);
      for (int i = tid; i < 4 * H; i += NT) gz[i] = 0.f;
      if (dd.dq) {
        float* qz = dd.dq + ((long long)tz * B + bz) * QT;
        for (int i = tid; i < QT; i += NT) qz[i] = 0.f;
      }
    }
  }

  PT_DECL
#pragma unroll 1
  for (int t = Te - 1; t >= 0; --t) {
    const int u = Te - 1 - t, cur = u & 1, nxt = cur ^ 1;
    const uint32_t par = (uint32_t)(u >> 1) & 1u;
    PT(15)
    prefetch(t - 1 - PFDB);
    cp_async_wait<PFDB>();                       // time t and t-1 are resident
    __syncthreads();                             // ring contents (written by other threads' cp.async) visible
    const float* rA = S.ringA + (size_t)(t % RINGB) * 3 * TtP;           // soft1[t], align1[t], align2[t]
    const float* rAp = S.ringA + (size_t)((t + RINGB - 1) % RINGB) * 3 * TtP;   // time t-1
    const float* rB = S.ringB + (size_t)(t % RINGB) * RB;
    const float* aS = rA;
    const float* alphaS = rA + TtP;
    const float* a2S = rA + 2 * TtP;
    const float* qs = rB;
    // u used by the recursion of step t (u_{t-1}); u_t = the factor produced by step t, used by step t+1
    const float u_tr = (AGENT && arow_ok) ? __ldg(d.u_save + (long long)t * B + arow) : 0.5f;
    float ctx_own = 0.f;
    if (AGENT) {
      if (tid == 0) {
        const float u_t = (t + 1 < Te && arow_ok) ? __ldg(d.u_save + (long long)(t + 1) * B + arow) : 0.5f;
        S.red[41] = S.red[40] * u_t * (1.f - u_t);         // d(pre-sigmoid) of the agent output of step t
      }
      if (tid < 64 && arow_ok) ctx_own = __ldg(d.x2 + ((long long)t * B + arow) * X2W + H + cq * 64 + tid);
    }
    for (int j = tid; j < TtP; j += NT) {
      const bool in = j < Tt && t > 0;
      // state of step t: a_{t-1}, or with cumulative weights state_{t+1} - a_t (walked back from the saved final state)
      S.aprev[HALO + j] = in ? (d.cumulative ? S.aprev[HALO + j] - rA[j] : rAp[j]) : 0.f;
      S.alphaPrevS[j] = in ? rAp[TtP + j] : ((j == 0 && d.mode == 2) ? 1.f : 0.f);
    }
    __syncthreads();
    // location features of this step (input: a_{t-1}); independent of the recurrent carries, so computed before waiting for them
    if (loc) location_features<AFT>(S.fS, S.aprev, S.wconv, S.bconv, Tt, d.att_kernel, pl, tid, NT);
    if (u > 0) cl::mbar_wait(&barC[cur], (uint32_t)((u - 1) >> 1) & 1u);   // recurrent carries of step t+1
    if (tid == 0) {
      const uint32_t rxw = 3u * NATT * (uint32_t)Tt4 * 4u + ((u > 0 && loc) ? 3u * (uint32_t)Tt4 * 4u : 0u);
      cl::mbar_arrive_expect_tx(&barW[cur], rxw);
      cl::mbar_arrive_expect_tx(&barQ[cur], RX_Q);
      if (t > 0) cl::mbar_arrive_expect_tx(&barG[cur], RX_G);
      if (t > 0) cl::mbar_arrive_expect_tx(&barC[nxt], RX_C);
    }
    PT(0)

    // ======================= BA1
    if (tid < VCW) {
      float v = S.dctx_in[tid] + rB[QC + tid];
      if (AGENT && tid < 64) {
        const float dpre = S.red[41];
        v = fmaf(dpre, wa, v);                   // u_t = sigmoid([ctx1_t, q1_t] . W + b): d(ctx1_t) += d(pre) W_ctx
        dwa_acc = fmaf(dpre, ctx_own, dwa_acc);
        if (tid == 0 && cq == 0) db_acc += dpre;
      }
      S.dctxS[tid] = v;                          // total d(ctx); saved for the dense dvalues GEMM by the saver warps
    }
    __syncthreads();
    {
      // partial d(weights)[j] = d(ctx) . values[j, this CTA's columns]: thread = (position pg + 64m, lane cl_ of 8); lane cl_ takes the
      // columns congruent to cl_ mod 8, so the 4 positions x 8 lanes of a warp hit 32 distinct banks (row stride 72)
      float dcx[8], dcx2 = HAS2 ? S.dctxS[64 + cl_] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) dcx[i] = S.dctxS[cl_ + 8 * i];
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        if (m >= nact) continue;                 // d(weights) of positions past the source length is never used
        const int j = pg + 64 * m;
        const float* vr = S.valS + (j < Tt ? j : 0) * KS + cl_;
        float a1 = 0.f, a1b = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          a1 = fmaf(dcx[i], vr[8 * i], a1);
          a1b = fmaf(dcx[i + 1], vr[8 * i + 8], a1b);
        }
        a1 += a1b;
        float a2 = HAS2 ? dcx2 * vr[64] : 0.f;
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 4);
        if (HAS2) {
          a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
          a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
          a2 += __shfl_xor_sync(0xffffffffu, a2, 4);
        }
        if (cl_ == 0 && j < Tt) {
          S.dwpart[(0 * 4 + cq) * TtP + j] = a1;
          if (HAS2) S.dwpart[(1 * 4 + cq) * TtP + j] = a2;
        }
      }
    }
    cl::fence_proxy_async();
    __syncthreads();
    if (tid < 3 * NATT) {
      const int att = tid / 3, q3 = tid % 3;
      const int r4 = q3 + (q3 >= cq ? 1 : 0);
      const uint32_t src = cl::smem_u32(&S.dwpart[(att * 4 + cq) * TtP]);
      cl::bulk_copy_to_cta(cl::mapa(src, ab * 4 + r4), src, (uint32_t)Tt4 * 4u, cl::mapa(cl::smem_u32(&barW[cur]), ab * 4 + r4));
    }
    PT(1)
    cl::mbar_wait(&barW[cur], par);
    PT(2)

    // ======================= BA2a: recursion / softmax backward, block-parallel (one position per thread)
    if (tid >= 256) {
      // attention 1: threads 256..511
      const int j = tid - 256;
      const bool in = j < TtP;
      const float a = in ? aS[j] : 0.f;
      float dw = 0.f, dst = 0.f;
      if (in) {
        dw = S.dwpart[(0 * 4 + 0) * TtP + j] + S.dwpart[(0 * 4 + 1) * TtP + j] + S.dwpart[(0 * 4 + 2) * TtP + j] + S.dwpart[(0 * 4 + 3) * TtP + j];
        if (u > 0 && loc)
          dst = S.dstate_part[(cur * 4 + 0) * TtP + j] + S.dstate_part[(cur * 4 + 1) * TtP + j] + S.dstate_part[(cur * 4 + 2) * TtP + j] +
                S.dstate_part[(cur * 4 + 3) * TtP + j];
      }
      if (d.cumulative) { Gacc += dst; dst = Gacc; }     // a_t feeds the state of EVERY later step
      float da;
      if (d.mode == 2) {
        const float dal = in ? dw + S.dalpha_carry[j] : 0.f;
        const float apm1 = (in && j > 0) ? S.alphaPrevS[j - 1] : 0.f;
        const float mix = in ? ((1.f - u_tr) * S.alphaPrevS[j] + u_tr * apm1 + 1e-7f) : 0.f;
        float s1 = mix * a, s2 = in ? dal * alphaS[j] : 0.f;
        group_sum2(s1, s2, S.red, warp & 7, lane, 2);
        const float dau = (in && j < alen) ? (dal - s2) / s1 : 0.f;
        da = dau * mix + dst;
        if (in) S.dmixS[j] = dau * a;
        if (AGENT) {
          // d(u_{t-1}) = sum_j d(mix_j) (alpha_{t-1}[j-1] - alpha_{t-1}[j])      (forward_attention.py:109)
          float s3 = in ? dau * a * (apm1 - S.alphaPrevS[j]) : 0.f, dummy3 = 0.f;
          group_sum2(s3, dummy3, S.red, warp & 7, lane, 2);
          if (tid == 256) S.red[40] = (t > 0) ? s3 : 0.f;
        }
      } else {
        da = dw + dst;
      }
      float dot2 = da * a, dummy = 0.f;
      group_sum2(dot2, dummy, S.red, warp & 7, lane, 2);
      if (in) {
        S.deS[0 * TtP + j] = a * (da - dot2);
        if (d.mode == 2) {
          const float nx = (j + 1 < TtP) ? S.dmixS[j + 1] : 0.f;
          S.dalpha_carry[j] = (1.f - u_tr) * S.dmixS[j] + u_tr * nx;  // adjoint of the shift (forward_attention.py:108-109)
        }
      }
    } else if (HAS2) {
      // attention 2: threads 0..255
      const int j = tid;
      const bool in = j < TtP;
      const float a = in ? a2S[j] : 0.f;
      float dw = 0.f;
      if (in) dw = S.dwpart[(1 * 4 + 0) * TtP + j] + S.dwpart[(1 * 4 + 1) * TtP + j] + S.dwpart[(1 * 4 + 2) * TtP + j] + S.dwpart[(1 * 4 + 3) * TtP + j];
      float dot = dw * a, dummy = 0.f;
      group_sum2(dot, dummy, S.red + 16, warp & 7, lane, 3);
      if (in) S.deS[1 * TtP + j] = a * (dw - dot);
    }
    __syncthreads();
    PT(3)

    // ======================= BA2b: energy backward over this CTA's channels
    float* stageW = S.dgx;  // aliased: the d(gates) buffer is idle during BA2
    {
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
        const int c = cl_ + 8 * i;
        float wf[AFT], P[AFT];
#pragma unroll
        for (int f = 0; f < AFT; ++f) { wf[f] = S.Wfs[f * QC + c]; P[f] = 0.f; }
        const float qc = qs[c], vc = S.vs[c];
        float dq = 0.f;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          if (m >= nact) continue;
          float s = S.keyS[jm[m] * KS + c] + qc;
#pragma unroll
          for (int f = 0; f < AFT; ++f) s = fmaf(fv[m][f], wf[f], s);
          const float th = ftanh(s);
          const float dsv = de1[m] * vc * (1.f - th * th);
          dv_acc[i] = fmaf(de1[m], th, dv_acc[i]);
          if (jok[m]) cl::sts_noalias(&S.dkeyS[jm[m] * KS + c], S.dkeyS[jm[m] * KS + c] + dsv);
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
          for (int f = 0; f < AFT; ++f) cl::sts_noalias(&stageW[(warp * 8 + cl_) * SW + i * AFT + f], P[f]);
        }
      }
      if (HAS2) {
        const int c = A1Q + cl_;
        const float qc = qs[c], vc = S.vs[c];
        float dq = 0.f;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          if (m >= nact) continue;
          const float th = ftanh(S.keyS[jm[m] * KS + c] + qc);
          const float dsv = de2[m] * vc * (1.f - th * th);
          dv_acc[NI1] = fmaf(de2[m], th, dv_acc[NI1]);
          if (jok[m]) cl::sts_noalias(&S.dkeyS[jm[m] * KS + c], S.dkeyS[jm[m] * KS + c] + dsv);
          dq += dsv;
        }
        dq_acc[NI1] = dq;
      }
      // dq: reduce over position lanes, stage per warp
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        dq_acc[i] += __shfl_xor_sync(0xffffffffu, dq_acc[i], 8);
        dq_acc[i] += __shfl_xor_sync(0xffffffffu, dq_acc[i], 16);
        if (lane < 8) cl::sts_noalias(&S.stageQ[warp * QC + cl_ + 8 * i], dq_acc[i]);
      }
      // d(location features): reduce over the 8 channel lanes
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        if (m >= nact) continue;                 // dfS of those positions stays at its initial zero
#pragma unroll
        for (int f = 0; f < AFT; ++f) {
          float v = dfp[m][f];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          if (cl_ == 0 && jok[m]) cl::sts_noalias(&S.dfS[f * DFW + HALO + jm[m]], v);
        }
      }
    }
    __syncthreads();
    PT(4)
    {
      if (tid < QC) {
        float q = 0.f;
        if (tid < A1Q || HAS2) {
#pragma unroll
          for (int w_ = 0; w_ < 16; ++w_) q += S.stageQ[w_ * QC + tid];
        }
        if (AGENT && tid < A1Q) {
          const float dpre = S.red[41];                   // d(q1_t) += d(pre) W_q ; d(W_q) += d(pre) q1_t
          dwaq_acc = fmaf(dpre, qs[tid], dwaq_acc);
          q = fmaf(dpre, waq, q);
        }
        S.dqS[tid] = q;
        const int qcol = (tid < A1Q) ? (cq * A1Q + tid) : (d.A1 + cq * 8 + (tid - A1Q));
        if (tid < A1Q || HAS2) S.dqB[ab * 256 + qcol] = q;    // own copy
      }
      if (loc) {
        if (tid < NI1 * 8 * AFT) {
          // thread owns (channel lane, i, f) of d(location_features_layer)
          const int c8 = tid / SW, rem = tid % SW;
          float acc = 0.f;
#pragma unroll
          for (int w_ = 0; w_ < 16; ++w_) acc += stageW[(w_ * 8 + c8) * SW + rem];
          dWf_acc += acc;
        }
        // partial d(state a_{t-1}) = conv-transpose of d(location features)
        if (t > 0) {
          for (int j = NT - 1 - tid; j < Tt4; j += NT) {   // last warps: the first ones carry the dq / dWf reductions
            float acc = 0.f;
            if (j < Tt) {
              const float* dfr = S.dfS + HALO + j + pl;   // columns outside [0,Tt) are zero
              if (d.att_kernel == 10) {
#pragma unroll
                for (int k = 0; k < 10; ++k)
#pragma unroll
                  for (int f = 0; f < AFT; ++f) acc = fmaf(dfr[f * DFW - k], S.wconv[k * MAXF + f], acc);
              } else {
                for (int k = 0; k < d.att_kernel; ++k)
#pragma unroll
                  for (int f = 0; f < AFT; ++f) acc = fmaf(dfr[f * DFW - k], S.wconv[k * MAXF + f], acc);
              }
            }
            S.dstate_own[j] = acc;
            S.dstate_part[(nxt * 4 + cq) * TtP + j] = acc;   // own contribution for step t-1
          }
        }
        // d(location conv kernel)[k][f] += sum_j a_prev[j+k-pl] * df[j][f]: 8 lanes per (k,f), accumulators in shared memory
        for (int e0 = 0; e0 < d.att_kernel * AFT; e0 += NT / 8) {   // warp-uniform trip count: the shuffles need all lanes
          const int e = e0 + (tid >> 3);
          const bool ok = e < d.att_kernel * AFT;
          const int k = ok ? e / AFT : 0, f = ok ? e % AFT : 0;
          float acc = 0.f;
          for (int j = lane & 7; j < alen; j += 8) acc = fmaf(S.aprev[HALO + j + k - pl], S.dfS[f * DFW + HALO + j], acc);
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          acc += __shfl_xor_sync(0xffffffffu, acc, 4);
          if (ok && (lane & 7) == 0) S.dwconvS[k * MAXF + f] += acc;
        }
        if (warp >= 6 && warp < 6 + AFT) {
          const int f = warp - 6;
          float acc = 0.f;
          for (int j = lane; j < Tt; j += 32) acc += S.dfS[f * DFW + HALO + j];
          acc = warp_sum(acc);
          if (lane == 0) S.dwconvS[MAXK * MAXF + f] += acc;
        }
      }
    }
    cl::fence_proxy_async();
    __syncthreads();
    if (tid < 16) {
      // dq slice -> every CTA (16-byte st.async per 4 columns)
      const int nq4 = (A1Q + (HAS2 ? 8 : 0)) / 4;
      if (tid < nq4) {
        const int c4 = tid * 4;
        const int qcol = (c4 < A1Q) ? (cq * A1Q + c4) : (d.A1 + cq * 8 + (c4 - A1Q));
        const uint32_t dsta = cl::smem_u32(&S.dqB[ab * 256 + qcol]), bara = cl::smem_u32(&barQ[cur]);
        const float q0 = S.dqS[c4], q1 = S.dqS[c4 + 1], q2 = S.dqS[c4 + 2], q3 = S.dqS[c4 + 3];
#pragma unroll
        for (int r = 0; r < CS; ++r)
          if (r != rank) st_async_v4(cl::mapa(dsta, r), q0, q1, q2, q3, cl::mapa(bara, r));
      }
    } else if (tid >= 32 && tid < 35 && t > 0 && loc) {
      const int q3 = tid - 32;
      const int r4 = q3 + (q3 >= cq ? 1 : 0);
      const uint32_t src = cl::smem_u32(S.dstate_own);
      cl::bulk_copy_to_cta(cl::mapa(cl::smem_u32(&S.dstate_part[(nxt * 4 + cq) * TtP]), ab * 4 + r4), src, (uint32_t)Tt4 * 4u,
                           cl::mapa(cl::smem_u32(&barW[nxt]), ab * 4 + r4));
    }
    PT(5)
    cl::mbar_wait(&barQ[cur], par);
    PT(6)

    // ======================= BB: d(out1), LSTM cell backward
    {
      // thread = (b = tid>>7, u = (tid>>3)&15, part = tid&7): 32 columns of dq each
      const int bb = tid >> 7, uu = (tid >> 3) & 15, part = tid & 7;
      float acc = 0.f;
      // lane `part` takes the columns congruent to part mod 8: consecutive lanes read consecutive banks
      const float* qrow = S.dqB + bb * 256 + part;
      const float* wrow = S.WqU + uu * WQS + part;
#pragma unroll 8
      for (int c = 0; c < 32; ++c) acc = fmaf(qrow[8 * c], wrow[8 * c], acc);
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
        const float* rp = rB + QC + VC + 8 + tid;
        const float gi = rp[0], gj = rp[64], gf = rp[128], go = rp[192], cp = rp[256], dx2o = rp[320];
        const uint8_t* mr = S.mk_ring + (t % RINGB) * 2 * BG * UH;
        const float mc = d.mask_c ? (float)mr[(0 * BG + pb) * UH + pu] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mr[(1 * BG + pb) * UH + pu] : (1.f - d.zh);
        const float c_new = gf * cp + gi * gj;
        const float tc = ftanh(c_new);
        const float dout1 = dx2o + S.dout1S[pb * UH + pu];
        const float dh_new = dout1 + mh * dh;
        dh = (1.f - mh) * dh;
        const float dcn = mc * dc + dh_new * go * (1.f - tc * tc);
        dgo = dh_new * tc * go * (1.f - go);
        dgi = dcn * gj * gi * (1.f - gi);
        dgj = dcn * gi * (1.f - gj * gj);
        dgf = dcn * cp * gf * (1.f - gf);
        dc = (1.f - mc) * dc + dcn * gf;
      }
      float* slot = &S.dgx[((size_t)pb * H + pidx) * 4];
      *reinterpret_cast<float4*>(slot) = make_float4(dgi, dgj, dgf, dgo);
      *reinterpret_cast<float4*>(&S.save_dg[tid * 4]) = make_float4(dgi, dgj, dgf, dgo);
      if (t > 0) {
        const uint32_t dsta = cl::smem_u32(slot), bara = cl::smem_u32(&barG[cur]);
#pragma unroll
        for (int r = 0; r < CS; ++r)
          if (r != rank) st_async_v4(cl::mapa(dsta, r), dgi, dgj, dgf, dgo, cl::mapa(bara, r));
      }
    }
    __syncthreads();
    PT(7)
    if (tid >= 256) {
      // saver warps: d(gates), dq slice and total d(ctx) of step t -> global memory
      const int e = tid - 256;
      if (e < 64) {
        const int sb = e >> 4, su = e & 15;
        if (b0 + sb < B) {
          const long long o4 = ((long long)t * B + b0 + sb) * (4 * H) + rank * UH + su;
          const float4 v = *reinterpret_cast<const float4*>(&S.save_dg[e * 4]);
          dd.dgates[o4] = v.x; dd.dgates[o4 + H] = v.y; dd.dgates[o4 + 2 * H] = v.z; dd.dgates[o4 + 3 * H] = v.w;
        }
      } else if (e < 64 + QC) {
        const int qi = e - 64;
        if (arow_ok && (qi < A1Q || HAS2)) {
          const int qcol = (qi < A1Q) ? (cq * A1Q + qi) : (d.A1 + cq * 8 + (qi - A1Q));
          dd.dq[((long long)t * B + arow) * QT + qcol] = S.dqS[qi];
        }
      } else if (e < 64 + QC + VCW) {
        const int ci = e - 64 - QC;
        if (arow_ok) {
          const int k = (ci < 64) ? (cq * 64 + ci) : (M1 + cq * 8 + (ci - 64));
          dd.dx2[((long long)t * B + arow) * X2W + H + k] = S.dctxS[ci];
        }
      }
    }
    if (t == 0) break;
    cl::mbar_wait(&barG[cur], par);
    PT(8)

    // ======================= BC: d(ctx, h)(t-1) = d(gates) . Wrec^T
    {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int b = 0; b < BG; ++b) {
          const float4 g4 = *reinterpret_cast<const float4*>(&S.dgx[((size_t)b * H + kq + 64 * i) * 4]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a = acc[c * 4 + b];
            a = fmaf(wr[i][c][0], g4.x, a);
            a = fmaf(wr[i][c][1], g4.y, a);
            a = fmaf(wr[i][c][2], g4.z, a);
            a = fmaf(wr[i][c][3], g4.w, a);
            acc[c * 4 + b] = a;
          }
        }
      }
      const float v = cl::reduce_scatter16(acc, lane);   // lane L: row og*4 + ((L>>2)&3), batch row L&3
      S.bcpart[(((kq >> 4) & 3) * 32 + og * 4 + ((lane >> 2) & 3)) * 4 + (lane & 3)] = v;
    }
    if (HAS2) {
      // ragged rows 512 + 2*rank + {0,1} of Wrec (hidden units 224..255; weights in shared memory), spread over all warps:
      // thread (kq, og) takes row og&1 and source unit kq + 64*(og>>1); per-warp totals land in bcpart[512 + warp*4 + b]
      const int un = kq + 64 * (og >> 1);
      const float* wrow = S.WragS + (og & 1) * 4 * H + un;
      const float w0 = wrow[0], w1 = wrow[H], w2 = wrow[2 * H], w3 = wrow[3 * H];
      float rg[4];
#pragma unroll
      for (int b = 0; b < BG; ++b) {
        const float4 g4 = *reinterpret_cast<const float4*>(&S.dgx[((size_t)b * H + un) * 4]);
        rg[b] = fmaf(w0, g4.x, fmaf(w1, g4.y, fmaf(w2, g4.z, w3 * g4.w)));
      }
      {
        const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
        const float s0 = up2 ? rg[0] : rg[2], s1 = up2 ? rg[1] : rg[3];
        float k0 = up2 ? rg[2] : rg[0], k1 = up2 ? rg[3] : rg[1];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        float v = (up1 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up1 ? k0 : k1, 1);   // lane L: batch row L&3
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 4) S.bcpart[512 + warp * 4 + lane] = v;
      }
    }
    __syncthreads();
    if (tid < 128) {
      // finalise: thread = (batch row b = tid >> 5, row k = tid & 31) -> 4 consecutive rows travel as one st.async.v4
      const int b = tid >> 5, k = tid & 31;
      const float v = S.bcpart[(0 * 32 + k) * 4 + b] + S.bcpart[(1 * 32 + k) * 4 + b] + S.bcpart[(2 * 32 + k) * 4 + b] +
                      S.bcpart[(3 * 32 + k) * 4 + b];
      const int l4 = lane & ~3;
      const float v0 = __shfl_sync(0xffffffffu, v, l4), v1 = __shfl_sync(0xffffffffu, v, l4 + 1);
      const float v2 = __shfl_sync(0xffffffffu, v, l4 + 2), v3 = __shfl_sync(0xffffffffu, v, l4 + 3);
      if ((lane & 3) == 0) {
        const int kg = 32 * rank + k;   // row of Wrec (< 512)
        int dst;
        uint32_t addr;
        if (kg < M1) {
          dst = b * 4 + (kg >> 6);
          addr = cl::smem_u32(&S.dctx_in[kg & 63]);
        } else if (kg < M1 + M2) {
          dst = b * 4 + ((kg - M1) >> 3);
          addr = cl::smem_u32(&S.dctx_in[64 + ((kg - M1) & 7)]);
        } else {
          const int un = kg - (M1 + M2);
          dst = un >> 4;
          addr = cl::smem_u32(&S.dh_in[b * UH + (un & 15)]);
        }
        st_async_v4(cl::mapa(addr, dst), v0, v1, v2, v3, cl::mapa(cl::smem_u32(&barC[nxt]), dst));
      }
    } else if (HAS2 && tid < 136) {
      // ragged rows: (row rr, batch row b) = sum over the 8 warps with ((warp >> 1) & 1) == rr
      const int rr = (tid - 128) >> 2, b = tid & 3;
      float v = 0.f;
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) v += S.bcpart[512 + (w4 * 4 + rr * 2) * 4 + b] + S.bcpart[512 + (w4 * 4 + rr * 2 + 1) * 4 + b];
      const int un = (512 + 2 * rank + rr) - (M1 + M2);
      const int dst = un >> 4;
      cl::st_async_f32(cl::mapa(cl::smem_u32(&S.dh_in[b * UH + (un & 15)]), dst), v, cl::mapa(cl::smem_u32(&barC[nxt]), dst));
    }
    PT(9)
  }
  PT_FLUSH(Td)
  cp_async_wait<0>();

  // ---------------- epilogue: flush accumulators
  __syncthreads();
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
        const int c = cl_ + 8 * i;
        if (i < NI1) atomicAdd(dd.dv1 + cq * A1Q + c, v);
        else atomicAdd(dd.dv2 + cq * 8 + cl_, v);
      }
    }
    if (AGENT) {
      if (tid < 64) atomicAdd(dd.dagent_w + cq * 64 + tid, dwa_acc);
      if (tid < A1Q) atomicAdd(dd.dagent_w + M1 + cq * A1Q + tid, dwaq_acc);
      if (tid == 0 && cq == 0) atomicAdd(dd.dagent_b, db_acc);
    }
    if (loc) {
      if (tid < NI1 * 8 * AFT) {
        const int c8 = tid / SW, rem = tid % SW;
        const int i_ = rem / AFT, f_ = rem % AFT;
        if (f_ < d.att_filters) atomicAdd(dd.dloc_layer_w + (long long)f_ * d.A1 + cq * A1Q + c8 + 8 * i_, dWf_acc);
      }
      for (int e = tid; e < d.att_kernel * AFT; e += NT) {
        const int k = e / AFT, f = e % AFT;
        if (f < d.att_filters) atomicAdd(dd.dloc_conv_w + k * d.att_filters + f, S.dwconvS[k * MAXF + f]);
      }
      if (tid < d.att_filters) atomicAdd(dd.dloc_conv_b + tid, S.dwconvS[MAXK * MAXF + tid]);
    }
  }
  cluster.sync();
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

namespace satk { namespace arnn2 {
bool v2_eligible(const satk_attn_rnn_fwd_desc* d);
int attn_rnn2_bwd_launch(const satk_attn_rnn_bwd_desc* d, float* de, int* prog, cudaStream_t st);
int attn_energy_grad_prepare(const satk_attn_rnn_bwd_desc* d, cudaStream_t st);
int* attn_energy_grad_progress(const satk_attn_rnn_bwd_desc* d);
int attn_energy_grad_launch(const satk_attn_rnn_bwd_desc* d, const float* de, int parts, int dependent, cudaStream_t st);
} }

static int bwd2_check(const satk_attn_rnn_bwd_desc* d, bool features_only = false) {
  bool has2;
  int rc = attn_rnn_check(&d->f, has2);
  if (rc) return rc;
  if (!d->de_ws || (!d->sync_ws && !features_only) || !arnn2::v2_eligible(&d->f)) {
    satk::set_error("attn_rnn_bwd: configuration not covered by the second-generation kernels (or de_ws / sync_ws missing)");
    return SATK_ERR_UNSUPPORTED;
  }
  if (features_only) return SATK_OK;     // reads the saved alignments and the location convolution only
  SATK_CHECK_ARG(d->f.gates && d->f.c_prev && d->f.q_save && d->f.soft1 && d->dkeys1 && d->dkeys2 && d->dv1 && d->dv2 &&
                 d->dloc_conv_w && d->dloc_conv_b && d->dloc_layer_w,
                 "attn_rnn_bwd: forward must have saved gates/c_prev/q_save/soft1 and every gradient buffer must be set");
  return SATK_OK;
}

extern "C" int satk_attn_rnn_bwd_recurrence(const satk_attn_rnn_bwd_desc* d, void* stream) {
  int rc = bwd2_check(d);
  if (rc) return rc;
  return arnn2::attn_rnn2_bwd_launch(d, d->de_ws, nullptr, (cudaStream_t)stream);
}

extern "C" int satk_attn_energy_grad_parts(const satk_attn_rnn_bwd_desc* d, int parts, void* stream) {
  SATK_CHECK_ARG(parts >= 1 && parts <= 3, "attn_energy_grad: parts=%d", parts);
  int rc = bwd2_check(d, /*features_only=*/parts == SATK_EG_FEATURES);
  if (rc) return rc;
  return arnn2::attn_energy_grad_launch(d, d->de_ws, parts, 0, (cudaStream_t)stream);
}

// Recurrence + energy gradients as ONE overlapped pair on one stream: the gradient grid is a programmatic dependent of the
// recurrence grid (it starts once every recurrence CTA is resident, runs on the SMs the clusters leave idle and follows the
// progress flags), so only its last chunks remain when the recurrence ends.
extern "C" int satk_attn_energy_grad_prepare(const satk_attn_rnn_bwd_desc* d, void* stream) {
  int rc = bwd2_check(d);
  if (rc) return rc;
  return arnn2::attn_energy_grad_prepare(d, (cudaStream_t)stream);
}

extern "C" int satk_attn_rnn_bwd_overlapped(const satk_attn_rnn_bwd_desc* d, int features, void* stream) {
  int rc = bwd2_check(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!(features & SATK_EG_PREPARED)) {
    rc = arnn2::attn_energy_grad_prepare(d, st);
    if (rc) return rc;
  }
  if (features & SATK_EG_FEATURES) {
    rc = arnn2::attn_energy_grad_launch(d, d->de_ws, SATK_EG_FEATURES, 0, st);
    if (rc) return rc;
  }
  rc = arnn2::attn_rnn2_bwd_launch(d, d->de_ws, arnn2::attn_energy_grad_progress(d), st);
  if (rc) return rc;
  return arnn2::attn_energy_grad_launch(d, d->de_ws, SATK_EG_GRADIENTS, 1, st);
}

extern "C" int satk_attn_energy_grad(const satk_attn_rnn_bwd_desc* d, void* stream) {
  return satk_attn_energy_grad_parts(d, SATK_EG_FEATURES | SATK_EG_GRADIENTS, stream);
}

extern "C" int satk_attn_rnn_bwd(const satk_attn_rnn_bwd_desc* d, void* stream) {
  bool has2;
  int rc = attn_rnn_check(&d->f, has2);
  if (rc) return rc;
  if (d->de_ws && d->sync_ws && arnn2::v2_eligible(&d->f)) {
    rc = satk_attn_rnn_bwd_recurrence(d, stream);
    if (rc) return rc;
    return satk_attn_energy_grad(d, stream);
  }
  SATK_CHECK_ARG(!d->f.cumulative || d->f.att_kernel == 0 || d->f.state_final, "attn_rnn_bwd: cumulative_weights needs the final state saved by forward");
  SATK_CHECK_ARG(d->f.gates && d->f.c_prev && d->f.q_save && d->f.soft1,
                 "attn_rnn_bwd: forward must have saved gates/c_prev/q_save/soft1");
  cudaStream_t st = (cudaStream_t)stream;
  const int np = (d->f.Tt + 63) / 64;
  const bool af5 = d->f.att_filters == 5 || d->f.att_kernel == 0;
  const bool agent = d->f.agent_w != nullptr;
  SATK_CHECK_ARG(!agent || (d->f.mode == 2 && d->f.u_save && d->dagent_w && d->dagent_b),
                 "attn_rnn_bwd: the transition agent needs forward attention, the saved factors and its gradient buffers");
  size_t smem = has2 ? bwd_smem_bytes<true>(d->f.Tt, af5 ? 5 : 8) : bwd_smem_bytes<false>(d->f.Tt, af5 ? 5 : 8);
  SATK_CHECK_ARG(smem <= 227 * 1024, "attn_rnn_bwd: Tt=%d needs %zu B of shared memory (> 227 KB)", d->f.Tt, smem);
#define SATK_ARNN_DISPATCH(H2, AF, NPV) \
  do { if (agent) return launch16b(attn_rnn_bwd_kernel<H2, AF, NPV, true>, *d, smem, st); \
       return launch16b(attn_rnn_bwd_kernel<H2, AF, NPV, false>, *d, smem, st); } while (0)
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
