// Attention-RNN forward, second generation (sm_100a): LSTM-1 + both attention mechanisms of the dual-source decoder for all Td
// steps in ONE launch, NB utterances per 16-CTA cluster (geometry: attn_rnn2.cuh).  Per step:
//   P1   gate GEMM [NB x 544] x [544 x 64 gate columns of this CTA] from register-resident weights            (all 16 warps)
//   P2a  LSTM cell + zoneout for the CTA's 16 units; carried h -> every CTA's next-step input row              [X]
//   P2b  partial queries  sum_{k in own 16 units} out1[u][k] Wq[k][:]  -> the G CTAs of utterance u            [Q]
//   P3   (group CTA) q = sum of the 16 partials; energies of its channel slice over all positions
//        (tanh through ex2/rcp on pre-scaled keys / weights); partial energies -> the group                    [E]
//   P4   masked softmax + forward-attention recursion (redundantly per group CTA; one position per thread)
//   P5   context slice (own value columns) -> every CTA's next-step input row                                  [X]
// [X]/[Q]/[E] = mbarrier transaction barriers completed by st.async (cluster_sync.cuh); there is no cluster-wide barrier in
// the loop and three point-to-point exchanges per step.  Warps 13..15 own all global-memory traffic of the loop (saved
// activations, alignment history, cp.async prefetch rings): warps that issue DSMEM stores never store to global memory.
// Reference semantics: forward_attention.py:88-136 (+ :13-26), TF BahdanauAttention / AttentionWrapper,
// ZoneoutLSTMCell (A.5, A.6), module.py:1011-1042.
#include <stdlib.h>
#include "attn_rnn2.cuh"

namespace satk {
namespace arnn2 {

using cl::cp_async_commit;
using cl::cp_async_wait;
using cl::st_async_v4;
constexpr int RING = 4;   // prefetch ring slots of the x-projection rows / zoneout masks
constexpr int PFD = 2;    // prefetch distance in steps (a step is ~5 us, far above the HBM latency)

constexpr int NT = 512;
constexpr int KREG = 13;            // k-slices (of 17) of the recurrent weights held in registers; the rest sit in shared memory

// per-phase cycle accounting of thread 0 / CTA 0 into shared memory (the generic PT macros keep 16 64-bit accumulators in
// registers, which this kernel cannot afford)
#ifdef SATK_PHASE_TIMING
#define PT2_DECL unsigned pt_last = (unsigned)clock();
#define PT2(i) if (tid == 0) { const unsigned pt_n = (unsigned)clock(); S.pt[i] += pt_n - pt_last; pt_last = pt_n; }
#define PT2_FLUSH(n) if (blockIdx.x == 0 && tid == 0) { for (int i_ = 0; i_ < 16; ++i_) satk::g_phase[i_] = (long long)(S.pt[i_] / (unsigned)(n)); }
#else
#define PT2_DECL
#define PT2(i)
#define PT2_FLUSH(n)
#endif

template <int NB>
struct FwdSmem2 {
  using GE = Geo<NB>;
  int TtP;
  float *xrec, *gsm, *hnS, *WqS, *qin, *qsS, *keyS, *valS, *fS, *wconv, *bconv, *epart, *estage, *aprev, *alphaS, *w1S, *w2S, *softS,
      *cpart, *save1, *red, *consts, *xg_ring, *hstS;
  float4 *packA, *packB, *Wsm;
  unsigned* pt;
  uint8_t* mk_ring;
  uint64_t* bars;   // [0..1] X, [2..3] Q, [4..5] E
  __host__ __device__ static size_t al4(size_t n) { return (n + 3) & ~(size_t)3; }
  __host__ __device__ size_t carve(float* base, int np) {
    TtP = np * PSL;                                     // 156 | 208 (multiple of 4)
    float* p = base;
    xrec = p; p += al4(2 * NB * KREC);                  // [buf][u][k]
    gsm = p; p += NB * 64;
    hnS = p; p += al4(NB * UH);
    hstS = p; p += al4(NB * UH);
    Wsm = reinterpret_cast<float4*>(p); p += 4 * (17 - KREG) * NT;   // [k-slice][warp][lane] x 4 gate columns
    pt = reinterpret_cast<unsigned*>(p); p += 16;
    WqS = p; p += UH * QT;                              // [unit][256]
    qin = p; p += 16 * GE::QC;                          // [source CTA][channel of my slice]
    qsS = p; p += GE::QC;
    packA = reinterpret_cast<float4*>(p); p += 4 * GE::NSLOT;   // {q', v, wf0', wf1'}
    packB = reinterpret_cast<float4*>(p); p += 4 * GE::NSLOT;   // {wf2', wf3', wf4', key column (int)}
    keyS = p; p += (size_t)TtP * GE::KSTR;
    valS = p; p += (size_t)TtP * GE::VC;
    fS = p; p += (size_t)TtP * MAXF;
    wconv = p; p += MAXK * MAXF;
    bconv = p; p += MAXF;
    epart = p; p += 2 * GE::G * (size_t)TtP;            // [att][source member][j]
    estage = p; p += 2 * (size_t)TtP;
    aprev = p; p += al4(TtP + 2 * HALO);
    alphaS = p; p += TtP;
    w1S = p; p += TtP;
    w2S = p; p += TtP;
    softS = p; p += TtP;
    cpart = p; p += CW * GE::VC;
    save1 = p; p += 7 * NB * UH;
    red = p; p += 64;
    consts = p; p += 8;
    xg_ring = p; p += RING * NB * 64;
    mk_ring = reinterpret_cast<uint8_t*>(p); p += RING * 2 * NB * UH / 4;
    bars = reinterpret_cast<uint64_t*>(p); p += 2 * 6;
    return (size_t)(p - base) * sizeof(float);
  }
};

// Partial energies of NACT position passes of one warp: lane = (position lane ep = lane >> 3, channel lane ecl = lane & 7); the
// sums over this CTA's channel slice land in estage[att][j].  e = sum_c v_c tanh(s_c) = Vs - 2 sum_c v_c / (1 + 2^(s'_c)).
template <int NB, int NACT>
__device__ __forceinline__ void energy_passes(const FwdSmem2<NB>& S, int slot0, int ecl, int Tt, float Vs1, float Vs2) {
  using GE = Geo<NB>;
  constexpr int NIA = GE::NIA, NIB = GE::NIB, KSTR = GE::KSTR;
  float fv[NACT][AFT], acc1[NACT], acc2[NACT];
  const float* krow[NACT];
#pragma unroll
  for (int m = 0; m < NACT; ++m) {
    const int j = min(slot0 + PSL * m, Tt - 1);
    krow[m] = S.keyS + j * KSTR;
    const float4 f4 = *reinterpret_cast<const float4*>(&S.fS[j * MAXF]);
    fv[m][0] = f4.x; fv[m][1] = f4.y; fv[m][2] = f4.z; fv[m][3] = f4.w;
    fv[m][4] = S.fS[j * MAXF + 4];
    acc1[m] = 0.f; acc2[m] = 0.f;
  }
#pragma unroll 2
  for (int i = 0; i < NIA; ++i) {
    const float4 pa = S.packA[ecl + 8 * i], pb = S.packB[ecl + 8 * i];
    const int col = __float_as_int(pb.w);
#pragma unroll
    for (int m = 0; m < NACT; ++m) {
      float s = krow[m][col] + pa.x;
      s = fmaf(fv[m][0], pa.z, s); s = fmaf(fv[m][1], pa.w, s);
      s = fmaf(fv[m][2], pb.x, s); s = fmaf(fv[m][3], pb.y, s); s = fmaf(fv[m][4], pb.z, s);
      acc1[m] = fmaf(pa.y, rcpf(1.f + ex2f(s)), acc1[m]);
    }
  }
#pragma unroll
  for (int i = 0; i < NIB; ++i) {
    const float4 pa = S.packA[8 * NIA + ecl + 8 * i];
    const int col = __float_as_int(S.packB[8 * NIA + ecl + 8 * i].w);
#pragma unroll
    for (int m = 0; m < NACT; ++m) acc2[m] = fmaf(pa.y, rcpf(1.f + ex2f(krow[m][col] + pa.x)), acc2[m]);
  }
#pragma unroll
  for (int m = 0; m < NACT; ++m) {
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
      acc1[m] += __shfl_xor_sync(0xffffffffu, acc1[m], o);
      acc2[m] += __shfl_xor_sync(0xffffffffu, acc2[m], o);
    }
    if (ecl == 0) {
      S.estage[0 * S.TtP + slot0 + PSL * m] = fmaf(-2.f, acc1[m], Vs1);
      S.estage[1 * S.TtP + slot0 + PSL * m] = fmaf(-2.f, acc2[m], Vs2);
    }
  }
}

template <int NB, int NP>
__global__ void __launch_bounds__(NT, 1) attn_rnn2_fwd_kernel(const satk_attn_rnn_fwd_desc d) {
  using GE = Geo<NB>;
  constexpr int G = GE::G, QA = GE::QA, QC = GE::QC, QAq = GE::QAq, QBq = GE::QBq, VC = GE::VC, VAq = GE::VAq, VCq = GE::VCq;
  constexpr int NIA = GE::NIA, NIB = GE::NIB, NSLOT = GE::NSLOT, KSTR = GE::KSTR;
  constexpr int KPT = KREC / 32;                       // 17 k's per lane in the gate GEMM
  constexpr uint32_t RX_X = (uint32_t)NB * (CS * UH + M1 + M2) * 4u;   // h slices of the 16 CTAs + all context columns
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CS) * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, Td = d.Td;

  extern __shared__ __align__(16) float smem_raw[];
  FwdSmem2<NB> S;
  S.carve(smem_raw, NP);
  const int TtP = S.TtP;
  uint64_t* barX = S.bars;
  uint64_t* barQ = S.bars + 2;
  uint64_t* barE = S.bars + 4;

  // attention role: group member `ag` of utterance `au`
  const bool grp = rank < NB * G;
  const int au = grp ? rank / G : 0, ag = grp ? rank % G : 0;
  const int arow = b0 + au;
  const bool arow_ok = grp && arow < B;
  const int alen = arow_ok ? min((int)d.lengths[arow], Tt) : 0;
  const int pl = (d.att_kernel - 1) / 2;
  const Slice<NB> sl(ag);

  // ---------------- one-time loads
  for (int i = tid; i < UH * QT; i += NT) {
    const int k = i / QT, c = i % QT;
    S.WqS[i] = (c < A1) ? __ldg(d.Wq1 + (long long)(rank * UH + k) * A1 + c) : __ldg(d.Wq2 + (long long)(rank * UH + k) * A2 + (c - A1));
  }
  for (int i = tid; i < TtP * KSTR; i += NT) {
    const int j = i / KSTR, c = i % KSTR;
    float kv = 0.f;
    if (arow_ok && j < Tt) {
      const long long rowi = (long long)j * B + arow;  // time-major memory
      if (c < QA) {
        if (c < 4 * sl.qan) kv = (__ldg(d.keys1 + rowi * A1 + 4 * sl.qa0 + c) + (d.b1 ? __ldg(d.b1 + 4 * sl.qa0 + c) : 0.f)) * K2LOG2E;
      } else if (c < QC) {
        if (c - QA < 4 * sl.qbn) kv = __ldg(d.keys2 + rowi * A2 + 4 * sl.qb0 + (c - QA)) * K2LOG2E;
      }
    }
    S.keyS[i] = kv;
  }
  for (int i = tid; i < TtP * VC; i += NT) {
    const int j = i / VC, c = i % VC;
    float vv = 0.f;
    if (arow_ok && j < Tt) {
      const long long rowi = (long long)j * B + arow;
      if (c < GE::VA) {
        if (c < 4 * sl.van) vv = __ldg(d.values1 + rowi * M1 + 4 * sl.va0 + c);
      } else if (c - GE::VA < 4 * sl.vbn) vv = __ldg(d.values2 + rowi * M2 + 4 * sl.vb0 + (c - GE::VA));
    }
    S.valS[i] = vv;
  }
  if (tid < NSLOT) {
    // channel slot -> {q', v, wf'} pack; padded slots carry v = 0 and read key column 0
    float v = 0.f, wf[AFT] = {0.f, 0.f, 0.f, 0.f, 0.f};
    int col = 0;
    if (grp) {
      if (tid < 8 * NIA) {
        if (tid < 4 * sl.qan) {
          col = tid;
          v = __ldg(d.v1 + 4 * sl.qa0 + tid);
#pragma unroll
          for (int f = 0; f < AFT; ++f)
            if (f < d.att_filters) wf[f] = __ldg(d.loc_layer_w + (long long)f * A1 + 4 * sl.qa0 + tid) * K2LOG2E;
        }
      } else {
        const int c2 = tid - 8 * NIA;
        if (c2 < 4 * sl.qbn) { col = QA + c2; v = __ldg(d.v2 + 4 * sl.qb0 + c2); }
      }
    }
    S.packA[tid] = make_float4(0.f, v, wf[0], wf[1]);
    S.packB[tid] = make_float4(wf[2], wf[3], wf[4], __int_as_float(col));
  }
  for (int i = tid; i < MAXK * MAXF; i += NT) {
    const int k = i / MAXF, f = i % MAXF;
    S.wconv[i] = (k < d.att_kernel && f < d.att_filters) ? __ldg(d.loc_conv_w + k * d.att_filters + f) : 0.f;
  }
  if (tid < MAXF) S.bconv[tid] = (tid < d.att_filters) ? __ldg(d.loc_conv_b + tid) : 0.f;
  for (int i = tid; i < 2 * NB * KREC; i += NT) S.xrec[i] = 0.f;
  for (int i = tid; i < TtP + 2 * HALO; i += NT) S.aprev[i] = 0.f;
  for (int i = tid; i < TtP; i += NT) {
    S.alphaS[i] = (d.mode == 2 && i == 0) ? 1.f : 0.f;   // alpha_0 = one-hot(0), forward_attention.py:131-133
    S.w1S[i] = 0.f; S.w2S[i] = 0.f; S.softS[i] = 0.f;
  }
  for (int i = tid; i < TtP * MAXF; i += NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * G * TtP; i += NT) S.epart[i] = 0.f;
  for (int i = tid; i < 2 * TtP; i += NT) S.estage[i] = 0.f;
  for (int i = tid; i < RING * 2 * NB * UH; i += NT) S.mk_ring[i] = 0;
  for (int i = tid; i < RING * NB * 64; i += NT) S.xg_ring[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) cl::mbar_init(&S.bars[i], 1);
    cl::fence_mbar_init();
  }
  if (warp == 0) {
    // energy bounds: |e1| <= sum |v1|, |e2| <= sum |v2| (over ALL channels): constant softmax stabilisers
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < A1; c += 32) s1 += fabsf(__ldg(d.v1 + c));
    for (int c = lane; c < A2; c += 32) s2 += fabsf(__ldg(d.v2 + c));
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { S.consts[2] = s1; S.consts[3] = s2; }
  }
  __syncthreads();
  if (warp == 1) {
    // sum of v over the real channels of this CTA's slices: e = sum v tanh = Vs - 2 sum v / (1 + 2^s')
    float s1 = 0.f, s2 = 0.f;
    for (int s = lane; s < NSLOT; s += 32) {
      const float v = S.packA[s].y;
      if (s < 8 * NIA) s1 += v; else s2 += v;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { S.consts[0] = s1; S.consts[1] = s2; }
  }

  // ---------------- P1 role: thread = (k slice = lane, 4 gate columns = warp): weights in registers for all steps
  float w[KREG][4];
#pragma unroll
  for (int i = 0; i < KPT; ++i) {
    float wv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int colc = warp * 4 + c;                                   // CTA-local gate column = gate*16 + unit
      wv[c] = __ldg(d.Wrec + (long long)(lane + 32 * i) * (4 * H) + (colc >> 4) * H + rank * UH + (colc & 15));
    }
    if (i < KREG) {
#pragma unroll
      for (int c = 0; c < 4; ++c) w[i][c] = wv[c];
    } else {
      S.Wsm[(i - KREG) * NT + tid] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    }
  }
  if (tid < 16) S.pt[tid] = 0u;
  // pointwise role: tid < NB*16 -> (pu_u, pu_k)
  const int pu_u = tid >> 4, pu_k = tid & 15;
  const bool pw_act = tid < NB * UH;
  const bool prow_ok = pw_act && (b0 + pu_u) < B;
  float c_st = 0.f, h_st = 0.f;
  // energy role
  const int ep = lane >> 3, ecl = lane & 7;
  const int nact = (warp < CW) ? min(NP, max(0, (alen - 4 * warp + PSL - 1) / PSL)) : 0;
  int echunks = 0;
  for (int w_ = 0; w_ < CW; ++w_) echunks += min(NP, max(0, (alen - 4 * w_ + PSL - 1) / PSL));
  const uint32_t RX_Q = 16u * (uint32_t)(sl.qan + sl.qbn) * 16u;
  const uint32_t RX_E = (uint32_t)G * 2u * 16u * (uint32_t)echunks;
  // saver role (warps CW..15)
  const int sv = tid - CW * 32;
  const bool saver = sv >= 0;
  constexpr int NSV = NT - CW * 32;   // 96

  __syncthreads();
  const float Vs1 = S.consts[0], Vs2 = S.consts[1];
  const float stab1 = S.consts[2], stab2 = S.consts[3];
  const bool use_max1 = stab1 > STAB_MAX, use_max2 = stab2 > STAB_MAX;
  cluster.sync();

  auto prefetch = [&](int t) {
    if (saver && t < Td) {
      const int slot = t % RING;
      for (int e = sv; e < NB * 16; e += NSV) {               // x-projection rows: (u, gate, quad of units)
        const int u = e >> 4, g4 = (e >> 2) & 3, q4 = e & 3;
        if (b0 + u < B)
          cp_async16(&S.xg_ring[(slot * NB + u) * 64 + g4 * 16 + 4 * q4],
                     d.xg + ((long long)t * B + b0 + u) * (4 * H) + g4 * H + rank * UH + 4 * q4);
      }
      if (sv < 2 * NB) {
        const int which = sv / NB, u = sv % NB;
        const uint8_t* src = which ? d.mask_h : d.mask_c;
        if (src && b0 + u < B) cp_async16(S.mk_ring + ((slot * 2 + which) * NB + u) * UH, src + ((long long)t * B + b0 + u) * H + rank * UH);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int t = 0; t < PFD; ++t) prefetch(t);

  PT2_DECL
#pragma unroll 1
  for (int t = 0; t < Td; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    const uint32_t par = (uint32_t)(t >> 1) & 1u;
    const bool last = (t + 1 == Td);
    PT2(15)
    prefetch(t + PFD);
    cp_async_wait<PFD>();
    // location features f = conv1d(a_{t-1}) of this step (forward_attention.py:98-100): independent of the exchange we wait for
    if (grp && warp < CW) arnn::location_features<AFT>(S.fS, S.aprev, S.wconv, S.bconv, alen, d.att_kernel, pl, tid, CW * 32);
    if (t > 0) cl::mbar_wait(&barX[cur], (uint32_t)((t - 1) >> 1) & 1u);    // h(t), ctx(t-1) of every CTA have landed
    if (tid == 0) {
      if (!last) cl::mbar_arrive_expect_tx(&barX[nxt], RX_X);
      if (grp) {
        cl::mbar_arrive_expect_tx(&barQ[cur], RX_Q);
        cl::mbar_arrive_expect_tx(&barE[cur], RX_E);
      }
    }
    PT2(0)
    // ======================= P1: gate GEMM
    {
      float acc[NB * 4];
#pragma unroll
      for (int j = 0; j < NB * 4; ++j) acc[j] = 0.f;
      const float* xr = S.xrec + cur * NB * KREC;
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const int k = lane + 32 * i;
        float x[NB], wc[4];
#pragma unroll
        for (int u = 0; u < NB; ++u) x[u] = xr[u * KREC + k];
        if (i < KREG) {
#pragma unroll
          for (int c = 0; c < 4; ++c) wc[c] = w[i][c];
        } else {
          const float4 w4 = S.Wsm[(i - KREG) * NT + tid];
          wc[0] = w4.x; wc[1] = w4.y; wc[2] = w4.z; wc[3] = w4.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int u = 0; u < NB; ++u) acc[u * 4 + c] = fmaf(wc[c], x[u], acc[u * 4 + c]);
      }
      float a16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a16[j] = acc[j];
      float v = cl::reduce_scatter16(a16, lane);          // lane L: utterance (L&15)>>2, column L&3
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 16) S.gsm[(lane >> 2) * 64 + warp * 4 + (lane & 3)] = v;
      if (NB == 5) {
        // fifth utterance: 4 values over 32 lanes
        const bool up = (lane & 16) != 0;
        float k0 = up ? acc[NB * 4 - 2] : acc[NB * 4 - 4], k1 = up ? acc[NB * 4 - 1] : acc[NB * 4 - 3];
        const float s0 = up ? acc[NB * 4 - 4] : acc[NB * 4 - 2], s1 = up ? acc[NB * 4 - 3] : acc[NB * 4 - 1];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        const bool up8 = (lane & 8) != 0;
        float kk = up8 ? k1 : k0;
        kk += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
        kk += __shfl_xor_sync(0xffffffffu, kk, 4);
        kk += __shfl_xor_sync(0xffffffffu, kk, 2);
        kk += __shfl_xor_sync(0xffffffffu, kk, 1);
        if ((lane & 7) == 0) S.gsm[(NB - 1) * 64 + warp * 4 + 2 * (lane >> 4) + ((lane >> 3) & 1)] = kk;
      }
    }
    PT2(1)
    __syncthreads();   // #1
    PT2(2)
    // ======================= P2a: LSTM cell, zoneout, state exchange
    if (warp < (NB * UH + 31) / 32) {
      float gi = 0.f, gj = 0.f, gf = 0.f, go = 0.f, h_new = 0.f;
      const float c_old = c_st, h_old = h_st;
      if (prow_ok) {
        const float* xg = S.xg_ring + ((t % RING) * NB + pu_u) * 64;
        const float* gs = S.gsm + pu_u * 64;
        const uint8_t* mr = S.mk_ring + (t % RING) * 2 * NB * UH;
        const float mc = d.mask_c ? (float)mr[(0 * NB + pu_u) * UH + pu_k] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mr[(1 * NB + pu_u) * UH + pu_k] : (1.f - d.zh);
        gi = fsigmoid(gs[0 * 16 + pu_k] + xg[0 * 16 + pu_k]);
        gj = ftanh(gs[1 * 16 + pu_k] + xg[1 * 16 + pu_k]);
        gf = fsigmoid(gs[2 * 16 + pu_k] + xg[2 * 16 + pu_k] + d.forget_bias);
        go = fsigmoid(gs[3 * 16 + pu_k] + xg[3 * 16 + pu_k]);
        const float c_new = gf * c_st + gi * gj;
        h_new = go * ftanh(c_new);
        c_st = c_st + mc * (c_new - c_st);
        h_st = h_st + mh * (h_new - h_st);
      }
      if (pw_act) {
        S.hnS[tid] = h_new;
        S.hstS[tid] = h_st;
        S.save1[0 * NB * UH + tid] = gi; S.save1[1 * NB * UH + tid] = gj; S.save1[2 * NB * UH + tid] = gf; S.save1[3 * NB * UH + tid] = go;
        S.save1[4 * NB * UH + tid] = c_old; S.save1[5 * NB * UH + tid] = h_old; S.save1[6 * NB * UH + tid] = h_new;
      }
    }
    PT2(3)
    __syncthreads();   // #2
    PT2(4)
    // ======================= P2b: partial queries of my 16 units -> the groups
    if (tid < NB * 64) {
      const int u = tid >> 6, quad = tid & 63;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* hn = S.hnS + u * UH;
#pragma unroll
      for (int k = 0; k < UH; ++k) {
        const float hk = hn[k];
        const float4 w4 = *reinterpret_cast<const float4*>(&S.WqS[k * QT + 4 * quad]);
        a.x = fmaf(hk, w4.x, a.x); a.y = fmaf(hk, w4.y, a.y); a.z = fmaf(hk, w4.z, a.z); a.w = fmaf(hk, w4.w, a.w);
      }
      int g, off;
      if (quad < A1 / 4) { g = quad / QAq; off = 4 * (quad - g * QAq); }
      else { const int q2 = quad - A1 / 4; g = q2 / QBq; off = QA + 4 * (q2 - g * QBq); }
      const int dst = u * G + g;
      st_async_v4(cl::mapa(cl::smem_u32(&S.qin[rank * QC + off]), dst), a.x, a.y, a.z, a.w, cl::mapa(cl::smem_u32(&barQ[cur]), dst));
    } else if (warp < CW) {
      // carried state h(t+1) of my 16 units -> every CTA's next-step input row: (utterance, unit quad) x destination
      if (!last) {
        for (int e = tid - NB * 64; e < NB * 4 * CS; e += CW * 32 - NB * 64) {
          const int chunk = e % (NB * 4), dst = e / (NB * 4);
          const int u = chunk >> 2, q4 = chunk & 3;
          const float4 h4 = *reinterpret_cast<const float4*>(&S.hstS[u * UH + 4 * q4]);
          st_async_v4(cl::mapa(cl::smem_u32(&S.xrec[(nxt * NB + u) * KREC + (M1 + M2) + rank * UH + 4 * q4]), dst), h4.x, h4.y, h4.z, h4.w,
                      cl::mapa(cl::smem_u32(&barX[nxt]), dst));
        }
      }
    } else if (saver) {
      // saver: LSTM activations of step t (7 arrays x NB utterances x 4 float4)
      for (int e = sv; e < 7 * NB * 4; e += NSV) {
        const int arr = e / (NB * 4), u = (e / 4) % NB, q4 = e & 3;
        const int row = b0 + u;
        if (row >= B) continue;
        const float4 v = *reinterpret_cast<const float4*>(&S.save1[(arr * NB + u) * UH + 4 * q4]);
        const long long rb = (long long)t * B + row;
        if (arr < 4) { if (d.gates) *reinterpret_cast<float4*>(d.gates + rb * (4 * H) + arr * H + rank * UH + 4 * q4) = v; }
        else if (arr == 4) { if (d.c_prev) *reinterpret_cast<float4*>(d.c_prev + rb * H + rank * UH + 4 * q4) = v; }
        else if (arr == 5) { if (d.h_prev) *reinterpret_cast<float4*>(d.h_prev + rb * H + rank * UH + 4 * q4) = v; }
        else *reinterpret_cast<float4*>(d.x2 + rb * X2W + rank * UH + 4 * q4) = v;
      }
    }
    PT2(5)
    if (grp) {
      cl::mbar_wait(&barQ[cur], par);   // the 16 partial queries of my slice
      PT2(6)
      // ======================= P3a: q = sum of the partials
      if (tid < 4 * QC) {
        const int s = tid >> 2, part = tid & 3;
        float v = (S.qin[part * QC + s] + S.qin[(part + 4) * QC + s]) + (S.qin[(part + 8) * QC + s] + S.qin[(part + 12) * QC + s]);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (part == 0) {
          S.qsS[s] = v;
          const int slot = (s < QA) ? s : 8 * NIA + (s - QA);
          reinterpret_cast<float*>(&S.packA[slot])[0] = v * K2LOG2E;
        }
      }
    }
    PT2(7)
    __syncthreads();   // #3
    if (grp) {
      // ======================= P3b: partial energies over my channel slice
      if (warp < CW) {
        const int slot0 = warp * 4 + ep;
        // the number of active passes is warp-uniform: one branch-free instance per count (a predicate per pass would
        // split the passes into basic blocks and serialise their ex2 / rcp chains)
        switch (nact) {
          case 1: energy_passes<NB, 1>(S, slot0, ecl, Tt, Vs1, Vs2); break;
          case 2: energy_passes<NB, 2>(S, slot0, ecl, Tt, Vs1, Vs2); break;
          case 3: energy_passes<NB, 3>(S, slot0, ecl, Tt, Vs1, Vs2); break;
          case 4: if (NP >= 4) energy_passes<NB, (NP >= 4 ? 4 : 1)>(S, slot0, ecl, Tt, Vs1, Vs2); break;
          default: break;
        }
        __syncwarp();
        // one 16-byte st.async per (pass, mechanism, group member): the warp's 4 consecutive positions
        if (lane < nact * 2 * G) {
          const int m = lane / (2 * G), rem = lane % (2 * G), att = rem / G, gd = rem % G;
          const int j0 = warp * 4 + PSL * m;
          const float4 e4 = *reinterpret_cast<const float4*>(&S.estage[att * TtP + j0]);
          const int dst = au * G + gd;
          st_async_v4(cl::mapa(cl::smem_u32(&S.epart[(att * G + ag) * TtP + j0]), dst), e4.x, e4.y, e4.z, e4.w,
                      cl::mapa(cl::smem_u32(&barE[cur]), dst));
        }
      } else if (saver && arow_ok && d.q_save) {
        // saver: processed queries (for the backward pass)
        for (int q = sv; q < GE::QAq + GE::QBq; q += NSV) {
          const bool a1 = q < QAq;
          const bool real = a1 ? (q < sl.qan) : (q - QAq < sl.qbn);
          if (!real) continue;
          const int gcol = a1 ? 4 * (sl.qa0 + q) : A1 + 4 * (sl.qb0 + q - QAq);
          *reinterpret_cast<float4*>(d.q_save + ((long long)t * B + arow) * QT + gcol) = *reinterpret_cast<const float4*>(&S.qsS[4 * q]);
        }
      }
      PT2(8)
      if (warp < 2 * SMW) cl::mbar_wait(&barE[cur], par);   // partial energies of the whole group
      PT2(9)
      // ======================= P4: softmax, forward recursion (one position per thread)
      if (warp < SMW) {
        const int j = tid;
        const bool in = j < alen;
        float e = -INFINITY;
        if (in) {
          e = 0.f;
#pragma unroll
          for (int g = 0; g < G; ++g) e += S.epart[(0 * G + g) * TtP + j];
        }
        const float mx = use_max1 ? gmax<SMW>(e, S.red, warp, lane, 2) : stab1;
        const float pexp = in ? __expf(e - mx) : 0.f;
        float mixp = 0.f;
        if (d.mode == 2 && j < TtP) {
          const float apm1 = (j > 0) ? S.alphaS[j - 1] : 0.f;
          mixp = (0.5f * S.alphaS[j] + 0.5f * apm1 + 1e-7f) * pexp;   // u stays 0.5 without the agent (forward_attention.py:109,116,135)
        }
        float s1 = pexp, s2 = mixp;
        gsum2<SMW>(s1, s2, S.red, warp, lane, 2);                    // the barrier inside also orders the alphaS reads above
        if (j < TtP) {
          const float a = (s1 > 0.f) ? pexp * __fdividef(1.f, s1) : 0.f;
          const float wgt = (d.mode == 2) ? ((s2 > 0.f) ? mixp * __fdividef(1.f, s2) : 0.f) : a;
          if (d.mode == 2) S.alphaS[j] = wgt;
          S.w1S[j] = wgt;
          S.softS[j] = a;
          S.aprev[HALO + j] = a;
        }
      } else if (warp < 2 * SMW) {
        const int j = tid - SMW * 32;
        const bool in = j < alen;
        float e = -INFINITY;
        if (in) {
          e = 0.f;
#pragma unroll
          for (int g = 0; g < G; ++g) e += S.epart[(1 * G + g) * TtP + j];
        }
        const float mx = use_max2 ? gmax<SMW>(e, S.red + 32, warp - SMW, lane, 3) : stab2;
        float pexp = in ? __expf(e - mx) : 0.f, dummy = 0.f;
        float s1 = pexp;
        gsum2<SMW>(s1, dummy, S.red + 32, warp - SMW, lane, 3);
        if (j < TtP) S.w2S[j] = (s1 > 0.f) ? pexp * __fdividef(1.f, s1) : 0.f;
      }
    }
    PT2(10)
    __syncthreads();   // #4
    PT2(11)
    if (grp) {
      // ======================= P5: context partial sums over my value columns: warp = position group, lane = column quad
      if (warp < CW) {
        if (lane < VCq) {
          const float* wS = (lane < VAq) ? S.w1S : S.w2S;
          // weights past the source length are exactly zero and rows up to TtP = CW * (TtP / CW) exist (zero-filled), so the
          // trip count is rounded up to 4 and the body needs no predicates
          float4 acc[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int nj = (alen - warp + CW - 1) / CW;
          for (int jj = 0; jj < nj; jj += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = warp + CW * (jj + q);
              const float w0 = wS[j];
              const float4 v0 = *reinterpret_cast<const float4*>(&S.valS[j * VC + 4 * lane]);
              acc[q].x = fmaf(w0, v0.x, acc[q].x); acc[q].y = fmaf(w0, v0.y, acc[q].y);
              acc[q].z = fmaf(w0, v0.z, acc[q].z); acc[q].w = fmaf(w0, v0.w, acc[q].w);
            }
          }
          *reinterpret_cast<float4*>(&S.cpart[warp * VC + 4 * lane]) =
              make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                          (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
        }
      } else if (saver && arow_ok) {
        // saver: alignment history of step t; the three rows are spread over the group members
        const float* src = (ag == 0) ? S.w1S : (ag == 1) ? S.softS : S.w2S;
        float* dst = (ag == 0) ? d.align1 : (ag == 1) ? d.soft1 : (ag == 2) ? d.align2 : nullptr;
        if (dst) {
          dst += ((long long)t * B + arow) * Tt;
          if ((Tt & 3) == 0) {
            for (int q = sv; q < Tt / 4; q += NSV) *reinterpret_cast<float4*>(dst + 4 * q) = *reinterpret_cast<const float4*>(src + 4 * q);
          } else {
            for (int j = sv; j < Tt; j += NSV) dst[j] = src[j];
          }
        }
      }
    }
    PT2(12)
    __syncthreads();   // #5
    PT2(13)
    if (grp) {
      if (tid < CS * VCq) {
        // context slice of my utterance -> every CTA's x[u][k] of the next step: thread = (column quad, destination CTA)
        const int quad = tid % VCq, dst = tid / VCq;
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w_ = 0; w_ < CW; ++w_) {
          const float4 p4 = *reinterpret_cast<const float4*>(&S.cpart[w_ * VC + 4 * quad]);
          c4.x += p4.x; c4.y += p4.y; c4.z += p4.z; c4.w += p4.w;
        }
        const bool a1 = quad < VAq;
        const bool real = a1 ? (quad < sl.van) : (quad - VAq < sl.vbn);
        if (real && !last) {
          const int k = a1 ? 4 * (sl.va0 + quad) : M1 + 4 * (sl.vb0 + quad - VAq);
          st_async_v4(cl::mapa(cl::smem_u32(&S.xrec[(nxt * NB + au) * KREC + k]), dst), c4.x, c4.y, c4.z, c4.w,
                      cl::mapa(cl::smem_u32(&barX[nxt]), dst));
        }
      } else if (saver && arow_ok) {
        // saver: context of step t -> x2[:, H:]
        for (int quad = sv; quad < VCq; quad += NSV) {
          const bool a1 = quad < VAq;
          const bool real = a1 ? (quad < sl.van) : (quad - VAq < sl.vbn);
          if (!real) continue;
          float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w_ = 0; w_ < CW; ++w_) {
            const float4 p4 = *reinterpret_cast<const float4*>(&S.cpart[w_ * VC + 4 * quad]);
            c4.x += p4.x; c4.y += p4.y; c4.z += p4.z; c4.w += p4.w;
          }
          const int k = a1 ? 4 * (sl.va0 + quad) : M1 + 4 * (sl.vb0 + quad - VAq);
          *reinterpret_cast<float4*>(d.x2 + ((long long)t * B + arow) * X2W + H + k) = c4;
        }
      }
    }
    PT2(14)
  }
  PT2_FLUSH(Td)
  cp_async_wait<0>();
  cluster.sync();
}

template <int NB>
static size_t fwd2_smem_bytes(int np) {
  FwdSmem2<NB> S;
  return S.carve(nullptr, np);
}

template <typename Kern>
static int launch16(Kern kern, const satk_attn_rnn_fwd_desc& d, int nb, size_t smem, cudaStream_t st) {
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((d.B + nb - 1) / nb) * CS);
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

// The second-generation kernels cover the dual-source decoder of the shipped configurations; everything else (single
// attention, additive first mechanism, cumulative weights, transition agent, > 5 location filters, Tt > 192) stays on the
// first-generation kernels.
bool v2_eligible(const satk_attn_rnn_fwd_desc* d) {
  const char* e = getenv("SATK_ATTN_GEN");
  if (e && e[0] == '1') return false;
  return d->H == H && d->M1 == M1 && d->A1 == A1 && d->A2 == A2 && d->M2 == M2 && d->att_kernel >= 1 && d->att_kernel <= MAXK &&
         d->att_filters >= 1 && d->att_filters <= AFT && (d->mode == 1 || d->mode == 2) && !d->cumulative && !d->agent_w &&
         d->Tt >= 1 && d->Tt <= SMW * 32 && d->lengths && d->keys2 && d->values2 && d->Wq2 && d->v2 && d->align2;
}

size_t attn_rnn2_fwd_smem(int nb, int Tt);

// utterances per cluster: 4 (groups of 4 CTAs) while the batch fits the co-resident clusters, else 5 (groups of 3)
int v2_pick_nb(const satk_attn_rnn_fwd_desc* d) {
  const char* e = getenv("SATK_ATTN_NB");
  if (e && (e[0] == '4' || e[0] == '5')) return e[0] - '0';
  if (d->B <= 28) return 4;
  return (attn_rnn2_fwd_smem(5, d->Tt) <= 227 * 1024) ? 5 : 4;   // long texts: the narrower slices of the 4-CTA groups
}

int attn_rnn2_fwd_launch(const satk_attn_rnn_fwd_desc* d, cudaStream_t st) {
  const int nb = v2_pick_nb(d);
  const int np = (d->Tt + PSL - 1) / PSL;
  SATK_CHECK_ARG(np <= 4, "attn_rnn2_fwd: Tt=%d out of range", d->Tt);
  const int npv = np <= 3 ? 3 : 4;
  const size_t smem = nb == 5 ? fwd2_smem_bytes<5>(npv) : fwd2_smem_bytes<4>(npv);
  SATK_CHECK_ARG(smem <= 227 * 1024, "attn_rnn2_fwd: Tt=%d needs %zu B of shared memory (> 227 KB)", d->Tt, smem);
  if (nb == 5) {
    if (npv == 3) return launch16(attn_rnn2_fwd_kernel<5, 3>, *d, nb, smem, st);
    return launch16(attn_rnn2_fwd_kernel<5, 4>, *d, nb, smem, st);
  }
  if (npv == 3) return launch16(attn_rnn2_fwd_kernel<4, 3>, *d, nb, smem, st);
  return launch16(attn_rnn2_fwd_kernel<4, 4>, *d, nb, smem, st);
}

size_t attn_rnn2_fwd_smem(int nb, int Tt) {
  const int np = (Tt + PSL - 1) / PSL;
  const int npv = np <= 3 ? 3 : 4;
  return nb == 5 ? fwd2_smem_bytes<5>(npv) : fwd2_smem_bytes<4>(npv);
}

int attn2_fwd_phase_cycles(long long* out16) {
#ifdef SATK_PHASE_TIMING
  SATK_CUDA(cudaMemcpyFromSymbol(out16, satk::g_phase, sizeof(long long) * 16));
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
#endif
  return 0;
}

}  // namespace arnn2
}  // namespace satk

