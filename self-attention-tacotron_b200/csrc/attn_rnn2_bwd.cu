// Attention-RNN backward, second generation (BPTT through LSTM-1 + both attention mechanisms), one launch for all Td steps;
// cluster geometry of attn_rnn2.cuh (NB utterances per 16-CTA cluster, groups of G CTAs per utterance).  Per step, descending t:
//   BA0  d(ctx_t) of my value columns / d(h_t) of my units = external + sum of the 16 partial products of step t+1       [C]
//   BA1  partial d(alignment weights) = d(ctx) . values^T over my value columns            -> the group                  [W]
//   BA2a forward-attention recursion + softmax backward (redundantly per group CTA, one position per thread) -> d(energies)
//   BA2b energy backward over my channel slice (recomputes tanh through ex2 / rcp): d(query) slice -> every CTA          [Q]
//        partial d(location features) -> partial d(previous alignments) -> the group (used by the next step)             [W]
//   BB   d(out1) = external + dq . Wq^T for my 16 units; LSTM cell backward -> d(gates) of my 64 gate columns
//   BC   d([ctx | h])(t-1) = d(gates) . Wrec^T restricted to MY gate columns: thread k owns row k of the recurrent kernel
//        (64 weights in registers), so the product needs no reduction inside the CTA; the 16 partial results are summed by
//        their consumers (reduce-scatter straight to the owners of the context columns / hidden units)                    [C]
// Three point-to-point exchanges per step.  Gradients that do not feed the recurrence (d(keys), d(v), d(location layer /
// convolution)) are NOT computed here: the kernel saves d(energies) and `satk_attn_energy_grad` (attn_energy_grad.cu)
// recomputes the tanh terms for all (step, utterance) pairs in parallel.  Weight gradients that are dense over time (dWrec,
// dWq, dW_memory, dvalues) are plain GEMMs of the caller over the saved d(gates), dq and total d(ctx).
// Reference semantics: see attn_rnn2_fwd.cu.
#include <stdlib.h>
#include "attn_rnn2.cuh"

namespace satk {
namespace arnn2 {

using cl::cp_async4;
using cl::cp_async_commit;
using cl::cp_async_wait;
using cl::st_async_v4;
using cl::st_async_f32;

constexpr int NTB = 512;     // thread k owns row k of the recurrent kernel; the last 32 rows sit in shared memory
constexpr int WRS = 72;      // row stride of those rows: 4 rows x 2 column halves x 16 bytes hit 32 distinct banks
constexpr int RINGB = 4;     // ring slots (time-indexed)
constexpr int PFDB = 2;      // prefetch distance (steps)
constexpr int WQS = 260;     // row stride of the per-unit query weights: 8 units x 16 bytes hit 32 distinct banks
constexpr int NSVB = NTB - CW * 32;   // 96 service threads (global-memory traffic)
constexpr int DGG = 20;      // floats per gate in the d(gates) staging: the 4 column quarters of BC then hit disjoint banks
constexpr int DGS = 4 * DGG; // floats per batch row

#ifdef SATK_PHASE_TIMING
#define PT2_DECL unsigned pt_last = (unsigned)clock();
#define PT2(i) if (tid == 0) { const unsigned pt_n = (unsigned)clock(); S.pt[i] += pt_n - pt_last; pt_last = pt_n; }
#define PT2_FLUSH(n) if (blockIdx.x == 0 && tid == 0) { for (int i_ = 0; i_ < 16; ++i_) satk::g_phase[i_] = (long long)(S.pt[i_] / (unsigned)(n)); }
#else
#define PT2_DECL
#define PT2(i)
#define PT2_FLUSH(n)
#endif


// ---- tensor memory as a parking space for the stationary recurrent weights -------------------------------------------------
// The 64 weights a thread needs for BC are read-only for the whole launch, but pinned in registers they leave the energy phase
// ~30 free registers (serialised ex2 / rcp chains, spills).  Tensor memory (256 KB per SM) is otherwise unused by this kernel:
// each warp stores its 32 x 64 weights once (tcgen05.st, 32x32b: lane i <-> TMEM lane 32*(warp%4)+i, 64 columns per warp) and
// reads them back right before BC of every step (tcgen05.ld, ~8 KB per warp from a ~TB/s datapath).
#define SATK_TM_REGS32(r, o) \
  "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), \
  "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]), "r"(r[o + 16]), \
  "r"(r[o + 17]), "r"(r[o + 18]), "r"(r[o + 19]), "r"(r[o + 20]), "r"(r[o + 21]), "r"(r[o + 22]), "r"(r[o + 23]), "r"(r[o + 24]), \
  "r"(r[o + 25]), "r"(r[o + 26]), "r"(r[o + 27]), "r"(r[o + 28]), "r"(r[o + 29]), "r"(r[o + 30]), "r"(r[o + 31])
#define SATK_TM_OUT32(r, o) \
  "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7]), \
  "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15]), \
  "=r"(r[o + 16]), "=r"(r[o + 17]), "=r"(r[o + 18]), "=r"(r[o + 19]), "=r"(r[o + 20]), "=r"(r[o + 21]), "=r"(r[o + 22]), "=r"(r[o + 23]), \
  "=r"(r[o + 24]), "=r"(r[o + 25]), "=r"(r[o + 26]), "=r"(r[o + 27]), "=r"(r[o + 28]), "=r"(r[o + 29]), "=r"(r[o + 30]), "=r"(r[o + 31])
#define SATK_TM_LIST32 \
  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, " \
  "%28, %29, %30, %31}"
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[64], int o) {
  if (o == 0)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], " SATK_TM_LIST32 ";" ::SATK_TM_REGS32(r, 0), "r"(taddr) : "memory");
  else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], " SATK_TM_LIST32 ";" ::SATK_TM_REGS32(r, 32), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " SATK_TM_LIST32 ", [%32];" : SATK_TM_OUT32(r, 0) : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " SATK_TM_LIST32 ", [%32];" : SATK_TM_OUT32(r, 32) : "r"(taddr + 32u));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
constexpr int TM_COLS = 256;   // 4 warps per TMEM lane quarter x 64 columns

template <int NB>
struct BwdSmem2 {
  using GE = Geo<NB>;
  static constexpr int RB = GE::QC + GE::VC + 6 * NB * UH;   // floats per ringB slot: q slice, external d(ctx) slice, pointwise inputs
  int TtP, DFW;
  float *keyS, *valS, *WqU, *fS, *dfS, *wconv, *bconv, *aprev, *alphaPrevS, *dalpha_carry, *dmixS, *deS, *dwst, *dwpart, *dstate_part,
      *inC, *inH, *dctxS, *dhS, *dqst, *dqS, *dqB, *dgS, *ringA, *ringB, *red, *WragS;
  float4 *packA, *packB;
  uint8_t* mk_ring;
  unsigned* pt;
  uint64_t* bars;   // [0..1] C, [2..3] W, [4..5] Q
  __host__ __device__ static size_t al4(size_t n) { return (n + 3) & ~(size_t)3; }
  __host__ __device__ size_t carve(float* base, int np) {
    TtP = np * PSL;
    DFW = TtP + 2 * HALO;
    float* p = base;
    keyS = p; p += (size_t)TtP * GE::KSTR;
    valS = p; p += (size_t)TtP * GE::VC;
    WqU = p; p += UH * WQS;
    WragS = p; p += (KREC - NTB) * WRS;                 // rows 512..543 of the recurrent kernel, my 64 gate columns
    packA = reinterpret_cast<float4*>(p); p += 4 * GE::NSLOT;
    packB = reinterpret_cast<float4*>(p); p += 4 * GE::NSLOT;
    fS = p; p += (size_t)TtP * MAXF;
    dfS = p; p += (size_t)AFT * DFW;                    // d(location features), filter-major [f][HALO + j]
    wconv = p; p += MAXK * MAXF;
    bconv = p; p += MAXF;
    aprev = p; p += al4(TtP + 2 * HALO);
    alphaPrevS = p; p += TtP + 8;
    dalpha_carry = p; p += TtP;
    dmixS = p; p += TtP + 8;
    deS = p; p += 2 * (size_t)TtP;
    dwst = p; p += 2 * (size_t)TtP;
    dwpart = p; p += 2 * GE::G * (size_t)TtP;           // [att][source member][j]
    dstate_part = p; p += 2 * GE::G * (size_t)TtP;      // [parity][source member][j]
    inC = p; p += 16 * GE::VC;                          // [source CTA][my value column]
    inH = p; p += 2 * 16 * NB * UH;                     // [parity][source CTA][u][unit]
    dctxS = p; p += GE::VC;
    dhS = p; p += NB * UH;
    dqst = p; p += CW * GE::NSLOT;                      // per-warp d(query) partials
    dqS = p; p += GE::QC;
    dqB = p; p += NB * QT;                              // d(query) of every utterance of the cluster
    dgS = p; p += NB * DGS;                             // d(gates) of my 64 gate columns, [row][gate][DGG] (16 used)
    ringA = p; p += RINGB * 3 * (size_t)TtP;            // soft1 / align1 / align2 of time tau
    ringB = p; p += RINGB * (size_t)RB;
    red = p; p += 96;
    mk_ring = reinterpret_cast<uint8_t*>(p); p += RINGB * 2 * NB * UH / 4;
    pt = reinterpret_cast<unsigned*>(p); p += 16;
    bars = reinterpret_cast<uint64_t*>(p); p += 2 * 6;
    return (size_t)(p - base) * sizeof(float);
  }
};

// BA1 for NACT position passes of one warp: partial d(weights)[j] = d(ctx) . values[j, my columns] (both mechanisms) -> dwst
template <int NB, int NACT>
__device__ __forceinline__ void dweights_passes(const BwdSmem2<NB>& S, int slot0, int ecl, int Tt) {
  using GE = Geo<NB>;
  constexpr int VC = GE::VC, VA = GE::VA, VAq = GE::VAq, VBq = GE::VBq;
  constexpr int NQ = (VAq + 7) / 8;
  float4 dc[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
    dc[i] = (ecl + 8 * i < VAq) ? *reinterpret_cast<const float4*>(&S.dctxS[4 * (ecl + 8 * i)]) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 d2 = (ecl < VBq) ? *reinterpret_cast<const float4*>(&S.dctxS[VA + 4 * ecl]) : make_float4(0.f, 0.f, 0.f, 0.f);
  float a1[NACT], a2[NACT];
#pragma unroll
  for (int m = 0; m < NACT; ++m) {
    const float* vr = S.valS + min(slot0 + PSL * m, Tt - 1) * VC;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float4 v4 = *reinterpret_cast<const float4*>(vr + 4 * min(ecl + 8 * i, VAq - 1));
      s = fmaf(dc[i].x, v4.x, s); s = fmaf(dc[i].y, v4.y, s); s = fmaf(dc[i].z, v4.z, s); s = fmaf(dc[i].w, v4.w, s);
    }
    const float4 w4 = *reinterpret_cast<const float4*>(vr + VA + 4 * min(ecl, VBq - 1));
    a1[m] = s;
    a2[m] = fmaf(d2.x, w4.x, fmaf(d2.y, w4.y, fmaf(d2.z, w4.z, d2.w * w4.w)));
  }
#pragma unroll
  for (int m = 0; m < NACT; ++m) {
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
      a1[m] += __shfl_xor_sync(0xffffffffu, a1[m], o);
      a2[m] += __shfl_xor_sync(0xffffffffu, a2[m], o);
    }
    if (ecl == 0) {
      S.dwst[0 * S.TtP + slot0 + PSL * m] = a1[m];
      S.dwst[1 * S.TtP + slot0 + PSL * m] = a2[m];
    }
  }
}

// BA2b for NACT position passes of one warp: d(pre-activation) ds = de * v * (1 - tanh^2) over my channel slice;
// d(query)[c] = sum_j ds (per-warp partials -> dqst), d(location features)[j][f] = sum_c ds Wf[f][c] (partial over my slice -> dfS)
template <int NB, int NACT>
__device__ __forceinline__ void energy_bwd_passes(const BwdSmem2<NB>& S, int slot0, int ecl, int lane, int warp, int Tt) {
  using GE = Geo<NB>;
  constexpr int NIA = GE::NIA, NIB = GE::NIB, KSTR = GE::KSTR, NSLOT = GE::NSLOT;
  float fv[NACT][AFT], dfp[NACT][AFT], de1[NACT], de2[NACT];
  const float* krow[NACT];
#pragma unroll
  for (int m = 0; m < NACT; ++m) {
    const int jr = slot0 + PSL * m, j = min(jr, Tt - 1);
    krow[m] = S.keyS + j * KSTR;
    const float4 f4 = *reinterpret_cast<const float4*>(&S.fS[j * MAXF]);
    fv[m][0] = f4.x; fv[m][1] = f4.y; fv[m][2] = f4.z; fv[m][3] = f4.w;
    fv[m][4] = S.fS[j * MAXF + 4];
    de1[m] = (jr < Tt) ? S.deS[0 * S.TtP + j] : 0.f;
    de2[m] = (jr < Tt) ? S.deS[1 * S.TtP + j] : 0.f;
#pragma unroll
    for (int f = 0; f < AFT; ++f) dfp[m][f] = 0.f;
  }
  float* dqw = S.dqst + warp * NSLOT;
  // two channel iterations at a time: 2 x NACT independent ex2 / rcp chains in flight
  static_assert(NIA % 2 == 0 || NIA == 7, "channel iterations are walked in pairs");
#pragma unroll 1
  for (int i = 0; i + 1 < NIA; i += 2) {
    const float4 pa0 = S.packA[ecl + 8 * i], pb0 = S.packB[ecl + 8 * i];
    const float4 pa1 = S.packA[ecl + 8 * i + 8], pb1 = S.packB[ecl + 8 * i + 8];
    const int col0 = __float_as_int(pb0.w), col1 = __float_as_int(pb1.w);
    float s0[NACT], s1[NACT];
#pragma unroll
    for (int m = 0; m < NACT; ++m) {
      float a = krow[m][col0] + pa0.x, c = krow[m][col1] + pa1.x;
      a = fmaf(fv[m][0], pa0.z, a); c = fmaf(fv[m][0], pa1.z, c);
      a = fmaf(fv[m][1], pa0.w, a); c = fmaf(fv[m][1], pa1.w, c);
      a = fmaf(fv[m][2], pb0.x, a); c = fmaf(fv[m][2], pb1.x, c);
      a = fmaf(fv[m][3], pb0.y, a); c = fmaf(fv[m][3], pb1.y, c);
      a = fmaf(fv[m][4], pb0.z, a); c = fmaf(fv[m][4], pb1.z, c);
      s0[m] = ex2f(a); s1[m] = ex2f(c);
    }
    float dq0 = 0.f, dq1 = 0.f;
#pragma unroll
    for (int m = 0; m < NACT; ++m) {
      const float r0 = rcpf(1.f + s0[m]), r1 = rcpf(1.f + s1[m]);      // tanh = 1 - 2r, 1 - tanh^2 = 4 r (1 - r)
      const float d0 = (de1[m] * pa0.y) * (4.f * r0 * (1.f - r0)), d1 = (de1[m] * pa1.y) * (4.f * r1 * (1.f - r1));
      dq0 += d0; dq1 += d1;
      dfp[m][0] = fmaf(d0, pa0.z, fmaf(d1, pa1.z, dfp[m][0])); dfp[m][1] = fmaf(d0, pa0.w, fmaf(d1, pa1.w, dfp[m][1]));
      dfp[m][2] = fmaf(d0, pb0.x, fmaf(d1, pb1.x, dfp[m][2])); dfp[m][3] = fmaf(d0, pb0.y, fmaf(d1, pb1.y, dfp[m][3]));
      dfp[m][4] = fmaf(d0, pb0.z, fmaf(d1, pb1.z, dfp[m][4]));
    }
    dq0 += __shfl_xor_sync(0xffffffffu, dq0, 8); dq1 += __shfl_xor_sync(0xffffffffu, dq1, 8);
    dq0 += __shfl_xor_sync(0xffffffffu, dq0, 16); dq1 += __shfl_xor_sync(0xffffffffu, dq1, 16);
    if (lane < 8) { dqw[ecl + 8 * i] = dq0; dqw[ecl + 8 * i + 8] = dq1; }
  }
  if (NIA & 1) {
    constexpr int i = NIA - 1;
    const float4 pa = S.packA[ecl + 8 * i], pb = S.packB[ecl + 8 * i];
    const int col = __float_as_int(pb.w);
    float dq = 0.f;
#pragma unroll
    for (int m = 0; m < NACT; ++m) {
      float s = krow[m][col] + pa.x;
      s = fmaf(fv[m][0], pa.z, s); s = fmaf(fv[m][1], pa.w, s);
      s = fmaf(fv[m][2], pb.x, s); s = fmaf(fv[m][3], pb.y, s); s = fmaf(fv[m][4], pb.z, s);
      const float r = rcpf(1.f + ex2f(s));
      const float ds = (de1[m] * pa.y) * (4.f * r * (1.f - r));
      dq += ds;
      dfp[m][0] = fmaf(ds, pa.z, dfp[m][0]); dfp[m][1] = fmaf(ds, pa.w, dfp[m][1]);
      dfp[m][2] = fmaf(ds, pb.x, dfp[m][2]); dfp[m][3] = fmaf(ds, pb.y, dfp[m][3]); dfp[m][4] = fmaf(ds, pb.z, dfp[m][4]);
    }
    dq += __shfl_xor_sync(0xffffffffu, dq, 8);
    dq += __shfl_xor_sync(0xffffffffu, dq, 16);
    if (lane < 8) dqw[ecl + 8 * i] = dq;
  }
#pragma unroll
  for (int i = 0; i < NIB; ++i) {
    const float4 pa = S.packA[8 * NIA + ecl + 8 * i];
    const int col = __float_as_int(S.packB[8 * NIA + ecl + 8 * i].w);
    float dq = 0.f;
#pragma unroll
    for (int m = 0; m < NACT; ++m) {
      const float r = rcpf(1.f + ex2f(krow[m][col] + pa.x));
      dq = fmaf(de2[m] * pa.y, 4.f * r * (1.f - r), dq);
    }
    dq += __shfl_xor_sync(0xffffffffu, dq, 8);
    dq += __shfl_xor_sync(0xffffffffu, dq, 16);
    if (lane < 8) dqw[8 * NIA + ecl + 8 * i] = dq;
  }
  // d(location features): reduce over the 8 channel lanes; the packs hold Wf scaled by K2LOG2E
#pragma unroll
  for (int m = 0; m < NACT; ++m) {
#pragma unroll
    for (int f = 0; f < AFT; ++f) {
      float v = dfp[m][f];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (ecl == 0 && slot0 + PSL * m < Tt) S.dfS[f * S.DFW + HALO + slot0 + PSL * m] = v * (1.f / K2LOG2E);
    }
  }
}

template <int NB, int NP>
__global__ void __launch_bounds__(NTB, 1) attn_rnn2_bwd_kernel(const satk_attn_rnn_bwd_desc dd, float* __restrict__ de_out,
                                                                 int* __restrict__ prog /* [B][2] progress flags or NULL */) {
  // a dependent launch (the streaming energy-gradient kernel) may start as soon as every CTA of this grid is resident
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  using GE = Geo<NB>;
  constexpr int G = GE::G, QA = GE::QA, QC = GE::QC, VA = GE::VA, VC = GE::VC, VAq = GE::VAq, VBq = GE::VBq, VCq = GE::VCq;
  constexpr int NIA = GE::NIA, NSLOT = GE::NSLOT, KSTR = GE::KSTR;
  constexpr int RB = BwdSmem2<NB>::RB;
  const satk_attn_rnn_fwd_desc& d = dd.f;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CS) * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, Td = d.Td;

  // steps [Te, Td) of this cluster's utterances carry exactly zero gradient (satk_attn_rnn_bwd_desc.step_end)
  int Te = Td;
  if (dd.step_end) {
    Te = 1;
    for (int r = 0; r < NB; ++r)
      if (b0 + r < B) Te = max(Te, min(Td, __ldg(dd.step_end + b0 + r)));
  }

  extern __shared__ __align__(16) float smem_raw[];
  BwdSmem2<NB> S;
  S.carve(smem_raw, NP);
  const int TtP = S.TtP, DFW = S.DFW;
  uint64_t* barC = S.bars;
  uint64_t* barW = S.bars + 2;
  uint64_t* barQ = S.bars + 4;

  const bool grp = rank < NB * G;
  const int au = grp ? rank / G : 0, ag = grp ? rank % G : 0;
  const int arow = b0 + au;
  const bool arow_ok = grp && arow < B;
  const int alen = arow_ok ? min((int)d.lengths[arow], Tt) : 0;
  const int pl = (d.att_kernel - 1) / 2;
  const Slice<NB> sl(ag);
  const int Tt4 = (Tt + 3) >> 2;     // 16-byte chunks of a per-position row
  const int DEL = de_row_stride(Tt); // row stride of the d(energies) workspace

  // ---------------- one-time loads
  for (int i = tid; i < TtP * KSTR; i += NTB) {
    const int j = i / KSTR, c = i % KSTR;
    float kv = 0.f;
    if (arow_ok && j < Tt) {
      const long long rowi = (long long)j * B + arow;
      if (c < QA) {
        if (c < 4 * sl.qan) kv = (__ldg(d.keys1 + rowi * A1 + 4 * sl.qa0 + c) + (d.b1 ? __ldg(d.b1 + 4 * sl.qa0 + c) : 0.f)) * K2LOG2E;
      } else if (c < QC) {
        if (c - QA < 4 * sl.qbn) kv = __ldg(d.keys2 + rowi * A2 + 4 * sl.qb0 + (c - QA)) * K2LOG2E;
      }
    }
    S.keyS[i] = kv;
  }
  for (int i = tid; i < TtP * VC; i += NTB) {
    const int j = i / VC, c = i % VC;
    float vv = 0.f;
    if (arow_ok && j < Tt) {
      const long long rowi = (long long)j * B + arow;
      if (c < VA) {
        if (c < 4 * sl.van) vv = __ldg(d.values1 + rowi * M1 + 4 * sl.va0 + c);
      } else if (c - VA < 4 * sl.vbn) vv = __ldg(d.values2 + rowi * M2 + 4 * sl.vb0 + (c - VA));
    }
    S.valS[i] = vv;
  }
  for (int i = tid; i < (KREC - NTB) * WRS; i += NTB) {
    const int r = i / WRS, c = i % WRS;
    S.WragS[i] = (c < 64) ? __ldg(d.Wrec + (long long)(NTB + r) * (4 * H) + (c >> 4) * H + rank * UH + (c & 15)) : 0.f;
  }
  for (int i = tid; i < UH * WQS; i += NTB) {
    const int k = i / WQS, c = i % WQS;
    float v = 0.f;
    if (c < A1) v = __ldg(d.Wq1 + (long long)(rank * UH + k) * A1 + c);
    else if (c < QT) v = __ldg(d.Wq2 + (long long)(rank * UH + k) * A2 + (c - A1));
    S.WqU[i] = v;
  }
  if (tid < NSLOT) {
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
  for (int i = tid; i < MAXK * MAXF; i += NTB) {
    const int k = i / MAXF, f = i % MAXF;
    S.wconv[i] = (k < d.att_kernel && f < d.att_filters) ? __ldg(d.loc_conv_w + k * d.att_filters + f) : 0.f;
  }
  if (tid < MAXF) S.bconv[tid] = (tid < d.att_filters) ? __ldg(d.loc_conv_b + tid) : 0.f;
  for (int i = tid; i < TtP + 2 * HALO; i += NTB) S.aprev[i] = 0.f;
  for (int i = tid; i < AFT * DFW; i += NTB) S.dfS[i] = 0.f;
  for (int i = tid; i < TtP * MAXF; i += NTB) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * G * TtP; i += NTB) { S.dstate_part[i] = 0.f; S.dwpart[i] = 0.f; }
  for (int i = tid; i < TtP; i += NTB) S.dalpha_carry[i] = 0.f;
  for (int i = tid; i < TtP + 8; i += NTB) { S.alphaPrevS[i] = 0.f; S.dmixS[i] = 0.f; }
  for (int i = tid; i < 2 * TtP; i += NTB) { S.deS[i] = 0.f; S.dwst[i] = 0.f; }
  for (int i = tid; i < RINGB * 3 * TtP; i += NTB) S.ringA[i] = 0.f;
  for (int i = tid; i < RINGB * RB; i += NTB) S.ringB[i] = 0.f;
  for (int i = tid; i < RINGB * 2 * NB * UH; i += NTB) S.mk_ring[i] = 0;
  for (int i = tid; i < 16 * VC; i += NTB) S.inC[i] = 0.f;
  for (int i = tid; i < 2 * 16 * NB * UH; i += NTB) S.inH[i] = 0.f;
  for (int i = tid; i < CW * NSLOT; i += NTB) S.dqst[i] = 0.f;
  for (int i = tid; i < NB * QT; i += NTB) S.dqB[i] = 0.f;
  if (tid < VC) S.dctxS[tid] = 0.f;
  if (tid < QC) S.dqS[tid] = 0.f;
  if (tid < NB * UH) S.dhS[tid] = 0.f;
  for (int i = tid; i < NB * DGS; i += NTB) S.dgS[i] = 0.f;
  if (tid < 16) S.pt[tid] = 0u;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) cl::mbar_init(&S.bars[i], 1);
    cl::fence_mbar_init();
  }

  // ---------------- BC role: thread = (quad of rows rq = tid >> 2 of the recurrent kernel (rows < 512), quarter cq = tid & 3 of my
  // 64 gate columns): 4 x 16 weights in registers.  (A thread per row with all 64 columns would need 80 broadcast LDS.128 of
  // d(gates) per step, and a broadcast 16-byte load still costs 4 shared-memory cycles: 5 K cycles per step for 16 warps.)
  const int bc_rq = tid >> 2, bc_cq = tid & 3;
  __shared__ uint32_t tmem_base_sh;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cl::smem_u32(&tmem_base_sh)), "n"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm_addr = tmem_base_sh + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  {
    uint32_t wbits[64];       // [row of my quad][column of my quarter]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int col = 16 * bc_cq + c;                  // CTA-local gate column = gate*16 + unit
        wbits[r * 16 + c] = __float_as_uint(__ldg(d.Wrec + (long long)(4 * bc_rq + r) * (4 * H) + (col >> 4) * H + rank * UH + (col & 15)));
      }
    tmem_st32(tm_addr, wbits, 0);
    tmem_st32(tm_addr + 32u, wbits, 32);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  // destination of my row quad's partial results: owner of context columns 4rq..4rq+3 / hidden units 4rq-288..
  const int bc_k0 = 4 * bc_rq;
  int bc_g, bc_off;          // group member (context rows) or -1, float offset inside the destination's inbox
  if (bc_k0 < M1) { bc_g = bc_rq / VAq; bc_off = rank * VC + (bc_k0 - 4 * bc_g * VAq); }
  else if (bc_k0 < M1 + M2) { const int c2 = bc_k0 - M1; bc_g = (c2 >> 2) / VBq; bc_off = rank * VC + VA + (c2 - 4 * bc_g * VBq); }
  else { bc_g = -1; bc_off = 0; }
  const int bc_unit = bc_k0 - (M1 + M2);

  // ---------------- BB role: tid < NB*64: (utterance, unit half, channel quarter, unit) ; the quarter-0 lanes own the cell state
  const int bb_u = tid >> 6, bb_unit = (tid & 7) + 8 * ((tid >> 5) & 1), bb_part = (tid >> 3) & 3;
  const bool bb_act = tid < NB * 64;
  const bool bb_own = bb_act && bb_part == 0;
  const bool bb_ok = bb_own && (b0 + bb_u) < B;
  float dc_st = 0.f, dh_st = 0.f;

  // ---------------- energy / d(weights) role
  const int ep = lane >> 3, ecl = lane & 7;
  const int nact = (warp < CW) ? min(NP, max(0, (alen - 4 * warp + PSL - 1) / PSL)) : 0;
  int echunks = 0;
  for (int w_ = 0; w_ < CW; ++w_) echunks += min(NP, max(0, (alen - 4 * w_ + PSL - 1) / PSL));
  const uint32_t RX_W = (uint32_t)G * 2u * 16u * (uint32_t)echunks;
  const uint32_t RX_WS = (uint32_t)G * 16u * (uint32_t)Tt4;               // partial d(previous alignments) of the group
  const uint32_t RX_Q = (uint32_t)NB * QT * 4u;
  const uint32_t RX_C = 16u * (uint32_t)((grp ? 4 * (sl.van + sl.vbn) : 0) + NB * UH) * 4u;
  // service role (warps CW..16)
  const int sv = tid - CW * 32;
  const bool service = sv >= 0;

  __syncthreads();
  cluster.sync();

  // prefetch of the global inputs of time tau into ring slot tau % RINGB (always commits a group)
  auto prefetch = [&](int tau) {
    if (service && tau >= 0) {
      const int slot = tau % RINGB;
      if (arow_ok) {
        float* ra = S.ringA + (size_t)slot * 3 * TtP;
        const long long oa = ((long long)tau * B + arow) * Tt;
        if ((Tt & 3) == 0) {
          for (int e = sv; e < 3 * Tt4; e += NSVB) {
            const int arr = e / Tt4, q = e % Tt4;
            const float* src = (arr == 0) ? d.soft1 : (arr == 1) ? d.align1 : d.align2;
            cp_async16(ra + arr * TtP + 4 * q, src + oa + 4 * q);
          }
        } else {
          for (int e = sv; e < 3 * Tt; e += NSVB) {
            const int arr = e / Tt, j = e % Tt;
            const float* src = (arr == 0) ? d.soft1 : (arr == 1) ? d.align1 : d.align2;
            cp_async4(ra + arr * TtP + j, src + oa + j);
          }
        }
        float* rb = S.ringB + (size_t)slot * RB;
        const long long rowq = ((long long)tau * B + arow) * QT;
        const long long rowx = ((long long)tau * B + arow) * X2W + H;
        for (int e = sv; e < sl.qan + sl.qbn + sl.van + sl.vbn; e += NSVB) {
          int q = e;
          if (q < sl.qan) { cp_async16(rb + 4 * q, d.q_save + rowq + 4 * (sl.qa0 + q)); continue; }
          q -= sl.qan;
          if (q < sl.qbn) { cp_async16(rb + QA + 4 * q, d.q_save + rowq + A1 + 4 * (sl.qb0 + q)); continue; }
          q -= sl.qbn;
          if (q < sl.van) { cp_async16(rb + QC + 4 * q, dd.dx2 + rowx + 4 * (sl.va0 + q)); continue; }
          q -= sl.van;
          cp_async16(rb + QC + VA + 4 * q, dd.dx2 + rowx + M1 + 4 * (sl.vb0 + q));
        }
      }
      {
        float* rp = S.ringB + (size_t)slot * RB + QC + VC;
        for (int e = sv; e < 6 * NB * 4; e += NSVB) {
          const int arr = e / (NB * 4), u = (e >> 2) % NB, q4 = e & 3;
          if (b0 + u >= B) continue;
          const long long rb_ = (long long)tau * B + b0 + u;
          const float* src = (arr < 4) ? d.gates + rb_ * (4 * H) + arr * H + rank * UH + 4 * q4
                           : (arr == 4) ? d.c_prev + rb_ * H + rank * UH + 4 * q4
                                        : dd.dx2 + rb_ * X2W + rank * UH + 4 * q4;
          cp_async16(rp + (arr * NB + u) * UH + 4 * q4, src);
        }
        if (sv < 2 * NB) {
          const int which = sv / NB, u = sv % NB;
          const uint8_t* src = which ? d.mask_h : d.mask_c;
          if (src && b0 + u < B) cp_async16(S.mk_ring + ((slot * 2 + which) * NB + u) * UH, src + ((long long)tau * B + b0 + u) * H + rank * UH);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int i = 0; i <= PFDB; ++i) prefetch(Te - 1 - i);
  // rows of the skipped steps: zero gradients (round-robin over the CTAs of the cluster; plain stores, long before the first exchange)
  for (int r = rank; r < (Td - Te) * NB; r += CS) {
    const int tz = Te + r / NB, bz = b0 + r % NB;
    if (bz < B) {
      float* gz = dd.dgates + ((long long)tz * B + bz) * (4 * H);
      for (int i = tid; i < 4 * H; i += NTB) gz[i] = 0.f;
      if (dd.dq) {
        float* qz = dd.dq + ((long long)tz * B + bz) * QT;
        for (int i = tid; i < QT; i += NTB) qz[i] = 0.f;
      }
      float* ez = de_out + ((long long)tz * B + bz) * 2 * DEL;
      for (int i = tid; i < 2 * DEL; i += NTB) ez[i] = 0.f;
    }
  }

  // staging of the chain-independent inputs of step ts (compute warps): a_{ts-1} (location input), alpha_{ts-1}, the processed
  // query of step ts in the channel packs; `location_features` follows after a barrier of the compute warps
  auto stage_inputs = [&](int ts) {
    const float* rAp = S.ringA + (size_t)((ts + RINGB - 1) % RINGB) * 3 * TtP;   // time ts-1
    const float* rBs = S.ringB + (size_t)(ts % RINGB) * RB;
    if (tid < TtP) {
      const int j = tid;
      const bool in = j < Tt && ts > 0;
      S.aprev[HALO + j] = in ? rAp[j] : 0.f;
      S.alphaPrevS[j] = in ? rAp[TtP + j] : ((j == 0 && d.mode == 2) ? 1.f : 0.f);
    } else if (grp && tid < TtP + QC) {
      const int c = tid - TtP;
      const int slot = (c < QA) ? c : 8 * NIA + (c - QA);
      reinterpret_cast<float*>(&S.packA[slot])[0] = rBs[c] * K2LOG2E;
    }
  };
  cp_async_wait<1>();
  __syncthreads();
  if (warp < CW) {
    stage_inputs(Te - 1);
    cl::named_bar_sync(5, CW * 32);
    if (grp) arnn::location_features<AFT>(S.fS, S.aprev, S.wconv, S.bconv, alen, d.att_kernel, pl, tid, CW * 32);
  }

  PT2_DECL
#pragma unroll 1
  for (int t = Te - 1; t >= 0; --t) {
    const int u = Te - 1 - t, cur = u & 1, nxt = cur ^ 1;
    const uint32_t par = (uint32_t)(u >> 1) & 1u;
    PT2(15)
    prefetch(t - 1 - PFDB);
    cp_async_wait<1>();                          // times t, t-1 and t-2 are resident (t-2 feeds the early staging of step t-1)
    __syncthreads();                             // #0: ring contents (written by other threads' cp.async) visible
    const float* rA = S.ringA + (size_t)(t % RINGB) * 3 * TtP;                  // soft1[t], align1[t], align2[t]
    const float* rB = S.ringB + (size_t)(t % RINGB) * RB;
    // (a_{t-1}, alpha_{t-1}, the query pack and the location features of this step were staged during step t+1 / the prologue)
    if (u > 0) cl::mbar_wait(&barC[cur], (uint32_t)((u - 1) >> 1) & 1u);   // partial products of step t+1
    if (tid == 0) {
      if (grp) cl::mbar_arrive_expect_tx(&barW[cur], RX_W + (u > 0 ? RX_WS : 0u));
      cl::mbar_arrive_expect_tx(&barQ[cur], RX_Q);
      if (t > 0) cl::mbar_arrive_expect_tx(&barC[nxt], RX_C);
    }
    PT2(0)
    // ======================= BA0: sum the 16 partial products
    if (grp && tid < VC) {
      const bool a1 = tid < VA;
      const bool real = a1 ? (tid < 4 * sl.van) : (tid - VA < 4 * sl.vbn);
      float v = 0.f;
      if (real) {
        v = rB[QC + tid];
        if (u > 0) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int r = 0; r < 16; r += 2) { s0 += S.inC[r * VC + tid]; s1 += S.inC[(r + 1) * VC + tid]; }
          v += s0 + s1;
        }
      }
      S.dctxS[tid] = v;                          // total d(ctx); saved for the dense dvalues GEMM by the service warps
    } else if (tid >= 128 && tid < 128 + NB * UH) {
      const int e = tid - 128;
      float v = 0.f;
      if (u > 0) {
        const float* ih = S.inH + cur * 16 * NB * UH + e;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int r = 0; r < 16; r += 2) { s0 += ih[r * NB * UH]; s1 += ih[(r + 1) * NB * UH]; }
        v = s0 + s1;
      }
      S.dhS[e] = v;
    }
    __syncthreads();                             // #2
    PT2(1)
    if (grp) {
      // ======================= BA1: partial d(weights) over my value columns -> the group
      if (warp < CW) {
        const int slot0 = warp * 4 + ep;
        switch (nact) {
          case 1: dweights_passes<NB, 1>(S, slot0, ecl, Tt); break;
          case 2: dweights_passes<NB, 2>(S, slot0, ecl, Tt); break;
          case 3: dweights_passes<NB, 3>(S, slot0, ecl, Tt); break;
          case 4: if (NP >= 4) dweights_passes<NB, (NP >= 4 ? 4 : 1)>(S, slot0, ecl, Tt); break;
          default: break;
        }
        __syncwarp();
        if (lane < nact * 2 * G) {
          const int m = lane / (2 * G), rem = lane % (2 * G), att = rem / G, gd = rem % G;
          const int j0 = warp * 4 + PSL * m;
          const float4 e4 = *reinterpret_cast<const float4*>(&S.dwst[att * TtP + j0]);
          const int dst = au * G + gd;
          st_async_v4(cl::mapa(cl::smem_u32(&S.dwpart[(att * G + ag) * TtP + j0]), dst), e4.x, e4.y, e4.z, e4.w,
                      cl::mapa(cl::smem_u32(&barW[cur]), dst));
        }
      } else if (arow_ok) {
        // service: total d(ctx) of step t -> dx2[:, H:]
        for (int q = sv; q < VCq; q += NSVB) {
          const bool a1 = q < VAq;
          const bool real = a1 ? (q < sl.van) : (q - VAq < sl.vbn);
          if (!real) continue;
          const int k = a1 ? 4 * (sl.va0 + q) : M1 + 4 * (sl.vb0 + q - VAq);
          *reinterpret_cast<float4*>(dd.dx2 + ((long long)t * B + arow) * X2W + H + k) = *reinterpret_cast<const float4*>(&S.dctxS[4 * q]);
        }
      }
      PT2(2)
      if (warp < 2 * SMW) cl::mbar_wait(&barW[cur], par);
      PT2(3)
      // ======================= BA2a: recursion / softmax backward (one position per thread)
      if (warp < SMW) {
        const int j = tid;
        const bool in = j < TtP;
        const float a = in ? rA[j] : 0.f;
        float dw = 0.f, dst = 0.f;
        if (in) {
#pragma unroll
          for (int g = 0; g < G; ++g) dw += S.dwpart[(0 * G + g) * TtP + j];
          if (u > 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) dst += S.dstate_part[(cur * G + g) * TtP + j];
          }
        }
        float da;
        if (d.mode == 2) {
          const float dal = in ? dw + S.dalpha_carry[j] : 0.f;
          const float apm1 = (in && j > 0) ? S.alphaPrevS[j - 1] : 0.f;
          const float mix = in ? (0.5f * S.alphaPrevS[j] + 0.5f * apm1 + 1e-7f) : 0.f;
          float s1 = mix * a, s2 = in ? dal * rA[TtP + j] : 0.f;
          gsum2<SMW>(s1, s2, S.red, warp, lane, 2);
          const float dau = (in && j < alen && s1 > 0.f) ? (dal - s2) * __fdividef(1.f, s1) : 0.f;
          da = dau * mix + dst;
          if (in) S.dmixS[j] = dau * a;
        } else {
          da = dw + dst;
        }
        float dot2 = da * a, dummy = 0.f;
        gsum2<SMW>(dot2, dummy, S.red, warp, lane, 2);   // the barrier inside also orders the dmixS writes above
        if (in) {
          S.deS[0 * TtP + j] = a * (da - dot2);
          if (d.mode == 2) {
            const float nx = (j + 1 < TtP) ? S.dmixS[j + 1] : 0.f;
            S.dalpha_carry[j] = 0.5f * S.dmixS[j] + 0.5f * nx;      // adjoint of the shift (forward_attention.py:108-109)
          }
        }
      } else if (warp < 2 * SMW) {
        const int j = tid - SMW * 32;
        const bool in = j < TtP;
        const float a = in ? rA[2 * TtP + j] : 0.f;
        float dw = 0.f;
        if (in) {
#pragma unroll
          for (int g = 0; g < G; ++g) dw += S.dwpart[(1 * G + g) * TtP + j];
        }
        float dot = dw * a, dummy = 0.f;
        gsum2<SMW>(dot, dummy, S.red + 32, warp - SMW, lane, 3);
        if (in) S.deS[1 * TtP + j] = a * (dw - dot);
      }
    }
    PT2(4)
    __syncthreads();                             // #3
    PT2(5)
    if (grp) {
      // ======================= BA2b: energy backward over my channel slice
      if (warp < CW) {
        const int slot0 = warp * 4 + ep;
        switch (nact) {
          case 1: energy_bwd_passes<NB, 1>(S, slot0, ecl, lane, warp, Tt); break;
          case 2: energy_bwd_passes<NB, 2>(S, slot0, ecl, lane, warp, Tt); break;
          case 3: energy_bwd_passes<NB, 3>(S, slot0, ecl, lane, warp, Tt); break;
          case 4: if (NP >= 4) energy_bwd_passes<NB, (NP >= 4 ? 4 : 1)>(S, slot0, ecl, lane, warp, Tt); break;
          default: break;                        // dqst / dfS of an idle warp keep their initial zeros
        }
      } else if (arow_ok && ag < 2) {
        // service: d(energies) of step t -> global (input of satk_attn_energy_grad): member 0 mechanism 1, member 1 mechanism 2
        if (prog && (t + 1) % EG_PUB == 0 && t + 1 < Te) {
          // the rows of the steps >= t + 1 left a whole step ago: make them visible, then publish the progress
          __threadfence();
          cl::named_bar_sync(7, NSVB);
          if (sv == 0) *reinterpret_cast<volatile int*>(prog + arow * 2 + ag) = t + 1;
        }
        float* dst = de_out + (((long long)t * B + arow) * 2 + ag) * DEL;
        const float* src = S.deS + ag * TtP;
        if ((Tt & 3) == 0) {
          for (int q = sv; q < Tt4; q += NSVB) *reinterpret_cast<float4*>(dst + 4 * q) = *reinterpret_cast<const float4*>(src + 4 * q);
        } else {
          for (int j = sv; j < Tt; j += NSVB) dst[j] = src[j];
        }
      }
    }
    PT2(6)
    __syncthreads();                             // #4
    PT2(7)
    if (grp) {
      if (warp < 3) {
        // ======================= BA3: d(query) slice = sum of the per-warp partials -> every CTA
        {
          const int slot = tid;
          float q = 0.f;
          if (slot < NSLOT) {
#pragma unroll
            for (int w_ = 0; w_ < CW; ++w_) q += S.dqst[w_ * NSLOT + slot];
            if (slot < 8 * NIA) { if (slot < QA) S.dqS[slot] = q; }
            else if (slot - 8 * NIA < GE::QB) S.dqS[QA + slot - 8 * NIA] = q;
          }
        }
        cl::named_bar_sync(4, 96);
        const int nq = sl.qan + sl.qbn;
        for (int e = tid; e < nq * CS; e += 96) {
          const int q = e % nq, dst = e / nq;
          const bool a1 = q < sl.qan;
          const float4 q4 = *reinterpret_cast<const float4*>(&S.dqS[a1 ? 4 * q : QA + 4 * (q - sl.qan)]);
          const int gcol = a1 ? 4 * (sl.qa0 + q) : A1 + 4 * (sl.qb0 + q - sl.qan);
          st_async_v4(cl::mapa(cl::smem_u32(&S.dqB[au * QT + gcol]), dst), q4.x, q4.y, q4.z, q4.w, cl::mapa(cl::smem_u32(&barQ[cur]), dst));
        }
      } else if (service && arow_ok && dd.dq) {
        // service: d(query) slice of step t -> global (dense dWq afterwards): sum of the per-warp partials
        for (int q = sv; q < sl.qan + sl.qbn; q += NSVB) {
          const bool a1 = q < sl.qan;
          const int s0 = a1 ? 4 * q : 8 * NIA + 4 * (q - sl.qan);
          float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w_ = 0; w_ < CW; ++w_) {
            const float4 p4 = *reinterpret_cast<const float4*>(&S.dqst[w_ * NSLOT + s0]);
            v4.x += p4.x; v4.y += p4.y; v4.z += p4.z; v4.w += p4.w;
          }
          const int gcol = a1 ? 4 * (sl.qa0 + q) : A1 + 4 * (sl.qb0 + q - sl.qan);
          *reinterpret_cast<float4*>(dd.dq + ((long long)t * B + arow) * QT + gcol) = v4;
        }
      }
    }
    PT2(8)
    cl::mbar_wait(&barQ[cur], par);
    PT2(9)
    // ======================= BB: d(out1) of my units, LSTM cell backward
    if (warp < (NB * 64 + 31) / 32) {
      float acc = 0.f;
      if (bb_act) {
        const float* qrow = S.dqB + bb_u * QT + 64 * bb_part;
        const float* wrow = S.WqU + bb_unit * WQS + 64 * bb_part;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float4 q0 = *reinterpret_cast<const float4*>(qrow + 4 * i), w0 = *reinterpret_cast<const float4*>(wrow + 4 * i);
          const float4 q1 = *reinterpret_cast<const float4*>(qrow + 4 * i + 4), w1 = *reinterpret_cast<const float4*>(wrow + 4 * i + 4);
          a0 = fmaf(q0.x, w0.x, a0); a0 = fmaf(q0.y, w0.y, a0); a0 = fmaf(q0.z, w0.z, a0); a0 = fmaf(q0.w, w0.w, a0);
          a1 = fmaf(q1.x, w1.x, a1); a1 = fmaf(q1.y, w1.y, a1); a1 = fmaf(q1.z, w1.z, a1); a1 = fmaf(q1.w, w1.w, a1);
        }
        acc = a0 + a1;
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      if (bb_own) {
        float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
        const int e = bb_u * UH + bb_unit;
        if (bb_ok) {
          dh_st += S.dhS[e];                     // recurrent carry: sum of the partial products of step t+1
          const float* rp = rB + QC + VC + e;
          const float gi = rp[0 * NB * UH], gj = rp[1 * NB * UH], gf = rp[2 * NB * UH], go = rp[3 * NB * UH];
          const float cp = rp[4 * NB * UH], dx2o = rp[5 * NB * UH];
          const uint8_t* mr = S.mk_ring + (t % RINGB) * 2 * NB * UH;
          const float mc = d.mask_c ? (float)mr[0 * NB * UH + e] : (1.f - d.zc);
          const float mh = d.mask_h ? (float)mr[1 * NB * UH + e] : (1.f - d.zh);
          const float c_new = gf * cp + gi * gj;
          const float tc = ftanh(c_new);
          const float dh_new = (dx2o + acc) + mh * dh_st;
          dh_st = (1.f - mh) * dh_st;
          const float dcn = mc * dc_st + dh_new * go * (1.f - tc * tc);
          dgo = dh_new * tc * go * (1.f - go);
          dgi = dcn * gj * gi * (1.f - gi);
          dgj = dcn * gi * (1.f - gj * gj);
          dgf = dcn * cp * gf * (1.f - gf);
          dc_st = (1.f - mc) * dc_st + dcn * gf;
        }
        float* dg = S.dgS + bb_u * DGS + bb_unit;
        dg[0] = dgi; dg[DGG] = dgj; dg[2 * DGG] = dgf; dg[3 * DGG] = dgo;
      }
    }
    PT2(10)
    __syncthreads();                             // #5
    PT2(11)
    if (t > 0) {
      // ======================= BC: partial d([ctx | h])(t-1) over my gate columns -> the consumers
      uint32_t wbits[64];
      tmem_ld64(tm_addr, wbits);
      float acc[NB][4];      // [utterance][row of my quad], partial over my 16 columns
#pragma unroll
      for (int uu = 0; uu < NB; ++uu) {
        acc[uu][0] = 0.f; acc[uu][1] = 0.f; acc[uu][2] = 0.f; acc[uu][3] = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 g4 = *reinterpret_cast<const float4*>(&S.dgS[uu * DGS + DGG * bc_cq + 4 * c4]);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[uu][r] = fmaf(__uint_as_float(wbits[r * 16 + 4 * c4]), g4.x, acc[uu][r]);
            acc[uu][r] = fmaf(__uint_as_float(wbits[r * 16 + 4 * c4 + 1]), g4.y, acc[uu][r]);
            acc[uu][r] = fmaf(__uint_as_float(wbits[r * 16 + 4 * c4 + 2]), g4.z, acc[uu][r]);
            acc[uu][r] = fmaf(__uint_as_float(wbits[r * 16 + 4 * c4 + 3]), g4.w, acc[uu][r]);
          }
        }
      }
      // reduce-scatter over the 4 column quarters (lanes cq): lane q ends with the 4 rows of utterance q, i.e. with exactly the
      // 16-byte packet it sends; the fifth utterance is reduced row-wise and gathered by lane 0 of the quad
      float o4[4];
      {
        const bool up2 = (bc_cq & 2) != 0, up1 = (bc_cq & 1) != 0;
        float k[2][4];       // utterances {0,1} (lanes 0,1) or {2,3} (lanes 2,3)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float s0 = up2 ? acc[0][r] : acc[2][r], s1 = up2 ? acc[1][r] : acc[3][r];
          k[0][r] = (up2 ? acc[2][r] : acc[0][r]) + __shfl_xor_sync(0xffffffffu, s0, 2);
          k[1][r] = (up2 ? acc[3][r] : acc[1][r]) + __shfl_xor_sync(0xffffffffu, s1, 2);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float sn = up1 ? k[0][r] : k[1][r];
          o4[r] = (up1 ? k[1][r] : k[0][r]) + __shfl_xor_sync(0xffffffffu, sn, 1);
        }
      }
      float e4[4] = {0.f, 0.f, 0.f, 0.f};
      if (NB == 5) {
        const bool up2 = (bc_cq & 2) != 0, up1 = (bc_cq & 1) != 0;
        float k0 = up2 ? acc[NB - 1][2] : acc[NB - 1][0], k1 = up2 ? acc[NB - 1][3] : acc[NB - 1][1];
        k0 += __shfl_xor_sync(0xffffffffu, up2 ? acc[NB - 1][0] : acc[NB - 1][2], 2);
        k1 += __shfl_xor_sync(0xffffffffu, up2 ? acc[NB - 1][1] : acc[NB - 1][3], 2);
        float kk = up1 ? k1 : k0;                      // lane q: row q of the fifth utterance
        kk += __shfl_xor_sync(0xffffffffu, up1 ? k0 : k1, 1);
        const int l4 = lane & ~3;
        e4[0] = __shfl_sync(0xffffffffu, kk, l4); e4[1] = __shfl_sync(0xffffffffu, kk, l4 + 1);
        e4[2] = __shfl_sync(0xffffffffu, kk, l4 + 2); e4[3] = __shfl_sync(0xffffffffu, kk, l4 + 3);
      }
      const uint32_t bara = cl::smem_u32(&barC[nxt]);
      if (bc_g >= 0) {
        const uint32_t dsta = cl::smem_u32(&S.inC[bc_off]);
        if (bc_cq < NB) st_async_v4(cl::mapa(dsta, bc_cq * G + bc_g), o4[0], o4[1], o4[2], o4[3], cl::mapa(bara, bc_cq * G + bc_g));
        if (NB == 5 && bc_cq == 0)
          st_async_v4(cl::mapa(dsta, (NB - 1) * G + bc_g), e4[0], e4[1], e4[2], e4[3], cl::mapa(bara, (NB - 1) * G + bc_g));
      } else {
        const int dst = bc_unit >> 4;
        const uint32_t dsta = cl::mapa(cl::smem_u32(&S.inH[(nxt * 16 + rank) * NB * UH + (bc_unit & 15)]), dst);
        const uint32_t barr = cl::mapa(bara, dst);
        if (bc_cq < NB) st_async_v4(dsta + bc_cq * UH * 4, o4[0], o4[1], o4[2], o4[3], barr);
        if (NB == 5 && bc_cq == 0) st_async_v4(dsta + (NB - 1) * UH * 4, e4[0], e4[1], e4[2], e4[3], barr);
      }
      PT2(14)
      // rows 512..543 (hidden units 224..255): thread = (utterance, row, half of the column quads)
      if (tid < NB * 64) {
        const int o = tid >> 1, hh = tid & 1, r = o & 31, uu = o >> 5;
        const float* wrow = S.WragS + r * WRS + 4 * hh;
        const float* grow = S.dgS + uu * DGS + 4 * hh;     // column quad 2i + hh = gate (i >> 1), quad ((2i + hh) & 3) of the gate
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float4 w0 = *reinterpret_cast<const float4*>(wrow + 8 * i), g0 = *reinterpret_cast<const float4*>(grow + (i >> 1) * DGG + 8 * (i & 1));
          const float4 w1 = *reinterpret_cast<const float4*>(wrow + 8 * i + 8),
                       g1 = *reinterpret_cast<const float4*>(grow + ((i + 1) >> 1) * DGG + 8 * ((i + 1) & 1));
          a0 = fmaf(w0.x, g0.x, a0); a0 = fmaf(w0.y, g0.y, a0); a0 = fmaf(w0.z, g0.z, a0); a0 = fmaf(w0.w, g0.w, a0);
          a1 = fmaf(w1.x, g1.x, a1); a1 = fmaf(w1.y, g1.y, a1); a1 = fmaf(w1.z, g1.z, a1); a1 = fmaf(w1.w, g1.w, a1);
        }
        float v = a0 + a1;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (hh == 0) {
          const int unit = NTB - (M1 + M2) + r;          // 224 + r
          const int dst = unit >> 4;
          st_async_f32(cl::mapa(cl::smem_u32(&S.inH[(nxt * 16 + rank) * NB * UH + uu * UH + (unit & 15)]), dst), v, cl::mapa(bara, dst));
        }
      }
    }
    PT2(12)
    // ---- off the dependency chain, in the shadow of the exchange the next step waits for:
    // partial d(a_{t-1}) for the group (consumed by BA2a of step t-1), then the staging / location features of step t-1
    // (nothing of the current step reads those buffers after barrier #4)
    if (grp && warp < CW) {
        // partial d(a_{t-1}) = conv-transpose of my partial d(location features) -> the group (consumed by step t-1)
        if (t > 0) {
          for (int base = warp * 32; base < 4 * Tt4; base += CW * 32) {
            const int j = base + lane;
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
            const int l4 = lane & ~3;
            const float a0 = __shfl_sync(0xffffffffu, acc, l4), a1 = __shfl_sync(0xffffffffu, acc, l4 + 1);
            const float a2 = __shfl_sync(0xffffffffu, acc, l4 + 2), a3 = __shfl_sync(0xffffffffu, acc, l4 + 3);
            if ((lane & 3) == 0 && j < 4 * Tt4) {
              const uint32_t dsta = cl::smem_u32(&S.dstate_part[(nxt * G + ag) * TtP + j]), bara = cl::smem_u32(&barW[nxt]);
#pragma unroll
              for (int gd = 0; gd < G; ++gd) st_async_v4(cl::mapa(dsta, au * G + gd), a0, a1, a2, a3, cl::mapa(bara, au * G + gd));
            }
          }
        }
    }
    if (warp < CW && t > 0) {
      stage_inputs(t - 1);
      cl::named_bar_sync(5, CW * 32);
      if (grp) arnn::location_features<AFT>(S.fS, S.aprev, S.wconv, S.bconv, alen, d.att_kernel, pl, tid, CW * 32);
    }
    PT2(13)
    if (service) {
      // service: d(gates) of step t -> global (after this warp's DSMEM stores of the step)
      for (int e = sv; e < NB * 16; e += NSVB) {
        const int uu = e >> 4, g4 = (e >> 2) & 3, q4 = e & 3;
        if (b0 + uu < B)
          *reinterpret_cast<float4*>(dd.dgates + ((long long)t * B + b0 + uu) * (4 * H) + g4 * H + rank * UH + 4 * q4) =
              *reinterpret_cast<const float4*>(&S.dgS[uu * DGS + g4 * DGG + 4 * q4]);
      }
    }
  }
  PT2_FLUSH(Te)
  if (prog && service && arow_ok && ag < 2) {      // every row of my utterance / mechanism is out
    __threadfence();
    cl::named_bar_sync(7, NSVB);
    if (sv == 0) *reinterpret_cast<volatile int*>(prog + arow * 2 + ag) = 0;
  }
  cp_async_wait<0>();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_sh), "n"(TM_COLS) : "memory");
  cluster.sync();
}

template <int NB>
static size_t bwd2_smem_bytes(int np) {
  BwdSmem2<NB> S;
  return S.carve(nullptr, np);
}

template <typename Kern>
static int launch16b(Kern kern, const satk_attn_rnn_bwd_desc& d, float* de, int* prog, int nb, size_t smem, cudaStream_t st) {
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((d.f.B + nb - 1) / nb) * CS);
  cfg.blockDim = dim3(NTB);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SATK_CUDA(cudaLaunchKernelEx(&cfg, kern, d, de, prog));
  return SATK_OK;
}

size_t attn_rnn2_bwd_smem(int nb, int Tt) {
  const int np = (Tt + PSL - 1) / PSL;
  const int npv = np <= 3 ? 3 : 4;
  return nb == 5 ? bwd2_smem_bytes<5>(npv) : bwd2_smem_bytes<4>(npv);
}

int v2_pick_nb_bwd(const satk_attn_rnn_fwd_desc* d) {
  const char* e = getenv("SATK_ATTN_NB");
  if (e && (e[0] == '4' || e[0] == '5')) return e[0] - '0';
  if (d->B <= 28) return 4;
  return (attn_rnn2_bwd_smem(5, d->Tt) <= 227 * 1024) ? 5 : 4;
}

// `de` [Td,B,2,Tt]: d(energies) of both mechanisms, consumed by satk_attn_energy_grad
int attn_rnn2_bwd_launch(const satk_attn_rnn_bwd_desc* d, float* de, int* prog, cudaStream_t st) {
  const int nb = v2_pick_nb_bwd(&d->f);
  const int np = (d->f.Tt + PSL - 1) / PSL;
  SATK_CHECK_ARG(np <= 4, "attn_rnn2_bwd: Tt=%d out of range", d->f.Tt);
  const int npv = np <= 3 ? 3 : 4;
  const size_t smem = attn_rnn2_bwd_smem(nb, d->f.Tt);
  SATK_CHECK_ARG(smem <= 227 * 1024, "attn_rnn2_bwd: Tt=%d needs %zu B of shared memory (> 227 KB)", d->f.Tt, smem);
  if (nb == 5) {
    if (npv == 3) return launch16b(attn_rnn2_bwd_kernel<5, 3>, *d, de, prog, nb, smem, st);
    return launch16b(attn_rnn2_bwd_kernel<5, 4>, *d, de, prog, nb, smem, st);
  }
  if (npv == 3) return launch16b(attn_rnn2_bwd_kernel<4, 3>, *d, de, prog, nb, smem, st);
  return launch16b(attn_rnn2_bwd_kernel<4, 4>, *d, de, prog, nb, smem, st);
}

int attn2_bwd_phase_cycles(long long* out16) {
#ifdef SATK_PHASE_TIMING
  SATK_CUDA(cudaMemcpyFromSymbol(out16, satk::g_phase, sizeof(long long) * 16));
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
#endif
  return 0;
}

}  // namespace arnn2
}  // namespace satk
