// Geometry of the second-generation attention-RNN cluster kernels (attn_rnn2_fwd.cu, attn_rnn2_bwd.cu) for the dual-source
// decoder (forward / location-sensitive attention on memory 1 + additive attention on memory 2).
//
// A cluster of 16 CTAs owns NB utterances for all Td steps.  With NB = 5 the seven co-resident 16-CTA clusters of a B200 hold
// 35 utterances, so a batch of 32 runs in ONE wave (the first-generation kernels own 4 utterances per cluster: 7 + 1 waves).
//   * every CTA owns 16 hidden units of LSTM-1: its 64 gate columns of the recurrent kernel live in registers;
//   * utterance u of the cluster is served by a GROUP of G CTAs (ranks u*G .. u*G+G-1; G = 3 for NB = 5, 4 for NB = 4) that split
//     the score channels and the value columns in units of 4 (one 16-byte DSMEM store);
//   * the query projection is computed by the UNIT OWNERS as partial sums over their 16 units and sent to the groups
//     (q = sum over the 16 CTAs), so no CTA needs more than 16 rows of the query weights and the cell output itself is
//     never exchanged.
#pragma once
#include "attn_rnn.cuh"
#include "cluster_sync.cuh"

namespace satk {
namespace arnn2 {

namespace cg = cooperative_groups;
using arnn::fsigmoid;
using arnn::ftanh;

constexpr int H = 256, M1 = 256, M2 = 32, A1 = 224, A2 = 32;
constexpr int KREC = M1 + M2 + H;   // 544 recurrent rows of dec.lstm1.W: [ctx1 | ctx2 | h]
constexpr int QT = A1 + A2;         // 256 processed-query columns
constexpr int X2W = H + M1 + M2;    // 544
constexpr int CS = 16, UH = 16;
constexpr int MAXF = 8, MAXK = 32, HALO = 16;
constexpr int AFT = 5;              // location filters held per channel pack (att_filters <= 5)
constexpr int CW = 13;              // compute warps (the remaining warps only store to / prefetch from global memory)
constexpr int PSL = CW * 4;         // position slots per pass of the energy phases
constexpr int SMW = 6;              // warps per softmax group (one position per thread): Tt <= 192
constexpr float K2LOG2E = 2.885390081777927f;   // 2*log2(e): tanh(x) = 1 - 2 / (1 + 2^(K2LOG2E x))

// d(energies) workspace of the backward pass: rows [Td][B][2] of de_row_stride(Tt) floats.  Rows start on 128-byte lines so that a
// consumer running beside the recurrence never pulls a line into L1 that holds part of a row which is not written yet.
__host__ __device__ inline int de_row_stride(int Tt) { return (Tt + 31) & ~31; }
constexpr int EG_PUB = 8;          // the recurrence publishes its progress (lowest finished step) every EG_PUB steps

template <int NB>
struct Geo {
  static constexpr int G = (NB == 5) ? 3 : 4;
  static constexpr int QAq = (A1 / 4 + G - 1) / G;   // attention-1 channel quads per CTA (19 | 14)
  static constexpr int QBq = (A2 / 4 + G - 1) / G;   // attention-2 channel quads per CTA (3 | 2)
  static constexpr int QA = 4 * QAq, QB = 4 * QBq, QC = QA + QB;       // 76, 12, 88 | 56, 8, 64
  static constexpr int VAq = (M1 / 4 + G - 1) / G;   // memory-1 value quads per CTA (22 | 16)
  static constexpr int VBq = (M2 / 4 + G - 1) / G;   // (3 | 2)
  static constexpr int VA = 4 * VAq, VB = 4 * VBq, VC = VA + VB;       // 88, 12, 100 | 64, 8, 72
  static constexpr int VCq = VAq + VBq;
  static constexpr int NIA = (QA + 7) / 8, NIB = (QB + 7) / 8;         // channel iterations of the 8 channel lanes (10, 2 | 7, 1)
  static constexpr int NSLOT = 8 * (NIA + NIB);                        // channel slots incl. padding (96 | 64)
  static constexpr int KSTR = (QC % 32 == 8 || QC % 32 == 24) ? QC : QC + 8;   // key row stride: 4 rows x 8 lanes hit 32 banks
};

// real (un-padded) part of group member g's slices, in quads
template <int NB>
struct Slice {
  int qa0, qan, qb0, qbn, va0, van, vb0, vbn;
  __device__ __forceinline__ explicit Slice(int g) {
    using GE = Geo<NB>;
    qa0 = g * GE::QAq; qan = max(0, min(A1 / 4, qa0 + GE::QAq) - qa0);
    qb0 = g * GE::QBq; qbn = max(0, min(A2 / 4, qb0 + GE::QBq) - qb0);
    va0 = g * GE::VAq; van = max(0, min(M1 / 4, va0 + GE::VAq) - va0);
    vb0 = g * GE::VBq; vbn = max(0, min(M2 / 4, vb0 + GE::VBq) - vb0);
  }
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16-byte cp.async
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cl::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// Reductions over a group of NW warps through shared memory + a named barrier (all NW*32 threads call them); `red` >= 2*NW floats.
template <int NW>
__device__ __forceinline__ void gsum2(float& x, float& y, float* red, int wig, int lane, int barid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  if (lane == 0) { red[wig] = x; red[NW + wig] = y; }
  cl::named_bar_sync(barid, NW * 32);
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int i = 0; i < NW; ++i) { sx += red[i]; sy += red[NW + i]; }
  cl::named_bar_sync(barid, NW * 32);
  x = sx; y = sy;
}
template <int NW>
__device__ __forceinline__ float gmax(float x, float* red, int wig, int lane, int barid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  if (lane == 0) red[wig] = x;
  cl::named_bar_sync(barid, NW * 32);
  float m = red[0];
#pragma unroll
  for (int i = 1; i < NW; ++i) m = fmaxf(m, red[i]);
  cl::named_bar_sync(barid, NW * 32);
  return m;
}

// bound of |energy| above which the softmax falls back to subtracting the running maximum instead of the constant bound
constexpr float STAB_MAX = 30.f;

bool v2_eligible(const satk_attn_rnn_fwd_desc* d);
int v2_pick_nb(const satk_attn_rnn_fwd_desc* d);

}  // namespace arnn2
}  // namespace satk
