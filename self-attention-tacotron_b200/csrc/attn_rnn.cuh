// Shared layout of the attention-RNN cluster kernels (forward + backward).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace satk {
namespace arnn {

namespace cg = cooperative_groups;

constexpr int H = 256;      // LSTM-1 units
constexpr int M1 = 256;     // memory-1 depth
constexpr int CS = 16;      // CTAs per cluster
constexpr int BG = 4;       // utterances per cluster
constexpr int UH = 16;      // hidden units per CTA
constexpr int NT = 512;     // threads per CTA
constexpr int QC = 64;      // query/key columns handled per CTA (A1/4 [+ A2/4])
constexpr int KS = 72;      // padded row stride of keyS/valS (bank-conflict-free for 4 positions x 8 lanes)
constexpr int VC = 72;      // value columns per CTA: 64 of memory-1 + 8 of memory-2
constexpr int MAXF = 8;     // max location filters
constexpr int MAXK = 32;    // max location kernel taps
constexpr int HALO = 16;     // zero halo around per-position arrays (>= half the widest location kernel)

__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  x = fminf(fmaxf(x, -15.f), 15.f);
  float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

template <bool HAS2>
struct Dims {
  static constexpr int M2 = HAS2 ? 32 : 0;
  static constexpr int KREC = M1 + M2 + H;   // rows of the recurrent part of dec.lstm1.W
  static constexpr int KPT = KREC / 8;       // rows per thread in the gate GEMM
  static constexpr int A1Q = HAS2 ? 56 : 64; // attention-1 score columns per CTA
  static constexpr int NI1 = A1Q / 8;
  static constexpr int A2Q = HAS2 ? 8 : 0;
  static constexpr int X2W = H + M1 + M2;    // width of x2 rows
};

// round Tt up to a multiple of 32 for per-lane loops
__host__ __device__ inline int tt_pad(int Tt) { return (Tt + 31) / 32 * 32; }

// location features f[j][.] = conv1d(a_prev) + bias for all positions (forward_attention.py:98-100); fully unrolled for the
// shipped configuration (10 taps), generic otherwise
template <int AFT>
__device__ __forceinline__ void location_features(float* fS, const float* aprev, const float* wconv, const float* bconv, int Tt,
                                                  int att_kernel, int pl, int tid, int nthreads) {
  for (int idx = tid; idx < Tt * AFT; idx += nthreads) {
    const int j = idx / AFT, f = idx % AFT;
    float acc = bconv[f];
    const float* ap = aprev + HALO + j - pl;
    if (att_kernel == 10) {
#pragma unroll
      for (int k = 0; k < 10; ++k) acc = fmaf(ap[k], wconv[k * MAXF + f], acc);
    } else {
      for (int k = 0; k < att_kernel; ++k) acc = fmaf(ap[k], wconv[k * MAXF + f], acc);
    }
    fS[j * MAXF + f] = acc;
  }
}

}  // namespace arnn
}  // namespace satk
