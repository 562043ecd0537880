// Gradients of the attention energies that do NOT feed the recurrence, for all (step, utterance) pairs in parallel:
//   e_t[j] = sum_c v_c tanh(keys[j,c] + q_t[c] + (f_t[j,:] . Wf[:,c]) + b_c),  f_t = conv1d(a_{t-1}) + bias
// (forward_attention.py:13-26,98-100; TF BahdanauAttention score, A.8).  The sequential backward kernel (attn_rnn2_bwd.cu)
// saves de_t[j] for both mechanisms; here the tanh terms are recomputed and
//   dkeys[j,c] = sum_t ds,  dv[c] += sum de tanh,  dWf[f,c] += sum f ds,  dconv[k,f] += sum a_{t-1}[j+k-pl] df[j,f],  dbias[f] += sum df
// with ds = de v (1 - tanh^2), df[j,f] = sum_c ds Wf[f,c].  Nothing here is on the step-to-step dependency chain, so the
// 400 x B x Tt x 256 element grid is spread over the whole chip: one CTA per (utterance, block of 32 score channels) walks
// the steps with its keys in shared memory and its dkeys tile in registers (written once, no atomics on dkeys).
#include "attn_rnn2.cuh"

namespace satk {
namespace arnn2 {

constexpr int EG_NT = 256;
constexpr int EG_CB = 16;          // score channels per CTA: lane = (position lane ep = lane >> 2, channel lane ecl = lane & 3), 4 channels per lane
constexpr int EG_KS = 20;          // key row stride (8 rows x 4 channel lanes -> 32 distinct banks)
constexpr int EG_MP = 3;           // position passes of 64 slots: Tt <= 192
constexpr int EG_TB = 4;           // decoder steps per iteration (amortises the barriers, hides the global loads)
constexpr int EG_JC = 8;           // positions per thread of the conv-gradient stage

struct EgSmem {
  int TtP, DFW;
  float *keyS, *fS, *dfS, *aprev, *deS, *qS;
  __host__ __device__ size_t carve(float* base, int Tt) {
    TtP = (Tt + 63) / 64 * 64;
    DFW = TtP + 2 * HALO;
    float* p = base;
    keyS = p; p += (size_t)TtP * EG_KS;
    fS = p; p += (size_t)2 * EG_TB * TtP * MAXF;   // [buffer][step][j][8] location features (precomputed, asynchronous copies)
    dfS = p; p += (size_t)EG_TB * AFT * DFW;
    aprev = p; p += 2 * EG_TB * DFW;           // [buffer][step][HALO + j]
    deS = p; p += 2 * EG_TB * TtP;
    qS = p; p += 2 * EG_TB * EG_CB;
    return (size_t)(p - base) * sizeof(float);
  }
};

// location features f_t[j][0..AFT) = conv1d(a_{t-1}) + bias for every (step, utterance): [Td][B][Tt][MAXF] (forward_attention.py:98-100)
__global__ void __launch_bounds__(256) loc_features_all_kernel(const satk_attn_rnn_fwd_desc d, float* __restrict__ fws) {
  __shared__ float ap[256 + 2 * HALO + MAXK];
  __shared__ float wc[MAXK * MAXF + MAXF];
  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, Tt = d.Tt;
  const int pl = (d.att_kernel - 1) / 2;
  for (int i = tid; i < MAXK * MAXF; i += 256) {
    const int k = i / MAXF, f = i % MAXF;
    wc[i] = (k < d.att_kernel && f < d.att_filters) ? __ldg(d.loc_conv_w + k * d.att_filters + f) : 0.f;
  }
  if (tid < MAXF) wc[MAXK * MAXF + tid] = (tid < d.att_filters) ? __ldg(d.loc_conv_b + tid) : 0.f;
  for (int i = tid; i < Tt + 2 * HALO + MAXK; i += 256) {
    const int j = i - HALO;
    ap[i] = (t > 0 && j >= 0 && j < Tt) ? __ldg(d.soft1 + ((long long)(t - 1) * d.B + b) * Tt + j) : 0.f;
  }
  __syncthreads();
  float* out = fws + ((long long)t * d.B + b) * Tt * MAXF;
  for (int idx = tid; idx < Tt * MAXF; idx += 256) {
    const int j = idx / MAXF, f = idx % MAXF;
    float acc = wc[MAXK * MAXF + f];
    for (int k = 0; k < d.att_kernel; ++k) acc = fmaf(ap[HALO + j - pl + k], wc[k * MAXF + f], acc);
    out[idx] = (f < AFT) ? acc : 0.f;
  }
}

template <bool LOC>
__device__ __forceinline__ void energy_grad_body(const satk_attn_rnn_bwd_desc& dd, const float* __restrict__ de,
                                                 const float* __restrict__ fws, const int cb /* channel block inside the mechanism */) {
  const satk_attn_rnn_fwd_desc& d = dd.f;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, Td = d.Td;
  const int alen = min((int)d.lengths[b], Tt);
  const int Te = dd.step_end ? max(1, min(Td, __ldg(dd.step_end + b))) : Td;
  const int pl = (d.att_kernel - 1) / 2;
  const int AW = LOC ? A1 : A2;                    // width of the mechanism's key rows
  const float* keys = LOC ? d.keys1 : d.keys2;
  const int qoff = (LOC ? 0 : A1) + cb * EG_CB;    // column of my channels inside the saved query rows

  extern __shared__ __align__(16) float smem_raw[];
  EgSmem S;
  S.carve(smem_raw, Tt);
  const int TtP = S.TtP, DFW = S.DFW;

  const int ep = lane >> 2, ecl = lane & 3;
  const int slot0 = warp * 8 + ep;                 // positions slot0 + 64 m
  const int npass = min(EG_MP, (alen - 8 * warp + 63) / 64);   // passes of this warp that touch the utterance (warp-uniform)

  for (int i = tid; i < TtP * EG_KS; i += EG_NT) {
    const int j = i / EG_KS, c = i % EG_KS;
    float kv = 0.f;
    if (j < Tt && c < EG_CB) kv = (__ldg(keys + ((long long)j * B + b) * AW + cb * EG_CB + c) + ((LOC && d.b1) ? __ldg(d.b1 + cb * EG_CB + c) : 0.f)) * K2LOG2E;
    S.keyS[i] = kv;
  }
  for (int i = tid; i < 2 * EG_TB * DFW; i += EG_NT) S.aprev[i] = 0.f;
  for (int i = tid; i < EG_TB * AFT * DFW; i += EG_NT) S.dfS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * TtP * MAXF; i += EG_NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * TtP; i += EG_NT) S.deS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * EG_CB; i += EG_NT) S.qS[i] = 0.f;

  // per-channel constants of my 4 channels (ecl + 4 i)
  float v4c[4], wf[4][AFT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cb * EG_CB + ecl + 4 * i;
    v4c[i] = 4.f * __ldg((LOC ? d.v1 : d.v2) + c);
#pragma unroll
    for (int f = 0; f < AFT; ++f) wf[i][f] = (LOC && f < d.att_filters) ? __ldg(d.loc_layer_w + (long long)f * A1 + c) * K2LOG2E : 0.f;
  }
  // accumulators: dkeys tile, sum de*r per channel (sum de*tanh = sum de - 2 sum de*r), d(location layer)
  float dk[EG_MP][4], dvr[4], dWf[4][AFT], sum_de = 0.f;
#pragma unroll
  for (int m = 0; m < EG_MP; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[m][i] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dvr[i] = 0.f;
#pragma unroll
    for (int f = 0; f < AFT; ++f) dWf[i][f] = 0.f;
  }
  // d(location convolution): thread = (filter f, chunk of EG_JC positions): the taps of the chunk accumulate in registers;
  // entry att_kernel of the accumulator row is the bias.  (generic tap count: the slower (tap, filter, quarter) mapping below)
  const bool conv10 = LOC && d.att_kernel == 10;
  const int nchunk = (Tt + EG_JC - 1) / EG_JC;
  const int cv_f = tid / nchunk, cv_c = tid % nchunk;
  const bool cv_act = conv10 && cv_f < AFT;
  float cacc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) cacc[k] = 0.f;
  const int cf_e = tid >> 2, cf_q = tid & 3;
  const int ntap = (LOC && !conv10) ? (d.att_kernel + 1) * AFT : 0;
  const int cf_k = cf_e / AFT, cf_f = cf_e % AFT;
  float dconv = 0.f;

  // asynchronous loads of de_t, a_{t-1}, f_t, q_t (my channels) of the EG_TB steps starting at t0 into buffer `buf`
  auto load_steps = [&](int t0, int buf) {
    for (int e = tid; e < EG_TB * Tt; e += EG_NT) {
      const int ts = e / Tt, j = e % Tt, t = t0 + ts;
      if (t < Te) {
        cl::cp_async4(&S.deS[(buf * EG_TB + ts) * TtP + j], de + (((long long)t * B + b) * 2 + (LOC ? 0 : 1)) * Tt + j);
        if (LOC) {
          if (t > 0) cl::cp_async4(&S.aprev[(buf * EG_TB + ts) * DFW + HALO + j], d.soft1 + ((long long)(t - 1) * B + b) * Tt + j);
          else S.aprev[(buf * EG_TB + ts) * DFW + HALO + j] = 0.f;
        }
      }
    }
    if (LOC) {
      for (int e = tid; e < EG_TB * Tt * 2; e += EG_NT) {          // two 16-byte halves per position
        const int ts = e / (2 * Tt), r = e % (2 * Tt), t = t0 + ts;
        if (t < Te) cp_async16(&S.fS[((size_t)(buf * EG_TB + ts) * TtP) * MAXF + 4 * r], fws + ((long long)t * B + b) * Tt * MAXF + 4 * r);
      }
    }
    if (tid < EG_TB * EG_CB) {
      const int ts = tid / EG_CB, c = tid % EG_CB, t = t0 + ts;
      if (t < Te) cl::cp_async4(&S.qS[(buf * EG_TB + ts) * EG_CB + c], d.q_save + ((long long)t * B + b) * QT + qoff + c);
    }
    cl::cp_async_commit();
  };

  __syncthreads();                         // the zero-fill above is ordered before the first asynchronous copies
  load_steps(0, 0);
#pragma unroll 1
  for (int t0 = 0, it = 0; t0 < Te; t0 += EG_TB, ++it) {
    const int buf = it & 1;
    const int nts = min(EG_TB, Te - t0);
    cl::cp_async_wait<0>();
    __syncthreads();                       // inputs of this iteration visible; everybody is done with the previous iteration
    load_steps(t0 + EG_TB, buf ^ 1);       // (commits an empty group past the end)
#pragma unroll 1
    for (int ts = 0; ts < nts; ++ts) {
      float q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = S.qS[(buf * EG_TB + ts) * EG_CB + ecl + 4 * i] * K2LOG2E;
      const float* des = S.deS + (buf * EG_TB + ts) * TtP;
      const float* fs = S.fS + (size_t)(buf * EG_TB + ts) * TtP * MAXF;
      float* dfs = S.dfS + (size_t)ts * AFT * DFW;
#pragma unroll
      for (int m = 0; m < EG_MP; ++m) {
        if (m < npass) {
          const int jr = slot0 + 64 * m, j = min(jr, Tt - 1);
          const float dej = (jr < Tt) ? des[j] : 0.f;
          sum_de += dej;
          float fv[AFT], dfp[AFT];
          if (LOC) {
            const float4 f4 = *reinterpret_cast<const float4*>(&fs[j * MAXF]);
            fv[0] = f4.x; fv[1] = f4.y; fv[2] = f4.z; fv[3] = f4.w; fv[4] = fs[j * MAXF + 4];
          }
#pragma unroll
          for (int f = 0; f < AFT; ++f) dfp[f] = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float s = S.keyS[j * EG_KS + ecl + 4 * i] + q[i];
            if (LOC) {
#pragma unroll
              for (int f = 0; f < AFT; ++f) s = fmaf(fv[f], wf[i][f], s);
            }
            const float r = rcpf(1.f + ex2f(s));            // tanh = 1 - 2r, 1 - tanh^2 = 4 r (1 - r)
            const float ds = (dej * v4c[i]) * fmaf(-r, r, r);
            dk[m][i] += ds;
            dvr[i] = fmaf(dej, r, dvr[i]);
            if (LOC) {
#pragma unroll
              for (int f = 0; f < AFT; ++f) {
                dWf[i][f] = fmaf(fv[f], ds, dWf[i][f]);
                dfp[f] = fmaf(ds, wf[i][f], dfp[f]);
              }
            }
          }
          if (LOC) {
#pragma unroll
            for (int f = 0; f < AFT; ++f) {
              float v = dfp[f];
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              if (ecl == 0 && jr < Tt) dfs[f * DFW + HALO + jr] = v * (1.f / K2LOG2E);
            }
          }
        }
      }
    }
    if (LOC) {
      __syncthreads();
      // d(conv kernel)[k][f] += sum_j a_{t-1}[j + k - pl] df[j][f], d(bias)[f] += sum_j df[j][f] (partial df over my channels: linear,
      // the channel blocks add up)
      if (cv_act) {
        const int j0 = cv_c * EG_JC;
        for (int ts = 0; ts < nts; ++ts) {
          const float* ap = S.aprev + (buf * EG_TB + ts) * DFW + HALO + j0 - pl;     // 16-byte aligned: HALO = 16, pl = 4, j0 % 8 = 0
          const float* dfs = S.dfS + (size_t)ts * AFT * DFW + cv_f * DFW + HALO + j0;
          float a[EG_JC + 12], g[EG_JC];
#pragma unroll
          for (int i = 0; i < (EG_JC + 12) / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(ap + 4 * i);
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
          }
#pragma unroll
          for (int i = 0; i < EG_JC / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(dfs + 4 * i);
            g[4 * i] = v.x; g[4 * i + 1] = v.y; g[4 * i + 2] = v.z; g[4 * i + 3] = v.w;
          }
#pragma unroll
          for (int jj = 0; jj < EG_JC; ++jj) {
#pragma unroll
            for (int k = 0; k < 10; ++k) cacc[k] = fmaf(a[jj + k], g[jj], cacc[k]);
            cacc[10] += g[jj];
          }
        }
      } else if (cf_e < ntap) {
        for (int ts = 0; ts < nts; ++ts) {
          const float* ap = S.aprev + (buf * EG_TB + ts) * DFW;
          const float* dfs = S.dfS + (size_t)ts * AFT * DFW + cf_f * DFW + HALO;
          float acc = 0.f;
          if (cf_k < d.att_kernel) {
            for (int j = cf_q; j < alen; j += 4) acc = fmaf(ap[HALO + j + cf_k - pl], dfs[j], acc);
          } else {
            for (int j = cf_q; j < alen; j += 4) acc += dfs[j];
          }
          dconv += acc;
        }
      }
    }
  }
  cl::cp_async_wait<0>();

  // ---------------- flush
  if (LOC) {
    if (conv10) {
      if (cv_act) {
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          if (cv_f < d.att_filters) {
            if (k < 10) atomicAdd(dd.dloc_conv_w + k * d.att_filters + cv_f, cacc[k]);
            else atomicAdd(dd.dloc_conv_b + cv_f, cacc[k]);
          }
        }
      }
    } else {
      dconv += __shfl_xor_sync(0xffffffffu, dconv, 1);
      dconv += __shfl_xor_sync(0xffffffffu, dconv, 2);
      if (cf_q == 0 && cf_e < ntap && cf_f < d.att_filters) {
        if (cf_k < d.att_kernel) atomicAdd(dd.dloc_conv_w + cf_k * d.att_filters + cf_f, dconv);
        else atomicAdd(dd.dloc_conv_b + cf_f, dconv);
      }
    }
  }
  float* dkeys = LOC ? dd.dkeys1 : dd.dkeys2;
#pragma unroll
  for (int m = 0; m < EG_MP; ++m) {
    const int jr = slot0 + 64 * m;
    if (jr < Tt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dkeys[((long long)jr * B + b) * AW + cb * EG_CB + ecl + 4 * i] = dk[m][i];
    }
  }
  // d(v), d(location layer): reduce over the 8 position lanes of the warp, then over the warps through shared memory
  __syncthreads();
  float* stage = S.fS;     // dead by now: [warp][4 channel lanes][4 + 4*AFT]
  constexpr int SW = 4 + 4 * AFT;
  // the channel lanes of one position share de: the sum over position lanes / warps of sum_de is the same for every channel
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) sum_de += __shfl_xor_sync(0xffffffffu, sum_de, o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = dvr[i];
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane < 4) stage[(warp * 4 + ecl) * SW + i] = fmaf(-2.f, v, sum_de);     // sum de * tanh
#pragma unroll
    for (int f = 0; f < AFT; ++f) {
      float w_ = dWf[i][f];
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) w_ += __shfl_xor_sync(0xffffffffu, w_, o);
      if (lane < 4) stage[(warp * 4 + ecl) * SW + 4 + i * AFT + f] = w_;
    }
  }
  __syncthreads();
  if (tid < 4 * SW) {
    const int cl_ = tid / SW, e = tid % SW;
    float acc = 0.f;
#pragma unroll
    for (int w_ = 0; w_ < EG_NT / 32; ++w_) acc += stage[(w_ * 4 + cl_) * SW + e];
    if (e < 4) {
      atomicAdd((LOC ? dd.dv1 : dd.dv2) + cb * EG_CB + cl_ + 4 * e, acc);
    } else if (LOC) {
      const int i = (e - 4) / AFT, f = (e - 4) % AFT;
      if (f < d.att_filters) atomicAdd(dd.dloc_layer_w + (long long)f * A1 + cb * EG_CB + cl_ + 4 * i, acc);
    }
  }
}

// One grid for both mechanisms: channel blocks 0..A1/EG_CB-1 belong to mechanism 1 (location features), the rest to mechanism 2.
// The short mechanism-2 CTAs fill the slots that the second, partial wave of mechanism-1 CTAs leaves idle.
__global__ void __launch_bounds__(EG_NT, 2) attn_energy_grad_kernel(const satk_attn_rnn_bwd_desc dd, const float* __restrict__ de,
                                                                    const float* __restrict__ fws) {
  if (blockIdx.x < A1 / EG_CB) energy_grad_body<true>(dd, de, fws, blockIdx.x);
  else energy_grad_body<false>(dd, de, fws, blockIdx.x - A1 / EG_CB);
}

// workspace: de [Td,B,2,Tt] followed by the location features [Td,B,Tt,MAXF]
int attn_energy_grad_launch(const satk_attn_rnn_bwd_desc* d, const float* de, int parts, cudaStream_t st) {
  SATK_CHECK_ARG(d->f.Tt <= 64 * EG_MP, "attn_energy_grad: Tt=%d out of range", d->f.Tt);
  EgSmem S;
  const size_t smem = S.carve(nullptr, d->f.Tt);
  float* fws = const_cast<float*>(de) + (((size_t)d->f.Td * d->f.B * 2 * d->f.Tt + 3) & ~(size_t)3);   // 16-byte aligned rows
  if (parts & 1) {      // location features of every step: they depend on the forward pass only
    loc_features_all_kernel<<<dim3(d->f.Td, d->f.B), 256, 0, st>>>(d->f, fws);
    SATK_LAUNCH_CHECK();
  }
  if (parts & 2) {
    SATK_CUDA(cudaFuncSetAttribute(attn_energy_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_energy_grad_kernel<<<dim3((A1 + A2) / EG_CB, d->f.B), EG_NT, smem, st>>>(*d, de, fws);
    SATK_LAUNCH_CHECK();
  }
  return SATK_OK;
}

}  // namespace arnn2
}  // namespace satk
