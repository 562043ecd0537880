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
constexpr int EG_CB = 32;          // score channels per CTA
constexpr int EG_KS = 40;          // key row stride (4 rows x 8 channel lanes -> 32 distinct banks)
constexpr int EG_MP = 6;           // position passes of 32 slots: Tt <= 192
constexpr int EG_TB = 4;           // decoder steps per iteration (amortises the barriers, hides the global loads)

struct EgSmem {
  int TtP, DFW;
  float *keyS, *fS, *dfS, *aprev, *deS, *qS, *wconv, *bconv;
  __host__ __device__ size_t carve(float* base, int Tt) {
    TtP = (Tt + 31) / 32 * 32;
    DFW = TtP + 2 * HALO;
    float* p = base;
    keyS = p; p += (size_t)TtP * EG_KS;
    fS = p; p += (size_t)EG_TB * TtP * MAXF;
    dfS = p; p += (size_t)EG_TB * AFT * DFW;
    aprev = p; p += 2 * EG_TB * DFW;           // [buffer][step][HALO + j]
    deS = p; p += 2 * EG_TB * TtP;
    qS = p; p += 2 * EG_TB * EG_CB;
    wconv = p; p += MAXK * MAXF;
    bconv = p; p += MAXF;
    return (size_t)(p - base) * sizeof(float);
  }
};

template <bool LOC>
__global__ void __launch_bounds__(EG_NT, 2) attn_energy_grad_kernel(const satk_attn_rnn_bwd_desc dd, const float* __restrict__ de) {
  const satk_attn_rnn_fwd_desc& d = dd.f;
  const int b = blockIdx.y;
  const int cb = LOC ? blockIdx.x : 0;             // channel block inside the mechanism
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B, Td = d.Td;
  const int alen = min((int)d.lengths[b], Tt);
  const int Te = dd.step_end ? max(1, min(Td, __ldg(dd.step_end + b))) : Td;
  const int pl = (d.att_kernel - 1) / 2;
  const int AW = LOC ? A1 : A2;                    // width of the mechanism's key rows
  const float* keys = LOC ? d.keys1 : d.keys2;
  const int qoff = LOC ? cb * EG_CB : A1;          // column of my channels inside the saved query rows

  extern __shared__ __align__(16) float smem_raw[];
  EgSmem S;
  S.carve(smem_raw, Tt);
  const int TtP = S.TtP, DFW = S.DFW;

  const int ep = lane >> 3, ecl = lane & 7;
  const int slot0 = warp * 4 + ep;                 // positions slot0 + 32 m
  const int npass = min(EG_MP, (alen - 4 * warp + 31) / 32);   // passes of this warp that touch the utterance (warp-uniform)

  for (int i = tid; i < TtP * EG_KS; i += EG_NT) {
    const int j = i / EG_KS, c = i % EG_KS;
    float kv = 0.f;
    if (j < Tt && c < EG_CB) kv = (__ldg(keys + ((long long)j * B + b) * AW + cb * EG_CB + c) + ((LOC && d.b1) ? __ldg(d.b1 + cb * EG_CB + c) : 0.f)) * K2LOG2E;
    S.keyS[i] = kv;
  }
  for (int i = tid; i < MAXK * MAXF; i += EG_NT) {
    const int k = i / MAXF, f = i % MAXF;
    S.wconv[i] = (LOC && k < d.att_kernel && f < d.att_filters) ? __ldg(d.loc_conv_w + k * d.att_filters + f) : 0.f;
  }
  if (tid < MAXF) S.bconv[tid] = (LOC && tid < d.att_filters) ? __ldg(d.loc_conv_b + tid) : 0.f;
  for (int i = tid; i < 2 * EG_TB * DFW; i += EG_NT) S.aprev[i] = 0.f;
  for (int i = tid; i < EG_TB * AFT * DFW; i += EG_NT) S.dfS[i] = 0.f;
  for (int i = tid; i < EG_TB * TtP * MAXF; i += EG_NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * TtP; i += EG_NT) S.deS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * EG_CB; i += EG_NT) S.qS[i] = 0.f;

  // per-channel constants of my 4 channels (ecl + 8 i)
  float v4c[4], wf[4][AFT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cb * EG_CB + ecl + 8 * i;
    v4c[i] = 4.f * __ldg((LOC ? d.v1 : d.v2) + c);
#pragma unroll
    for (int f = 0; f < AFT; ++f) wf[i][f] = (LOC && f < d.att_filters) ? __ldg(d.loc_layer_w + (long long)f * A1 + c) * K2LOG2E : 0.f;
  }
  // accumulators: dkeys tile, sum de*r per channel (dv = v-free: sum de*tanh = sum de - 2 sum de*r), d(location layer)
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
  // d(location convolution): thread = (tap k, filter f, quarter of the positions); tap index att_kernel = the bias
  const int cf_e = tid >> 2, cf_q = tid & 3;
  const int ntap = LOC ? (d.att_kernel + 1) * AFT : 0;
  const int cf_k = cf_e / AFT, cf_f = cf_e % AFT;
  float dconv = 0.f;

  // asynchronous loads of de_t, a_{t-1}, q_t (my channels) of the EG_TB steps starting at t0 into buffer `buf`
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
    if (LOC) {
      for (int ts = 0; ts < nts; ++ts)
        arnn::location_features<AFT>(S.fS + (size_t)ts * TtP * MAXF, S.aprev + (buf * EG_TB + ts) * DFW, S.wconv, S.bconv, alen, d.att_kernel, pl,
                                     tid, EG_NT);
      __syncthreads();
    }
#pragma unroll 1
    for (int ts = 0; ts < nts; ++ts) {
      float q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = S.qS[(buf * EG_TB + ts) * EG_CB + ecl + 8 * i] * K2LOG2E;
      const float* des = S.deS + (buf * EG_TB + ts) * TtP;
      const float* fs = S.fS + (size_t)ts * TtP * MAXF;
      float* dfs = S.dfS + (size_t)ts * AFT * DFW;
#pragma unroll
      for (int m = 0; m < EG_MP; ++m) {
        if (m < npass) {
          const int jr = slot0 + 32 * m, j = min(jr, Tt - 1);
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
            float s = S.keyS[j * EG_KS + ecl + 8 * i] + q[i];
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
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              if (ecl == 0 && jr < Tt) dfs[f * DFW + HALO + jr] = v * (1.f / K2LOG2E);
            }
          }
        }
      }
    }
    if (LOC) {
      __syncthreads();
      if (cf_e < ntap) {
        // d(conv kernel)[k][f] += sum_j a_{t-1}[j + k - pl] df[j][f] (partial df over my 32 channels: linear, the blocks add up)
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
    dconv += __shfl_xor_sync(0xffffffffu, dconv, 1);
    dconv += __shfl_xor_sync(0xffffffffu, dconv, 2);
    if (cf_q == 0 && cf_e < ntap && cf_f < d.att_filters) {
      if (cf_k < d.att_kernel) atomicAdd(dd.dloc_conv_w + cf_k * d.att_filters + cf_f, dconv);
      else atomicAdd(dd.dloc_conv_b + cf_f, dconv);
    }
  }
  float* dkeys = LOC ? dd.dkeys1 : dd.dkeys2;
#pragma unroll
  for (int m = 0; m < EG_MP; ++m) {
    const int jr = slot0 + 32 * m;
    if (jr < Tt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dkeys[((long long)jr * B + b) * AW + cb * EG_CB + ecl + 8 * i] = dk[m][i];
    }
  }
  // d(v), d(location layer): reduce over the 4 position lanes of the warp, then over the warps through shared memory
  __syncthreads();
  float* stage = S.keyS;   // keys are dead: [warp][8 lanes][4 + 4*AFT]
  constexpr int SW = 4 + 4 * AFT;
  // the channel lanes of one position share de: sum over the position lanes / warps of sum_de is the same for every channel
  sum_de += __shfl_xor_sync(0xffffffffu, sum_de, 8);
  sum_de += __shfl_xor_sync(0xffffffffu, sum_de, 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = dvr[i];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 8) stage[(warp * 8 + ecl) * SW + i] = fmaf(-2.f, v, sum_de);     // sum de * tanh
#pragma unroll
    for (int f = 0; f < AFT; ++f) {
      float w_ = dWf[i][f];
      w_ += __shfl_xor_sync(0xffffffffu, w_, 8);
      w_ += __shfl_xor_sync(0xffffffffu, w_, 16);
      if (lane < 8) stage[(warp * 8 + ecl) * SW + 4 + i * AFT + f] = w_;
    }
  }
  __syncthreads();
  if (tid < 8 * SW) {
    const int cl_ = tid / SW, e = tid % SW;
    float acc = 0.f;
#pragma unroll
    for (int w_ = 0; w_ < EG_NT / 32; ++w_) acc += stage[(w_ * 8 + cl_) * SW + e];
    if (e < 4) {
      atomicAdd((LOC ? dd.dv1 : dd.dv2) + cb * EG_CB + cl_ + 8 * e, acc);
    } else if (LOC) {
      const int i = (e - 4) / AFT, f = (e - 4) % AFT;
      if (f < d.att_filters) atomicAdd(dd.dloc_layer_w + (long long)f * A1 + cb * EG_CB + cl_ + 8 * i, acc);
    }
  }
}

int attn_energy_grad_launch(const satk_attn_rnn_bwd_desc* d, const float* de, cudaStream_t st) {
  SATK_CHECK_ARG(d->f.Tt <= 32 * EG_MP, "attn_energy_grad: Tt=%d out of range", d->f.Tt);
  EgSmem S;
  const size_t smem = S.carve(nullptr, d->f.Tt);
  SATK_CUDA(cudaFuncSetAttribute(attn_energy_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SATK_CUDA(cudaFuncSetAttribute(attn_energy_grad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_energy_grad_kernel<true><<<dim3(A1 / EG_CB, d->f.B), EG_NT, smem, st>>>(*d, de);
  SATK_LAUNCH_CHECK();
  attn_energy_grad_kernel<false><<<dim3(1, d->f.B), EG_NT, smem, st>>>(*d, de);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

}  // namespace arnn2
}  // namespace satk
