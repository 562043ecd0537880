// Gradients of the attention energies that do NOT feed the recurrence, for all (step, utterance) pairs in parallel:
//   e_t[j] = sum_c v_c tanh(keys[j,c] + q_t[c] + (f_t[j,:] . Wf[:,c]) + b_c),  f_t = conv1d(a_{t-1}) + bias
// (forward_attention.py:13-26,98-100; TF BahdanauAttention score, A.8).  The sequential backward kernel (attn_rnn2_bwd.cu)
// saves de_t[j] for both mechanisms; here the tanh terms are recomputed and
//   dkeys[j,c] = sum_t ds,  dv[c] += sum de tanh,  dWf[f,c] += sum f ds,  dconv[k,f] += sum a_{t-1}[j+k-pl] df[j,f],  dbias[f] += sum df
// with ds = de v (1 - tanh^2), df[j,f] = sum_c ds Wf[f,c].  Nothing here is on the step-to-step dependency chain.
//
// Work unit = (chunk of EG chunk steps, utterance, block of 16 score channels); persistent CTAs (two per SM) pull units from a
// queue in the order in which the recurrence finishes them (it walks the steps downwards) and add their partial sums with
// atomics.  Launched as a programmatic dependent of the recurrence kernel, the grid starts while the recurrence still runs — on
// the 36 SMs its 7 clusters leave idle — and waits per unit for the progress flag the recurrence publishes (EG_PUB); launched
// the ordinary way it runs after the recurrence and never waits.
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
  float *keyS, *fS, *dfS, *aprev, *deS, *qS, *stage;
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
    stage = p; p += (EG_NT / 32) * 4 * (4 + 4 * AFT);   // flush of d(v) / d(location layer): [warp][4 channel lanes][4 + 4*AFT]
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

// Conv-gradient accumulators of a worker: they do not depend on the utterance or the channel block, so they live in registers
// across units and are flushed once.
struct EgConvAcc {
  float cacc[11];
  float dconv;
};

// One unit: steps [ta, tb) of utterance b, channel block cb of mechanism 1 (LOC) or 2.
template <bool LOC>
__device__ __forceinline__ void energy_grad_unit(const satk_attn_rnn_bwd_desc& dd, const float* __restrict__ de,
                                                 const float* __restrict__ fws, const EgSmem& S, const int cb, const int b,
                                                 const int ta, const int tb, EgConvAcc& cv) {
  const satk_attn_rnn_fwd_desc& d = dd.f;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tt = d.Tt, B = d.B;
  const int DEL = de_row_stride(Tt);
  const int alen = min((int)d.lengths[b], Tt);
  const int pl = (d.att_kernel - 1) / 2;
  const int AW = LOC ? A1 : A2;                    // width of the mechanism's key rows
  const float* keys = LOC ? d.keys1 : d.keys2;
  const int qoff = (LOC ? 0 : A1) + cb * EG_CB;    // column of my channels inside the saved query rows
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
  if (LOC) {
    // positions past this utterance's length keep the zeros the conv-gradient stage expects (an earlier unit may have been longer)
    for (int i = tid; i < EG_TB * AFT * DFW; i += EG_NT) S.dfS[i] = 0.f;
  }

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
  float (&cacc)[11] = cv.cacc;
  const int cf_e = tid >> 2, cf_q = tid & 3;
  const int ntap = (LOC && !conv10) ? (d.att_kernel + 1) * AFT : 0;
  const int cf_k = cf_e / AFT, cf_f = cf_e % AFT;
  float& dconv = cv.dconv;

  // asynchronous loads of de_t, a_{t-1}, f_t, q_t (my channels) of the EG_TB steps starting at t0 into buffer `buf`
  auto load_steps = [&](int t0, int buf) {
    for (int e = tid; e < EG_TB * Tt; e += EG_NT) {
      const int ts = e / Tt, j = e % Tt, t = t0 + ts;
      if (t < tb) {
        cl::cp_async4(&S.deS[(buf * EG_TB + ts) * TtP + j], de + (((long long)t * B + b) * 2 + (LOC ? 0 : 1)) * DEL + j);
        if (LOC) {
          if (t > 0) cl::cp_async4(&S.aprev[(buf * EG_TB + ts) * DFW + HALO + j], d.soft1 + ((long long)(t - 1) * B + b) * Tt + j);
          else S.aprev[(buf * EG_TB + ts) * DFW + HALO + j] = 0.f;
        }
      }
    }
    if (LOC) {
      for (int e = tid; e < EG_TB * Tt * 2; e += EG_NT) {          // two 16-byte halves per position
        const int ts = e / (2 * Tt), r = e % (2 * Tt), t = t0 + ts;
        if (t < tb) cp_async16(&S.fS[((size_t)(buf * EG_TB + ts) * TtP) * MAXF + 4 * r], fws + ((long long)t * B + b) * Tt * MAXF + 4 * r);
      }
    }
    if (tid < EG_TB * EG_CB) {
      const int ts = tid / EG_CB, c = tid % EG_CB, t = t0 + ts;
      if (t < tb) cl::cp_async4(&S.qS[(buf * EG_TB + ts) * EG_CB + c], d.q_save + ((long long)t * B + b) * QT + qoff + c);
    }
    cl::cp_async_commit();
  };

  __syncthreads();                         // the fills above are ordered before the first asynchronous copies
  load_steps(ta, 0);
#pragma unroll 1
  for (int t0 = ta, it = 0; t0 < tb; t0 += EG_TB, ++it) {
    const int buf = it & 1;
    const int nts = min(EG_TB, tb - t0);
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

  // ---------------- flush (partial sums of this unit)
  float* dkeys = LOC ? dd.dkeys1 : dd.dkeys2;
#pragma unroll
  for (int m = 0; m < EG_MP; ++m) {
    const int jr = slot0 + 64 * m;
    if (m < npass && jr < Tt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(&dkeys[((long long)jr * B + b) * AW + cb * EG_CB + ecl + 4 * i], dk[m][i]);
    }
  }
  // d(v), d(location layer): reduce over the 8 position lanes of the warp, then over the warps through shared memory
  float* stage = S.stage;
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
  __syncthreads();                         // the next unit overwrites keyS / dfS / stage
}

// queue / progress words shared with the recurrence kernel (satk_attn_rnn_bwd_desc.sync_ws)
constexpr int EGQ_NEXT = 0;        // next unit
constexpr int EGQ_PROG = 4;        // [B][2]: lowest step whose d(energies) row is complete (16-byte aligned start)

// Persistent worker.  Unit order: chunks from the last steps down (the order in which the recurrence finishes them), inside a chunk
// utterance-major so that the 16 channel blocks of one (chunk, utterance) run at about the same time and share its rows in L2.
__global__ void __launch_bounds__(EG_NT, 2) attn_energy_grad_kernel(const satk_attn_rnn_bwd_desc dd, const float* __restrict__ de,
                                                                    const float* __restrict__ fws, int* __restrict__ sync,
                                                                    const int chunk, const int wait_progress) {
  const satk_attn_rnn_fwd_desc& d = dd.f;
  extern __shared__ __align__(16) float smem_raw[];
  __shared__ int unit_sh;
  EgSmem S;
  S.carve(smem_raw, d.Tt);
  const int tid = threadIdx.x;
  const int TtP = S.TtP, DFW = S.DFW;
  // pads and halos stay zero for the whole launch: the asynchronous copies only write positions < Tt
  for (int i = tid; i < 2 * EG_TB * DFW; i += EG_NT) S.aprev[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * TtP * MAXF; i += EG_NT) S.fS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * TtP; i += EG_NT) S.deS[i] = 0.f;
  for (int i = tid; i < 2 * EG_TB * EG_CB; i += EG_NT) S.qS[i] = 0.f;
  for (int i = tid; i < EG_TB * AFT * DFW; i += EG_NT) S.dfS[i] = 0.f;
  EgConvAcc cv;
#pragma unroll
  for (int k = 0; k < 11; ++k) cv.cacc[k] = 0.f;
  cv.dconv = 0.f;
  constexpr int NCB = (A1 + A2) / EG_CB;
  const int nchunk = (d.Td + chunk - 1) / chunk;
  const int total = nchunk * d.B * NCB;
#pragma unroll 1
  for (;;) {
    __syncthreads();
    if (tid == 0) unit_sh = atomicAdd(sync + EGQ_NEXT, 1);
    __syncthreads();
    const int u = unit_sh;
    if (u >= total) break;
    const int c = nchunk - 1 - u / (d.B * NCB);
    const int r = u % (d.B * NCB);
    const int b = r / NCB, cbx = r % NCB;
    const int Te = dd.step_end ? max(1, min(d.Td, __ldg(dd.step_end + b))) : d.Td;
    const int ta = c * chunk, tb = min(Te, ta + chunk);
    if (ta >= tb) continue;                  // steps without a loss
    const bool loc = cbx < A1 / EG_CB;
    if (wait_progress) {
      if (tid == 0) {
        const volatile int* flag = sync + EGQ_PROG + b * 2 + (loc ? 0 : 1);
        // bounded: the recurrence needs ~4 ms for all its steps; if its flag has not moved after ~1 s something upstream died —
        // fail loudly (a trap surfaces as a CUDA error at the next synchronisation) instead of spinning forever
        long long spins = 0;
        while (*flag > ta) {
          __nanosleep(256);
          if (++spins > (1ll << 22)) __trap();
        }
        __threadfence();
      }
      __syncthreads();
    }
    if (loc) energy_grad_unit<true>(dd, de, fws, S, cbx, b, ta, tb, cv);
    else energy_grad_unit<false>(dd, de, fws, S, cbx - A1 / EG_CB, b, ta, tb, cv);
  }
  // ---------------- d(location convolution) of everything this worker has seen
  if (d.att_kernel > 0 && dd.dloc_conv_w) {
    const int Tt = d.Tt;
    const bool conv10 = d.att_kernel == 10;
    const int nch = (Tt + EG_JC - 1) / EG_JC;
    const int cv_f = tid / nch;
    const int cf_e = tid >> 2, cf_q = tid & 3;
    const int ntap = conv10 ? 0 : (d.att_kernel + 1) * AFT;
    const int cf_k = cf_e / AFT, cf_f = cf_e % AFT;
    if (conv10) {
      if (cv_f < AFT && cv_f < d.att_filters) {
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          if (k < 10) atomicAdd(dd.dloc_conv_w + k * d.att_filters + cv_f, cv.cacc[k]);
          else atomicAdd(dd.dloc_conv_b + cv_f, cv.cacc[k]);
        }
      }
    } else {
      float dconv = cv.dconv;
      dconv += __shfl_xor_sync(0xffffffffu, dconv, 1);
      dconv += __shfl_xor_sync(0xffffffffu, dconv, 2);
      if (cf_q == 0 && cf_e < ntap && cf_f < d.att_filters) {
        if (cf_k < d.att_kernel) atomicAdd(dd.dloc_conv_w + cf_k * d.att_filters + cf_f, dconv);
        else atomicAdd(dd.dloc_conv_b + cf_f, dconv);
      }
    }
  }
  // as a programmatic dependent: this grid is complete only when the recurrence grid is (and its results are visible)
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

size_t attn_energy_grad_sync_ints(int B) { return EGQ_PROG + 2 * (size_t)B; }
int* attn_energy_grad_progress(const satk_attn_rnn_bwd_desc* d) { return d->sync_ws + EGQ_PROG; }   // what the recurrence publishes into

static int eg_chunk() {
  const char* e = getenv("SATK_EG_CHUNK");
  int c = e ? atoi(e) : 32;
  c = (c + EG_PUB - 1) / EG_PUB * EG_PUB;      // chunk starts are published progress values
  return c < EG_PUB ? EG_PUB : c;
}

// Queue / flag words and the dkeys accumulators, zeroed on the stream BEFORE the recurrence kernel of an overlapped launch
// (nothing may sit between the recurrence and its programmatic dependent).
int attn_energy_grad_prepare(const satk_attn_rnn_bwd_desc* d, cudaStream_t st) {
  const int B = d->f.B, Tt = d->f.Tt;
  SATK_CUDA(cudaMemsetAsync(d->sync_ws, 0x7f, attn_energy_grad_sync_ints(B) * sizeof(int), st));   // progress: "nothing yet"
  SATK_CUDA(cudaMemsetAsync(d->sync_ws, 0, EGQ_PROG * sizeof(int), st));
  SATK_CUDA(cudaMemsetAsync(d->dkeys1, 0, (size_t)Tt * B * A1 * sizeof(float), st));
  SATK_CUDA(cudaMemsetAsync(d->dkeys2, 0, (size_t)Tt * B * A2 * sizeof(float), st));
  return SATK_OK;
}

// workspace: de [Td,B,2,de_row_stride(Tt)] followed by the location features [Td,B,Tt,MAXF]
// parts: 1 = location features, 2 = gradients; dependent != 0: the gradient launch is a programmatic dependent of the kernel
// issued just before it on `st` (the recurrence; `attn_energy_grad_prepare` has run before that one)
int attn_energy_grad_launch(const satk_attn_rnn_bwd_desc* d, const float* de, int parts, int dependent, cudaStream_t st) {
  SATK_CHECK_ARG(d->f.Tt <= 64 * EG_MP, "attn_energy_grad: Tt=%d out of range", d->f.Tt);
  EgSmem S;
  const size_t smem = S.carve(nullptr, d->f.Tt);
  float* fws = const_cast<float*>(de) + (size_t)d->f.Td * d->f.B * 2 * de_row_stride(d->f.Tt);   // 128-byte aligned rows
  if (parts & 1) {      // location features of every step: they depend on the forward pass only
    loc_features_all_kernel<<<dim3(d->f.Td, d->f.B), 256, 0, st>>>(d->f, fws);
    SATK_LAUNCH_CHECK();
  }
  if (parts & 2) {
    SATK_CHECK_ARG(d->sync_ws != nullptr, "attn_energy_grad: sync_ws missing");
    if (!dependent) {
      int rc = attn_energy_grad_prepare(d, st);
      if (rc) return rc;
    }
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      SATK_CUDA(cudaGetDevice(&dev));
      SATK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int chunk = eg_chunk();
    const int units = ((d->f.Td + chunk - 1) / chunk) * d->f.B * ((A1 + A2) / EG_CB);
    SATK_CUDA(cudaFuncSetAttribute(attn_energy_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(units < 2 * sms ? units : 2 * sms);
    cfg.blockDim = dim3(EG_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = dependent ? 1 : 0;
    SATK_CUDA(cudaLaunchKernelEx(&cfg, attn_energy_grad_kernel, *d, de, (const float*)fws, d->sync_ws, chunk, dependent ? 1 : 0));
  }
  return SATK_OK;
}

}  // namespace arnn2
}  // namespace satk
