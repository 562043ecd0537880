// Zoneout-LSTM backward for the decoder layers (H = 256, no per-sequence lengths, forward time order), second generation:
// a cluster of 16 CTAs owns 5 batch rows, so a batch of 32 needs 7 clusters = the number of co-resident 16-CTA clusters of a
// B200 (the first-generation kernel owns 4 rows per cluster: 8 clusters, two of which share SMs and run at half speed).
//
// Per step, descending t (same scheme as BB / BC of attn_rnn2_bwd.cu):
//   * the CTA that owns 16 hidden units sums the 16 partial products of the previous step (+ its own zoneout carry) into
//     d(h_t), runs the LSTM cell backward and leaves d(gates) of its 64 gate columns in shared memory (and in global memory
//     for the dense weight-gradient GEMMs of the caller);
//   * d(h_{t-1}) = d(gates) . Wh^T restricted to MY 64 gate columns: thread = (quad of rows of Wh, quarter of my columns),
//     4 x 16 weights in registers, no reduction inside the CTA; a 4-lane reduce-scatter leaves each lane with the 16-byte
//     packet it sends to the owner of those 4 units (st.async completing an mbarrier transaction: one exchange per step, no
//     cluster-wide barrier in the loop).
// Math: TF LSTMCell (i,j,f,o; forget_bias) + tacotron2 ZoneoutLSTMCell; module.py:1525-1534.
#include <cooperative_groups.h>
#include "cluster_sync.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace satk {
namespace lstm5 {

using cl::cp_async_commit;
using cl::cp_async_wait;
using cl::st_async_v4;

constexpr int H = 256, CS = 16, UH = 16, NB = 5, NT = 256;
constexpr int RINGL = 4, PFDL = 2;
constexpr int RB = 6 * NB * UH;     // floats per ring slot: gates (4), c_prev, dout of my 16 units x NB rows
constexpr int DGG = 20;             // floats per gate in the d(gates) staging: the 4 column quarters then hit disjoint banks
constexpr int DGS = 4 * DGG;        // floats per batch row

__device__ __forceinline__ float fast_tanh(float x) {
  float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cl::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(NT, 1) lstm5_bwd_kernel(const satk_lstm_bwd_desc d) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CS) * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = d.T, B = d.B;
  constexpr int K4 = 4 * H;

  __shared__ __align__(16) float inH[2][CS][NB * UH];      // [parity][source CTA][row][unit]: partial d(h) of my units
  __shared__ __align__(16) float dgS[NB * DGS];            // d(gates) of my 64 gate columns [row][gate][DGG] (16 used)
  __shared__ __align__(16) float ring[RINGL][RB];
  __shared__ __align__(16) uint8_t mk_ring[RINGL][2][NB][UH];
  __shared__ __align__(8) uint64_t bars[2];

  int Te = T;
  if (d.step_end) {
    Te = 1;
    for (int r = 0; r < NB; ++r)
      if (b0 + r < B) Te = max(Te, min(T, __ldg(d.step_end + b0 + r)));
  }

  // BC role: thread = (quad of rows rq of Wh, quarter cq of my 64 gate columns)
  const int rq = tid >> 2, cq = tid & 3;
  float w[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int col = 16 * cq + c;                       // CTA-local gate column = gate*16 + unit
      w[r][c] = __ldg(d.Wh + (long long)(4 * rq + r) * K4 + (col >> 4) * H + rank * UH + (col & 15));
    }
  const int dst_cta = (4 * rq) >> 4, dst_unit = (4 * rq) & 15;
  // pointwise role: tid < NB*16 -> (row pu_u, unit pu_k)
  const int pu_u = tid >> 4, pu_k = tid & 15;
  const bool pw = tid < NB * UH;
  const bool pw_ok = pw && (b0 + pu_u) < B;
  float dc = 0.f, dh = 0.f;

  for (int i = tid; i < 2 * CS * NB * UH; i += NT) (&inH[0][0][0])[i] = 0.f;
  for (int i = tid; i < NB * DGS; i += NT) dgS[i] = 0.f;
  for (int i = tid; i < RINGL * RB; i += NT) (&ring[0][0])[i] = 0.f;
  for (int i = tid; i < RINGL * 2 * NB * UH; i += NT) (&mk_ring[0][0][0][0])[i] = 0;
  if (tid == 0) {
    cl::mbar_init(&bars[0], 1);
    cl::mbar_init(&bars[1], 1);
    cl::fence_mbar_init();
  }
  __syncthreads();
  cluster.sync();

  // rows of the skipped steps: zero gradients
  for (int r = rank; r < (T - Te) * NB; r += CS) {
    const int tz = Te + r / NB, bz = b0 + r % NB;
    if (bz < B) {
      float* gz = d.dgates + ((long long)tz * B + bz) * K4;
      for (int i = tid; i < K4; i += NT) gz[i] = 0.f;
    }
  }

  // prefetch of the pointwise inputs of time s into ring slot s % RINGL (always commits a group)
  auto prefetch = [&](int s) {
    if (s >= 0) {
      const int slot = s % RINGL;
      for (int e = tid; e < 6 * NB * 4; e += NT) {
        const int arr = e / (NB * 4), u = (e >> 2) % NB, q4 = e & 3;
        if (b0 + u >= B) continue;
        const long long rb = (long long)s * B + b0 + u;
        const float* src = (arr < 4) ? d.gates + rb * K4 + arr * H + rank * UH + 4 * q4
                         : (arr == 4) ? d.c_prev + rb * H + rank * UH + 4 * q4
                                      : d.dout + rb * d.ld_dout + rank * UH + 4 * q4;
        cp_async16(&ring[slot][(arr * NB + u) * UH + 4 * q4], src);
      }
      if (tid >= 128 && tid < 128 + 2 * NB) {
        const int which = (tid - 128) / NB, u = (tid - 128) % NB;
        const uint8_t* src = which ? d.mask_h : d.mask_c;
        if (src && b0 + u < B) cp_async16(&mk_ring[slot][which][u][0], src + ((long long)s * B + b0 + u) * H + rank * UH);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int i = 0; i <= PFDL; ++i) prefetch(Te - 1 - i);

#pragma unroll 1
  for (int s = Te - 1; s >= 0; --s) {
    const int u = Te - 1 - s, cur = u & 1, nxt = cur ^ 1;
    prefetch(s - 1 - PFDL);
    cp_async_wait<PFDL + 1>();
    __syncthreads();                                   // #0: ring slot of time s visible; d(gates) staging of the previous step is free
    if (u > 0) cl::mbar_wait(&bars[cur], (uint32_t)((u - 1) >> 1) & 1u);   // partial products of step s+1
    if (tid == 0 && s > 0) cl::mbar_arrive_expect_tx(&bars[nxt], (uint32_t)(CS * NB * UH * 4));
    if (pw) {
      float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
      if (pw_ok) {
        if (u > 0) {
          const float* ih = &inH[cur][0][tid];
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int r = 0; r < CS; r += 2) { s0 += ih[r * NB * UH]; s1 += ih[(r + 1) * NB * UH]; }
          dh += s0 + s1;                                 // recurrent carry: sum of the 16 partial products
        }
        const float* rp = &ring[s % RINGL][tid];
        const float gi = rp[0 * NB * UH], gj = rp[1 * NB * UH], gf = rp[2 * NB * UH], go = rp[3 * NB * UH];
        const float cp = rp[4 * NB * UH], dout = rp[5 * NB * UH];
        const float mc = d.mask_c ? (float)mk_ring[s % RINGL][0][pu_u][pu_k] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mk_ring[s % RINGL][1][pu_u][pu_k] : (1.f - d.zh);
        const float c_new = gf * cp + gi * gj;
        const float tc = fast_tanh(c_new);
        const float dh_new = dout + mh * dh;
        dh = (1.f - mh) * dh;
        const float dcn = mc * dc + dh_new * go * (1.f - tc * tc);
        dgo = dh_new * tc * go * (1.f - go);
        dgi = dcn * gj * gi * (1.f - gi);
        dgj = dcn * gi * (1.f - gj * gj);
        dgf = dcn * cp * gf * (1.f - gf);
        dc = (1.f - mc) * dc + dcn * gf;
      }
      float* dg = dgS + pu_u * DGS + pu_k;
      dg[0] = dgi; dg[DGG] = dgj; dg[2 * DGG] = dgf; dg[3 * DGG] = dgo;
    }
    __syncthreads();                                   // #1
    if (s > 0) {
      float acc[NB][4];
#pragma unroll
      for (int uu = 0; uu < NB; ++uu) {
        acc[uu][0] = 0.f; acc[uu][1] = 0.f; acc[uu][2] = 0.f; acc[uu][3] = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 g4 = *reinterpret_cast<const float4*>(&dgS[uu * DGS + DGG * cq + 4 * c4]);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[uu][r] = fmaf(w[r][4 * c4], g4.x, acc[uu][r]); acc[uu][r] = fmaf(w[r][4 * c4 + 1], g4.y, acc[uu][r]);
            acc[uu][r] = fmaf(w[r][4 * c4 + 2], g4.z, acc[uu][r]); acc[uu][r] = fmaf(w[r][4 * c4 + 3], g4.w, acc[uu][r]);
          }
        }
      }
      // reduce-scatter over the 4 column quarters: lane q ends with the 4 rows of batch row q; the fifth row is gathered by lane 0
      const bool up2 = (cq & 2) != 0, up1 = (cq & 1) != 0;
      float o4[4], k[2][4];
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
      float k0 = up2 ? acc[4][2] : acc[4][0], k1 = up2 ? acc[4][3] : acc[4][1];
      k0 += __shfl_xor_sync(0xffffffffu, up2 ? acc[4][0] : acc[4][2], 2);
      k1 += __shfl_xor_sync(0xffffffffu, up2 ? acc[4][1] : acc[4][3], 2);
      float kk = up1 ? k1 : k0;
      kk += __shfl_xor_sync(0xffffffffu, up1 ? k0 : k1, 1);
      const int l4 = lane & ~3;
      const float e0 = __shfl_sync(0xffffffffu, kk, l4), e1 = __shfl_sync(0xffffffffu, kk, l4 + 1);
      const float e2 = __shfl_sync(0xffffffffu, kk, l4 + 2), e3 = __shfl_sync(0xffffffffu, kk, l4 + 3);
      const uint32_t dsta = cl::mapa(cl::smem_u32(&inH[nxt][rank][dst_unit]), dst_cta);
      const uint32_t barr = cl::mapa(cl::smem_u32(&bars[nxt]), dst_cta);
      st_async_v4(dsta + cq * UH * 4, o4[0], o4[1], o4[2], o4[3], barr);
      if (cq == 0) st_async_v4(dsta + 4 * UH * 4, e0, e1, e2, e3, barr);
    }
    if (warp == NT / 32 - 1) {
      // d(gates) of step s -> global (after this warp's DSMEM stores of the step)
      for (int e = lane; e < NB * 16; e += 32) {
        const int uu = e >> 4, g4 = (e >> 2) & 3, q4 = e & 3;
        if (b0 + uu < B)
          *reinterpret_cast<float4*>(d.dgates + ((long long)s * B + b0 + uu) * K4 + g4 * H + rank * UH + 4 * q4) =
              *reinterpret_cast<const float4*>(&dgS[uu * DGS + g4 * DGG + 4 * q4]);
      }
    }
  }
  cp_async_wait<0>();
  cluster.sync();
}

// ------------------------------------------------------------------------------------------------------------------
// Forward, same geometry: a cluster of 16 CTAs owns 5 batch rows; a CTA owns 16 hidden units (64 gate columns, its slice of Wh in
// registers).  Per step: gates = xg[t] + h_{t-1} . Wh over the full 256-wide state (thread = 16 k's x 4 columns, the 20 partial
// sums reduce-scattered over the 16 k lanes), the LSTM cell on my 16 units, then the new 16 x 5 state slice goes to every CTA of
// the cluster (st.async completing the mbarrier transaction of the next step; rows 0..3 of a unit are one 16-byte packet, the
// fifth rows of 4 neighbouring units another).
constexpr int FRING = cl::RING, FPFD = cl::PFD;
constexpr int PWF = NB * UH;        // pointwise threads: tid < 64 -> (row tid & 3, unit tid >> 2); 64..79 -> (row 4, unit tid - 64)

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__global__ void __launch_bounds__(NT, 1) lstm5_fwd_kernel(const satk_lstm_fwd_desc d) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / CS) * NB;
  const int tid = threadIdx.x, lane = tid & 31;
  const int T = d.T, B = d.B;
  constexpr int K4 = 4 * H;
  constexpr uint32_t RX_BYTES = (CS - 1) * UH * NB * 4;

  __shared__ __align__(16) float hq[2][H][4];             // hidden state, rows 0..3
  __shared__ __align__(16) float h4[2][H];                // hidden state, row 4
  __shared__ float gsm[NB][4 * UH];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ float xg_ring[FRING][NT + 64];               // x-projection elements of the lanes that finalise a (row, column)
  __shared__ __align__(4) uint8_t mk_ring[FRING][2][NB][UH];
  __shared__ float save_st[7][PWF];

  // gate-GEMM role: thread = (kq = tid & 15, column group cgp = tid >> 4): 4 gate columns x the 16 k's congruent to kq mod 16
  const int kq = tid & 15, cgp = tid >> 4;
  float w[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int colc = cgp * 4 + c;                                  // CTA-local gate column = gate*UH + unit
      w[i][c] = __ldg(d.Wh + (long long)(kq + 16 * i) * K4 + (colc / UH) * H + rank * UH + (colc % UH));
    }
  const int col = cgp * 4 + ((kq >> 2) & 3);                         // the column this lane finalises (rows kq & 3 and, lanes kq & 3 == 0, row 4)
  const int gcol = (col / UH) * H + rank * UH + (col % UH);
  const int myb = b0 + (kq & 3);
  const bool myb_ok = myb < B;
  const bool my4 = (kq & 3) == 0 && (b0 + 4) < B;

  // pointwise role
  const bool pw = tid < PWF;
  const int pb = tid < 64 ? (tid & 3) : 4, pu = tid < 64 ? (tid >> 2) : ((tid - 64) & 15);
  const bool prow_ok = pw && (b0 + pb) < B;
  const int pidx = rank * UH + pu;
  const int pslot = pb * UH + pu;
  float c_st = 0.f, h_st = 0.f;
  // saver / mask-prefetch role: tid in [96, 96 + 80) -> (row sb, unit su)
  const int se = tid - 96;
  const int sb = se >= 0 ? se / UH : 0, su = se >= 0 ? se % UH : 0;
  const int srow = b0 + sb;
  const bool srow_ok = se >= 0 && se < PWF && srow < B;

  if (tid == 0) {
    cl::mbar_init(&bars[0], 1);
    cl::mbar_init(&bars[1], 1);
    cl::fence_mbar_init();
  }
  for (int i = tid; i < 2 * H * 4; i += NT) (&hq[0][0][0])[i] = 0.f;
  for (int i = tid; i < 2 * H; i += NT) (&h4[0][0])[i] = 0.f;
  for (int i = tid; i < FRING * 2 * NB * UH; i += NT) (&mk_ring[0][0][0][0])[i] = 0;
  for (int i = tid; i < FRING * (NT + 64); i += NT) (&xg_ring[0][0])[i] = 0.f;
  __syncthreads();
  cluster.sync();

  auto prefetch = [&](int s) {
    if (s < T) {
      if (myb_ok) cl::cp_async4(&xg_ring[s % FRING][tid], d.xg + ((long long)s * B + myb) * K4 + gcol);
      if (my4) cl::cp_async4(&xg_ring[s % FRING][NT + (tid >> 2)], d.xg + ((long long)s * B + b0 + 4) * K4 + gcol);
      if (srow_ok && (su & 3) == 0) {
        const long long om = ((long long)s * B + srow) * H + rank * UH + su;
        if (d.mask_c) cl::cp_async4(&mk_ring[s % FRING][0][sb][su], d.mask_c + om);
        if (d.mask_h) cl::cp_async4(&mk_ring[s % FRING][1][sb][su], d.mask_h + om);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int s = 0; s < FPFD; ++s) prefetch(s);

#pragma unroll 1
  for (int s = 0; s < T; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    prefetch(s + FPFD);
    cp_async_wait<FPFD>();
    if (s > 0) cl::mbar_wait(&bars[cur], ((s - 1) >> 1) & 1);         // the peers' slices of h(s) have landed
    if (tid == 0 && s + 1 < T) cl::mbar_arrive_expect_tx(&bars[nxt], RX_BYTES);
    float acc[16], a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 hv = *reinterpret_cast<const float4*>(&hq[cur][kq + 16 * i][0]);
      const float he = h4[cur][kq + 16 * i];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[c * 4 + 0] = fmaf(w[i][c], hv.x, acc[c * 4 + 0]);
        acc[c * 4 + 1] = fmaf(w[i][c], hv.y, acc[c * 4 + 1]);
        acc[c * 4 + 2] = fmaf(w[i][c], hv.z, acc[c * 4 + 2]);
        acc[c * 4 + 3] = fmaf(w[i][c], hv.w, acc[c * 4 + 3]);
        a4[c] = fmaf(w[i][c], he, a4[c]);
      }
    }
    const float mine = cl::reduce_scatter16(acc, lane);               // (column (kq >> 2) & 3, row kq & 3)
    // fifth row: 4 columns reduce-scattered over the 4 lane quads, then summed inside the quad
    {
      const bool up8 = (kq & 8) != 0, up4 = (kq & 4) != 0;
      const float k0 = (up8 ? a4[2] : a4[0]) + __shfl_xor_sync(0xffffffffu, up8 ? a4[0] : a4[2], 8);
      const float k1 = (up8 ? a4[3] : a4[1]) + __shfl_xor_sync(0xffffffffu, up8 ? a4[1] : a4[3], 8);
      float m = (up4 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up4 ? k0 : k1, 4);
      m += __shfl_xor_sync(0xffffffffu, m, 2);
      m += __shfl_xor_sync(0xffffffffu, m, 1);
      if ((kq & 3) == 0) gsm[4][col] = m + xg_ring[s % FRING][NT + (tid >> 2)];
    }
    gsm[kq & 3][col] = mine + xg_ring[s % FRING][tid];
    __syncthreads();
    if (pw) {
      float gi = 0.f, gj = 0.f, gf = 0.f, go = 0.f, h_new = 0.f;
      const float c_old = c_st, h_old = h_st;
      if (prow_ok) {
        const float mc = d.mask_c ? (float)mk_ring[s % FRING][0][pb][pu] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mk_ring[s % FRING][1][pb][pu] : (1.f - d.zh);
        gi = fast_sigmoid(gsm[pb][0 * UH + pu]);
        gj = fast_tanh(gsm[pb][1 * UH + pu]);
        gf = fast_sigmoid(gsm[pb][2 * UH + pu] + d.forget_bias);
        go = fast_sigmoid(gsm[pb][3 * UH + pu]);
        const float c_new = gf * c_st + gi * gj;
        h_new = go * fast_tanh(c_new);
        c_st = c_st + mc * (c_new - c_st);
        h_st = h_st + mh * (h_new - h_st);
      }
      if (s + 1 < T) {
        // publish h(s+1): the 4 lanes of a packet (4 rows of a unit / fifth rows of 4 units) all hold it and share the 15 peers
        const int l4 = lane & ~3;
        const unsigned pm = tid < 64 ? 0xffffffffu : 0x0000ffffu;      // the third warp is only half in this role
        const float p0 = __shfl_sync(pm, h_st, l4), p1 = __shfl_sync(pm, h_st, l4 + 1);
        const float p2 = __shfl_sync(pm, h_st, l4 + 2), p3 = __shfl_sync(pm, h_st, l4 + 3);
        float* dstp = tid < 64 ? &hq[nxt][pidx][0] : &h4[nxt][rank * UH + (pu & ~3)];
        if (tid < 64) hq[nxt][pidx][pb] = h_st; else h4[nxt][pidx] = h_st;
        const uint32_t dsta = cl::smem_u32(dstp), bara = cl::smem_u32(&bars[nxt]);
#pragma unroll
        for (int r4 = 0; r4 < CS; r4 += 4) {
          const int r = r4 + (lane & 3);
          if (r != rank) st_async_v4(cl::mapa(dsta, r), p0, p1, p2, p3, cl::mapa(bara, r));
        }
      }
      save_st[0][pslot] = gi; save_st[1][pslot] = gj; save_st[2][pslot] = gf; save_st[3][pslot] = go;
      save_st[4][pslot] = prow_ok ? c_old : 0.f;
      save_st[5][pslot] = prow_ok ? h_old : 0.f;
      save_st[6][pslot] = h_new;
    }
    __syncthreads();
    if (srow_ok) {
      // saver warps: staged activations -> global memory, off the exchange's critical path
      const int sidx = rank * UH + su;
      const long long o1 = ((long long)s * B + srow) * H + sidx;
      d.out[((long long)s * B + srow) * d.ld_out + sidx] = save_st[6][se];
      if (d.gates) {
        const long long o4 = ((long long)s * B + srow) * K4 + sidx;
        d.gates[o4] = save_st[0][se];
        d.gates[o4 + H] = save_st[1][se];
        d.gates[o4 + 2 * H] = save_st[2][se];
        d.gates[o4 + 3 * H] = save_st[3][se];
        d.c_prev[o1] = save_st[4][se];
        d.h_prev[o1] = save_st[5][se];
      }
    }
  }
  cp_async_wait<0>();
  cluster.sync();  // nobody leaves while a peer may still be writing into its shared memory
}

template <typename Kern, typename Desc>
static int launch5(Kern kern, const Desc& d, cudaStream_t st) {
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((d.B + NB - 1) / NB) * CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = 0;
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

int lstm5_fwd_launch(const satk_lstm_fwd_desc* d, cudaStream_t st) { return launch5(lstm5_fwd_kernel, *d, st); }

int lstm5_bwd_launch(const satk_lstm_bwd_desc* d, cudaStream_t st) {
  SATK_CUDA(cudaFuncSetAttribute(lstm5_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((d->B + NB - 1) / NB) * CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SATK_CUDA(cudaLaunchKernelEx(&cfg, lstm5_bwd_kernel, *d));
  return SATK_OK;
}

}  // namespace lstm5
}  // namespace satk
