// Persistent zoneout-LSTM recurrence (sm_100a): one thread-block cluster of H/16 CTAs owns 4 batch
// rows for the whole sequence.  Each CTA owns 16 hidden units (= 64 gate columns); its slice of the
// recurrent kernel Wh stays in REGISTERS for all T steps; the 4x H hidden state is exchanged through
// distributed shared memory with one hardware cluster barrier per step.  No global synchronisation,
// no re-reading of weights: per step the only HBM/L2 traffic is the pre-computed input projection
// xg[t] (prefetched one step ahead) and the saved activations.
//
// Math: TF LSTMCell (i,j,f,o; forget_bias) + tacotron2 ZoneoutLSTMCell, SURVEY.md A.5/A.6;
// sequence-length semantics of tf.nn.bidirectional_dynamic_rnn (module.py:93-108).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "cluster_sync.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace satk {

constexpr int LBG = 4;    // batch rows per cluster
constexpr int LUH = 16;   // hidden units per CTA

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2/(1+e^{2x}); absolute error ~2e-7, saturates correctly for |x| large
  float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

using cl::cp_async4;
using cl::cp_async_commit;
using cl::cp_async_wait;
using cl::st_async_v4;
using cl::RING;
using cl::PFD;

// UH = hidden units per CTA: 16 (cluster of H/16 CTAs x 256 threads, two CTAs per SM) or 32 (cluster of H/32 CTAs x 512 threads, one CTA
// per SM: with H = 256 and 8 clusters this gives every cluster its own SMs — 15 clusters of 8 are co-resident, only 7 of 16)
template <int H, int UH>
__global__ void __launch_bounds__(UH * 16, UH == 16 ? 2 : 1) lstm_fwd_kernel(const satk_lstm_fwd_desc d) {
  constexpr int LUH = UH;
  constexpr int NT = UH * 16;          // threads: 16 reduction lanes x UH column groups of 4 gate columns
  constexpr int PW = LBG * UH;         // pointwise threads (row, unit)
  constexpr int GC = 4 * UH;           // gate columns of this CTA
  constexpr int CS = H / LUH;
  constexpr int KPT = H / 4;  // weights per thread
  constexpr uint32_t SLICE_BYTES = LUH * LBG * 4;          // one CTA's slice of the hidden state
  constexpr uint32_t RX_BYTES = (CS - 1) * SLICE_BYTES;    // received from the peers every step
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * LBG;
  const int tid = threadIdx.x;

  __shared__ __align__(16) float hbuf[2][H][LBG];
  __shared__ float gsm[LBG][GC];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ float xg_ring[RING][NT];                      // this thread's x-projection element, PFD steps ahead
  __shared__ __align__(4) uint8_t mk_ring[RING][2][LBG][LUH];  // zoneout keep masks (c, h) of this CTA's units
  __shared__ float save_st[7][PW];                          // activations saved for backward, staged for the saver warps

  // --- gate-GEMM role: thread = (kq = lane & 15, column group cgp = tid >> 4): 4 gate columns x the k's
  // congruent to kq mod 16.  One LDS.128 of h[k][0..3] feeds 16 FMAs (4 columns x 4 rows); the 16 partial
  // sums are reduce-scattered over the 16 kq lanes, so lane L ends with (column cgp*4 + (L>>2)&3, row L&3).
  const int lane = tid & 31;
  const int kq = tid & 15, cgp = tid >> 4;
  float w[KPT / 4][4];
#pragma unroll
  for (int i = 0; i < KPT / 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int colc = cgp * 4 + c;                                  // CTA-local gate column = gate*UH + unit
      const int gc = (colc / LUH) * H + rank * LUH + (colc % LUH);   // column in the [.,4H] kernel
      w[i][c] = __ldg(d.Wh + (long long)(kq + 16 * i) * (4 * H) + gc);
    }
  const int col = cgp * 4 + ((kq >> 2) & 3);                         // the column / row this lane finalises
  const int gcol = (col / LUH) * H + rank * LUH + (col % LUH);
  const int myb = b0 + (kq & 3);
  const bool myb_ok = myb < d.B;
  const int mylen = myb_ok ? (d.lengths ? (int)d.lengths[myb] : d.T) : 0;

  // --- pointwise role: tid < 64 -> (pb, pu).  These two warps touch NO global memory: stores issued by the warp
  // that also issues the st.async exchange would sit in front of it in the LSU and delay every peer.
  // (unit-major: the 4 rows of a unit sit in 4 adjacent lanes, so one 16-byte st.async per peer carries them)
  const int pb = tid & 3, pu = (tid >> 2) % LUH;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < PW) && prow < d.B;
  const int plen = prow_ok ? (d.lengths ? (int)d.lengths[prow] : d.T) : 0;
  const int pidx = rank * LUH + pu;
  const int pslot = pb * LUH + pu;                           // index of (row, unit) in the staging arrays read by the saver warps
  float c_st = 0.f, h_st = 0.f;

  // --- saver / mask-prefetch role: tid in [64,128) -> (sb, su)
  const int sb = (tid - PW) / LUH, su = (tid - PW) % LUH;
  const int srow = b0 + sb;
  const bool srow_ok = (tid >= PW && tid < 2 * PW) && srow < d.B;
  const int slen = srow_ok ? (d.lengths ? (int)d.lengths[srow] : d.T) : 0;

  if (tid == 0) {
    cl::mbar_init(&bars[0], 1);
    cl::mbar_init(&bars[1], 1);
    cl::fence_mbar_init();
  }
  for (int i = tid; i < 2 * H * LBG; i += NT) (&hbuf[0][0][0])[i] = 0.f;
  for (int i = tid; i < RING * 2 * LBG * LUH; i += NT) (&mk_ring[0][0][0][0])[i] = 0;
  cluster.sync();

  // issue the prefetch of processing step s into ring slot s % RING (always commits one group)
  auto prefetch = [&](int s) {
    if (s < d.T) {
      if (s < mylen) {
        const int p = d.reverse ? (mylen - 1 - s) : s;
        cp_async4(&xg_ring[s % RING][tid], d.xg + ((long long)p * d.B + myb) * (4 * H) + gcol);
      }
      // masks are indexed by processing step; 4 units (4 bytes) per copy: saver threads with su % 4 == 0
      if (srow_ok && (su & 3) == 0 && s < slen) {
        const long long om = ((long long)s * d.B + srow) * H + rank * LUH + su;
        if (d.mask_c) cp_async4(&mk_ring[s % RING][0][sb][su], d.mask_c + om);
        if (d.mask_h) cp_async4(&mk_ring[s % RING][1][sb][su], d.mask_h + om);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int s = 0; s < PFD; ++s) prefetch(s);

  PT_DECL
#pragma unroll 1
  for (int s = 0; s < d.T; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    PT(5)
    prefetch(s + PFD);
    cp_async_wait<PFD>();                                              // step s's ring slot has landed
    if (s > 0) cl::mbar_wait(&bars[cur], ((s - 1) >> 1) & 1);         // peers' h(s) has landed
    if (tid == 0 && s + 1 < d.T) cl::mbar_arrive_expect_tx(&bars[nxt], RX_BYTES);  // arm the barrier of step s+1
    PT(0)
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < KPT / 4; ++i) {
      const float4 hv = *reinterpret_cast<const float4*>(&hbuf[cur][kq + 16 * i][0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[c * 4 + 0] = fmaf(w[i][c], hv.x, acc[c * 4 + 0]);
        acc[c * 4 + 1] = fmaf(w[i][c], hv.y, acc[c * 4 + 1]);
        acc[c * 4 + 2] = fmaf(w[i][c], hv.z, acc[c * 4 + 2]);
        acc[c * 4 + 3] = fmaf(w[i][c], hv.w, acc[c * 4 + 3]);
      }
    }
    const float mine = cl::reduce_scatter16(acc, lane);
    gsm[kq & 3][col] = mine + ((s < mylen) ? xg_ring[s % RING][tid] : 0.f);
    PT(1)
    __syncthreads();
    PT(2)
    if (tid < PW) {
      const bool valid = prow_ok && (s < plen);
      float gi = 0.f, gj = 0.f, gf = 0.f, go = 0.f, h_new = 0.f;
      const float c_old = c_st, h_old = h_st;
      if (valid) {
        const float mc = d.mask_c ? (float)mk_ring[s % RING][0][pb][pu] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mk_ring[s % RING][1][pb][pu] : (1.f - d.zh);
        gi = fast_sigmoid(gsm[pb][0 * LUH + pu]);
        gj = fast_tanh(gsm[pb][1 * LUH + pu]);
        gf = fast_sigmoid(gsm[pb][2 * LUH + pu] + d.forget_bias);
        go = fast_sigmoid(gsm[pb][3 * LUH + pu]);
        const float c_new = gf * c_st + gi * gj;
        h_new = go * fast_tanh(c_new);
        c_st = c_st + mc * (c_new - c_st);
        h_st = h_st + mh * (h_new - h_st);
      }
      if (s + 1 < d.T) {
        // publish h(s+1): local store + ONE 16-byte st.async per peer for the 4 rows of a unit (issued by the row-0 lane)
        hbuf[nxt][pidx][pb] = h_st;
        const int l4 = lane & ~3;
        const float h0 = __shfl_sync(0xffffffffu, h_st, l4), h1 = __shfl_sync(0xffffffffu, h_st, l4 + 1);
        const float h2 = __shfl_sync(0xffffffffu, h_st, l4 + 2), h3 = __shfl_sync(0xffffffffu, h_st, l4 + 3);
        if (pb == 0) {
          const uint32_t dsta = cl::smem_u32(&hbuf[nxt][pidx][0]);
          const uint32_t bara = cl::smem_u32(&bars[nxt]);
#pragma unroll
          for (int r = 0; r < CS; ++r)
            if (r != rank) st_async_v4(cl::mapa(dsta, r), h0, h1, h2, h3, cl::mapa(bara, r));
        }
      }
      // stage what the backward pass needs; zero past the row's length (dynamic_rnn semantics)
      save_st[0][pslot] = gi; save_st[1][pslot] = gj; save_st[2][pslot] = gf; save_st[3][pslot] = go;
      save_st[4][pslot] = valid ? c_old : 0.f;
      save_st[5][pslot] = valid ? h_old : 0.f;
      save_st[6][pslot] = h_new;
    }
    PT(3)
    __syncthreads();
    PT(4)
    if (srow_ok) {
      // saver warps: staged activations -> global memory (position-indexed), off the exchange's critical path
      const int e = tid - PW;
      const bool valid = s < slen;
      const int p = valid ? (d.reverse ? (slen - 1 - s) : s) : s;
      const int sidx = rank * LUH + su;
      const long long o1 = ((long long)p * d.B + srow) * H + sidx;
      d.out[((long long)p * d.B + srow) * d.ld_out + sidx] = save_st[6][e];
      if (d.gates) {
        const long long o4 = ((long long)p * d.B + srow) * (4 * H) + sidx;
        d.gates[o4] = save_st[0][e];
        d.gates[o4 + H] = save_st[1][e];
        d.gates[o4 + 2 * H] = save_st[2][e];
        d.gates[o4 + 3 * H] = save_st[3][e];
        d.c_prev[o1] = save_st[4][e];
        d.h_prev[o1] = save_st[5][e];
      }
    }
  }
  PT_FLUSH(d.T)
  cp_async_wait<0>();
  cluster.sync();  // nobody leaves while a peer may still be writing into its shared memory
}

template <int H>
__global__ void __launch_bounds__(256, 2) lstm_bwd_kernel(const satk_lstm_bwd_desc d) {
  constexpr int CS = H / LUH;
  constexpr int K4 = 4 * H;
  constexpr int NI = H / 64;                               // source units per thread in the GEMM
  constexpr uint32_t SLICE_BYTES = 4 * LUH * LBG * 4;      // one CTA's d(gates): 16 units x 4 rows x 4 gates
  constexpr uint32_t RX_BYTES = (CS - 1) * SLICE_BYTES;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int bg = blockIdx.x / CS;
  const int b0 = bg * LBG;
  const int tid = threadIdx.x;

  // d(gates) exchange buffer: [source unit][row][gate] -> one float4 per (unit, row), written with ONE st.async.v4 per peer
  __shared__ __align__(16) float dgx[2][H][LBG][4];
  __shared__ float dhpart[4][LBG][LUH];                    // [16-lane group of the k split][row][unit]
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ float in_ring[RING][6][64];                   // gates i,j,f,o, c_prev, dout of this CTA's units, PFD steps ahead
  __shared__ __align__(4) uint8_t mk_ring[RING][2][LBG][LUH];
  __shared__ __align__(16) float save_st[64][4];           // d(gates) staged for the saver warps

  // --- GEMM role (dh_prev = dg . Wh^T restricted to my 16 units): thread = (kq = tid & 63, out-unit group ug = tid >> 6)
  const int lane = tid & 31;
  const int kq = tid & 63, ug = tid >> 6;
  float w[NI][4][4];                                        // [i][out unit c][gate]
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4)
        w[i][c][g4] = __ldg(d.Wh + (long long)(rank * LUH + ug * 4 + c) * K4 + g4 * H + kq + 64 * i);

  // --- pointwise role (no global memory traffic in these two warps besides the async prefetch)
  const int pb = tid >> 4, pu = tid & 15;
  const int prow = b0 + pb;
  const bool prow_ok = (tid < 64) && prow < d.B;
  const int plen = prow_ok ? (d.lengths ? (int)d.lengths[prow] : d.T) : 0;
  const int pidx = rank * LUH + pu;
  float dc = 0.f, dh = 0.f;
  // --- saver role: tid in [64,128)
  const int sb = (tid - 64) >> 4, su = tid & 15;
  const int srow = b0 + sb;
  const bool srow_ok = (tid >= 64 && tid < 128) && srow < d.B;
  const int slen = srow_ok ? (d.lengths ? (int)d.lengths[srow] : d.T) : 0;

  if (tid == 0) {
    cl::mbar_init(&bars[0], 1);
    cl::mbar_init(&bars[1], 1);
    cl::fence_mbar_init();
  }
  for (int i = tid; i < RING * 2 * LBG * LUH; i += 256) (&mk_ring[0][0][0][0])[i] = 0;
  cluster.sync();

  // steps [Te, T) of this cluster's rows carry exactly zero gradient (satk_lstm_bwd_desc.step_end): the walk starts at Te - 1 and
  // their d(gates) rows are zero-filled here.  With 8 clusters on 7 cluster slots the clusters that share SMs finish earlier.
  int Te = d.T;
  if (d.step_end && !d.reverse && !d.lengths) {
    Te = 1;
    for (int r = 0; r < LBG; ++r)
      if (b0 + r < d.B) Te = max(Te, min(d.T, __ldg(d.step_end + b0 + r)));
    for (int r = rank; r < (d.T - Te) * LBG; r += CS) {
      const int tz = Te + r / LBG, bz = b0 + r % LBG;
      if (bz < d.B) {
        float* gz = d.dgates + ((long long)tz * d.B + bz) * K4;
        for (int i = tid; i < K4; i += 256) gz[i] = 0.f;
      }
    }
  }

  // prefetch of everything the pointwise step u (processing step s = Te-1-u) reads from global memory
  auto prefetch = [&](int u) {
    const int s = Te - 1 - u;
    if (s >= 0) {
      if (prow_ok && s < plen) {
        const int p = d.reverse ? (plen - 1 - s) : s;
        const long long o1 = ((long long)p * d.B + prow) * H + pidx;
        const long long o4 = ((long long)p * d.B + prow) * K4 + pidx;
        float* slot = &in_ring[u % RING][0][tid];
        cp_async4(slot + 0 * 64, d.gates + o4);
        cp_async4(slot + 1 * 64, d.gates + o4 + H);
        cp_async4(slot + 2 * 64, d.gates + o4 + 2 * H);
        cp_async4(slot + 3 * 64, d.gates + o4 + 3 * H);
        cp_async4(slot + 4 * 64, d.c_prev + o1);
        cp_async4(slot + 5 * 64, d.dout + ((long long)p * d.B + prow) * d.ld_dout + pidx);
        if ((pu & 3) == 0) {
          const long long om = ((long long)s * d.B + prow) * H + pidx;
          if (d.mask_c) cp_async4(&mk_ring[u % RING][0][pb][pu], d.mask_c + om);
          if (d.mask_h) cp_async4(&mk_ring[u % RING][1][pb][pu], d.mask_h + om);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int u = 0; u < PFD; ++u) prefetch(u);

#pragma unroll 1
  for (int s = Te - 1; s >= 0; --s) {
    const int u = Te - 1 - s, cur = u & 1;
    prefetch(u + PFD);
    cp_async_wait<PFD>();
    // the 4-byte mask words are fetched by every 4th lane: make them visible to the other pointwise lanes
    if (tid < 64) cl::named_bar_sync(1, 64);
    float dh_part = dh;
    if (tid < 64) {
      float dgi = 0.f, dgj = 0.f, dgf = 0.f, dgo = 0.f;
      const bool valid = prow_ok && (s < plen);
      if (valid) {
        const float* slot = &in_ring[u % RING][0][tid];
        const float gi = slot[0], gj = slot[64], gf = slot[128], go = slot[192], cp = slot[256], dout = slot[320];
        const float mc = d.mask_c ? (float)mk_ring[u % RING][0][pb][pu] : (1.f - d.zc);
        const float mh = d.mask_h ? (float)mk_ring[u % RING][1][pb][pu] : (1.f - d.zh);
        const float c_new = gf * cp + gi * gj;
        const float tc = fast_tanh(c_new);
        const float dh_new = dout + mh * dh;
        dh_part = (1.f - mh) * dh;
        const float dcn = mc * dc + dh_new * go * (1.f - tc * tc);
        dgo = dh_new * tc * go * (1.f - go);
        dgi = dcn * gj * gi * (1.f - gi);
        dgj = dcn * gi * (1.f - gj * gj);
        dgf = dcn * cp * gf * (1.f - gf);
        dc = (1.f - mc) * dc + dcn * gf;
      }
      if (s > 0) {
        // publish d(gates): local float4 + ONE st.async.v4 per peer
        if (tid == 0) cl::mbar_arrive_expect_tx(&bars[cur], RX_BYTES);
        float* slot = &dgx[cur][pidx][pb][0];
        *reinterpret_cast<float4*>(slot) = make_float4(dgi, dgj, dgf, dgo);
        const uint32_t dsta = cl::smem_u32(slot), bara = cl::smem_u32(&bars[cur]);
#pragma unroll
        for (int r = 0; r < CS; ++r)
          if (r != rank) st_async_v4(cl::mapa(dsta, r), dgi, dgj, dgf, dgo, cl::mapa(bara, r));
      }
      *reinterpret_cast<float4*>(&save_st[tid][0]) = make_float4(dgi, dgj, dgf, dgo);
    }
    __syncthreads();     // own slice + staged d(gates) visible to every thread of this CTA
    if (srow_ok) {
      const int e = tid - 64;
      const bool valid = s < slen;
      const int p = valid ? (d.reverse ? (slen - 1 - s) : s) : s;
      const long long o4 = ((long long)p * d.B + srow) * K4 + rank * LUH + su;
      const float4 v = *reinterpret_cast<const float4*>(&save_st[e][0]);
      d.dgates[o4] = v.x; d.dgates[o4 + H] = v.y; d.dgates[o4 + 2 * H] = v.z; d.dgates[o4 + 3 * H] = v.w;
    }
    if (s == 0) break;  // no earlier step consumes dh_prev
    cl::mbar_wait(&bars[cur], (u >> 1) & 1);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
      for (int b = 0; b < LBG; ++b) {
        const float4 g4 = *reinterpret_cast<const float4*>(&dgx[cur][kq + 64 * i][b][0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float a = acc[c * 4 + b];
          a = fmaf(w[i][c][0], g4.x, a);
          a = fmaf(w[i][c][1], g4.y, a);
          a = fmaf(w[i][c][2], g4.z, a);
          a = fmaf(w[i][c][3], g4.w, a);
          acc[c * 4 + b] = a;
        }
      }
    }
    {
      const float v = cl::reduce_scatter16(acc, lane);       // lane L: unit ug*4 + ((L>>2)&3), row L&3
      dhpart[(kq >> 4) & 3][lane & 3][ug * 4 + ((lane >> 2) & 3)] = v;
    }
    __syncthreads();
    if (tid < 64) dh = dh_part + dhpart[0][pb][pu] + dhpart[1][pb][pu] + dhpart[2][pb][pu] + dhpart[3][pb][pu];
  }
  cp_async_wait<0>();
  cluster.sync();
}

template <typename Kern, typename Desc>
static int launch_cluster(Kern kern, const Desc& d, int H, int B, cudaStream_t st, int uh = LUH) {
  const int CS = H / uh;
  const int groups = (B + LBG - 1) / LBG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CS);
  cfg.blockDim = dim3(uh * 16);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CS > 8) SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // two CTAs per SM (two clusters interleave on the same SMs and hide each other's exchange latency)
  SATK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  SATK_CUDA(cudaLaunchKernelEx(&cfg, kern, d));
  return SATK_OK;
}

int lstm_max_clusters_h256() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(8 * 64);
  cfg.blockDim = dim3(512);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 8;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_fwd_kernel<256, 32>, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

}  // namespace satk

namespace satk { namespace tc { int tc_trace(long long* out16); } }
namespace satk { namespace lstm5 {
int lstm5_bwd_launch(const satk_lstm_bwd_desc* d, cudaStream_t st);
int lstm5_fwd_launch(const satk_lstm_fwd_desc* d, cudaStream_t st);
} }
namespace satk { namespace arnn { int attn_fwd_phase_cycles(long long* out16); int attn_bwd_phase_cycles(long long* out16); } }
namespace satk { namespace arnn2 { int attn2_fwd_phase_cycles(long long* out16); int attn2_bwd_phase_cycles(long long* out16); } }
using namespace satk;

extern "C" {

int satk_debug_phase_cycles(int which, long long* out16) {
#ifdef SATK_PHASE_TIMING
  if (which == 1) return satk::arnn::attn_fwd_phase_cycles(out16);
  if (which == 2) return satk::arnn::attn_bwd_phase_cycles(out16);
  if (which == 3) return satk::tc::tc_trace(out16);
  if (which == 4) return satk::arnn2::attn2_fwd_phase_cycles(out16);
  if (which == 5) return satk::arnn2::attn2_bwd_phase_cycles(out16);
  SATK_CUDA(cudaMemcpyFromSymbol(out16, satk::g_phase, sizeof(long long) * 16));
  return 0;
#else
  (void)out16; (void)which;
  satk::set_error("library built without -DSATK_PHASE_TIMING");
  return SATK_ERR_UNSUPPORTED;
#endif
}

int satk_lstm_seq_fwd(const satk_lstm_fwd_desc* d, void* stream) {
  SATK_CHECK_ARG(d->H == 128 || d->H == 256, "lstm_seq_fwd: H=%d unsupported (128 or 256)", d->H);
  SATK_CHECK_ARG(d->T > 0 && d->B > 0, "lstm_seq_fwd: empty T=%d B=%d", d->T, d->B);
  SATK_CHECK_ARG(d->ld_out >= d->H, "lstm_seq_fwd: ld_out=%lld < H", d->ld_out);
  SATK_CHECK_ARG((d->gates == nullptr) == (d->c_prev == nullptr) && (d->gates == nullptr) == (d->h_prev == nullptr),
                 "lstm_seq_fwd: gates/c_prev/h_prev must be all set or all NULL");
  // H = 256: 32 units per CTA (clusters of 8 x 512 threads, one CTA per SM); H = 128: 16 units per CTA (clusters of 8 x 256 threads)
  // decoder layers (H = 256, forward order, no length masking): 5 rows per 16-CTA cluster, one wave at B = 32 (lstm_seq5.cu)
  const char* gen = getenv("SATK_LSTM_GEN");
  if (d->H == 256 && !d->lengths && !d->reverse && !(gen && gen[0] == '1')) return lstm5::lstm5_fwd_launch(d, (cudaStream_t)stream);
  if (d->H == 256) return launch_cluster(lstm_fwd_kernel<256, 32>, *d, 256, d->B, (cudaStream_t)stream, 32);
  return launch_cluster(lstm_fwd_kernel<128, 16>, *d, 128, d->B, (cudaStream_t)stream, 16);
}

int satk_lstm_seq_bwd(const satk_lstm_bwd_desc* d, void* stream) {
  SATK_CHECK_ARG(d->H == 128 || d->H == 256, "lstm_seq_bwd: H=%d unsupported (128 or 256)", d->H);
  SATK_CHECK_ARG(d->T > 0 && d->B > 0, "lstm_seq_bwd: empty T=%d B=%d", d->T, d->B);
  SATK_CHECK_ARG(d->ld_dout >= d->H, "lstm_seq_bwd: ld_dout=%lld < H", d->ld_dout);
  // decoder layers (H = 256, forward order, no length masking): 5 rows per cluster, one wave at B = 32 (lstm_seq5.cu)
  const char* gen = getenv("SATK_LSTM_GEN");
  if (d->H == 256 && !d->lengths && !d->reverse && !(gen && gen[0] == '1')) return lstm5::lstm5_bwd_launch(d, (cudaStream_t)stream);
  if (d->H == 256) return launch_cluster(lstm_bwd_kernel<256>, *d, 256, d->B, (cudaStream_t)stream);
  return launch_cluster(lstm_bwd_kernel<128>, *d, 128, d->B, (cudaStream_t)stream);
}

}  // extern "C"
