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
