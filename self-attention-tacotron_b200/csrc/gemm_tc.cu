// tcgen05 GEMM tile with fp32-class accuracy ("3xTF32"), engine 2 of satk_gemm (sm_100a).
//
//   C[M,N] = epilogue( alpha * sum_tap A_tap[M,K] . B_tap[N,K]^T )        A, B K-contiguous ("K-major")
//
// * operands arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 3-stage shared-memory ring;
// * fp32 inputs are split in shared memory into hi = tf32(x) and lo = x - hi (element-wise, so the swizzled
//   layout is preserved); three tcgen05.mma kind::tf32 per k-step accumulate hi*hi + hi*lo + lo*hi in TMEM:
//   relative error ~2^-21 instead of TF32's 2^-11 — the 1e-3 parity budget over 400 recurrent steps needs it;
// * one elected thread issues the MMAs; tcgen05.commit releases ring slots / signals the epilogue;
// * the epilogue reads the 128x128 fp32 accumulator with tcgen05.ld and applies bias / activation /
//   dropout mask / residual / beta exactly like the SIMT tile (gemm_simt.cu).
// Weight-gradient form (round 2): C[M,N] = sum_k A[k,m] . B[k,n] with BOTH operands "MN-major" (row-major [K, M] / [K, N], the
// reduction index k = the row): X^T.dY straight from the activations and their gradients, no transposed copies.  A 128 x 32 tile is
// four TMA boxes of 32 k-rows x 128 bytes (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); the shared-memory descriptors say MN-major with
// layout type 128B_BASE32B (leading byte offset = distance of the 32-float MN blocks, stride byte offset = distance of the 4-row k
// groups) and the instruction descriptor sets the a_major / b_major bits.
// Convolution taps = extra iterations of the K loop with a row-shifted TMA coordinate for A (time-major rows:
// a tap is a shift by B rows; out-of-range rows are zero-filled by TMA) and the tap index as third coordinate of B.
#include <cuda.h>
#include <stdlib.h>
#include "cluster_sync.cuh"
#include "common.cuh"

namespace satk {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;      // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;         // 16 KB (A and B tiles have the same size)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;     // A_hi(raw) | A_lo | B_hi(raw) | B_lo
constexpr int NTHREADS = 256;
constexpr int XFORM_THREADS = 192;              // warps 2..7
constexpr int XFORM_WARPS = 6;
constexpr int TMEM_COLS = 128;

#if defined(SATK_PHASE_TIMING) && defined(SATK_TC_LIFE)
// lifecycle marks of CTA (0,0,0): cycles since kernel entry at setup-done / first TMA issued / first tile landed / first split
// done / first MMA issue / last commit issued / accumulator complete / epilogue done / teardown
#define TC_TRACE(ev)
#define TC_MARK(i) if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) satk::g_phase[i] = clock64() - t_start;
#elif defined(SATK_PHASE_TIMING)
#define TC_MARK(i)
#define TC_TRACE(ev) if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it >= 6 && it < 10) satk::g_phase[(it - 6) * 4 + (ev)] = clock64() - t_start;
#else
#define TC_TRACE(ev)
#define TC_MARK(i)
#endif

struct Params {
  int M, N, K;
  int taps, shift0, tap_dir;
  float* C; long long ldc;
  float alpha, beta;
  const float* bias;
  int act;
  const float* residual; long long ldres;
  const uint8_t* keep_mask; float keep_scale;
  int split_k;
  int zbatch, kshift0, kshift_step;   // z-batches with a shifted reduction coordinate of A (conv weight gradients)
  long long c_zstride;
  int bank, bank_a_kstep, bank_c_nstep;   // conv bank: z-batch entry = conv width (see satk_gemm_desc.bank_widths)
  int tma_store;            // 1: plain overwrite epilogue (bias + activation only) leaves through TMA bulk stores
  // batched self-attention products (satk_gemm_desc.zcoord): per-entry coordinate shifts into shared 2-D operand views
  int zcoord, za_row, za_k, zb_row, zb_k, zc_col;
  int causal;               // satk_gemm_desc.causal_skip (zcoord form only)
  int c_rank3;              // output map is [z][M][N] (rank 3): rows / columns outside an entry's slab are clipped by the TMA unit
  int prec;                 // satk_gemm_desc.precision: 1 = one TF32 pass on the raw operands (no hi/lo split)
  int a_mn, b_mn;           // 1: the operand is MN-major (stored row-major [K, M] / [K, N], the reduction index is the row): both for the
                            // weight-gradient products; B (P.V, dS.K) or both (P^T.dO, dS^T.Q) for the self-attention products
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cl::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cl::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

// MN-major fp32 / tf32 operand: layout type 128B_BASE32B (the 128-byte swizzle whose atom is 32 bytes: Swizzle<2,5,2>, what a TMA map
// with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes), 32-float MN blocks of [32 k-rows x 128 B] 4096 bytes apart (leading byte offset),
// 4-row k groups 512 bytes apart (stride byte offset); an 8-deep k step advances the start address by 1024 bytes.  Measured with
// tools/micro/mn_probe.cu: with the plain SWIZZLE_128B layout type a tf32 MMA with a_major / b_major = MN returns zeros on sm_100a.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(4096 >> 4) << 16;                  // leading byte offset
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return d;
}

// Epilogue of one 32-row x 32-column block held in shared memory (stg, row stride 33): lane = column.  Residual / accumulate
// operands of all 32 rows are loaded before the first use so the loads overlap.
template <int ACT, bool MASK, bool ADD>
__device__ __forceinline__ void epi_rows(const Params& p, const float* stg, int lane, int mbase, int nrows, int n, float bias, long long zoff) {
  const float alpha = p.alpha, beta = p.beta, keep_scale = p.keep_scale;
  const long long ldc = p.ldc, ldres = p.ldres;
  float* cp = p.C + zoff + (long long)mbase * ldc + n;
  const float* rp = p.residual ? p.residual + (long long)mbase * ldres + n : nullptr;
  const uint8_t* kp = MASK ? p.keep_mask + (long long)mbase * p.N + n : nullptr;
  const int N = p.N;
  float add[32];
  if (ADD) {
    if (rp) {
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) add[rr] = (rr < nrows) ? __ldg(rp + rr * ldres) : 0.f;
    } else {
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) add[rr] = 0.f;
    }
    if (beta != 0.0f) {
      float cv[32];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) cv[rr] = (rr < nrows) ? cp[rr * ldc] : 0.f;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) add[rr] = fmaf(beta, cv[rr], add[rr]);
    }
  }
  uint8_t keep[32];
  if (MASK) {
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) keep[rr] = (rr < nrows) ? kp[(long long)rr * N] : (uint8_t)0;
  }
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    if (rr < nrows) {
      float v = apply_act(fmaf(alpha, stg[rr * 33 + lane], bias), ACT);
      if (MASK) v = keep[rr] ? v * keep_scale : 0.0f;
      if (ADD) v += add[rr];
      cp[rr * ldc] = v;
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC,
               const Params p, const int b_rank3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], xform_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_sh;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ks = blockIdx.z % p.split_k, zb = blockIdx.z / p.split_k;
  if (p.causal == 1 && n0 >= m0 + BM) return;       // QK^T / dP under a causal mask: the softmax never reads this tile
  int a_kshift = p.kshift0 + zb * p.kshift_step;
  int a_rowoff = 0, b_rowoff = 0, b_kshift = 0;
  int taps_eff = p.taps, shift0_eff = p.shift0, tap_base = 0, c_noff = 0, c_rowz = zb * p.M;
  long long c_zoff = zb * p.c_zstride;
  if (p.bank) {
    const int w = p.zbatch - 1 - zb;               // widest conv first: the long tiles are scheduled before the short ones
    taps_eff = w + 1;
    shift0_eff = -(w / 2) * p.tap_dir;             // SAME padding: (k-1)/2 taps to the left
    tap_base = w * (w + 1) / 2;
    a_kshift = w * p.bank_a_kstep;
    c_noff = w * p.bank_c_nstep;
    c_rowz = 0;
    c_zoff = 0;
  }
  if (p.zcoord) {
    a_kshift = zb * p.za_k; a_rowoff = zb * p.za_row;
    b_kshift = zb * p.zb_k; b_rowoff = zb * p.zb_row;
    c_noff = zb * p.zc_col;
    c_rowz = 0;
  }
  // carve: align the dynamic region to 1024 B (swizzle atom alignment)
  const uint32_t smem_base = (cl::smem_u32(smem) + 1023u) & ~1023u;

  // causal structure of the attention products: P[m, k] = 0 for k > m, so P.V reduces over k < m0 + BM only (2) and
  // P^T.dO over k >= m0 only (3)
  const int k_eff = p.causal == 2 ? min(p.K, m0 + BM) : p.K;
  const int kb_first = p.causal == 3 ? m0 / BK : 0;
  const int kblocks_total = (k_eff + BK - 1) / BK - kb_first;
  // split-K partitions the flattened (tap, k-block) iteration space, so tapped (conv) products with a short K still split
  const int it_total = kblocks_total * taps_eff;
  const int it_per = (it_total + p.split_k - 1) / p.split_k;
  const int it_beg = ks * it_per;
  const int iters = max(0, min(it_total, it_beg + it_per) - it_beg);
#ifdef SATK_PHASE_TIMING
  const long long t_start = clock64();
#endif

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      cl::mbar_init(&full_bar[s], 1);
      cl::mbar_init(&xform_bar[s], XFORM_WARPS);
      cl::mbar_init(&empty_bar[s], 1);
    }
    cl::mbar_init(&accum_bar, 1);
    cl::fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cl::smem_u32(&tmem_base_sh)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  if (tid == 0) { TC_MARK(0) }

  if (warp == 0 && lane == 0) {
    // ===== TMA producer
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES, use = it / STAGES;
      if (use > 0) cl::mbar_wait(&empty_bar[s], (use - 1) & 1);
      const int tap = (it_beg + it) / kblocks_total, kb = kb_first + (it_beg + it) % kblocks_total;
      const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + 2 * TILE_BYTES;
      TC_TRACE(0)
      cl::mbar_arrive_expect_tx(&full_bar[s], 2 * TILE_BYTES);
      // MN-major operand: coordinate 0 = the M / N column, coordinate 1 = the reduction row; one box per 32-float column block
      if (p.a_mn) {
#pragma unroll
        for (int j = 0; j < BM / 32; ++j)
          tma_load_2d(sa + j * 4096, &mapA, m0 + a_rowoff + 32 * j, kb * BK + a_kshift, cl::smem_u32(&full_bar[s]));
      } else {
        tma_load_2d(sa, &mapA, kb * BK + a_kshift, m0 + a_rowoff + shift0_eff + tap * p.tap_dir, cl::smem_u32(&full_bar[s]));
      }
      if (p.b_mn) {
#pragma unroll
        for (int j = 0; j < BN / 32; ++j)
          tma_load_2d(sb + j * 4096, &mapB, n0 + b_rowoff + 32 * j, kb * BK + b_kshift, cl::smem_u32(&full_bar[s]));
      } else if (b_rank3) {
        tma_load_3d(sb, &mapB, kb * BK, n0, tap_base + tap, cl::smem_u32(&full_bar[s]));
      } else {
        tma_load_2d(sb, &mapB, kb * BK + b_kshift, n0 + b_rowoff, cl::smem_u32(&full_bar[s]));
      }
      if (it == 0) { TC_MARK(1) }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24) |
                           (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);       // a_major / b_major = MN
    // bytes per 8-deep k step: two 4-row groups (MN-major) / 32 bytes of the row (K-major)
    const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES, use = it / STAGES;
      cl::mbar_wait(&xform_bar[s], use & 1);
      if (it == 0) { TC_MARK(4) }
      TC_TRACE(2)
      tc_fence_after();
      const uint32_t sa = smem_base + s * STAGE_BYTES;
      const uint64_t a_hi = p.a_mn ? make_desc_mn(sa) : make_desc(sa), a_lo = p.a_mn ? make_desc_mn(sa + TILE_BYTES) : make_desc(sa + TILE_BYTES);
      const uint64_t b_hi = p.b_mn ? make_desc_mn(sa + 2 * TILE_BYTES) : make_desc(sa + 2 * TILE_BYTES);
      const uint64_t b_lo = p.b_mn ? make_desc_mn(sa + 3 * TILE_BYTES) : make_desc(sa + 3 * TILE_BYTES);
      if (p.prec == 1) {
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t ak = (uint64_t)((k * a_kstep) >> 4), bk = (uint64_t)((k * b_kstep) >> 4);
          tc_mma_tf32(tmem_base, a_hi + ak, b_hi + bk, idesc, (it > 0 || k > 0) ? 1u : 0u);   // raw fp32: the tensor core keeps 10 mantissa bits
        }
      } else {
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t ak = (uint64_t)((k * a_kstep) >> 4), bk = (uint64_t)((k * b_kstep) >> 4);
          tc_mma_tf32(tmem_base, a_lo + ak, b_hi + bk, idesc, (it > 0 || k > 0) ? 1u : 0u);
          tc_mma_tf32(tmem_base, a_hi + ak, b_lo + bk, idesc, 1u);
          tc_mma_tf32(tmem_base, a_hi + ak, b_hi + bk, idesc, 1u);
        }
      }
      tc_commit(&empty_bar[s]);          // slot reusable once these MMAs have read it
      TC_TRACE(3)
    }
    tc_commit(&accum_bar);               // accumulator complete
    TC_MARK(5)
  } else if (warp >= 2) {
    // ===== hi/lo split of both operand tiles (element-wise: the swizzled layout is untouched)
    const int xt = tid - 64;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES, use = it / STAGES;
      cl::mbar_wait(&full_bar[s], use & 1);
      if (tid == 64) { TC_TRACE(1) }
      if (tid == 64 && it == 0) { TC_MARK(2) }
      uint8_t* base = smem + (smem_base - cl::smem_u32(smem)) + s * STAGE_BYTES;
#pragma unroll 2
      for (int i = xt; i < (p.prec == 1 ? 0 : 2 * (TILE_BYTES / 16)); i += XFORM_THREADS) {
        const int which = i / (TILE_BYTES / 16), c = i % (TILE_BYTES / 16);
        float4* hi = reinterpret_cast<float4*>(base + which * 2 * TILE_BYTES) + c;
        float4* lo = reinterpret_cast<float4*>(base + which * 2 * TILE_BYTES + TILE_BYTES) + c;
        float4 v = *hi;
        float4 h;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        *hi = h;
        *lo = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      cl::fence_proxy_async();           // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&xform_bar[s]);
      if (tid == 64 && it == 0) { TC_MARK(3) }
    }
  }

  // ===== epilogue: warps 4..7 own TMEM lanes 32*(warp-4) .. +31 (row = lane)
  if (warp >= 4) {
    cl::mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (tid == 128) { TC_MARK(6) }
    const int q = warp - 4;
    if (p.tma_store) {
      // ---- fast path: all four 32-column blocks of this warp's 32 rows are read from TMEM at once (one TMEM round trip
      // instead of four), bias + activation applied in registers (lane = row), written to shared memory in the 128-byte
      // swizzle layout and handed to the TMA unit: the warps never wait on global stores (a plain STG row loop costs ~120
      // cycles per 128-byte row when every SM drains its tile at the same time) and TMA clips rows / columns outside C.
      uint32_t r[4][32];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
            "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[cb][0]), "=r"(r[cb][1]), "=r"(r[cb][2]), "=r"(r[cb][3]), "=r"(r[cb][4]), "=r"(r[cb][5]), "=r"(r[cb][6]), "=r"(r[cb][7]),
              "=r"(r[cb][8]), "=r"(r[cb][9]), "=r"(r[cb][10]), "=r"(r[cb][11]), "=r"(r[cb][12]), "=r"(r[cb][13]), "=r"(r[cb][14]),
              "=r"(r[cb][15]), "=r"(r[cb][16]), "=r"(r[cb][17]), "=r"(r[cb][18]), "=r"(r[cb][19]), "=r"(r[cb][20]), "=r"(r[cb][21]),
              "=r"(r[cb][22]), "=r"(r[cb][23]), "=r"(r[cb][24]), "=r"(r[cb][25]), "=r"(r[cb][26]), "=r"(r[cb][27]), "=r"(r[cb][28]),
              "=r"(r[cb][29]), "=r"(r[cb][30]), "=r"(r[cb][31])
            : "r"(taddr));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (tid == 128) { TC_MARK(9) }
      const float alpha = p.alpha;
      const int act = p.split_k == 1 ? p.act : 0;
      const bool has_bias = p.bias != nullptr && p.split_k == 1;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const int nb = n0 + cb * 32;
        if (iters > 0 && nb < p.N && m0 + q * 32 < p.M) {
          // staging block: 32 rows x 128 B, 1024-byte aligned, one per (warp, column block)
          const uint32_t sblk = smem_base + (uint32_t)((q * 4 + cb) * 4096);
          const float bl = (has_bias && nb + lane < p.N) ? __ldg(p.bias + nb + lane) : 0.f;   // lane j holds bias[nb + j]
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float b = __shfl_sync(0xffffffffu, bl, c * 4 + e);
              v[e] = fmaf(alpha, __uint_as_float(r[cb][c * 4 + e]), b);
            }
            if (act == SATK_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            } else if (act == SATK_ACT_TANH) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = tanhf_(v[e]);
            } else if (act == SATK_ACT_SIGMOID) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = sigmoidf_(v[e]);
            }
            const uint32_t addr = sblk + (uint32_t)(lane * 128) + (uint32_t)(((c ^ (lane & 7)) & 7) * 16);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          }
          cl::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (p.c_rank3)          // stacked [z][M][N] output: the slab index is the third coordinate
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&mapC), "r"(nb),
                           "r"(m0 + q * 32), "r"(zb), "r"(sblk)
                           : "memory");
            else if (p.tma_store == 2)   // split-K partial sums / accumulation onto C: the TMA unit adds into global memory
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&mapC),
                           "r"(c_noff + nb), "r"(c_rowz + m0 + q * 32), "r"(sblk)
                           : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&mapC), "r"(c_noff + nb),
                           "r"(c_rowz + m0 + q * 32), "r"(sblk)
                           : "memory");
          }
        }
      }
      if (lane == 0) {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must stay valid until the TMA unit has read it
      }
      __syncwarp();
    } else
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
          "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (tid == 128 && c0 == 0) { TC_MARK(9) }
      // stage the 32x32 block of this warp in shared memory (the operand ring is idle now), then write rows coalesced
      float* stg = reinterpret_cast<float*>(smem + (smem_base - cl::smem_u32(smem))) + q * (32 * 33);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      if (tid == 128 && c0 == 0) { TC_MARK(10) }
      const int n = n0 + c0 + lane;
      if (iters > 0 && n < p.N) {
        const float bias = (p.bias && p.split_k == 1) ? __ldg(p.bias + n) : 0.f;
        const int mbase = m0 + q * 32;
        const int nrows = min(32, p.M - mbase);
        if (p.split_k > 1) {
          float* cp = p.C + c_zoff + (long long)mbase * p.ldc + n;
          const float alpha = p.alpha;
          const long long ldc = p.ldc;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr)
            if (rr < nrows) atomicAdd(cp + rr * ldc, alpha * stg[rr * 33 + lane]);
        } else {
          // every epilogue option is resolved OUTSIDE the row loop (a per-row switch on kernel parameters costs a chain of
          // uniform-datapath loads and branches per row, ~270 cycles each with one warp per scheduler: measured 34 K cycles)
          const bool has_add = p.residual != nullptr || p.beta != 0.0f, has_mask = p.keep_mask != nullptr;
#define SATK_EPI(A)                                                                                   \
  do {                                                                                                \
    if (has_mask) { if (has_add) epi_rows<A, true, true>(p, stg, lane, mbase, nrows, n, bias, c_zoff); else epi_rows<A, true, false>(p, stg, lane, mbase, nrows, n, bias, c_zoff); } \
    else { if (has_add) epi_rows<A, false, true>(p, stg, lane, mbase, nrows, n, bias, c_zoff); else epi_rows<A, false, false>(p, stg, lane, mbase, nrows, n, bias, c_zoff); } \
  } while (0)
          switch (p.act) {
            case SATK_ACT_RELU: SATK_EPI(SATK_ACT_RELU); break;
            case SATK_ACT_TANH: SATK_EPI(SATK_ACT_TANH); break;
            case SATK_ACT_SIGMOID: SATK_EPI(SATK_ACT_SIGMOID); break;
            default: SATK_EPI(SATK_ACT_NONE); break;
          }
#undef SATK_EPI
        }
      }
      if (tid == 128 && c0 == 0) { TC_MARK(11) }
      if (tid == 128 && c0 == 32) { TC_MARK(12) }
    }
  }
  if (tid == 128) { TC_MARK(7) }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) { TC_MARK(8) }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// rank-2 map over a K-contiguous matrix [rows, K] (row stride ld floats), box BK x BM, 128-byte swizzle
// rank-2 map over the output C [M, N] (row stride ldc floats), box 32 columns x 32 rows, 128-byte swizzle (epilogue TMA stores)
static bool make_map_c(CUtensorMap* map, float* base, long long M, long long N, long long ldc) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ldc * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static bool make_map(CUtensorMap* map, const float* base, long long rows, long long K, long long ld, int rank3_taps, long long tap_stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)rank3_taps};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)tap_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const int rank = rank3_taps > 0 ? 3 : 2;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// rank-2 map over a row-major [K, MN] operand (the reduction index is the row): box = 32 columns (128 B) x BK rows, 128-byte swizzle
// with 32-byte atoms (the layout the MN-major tf32 descriptors read)
static bool make_map_mn(CUtensorMap* map, const float* base, long long krows, long long mn, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)krows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int tc_trace(long long* out16) {
#ifdef SATK_PHASE_TIMING
  SATK_CUDA(cudaMemcpyFromSymbol(out16, satk::g_phase, sizeof(long long) * 16));
#endif
  return 0;
}

}  // namespace tc

// Convolution bank: every width of the bank as a z-batch of ONE launch (see satk_gemm_desc.bank_widths)
static int gemm_tc_bank_launch(const satk_gemm_desc* d, cudaStream_t st, bool* supported) {
  using namespace tc;
  const int W = d->bank_widths;
  SATK_CHECK_ARG(W <= 64 && d->transA == 0 && d->transB == 1 && !d->bias && d->act == 0 && !d->residual && !d->keep_mask &&
                     (d->beta == 0.0f || d->beta == 1.0f) && d->batch1 <= 1 && d->batch2 <= 1 && d->split_k <= 1 && d->seq_len == 0,
                 "satk_gemm: a conv bank takes K-contiguous operands, no epilogue options and beta in {0, 1}");
  SATK_CHECK_ARG((d->lda % 4) == 0 && (d->ldb % 4) == 0 && (d->K % 4) == 0 && ((uintptr_t)d->A % 16) == 0 && ((uintptr_t)d->B % 16) == 0 &&
                     (d->sBtap % 4) == 0 && (d->ldc % 4) == 0 && ((uintptr_t)d->C % 16) == 0,
                 "satk_gemm: conv bank operands must be 16-byte addressable");
  const long long a_cols = (long long)d->K + (long long)(W - 1) * d->bank_a_kstep;       // widest reduction coordinate read from A
  const long long c_cols = (long long)d->N + (long long)(W - 1) * d->bank_c_nstep;
  CUtensorMap mapA, mapB, mapC;
  if (!make_map(&mapA, d->A, d->M, a_cols, d->lda, 0, 0) || !make_map(&mapB, d->B, d->N, d->K, d->ldb, W * (W + 1) / 2, d->sBtap) ||
      !make_map_c(&mapC, d->C, d->M, c_cols, d->ldc)) {
    set_error("satk_gemm: cuTensorMapEncodeTiled failed for the conv bank");
    return SATK_ERR_CUDA;
  }
  *supported = true;
  Params p = {};
  p.prec = d->precision == 1 ? 1 : 0;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.taps = 1; p.shift0 = 0; p.tap_dir = d->tap_dir;
  p.C = d->C; p.ldc = d->ldc;
  p.alpha = d->alpha; p.beta = d->beta;
  p.split_k = 1;
  p.zbatch = W; p.bank = 1; p.bank_a_kstep = d->bank_a_kstep; p.bank_c_nstep = d->bank_c_nstep;
  p.tma_store = d->beta == 0.0f ? 1 : 2;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  SATK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(d->M, BM), ceil_div(d->N, BN), W);
  gemm_tc_kernel<<<grid, NTHREADS, smem, st>>>(mapA, mapB, mapC, p, 1);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

// Batched self-attention products (see satk_gemm_desc.zcoord): QK^T, P.V and their four gradient products as z-batches over
// shared 2-D operand views; outputs leave through TMA stores, which clip the ragged last tile of each entry.
static int gemm_tc_zcoord_launch(const satk_gemm_desc* d, cudaStream_t st, bool* supported) {
  using namespace tc;
  const int batches = (d->batch1 < 1 ? 1 : d->batch1) * (d->batch2 < 1 ? 1 : d->batch2);
  const int taps = d->taps < 1 ? 1 : d->taps;
  // operand storage: transA = 0 -> A is [M-index rows, reduction columns] (K-major), 1 -> [reduction rows, M-index columns] (MN-major);
  // transB = 1 -> B is [N-index rows, reduction columns] (K-major), 0 -> [reduction rows, N-index columns] (MN-major).  The per-entry
  // shifts keep their meaning (za_row / zb_row: M / N index, za_k / zb_k: reduction index) whichever way an operand is stored.
  const bool a_mn = d->transA == 1, b_mn = d->transB == 0;
  SATK_CHECK_ARG(taps == 1 && d->seq_len == 0 && !d->bias && d->act == 0 && !d->residual &&
                     !d->keep_mask && d->beta == 0.0f && d->split_k <= 1 && d->causal_skip >= 0 && d->causal_skip <= 3,
                 "satk_gemm: the zcoord form takes no epilogue options, beta = 0");
  SATK_CHECK_ARG(d->a_rows > 0 && d->a_cols > 0 && d->b_rows > 0 && d->b_cols > 0 && (d->sC1 != 0 || d->c_cols > 0),
                 "satk_gemm: the zcoord form needs the extents of the operand views");
  SATK_CHECK_ARG((d->lda % 4) == 0 && (d->ldb % 4) == 0 && (d->ldc % 4) == 0 && (d->sC1 % 4) == 0 && ((uintptr_t)d->A % 16) == 0 &&
                     ((uintptr_t)d->B % 16) == 0 && ((uintptr_t)d->C % 16) == 0 && ((a_mn ? d->za_row : d->za_k) % 4) == 0 &&
                     ((b_mn ? d->zb_row : d->zb_k) % 4) == 0 && (d->zc_col % 4) == 0,
                 "satk_gemm: zcoord operands must be 16-byte addressable");
  SATK_CHECK_ARG(d->causal_skip == 0 || (d->causal_skip == 1 ? d->M == d->N : d->M == d->K),
                 "satk_gemm: causal_skip needs a square (query, key) index pair");
  CUtensorMap mapA, mapB, mapC;
  bool ok = (a_mn ? make_map_mn(&mapA, d->A, d->a_rows, d->a_cols, d->lda) : make_map(&mapA, d->A, d->a_rows, d->a_cols, d->lda, 0, 0)) &&
            (b_mn ? make_map_mn(&mapB, d->B, d->b_rows, d->b_cols, d->ldb) : make_map(&mapB, d->B, d->b_rows, d->b_cols, d->ldb, 0, 0));
  if (ok) {
    if (d->sC1 != 0) {
      EncodeTiledFn enc = get_encode();
      cuuint64_t dims[3] = {(cuuint64_t)d->N, (cuuint64_t)d->M, (cuuint64_t)batches};
      cuuint64_t strides[2] = {(cuuint64_t)d->ldc * 4, (cuuint64_t)d->sC1 * 4};
      cuuint32_t box[3] = {32, 32, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      ok = enc && enc(&mapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d->C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    } else {
      ok = make_map_c(&mapC, d->C, d->M, d->c_cols, d->ldc);
    }
  }
  if (!ok) {
    set_error("satk_gemm: cuTensorMapEncodeTiled failed for the zcoord form");
    return SATK_ERR_CUDA;
  }
  *supported = true;
  Params p = {};
  p.prec = d->precision == 1 ? 1 : 0;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.taps = 1; p.shift0 = 0; p.tap_dir = 1;
  p.C = d->C; p.ldc = d->ldc;
  p.alpha = d->alpha; p.beta = 0.0f;
  p.split_k = 1;
  p.zbatch = batches;
  p.zcoord = 1; p.za_row = d->za_row; p.za_k = d->za_k; p.zb_row = d->zb_row; p.zb_k = d->zb_k; p.zc_col = d->zc_col;
  p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.causal = d->causal_skip;
  p.c_rank3 = d->sC1 != 0 ? 1 : 0;
  p.tma_store = 1;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  SATK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(d->M, BM), ceil_div(d->N, BN), batches);
  gemm_tc_kernel<<<grid, NTHREADS, smem, st>>>(mapA, mapB, mapC, p, 0);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

// Shapes served by the tensor-core tile: plain or tapped (conv) products with K-contiguous operands.
int gemm_tc_launch(const satk_gemm_desc* d, cudaStream_t st, bool* supported) {
  using namespace tc;
  *supported = false;
  const int batches = (d->batch1 < 1 ? 1 : d->batch1) * (d->batch2 < 1 ? 1 : d->batch2);
  const int taps = d->taps < 1 ? 1 : d->taps;
  // z-batches are served in one form only: same A and B for every entry, A read with a per-entry shift of the reduction
  // coordinate, one output per entry (conv weight gradients)
  if (d->bank_widths > 0) return gemm_tc_bank_launch(d, st, supported);
  if (d->zcoord) return gemm_tc_zcoord_launch(d, st, supported);
  const bool zform = batches > 1 && (d->batch2 <= 1) && d->sA1 == 0 && d->sB1 == 0 && taps == 1 && !d->bias && d->act == 0 &&
                     !d->residual && !d->keep_mask;
  // weight-gradient form: A [K, M] and B [K, N] row-major (transA, !transB), reduction over the rows, no epilogue options; the
  // SIMT tile's shift0 / shift_per_batch1 (row shift of A per z-batch) are shifts of the reduction coordinate here
  const bool mn = d->transA == 1 && d->transB == 0 && taps == 1 && !d->bias && d->act == 0 && !d->residual && !d->keep_mask &&
                  (batches == 1 || zform) && getenv("SATK_TC_MN") == nullptr;
  if ((batches != 1 && !zform) || (!mn && (d->transA != 0 || d->transB != 1)) || d->causal_skip != 0 || d->seq_len != 0 ||
      (!mn && d->shift_per_batch1 != 0))
    return SATK_OK;
  if (d->M < 64 || d->N < 48 || d->K < 32) return SATK_OK;                       // tiny problems stay on the SIMT tile
  if ((d->lda % 4) || (d->ldb % 4) || (!mn && (d->K % 4)) || ((uintptr_t)d->A % 16) || ((uintptr_t)d->B % 16)) return SATK_OK;
  if (mn && (d->beta != 0.0f && d->beta != 1.0f)) return SATK_OK;
  if (taps > 1 && (d->sBtap % 4)) return SATK_OK;
  if (d->keep_mask && d->split_k > 1) return SATK_OK;
  CUtensorMap mapA, mapB;
  if (mn) {
    if (!make_map_mn(&mapA, d->A, d->K, d->M, d->lda) || !make_map_mn(&mapB, d->B, d->K, d->N, d->ldb)) return SATK_OK;
  } else {
    if (!make_map(&mapA, d->A, d->M, d->K, d->lda, 0, 0)) return SATK_OK;
    if (!make_map(&mapB, d->B, d->N, d->K, d->ldb, taps > 1 ? taps : 0, d->sBtap)) return SATK_OK;
  }
  *supported = true;
  Params p;
  p.a_mn = p.b_mn = mn ? 1 : 0;
  p.prec = d->precision == 1 ? 1 : 0;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.taps = taps; p.shift0 = d->shift0; p.tap_dir = d->tap_dir;
  p.C = d->C; p.ldc = d->ldc;
  p.alpha = d->alpha; p.beta = d->beta;
  p.bias = d->bias; p.act = d->act;
  p.residual = d->residual; p.ldres = d->ldres;
  p.keep_mask = d->keep_mask; p.keep_scale = d->keep_scale;
  p.split_k = d->split_k < 1 ? 1 : d->split_k;
  p.bank = 0; p.bank_a_kstep = 0; p.bank_c_nstep = 0;
  p.zcoord = 0; p.za_row = p.za_k = p.zb_row = p.zb_k = p.zc_col = 0; p.causal = 0; p.c_rank3 = 0;
  p.zbatch = batches; p.kshift0 = d->kshift0; p.kshift_step = d->kshift_per_batch1; p.c_zstride = d->sC1;
  if (mn) {
    p.kshift0 += d->shift0; p.kshift_step += d->shift_per_batch1;
    p.shift0 = 0; p.tap_dir = 0;
  }
  const int kblocks = (d->K + BK - 1) / BK;
  const int tiles = ceil_div(d->M, BM) * ceil_div(d->N, BN);
  const bool linear_epi = !d->bias && d->act == 0 && !d->residual && !d->keep_mask && (d->beta == 0.0f || d->beta == 1.0f);
  if (batches == 1 && p.split_k == 1 && linear_epi && tiles < 74 && kblocks * taps >= 16) {
    // too few tiles for 148 SMs: split K, accumulate with atomics into a zeroed C
    int sk = (148 + tiles - 1) / tiles;
    if (sk > kblocks * taps / 4) sk = kblocks * taps / 4;
    if (sk > 1) {
      // beta == 1: the partial sums are added straight onto the existing C; beta == 0: onto a zeroed C
      if (d->beta == 0.0f) SATK_CUDA(cudaMemset2DAsync(d->C, (size_t)d->ldc * 4, 0, (size_t)d->N * 4, (size_t)d->M, st));
      p.split_k = sk;
    }
  }
  if (p.split_k > kblocks * taps) p.split_k = kblocks * taps;
  if (d->split_k > 1 && p.split_k == 1) p.beta = 1.0f;   // split-K semantics = accumulate onto C
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    SATK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  // plain overwrite epilogues (bias + activation at most) leave through TMA stores when C is 16-byte addressable
  CUtensorMap mapC = mapA;
  p.tma_store = 0;
  if ((d->ldc % 4) == 0 && ((uintptr_t)d->C % 16) == 0 && getenv("SATK_NO_TMA_STORE") == nullptr) {
    int mode = 0;
    if (p.split_k > 1) mode = 2;                                                          // atomically added partial sums
    else if (!d->residual && !d->keep_mask && p.beta == 0.0f) mode = 1;                   // overwrite
    else if (!d->residual && !d->keep_mask && p.beta == 1.0f && d->act == 0) mode = 2;    // C += alpha*A.B (+bias)
    // z-batched outputs are addressed as extra rows of one map: needs whole tiles per entry and densely stacked outputs
    if (batches > 1 && (d->M % BM != 0 || d->sC1 != (long long)d->M * d->ldc)) mode = 0;
    if (mode && make_map_c(&mapC, d->C, (long long)d->M * batches, d->N, d->ldc)) p.tma_store = mode;
    else mapC = mapA;
  }
  dim3 grid(ceil_div(d->M, BM), ceil_div(d->N, BN), p.split_k * batches);
  gemm_tc_kernel<<<grid, NTHREADS, smem, st>>>(mapA, mapB, mapC, p, taps > 1 ? 1 : 0);
  SATK_LAUNCH_CHECK();
  return SATK_OK;
}

}  // namespace satk
