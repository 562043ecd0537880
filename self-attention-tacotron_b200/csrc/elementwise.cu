// HBM-bound kernels: embedding, batch-norm (+ReLU, +max-pool), highway combine, activation
// backward, column sums, masks, softmax rows, teacher inputs, losses, optimiser.
// Grid sizes are multiples of the SM count (148) where the problem is large enough; reductions use
// warp shuffles.
#include <initializer_list>
#include <stdint.h>
#include "common.cuh"

namespace satk {

constexpr int kSMs = 148;
static inline int grid_for(long long n, int per_block) {
  long long g = (n + per_block - 1) / per_block;
  long long cap = (long long)kSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------- embedding
__global__ void embedding_fwd_k(const long long* __restrict__ ids, int rows, int offset, const float* __restrict__ table,
                                int dim, float* __restrict__ out) {
  int dv = dim >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)rows * dv; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / dv), c = (int)(i % dv);
    long long id = ids[r] - offset;
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(table + id * dim) + c);
  }
}
__global__ void embedding_bwd_k(const long long* __restrict__ ids, int rows, int offset, const float* __restrict__ dout,
                                int dim, float* __restrict__ dtable) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)rows * dim; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / dim), c = (int)(i % dim);
    long long id = ids[r] - offset;
    atomicAdd(dtable + id * dim + c, dout[i]);
  }
}

// ---------------------------------------------------------------- batch norm
// stats: grid (C/32, RS) ; block (32, 8): each block reduces a row slice for 32 channels, atomics into sum/sumsq
__global__ void bn_partial_k(const float* __restrict__ x, long long ldx, int rows, int C, float* __restrict__ sum,
                             float* __restrict__ sumsq, const float* __restrict__ shift) {
  __shared__ float s1[8][33], s2[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  float sh = (c < C && shift) ? shift[c] : 0.f;
  if (c < C) {
    for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) {
      float v = __ldg(x + (long long)r * ldx + c) - sh;
      a += v;
      b += v * v;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    if (sum) atomicAdd(sum + c, a);
    if (sumsq) atomicAdd(sumsq + c, b);
  }
}
// two-pass variance: pass 1 mean (shift=NULL), pass 2 centered second moment (shift=mean)
__global__ void bn_mean_finish_k(float* sum, int C, int rows, float* mean) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) mean[c] = sum[c] / rows;
}
__global__ void bn_var_finish_k(int C, int rows, const float* mean, float* var,
                                float* mov_mean, float* mov_var, float mom, int bessel) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    // var holds the centered second moment sum((x-mean)^2)
    float v = var[c] / rows;
    var[c] = v;
    if (mov_mean) {
      float vb = bessel ? v * ((float)rows / (float)max(rows - 1, 1)) : v;
      mov_mean[c] = mom * mov_mean[c] + (1.f - mom) * mean[c];
      mov_var[c] = mom * mov_var[c] + (1.f - mom) * vb;
    }
  }
}

__global__ void bn_apply_k(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ mean,
                           const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float eps, int act, const float* __restrict__ residual, int mp_len, int ps, float* __restrict__ y, long long ldy) {
  long long n = (long long)rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / C), c = (int)(i % C);
    float inv = rsqrtf(var[c] + eps) * gamma[c];
    float sh = beta[c] - mean[c] * inv;
    float v = apply_act(fmaf(__ldg(x + (long long)r * ldx + c), inv, sh), act);
    if (mp_len > 0 && ((r / ps) % mp_len) != mp_len - 1) {
      float v2 = apply_act(fmaf(__ldg(x + (long long)(r + ps) * ldx + c), inv, sh), act);
      v = fmaxf(v, v2);
    }
    if (residual) v += residual[(long long)r * C + c];
    y[(long long)r * ldy + c] = v;
  }
}

// dz (gradient wrt the BN output z, before act/maxpool) given dy; returns value for (r,c)
__device__ __forceinline__ float bn_dz(const float* __restrict__ x, long long ldx, int r, int c, float inv, float sh, int act,
                                       int mp_len, int ps, const float* __restrict__ dy, long long lddy) {
  float z = fmaf(__ldg(x + (long long)r * ldx + c), inv, sh);
  float a = apply_act(z, act);
  float g = 0.f;
  if (mp_len > 0) {
    int t = (r / ps) % mp_len;
    // y[t] = max(a[t], a[t+1]) (t<T-1), y[T-1] = a[T-1];  a[t] receives from y[t] if a[t] >= a[t+1], and from y[t-1] if a[t] > a[t-1]
    if (t == mp_len - 1) {
      g += dy[(long long)r * lddy + c];
    } else {
      float an = apply_act(fmaf(__ldg(x + (long long)(r + ps) * ldx + c), inv, sh), act);
      if (a >= an) g += dy[(long long)r * lddy + c];
    }
    if (t > 0) {
      float ap = apply_act(fmaf(__ldg(x + (long long)(r - ps) * ldx + c), inv, sh), act);
      if (a > ap) g += dy[(long long)(r - ps) * lddy + c];
    }
  } else {
    g = dy[(long long)r * lddy + c];
  }
  if (act == SATK_ACT_RELU) g = (z > 0.f) ? g : 0.f;
  else if (act == SATK_ACT_TANH) g *= (1.f - a * a);
  else if (act == SATK_ACT_SIGMOID) g *= a * (1.f - a);
  return g;
}
// pass 1: per-channel sums of dz and dz*xhat
__global__ void bn_bwd_reduce_k(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int act, int mp_len, int ps, const float* __restrict__ dy, long long lddy,
                                float* __restrict__ sum_dz, float* __restrict__ sum_dzx) {
  __shared__ float s1[8][33], s2[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (c < C) {
    float rstd = rsqrtf(var[c] + eps);
    float inv = rstd * gamma[c];
    float sh = beta[c] - mean[c] * inv;
    float mu = mean[c];
    for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) {
      float dz = bn_dz(x, ldx, r, c, inv, sh, act, mp_len, ps, dy, lddy);
      float xh = (__ldg(x + (long long)r * ldx + c) - mu) * rstd;
      a += dz;
      b += dz * xh;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    atomicAdd(sum_dz + c, a);
    atomicAdd(sum_dzx + c, b);
  }
}
__global__ void bn_bwd_apply_k(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ mean,
                               const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                               float eps, int act, int mp_len, int ps, int batch_stats, const float* __restrict__ dy, long long lddy,
                               const float* __restrict__ sum_dz, const float* __restrict__ sum_dzx, float* __restrict__ dx,
                               long long lddx) {
  long long n = (long long)rows * C;
  float invn = 1.f / rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / C), c = (int)(i % C);
    float rstd = rsqrtf(var[c] + eps);
    float inv = rstd * gamma[c];
    float sh = beta[c] - mean[c] * inv;
    float dz = bn_dz(x, ldx, r, c, inv, sh, act, mp_len, ps, dy, lddy);
    float g;
    if (batch_stats) {
      float xh = (__ldg(x + (long long)r * ldx + c) - mean[c]) * rstd;
      g = inv * (dz - invn * sum_dz[c] - xh * invn * sum_dzx[c]);
    } else {
      g = inv * dz;
    }
    dx[(long long)r * lddx + c] = g;
  }
}
// ---- 16-byte variants (C, every leading dimension and every base pointer multiples of 4 floats): a thread owns 4 consecutive
// channels; the per-channel constants are hoisted when the grid stride keeps the channel quad fixed.
struct BnQuad { float inv[4], sh[4], mu[4], rstd[4]; };
__device__ __forceinline__ void f4_to(const float4 v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ BnQuad bn_quad(const float* mean, const float* var, const float* gamma, const float* beta, float eps, int c) {
  BnQuad q;
  float m[4], v[4], g[4], b[4];
  f4_to(ldg4(mean + c), m); f4_to(ldg4(var + c), v); f4_to(ldg4(gamma + c), g); f4_to(ldg4(beta + c), b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    q.rstd[i] = rsqrtf(v[i] + eps);
    q.inv[i] = q.rstd[i] * g[i];
    q.sh[i] = b[i] - m[i] * q.inv[i];
    q.mu[i] = m[i];
  }
  return q;
}
// dz of 4 channels of row r (see bn_dz); xr receives the row's x values
__device__ __forceinline__ void bn_dz4(const float* __restrict__ x, long long ldx, int r, int c, const BnQuad& q, int act, int mp_len,
                                       int ps, const float* __restrict__ dy, long long lddy, float* dz, float* xr) {
  float z[4], a[4], g[4];
  f4_to(ldg4(x + (long long)r * ldx + c), xr);
#pragma unroll
  for (int i = 0; i < 4; ++i) { z[i] = fmaf(xr[i], q.inv[i], q.sh[i]); a[i] = apply_act(z[i], act); g[i] = 0.f; }
  if (mp_len > 0) {
    const int t = (r / ps) % mp_len;
    float d0[4];
    f4_to(ldg4(dy + (long long)r * lddy + c), d0);
    if (t == mp_len - 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = d0[i];
    } else {
      float xn[4];
      f4_to(ldg4(x + (long long)(r + ps) * ldx + c), xn);
#pragma unroll
      for (int i = 0; i < 4; ++i) if (a[i] >= apply_act(fmaf(xn[i], q.inv[i], q.sh[i]), act)) g[i] = d0[i];
    }
    if (t > 0) {
      float xp[4], dp[4];
      f4_to(ldg4(x + (long long)(r - ps) * ldx + c), xp);
      f4_to(ldg4(dy + (long long)(r - ps) * lddy + c), dp);
#pragma unroll
      for (int i = 0; i < 4; ++i) if (a[i] > apply_act(fmaf(xp[i], q.inv[i], q.sh[i]), act)) g[i] += dp[i];
    }
  } else {
    f4_to(ldg4(dy + (long long)r * lddy + c), g);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (act == SATK_ACT_RELU) g[i] = (z[i] > 0.f) ? g[i] : 0.f;
    else if (act == SATK_ACT_TANH) g[i] *= (1.f - a[i] * a[i]);
    else if (act == SATK_ACT_SIGMOID) g[i] *= a[i] * (1.f - a[i]);
    dz[i] = g[i];
  }
}
// grid (C/128, RS), block (32, 8): thread = channel quad x row slice
__global__ void bn_bwd_reduce4_k(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ mean,
                                 const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, int act, int mp_len, int ps, const float* __restrict__ dy, long long lddy,
                                 float* __restrict__ sum_dz, float* __restrict__ sum_dzx) {
  __shared__ float4 s1[8][33], s2[8][33];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    const BnQuad q = bn_quad(mean, var, gamma, beta, eps, c);
    for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) {
      float dz[4], xr[4];
      bn_dz4(x, ldx, r, c, q, act, mp_len, ps, dy, lddy, dz, xr);
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] += dz[i]; b[i] = fmaf(dz[i], (xr[i] - q.mu[i]) * q.rstd[i], b[i]); }
    }
  }
  s1[threadIdx.y][threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
  s2[threadIdx.y][threadIdx.x] = make_float4(b[0], b[1], b[2], b[3]);
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) {
      const float4 u = s1[i][threadIdx.x], v = s2[i][threadIdx.x];
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      b[0] += v.x; b[1] += v.y; b[2] += v.z; b[3] += v.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { atomicAdd(sum_dz + c + i, a[i]); atomicAdd(sum_dzx + c + i, b[i]); }
  }
}
// thread = channel quad, rows by grid stride (a whole number of rows per stride: the quad of a thread is fixed)
__global__ void bn_bwd_apply4_k(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int act, int mp_len, int ps, int batch_stats, const float* __restrict__ dy, long long lddy,
                                const float* __restrict__ sum_dz, const float* __restrict__ sum_dzx, float* __restrict__ dx,
                                long long lddx) {
  const int C4 = C >> 2;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;     // nt is a multiple of C4
  const int c = (gt % C4) * 4;
  const BnQuad q = bn_quad(mean, var, gamma, beta, eps, c);
  const float invn = 1.f / rows;
  float k1[4], k2[4];
  f4_to(ldg4(sum_dz + c), k1); f4_to(ldg4(sum_dzx + c), k2);
#pragma unroll
  for (int i = 0; i < 4; ++i) { k1[i] *= invn; k2[i] *= invn; }
  for (int r = gt / C4; r < rows; r += nt / C4) {
    float dz[4], xr[4], g[4];
    bn_dz4(x, ldx, r, c, q, act, mp_len, ps, dy, lddy, dz, xr);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      g[i] = batch_stats ? q.inv[i] * (dz[i] - k1[i] - (xr[i] - q.mu[i]) * q.rstd[i] * k2[i]) : q.inv[i] * dz[i];
    *reinterpret_cast<float4*>(dx + (long long)r * lddx + c) = make_float4(g[0], g[1], g[2], g[3]);
  }
}
__global__ void bn_apply4_k(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ mean,
                            const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                            float eps, int act, const float* __restrict__ residual, int mp_len, int ps, float* __restrict__ y, long long ldy) {
  const int C4 = C >> 2;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;     // nt is a multiple of C4
  const int c = (gt % C4) * 4;
  const BnQuad q = bn_quad(mean, var, gamma, beta, eps, c);
  for (int r = gt / C4; r < rows; r += nt / C4) {
    float xr[4], v[4];
    f4_to(ldg4(x + (long long)r * ldx + c), xr);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = apply_act(fmaf(xr[i], q.inv[i], q.sh[i]), act);
    if (mp_len > 0 && ((r / ps) % mp_len) != mp_len - 1) {
      float xn[4];
      f4_to(ldg4(x + (long long)(r + ps) * ldx + c), xn);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], apply_act(fmaf(xn[i], q.inv[i], q.sh[i]), act));
    }
    if (residual) {
      float rr[4];
      f4_to(ldg4(residual + (long long)r * C + c), rr);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += rr[i];
    }
    *reinterpret_cast<float4*>(y + (long long)r * ldy + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// launch geometry of the quad kernels: 256-thread blocks, a whole number of rows per grid stride, about 8 blocks per SM at most
static inline bool bn_quad_ok(int C, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  if (C % 4 != 0 || (C / 4) > 256 * 1184 || (256 % (C / 4) != 0 && (C / 4) % 256 != 0)) return false;
  for (long long l : lds) if (l % 4 != 0) return false;
  for (const void* q : ptrs) if (q && ((uintptr_t)q & 15)) return false;
  return true;
}
static inline int bn_quad_grid(int rows, int C) {
  const int C4 = C / 4;
  const long long quads = (long long)rows * C4;
  const int unit = C4 > 256 ? C4 / 256 : 1;            // blocks per row when a row is wider than a block
  long long blocks = (quads + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  blocks = (blocks + unit - 1) / unit * unit;
  return (int)blocks;
}

__global__ void bn_bwd_param_k(const float* sum_dz, const float* sum_dzx, int C, float* dgamma, float* dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    dgamma[c] += sum_dzx[c];
    dbeta[c] += sum_dz[c];
  }
}

// ---------------------------------------------------------------- highway / activations / misc
__global__ void highway_fwd_k(const float* __restrict__ H, const float* __restrict__ T, const float* __restrict__ x,
                              float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = T[i];
    y[i] = H[i] * t + x[i] * (1.f - t);
  }
}
__global__ void highway_bwd_k(const float* __restrict__ H, const float* __restrict__ T, const float* __restrict__ x,
                              const float* __restrict__ dy, float* __restrict__ dH, float* __restrict__ dT,
                              float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = T[i], h = H[i], g = dy[i];
    dH[i] = (h > 0.f) ? g * t : 0.f;
    dT[i] = g * (h - x[i]) * t * (1.f - t);
    dx[i] = g * (1.f - t);
  }
}
__global__ void act_bwd_k(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz, long long n,
                          int act, const uint8_t* __restrict__ mask, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) { dz[i] = 0.f; continue; }
    // y = act(z) * scale (dropout scaling of kept units); dropped units have y == 0
    float g = dy[i] * scale;
    float yv = y[i] / scale;
    if (act == SATK_ACT_RELU) g = (yv > 0.f) ? g : 0.f;
    else if (act == SATK_ACT_TANH) g *= (1.f - yv * yv);
    else if (act == SATK_ACT_SIGMOID) g *= yv * (1.f - yv);
    dz[i] = g;
  }
}
__global__ void mask_scale_k(const float* __restrict__ x, const uint8_t* __restrict__ mask, float scale, float* __restrict__ y, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const uchar4 m = reinterpret_cast<const uchar4*>(mask)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(m.x ? v.x * scale : 0.f, m.y ? v.y * scale : 0.f, m.z ? v.z * scale : 0.f,
                                                  m.w ? v.w * scale : 0.f);
  }
}
__global__ void colsum_k(const float* __restrict__ x, long long ldx, int rows, int C, float* __restrict__ out) {
  __shared__ float s1[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < C)
    for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) a += __ldg(x + (long long)r * ldx + c);
  s1[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) a += s1[i][threadIdx.x];
    atomicAdd(out + c, a);
  }
}
__global__ void add_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = a[i] + b[i];
}
__global__ void axpy_k(float alpha, const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += alpha * x[i];
}
__global__ void transpose_k(const float* __restrict__ x, int rows, int cols, float* __restrict__ y) {
  __shared__ float t[32][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) t[j][threadIdx.x] = x[(long long)r * cols + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) y[(long long)c2 * rows + r2] = t[threadIdx.x][j];
  }
}
// strided variant: y[c * ldy + r] = x[r * ldx + c]
__global__ void transpose_strided_k(const float* __restrict__ x, long long ldx, int rows, int cols, float* __restrict__ y,
                                    long long ldy) {
  __shared__ float t[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) t[j][threadIdx.x] = x[(long long)r * ldx + c];
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) y[(long long)c2 * ldy + r2] = t[threadIdx.x][j];
  }
}
// many small transposes in one launch: desc[3*i] = {offset, rows, cols}; dst[off + c*rows + r] = src[off + r*cols + c]
__global__ void transpose_batched_k(const float* __restrict__ src, float* __restrict__ dst, const int* __restrict__ desc) {
  __shared__ float t[32][33];
  const int off = desc[3 * blockIdx.y], rows = desc[3 * blockIdx.y + 1], cols = desc[3 * blockIdx.y + 2];
  const int tx_n = (cols + 31) / 32, ty_n = (rows + 31) / 32;
  for (int tile = blockIdx.x; tile < tx_n * ty_n; tile += gridDim.x) {
    const int bx = tile % tx_n, by = tile / tx_n;
    const int c = bx * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int r = by * 32 + j;
      if (r < rows && c < cols) t[j][threadIdx.x] = src[off + (long long)r * cols + c];
    }
    __syncthreads();
    const int r2 = by * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int c2 = bx * 32 + j;
      if (r2 < rows && c2 < cols) dst[off + (long long)c2 * rows + r2] = t[threadIdx.x][j];
    }
    __syncthreads();
  }
}
__global__ void mask_rows_k(const float* __restrict__ x, const long long* __restrict__ len, int B, int T, int C, int tm,
                            float* __restrict__ y) {
  long long n = (long long)B * T * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / C;
    int b, t;
    if (tm) { t = (int)(row / B); b = (int)(row % B); } else { b = (int)(row / T); t = (int)(row % T); }
    y[i] = (t < len[b]) ? x[i] : 0.f;
  }
}
__global__ void softsign_fwd_k(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    y[i] = v / (1.f + fabsf(v));
  }
}
__global__ void softsign_bwd_k(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = 1.f + fabsf(x[i]);
    dx[i] = dy[i] / (d * d);
  }
}
__global__ void add_rowvec_tb_k(float* __restrict__ y, const float* __restrict__ v, int T, int B, int C) {
  long long n = (long long)T * B * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long bc = i % ((long long)B * C);
    y[i] += v[bc];
  }
}
__global__ void sum_over_t_k(const float* __restrict__ dy, int T, int B, int C, float* __restrict__ dv) {
  long long bc = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (bc < (long long)B * C) {
    float a = 0.f;
    for (int t = 0; t < T; ++t) a += dy[(long long)t * B * C + bc];
    dv[bc] = a;
  }
}
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
__global__ void bernoulli_k(uint8_t* __restrict__ out, long long n, float keep, unsigned long long seed) {
  uint32_t thr = (keep >= 1.f) ? 0xffffffffu : (uint32_t)(keep * 4294967296.0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = hash32(seed * 0x9e3779b97f4a7c15ULL + (uint64_t)i) < thr ? 1 : 0;
}
// the same with the step's seed read from device memory (a captured CUDA graph replays the launch with a new seed every step)
__global__ void bernoulli_dev_k(uint8_t* __restrict__ out, long long n, float keep, const unsigned long long* __restrict__ seed_dev,
                                unsigned long long salt) {
  const unsigned long long seed = seed_dev[0] * 1000003ULL + salt;
  uint32_t thr = (keep >= 1.f) ? 0xffffffffu : (uint32_t)(keep * 4294967296.0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = hash32(seed * 0x9e3779b97f4a7c15ULL + (uint64_t)i) < thr ? 1 : 0;
}

// ---------------------------------------------------------------- softmax rows (T <= 1024)
// one warp per row; S row-major [nmat*T, T]
__global__ void softmax_fwd_k(float* __restrict__ S, int nmat, int T, int causal, const uint8_t* __restrict__ mask, float scale,
                              float* __restrict__ Pd) {
  int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  long long rows = (long long)nmat * T;
  for (long long row = blockIdx.x * (long long)warps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * warps) {
    int q = (int)(row % T);
    int len = causal ? q + 1 : T;
    float* s = S + row * T;
    float mx = -INFINITY;
    for (int j = lane; j < len; j += 32) mx = fmaxf(mx, s[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < len; j += 32) {
      float e = expf(s[j] - mx);
      s[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int j = lane; j < T; j += 32) {
      float p = (j < len) ? s[j] * inv : 0.f;
      s[j] = p;
      if (Pd) Pd[row * T + j] = mask ? (mask[row * T + j] ? p * scale : 0.f) : p;
    }
  }
}
__global__ void softmax_bwd_k(const float* __restrict__ P, const float* __restrict__ dPd, int nmat, int T, int causal,
                              const uint8_t* __restrict__ mask, float scale, float* __restrict__ dS) {
  int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  long long rows = (long long)nmat * T;
  for (long long row = blockIdx.x * (long long)warps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * warps) {
    int q = (int)(row % T);
    int len = causal ? q + 1 : T;
    const float* p = P + row * T;
    const float* g = dPd + row * T;
    float dot = 0.f;
    for (int j = lane; j < len; j += 32) {
      float gj = mask ? (mask[row * T + j] ? g[j] * scale : 0.f) : g[j];
      dot += gj * p[j];
    }
    dot = warp_sum(dot);
    for (int j = lane; j < T; j += 32) {
      float v = 0.f;
      if (j < len) {
        float gj = mask ? (mask[row * T + j] ? g[j] * scale : 0.f) : g[j];
        v = p[j] * (gj - dot);
      }
      dS[row * T + j] = v;
    }
  }
}

// ---------------------------------------------------------------- teacher inputs / losses
__global__ void teacher_inputs_k(const float* __restrict__ mel, int B, int Tm, int nm, int r, int nf, float* __restrict__ out) {
  int Td = Tm / r, W = nm * nf;
  long long n = (long long)Td * B * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    long long tb = i / W;
    int b = (int)(tb % B), t = (int)(tb / B);
    float v = 0.f;
    if (t > 0) {
      int f = w / nm, c = w % nm;                 // frame f of the fed frames, channel c
      int frame = (t - 1) * r + (r - nf) + f;     // last n_feed frames of group t-1
      v = mel[((long long)b * Tm + frame) * nm + c];
    }
    out[i] = v;
  }
}
// scratch4: [0]=sum |d| * m, [1]=nnz(spec mask), [2]=sum ce*m, [3]=nnz(bin mask)
__global__ void loss_reduce_k(const float* __restrict__ pred, const float* __restrict__ stop, const float* __restrict__ mel,
                              const float* __restrict__ done, const float* __restrict__ smask, const float* __restrict__ bmask,
                              int B, int Tm, int nm, int r, float* __restrict__ scratch) {
  __shared__ float sh[32];
  int Td = Tm / r;
  long long n = (long long)B * Tm * nm;
  float a = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % nm);
    long long bt = i / nm;
    int tm = (int)(bt % Tm), b = (int)(bt / Tm);
    int t = tm / r, f = tm % r;
    float p = pred[((long long)t * B + b) * (r * nm) + f * nm + c];
    a += fabsf(p - mel[i]) * smask[bt];
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0) atomicAdd(scratch + 0, a);
  float m1 = 0.f, ce = 0.f, m2 = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * Tm; i += (long long)gridDim.x * blockDim.x)
    m1 += (smask[i] != 0.f) ? 1.f : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * Td; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % Td), b = (int)(i / Td);
    float x = stop[(long long)t * B + b], z = done[i], w = bmask[i];
    ce += (fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)))) * w;
    m2 += (w != 0.f) ? 1.f : 0.f;
  }
  m1 = block_sum(m1, sh);
  if (threadIdx.x == 0) atomicAdd(scratch + 1, m1);
  ce = block_sum(ce, sh);
  if (threadIdx.x == 0) atomicAdd(scratch + 2, ce);
  m2 = block_sum(m2, sh);
  if (threadIdx.x == 0) atomicAdd(scratch + 3, m2);
}
__global__ void loss_grad_k(const float* __restrict__ pred, const float* __restrict__ stop, const float* __restrict__ mel,
                            const float* __restrict__ done, const float* __restrict__ smask, const float* __restrict__ bmask,
                            int B, int Tm, int nm, int r, const float* __restrict__ scratch, float* __restrict__ out3,
                            float* __restrict__ dpred, float* __restrict__ dstop) {
  int Td = Tm / r;
  float n1 = scratch[1] * nm, n2 = scratch[3];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float l1 = scratch[0] / n1, l2 = scratch[2] / n2;
    out3[0] = l1; out3[1] = l2; out3[2] = l1 + l2;
  }
  long long n = (long long)B * Tm * nm;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % nm);
    long long bt = i / nm;
    int tm = (int)(bt % Tm), b = (int)(bt / Tm);
    int t = tm / r, f = tm % r;
    long long pi = ((long long)t * B + b) * (r * nm) + f * nm + c;
    float dlt = pred[pi] - mel[i];
    float sg = (dlt > 0.f) ? 1.f : ((dlt < 0.f) ? -1.f : 0.f);
    dpred[pi] = sg * smask[bt] / n1;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * Td; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % Td), b = (int)(i / Td);
    float x = stop[(long long)t * B + b];
    dstop[(long long)t * B + b] = (sigmoidf_(x) - done[i]) * bmask[i] / n2;
  }
}

// ---------------------------------------------------------------- optimiser
// Deterministic global sum of squares: every block leaves its partial in the scratch, the last block to finish adds them in
// a fixed order (data-parallel replicas must derive bit-identical clip factors from bit-identical all-reduced gradients, or
// their weights drift apart).  scratch: [0] result, [1] ticket counter, [2 ..] one partial per block.
__global__ void sumsq_k(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  __shared__ bool last;
  float a = 0.f;
  long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(g4 + i);
    a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a += g[i] * g[i];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) {
    out[2 + blockIdx.x] = a;
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(out + 1), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(out + 2 + i);
    t = block_sum(t, sh);
    if (threadIdx.x == 0) out[0] = t;
  }
}
__global__ void l2_reg_k(const float* __restrict__ p, const float* __restrict__ mask, long long n, float scale, float* __restrict__ g,
                         float* __restrict__ loss_acc) {
  __shared__ float sh[32];
  float a = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = mask[i] * p[i];
    a += w * w;
    if (g) g[i] += scale * w;
  }
  if (loss_acc) {
    a = block_sum(a, sh);
    if (threadIdx.x == 0) atomicAdd(loss_acc, 0.5f * scale * a);
  }
}
__global__ void adam_clip_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, const float* __restrict__ sumsq, float gscale, float clip, float lr_t, float b1,
                            float b2, float eps) {
  // tf.clip_by_global_norm: g * clip / max(norm, clip); norm of the (already averaged) gradient
  float norm = sqrtf(sumsq[0]) * gscale;
  float sc = gscale * clip / fmaxf(norm, clip);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * sc;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace satk

using namespace satk;
#define ST ((cudaStream_t)stream)

extern "C" {

int satk_embedding_fwd(const long long* ids, int rows, int offset, const float* table, int dim, float* out, void* stream) {
  SATK_CHECK_ARG(dim % 4 == 0, "embedding dim %d not a multiple of 4", dim);
  embedding_fwd_k<<<grid_for((long long)rows * dim / 4, 256), 256, 0, ST>>>(ids, rows, offset, table, dim, out);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_embedding_bwd(const long long* ids, int rows, int offset, const float* dout, int dim, float* dtable, void* stream) {
  embedding_bwd_k<<<grid_for((long long)rows * dim, 256), 256, 0, ST>>>(ids, rows, offset, dout, dim, dtable);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_bn_stats(const float* x, long long ldx, int rows, int C, float* mean, float* var, float* mov_mean, float* mov_var,
                  float momentum, int bessel, void* stream) {
  // two passes: mean, then centered second moment (no E[x^2]-E[x]^2 cancellation)
  SATK_CUDA(cudaMemsetAsync(mean, 0, sizeof(float) * C, ST));
  SATK_CUDA(cudaMemsetAsync(var, 0, sizeof(float) * C, ST));
  int rs = min(64, max(1, rows / 64));
  dim3 grid(ceil_div(C, 32), rs), block(32, 8);
  bn_partial_k<<<grid, block, 0, ST>>>(x, ldx, rows, C, mean, nullptr, nullptr);
  bn_mean_finish_k<<<ceil_div(C, 128), 128, 0, ST>>>(mean, C, rows, mean);
  bn_partial_k<<<grid, block, 0, ST>>>(x, ldx, rows, C, nullptr, var, mean);
  bn_var_finish_k<<<ceil_div(C, 128), 128, 0, ST>>>(C, rows, mean, var, mov_mean, mov_var, momentum, bessel);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_bn_apply(const float* x, long long ldx, int rows, int C, const float* mean, const float* var, const float* gamma,
                  const float* beta, float eps, int act, const float* residual, int maxpool_seq_len, int pos_stride, float* y, long long ldy,
                  void* stream) {
  SATK_CHECK_ARG(maxpool_seq_len == 0 || (pos_stride >= 1 && rows % maxpool_seq_len == 0), "bn_apply: rows %d not a multiple of seq_len %d", rows, maxpool_seq_len);
  if (pos_stride < 1) pos_stride = 1;
  if (bn_quad_ok(C, {ldx, ldy}, {x, y, mean, var, gamma, beta, residual}))
    bn_apply4_k<<<bn_quad_grid(rows, C), 256, 0, ST>>>(x, ldx, rows, C, mean, var, gamma, beta, eps, act, residual, maxpool_seq_len,
                                                       pos_stride, y, ldy);
  else
    bn_apply_k<<<grid_for((long long)rows * C, 256), 256, 0, ST>>>(x, ldx, rows, C, mean, var, gamma, beta, eps, act, residual,
                                                                  maxpool_seq_len, pos_stride, y, ldy);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_bn_bwd(const float* x, long long ldx, int rows, int C, const float* mean, const float* var, const float* gamma,
                const float* beta, float eps, int act, int maxpool_seq_len, int pos_stride, int use_batch_stats, const float* dy, long long lddy,
                float* dx, long long lddx, float* dgamma, float* dbeta, float* scratch, void* stream) {
  if (pos_stride < 1) pos_stride = 1;
  SATK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * C, ST));
  int rs = min(64, max(1, rows / 64));
  dim3 grid(ceil_div(C, 32), rs), block(32, 8);
  if (bn_quad_ok(C, {ldx, lddy, lddx}, {x, dy, dx, mean, var, gamma, beta, scratch})) {
    dim3 grid4(ceil_div(C, 128), rs);
    bn_bwd_reduce4_k<<<grid4, block, 0, ST>>>(x, ldx, rows, C, mean, var, gamma, beta, eps, act, maxpool_seq_len, pos_stride, dy, lddy,
                                              scratch, scratch + C);
    bn_bwd_apply4_k<<<bn_quad_grid(rows, C), 256, 0, ST>>>(x, ldx, rows, C, mean, var, gamma, beta, eps, act, maxpool_seq_len, pos_stride,
                                                           use_batch_stats, dy, lddy, scratch, scratch + C, dx, lddx);
  } else {
    bn_bwd_reduce_k<<<grid, block, 0, ST>>>(x, ldx, rows, C, mean, var, gamma, beta, eps, act, maxpool_seq_len, pos_stride, dy, lddy, scratch,
                                            scratch + C);
    bn_bwd_apply_k<<<grid_for((long long)rows * C, 256), 256, 0, ST>>>(x, ldx, rows, C, mean, var, gamma, beta, eps, act,
                                                                      maxpool_seq_len, pos_stride, use_batch_stats, dy, lddy, scratch,
                                                                      scratch + C, dx, lddx);
  }
  bn_bwd_param_k<<<ceil_div(C, 128), 128, 0, ST>>>(scratch, scratch + C, C, dgamma, dbeta);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_highway_fwd(const float* H, const float* T, const float* x, float* y, long long n, void* stream) {
  highway_fwd_k<<<grid_for(n, 256), 256, 0, ST>>>(H, T, x, y, n);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_highway_bwd(const float* H, const float* T, const float* x, const float* dy, float* dHpre, float* dTpre, float* dx,
                     long long n, void* stream) {
  highway_bwd_k<<<grid_for(n, 256), 256, 0, ST>>>(H, T, x, dy, dHpre, dTpre, dx, n);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_act_bwd(const float* y, const float* dy, float* dz, long long n, int act, const uint8_t* keep_mask, float keep_scale,
                 void* stream) {
  act_bwd_k<<<grid_for(n, 256), 256, 0, ST>>>(y, dy, dz, n, act, keep_mask, keep_scale);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_mask_scale(const float* x, const uint8_t* keep_mask, float keep_scale, float* y, long long n, void* stream) {
  SATK_CHECK_ARG(n % 4 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && ((uintptr_t)keep_mask & 3) == 0,
                 "mask_scale: n=%lld and the buffers must be multiples of 4 elements / 16-byte aligned", n);
  mask_scale_k<<<grid_for(n / 4, 256), 256, 0, ST>>>(x, keep_mask, keep_scale, y, n / 4);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_colsum_acc(const float* x, long long ldx, int rows, int C, float* out, void* stream) {
  int rs = min(64, max(1, rows / 64));
  dim3 grid(ceil_div(C, 32), rs), block(32, 8);
  colsum_k<<<grid, block, 0, ST>>>(x, ldx, rows, C, out);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_add(const float* a, const float* b, float* out, long long n, void* stream) {
  add_k<<<grid_for(n, 256), 256, 0, ST>>>(a, b, out, n);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_axpy(float alpha, const float* x, float* y, long long n, void* stream) {
  axpy_k<<<grid_for(n, 256), 256, 0, ST>>>(alpha, x, y, n);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_transpose(const float* x, int rows, int cols, float* y, void* stream) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  transpose_k<<<grid, block, 0, ST>>>(x, rows, cols, y);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_transpose_strided(const float* x, long long ldx, int rows, int cols, float* y, long long ldy, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  SATK_CHECK_ARG(ldx >= cols && ldy >= rows, "transpose_strided: ldx=%lld < cols=%d or ldy=%lld < rows=%d", ldx, cols, ldy, rows);
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  transpose_strided_k<<<grid, block, 0, ST>>>(x, ldx, rows, cols, y, ldy);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_transpose_batched(const float* src, float* dst, const int* desc, int n, void* stream) {
  if (n <= 0) return 0;
  dim3 grid(64, n), block(32, 8);
  transpose_batched_k<<<grid, block, 0, ST>>>(src, dst, desc);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_mask_rows(const float* x, const long long* lengths, int B, int T, int C, int time_major, float* y, void* stream) {
  mask_rows_k<<<grid_for((long long)B * T * C, 256), 256, 0, ST>>>(x, lengths, B, T, C, time_major, y);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_softsign_fwd(const float* x, float* y, long long n, void* stream) {
  softsign_fwd_k<<<grid_for(n, 256), 256, 0, ST>>>(x, y, n);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_softsign_bwd(const float* x, const float* dy, float* dx, long long n, void* stream) {
  softsign_bwd_k<<<grid_for(n, 256), 256, 0, ST>>>(x, dy, dx, n);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_add_rowvec_tb(float* y, const float* v, int T, int B, int C, void* stream) {
  add_rowvec_tb_k<<<grid_for((long long)T * B * C, 256), 256, 0, ST>>>(y, v, T, B, C);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_sum_over_t(const float* dy, int T, int B, int C, float* dv, void* stream) {
  sum_over_t_k<<<ceil_div((long long)B * C, 128), 128, 0, ST>>>(dy, T, B, C, dv);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_bernoulli_mask(uint8_t* out, long long n, float keep_prob, unsigned long long seed, void* stream) {
  bernoulli_k<<<grid_for(n, 256), 256, 0, ST>>>(out, n, keep_prob, seed);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_bernoulli_mask_dev(uint8_t* out, long long n, float keep_prob, const unsigned long long* seed_dev, unsigned long long salt,
                            void* stream) {
  SATK_CHECK_ARG(seed_dev != nullptr, "bernoulli_mask_dev: seed_dev is NULL%s", "");
  bernoulli_dev_k<<<grid_for(n, 256), 256, 0, ST>>>(out, n, keep_prob, seed_dev, salt);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_softmax_fwd(float* S, int nmat, int T, int causal, const uint8_t* keep_mask, float keep_scale, float* Pd, void* stream) {
  long long rows = (long long)nmat * T;
  softmax_fwd_k<<<grid_for(rows, 8), 256, 0, ST>>>(S, nmat, T, causal, keep_mask, keep_scale, Pd);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_softmax_bwd(const float* P, const float* dPd, int nmat, int T, int causal, const uint8_t* keep_mask, float keep_scale,
                     float* dS, void* stream) {
  long long rows = (long long)nmat * T;
  softmax_bwd_k<<<grid_for(rows, 8), 256, 0, ST>>>(P, dPd, nmat, T, causal, keep_mask, keep_scale, dS);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_teacher_inputs(const float* mel, int B, int Tm, int n_mels, int r, int n_feed, float* out, void* stream) {
  SATK_CHECK_ARG(Tm % r == 0 && n_feed <= r, "teacher_inputs: Tm=%d r=%d n_feed=%d", Tm, r, n_feed);
  teacher_inputs_k<<<grid_for((long long)(Tm / r) * B * n_mels * n_feed, 256), 256, 0, ST>>>(mel, B, Tm, n_mels, r, n_feed, out);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_losses(const float* pred_tm, const float* stop_tm, const float* mel, const float* done, const float* spec_mask,
                const float* bin_mask, int B, int Tm, int n_mels, int r, float* out3, float* dpred_tm, float* dstop_tm,
                float* scratch4, void* stream) {
  SATK_CUDA(cudaMemsetAsync(scratch4, 0, sizeof(float) * 4, ST));
  long long n = (long long)B * Tm * n_mels;
  loss_reduce_k<<<grid_for(n, 1024), 256, 0, ST>>>(pred_tm, stop_tm, mel, done, spec_mask, bin_mask, B, Tm, n_mels, r, scratch4);
  loss_grad_k<<<grid_for(n, 256), 256, 0, ST>>>(pred_tm, stop_tm, mel, done, spec_mask, bin_mask, B, Tm, n_mels, r, scratch4, out3,
                                                dpred_tm, dstop_tm);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_grad_sumsq(const float* g, long long n, float* sumsq, void* stream) {
  SATK_CUDA(cudaMemsetAsync(sumsq, 0, 2 * sizeof(float), ST));
  sumsq_k<<<kSMs * 4, 256, 0, ST>>>(g, n, sumsq);      // SATK_SUMSQ_SCRATCH = 2 + 4 * 148 floats
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_l2_reg(const float* p, const float* mask, long long n, float scale, float* g, float* loss_acc, void* stream) {
  SATK_CHECK_ARG(p && mask && n > 0 && (g || loss_acc), "l2_reg: needs parameters, a mask and at least one output");
  l2_reg_k<<<kSMs * 4, 256, 0, ST>>>(p, mask, n, scale, g, loss_acc);
  SATK_LAUNCH_CHECK();
  return 0;
}
int satk_adam_clip(float* p, const float* g, float* m, float* v, long long n, const float* sumsq, float grad_scale,
                   float clip_norm, float lr, float beta1, float beta2, float eps, int step, void* stream) {
  SATK_CHECK_ARG(step >= 1, "adam: step must be >= 1");
  float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
  adam_clip_k<<<kSMs * 8, 256, 0, ST>>>(p, g, m, v, n, sumsq, grad_scale, clip_norm, lr_t, beta1, beta2, eps);
  SATK_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
