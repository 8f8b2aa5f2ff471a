// Bandwidth-bound sm_100a kernels: norms, activations, RoPE, bias/dropout/residual, cross
// entropy, optimizer and multi-tensor utilities.  All are 16-byte vectorised, keep per-row data
// in registers (one HBM read + one write per element) and do their statistics in fp32.
//
// Reference call sites these replace (capability parity, not code): fused_layer_norm.py:30-169,
// fused_bias_gelu.py:16-33, fused_bias_swiglu.py:14-58, fused_bias_dropout.py:9-73,
// rope_utils.py:24-42, tensor_parallel/cross_entropy.py:123-220, optimizer/__init__.py:11-26
// (FusedAdam), clip_grads.py:12-44 (multi-tensor l2norm / scale).
#include "b200_api.h"
#include "common.cuh"

#include <curand_kernel.h>

namespace b200 {

namespace {

template <typename T> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float2 up(uint32_t u) { return unpack_bf16x2(u); }
  static __device__ __forceinline__ uint32_t down(float a, float b) { return pack_bf16x2(a, b); }
  static __device__ __forceinline__ float up1(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 down1(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float2 up(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t down(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float up1(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half down1(float v) { return __float2half_rn(v); }
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = Cvt<T>::up(u.x), b = Cvt<T>::up(u.y), c = Cvt<T>::up(u.z), d = Cvt<T>::up(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]) {
  uint4 u;
  u.x = Cvt<T>::down(f[0], f[1]); u.y = Cvt<T>::down(f[2], f[3]);
  u.z = Cvt<T>::down(f[4], f[5]); u.w = Cvt<T>::down(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : -INFINITY;
  t = warp_max(t);
  return t;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  return 0.5f * x * (1.f + tanhf(0.79788456f * x * (1.f + 0.044715f * x * x)));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float t = tanhf(0.79788456f * x * (1.f + 0.044715f * x * x));
  return 0.5f * x * ((1.f - t * t) * (0.79788456f + 0.1070322243f * x * x)) + 0.5f * (1.f + t);
}

// ------------------------------------------------------------------ LayerNorm / RMSNorm
constexpr int kLnThreads = 256;
constexpr int kLnMaxVec = 4;  // cached 8-element vectors per thread => cols <= 8192 stay in registers

template <typename T, bool kRms>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const T* __restrict__ x,
                                                            const T* __restrict__ gamma,
                                                            const T* __restrict__ beta,
                                                            T* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int rows,
                                                            int cols, float eps, bool zero_centered) {
  __shared__ float red[32];
  const int nvec = cols / 8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + static_cast<size_t>(row) * cols;
    float cache[kLnMaxVec][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = threadIdx.x + i * kLnThreads;
      if (v < nvec) {
        load8(xr + v * 8, cache[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += kRms ? cache[i][j] * cache[i][j] : cache[i][j];
      }
    }
    for (int v = threadIdx.x + kLnMaxVec * kLnThreads; v < nvec; v += kLnThreads) {
      float t[8];
      load8(xr + v * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += kRms ? t[j] * t[j] : t[j];
    }
    s = block_sum(s, red);
    float mean = 0.f, rstd;
    if (kRms) {
      rstd = rsqrtf(s / cols + eps);
    } else {
      mean = s / cols;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < kLnMaxVec; ++i) {
        const int v = threadIdx.x + i * kLnThreads;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = cache[i][j] - mean; q += d * d; }
        }
      }
      for (int v = threadIdx.x + kLnMaxVec * kLnThreads; v < nvec; v += kLnThreads) {
        float t[8];
        load8(xr + v * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = t[j] - mean; q += d * d; }
      }
      q = block_sum(q, red);
      rstd = rsqrtf(q / cols + eps);
    }
    if (threadIdx.x == 0) {
      if (mean_out) mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    T* yr = y + static_cast<size_t>(row) * cols;
    auto emit = [&](int v, const float (&xv)[8]) {
      float g[8], b[8], o[8];
      load8(gamma + v * 8, g);
      if (beta) load8(beta + v * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = zero_centered ? g[j] + 1.f : g[j];
        o[j] = (xv[j] - mean) * rstd * gg + (beta ? b[j] : 0.f);
      }
      store8(yr + v * 8, o);
    };
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = threadIdx.x + i * kLnThreads;
      if (v < nvec) emit(v, cache[i]);
    }
    for (int v = threadIdx.x + kLnMaxVec * kLnThreads; v < nvec; v += kLnThreads) {
      float t[8];
      load8(xr + v * 8, t);
      emit(v, t);
    }
  }
}

// dx plus per-CTA partial column sums of dgamma/dbeta (finalised by a second tiny kernel)
template <typename T, bool kRms>
__global__ void __launch_bounds__(kLnThreads) ln_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
    float* __restrict__ dgamma_partial, float* __restrict__ dbeta_partial, int rows, int cols,
    bool zero_centered) {
  __shared__ float red[32];
  const int nvec = cols / 8;
  float dg_acc[kLnMaxVec][8], db_acc[kLnMaxVec][8];
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg_acc[i][j] = 0.f; db_acc[i][j] = 0.f; }
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + static_cast<size_t>(row) * cols;
    const T* dyr = dy + static_cast<size_t>(row) * cols;
    const float mu = kRms ? 0.f : mean[row];
    const float rs = rstd[row];
    float xh[kLnMaxVec][8], gd[kLnMaxVec][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = threadIdx.x + i * kLnThreads;
      if (v < nvec) {
        float xv[8], dv[8], g[8];
        load8(xr + v * 8, xv);
        load8(dyr + v * 8, dv);
        load8(gamma + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = zero_centered ? g[j] + 1.f : g[j];
          xh[i][j] = (xv[j] - mu) * rs;
          gd[i][j] = gg * dv[j];
          s1 += gd[i][j];
          s2 += gd[i][j] * xh[i][j];
          dg_acc[i][j] += dv[j] * xh[i][j];
          db_acc[i][j] += dv[j];
        }
      }
    }
    if (!kRms) s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    const float m1 = kRms ? 0.f : s1 / cols, m2 = s2 / cols;
    T* dxr = dx + static_cast<size_t>(row) * cols;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = threadIdx.x + i * kLnThreads;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs * (gd[i][j] - m1 - xh[i][j] * m2);
        store8(dxr + v * 8, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = threadIdx.x + i * kLnThreads;
    if (v < nvec) {
      float* pg = dgamma_partial + static_cast<size_t>(blockIdx.x) * cols + v * 8;
      *reinterpret_cast<float4*>(pg) = make_float4(dg_acc[i][0], dg_acc[i][1], dg_acc[i][2], dg_acc[i][3]);
      *reinterpret_cast<float4*>(pg + 4) = make_float4(dg_acc[i][4], dg_acc[i][5], dg_acc[i][6], dg_acc[i][7]);
      if (dbeta_partial) {
        float* pb = dbeta_partial + static_cast<size_t>(blockIdx.x) * cols + v * 8;
        *reinterpret_cast<float4*>(pb) = make_float4(db_acc[i][0], db_acc[i][1], db_acc[i][2], db_acc[i][3]);
        *reinterpret_cast<float4*>(pb + 4) = make_float4(db_acc[i][4], db_acc[i][5], db_acc[i][6], db_acc[i][7]);
      }
    }
  }
}

// ---- fast path: cols == 8 * VPT * blockDim.x; every thread owns VPT 8-element vectors of a row,
// gamma/beta live in registers across the persistent row loop and the next row's loads are issued
// before the reductions of the current row (more bytes in flight per SM -> closer to HBM peak)
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = make_float2(a, b);
  __syncthreads();
  float2 t = (l < nw) ? red[l] : make_float2(0.f, 0.f);
  t.x = warp_sum(t.x);
  t.y = warp_sum(t.y);
  return t;
}

template <typename T, bool kRms, int VPT>
__global__ void __launch_bounds__(512) ln_fwd_fast_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                          const T* __restrict__ beta, T* __restrict__ y,
                                                          float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                          int rows, int cols, float eps, bool zero_centered) {
  __shared__ float2 red[32];
  float g[VPT][8], b[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    load8(gamma + v * 8, g[i]);
    if (beta) load8(beta + v * 8, b[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (zero_centered) g[i][j] += 1.f;
      if (!beta) b[i][j] = 0.f;
    }
  }
  uint4 nxt[VPT];
  int row = blockIdx.x;
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      nxt[i] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * cols + (threadIdx.x + i * blockDim.x) * 8);
  }
  const float inv_cols = 1.f / cols;
  for (; row < rows; row += gridDim.x) {
    float c[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const uint4 u = nxt[i];
      float2 p0 = Cvt<T>::up(u.x), p1 = Cvt<T>::up(u.y), p2 = Cvt<T>::up(u.z), p3 = Cvt<T>::up(u.w);
      c[i][0] = p0.x; c[i][1] = p0.y; c[i][2] = p1.x; c[i][3] = p1.y;
      c[i][4] = p2.x; c[i][5] = p2.y; c[i][6] = p3.x; c[i][7] = p3.y;
    }
    const int nrow = row + gridDim.x;
    if (nrow < rows) {
#pragma unroll
      for (int i = 0; i < VPT; ++i)
        nxt[i] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(nrow) * cols + (threadIdx.x + i * blockDim.x) * 8);
    }
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { s += c[i][j]; q += c[i][j] * c[i][j]; }
    float mean = 0.f, rstd;
    if (kRms) {
      const float2 t = block_sum2(q, 0.f, red);
      rstd = rsqrtf(t.x * inv_cols + eps);
    } else {
      // two-pass variance on register-resident values (exact, no E[x^2]-mean^2 cancellation)
      const float2 t = block_sum2(s, 0.f, red);
      mean = t.x * inv_cols;
      float d2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPT; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = c[i][j] - mean; d2 += d * d; }
      const float2 t2 = block_sum2(d2, 0.f, red);
      rstd = rsqrtf(t2.x * inv_cols + eps);
    }
    if (threadIdx.x == 0) {
      if (mean_out) mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    T* yr = y + static_cast<size_t>(row) * cols;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (c[i][j] - mean) * rstd * g[i][j] + b[i][j];
      store8(yr + (threadIdx.x + i * blockDim.x) * 8, o);
    }
  }
}

template <typename T, bool kRms, int VPT>
__global__ void __launch_bounds__(512) ln_bwd_fast_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ rstd, T* __restrict__ dx, float* __restrict__ dgamma_partial,
    float* __restrict__ dbeta_partial, int rows, int cols, bool zero_centered) {
  __shared__ float2 red[32];
  float g[VPT][8], dg_acc[VPT][8], db_acc[VPT][8];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    load8(gamma + (threadIdx.x + i * blockDim.x) * 8, g[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (zero_centered) g[i][j] += 1.f;
      dg_acc[i][j] = 0.f;
      db_acc[i][j] = 0.f;
    }
  }
  uint4 nx[VPT], nd[VPT];
  float nmu = 0.f, nrs = 0.f;
  int row = blockIdx.x;
  auto prefetch = [&](int r) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const size_t off = static_cast<size_t>(r) * cols + (threadIdx.x + i * blockDim.x) * 8;
      nx[i] = *reinterpret_cast<const uint4*>(x + off);
      nd[i] = *reinterpret_cast<const uint4*>(dy + off);
    }
    nmu = kRms ? 0.f : mean[r];
    nrs = rstd[r];
  };
  if (row < rows) prefetch(row);
  const float inv_cols = 1.f / cols;
  for (; row < rows; row += gridDim.x) {
    float xh[VPT][8], gd[VPT][8];
    const float mu = nmu, rs = nrs;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const uint4 ux = nx[i], ud = nd[i];
      const float2 x0 = Cvt<T>::up(ux.x), x1 = Cvt<T>::up(ux.y), x2 = Cvt<T>::up(ux.z), x3 = Cvt<T>::up(ux.w);
      const float2 d0 = Cvt<T>::up(ud.x), d1 = Cvt<T>::up(ud.y), d2 = Cvt<T>::up(ud.z), d3 = Cvt<T>::up(ud.w);
      const float xv[8] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y};
      const float dv[8] = {d0.x, d0.y, d1.x, d1.y, d2.x, d2.y, d3.x, d3.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xv[j] - mu) * rs;
        gd[i][j] = g[i][j] * dv[j];
        s1 += gd[i][j];
        s2 += gd[i][j] * xh[i][j];
        dg_acc[i][j] += dv[j] * xh[i][j];
        db_acc[i][j] += dv[j];
      }
    }
    const int nrow = row + gridDim.x;
    if (nrow < rows) prefetch(nrow);
    const float2 t = block_sum2(s1, s2, red);
    const float m1 = kRms ? 0.f : t.x * inv_cols, m2 = t.y * inv_cols;
    T* dxr = dx + static_cast<size_t>(row) * cols;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rs * (gd[i][j] - m1 - xh[i][j] * m2);
      store8(dxr + (threadIdx.x + i * blockDim.x) * 8, o);
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    float* pg = dgamma_partial + static_cast<size_t>(blockIdx.x) * cols + v * 8;
    *reinterpret_cast<float4*>(pg) = make_float4(dg_acc[i][0], dg_acc[i][1], dg_acc[i][2], dg_acc[i][3]);
    *reinterpret_cast<float4*>(pg + 4) = make_float4(dg_acc[i][4], dg_acc[i][5], dg_acc[i][6], dg_acc[i][7]);
    if (dbeta_partial) {
      float* pb = dbeta_partial + static_cast<size_t>(blockIdx.x) * cols + v * 8;
      *reinterpret_cast<float4*>(pb) = make_float4(db_acc[i][0], db_acc[i][1], db_acc[i][2], db_acc[i][3]);
      *reinterpret_cast<float4*>(pb + 4) = make_float4(db_acc[i][4], db_acc[i][5], db_acc[i][6], db_acc[i][7]);
    }
  }
}

// vectors-per-thread / block size for the fast path; vpt = 0 -> generic kernels
static inline void ln_fast_shape(int cols, int& vpt, int& threads) {
  vpt = 0; threads = 0;
  if (cols % 8) return;
  const int nvec = cols / 8;
  for (int v : {1, 2, 4}) {
    if (nvec % v == 0) {
      const int t = nvec / v;
      if (t <= 512 && t % 32 == 0) { vpt = v; threads = t; return; }
    }
  }
}

__global__ void __launch_bounds__(512) ln_bwd_finalize_kernel(const float* __restrict__ pg, const float* __restrict__ pb,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       int n_partial, int cols, bool accumulate) {
  // 32 columns x 16 partial-row lanes per CTA: coalesced 128B rows, smem tree over the 16 lanes
  __shared__ float sg[16][33], sb[16][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float g = 0.f, b = 0.f;
  if (c < cols) {
    for (int i = ty; i < n_partial; i += 16) {
      g += pg[static_cast<size_t>(i) * cols + c];
      if (pb) b += pb[static_cast<size_t>(i) * cols + c];
    }
  }
  sg[ty][tx] = g; sb[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 16; ++i) { g += sg[i][tx]; b += sb[i][tx]; }
    dgamma[c] = accumulate ? dgamma[c] + g : g;
    if (pb && dbeta) dbeta[c] = accumulate ? dbeta[c] + b : b;
  }
}

// ------------------------------------------------------------------ activations
template <typename T>
__global__ void bias_gelu_fwd_kernel(const T* __restrict__ x, const T* __restrict__ bias,
                                     T* __restrict__ y, int64_t nvec, int cols) {
  const int vpr = cols / 8;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    load8(x + v * 8, a);
    if (bias) {
      load8(bias + (v % vpr) * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = gelu_tanh_f(a[j]);
    store8(y + v * 8, a);
  }
}
template <typename T>
__global__ void bias_gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                     const T* __restrict__ bias, T* __restrict__ dx, int64_t nvec,
                                     int cols) {
  const int vpr = cols / 8;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8], d[8];
    load8(x + v * 8, a);
    load8(dy + v * 8, d);
    if (bias) {
      load8(bias + (v % vpr) * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] *= gelu_tanh_grad(a[j]);
    store8(dx + v * 8, d);
  }
}
// x: [rows, 2*cols_out] = [gate | up];  y = silu(gate) * up
template <typename T>
__global__ void swiglu_fwd_kernel(const T* __restrict__ x, const T* __restrict__ bias,
                                  T* __restrict__ y, int64_t rows, int cols_out) {
  const int vpr = cols_out / 8;
  const int64_t nvec = rows * vpr;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = v / vpr;
    const int c = static_cast<int>(v - r * vpr) * 8;
    float g[8], u[8], b[8];
    load8(x + r * 2 * cols_out + c, g);
    load8(x + r * 2 * cols_out + cols_out + c, u);
    if (bias) {
      load8(bias + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += b[j];
      load8(bias + cols_out + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] += b[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = g[j] / (1.f + __expf(-g[j])) * u[j];
    store8(y + r * cols_out + c, g);
  }
}
template <typename T>
__global__ void swiglu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                  const T* __restrict__ bias, T* __restrict__ dx, int64_t rows,
                                  int cols_out) {
  const int vpr = cols_out / 8;
  const int64_t nvec = rows * vpr;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = v / vpr;
    const int c = static_cast<int>(v - r * vpr) * 8;
    float g[8], u[8], b[8], d[8], dg[8], du[8];
    load8(x + r * 2 * cols_out + c, g);
    load8(x + r * 2 * cols_out + cols_out + c, u);
    load8(dy + r * cols_out + c, d);
    if (bias) {
      load8(bias + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += b[j];
      load8(bias + cols_out + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] += b[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-g[j]));
      const float silu = g[j] * sg;
      dg[j] = d[j] * u[j] * (sg * (1.f + g[j] * (1.f - sg)));
      du[j] = d[j] * silu;
    }
    store8(dx + r * 2 * cols_out + c, dg);
    store8(dx + r * 2 * cols_out + cols_out + c, du);
  }
}

// out = residual + dropout(x + bias); mask (1 = kept) saved as bytes when p > 0
template <typename T>
__global__ void bias_dropout_add_kernel(const T* __restrict__ x, const T* __restrict__ bias,
                                        const T* __restrict__ residual, T* __restrict__ out,
                                        uint8_t* __restrict__ mask, int64_t nvec, int cols, float p,
                                        uint64_t seed, uint64_t offset) {
  const int vpr = cols / 8;
  const float keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8], r[8];
    load8(x + v * 8, a);
    if (bias) {
      load8(bias + (v % vpr) * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
    if (p > 0.f) {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, static_cast<unsigned long long>(v), offset, &st);
      float4 r0 = curand_uniform4(&st), r1 = curand_uniform4(&st);
      const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      uint8_t m[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m[j] = rr[j] >= p ? 1 : 0;
        a[j] = m[j] ? a[j] * keep_scale : 0.f;
      }
      if (mask) *reinterpret_cast<uint2*>(mask + v * 8) = *reinterpret_cast<uint2*>(m);
    }
    if (residual) {
      load8(residual + v * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += r[j];
    }
    store8(out + v * 8, a);
  }
}
template <typename T>
__global__ void dropout_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ mask,
                                   T* __restrict__ dx, int64_t nvec, float keep_scale) {
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float d[8];
    load8(dy + v * 8, d);
    uint2 mm = *reinterpret_cast<const uint2*>(mask + v * 8);
    const uint8_t* m = reinterpret_cast<const uint8_t*>(&mm);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = m[j] ? d[j] * keep_scale : 0.f;
    store8(dx + v * 8, d);
  }
}

// RoPE on x[s, b, h, d]; cos_sin[s, rot_dim] holds cos in [0,rot/2) and sin in [rot/2, rot)
template <typename T>
__global__ void rope_kernel(const T* __restrict__ x, const float* __restrict__ cos_sin,
                            T* __restrict__ y, int s, int bh, int d, int rot_dim, bool interleaved,
                            float sign) {
  const int half = rot_dim / 2;
  const int64_t total = static_cast<int64_t>(s) * bh * (d / 2);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int pair = static_cast<int>(i % (d / 2));
    const int64_t vec = i / (d / 2);  // (s, b, h) index
    const int si = static_cast<int>(vec / bh);
    const T* xp = x + vec * d;
    T* yp = y + vec * d;
    if (pair < half) {
      const float c = cos_sin[static_cast<size_t>(si) * rot_dim + pair];
      const float sn = sign * cos_sin[static_cast<size_t>(si) * rot_dim + half + pair];
      const int i0 = interleaved ? 2 * pair : pair;
      const int i1 = interleaved ? 2 * pair + 1 : pair + half;
      const float x0 = Cvt<T>::up1(xp[i0]), x1 = Cvt<T>::up1(xp[i1]);
      yp[i0] = Cvt<T>::down1(x0 * c - x1 * sn);
      yp[i1] = Cvt<T>::down1(x1 * c + x0 * sn);
    } else {
      // pass-through tail (rot_dim < d): pairs [half, d/2) map to elements [rot_dim, d)
      const int e = rot_dim + 2 * (pair - half);
      yp[e] = xp[e];
      yp[e + 1] = xp[e + 1];
    }
  }
}

// dbias[c] (+)= sum_r dy[r, c]
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dy, float* __restrict__ dbias, int64_t rows,
                                 int cols, int rows_per_cta) {
  __shared__ float sm[32][65];
  const int cv = threadIdx.x & 7;        // 8 vectors of 8 columns = 64 columns per CTA
  const int rl = threadIdx.x >> 3;       // 32 row lanes
  const int col0 = blockIdx.x * 64 + cv * 8;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col0 < cols) {
    for (int64_t r = r0 + rl; r < r1; r += 32) {
      float d[8];
      load8(dy + r * cols + col0, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[rl][cv * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < cols) atomicAdd(dbias + c, t);
  }
}

// ------------------------------------------------------------------ cross entropy
template <typename T>
__device__ __forceinline__ float ld_logit(const T* p, int64_t i) { return Cvt<T>::up1(p[i]); }
template <>
__device__ __forceinline__ float ld_logit<float>(const float* p, int64_t i) { return p[i]; }

template <typename T> struct VecLd {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const T* p, float (&f)[8]) { load8(p, f); }
  static __device__ __forceinline__ void st(T* p, const float (&f)[8]) { store8(p, f); }
};
template <> struct VecLd<float> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const float* p, float (&f)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// one CTA per row, single pass: per-thread online (max, sum) over 8-wide vectors, then a block merge
template <typename T>
__global__ void __launch_bounds__(512) ce_stats_kernel(const T* __restrict__ logits,
                                                       const int64_t* __restrict__ labels,
                                                       float* __restrict__ row_max,
                                                       float* __restrict__ row_sumexp,
                                                       float* __restrict__ tgt, int64_t rows,
                                                       int vocab_local, int vocab_start, int ldl) {
  __shared__ float red[32];
  const bool vec_ok = (ldl % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 31) == 0);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* lr = logits + row * ldl;
    float m = -INFINITY, s = 0.f;
    const int nvec = vec_ok ? vocab_local / 8 : 0;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      float f[8];
      VecLd<T>::ld(lr + v * 8, f);
      float vm = f[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) vm = fmaxf(vm, f[j]);
      const float nm = fmaxf(m, vm);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += __expf(f[j] - nm);
      s = s * __expf(m - nm) + acc;
      m = nm;
    }
    for (int c = nvec * 8 + threadIdx.x; c < vocab_local; c += blockDim.x) {
      const float x = ld_logit(lr, c);
      const float nm = fmaxf(m, x);
      s = s * __expf(m - nm) + __expf(x - nm);
      m = nm;
    }
    const float gm = block_max(m, red);
    s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
      row_max[row] = gm;
      row_sumexp[row] = s;
      const int64_t lab = labels[row] - vocab_start;
      tgt[row] = (lab >= 0 && lab < vocab_local) ? ld_logit(lr, lab) : 0.f;
    }
  }
}
__global__ void ce_rescale_kernel(float* __restrict__ sumexp, const float* __restrict__ local_max,
                                  const float* __restrict__ global_max, int64_t rows) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < rows) sumexp[i] *= __expf(local_max[i] - global_max[i]);
}
template <typename T> __device__ __forceinline__ void st_logit(T* p, int64_t i, float v) { p[i] = Cvt<T>::down1(v); }
template <> __device__ __forceinline__ void st_logit<float>(float* p, int64_t i, float v) { p[i] = v; }

template <typename T>
__global__ void __launch_bounds__(512) ce_grad_kernel(const T* __restrict__ logits,
                                                      const int64_t* __restrict__ labels,
                                                      const float* __restrict__ gmax,
                                                      const float* __restrict__ gsum,
                                                      const float* __restrict__ grad_loss,
                                                      T* __restrict__ dlogits, int64_t rows,
                                                      int vocab_local, int vocab_start, int ldl,
                                                      float label_smoothing, int64_t vocab_total) {
  const bool vec_ok = (ldl % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 31) == 0) &&
                      ((reinterpret_cast<uintptr_t>(dlogits) & 31) == 0);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* lr = logits + row * ldl;
    T* dr = dlogits + row * ldl;
    const float m = gmax[row], inv = 1.f / gsum[row], g = grad_loss[row];
    const int64_t lab = labels[row] - vocab_start;
    const float smooth = label_smoothing > 0.f ? label_smoothing / static_cast<float>(vocab_total - 1) : 0.f;
    const float on_target = label_smoothing > 0.f ? 1.f - label_smoothing - smooth : 1.f;
    const int nvec = vec_ok ? vocab_local / 8 : 0;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      float f[8];
      VecLd<T>::ld(lr + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = (v * 8 + j == lab) ? on_target : 0.f;
        f[j] = (__expf(f[j] - m) * inv - t - smooth) * g;
      }
      VecLd<T>::st(dr + v * 8, f);
    }
    for (int c = nvec * 8 + threadIdx.x; c < vocab_local; c += blockDim.x) {
      float pr = __expf(ld_logit(lr, c) - m) * inv;
      float t = (c == lab) ? on_target : 0.f;
      st_logit(dr, c, (pr - t - smooth) * g);
    }
  }
}

// ------------------------------------------------------------------ optimizer & multi-tensor
template <typename T>
__global__ void adam_kernel(AdamArgs a) {
  if (a.found_inf && *a.found_inf != 0.f) return;
  const float inv_bc1 = 1.f / a.bias_correction1;
  const float inv_sqrt_bc2 = rsqrtf(a.bias_correction2);
  const int64_t n4 = a.n / 4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 g = reinterpret_cast<const float4*>(a.grad)[i];
    float4 p = reinterpret_cast<float4*>(a.master)[i];
    float4 m = reinterpret_cast<float4*>(a.exp_avg)[i];
    float4 v = reinterpret_cast<float4*>(a.exp_avg_sq)[i];
    float gg[4] = {g.x, g.y, g.z, g.w}, pp[4] = {p.x, p.y, p.z, p.w};
    float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * a.grad_scale;
      mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gr;
      vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gr * gr;
      const float denom = sqrtf(vv[j]) * inv_sqrt_bc2 + a.eps;
      const float upd = (mm[j] * inv_bc1) / denom + a.weight_decay * pp[j];  // decoupled (AdamW)
      pp[j] -= a.lr * upd;
    }
    reinterpret_cast<float4*>(a.master)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(a.exp_avg)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.exp_avg_sq)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (a.param16) {
      uint2 o;
      o.x = Cvt<T>::down(pp[0], pp[1]);
      o.y = Cvt<T>::down(pp[2], pp[3]);
      reinterpret_cast<uint2*>(a.param16)[i] = o;
    }
  }
  // tail
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    const float gr = a.grad[i] * a.grad_scale;
    float m = a.beta1 * a.exp_avg[i] + (1.f - a.beta1) * gr;
    float v = a.beta2 * a.exp_avg_sq[i] + (1.f - a.beta2) * gr * gr;
    float p = a.master[i];
    p -= a.lr * ((m * inv_bc1) / (sqrtf(v) * inv_sqrt_bc2 + a.eps) + a.weight_decay * p);
    a.master[i] = p; a.exp_avg[i] = m; a.exp_avg_sq[i] = v;
    if (a.param16) reinterpret_cast<T*>(a.param16)[i] = Cvt<T>::down1(p);
  }
}

template <typename T> __device__ __forceinline__ float ld_any(const T* p, int64_t i) { return Cvt<T>::up1(p[i]); }
template <> __device__ __forceinline__ float ld_any<float>(const float* p, int64_t i) { return p[i]; }
template <typename T> __device__ __forceinline__ void st_any(T* p, int64_t i, float v) { p[i] = Cvt<T>::down1(v); }
template <> __device__ __forceinline__ void st_any<float>(float* p, int64_t i, float v) { p[i] = v; }

template <typename T>
__global__ void l2norm_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = ld_any(x, i);
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
template <typename T>
__global__ void scale_kernel(T* __restrict__ x, int64_t n, float scale, const float* __restrict__ sd) {
  const float sc = sd ? scale * (*sd) : scale;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    st_any(x, i, ld_any(x, i) * sc);
}
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n, float scale) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    st_any(dst, i, ld_any(src, i) * scale);
}
template <typename S>
__global__ void add_into_kernel(float* __restrict__ dst, const S* __restrict__ src, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] += ld_any(src, i);
}
template <typename T>
__global__ void finite_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ found) {
  bool bad = false;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = ld_any(x, i);
    bad |= !isfinite(v);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found = 1.f;
}
// MegaScope compressor: mean over contiguous chunks of the hidden dim -> [rows, pixels]
template <typename T>
__global__ void chunk_mean_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows,
                                  int cols, int pixels) {
  const int64_t total = rows * pixels;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / pixels;
    const int px = static_cast<int>(i - r * pixels);
    const int c0 = static_cast<int>(static_cast<int64_t>(px) * cols / pixels);
    const int c1 = static_cast<int>(static_cast<int64_t>(px + 1) * cols / pixels);
    float s = 0.f;
    for (int c = c0; c < c1; ++c) s += Cvt<T>::up1(x[r * cols + c]);
    out[i] = s / max(1, c1 - c0);
  }
}

inline int grid_for(int64_t work_items, int threads, int max_ctas = 148 * 8) {
  int64_t g = (work_items + threads - 1) / threads;
  if (g > max_ctas) g = max_ctas;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

#define DISPATCH_16(is_bf16, ...)                       \
  if (is_bf16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else { using T = __half; __VA_ARGS__; }

#define DISPATCH_ANY(dtype, ...)                                  \
  if ((dtype) == 0) { using T = __nv_bfloat16; __VA_ARGS__; }     \
  else if ((dtype) == 1) { using T = __half; __VA_ARGS__; }       \
  else { using T = float; __VA_ARGS__; }

}  // namespace

int layernorm_bwd_partials(int rows) { return rows < 296 ? rows : 296; }

void layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                   float* rstd, int rows, int cols, float eps, bool rms, bool zero_centered,
                   bool is_bf16, cudaStream_t s) {
  int vpt, threads;
  ln_fast_shape(cols, vpt, threads);
  if (vpt) {
    // enough CTAs to fill every SM to its register/thread limit; rows are grid-strided
    const int per_sm = 2048 / threads > 4 ? 4 : (2048 / threads < 1 ? 1 : 2048 / threads);
    const int grid = rows < 148 * per_sm ? rows : 148 * per_sm;
#define LN_FWD_FAST(V)                                                                                     \
    DISPATCH_16(is_bf16, {                                                                                 \
      if (rms) ln_fwd_fast_kernel<T, true, V><<<grid, threads, 0, s>>>((const T*)x, (const T*)gamma,      \
               (const T*)beta, (T*)y, mean, rstd, rows, cols, eps, zero_centered);                        \
      else ln_fwd_fast_kernel<T, false, V><<<grid, threads, 0, s>>>((const T*)x, (const T*)gamma,         \
               (const T*)beta, (T*)y, mean, rstd, rows, cols, eps, zero_centered);                        \
    })
    if (vpt == 1) { LN_FWD_FAST(1); } else if (vpt == 2) { LN_FWD_FAST(2); } else { LN_FWD_FAST(4); }
#undef LN_FWD_FAST
    return;
  }
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
  DISPATCH_16(is_bf16, {
    if (rms)
      ln_fwd_kernel<T, true><<<grid, kLnThreads, 0, s>>>((const T*)x, (const T*)gamma, (const T*)beta,
                                                         (T*)y, mean, rstd, rows, cols, eps, zero_centered);
    else
      ln_fwd_kernel<T, false><<<grid, kLnThreads, 0, s>>>((const T*)x, (const T*)gamma, (const T*)beta,
                                                          (T*)y, mean, rstd, rows, cols, eps, zero_centered);
  });
}

void layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                   const float* rstd, void* dx, float* dgamma_partial, float* dbeta_partial,
                   int rows, int cols, bool rms, bool zero_centered, bool is_bf16, int n_partial,
                   cudaStream_t s) {
  int vpt, threads;
  ln_fast_shape(cols, vpt, threads);
  if (vpt == 1 || vpt == 2) {   // 4 vectors per thread would spill under the 512-thread register cap
#define LN_BWD_FAST(V)                                                                                     \
    DISPATCH_16(is_bf16, {                                                                                 \
      if (rms) ln_bwd_fast_kernel<T, true, V><<<n_partial, threads, 0, s>>>((const T*)dy, (const T*)x,    \
               (const T*)gamma, mean, rstd, (T*)dx, dgamma_partial, dbeta_partial, rows, cols,            \
               zero_centered);                                                                            \
      else ln_bwd_fast_kernel<T, false, V><<<n_partial, threads, 0, s>>>((const T*)dy, (const T*)x,       \
               (const T*)gamma, mean, rstd, (T*)dx, dgamma_partial, dbeta_partial, rows, cols,            \
               zero_centered);                                                                            \
    })
    if (vpt == 1) { LN_BWD_FAST(1); } else { LN_BWD_FAST(2); }
#undef LN_BWD_FAST
    return;
  }
  DISPATCH_16(is_bf16, {
    if (rms)
      ln_bwd_kernel<T, true><<<n_partial, kLnThreads, 0, s>>>((const T*)dy, (const T*)x, (const T*)gamma,
                                                              mean, rstd, (T*)dx, dgamma_partial,
                                                              dbeta_partial, rows, cols, zero_centered);
    else
      ln_bwd_kernel<T, false><<<n_partial, kLnThreads, 0, s>>>((const T*)dy, (const T*)x, (const T*)gamma,
                                                               mean, rstd, (T*)dx, dgamma_partial,
                                                               dbeta_partial, rows, cols, zero_centered);
  });
}

void layernorm_bwd_finalize(const float* pg, const float* pb, float* dgamma, float* dbeta,
                            int n_partial, int cols, bool accumulate, cudaStream_t s) {
  ln_bwd_finalize_kernel<<<(cols + 31) / 32, 512, 0, s>>>(pg, pb, dgamma, dbeta, n_partial, cols,
                                                            accumulate);
}

void bias_gelu_fwd(const void* x, const void* bias, void* y, int64_t rows, int cols, bool is_bf16,
                   cudaStream_t s) {
  const int64_t nvec = rows * cols / 8;
  DISPATCH_16(is_bf16, (bias_gelu_fwd_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           (const T*)x, (const T*)bias, (T*)y, nvec, cols)));
}
void bias_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, int64_t rows,
                   int cols, bool is_bf16, cudaStream_t s) {
  const int64_t nvec = rows * cols / 8;
  DISPATCH_16(is_bf16, (bias_gelu_bwd_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           (const T*)dy, (const T*)x, (const T*)bias, (T*)dx, nvec, cols)));
}
void swiglu_fwd(const void* x, const void* bias, void* y, int64_t rows, int cols_out, bool is_bf16,
                cudaStream_t s) {
  const int64_t nvec = rows * cols_out / 8;
  DISPATCH_16(is_bf16, (swiglu_fwd_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           (const T*)x, (const T*)bias, (T*)y, rows, cols_out)));
}
void swiglu_bwd(const void* dy, const void* x, const void* bias, void* dx, int64_t rows,
                int cols_out, bool is_bf16, cudaStream_t s) {
  const int64_t nvec = rows * cols_out / 8;
  DISPATCH_16(is_bf16, (swiglu_bwd_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           (const T*)dy, (const T*)x, (const T*)bias, (T*)dx, rows, cols_out)));
}
void bias_dropout_add(const void* x, const void* bias, const void* residual, void* out,
                      uint8_t* mask_out, int64_t rows, int cols, float p, uint64_t seed,
                      uint64_t offset, bool is_bf16, cudaStream_t s) {
  const int64_t nvec = rows * cols / 8;
  DISPATCH_16(is_bf16, (bias_dropout_add_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           (const T*)x, (const T*)bias, (const T*)residual, (T*)out, mask_out, nvec,
                           cols, p, seed, offset)));
}
void dropout_bwd(const void* dy, const uint8_t* mask, void* dx, int64_t n, float p, bool is_bf16,
                 cudaStream_t s) {
  const int64_t nvec = n / 8;
  DISPATCH_16(is_bf16, (dropout_bwd_kernel<T><<<grid_for(nvec, 256), 256, 0, s>>>(
                           (const T*)dy, mask, (T*)dx, nvec, 1.f / (1.f - p))));
}
void rope_apply(const void* x, const float* cos_sin, void* y, int s, int b, int h, int d,
                int rot_dim, bool interleaved, bool inverse, bool is_bf16, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(s) * b * h * (d / 2);
  DISPATCH_16(is_bf16, (rope_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
                           (const T*)x, cos_sin, (T*)y, s, b * h, d, rot_dim, interleaved,
                           inverse ? -1.f : 1.f)));
}
void bias_grad_reduce(const void* dy, float* dbias, int64_t rows, int cols, bool is_bf16,
                      bool accumulate, cudaStream_t s) {
  if (!accumulate) cudaMemsetAsync(dbias, 0, sizeof(float) * cols, s);
  const int col_blocks = (cols + 63) / 64;
  int row_splits = (148 * 4 + col_blocks - 1) / col_blocks;
  if (row_splits < 1) row_splits = 1;
  int rows_per_cta = static_cast<int>((rows + row_splits - 1) / row_splits);
  rows_per_cta = (rows_per_cta + 31) / 32 * 32;
  row_splits = static_cast<int>((rows + rows_per_cta - 1) / rows_per_cta);
  dim3 grid(col_blocks, row_splits);
  DISPATCH_16(is_bf16, (bias_grad_kernel<T><<<grid, 256, 0, s>>>((const T*)dy, dbias, rows, cols, rows_per_cta)));
}

void ce_local_stats(const void* logits, const int64_t* labels, float* row_max, float* row_sumexp,
                    float* tgt_logit, int64_t rows, int vocab_local, int vocab_start, int ldl,
                    int logits_dtype, cudaStream_t s) {
  const int grid = rows < 148 * 8 ? static_cast<int>(rows) : 148 * 8;
  DISPATCH_ANY(logits_dtype, (ce_stats_kernel<T><<<grid, 512, 0, s>>>((const T*)logits, labels, row_max,
                                                                      row_sumexp, tgt_logit, rows,
                                                                      vocab_local, vocab_start, ldl)));
}
void ce_rescale_sumexp(float* row_sumexp, const float* local_max, const float* global_max,
                       int64_t rows, cudaStream_t s) {
  ce_rescale_kernel<<<static_cast<int>((rows + 255) / 256), 256, 0, s>>>(row_sumexp, local_max,
                                                                         global_max, rows);
}
void ce_grad(const void* logits, const int64_t* labels, const float* global_max,
             const float* global_sumexp, const float* grad_loss, void* dlogits, int64_t rows,
             int vocab_local, int vocab_start, int ldl, int logits_dtype, float label_smoothing,
             int64_t vocab_total, cudaStream_t s) {
  const int grid = rows < 148 * 8 ? static_cast<int>(rows) : 148 * 8;
  DISPATCH_ANY(logits_dtype, (ce_grad_kernel<T><<<grid, 512, 0, s>>>(
                                 (const T*)logits, labels, global_max, global_sumexp, grad_loss,
                                 (T*)dlogits, rows, vocab_local, vocab_start, ldl, label_smoothing,
                                 vocab_total)));
}

void adam_step(const AdamArgs& a, cudaStream_t s) {
  const int grid = grid_for(a.n / 4 + 1, 256, 148 * 16);
  if (a.is_bf16) adam_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a);
  else adam_kernel<__half><<<grid, 256, 0, s>>>(a);
}
void l2norm_sq(const void* x, int64_t n, int dtype, float* out_accum, cudaStream_t s) {
  DISPATCH_ANY(dtype, (l2norm_kernel<T><<<grid_for(n, 256, 148 * 4), 256, 0, s>>>((const T*)x, n, out_accum)));
}
void scale_inplace(void* x, int64_t n, int dtype, float scale, const float* scale_dev,
                   cudaStream_t s) {
  DISPATCH_ANY(dtype, (scale_kernel<T><<<grid_for(n, 256), 256, 0, s>>>((T*)x, n, scale, scale_dev)));
}
void cast_copy(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, float scale,
               cudaStream_t s) {
  const int grid = grid_for(n, 256);
#define CAST_CASE(SD, ST, DD, DT) \
  if (src_dtype == SD && dst_dtype == DD) { cast_kernel<ST, DT><<<grid, 256, 0, s>>>((const ST*)src, (DT*)dst, n, scale); return; }
  CAST_CASE(2, float, 0, __nv_bfloat16)
  CAST_CASE(2, float, 1, __half)
  CAST_CASE(0, __nv_bfloat16, 2, float)
  CAST_CASE(1, __half, 2, float)
  CAST_CASE(2, float, 2, float)
  CAST_CASE(0, __nv_bfloat16, 0, __nv_bfloat16)
  CAST_CASE(1, __half, 1, __half)
  CAST_CASE(0, __nv_bfloat16, 1, __half)
  CAST_CASE(1, __half, 0, __nv_bfloat16)
#undef CAST_CASE
}
void add_into_f32(float* dst, const void* src, int src_dtype, int64_t n, cudaStream_t s) {
  DISPATCH_ANY(src_dtype, (add_into_kernel<T><<<grid_for(n, 256), 256, 0, s>>>(dst, (const T*)src, n)));
}
void check_finite(const void* x, int64_t n, int dtype, float* found_inf, cudaStream_t s) {
  DISPATCH_ANY(dtype, (finite_kernel<T><<<grid_for(n, 256, 148 * 4), 256, 0, s>>>((const T*)x, n, found_inf)));
}
void chunk_mean(const void* x, float* out, int64_t rows, int cols, int pixels, bool is_bf16,
                cudaStream_t s) {
  DISPATCH_16(is_bf16, (chunk_mean_kernel<T><<<grid_for(rows * pixels, 256), 256, 0, s>>>(
                           (const T*)x, out, rows, cols, pixels)));
}

}  // namespace b200
