// Paged-KV attention for the dynamic (continuous-batching) inference engine.
//
// One step of the engine carries a ragged mix of queries: single decode tokens of running requests and chunks of prompt
// tokens of requests being prefilled.  Every query token t belongs to request r = tok_req[t] and sees the first
// tok_len[t] keys of that request; keys live in fixed-size pages of a shared pool, found through the request's block
// table.  A CTA owns one (query token, KV head) pair -- all G query heads of the group, so each key/value row is read
// from HBM once per group -- and, for long contexts, one of `splits` key ranges; partial (max, sum, accumulator)
// triples are merged by a second tiny kernel (flash-decoding).  Decode attention moves 2 * len * d * 2 bytes per
// (token, KV head) and does O(G) flops per byte: it is HBM bound, so the kernel is built around 16-byte loads with
// several keys in flight per warp, not around the tensor cores.
//
// Replaces the reference's dynamic-context attention (core/inference/contexts/dynamic_context.py + flash_attn_with_kvcache
// in attention.py:590-603) and this repo's earlier gather + einsum path (inference/dynamic.py).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "b200_api.h"
#include "common.cuh"

namespace b200 {
namespace {

template <bool kBf16>
__device__ __forceinline__ void unpack8(const int4& r, float (&f)[8]) {
  const uint32_t w[4] = {static_cast<uint32_t>(r.x), static_cast<uint32_t>(r.y), static_cast<uint32_t>(r.z), static_cast<uint32_t>(r.w)};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kBf16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 v = __half22float2(h);
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
  }
}

template <bool kBf16>
__device__ __forceinline__ void store_out(void* base, int64_t idx, float v) {
  if constexpr (kBf16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16(v);
  else reinterpret_cast<__half*>(base)[idx] = __float2half(v);
}

template <int D, int G, bool kBf16>
__global__ void __launch_bounds__(128) paged_attn_kernel(PagedAttnArgs a) {
  constexpr int kLPK = D / 8;        // lanes that share one key row (16 B each)
  constexpr int kKPW = 32 / kLPK;    // keys per warp per iteration
  constexpr int kGroups = 4 * kKPW;  // key groups per CTA
  const int t = blockIdx.x, hk = blockIdx.y, split = blockIdx.z, nsplit = gridDim.z;
  const int r = a.tok_req[t];
  const int len = a.tok_len[t];
  const int per = ((len + nsplit - 1) / nsplit + kGroups - 1) / kGroups * kGroups;
  const int k0 = split * per, k1 = min(len, k0 + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lk = lane % kLPK, grp = warp * kKPW + lane / kLPK;
  // the key groups of a warp run different trip counts at the tail of a context: shuffles name only the group's lanes
  const unsigned gmask = (kLPK == 32) ? 0xffffffffu : (((1u << kLPK) - 1u) << (lane / kLPK * kLPK));
  const int np = a.ng * G;

  __shared__ float s_m[kGroups][G], s_l[kGroups][G];
  __shared__ float s_acc[kGroups][G][D];

  // this lane's 8 dims of every query head of the group, pre-scaled into the exp2 domain
  float q[G][8];
  {
    const uint16_t* qp = reinterpret_cast<const uint16_t*>(a.q) + (static_cast<int64_t>(t) * np + hk * G) * D + lk * 8;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int4 raw = *reinterpret_cast<const int4*>(qp + g * D);
      unpack8<kBf16>(raw, q[g]);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[g][j] *= a.scale_log2;
    }
  }
  float m[G], l[G], acc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[g][j] = 0.f;
  }
  const int32_t* bt = a.block_table + static_cast<int64_t>(r) * a.max_blocks;
  const uint16_t* kc = reinterpret_cast<const uint16_t*>(a.k_cache);
  const uint16_t* vc = reinterpret_cast<const uint16_t*>(a.v_cache);
  for (int k = k0 + grp; k < k1; k += kGroups) {
    const int64_t slot = static_cast<int64_t>(bt[k / a.block_size]) * a.block_size + k % a.block_size;
    const int64_t off = (slot * a.ng + hk) * D + lk * 8;
    const int4 kraw = ld_nc_v4(reinterpret_cast<const int4*>(kc + off));
    const int4 vraw = ld_nc_v4(reinterpret_cast<const int4*>(vc + off));
    float kf[8], vf[8];
    unpack8<kBf16>(kraw, kf);
    unpack8<kBf16>(vraw, vf);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(q[g][j], kf[j], s);
#pragma unroll
      for (int o = kLPK / 2; o > 0; o >>= 1) s += __shfl_xor_sync(gmask, s, o, kLPK);
      const float mn = fmaxf(m[g], s);
      const float corr = exp2f(m[g] - mn), p = exp2f(s - mn);
      l[g] = l[g] * corr + p;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[g][j] = fmaf(acc[g][j], corr, p * vf[j]);
      m[g] = mn;
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (lk == 0) {
      s_m[grp][g] = m[g];
      s_l[grp][g] = l[g];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[grp][g][lk * 8 + j] = acc[g][j];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * D; idx += blockDim.x) {
    const int g = idx / D, d = idx % D;
    float M = -INFINITY;
#pragma unroll
    for (int c = 0; c < kGroups; ++c) M = fmaxf(M, s_m[c][g]);
    float L = 0.f, O = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int c = 0; c < kGroups; ++c) {
        const float w = exp2f(s_m[c][g] - M);
        L = fmaf(s_l[c][g], w, L);
        O = fmaf(s_acc[c][g][d], w, O);
      }
    }
    const int64_t head = static_cast<int64_t>(t) * np + hk * G + g;
    if (nsplit == 1) {
      store_out<kBf16>(a.out, head * D + d, L > 0.f ? O / L : 0.f);
    } else {
      a.ws_o[(head * nsplit + split) * D + d] = O;
      if (d == 0) {
        a.ws_ml[(head * nsplit + split) * 2] = M;
        a.ws_ml[(head * nsplit + split) * 2 + 1] = L;
      }
    }
  }
}

template <int D, bool kBf16>
__global__ void __launch_bounds__(D) paged_attn_combine_kernel(PagedAttnArgs a, int nsplit) {
  const int64_t head = blockIdx.x;
  const int d = threadIdx.x;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, a.ws_ml[(head * nsplit + s) * 2]);
  float L = 0.f, O = 0.f;
  if (M > -INFINITY) {
    for (int s = 0; s < nsplit; ++s) {
      const float w = exp2f(a.ws_ml[(head * nsplit + s) * 2] - M);
      L = fmaf(a.ws_ml[(head * nsplit + s) * 2 + 1], w, L);
      O = fmaf(a.ws_o[(head * nsplit + s) * D + d], w, O);
    }
  }
  store_out<kBf16>(a.out, head * D + d, L > 0.f ? O / L : 0.f);
}

template <int D, int G>
void launch(const PagedAttnArgs& a, int splits, cudaStream_t s) {
  const dim3 grid(a.n_tokens, a.ng, splits);
  if (a.is_bf16) paged_attn_kernel<D, G, true><<<grid, 128, 0, s>>>(a);
  else paged_attn_kernel<D, G, false><<<grid, 128, 0, s>>>(a);
  if (splits > 1) {
    const int heads = a.n_tokens * a.ng * G;
    if (a.is_bf16) paged_attn_combine_kernel<D, true><<<heads, D, 0, s>>>(a, splits);
    else paged_attn_combine_kernel<D, false><<<heads, D, 0, s>>>(a, splits);
  }
}

}  // namespace

bool paged_attention_supported(int head_dim, int group) {
  return (head_dim == 64 || head_dim == 128) && (group == 1 || group == 2 || group == 4 || group == 8);
}

void paged_attention(const PagedAttnArgs& a, int splits, cudaStream_t s) {
#define B200_PA(D, G) if (a.head_dim == D && a.group == G) return launch<D, G>(a, splits, s);
  B200_PA(128, 1) B200_PA(128, 2) B200_PA(128, 4) B200_PA(128, 8)
  B200_PA(64, 1) B200_PA(64, 2) B200_PA(64, 4) B200_PA(64, 8)
#undef B200_PA
}

}  // namespace b200
