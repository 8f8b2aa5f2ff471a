// Mixture-of-experts device-side plumbing for sm_100a: routing, permutation and combine without a single host
// synchronisation.  The expert FFNs themselves run as ONE grouped tcgen05 GEMM per matrix (gemm2_sm100.cu,
// kModeGroupM / kModeGroupK) over the ragged layout these kernels produce:
//
//   rows of expert e occupy [offsets[e], offsets[e] + count[e]) of a [capacity, hidden] buffer, every expert segment
//   is padded to a multiple of 256 rows (the CTA-pair tile height), padding rows are zero, offsets[] lives on the
//   device.  capacity = roundup(n_tokens * topk, 256) + 256 * n_experts is a static upper bound, so no shape depends
//   on the routing decision.
//
// Kernels (reference call sites in parentheses):
//   router_topk_fwd/bwd   fused softmax + top-k (+ renormalisation, scaling) producing the dense probs / routing map
//                         the dispatchers consume            (moe/router.py:252-330, moe_utils.topk_softmax_with_capacity)
//   route_count           per-expert stable positions of the routed tokens (block-wide scans, deterministic)
//   build_maps            offsets[], row -> (token, expert) map, (token, expert) -> row map, per-row probability
//   gather_rows           out[r] = scale[r] * in[src[r]]  (permute fwd; combine bwd with the prob and d(prob) fused)
//   gather_sum            out[t] = sum_e y[row(t, e)]     (combine fwd; permute bwd)      (moe_utils.permute/unpermute)
//   dispatch_rows / ...   expert-parallel variants over the EP symmetric arena (rows land directly in the owner's
//                         per-(expert, source) region, counts and source slots travel with them)
//                                                            (token_dispatcher.py:248-706, fused_a2a.py:66-202)
#include "b200_api.h"
#include "common.cuh"

#include <algorithm>

namespace b200 {

namespace {

constexpr int kMaxExpertsPerLane = 8;   // router: E <= 256

// ------------------------------------------------------------------------------------------------ router
struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax warp_argmax(float v, int i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  return {v, i};
}

__global__ void __launch_bounds__(256) router_topk_fwd_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                                              uint8_t* __restrict__ map, int32_t* __restrict__ tokens_per_expert,
                                                              int n, int E, int k, int pre_softmax, float scaling) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* row = logits + static_cast<size_t>(warp) * E;
  float v[kMaxExpertsPerLane];
  bool sel[kMaxExpertsPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxExpertsPerLane; ++j) {
    const int e = lane + 32 * j;
    v[j] = e < E ? row[e] : -INFINITY;
    sel[j] = false;
    mx = fmaxf(mx, v[j]);
  }
  float denom = 1.f;
  if (pre_softmax) {
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxExpertsPerLane; ++j) {
      v[j] = (lane + 32 * j) < E ? __expf(v[j] - mx) : -INFINITY;
      if ((lane + 32 * j) < E) s += v[j];
    }
    denom = warp_sum(s);
  }
  float topv[8];
  int topi[8];
  for (int t = 0; t < k; ++t) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kMaxExpertsPerLane; ++j) {
      const int e = lane + 32 * j;
      if (e < E && !sel[j] && (v[j] > bv || (v[j] == bv && e < bi))) { bv = v[j]; bi = e; }
    }
    const ArgMax w = warp_argmax(bv, bi);
    topv[t] = w.v; topi[t] = w.i;
    if ((w.i & 31) == lane) sel[w.i >> 5] = true;
  }
  // probabilities of the selected experts
  float p[8];
  if (pre_softmax) {
    for (int t = 0; t < k; ++t) p[t] = topv[t] / denom * scaling;
  } else {
    float m2 = topv[0], s = 0.f;
    for (int t = 0; t < k; ++t) { p[t] = __expf(topv[t] - m2); s += p[t]; }
    for (int t = 0; t < k; ++t) p[t] = p[t] / s * scaling;
  }
  float* prow = probs + static_cast<size_t>(warp) * E;
  uint8_t* mrow = map + static_cast<size_t>(warp) * E;
#pragma unroll
  for (int j = 0; j < kMaxExpertsPerLane; ++j) {
    const int e = lane + 32 * j;
    if (e < E) {
      float out = 0.f;
      if (sel[j]) {
        for (int t = 0; t < k; ++t) if (topi[t] == e) out = p[t];
        if (tokens_per_expert != nullptr) atomicAdd(tokens_per_expert + e, 1);
      }
      prow[e] = out;
      mrow[e] = sel[j] ? 1 : 0;
    }
  }
}

__global__ void __launch_bounds__(256) router_topk_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ probs,
                                                              const uint8_t* __restrict__ map, const float* __restrict__ dprobs,
                                                              float* __restrict__ dlogits, int n, int E, int pre_softmax,
                                                              float scaling) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const size_t base = static_cast<size_t>(warp) * E;
  if (!pre_softmax) {
    // out_j = scaling * softmax_over_selected(logits)_j ; p_j = out_j / scaling
    float dot = 0.f;
    for (int e = lane; e < E; e += 32)
      if (map[base + e]) dot += probs[base + e] * dprobs[base + e];
    dot = warp_sum(dot);          // = scaling * sum_j p_j g_j
    for (int e = lane; e < E; e += 32) {
      float d = 0.f;
      if (map[base + e]) d = probs[base + e] * (dprobs[base + e] - dot / scaling);
      dlogits[base + e] = d;
    }
  } else {
    // out_j = scaling * softmax_all(logits)_j for selected j
    float mx = -INFINITY;
    for (int e = lane; e < E; e += 32) mx = fmaxf(mx, logits[base + e]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int e = lane; e < E; e += 32) s += __expf(logits[base + e] - mx);
    s = warp_sum(s);
    float dot = 0.f;
    for (int e = lane; e < E; e += 32)
      if (map[base + e]) dot += __expf(logits[base + e] - mx) / s * scaling * dprobs[base + e];
    dot = warp_sum(dot);
    for (int e = lane; e < E; e += 32) {
      const float sm = __expf(logits[base + e] - mx) / s;
      const float g = map[base + e] ? scaling * dprobs[base + e] : 0.f;
      dlogits[base + e] = sm * (g - dot);
    }
  }
}

// ------------------------------------------------------------------------------------------------ permutation maps
// block e: stable (token order) positions of the tokens routed to local expert e
__global__ void __launch_bounds__(1024) route_count_kernel(const uint8_t* __restrict__ map, int n, int E_total, int e_begin,
                                                           int E_local, int32_t* __restrict__ pos, int32_t* __restrict__ counts) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int t = t0 + tid;
    const bool f = t < n && map[static_cast<size_t>(t) * E_total + e_begin + e] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lp = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(b);
    __syncthreads();
    int wp = 0, tot = 0;
    for (int i = 0; i < 32; ++i) {
      const int c = warp_tot[i];
      if (i < w) wp += c;
      tot += c;
    }
    const int base = base_s;
    if (t < n) pos[static_cast<size_t>(t) * E_local + e] = f ? base + wp + lp : -1;
    __syncthreads();
    if (tid == 0) base_s = base + tot;
    __syncthreads();
  }
  if (tid == 0) counts[e] = base_s;
}

// thread per (token, local expert): row maps + per-row probability; block 0 publishes offsets[]
__global__ void __launch_bounds__(256) build_maps_kernel(const int32_t* __restrict__ pos, const int32_t* __restrict__ counts,
                                                         const float* __restrict__ probs, int n, int E_total, int e_begin,
                                                         int E_local, int pad, int32_t* __restrict__ offsets,
                                                         int32_t* __restrict__ row_slot, float* __restrict__ row_prob,
                                                         int32_t* __restrict__ dest_row) {
  extern __shared__ int offs_s[];
  if (threadIdx.x == 0) {
    int o = 0;
    for (int e = 0; e < E_local; ++e) {
      offs_s[e] = o;
      o += (counts[e] + pad - 1) / pad * pad;
    }
    offs_s[E_local] = o;
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int e = threadIdx.x; e <= E_local; e += blockDim.x) offsets[e] = offs_s[e];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(n) * E_local) return;
  const int t = static_cast<int>(i / E_local), e = static_cast<int>(i - static_cast<int64_t>(t) * E_local);
  const int p = pos[i];
  int r = -1;
  if (p >= 0) {
    r = offs_s[e] + p;
    const int slot = t * E_total + e_begin + e;
    row_slot[r] = slot;
    if (row_prob != nullptr) row_prob[r] = probs != nullptr ? probs[slot] : 1.f;
  }
  dest_row[i] = r;
}

// ------------------------------------------------------------------------------------------------ row movers
// warp per row of the padded layout: out[r] = scale[r] * in[slot[r] / slot_div]   (zero rows where slot[r] < 0).
// kDot: also dot_out[slot[r]] = <in_row, other[r]> / scale[r]   (d prob of the combine backward; other = saved z)
template <bool kScale, bool kDot>
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint16_t* __restrict__ in, const int32_t* __restrict__ row_slot,
                                                          int slot_div, const float* __restrict__ row_scale,
                                                          const int32_t* __restrict__ offsets, int n_groups, int cap, int cols,
                                                          uint16_t* __restrict__ out, const uint16_t* __restrict__ other,
                                                          float* __restrict__ dot_out, int is_bf16) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = min(cap, __ldg(offsets + n_groups));
  const int vecs = cols / 8;
  for (int r = blockIdx.x * warps + warp; r < total; r += gridDim.x * warps) {
    const int slot = __ldg(row_slot + r);
    int4* dst = reinterpret_cast<int4*>(out + static_cast<size_t>(r) * cols);
    if (slot < 0) {
      for (int v = lane; v < vecs; v += 32) dst[v] = make_int4(0, 0, 0, 0);
      continue;
    }
    const int4* src = reinterpret_cast<const int4*>(in + static_cast<size_t>(slot / slot_div) * cols);
    const float sc = kScale ? __ldg(row_scale + r) : 1.f;
    float dot = 0.f;
    for (int v = lane; v < vecs; v += 32) {
      int4 x = ld_nc_v4(src + v);
      if constexpr (kScale || kDot) {
        uint32_t* u = reinterpret_cast<uint32_t*>(&x);
        int4 oth = make_int4(0, 0, 0, 0);
        if constexpr (kDot) oth = ld_nc_v4(reinterpret_cast<const int4*>(other + static_cast<size_t>(r) * cols) + v);
        const uint32_t* ou = reinterpret_cast<const uint32_t*>(&oth);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 f = is_bf16 ? unpack_bf16x2(u[q]) : __half22float2(*reinterpret_cast<__half2*>(&u[q]));
          if constexpr (kDot) {
            const float2 g = is_bf16 ? unpack_bf16x2(ou[q]) : __half22float2(*reinterpret_cast<const __half2*>(&ou[q]));
            dot += f.x * g.x + f.y * g.y;
          }
          if constexpr (kScale) {
            f.x *= sc; f.y *= sc;
            if (is_bf16) u[q] = pack_bf16x2(f.x, f.y);
            else { __half2 h = __floats2half2_rn(f.x, f.y); u[q] = *reinterpret_cast<uint32_t*>(&h); }
          }
        }
      }
      dst[v] = x;
    }
    if constexpr (kDot) {
      dot = warp_sum(dot);
      if (lane == 0) dot_out[slot] = (sc != 0.f) ? dot / sc : 0.f;
    }
  }
}

// block (128 threads) per token: out[t] = sum over local experts of y[dest_row[t, e]] (fp32 accumulation, fixed order)
__global__ void __launch_bounds__(128) gather_sum_kernel(const uint16_t* __restrict__ y, const int32_t* __restrict__ dest_row,
                                                         int n, int E_local, int cols, uint16_t* __restrict__ out, int is_bf16) {
  const int t = blockIdx.x;
  if (t >= n) return;
  const int vecs = cols / 8;
  const int32_t* dr = dest_row + static_cast<size_t>(t) * E_local;
  for (int v0 = threadIdx.x; v0 < vecs; v0 += 128) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < E_local; ++e) {
      const int r = __ldg(dr + e);
      if (r < 0) continue;
      const int4 x = ld_nc_v4(reinterpret_cast<const int4*>(y + static_cast<size_t>(r) * cols) + v0);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = is_bf16 ? unpack_bf16x2(u[q]) : __half22float2(*reinterpret_cast<const __half2*>(&u[q]));
        acc[2 * q] += f.x; acc[2 * q + 1] += f.y;
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (is_bf16) pk[q] = pack_bf16x2(acc[2 * q], acc[2 * q + 1]);
      else { __half2 h = __floats2half2_rn(acc[2 * q], acc[2 * q + 1]); pk[q] = *reinterpret_cast<uint32_t*>(&h); }
    }
    reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * cols)[v0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ------------------------------------------------------------------------------------------------ expert parallel (arena)
// Sources push routed rows straight into the expert owner's staging buffer over NVLink; per-(source, local expert)
// counts, the sender's compact slot index and the router probability travel with the rows.  One epoch flag per
// (owner, source) pair, raised by the last CTA of the sender ("last block" pattern), tells the owner that everything
// from that source has landed.  Staging layout on an owner: [source s][rows of local expert 0 | 1 | ...] (capacity Rs
// rows per source).
struct EpPeers {
  void* stage_x[kMaxPeers];
  int32_t* stage_idx[kMaxPeers];
  float* stage_prob[kMaxPeers];
  int32_t* recv_counts[kMaxPeers];   // [ep][E_local] on each owner
  uint32_t* flags[kMaxPeers];        // [ep] on each owner
};

__global__ void __launch_bounds__(256) ep_dispatch_kernel(const uint16_t* __restrict__ x, const int32_t* __restrict__ send_offsets,
                                                          const int32_t* __restrict__ send_row, const int32_t* __restrict__ counts_send,
                                                          const float* __restrict__ probs, EpPeers peers, int n, int E_total,
                                                          int E_local, int me, int ep, int Rs, int cols, uint32_t epoch,
                                                          int32_t* __restrict__ done_counter) {
  extern __shared__ int pre_s[];      // per global expert: rows of my region on its owner that precede this expert
  for (int ge = threadIdx.x; ge < E_total; ge += blockDim.x) {
    const int d = ge / E_local, le = ge - d * E_local;
    int o = 0;
    for (int j = 0; j < le; ++j) o += counts_send[d * E_local + j];
    pre_s[ge] = o;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // my per-expert counts -> every owner (including myself)
    for (int ge = threadIdx.x; ge < E_total; ge += blockDim.x) {
      const int d = ge / E_local, le = ge - d * E_local;
      peers.recv_counts[d][me * E_local + le] = counts_send[ge];
    }
  }
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vecs = cols / 8;
  const int64_t pairs = static_cast<int64_t>(n) * E_total;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * warps + warp; i < pairs; i += static_cast<int64_t>(gridDim.x) * warps) {
    const int sr = __ldg(send_row + i);
    if (sr < 0) continue;
    const int t = static_cast<int>(i / E_total), ge = static_cast<int>(i - static_cast<int64_t>(t) * E_total);
    const int d = ge / E_local;
    const int row = me * Rs + pre_s[ge] + (sr - __ldg(send_offsets + ge));   // position among my rows for this expert
    const int4* src = reinterpret_cast<const int4*>(x + static_cast<size_t>(t) * cols);
    int4* dst = reinterpret_cast<int4*>(reinterpret_cast<uint16_t*>(peers.stage_x[d]) + static_cast<size_t>(row) * cols);
    for (int v = lane; v < vecs; v += 32) st_na_v4(dst + v, ld_nc_v4(src + v));
    if (lane == 0) {
      peers.stage_idx[d][row] = sr;
      peers.stage_prob[d][row] = probs != nullptr ? __ldg(probs + i) : 1.f;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(done_counter, 1);
    if (old == static_cast<int>(gridDim.x) - 1) {
      __threadfence_system();
      for (int d = 0; d < ep; ++d) st_release_sys(peers.flags[d] + me, epoch);
      *done_counter = 0;
    }
  }
}

// owner: one block.  Waits for every source, then turns the per-(source, expert) counts into the ragged layout:
// offsets[] (256-row padded segments, expert-major, sources in rank order inside a segment) and, for every compact
// row, the staging row it comes from (-1 = padding).  Rows that do not fit the compact capacity are dropped and
// *overflow is raised (the caller re-runs with more slack).
__global__ void __launch_bounds__(1024) ep_recv_layout_kernel(const uint32_t* __restrict__ flags, const int32_t* __restrict__ recv_counts,
                                                              int ep, int E_local, int Rs, int pad, int cap, uint32_t epoch,
                                                              int32_t* __restrict__ offsets, int32_t* __restrict__ counts,
                                                              int32_t* __restrict__ row_stage, int32_t* __restrict__ overflow) {
  __shared__ int cnt[kMaxPeers * 64];
  __shared__ int offs[65];
  if (threadIdx.x < ep) {
    while (static_cast<int32_t>(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) __nanosleep(64);
    __threadfence_system();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ep * E_local; i += blockDim.x)
    cnt[i] = static_cast<int>(ld_relaxed_sys(reinterpret_cast<const uint32_t*>(recv_counts + i)));
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0;
    bool ovf = false;
    for (int le = 0; le < E_local; ++le) {
      offs[le] = o;
      int c = 0;
      for (int s2 = 0; s2 < ep; ++s2) c += cnt[s2 * E_local + le];
      int padded = (c + pad - 1) / pad * pad;
      if (o + padded > cap) {
        ovf = true;
        padded = (cap - o) / pad * pad;
        c = min(c, padded);
      }
      counts[le] = c;
      o += padded;
    }
    offs[E_local] = o;
    if (ovf) *overflow = 1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e <= E_local; e += blockDim.x) offsets[e] = offs[e];
  for (int r = threadIdx.x; r < cap; r += blockDim.x) row_stage[r] = -1;
  __syncthreads();
  for (int le = 0; le < E_local; ++le) {
    const int seg_rows = offs[le + 1] - offs[le];
    int base = offs[le];
    for (int s2 = 0; s2 < ep; ++s2) {
      int sbase = s2 * Rs;
      for (int j = 0; j < le; ++j) sbase += cnt[s2 * E_local + j];
      const int c = cnt[s2 * E_local + le];
      for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const int r = base + i;
        if (r - offs[le] < seg_rows) row_stage[r] = sbase + i;
      }
      base += c;
    }
  }
}

struct EpRet {
  void* ret_x[kMaxPeers];       // [n * k, cols] on each source
  float* ret_f[kMaxPeers];      // optional fp32 side payload (d prob), same indexing
  uint32_t* flags[kMaxPeers];   // [ep] on each source
};

// owner -> sources: compact row r returns to source (row_stage[r] / Rs) at the slot the source named (row_idx[r])
__global__ void __launch_bounds__(256) ep_combine_push_kernel(const uint16_t* __restrict__ z, const int32_t* __restrict__ row_stage,
                                                              const int32_t* __restrict__ row_idx, const float* __restrict__ side,
                                                              const int32_t* __restrict__ offsets, int n_groups, int cap, int Rs,
                                                              int cols, EpRet ret, int me, int ep, uint32_t epoch,
                                                              int32_t* __restrict__ done_counter) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = min(cap, __ldg(offsets + n_groups));
  const int vecs = cols / 8;
  for (int r = blockIdx.x * warps + warp; r < total; r += gridDim.x * warps) {
    const int sr = __ldg(row_stage + r);
    if (sr < 0) continue;
    const int s2 = sr / Rs;
    const int idx = __ldg(row_idx + r);
    const int4* src = reinterpret_cast<const int4*>(z + static_cast<size_t>(r) * cols);
    int4* dst = reinterpret_cast<int4*>(reinterpret_cast<uint16_t*>(ret.ret_x[s2]) + static_cast<size_t>(idx) * cols);
    for (int v = lane; v < vecs; v += 32) st_na_v4(dst + v, ld_nc_v4(src + v));
    if (side != nullptr && lane == 0) ret.ret_f[s2][idx] = __ldg(side + r);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(done_counter, 1);
    if (old == static_cast<int>(gridDim.x) - 1) {
      __threadfence_system();
      for (int s2 = 0; s2 < ep; ++s2) st_release_sys(ret.flags[s2] + me, epoch);
      *done_counter = 0;
    }
  }
}

__global__ void ep_wait_kernel(const uint32_t* __restrict__ flags, int ep, uint32_t epoch) {
  if (threadIdx.x < ep) {
    while (static_cast<int32_t>(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) __nanosleep(64);
    __threadfence_system();
  }
}

// dense d(probs)[t, e] = ret_f[send_row[t, e]] for routed pairs (0 elsewhere)
__global__ void scatter_dprobs_kernel(const float* __restrict__ ret_f, const int32_t* __restrict__ send_row, float* __restrict__ dprobs,
                                      int64_t pairs) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= pairs) return;
  const int r = send_row[i];
  dprobs[i] = r >= 0 ? ret_f[r] : 0.f;
}

}  // namespace

void moe_ep_dispatch(const void* x, const int32_t* send_offsets, const int32_t* send_row, const int32_t* counts_send, const float* probs,
                     void* const* stage_x, int32_t* const* stage_idx, float* const* stage_prob, int32_t* const* recv_counts,
                     uint32_t* const* flags, int n, int E_total, int E_local, int me, int ep, int Rs, int cols, uint32_t epoch,
                     int32_t* done_counter, cudaStream_t s) {
  EpPeers p{};
  for (int i = 0; i < ep; ++i) {
    p.stage_x[i] = stage_x[i]; p.stage_idx[i] = stage_idx[i]; p.stage_prob[i] = stage_prob[i];
    p.recv_counts[i] = recv_counts[i]; p.flags[i] = flags[i];
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t pairs = static_cast<int64_t>(n) * E_total;
  int blocks = static_cast<int>(std::min<int64_t>((pairs + 7) / 8, sms * 4));
  if (blocks < 1) blocks = 1;
  ep_dispatch_kernel<<<blocks, 256, E_total * sizeof(int), s>>>(reinterpret_cast<const uint16_t*>(x), send_offsets, send_row, counts_send, probs, p,
                                                               n, E_total, E_local, me, ep, Rs, cols, epoch, done_counter);
}

void moe_ep_recv_layout(const uint32_t* flags, const int32_t* recv_counts, int ep, int E_local, int Rs, int pad, int cap,
                        uint32_t epoch, int32_t* offsets, int32_t* counts, int32_t* row_stage, int32_t* overflow, cudaStream_t s) {
  ep_recv_layout_kernel<<<1, 1024, 0, s>>>(flags, recv_counts, ep, E_local, Rs, pad, cap, epoch, offsets, counts, row_stage, overflow);
}

void moe_ep_combine_push(const void* z, const int32_t* row_stage, const int32_t* row_idx, const float* side, const int32_t* offsets,
                         int n_groups, int cap, int Rs, int cols, void* const* ret_x, float* const* ret_f, uint32_t* const* flags,
                         int me, int ep, uint32_t epoch, int32_t* done_counter, cudaStream_t s) {
  EpRet r{};
  for (int i = 0; i < ep; ++i) { r.ret_x[i] = ret_x[i]; r.ret_f[i] = ret_f[i]; r.flags[i] = flags[i]; }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = std::min((cap + 7) / 8, sms * 4);
  if (blocks < 1) blocks = 1;
  ep_combine_push_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(z), row_stage, row_idx, side, offsets, n_groups, cap,
                                                Rs, cols, r, me, ep, epoch, done_counter);
}

void moe_ep_wait(const uint32_t* flags, int ep, uint32_t epoch, cudaStream_t s) { ep_wait_kernel<<<1, 32, 0, s>>>(flags, ep, epoch); }

void moe_scatter_dprobs(const float* ret_f, const int32_t* send_row, float* dprobs, int64_t pairs, cudaStream_t s) {
  if (pairs <= 0) return;
  scatter_dprobs_kernel<<<static_cast<int>((pairs + 255) / 256), 256, 0, s>>>(ret_f, send_row, dprobs, pairs);
}

void moe_router_topk_fwd(const float* logits, float* probs, uint8_t* map, int32_t* tokens_per_expert, int n, int E, int k,
                         bool pre_softmax, float scaling, cudaStream_t s) {
  const int blocks = (n * 32 + 255) / 256;
  router_topk_fwd_kernel<<<blocks, 256, 0, s>>>(logits, probs, map, tokens_per_expert, n, E, k, pre_softmax ? 1 : 0, scaling);
}

void moe_router_topk_bwd(const float* logits, const float* probs, const uint8_t* map, const float* dprobs, float* dlogits,
                         int n, int E, bool pre_softmax, float scaling, cudaStream_t s) {
  const int blocks = (n * 32 + 255) / 256;
  router_topk_bwd_kernel<<<blocks, 256, 0, s>>>(logits, probs, map, dprobs, dlogits, n, E, pre_softmax ? 1 : 0, scaling);
}

void moe_build_layout(const uint8_t* map, const float* probs, int n, int E_total, int e_begin, int E_local, int pad,
                      int32_t* pos_scratch, int32_t* counts, int32_t* offsets, int32_t* row_slot, float* row_prob,
                      int32_t* dest_row, int cap, cudaStream_t s) {
  cudaMemsetAsync(row_slot, 0xFF, static_cast<size_t>(cap) * sizeof(int32_t), s);   // -1: padding rows
  route_count_kernel<<<E_local, 1024, 0, s>>>(map, n, E_total, e_begin, E_local, pos_scratch, counts);
  const int64_t pairs = static_cast<int64_t>(n) * E_local;
  const int blocks = static_cast<int>((pairs + 255) / 256);
  build_maps_kernel<<<blocks > 0 ? blocks : 1, 256, (E_local + 1) * sizeof(int), s>>>(pos_scratch, counts, probs, n, E_total, e_begin,
                                                                                     E_local, pad, offsets, row_slot, row_prob, dest_row);
}

void moe_gather_rows(const void* in, const int32_t* row_slot, int slot_div, const float* row_scale, const int32_t* offsets,
                     int n_groups, int cap, int cols, void* out, const void* other, float* dot_out, bool is_bf16, cudaStream_t s) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = min((cap + 7) / 8, sms * 8);
  const uint16_t* i16 = reinterpret_cast<const uint16_t*>(in);
  uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
  const uint16_t* oth = reinterpret_cast<const uint16_t*>(other);
  if (dot_out != nullptr)
    gather_rows_kernel<true, true><<<blocks, 256, 0, s>>>(i16, row_slot, slot_div, row_scale, offsets, n_groups, cap, cols, o16, oth, dot_out, is_bf16);
  else if (row_scale != nullptr)
    gather_rows_kernel<true, false><<<blocks, 256, 0, s>>>(i16, row_slot, slot_div, row_scale, offsets, n_groups, cap, cols, o16, nullptr, nullptr, is_bf16);
  else
    gather_rows_kernel<false, false><<<blocks, 256, 0, s>>>(i16, row_slot, slot_div, nullptr, offsets, n_groups, cap, cols, o16, nullptr, nullptr, is_bf16);
}

void moe_gather_sum(const void* y, const int32_t* dest_row, int n, int E_local, int cols, void* out, bool is_bf16, cudaStream_t s) {
  if (n <= 0) return;
  gather_sum_kernel<<<n, 128, 0, s>>>(reinterpret_cast<const uint16_t*>(y), dest_row, n, E_local, cols,
                                      reinterpret_cast<uint16_t*>(out), is_bf16);
}

}  // namespace b200
