// NVLink / NVSwitch communication kernels that work on a CUDA-IPC "symmetric arena":
// every rank cudaMalloc's one arena, exports it with cudaIpcGetMemHandle, and maps all peers'
// arenas, so kernels address remote HBM with plain ld/st (routed over NVLink 5 through NVSwitch).
// Signalling is flag based: a writer publishes the call epoch with st.release.sys after its
// payload stores; readers spin with ld.acquire.sys.  Epochs are monotonic per slot, so flags
// never need resetting and double buffering by call parity makes slot reuse safe.
//
// Kernels:
//   push_flags_barrier     - device-side barrier across the group (no host sync, no NCCL)
//   bucket_reduce_scatter  - DP gradient reduce-scatter fused with the 1/dp scale and the
//                            fp32->wire cast  (reference param_and_grad_buffer.py:301-355)
//   shard_all_gather_cast  - ZeRO-1 param all-gather fused with the fp32->bf16 cast (:194-240)
//   embed_scatter          - vocab-parallel embedding lookup fused with the sequence-parallel scatter (K14)
//   a2a_scatter_rows       - MoE dispatch/combine: routed rows go straight into the expert
//                            owner's buffer (token_dispatcher.py:512-520), no host-synced splits
#include "b200_api.h"
#include "common.cuh"

namespace b200 {

namespace {

__global__ void barrier_kernel(uint32_t* const* peer_flags_dev, int rank, int size, int slot,
                               uint32_t epoch) {
  // flags layout per rank: [slot][src]
  if (threadIdx.x < size) {
    uint32_t* remote = peer_flags_dev[threadIdx.x] + slot * kMaxPeers + rank;
    __threadfence_system();
    st_release_sys(remote, epoch);
    const uint32_t* mine = peer_flags_dev[rank] + slot * kMaxPeers + threadIdx.x;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) __nanosleep(32);
  }
}

struct PeerPtrs {
  uint32_t* flags[kMaxPeers];
};
__global__ void barrier_kernel_v(PeerPtrs pp, int rank, int size, int slot, uint32_t epoch) {
  if (threadIdx.x < size) {
    uint32_t* remote = pp.flags[threadIdx.x] + slot * kMaxPeers + rank;
    __threadfence_system();
    st_release_sys(remote, epoch);
    const uint32_t* mine = pp.flags[rank] + slot * kMaxPeers + threadIdx.x;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) __nanosleep(32);
  }
  __syncthreads();
}

template <typename W> struct Wire;
template <> struct Wire<float> {
  static __device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    int4 r = make_int4(__float_as_int(v[0]), __float_as_int(v[1]), __float_as_int(v[2]), __float_as_int(v[3]));
    st_na_v4(reinterpret_cast<int4*>(p), r);
  }
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    int4 r = ld_relaxed_sys_v4(reinterpret_cast<const int4*>(p));
    v[0] = __int_as_float(r.x); v[1] = __int_as_float(r.y); v[2] = __int_as_float(r.z); v[3] = __int_as_float(r.w);
  }
};
template <> struct Wire<__nv_bfloat16> {
  static __device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 r = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(r.x), "r"(r.y) : "memory");
  }
  static __device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 r;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};

// grid = n_ctas; CTA c owns slice c of every rank's shard, so rank q / CTA c only ever waits for
// CTA c of the other ranks (tile-granular overlap, no grid-wide sync).
//
// Bandwidth: a thread that loads, scales and stores one float4 per iteration keeps ONE 16-byte request in flight, and
// 32 CTAs x 512 threads x 16 B per ~1 us HBM round trip is ~0.25 TB/s -- what round 1 measured (219 GB/s).  Each thread
// therefore issues kU independent loads before the first dependent store (phase 1) and the loads of all sources of kR
// elements before the first add (phase 2): the NVLink stores are posted, so with 4-8 requests per thread in flight the
// kernel is bound by the link, not by latency.
constexpr int kRsUnroll = 8;   // phase 1: independent 16-byte loads per thread
constexpr int kRsRedUnroll = 2;  // phase 2: elements x sources (<= 8) loads per thread

template <typename W, int kSize>   // kSize: group size known at compile time (2 / 4 / 8) or 0 = runtime
__global__ void __launch_bounds__(512) bucket_rs_kernel(RsArgs a) {
  const int64_t shard = a.n / a.size;
  const int64_t per_cta = ((shard + gridDim.x - 1) / gridDim.x + 3) / 4 * 4;
  const int64_t lo = min(shard, static_cast<int64_t>(blockIdx.x) * per_cta);
  const int64_t hi = min(shard, lo + per_cta);
  const int64_t step = static_cast<int64_t>(blockDim.x) * 4;
  const float scale = a.scale;
  // phase 1: push my (scaled, cast) contribution for every peer's shard slice
  for (int d = 1; d < a.size; ++d) {
    const int peer = (a.rank + d) % a.size;
    const float* src = a.src + peer * shard;
    W* dst = reinterpret_cast<W*>(a.peer_stage[peer]) + static_cast<int64_t>(a.rank) * shard;
    int64_t i = lo + threadIdx.x * 4;
    for (; i + (kRsUnroll - 1) * step < hi; i += kRsUnroll * step) {
      float4 g[kRsUnroll];
#pragma unroll
      for (int u = 0; u < kRsUnroll; ++u) g[u] = __ldcs(reinterpret_cast<const float4*>(src + i + u * step));
#pragma unroll
      for (int u = 0; u < kRsUnroll; ++u) {
        float v[4] = {g[u].x * scale, g[u].y * scale, g[u].z * scale, g[u].w * scale};
        Wire<W>::store4(dst + i + u * step, v);
      }
    }
    for (; i < hi; i += step) {
      float4 g = __ldcs(reinterpret_cast<const float4*>(src + i));
      float v[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
      Wire<W>::store4(dst + i, v);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      st_release_sys(a.peer_flags[peer] + static_cast<int64_t>(a.rank) * gridDim.x + blockIdx.x, a.epoch);
  }
  // phase 2: reduce my shard slice in fixed source order (deterministic)
  if (threadIdx.x < a.size && threadIdx.x != a.rank) {
    const uint32_t* f = a.peer_flags[a.rank] + static_cast<int64_t>(threadIdx.x) * gridDim.x + blockIdx.x;
    while (static_cast<int32_t>(ld_acquire_sys(f) - a.epoch) < 0) __nanosleep(64);
  }
  __threadfence_system();
  __syncthreads();
  const W* stage = reinterpret_cast<const W*>(a.peer_stage[a.rank]);
  const float* mine = a.src + a.rank * shard;
  auto reduce_at = [&](int64_t i, float (&acc)[4]) {
    if constexpr (kSize > 0) {
      // every source's load is issued before the first add: kSize requests in flight per element group
      float v[kSize][4];
#pragma unroll
      for (int s_ = 0; s_ < kSize; ++s_) {
        if (s_ == a.rank) {
          float4 g = __ldcs(reinterpret_cast<const float4*>(mine + i));
          v[s_][0] = g.x * scale; v[s_][1] = g.y * scale; v[s_][2] = g.z * scale; v[s_][3] = g.w * scale;
        } else {
          Wire<W>::load4(stage + static_cast<int64_t>(s_) * shard + i, v[s_]);
        }
      }
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
#pragma unroll
      for (int s_ = 0; s_ < kSize; ++s_) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += v[s_][j];
      }
    } else {
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      for (int s_ = 0; s_ < a.size; ++s_) {
        float v[4];
        if (s_ == a.rank) {
          float4 g = __ldcs(reinterpret_cast<const float4*>(mine + i));
          v[0] = g.x * scale; v[1] = g.y * scale; v[2] = g.z * scale; v[3] = g.w * scale;
        } else {
          Wire<W>::load4(stage + static_cast<int64_t>(s_) * shard + i, v);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += v[j];
      }
    }
  };
  int64_t i = lo + threadIdx.x * 4;
  for (; i + (kRsRedUnroll - 1) * step < hi; i += kRsRedUnroll * step) {
    float acc[kRsRedUnroll][4];
#pragma unroll
    for (int u = 0; u < kRsRedUnroll; ++u) reduce_at(i + u * step, acc[u]);
#pragma unroll
    for (int u = 0; u < kRsRedUnroll; ++u)
      __stcs(reinterpret_cast<float4*>(a.dst + i + u * step), make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]));
  }
  for (; i < hi; i += step) {
    float acc[4];
    reduce_at(i, acc);
    __stcs(reinterpret_cast<float4*>(a.dst + i), make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

template <typename D>
__global__ void __launch_bounds__(512) shard_ag_kernel(AgArgs a) {
  const int64_t per_cta = ((a.shard_n + gridDim.x - 1) / gridDim.x + 3) / 4 * 4;
  const int64_t lo = min(a.shard_n, static_cast<int64_t>(blockIdx.x) * per_cta);
  const int64_t hi = min(a.shard_n, lo + per_cta);
  for (int d = 0; d < a.size; ++d) {
    const int peer = (a.rank + d) % a.size;
    D* dst = reinterpret_cast<D*>(a.peer_dst[peer]) + static_cast<int64_t>(a.rank) * a.shard_n;
    for (int64_t i = lo + threadIdx.x * 4; i < hi; i += blockDim.x * 4) {
      float4 g = *reinterpret_cast<const float4*>(a.src_shard + i);
      float v[4] = {g.x, g.y, g.z, g.w};
      Wire<D>::store4(dst + i, v);
    }
    if (peer != a.rank) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0)
        st_release_sys(a.peer_flags[peer] + static_cast<int64_t>(a.rank) * gridDim.x + blockIdx.x, a.epoch);
    }
  }
  // kernel completion == every peer's slice c has landed here
  if (threadIdx.x < a.size && threadIdx.x != a.rank) {
    const uint32_t* f = a.peer_flags[a.rank] + static_cast<int64_t>(threadIdx.x) * gridDim.x + blockIdx.x;
    while (static_cast<int32_t>(ld_acquire_sys(f) - a.epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
}

// one warp per routed row; dst_rank < 0 drops the row (capacity overflow)
__global__ void __launch_bounds__(256) a2a_scatter_kernel(A2aArgs a) {
  const int warps_per_cta = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int row_vec = a.cols * 2 / 16;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; r < a.rows;
       r += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
    const int pr = a.dst_rank[r];
    if (pr < 0) continue;
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const uint16_t*>(a.src) + r * a.cols);
    int4* dst = reinterpret_cast<int4*>(reinterpret_cast<uint16_t*>(a.peer_buf[pr]) +
                                        static_cast<int64_t>(a.dst_row[r]) * a.cols);
    for (int v = lane; v < row_vec; v += 32) st_na_v4(dst + v, ld_nc_v4(src + v));
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.size && threadIdx.x != a.rank)
    st_release_sys(a.peer_flags[threadIdx.x] + static_cast<int64_t>(a.rank) * gridDim.x + blockIdx.x, a.epoch);
  if (threadIdx.x < a.size && threadIdx.x != a.rank) {
    const uint32_t* f = a.peer_flags[a.rank] + static_cast<int64_t>(threadIdx.x) * gridDim.x + blockIdx.x;
    while (static_cast<int32_t>(ld_acquire_sys(f) - a.epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
}

// K14: vocabulary-parallel embedding fused with the sequence-parallel scatter.  Every token's row lives on exactly one rank
// of the group, so "masked lookup -> reduce-scatter" (reference layers.py:242-273) needs no reduction at all: the owner
// writes the row straight into the [s/tp, b, h] output of the rank that owns that sequence position (plain st over NVLink,
// one warp per token).  CTA c signals CTA c of every peer when its rows are out and waits for theirs: when the kernel
// retires on a rank, every (peer, CTA) pair has delivered, i.e. the output is complete.
__global__ void __launch_bounds__(256) embed_scatter_kernel(EmbedArgs a) {
  const int warps_per_cta = blockDim.x / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int row_vec = a.h * 2 / 16;
  const int s_local = a.s / a.size;
  const int64_t n = static_cast<int64_t>(a.b) * a.s;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * warps_per_cta + warp; t < n;
       t += static_cast<int64_t>(gridDim.x) * warps_per_cta) {
    const int64_t id = a.tokens[t];
    if (id < a.vocab_start || id >= a.vocab_end) continue;
    const int bi = static_cast<int>(t / a.s), si = static_cast<int>(t % a.s);
    const int dest = si / s_local;
    const int64_t row = static_cast<int64_t>(si % s_local) * a.b + bi;
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const uint16_t*>(a.weight) + (id - a.vocab_start) * a.h);
    int4* dst = reinterpret_cast<int4*>(reinterpret_cast<uint16_t*>(a.peer_out[dest]) + row * a.h);
    for (int v = lane; v < row_vec; v += 32) st_na_v4(dst + v, ld_nc_v4(src + v));
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.size && threadIdx.x != a.rank)
    st_release_sys(a.peer_flags[threadIdx.x] + static_cast<int64_t>(a.rank) * gridDim.x + blockIdx.x, a.epoch);
  if (threadIdx.x < a.size && threadIdx.x != a.rank) {
    const uint32_t* f = a.peer_flags[a.rank] + static_cast<int64_t>(threadIdx.x) * gridDim.x + blockIdx.x;
    while (static_cast<int32_t>(ld_acquire_sys(f) - a.epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
}

__global__ void wait_flags_kernel(const uint32_t* flags, int n, uint32_t target) {
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    while (static_cast<int32_t>(ld_acquire_sys(flags + i) - target) < 0) __nanosleep(64);
}

__global__ void spin_kernel_impl(int64_t cycles) {
  const int64_t t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}

__global__ void write_flag_kernel(uint32_t* flag, uint32_t value) {
  __threadfence_system();
  st_release_sys(flag, value);
}

}  // namespace

void push_flags_barrier(uint32_t* const* peer_flags, int rank, int size, int slot, uint32_t epoch,
                        cudaStream_t s) {
  PeerPtrs pp{};
  for (int i = 0; i < size; ++i) pp.flags[i] = peer_flags[i];
  barrier_kernel_v<<<1, 32, 0, s>>>(pp, rank, size, slot, epoch);
}

void bucket_reduce_scatter(const RsArgs& a, cudaStream_t s) {
  const int ctas = a.n_ctas > 0 ? a.n_ctas : 32;
#define B200_RS_LAUNCH(W)                                                              \
  switch (a.size) {                                                                   \
    case 2: bucket_rs_kernel<W, 2><<<ctas, 512, 0, s>>>(a); break;                    \
    case 4: bucket_rs_kernel<W, 4><<<ctas, 512, 0, s>>>(a); break;                    \
    case 8: bucket_rs_kernel<W, 8><<<ctas, 512, 0, s>>>(a); break;                    \
    default: bucket_rs_kernel<W, 0><<<ctas, 512, 0, s>>>(a);                          \
  }
  if (a.wire_dtype == 2) { B200_RS_LAUNCH(float) } else { B200_RS_LAUNCH(__nv_bfloat16) }
#undef B200_RS_LAUNCH
}

void shard_all_gather_cast(const AgArgs& a, cudaStream_t s) {
  const int ctas = a.n_ctas > 0 ? a.n_ctas : 32;
  if (a.dst_dtype == 2) shard_ag_kernel<float><<<ctas, 512, 0, s>>>(a);
  else shard_ag_kernel<__nv_bfloat16><<<ctas, 512, 0, s>>>(a);
}

void a2a_scatter_rows(const A2aArgs& a, cudaStream_t s) {
  const int ctas = a.n_ctas > 0 ? a.n_ctas : 32;
  a2a_scatter_kernel<<<ctas, 256, 0, s>>>(a);
}

void embed_scatter(const EmbedArgs& a, cudaStream_t s) {
  const int ctas = a.n_ctas > 0 ? a.n_ctas : 32;
  embed_scatter_kernel<<<ctas, 256, 0, s>>>(a);
}

void wait_flags(const uint32_t* flags, int n, uint32_t target, cudaStream_t s) {
  wait_flags_kernel<<<1, 128, 0, s>>>(flags, n, target);
}

void spin_kernel(int64_t cycles, cudaStream_t s) { spin_kernel_impl<<<1, 1, 0, s>>>(cycles); }

void write_flag(uint32_t* flag, uint32_t value, cudaStream_t s) {
  write_flag_kernel<<<1, 1, 0, s>>>(flag, value);
}

}  // namespace b200
