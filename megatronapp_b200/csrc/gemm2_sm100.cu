// 2-CTA (cta_group::2) persistent tcgen05 GEMM for sm_100a, with fused collectives and grouped (MoE) modes.
//
// A CTA pair (cluster 2x1x1, same TPC) computes one 256 x 256 output tile per step:
//   * CTA r loads rows [128r, 128r+128) of the A tile and rows [128r, 128r+128) of the B tile
//     (its half of N) -> each SM stages 32 KB per k-block instead of 48 KB, and the tensor core
//     pair fetches the other half of B from the peer's smem.  Per-SM smem traffic drops from
//     ~192 B/clk (reads + TMA writes, measured limiter of the 1-CTA kernel) to ~128 B/clk.
//   * one elected thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256);
//     tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs' mbarriers.
//   * each CTA drains its own 128 TMEM lanes with its own epilogue warps.
// Six 32 KB stages (192 KB) + 2 x 16 KB epilogue staging; TMEM 2 x 256 columns per CTA.
//
// Modes (template kMode):
//   kModePlain    D = A B^T (+bias, +GELU, +pre-activation) / fp32 store / fp32 TMA reduce-add (wgrad)
//   kModeAG       all-gather(A rows) -> GEMM.  The first n_comm_ctas/2 clusters push this rank's activation shard
//                 into every peer's gather buffer over NVLink (TMA bulk stores to IPC-mapped peer memory) and publish
//                 per-row-block epoch flags (st.release.sys); both producers of a GEMM pair wait (ld.acquire.sys) on
//                 the flag of the 128-row block they are about to load.  Own shard first, then arrival order.
//                 bias / GELU / pre-activation stay in the epilogue (reference layers.py:429-443 + bias_gelu fusion).
//   kModeRS       GEMM -> reduce-scatter(rows).  Tiles of rows owned by a peer leave as bf16 partials through TMA
//                 stores into the owner's staging buffer (flag published one tile later, when the bulk group has
//                 retired); the tiles of this rank's OWN rows are visited last and their epilogue folds the reduction
//                 in: it waits for the tp-1 peer partials of that tile (already landed in the common case), adds
//                 them in fp32 to the accumulator it just read from TMEM, adds the bias and stores the final bf16
//                 output.  No separate reduce phase, no serialised tail, fixed summation order (deterministic).
//                 (reference layers.py:1200-1217; dgrad dual :496-507 with B MN-major)
//   kModeAGK      all-gather along the REDUCTION dimension -> GEMM: the column-parallel wgrad dW = dY^T . AG(X) under sequence
//                 parallelism.  Comm clusters push the local X shard to the peers exactly as in kModeAG; the producers walk
//                 the token (K) axis own shard first, then in arrival order, and wait on the 128-token block flag before
//                 loading B from the gather buffer.  Replaces the NCCL re-gather of X that used to run next to the dgrad
//                 kernel and fight it for NVLink and SMs (reference layers.py:461-471, 509-541).
//   kModeGroupM   grouped GEMM over a 256-row-padded ragged row layout (MoE experts, forward + dgrad): m-block ->
//                 group through a device-side offsets array (no host sync), B/bias selected per group, optional
//                 per-row scale (router probability) in the epilogue.  (reference moe/experts.py:90-265 gmm)
//   kModeGroupK   grouped wgrad: per group D_g[M,N] += A_g^T B_g with the K range of group g taken from the same
//                 offsets array; fp32 TMA reduce-add into main_grad.
#include "b200_api.h"
#include "common.cuh"

#include <mutex>
#include <string>

namespace b200 {

namespace {

constexpr int BM = 128;         // rows per CTA (256 per pair)
constexpr int BN = 256;         // tile N (128 rows of B per CTA)
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int kStagesDefault = 6;
enum : int { kModePlain = 0, kModeAG = 1, kModeRS = 2, kModeGroupM = 3, kModeGroupK = 4, kModeAGK = 5 };
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kGroupN2 = 8;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address
constexpr int kBulkChunk2 = 32768;            // comm-CTA TMA bulk copy granule
constexpr int kBulkStages2 = 6;


struct P2 {
  int M, N, K, ldd;
  void* D;
  void* D_pre;
  const void* bias;
  int act, is_bf16;
  int num_m_blocks, num_n_blocks, num_k_blocks;   // m blocks of 256
  // chunked tile order (comm modes): chunk(i) = (chunk_base + chunk_dir * i) mod n_chunks
  int n_chunks, m_blocks_per_chunk, chunk_base, chunk_dir;
  // fused collectives
  int tp_rank, tp_size, n_comm_ctas;
  uint32_t epoch;
  int rows_per_rank;
  int col_splits;
  const void* A_local;
  int lda_local;
  int ag_cols;                    // columns of the gathered matrix (kModeAG: K, kModeAGK: N)
  void* peer_buf[kMaxPeers];
  uint32_t* peer_flags[kMaxPeers];
  // grouped
  const int32_t* group_offsets;   // [n_groups + 1] row offsets, multiples of 256 (device memory)
  int n_groups;
  int group_stride;               // GroupM: rows of B (K-major) or K rows (MN-major) per group; GroupK: D rows per group
  const float* row_scale;         // GroupM: optional per-row multiplier applied after bias/activation
};

struct PeerMaps2 {
  CUtensorMap m[kMaxPeers];
};

// Shared-memory plan: six 32 KB operand stages + two 16 KB epilogue staging buffers.  (Measured at TP4: giving the
// reduce-scatter mode 6 staging buffers / 4 stages does NOT speed up the peer pushes -- 50 MB of 128-byte row segments
// leave a GPU at ~400 GB/s whatever the bytes in flight per SM, NCCL's ring moves the same bytes at ~470 GB/s -- so the
// link-bound attention-projection case stays transfer-bound and the plan is the same for every mode.)
template <int kMode>
struct Smem2T {
  static constexpr int kStages = kStagesDefault;
  static constexpr int kCBufs = 2;
  static constexpr int kA = BM * BK * 2;          // 16 KB
  static constexpr int kB = (BN / 2) * BK * 2;    // 16 KB (this CTA's half of the B tile)
  static constexpr int kC = BM * 64 * 2;          // 16 KB
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAOff + kStages * kA;
  static constexpr int kCOff = kBOff + kStages * kB;
  static constexpr int kBarOff = kCOff + kCBufs * kC;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // non-.aligned forms: some warps reach this point with lanes arriving from divergent role code
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const void* tmap, uint64_t* leader_bar_local_alias, void* smem_dst,
                                                int32_t c0, int32_t c1) {
  // data lands in this CTA's smem; the transaction bytes are credited to the leader CTA's mbarrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(leader_bar_local_alias) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}

__device__ __forceinline__ float gelu2(float x) {
  return 0.5f * x * (1.f + tanhf(0.79788456f * x * (1.f + 0.044715f * x * x)));
}

// One unit of work of a CTA pair.  Every role (producer / MMA / epilogue) of both CTAs derives it from the tile index
// with the same arithmetic (and, in the grouped modes, the same device-side offsets), so the roles stay in lock step.
struct Tile2 {
  int m_blk, n_blk;     // output tile (m in units of 256 rows)
  int kb0, kb1;         // k-block range
  int group;            // grouped modes: group (expert) index, else 0
  bool valid;
};

template <int kMode>
__device__ __forceinline__ Tile2 tile_info(const P2& p, int tile) {
  Tile2 t;
  t.kb0 = 0; t.kb1 = p.num_k_blocks; t.group = 0; t.valid = true;
  if constexpr (kMode == kModeGroupK) {
    const int per_group = p.num_m_blocks * p.num_n_blocks;
    t.group = tile / per_group;
    const int r = tile - t.group * per_group;
    t.m_blk = r / p.num_n_blocks;
    t.n_blk = r - t.m_blk * p.num_n_blocks;
    const int r0 = __ldg(p.group_offsets + t.group), r1 = __ldg(p.group_offsets + t.group + 1);
    t.kb0 = r0 / BK; t.kb1 = r1 / BK;
    t.valid = t.kb1 > t.kb0;
    return t;
  }
  const int tiles_per_chunk = p.m_blocks_per_chunk * p.num_n_blocks;
  const int ci = tile / tiles_per_chunk;
  int r = tile - ci * tiles_per_chunk;
  int chunk = (p.chunk_base + p.chunk_dir * ci) % p.n_chunks;
  if (chunk < 0) chunk += p.n_chunks;
  const int group_tiles = p.m_blocks_per_chunk * kGroupN2;
  const int g = r / group_tiles;
  r -= g * group_tiles;
  const int n_in_group = min(kGroupN2, p.num_n_blocks - g * kGroupN2);
  const int m_local = r / n_in_group;
  t.n_blk = g * kGroupN2 + (r - m_local * n_in_group);
  t.m_blk = chunk * p.m_blocks_per_chunk + m_local;
  t.valid = t.m_blk < p.num_m_blocks;          // the last L2 band may be partial
  if constexpr (kMode == kModeGroupM) {
    const int row0 = t.m_blk * 256;
    const int total = __ldg(p.group_offsets + p.n_groups);
    t.valid = t.valid && row0 < total;
    int e = 0;
    if (t.valid) {
      while (e + 1 < p.n_groups && __ldg(p.group_offsets + e + 1) <= row0) ++e;
    }
    t.group = e;
  }
  return t;
}

template <bool kAMn, bool kBMn, int kEpi, int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_local,
             const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_d,
             const __grid_constant__ CUtensorMap tmap_d_pre, const __grid_constant__ PeerMaps2 peer_maps, const P2 p) {
  using L = Smem2T<kMode>;
  constexpr int kStages = L::kStages;
  constexpr int kCBufs = L::kCBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem + L::kAOff;
  uint8_t* smem_b = smem + L::kBOff;
  uint8_t* smem_c = smem + L::kCOff;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ------------------------------------------------------------------ comm CTAs (all-gather push over NVLink)
  if constexpr (kMode == kModeAG || kMode == kModeAGK) {
    if (static_cast<int>(blockIdx.x) < p.n_comm_ctas) {
      // work item = (row block of 128 rows, destination, column split); all remote destinations first (they gate
      // the peers' GEMM tiles), the local copy (d == 0: complete gathered view for the caller's wgrad) last
      const int rb_per_rank = p.rows_per_rank / BM;
      const int per_dst = rb_per_rank * p.col_splits;
      const int n_items = per_dst * (kMode == kModeAGK ? p.tp_size - 1 : p.tp_size);   // AG-K reads its own shard in place
      const int row_bytes = p.ag_cols * 2;
      const int split_bytes = row_bytes / p.col_splits;  // multiple of 16 (checked on host)
      const bool contiguous = (p.lda_local == p.ag_cols) && (p.col_splits == 1) && ((BM * row_bytes) % kBulkChunk2 == 0);
      if (contiguous) {
        // TMA bulk pipeline: thread 0 streams gmem -> smem stages, thread 32 streams smem -> peer gmem
        uint64_t* c_full = reinterpret_cast<uint64_t*>(smem + kBulkStages2 * kBulkChunk2);
        uint64_t* c_empty = c_full + kBulkStages2;
        if (threadIdx.x == 0) {
          for (int i = 0; i < kBulkStages2; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 1); }
          mbar_fence_init();
        }
        __syncthreads();
        const int chunks_per_item = BM * row_bytes / kBulkChunk2;
        if (threadIdx.x == 0) {
          int stage = 0; uint32_t phase = 0;
          for (int item = blockIdx.x; item < n_items; item += p.n_comm_ctas) {
            const int rb = (item % per_dst);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.A_local) + static_cast<size_t>(rb) * BM * row_bytes;
            for (int c = 0; c < chunks_per_item; ++c) {
              mbar_wait(&c_empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&c_full[stage], kBulkChunk2);
              bulk_load_1d(smem + stage * kBulkChunk2, src + static_cast<size_t>(c) * kBulkChunk2, kBulkChunk2, &c_full[stage]);
              if (++stage == kBulkStages2) { stage = 0; phase ^= 1; }
            }
          }
        } else if (threadIdx.x == 32) {
          int stage = 0; uint32_t phase = 0;
          int prev_stage = -1;
          for (int item = blockIdx.x; item < n_items; item += p.n_comm_ctas) {
            const int di = item / per_dst;                 // 0 .. tp-1
            const int d = (di + 1) % p.tp_size;            // 1, 2, ..., tp-1, 0
            const int rb = item % per_dst;
            const int peer = (p.tp_rank + d) % p.tp_size;
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.peer_buf[peer]) +
                           (static_cast<size_t>(p.tp_rank) * p.rows_per_rank + rb * BM) * row_bytes;
            for (int c = 0; c < chunks_per_item; ++c) {
              mbar_wait(&c_full[stage], phase);
              bulk_store_1d(dst + static_cast<size_t>(c) * kBulkChunk2, smem + stage * kBulkChunk2, kBulkChunk2);
              tma_store_commit();
              if (prev_stage >= 0) {          // the previous store has finished *reading* its stage
                tma_store_wait_read<1>();
                mbar_arrive(&c_empty[prev_stage]);
              }
              prev_stage = stage;
              if (++stage == kBulkStages2) { stage = 0; phase ^= 1; }
            }
            tma_store_wait_all<0>();           // payload of this item is complete at the destination
            if (prev_stage >= 0) { mbar_arrive(&c_empty[prev_stage]); prev_stage = -1; }
            if (d != 0) {
              fence_proxy_async_all();
              __threadfence_system();
              st_release_sys(p.peer_flags[peer] + (static_cast<size_t>(p.tp_rank) * rb_per_rank + rb), p.epoch);
            }
          }
        }
        return;
      }
      const int vec_per_row = split_bytes / 16;
      for (int item = blockIdx.x; item < n_items; item += p.n_comm_ctas) {
        const int di = item / per_dst;
        const int d = (di + 1) % p.tp_size;
        const int r2 = item % per_dst;
        const int rb = r2 / p.col_splits;
        const int cs = r2 % p.col_splits;
        const int peer = (p.tp_rank + d) % p.tp_size;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.A_local) +
                             static_cast<size_t>(rb) * BM * p.lda_local * 2 + cs * split_bytes;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.peer_buf[peer]) +
                       (static_cast<size_t>(p.tp_rank) * p.rows_per_rank + rb * BM) * row_bytes + cs * split_bytes;
        const int total_vec = BM * vec_per_row;
        for (int v = threadIdx.x; v < total_vec; v += 8 * kThreads) {
          int4 regs[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int vv = v + u * kThreads;
            if (vv < total_vec) {
              const int r = vv / vec_per_row, c = vv - r * vec_per_row;
              regs[u] = ld_nc_v4(reinterpret_cast<const int4*>(src + static_cast<size_t>(r) * p.lda_local * 2 + c * 16));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int vv = v + u * kThreads;
            if (vv < total_vec) {
              const int r = vv / vec_per_row, c = vv - r * vec_per_row;
              st_na_v4(reinterpret_cast<int4*>(dst + static_cast<size_t>(r) * row_bytes + c * 16), regs[u]);
            }
          }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && d != 0) {
          st_release_sys(p.peer_flags[peer] + (static_cast<size_t>(p.tp_rank) * rb_per_rank + rb) * p.col_splits + cs,
                         p.epoch);
        }
      }
      return;
    }
  }

  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int comm_clusters = (kMode == kModeAG || kMode == kModeAGK) ? (p.n_comm_ctas >> 1) : 0;
  const int cluster_id = static_cast<int>(blockIdx.x >> 1) - comm_clusters;
  const int n_clusters = static_cast<int>(gridDim.x >> 1) - comm_clusters;
  const int num_tiles = (kMode == kModeGroupK) ? p.n_groups * p.num_m_blocks * p.num_n_blocks
                                               : p.n_chunks * p.m_blocks_per_chunk * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    if constexpr (kMode == kModeAG || kMode == kModeAGK) tma_prefetch_desc(&tmap_a_local);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);     // one arrive per CTA's producer (leader's copy is the live one)
      mbar_init(&empty_bar[i], 1);    // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);    // multicast tcgen05.commit
      mbar_init(&tmem_empty_bar[i], 8);   // 4 epilogue warps x 2 CTAs (leader's copy is the live one)
    }
    mbar_fence_init();
  }
  cluster_sync_all();                    // peer barriers are initialised before any remote arrive
  if (warp_idx == 2) tmem_alloc2<2 * BN>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kPairStageBytes = 2 * (L::kA + L::kB);

  if (warp_idx == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const Tile2 t = tile_info<kMode>(p, tile);
        if (!t.valid) continue;
        int m0 = t.m_blk * 256 + cta_rank * BM;
        int n0 = t.n_blk * BN + cta_rank * (BN / 2);
        int b_row_off = 0, b_k_off = 0;     // grouped: which expert's weight
        if constexpr (kMode == kModeGroupM) {
          if (!kBMn) b_row_off = t.group * p.group_stride;
          else b_k_off = t.group * p.group_stride;
        }
        const CUtensorMap* ta = &tmap_a;
        if constexpr (kMode == kModeAG) {
          const int src = m0 / p.rows_per_rank;
          if (src == p.tp_rank) {
            ta = &tmap_a_local;
            m0 -= src * p.rows_per_rank;
          } else {
            const int rb_per_rank = p.rows_per_rank / BM;
            const int rb = (m0 - src * p.rows_per_rank) / BM;
            const uint32_t* flag = p.peer_flags[p.tp_rank] + (static_cast<size_t>(src) * rb_per_rank + rb) * p.col_splits;
            for (int cs = 0; cs < p.col_splits; ++cs) {
              while (static_cast<int32_t>(ld_acquire_sys(flag + cs) - p.epoch) < 0) __nanosleep(64);
            }
            fence_proxy_async_all();
          }
        }
        for (int kbi = t.kb0; kbi < t.kb1; ++kbi) {
          int kb = kbi;
          const CUtensorMap* tb = &tmap_b;
          int kb_b = kbi;                      // k-block coordinate inside the B tensor map
          if constexpr (kMode == kModeAGK) {
            // token axis: own shard first, then rank-1, rank-2, ... (the order the peers' pushes arrive in)
            const int kb_per_rank = p.rows_per_rank / BK;
            const int hop = kbi / kb_per_rank, kin = kbi - hop * kb_per_rank;
            int src = (p.tp_rank - hop) % p.tp_size;
            if (src < 0) src += p.tp_size;
            kb = src * kb_per_rank + kin;
            kb_b = kb;
            if (src == p.tp_rank) {
              tb = &tmap_a_local;
              kb_b = kin;
            } else if ((kin & 1) == 0) {       // a flag covers 128 token rows = two k-blocks
              const uint32_t* flag = p.peer_flags[p.tp_rank] +
                                     (static_cast<size_t>(src) * (p.rows_per_rank / BM) + (kin >> 1)) * p.col_splits;
              for (int cs = 0; cs < p.col_splits; ++cs) {
                while (static_cast<int32_t>(ld_acquire_sys(flag + cs) - p.epoch) < 0) __nanosleep(64);
              }
              fence_proxy_async_all();
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], kPairStageBytes);
          else mbar_arrive_remote(&full_bar[stage], 0);
          uint8_t* sa = smem_a + stage * L::kA;
          uint8_t* sb = smem_b + stage * L::kB;
          if constexpr (!kAMn) {
            tma_load_2d_2sm(ta, &full_bar[stage], sa, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_2sm(ta, &full_bar[stage], sa + j * (BK * 128), m0 + j * 64, kb * BK);
          }
          if constexpr (!kBMn) {
            tma_load_2d_2sm(tb, &full_bar[stage], sb, kb_b * BK, n0 + b_row_off);
          } else {
#pragma unroll
            for (int j = 0; j < (BN / 2) / 64; ++j)
              tma_load_2d_2sm(tb, &full_bar[stage], sb + j * (BK * 128), n0 + j * 64, kb_b * BK + b_k_off);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, 256, BN, kAMn, kBMn);
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const Tile2 t = tile_info<kMode>(p, tile);
        if (!t.valid) continue;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * L::kA);
          const uint32_t sb = smem_u32(smem_b + stage * L::kB);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = kAMn ? make_smem_desc_sw128(sa + k * UK * 128, 1024, BK * 128)
                                     : make_smem_desc_sw128(sa + k * UK * 2, 1024, 0);
            const uint64_t db = kBMn ? make_smem_desc_sw128(sb + k * UK * 128, 1024, BK * 128)
                                     : make_smem_desc_sw128(sb + k * UK * 2, 1024, 0);
            umma2_f16_ss(tmem_d, da, db, idesc, (kb != t.kb0 || k != 0) ? 1u : 0u);
          }
          umma2_commit_mc(&empty_bar[stage]);
          if (kb == t.kb1 - 1) umma2_commit_mc(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp_idx >= 4) {
    // ================================================================ epilogue (both CTAs, own TMEM lanes)
    const int ew = warp_idx - 4;
    const int epi_tid = threadIdx.x - 128;
    const int row_in_tile = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    uint32_t* pending_flag = nullptr;   // RS: flag of the previous pushed tile, not yet published
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const Tile2 t = tile_info<kMode>(p, tile);
      if (!t.valid) continue;
      const int m_blk = t.m_blk, n_blk = t.n_blk;
      const int m0 = m_blk * 256 + cta_rank * BM;
      // RS bookkeeping: who owns these rows, and where they sit inside the owner's shard
      int owner = 0, lrow0 = m0;
      bool own = true;
      if constexpr (kMode == kModeRS) {
        owner = m0 / p.rows_per_rank;
        lrow0 = m0 - owner * p.rows_per_rank;
        own = owner == p.tp_rank;
        if (own) {
          // Publish the flag still held for the previous (pushed) tile BEFORE blocking on peers: every rank reaches
          // its own rows holding one unpublished flag for its predecessor - waiting first would deadlock the ring.
          if (epi_tid == 0 && pending_flag != nullptr) {
            tma_store_wait_all<0>();
            __threadfence_system();
            st_release_sys(pending_flag, p.epoch);
            pending_flag = nullptr;
          }
          if (epi_tid < p.tp_size && epi_tid != p.tp_rank) {
            const int lm = lrow0 / BM;
            const uint32_t* flag = p.peer_flags[p.tp_rank] +
                                   (static_cast<size_t>(epi_tid) * (p.rows_per_rank / BM) + lm) * p.num_n_blocks + n_blk;
            while (static_cast<int32_t>(ld_acquire_sys(flag) - p.epoch) < 0) __nanosleep(32);
            __threadfence_system();
          }
          named_bar_sync(3, kEpiThreads);
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      float rscale = 1.f;
      if constexpr (kMode == kModeGroupM) {
        if (p.row_scale != nullptr) rscale = __ldg(p.row_scale + m0 + row_in_tile);
      }
      const uint16_t* bias_base = reinterpret_cast<const uint16_t*>(p.bias);
      if constexpr (kMode == kModeGroupM) {
        if (bias_base != nullptr) bias_base += static_cast<size_t>(t.group) * p.N;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr + c * 64, v0);
        tmem_ld_32x32(taddr + c * 64 + 32, v1);
        tmem_ld_wait();
        if (c == BN / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&tmem_empty_bar[acc]);
            else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
          }
        }
        const int gcol0 = n_blk * BN + c * 64;
        if constexpr (kEpi == kEpiStore16 || kEpi == kEpiReduceScatter) {
          float f[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            f[i] = __uint_as_float(v0[i]);
            f[32 + i] = __uint_as_float(v1[i]);
          }
          if constexpr (kMode == kModeRS) {
            if (own) {
              // fold the reduction in: partials of the tp-1 peers for this row, bf16 in this rank's staging
              const uint16_t* stage_base = reinterpret_cast<const uint16_t*>(p.peer_buf[p.tp_rank]);
              for (int s = 0; s < p.tp_size; ++s) {
                if (s == p.tp_rank) continue;
                const uint16_t* src = stage_base +
                                      (static_cast<size_t>(s) * p.rows_per_rank + lrow0 + row_in_tile) * p.N + gcol0;
                int4 raw[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (gcol0 + j * 8 < p.N) raw[j] = ld_cg_v4(reinterpret_cast<const int4*>(src + j * 8));
                  else raw[j] = make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint32_t* uu = reinterpret_cast<const uint32_t*>(&raw[j]);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    float2 f2 = p.is_bf16 ? unpack_bf16x2(uu[q]) : __half22float2(*reinterpret_cast<const __half2*>(&uu[q]));
                    f[j * 8 + 2 * q] += f2.x;
                    f[j * 8 + 2 * q + 1] += f2.y;
                  }
                }
              }
            }
          }
          if (bias_base != nullptr && own) {
            const uint32_t* b32 = reinterpret_cast<const uint32_t*>(bias_base + gcol0);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (gcol0 + 2 * i < p.N) {
                uint32_t bb = __ldg(b32 + i);
                float2 bf = p.is_bf16 ? unpack_bf16x2(bb) : __half22float2(*reinterpret_cast<__half2*>(&bb));
                f[2 * i] += bf.x;
                f[2 * i + 1] += bf.y;
              }
            }
          }
          const int n_out = (kMode != kModeRS && p.D_pre != nullptr) ? 2 : 1;
          for (int o = 0; o < n_out; ++o) {
            const bool pre = (n_out == 2 && o == 0);
            if (epi_tid == 0) tma_store_wait_read<kCBufs - 1>();
            named_bar_sync(1, kEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kC + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float a = f[j * 8 + 2 * q], b = f[j * 8 + 2 * q + 1];
                if (!pre && p.act == 1) { a = gelu2(a); b = gelu2(b); }
                if constexpr (kMode == kModeGroupM) { if (!pre) { a *= rscale; b *= rscale; } }
                if (p.is_bf16) pk[q] = pack_bf16x2(a, b);
                else {
                  __half2 h = __floats2half2_rn(a, b);
                  pk[q] = *reinterpret_cast<uint32_t*>(&h);
                }
              }
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, kEpiThreads);
            if (epi_tid == 0) {
              if constexpr (kMode == kModeRS) {
                if (own) tma_store_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0, lrow0);
                else tma_store_2d(&peer_maps.m[owner], smem_c + cbuf * L::kC, gcol0, p.tp_rank * p.rows_per_rank + lrow0);
              } else {
                tma_store_2d(pre ? &tmap_d_pre : &tmap_d, smem_c + cbuf * L::kC, gcol0, m0);
              }
              tma_store_commit();
            }
            if (++cbuf == kCBufs) cbuf = 0;
          }
        } else {
          const int drow0 = (kMode == kModeGroupK) ? t.group * p.group_stride + m0 : m0;
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            const uint32_t* v = hseg == 0 ? v0 : v1;
            if (epi_tid == 0) tma_store_wait_read<kCBufs - 1>();
            named_bar_sync(1, kEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kC + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) =
                  make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
            fence_proxy_async_smem();
            named_bar_sync(2, kEpiThreads);
            if (epi_tid == 0) {
              if constexpr (kEpi == kEpiAccumF32) tma_reduce_add_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0 + hseg * 32, drow0);
              else tma_store_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0 + hseg * 32, drow0);
              tma_store_commit();
            }
            if (++cbuf == kCBufs) cbuf = 0;
          }
        }
      }
      if constexpr (kMode == kModeRS) {
        // The flag of a pushed tile is published one tile later: bulk-store groups retire in order, so once at most
        // this tile's BN/64 groups are pending the previous tile's data has landed at its owner.
        if (epi_tid == 0) {
          if (pending_flag != nullptr) {
            tma_store_wait_all<BN / 64>();
            __threadfence_system();
            st_release_sys(pending_flag, p.epoch);
            pending_flag = nullptr;
          }
          if (!own) {
            // flags layout on the owner: [src rank][local 128-row block][n block]
            pending_flag = p.peer_flags[owner] +
                           (static_cast<size_t>(p.tp_rank) * (p.rows_per_rank / BM) + lrow0 / BM) * p.num_n_blocks + n_blk;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (epi_tid == 0) {
      tma_store_wait_all<0>();
      if constexpr (kMode == kModeRS) {
        if (pending_flag != nullptr) {
          __threadfence_system();
          st_release_sys(pending_flag, p.epoch);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // both CTAs are done with TMEM and with each other's barriers
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc2<2 * BN>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn2 encode2() {
  static EncodeTiledFn2 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(ptr);
  });
  return fn;
}

thread_local std::string g_err2;

bool tmap2(CUtensorMap* map, const void* base, int dtype /*0 bf16 1 f16 2 f32*/, uint64_t inner, uint64_t outer,
           uint64_t ld, uint32_t bi, uint32_t bo) {
  ensure_cuda_context();
  auto fn = encode2();
  if (!fn) { g_err2 = "cuTensorMapEncodeTiled unavailable"; return false; }
  const int es = dtype == 2 ? 4 : 2;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * es};
  cuuint32_t box[2] = {bi, bo};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_err2 = "tensor map encode failed: " + std::to_string(static_cast<int>(r)); return false; }
  return true;
}

template <bool AM, bool BMj, int EPI, int MODE>
int launch2(const GemmArgs& a, const P2& p, const CUtensorMap& ta, const CUtensorMap& tal, const CUtensorMap& tb,
            const CUtensorMap& td, const CUtensorMap& tdp, const PeerMaps2& pm, int grid) {
  auto kern = gemm2_kernel<AM, BMj, EPI, MODE>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2T<MODE>::kTotal) != cudaSuccess) {
      g_err2 = "cudaFuncSetAttribute failed"; return -10;
    }
    set = true;
  }
  kern<<<grid, kThreads, Smem2T<MODE>::kTotal, a.stream>>>(ta, tal, tb, td, tdp, pm, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err2 = std::string("launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}

}  // namespace

const char* gemm2_last_error() { return g_err2.c_str(); }

bool gemm2_supported(const GemmArgs& a) {
  if (a.group_mode != kGroupNone) return true;   // grouped GEMMs exist only on this kernel (launcher validates)
  if (a.comm == kCommAllGather) {
    return a.epilogue == kEpiStore16 && !a.a_mn_major && a.tp_size >= 2 && a.M % (a.tp_size * 256) == 0 && a.N >= 128;
  }
  if (a.comm == kCommAllGatherK) {
    return (a.epilogue == kEpiAccumF32 || a.epilogue == kEpiStoreF32) && a.a_mn_major && a.b_mn_major && a.tp_size >= 2 &&
           a.K % (a.tp_size * 128) == 0 && a.M >= 256 && a.N >= 128 && a.N % 8 == 0;
  }
  if (a.comm == kCommReduceScatter) {
    return a.epilogue == kEpiReduceScatter && !a.a_mn_major && a.tp_size >= 2 && a.M % (a.tp_size * 256) == 0 &&
           a.N >= 128 && a.N % 8 == 0;
  }
  if (a.epilogue != kEpiStore16 && a.epilogue != kEpiAccumF32 && a.epilogue != kEpiStoreF32) return false;
  return a.M >= 256 && a.N >= 128;
}

int gemm2_bf16(const GemmArgs& a) {
  P2 p{};
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldd = a.ldd; p.D = a.D; p.D_pre = a.D_pre; p.bias = a.bias; p.act = a.act;
  p.is_bf16 = a.is_bf16 ? 1 : 0;
  p.num_m_blocks = (a.M + 255) / 256;
  p.num_n_blocks = (a.N + BN - 1) / BN;
  p.num_k_blocks = (a.K + BK - 1) / BK;
  p.n_chunks = 1; p.m_blocks_per_chunk = p.num_m_blocks; p.chunk_base = 0; p.chunk_dir = 1;
  p.tp_rank = a.tp_rank; p.tp_size = a.tp_size; p.epoch = a.epoch; p.col_splits = 1;
  p.group_offsets = a.group_offsets; p.n_groups = a.n_groups; p.group_stride = a.group_stride; p.row_scale = a.row_scale;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (a.num_sms > 0) sms = a.num_sms;
  int clusters = sms / 2;
  int mode = kModePlain;
  if (a.group_mode == kGroupM) {
    mode = kModeGroupM;
    if (a.group_offsets == nullptr || a.n_groups < 1) { g_err2 = "grouped GEMM needs group_offsets"; return -4; }
    if (a.a_mn_major || a.epilogue != kEpiStore16) { g_err2 = "GroupM: K-major A, 16-bit store epilogue"; return -4; }
  } else if (a.group_mode == kGroupK) {
    mode = kModeGroupK;
    if (a.group_offsets == nullptr || a.n_groups < 1) { g_err2 = "grouped GEMM needs group_offsets"; return -4; }
    if (!a.a_mn_major || !a.b_mn_major || a.epilogue != kEpiAccumF32) {
      g_err2 = "GroupK: MN-major A and B, fp32 accumulate epilogue"; return -4;
    }
    if (a.M % 256 != 0) { g_err2 = "GroupK: M (output features) must be a multiple of 256"; return -4; }
  } else if (a.comm == kCommAllGatherK) {
    mode = kModeAGK;
    if (a.tp_size < 2 || a.tp_size > kMaxPeers) { g_err2 = "fused collective needs 2 <= tp <= 8"; return -4; }
    if (a.K % (a.tp_size * 128) != 0) { g_err2 = "AG-K fusion needs tokens % (tp*128) == 0"; return -4; }
    p.rows_per_rank = a.K / a.tp_size;
    for (int i = 0; i < a.tp_size; ++i) { p.peer_buf[i] = a.peer_buf[i]; p.peer_flags[i] = a.peer_flags[i]; }
    p.A_local = a.A_local; p.lda_local = a.lda_local; p.ag_cols = a.N;
    p.n_comm_ctas = a.n_comm_ctas > 0 ? (a.n_comm_ctas + 1) / 2 * 2 : 12;
    const int base_items = (p.rows_per_rank / BM) * a.tp_size;
    int cs = 1;
    const bool bulk_ok = (a.lda_local == a.N) && ((BM * a.N * 2) % kBulkChunk2 == 0);
    while (!bulk_ok && base_items * cs < p.n_comm_ctas && (a.N * 2) % (32 * cs) == 0 && cs < 8) cs *= 2;
    p.col_splits = cs;
    if (p.n_comm_ctas / 2 >= clusters) { g_err2 = "too many comm CTAs"; return -4; }
  } else if (a.comm != kCommNone) {
    if (a.tp_size < 2 || a.tp_size > kMaxPeers) { g_err2 = "fused collective needs 2 <= tp <= 8"; return -4; }
    if (a.M % (a.tp_size * 256) != 0) { g_err2 = "2-CTA fused collective needs M % (tp*256) == 0"; return -4; }
    p.rows_per_rank = a.M / a.tp_size;
    p.n_chunks = a.tp_size;
    p.m_blocks_per_chunk = p.rows_per_rank / 256;
    for (int i = 0; i < a.tp_size; ++i) { p.peer_buf[i] = a.peer_buf[i]; p.peer_flags[i] = a.peer_flags[i]; }
    if (a.comm == kCommAllGather) {
      mode = kModeAG;
      if ((a.K * 2) % 16 != 0) { g_err2 = "AG fusion needs K % 8 == 0"; return -4; }
      p.chunk_base = a.tp_rank; p.chunk_dir = -1;        // own shard first, then arrival order
      p.A_local = a.A_local; p.lda_local = a.lda_local; p.ag_cols = a.K;
      p.n_comm_ctas = a.n_comm_ctas > 0 ? (a.n_comm_ctas + 1) / 2 * 2 : 12;
      const int base_items = (p.rows_per_rank / BM) * a.tp_size;
      int cs = 1;
      const bool bulk_ok = (a.lda_local == a.K) && ((BM * a.K * 2) % kBulkChunk2 == 0);
      while (!bulk_ok && base_items * cs < p.n_comm_ctas && (a.K * 2) % (32 * cs) == 0 && cs < 8) cs *= 2;
      p.col_splits = cs;
      if (p.n_comm_ctas / 2 >= clusters) { g_err2 = "too many comm CTAs"; return -4; }
    } else {
      mode = kModeRS;
      if (a.epilogue != kEpiReduceScatter) { g_err2 = "RS fusion needs the RS epilogue"; return -4; }
      p.chunk_base = (a.tp_rank + 1) % a.tp_size; p.chunk_dir = 1;  // own rows last
    }
  }
  if (mode == kModePlain || mode == kModeGroupM) {
    // L2 banding: tiles walk n-groups (8 x 256 columns) inside a band of m-blocks.  Without bands the whole A operand is
    // re-streamed from HBM once per n-group as soon as it no longer fits L2; with bands B is re-read once per band.
    // Pick whichever moves fewer bytes (l2_eff: what one operand can hold on to while the other streams through).
    const double l2_eff = 48.0 * (1 << 20);
    const double a_blk = 256.0 * a.K * 2, a_total = a_blk * p.num_m_blocks;
    const double b_one = static_cast<double>(a.N) * a.K * 2;
    const int n_ngroups = (p.num_n_blocks + kGroupN2 - 1) / kGroupN2;
    const int groups = mode == kModeGroupM ? a.n_groups : 1;
    int band = static_cast<int>(l2_eff / a_blk);
    if (band < 2) band = 2;
    if (a_total > l2_eff && band < p.num_m_blocks) {
      const int n_bands = (p.num_m_blocks + band - 1) / band;
      const double cost_plain = groups * b_one + a_total * n_ngroups;
      const double cost_band = (groups - 1 + n_bands) * b_one + a_total;
      if (cost_band < cost_plain) { p.n_chunks = n_bands; p.m_blocks_per_chunk = band; }
    }
  }
  const long tiles = mode == kModeGroupK ? static_cast<long>(a.n_groups) * p.num_m_blocks * p.num_n_blocks
                                         : static_cast<long>(p.n_chunks) * p.m_blocks_per_chunk * p.num_n_blocks;
  const int comm_clusters = p.n_comm_ctas / 2;
  if (tiles < clusters - comm_clusters) clusters = static_cast<int>(tiles) + comm_clusters;
  const int dt = a.is_bf16 ? 0 : 1;
  CUtensorMap ta{}, tal{}, tb{}, td{}, tdp{};
  bool ok = true;
  if (!a.a_mn_major) ok = ok && tmap2(&ta, a.A, dt, a.K, a.M, a.lda, BK, BM);
  else ok = ok && tmap2(&ta, a.A, dt, a.M, a.K, a.lda, 64, BK);
  if (mode == kModeAG) ok = ok && tmap2(&tal, a.A_local, dt, a.K, p.rows_per_rank, a.lda_local, BK, BM);
  else if (mode == kModeAGK) ok = ok && tmap2(&tal, a.A_local, dt, a.N, p.rows_per_rank, a.lda_local, 64, BK);
  else tal = ta;
  const uint64_t b_groups = (mode == kModeGroupM) ? static_cast<uint64_t>(a.n_groups) : 1;
  if (!a.b_mn_major) ok = ok && tmap2(&tb, a.B, dt, a.K, static_cast<uint64_t>(a.N) * b_groups, a.ldb, BK, BN / 2);
  else ok = ok && tmap2(&tb, a.B, dt, a.N, static_cast<uint64_t>(a.K) * b_groups, a.ldb, 64, BK);
  if (a.epilogue == kEpiStore16) {
    ok = ok && tmap2(&td, a.D, dt, a.N, a.M, a.ldd, 64, BM);
    if (a.D_pre) ok = ok && tmap2(&tdp, a.D_pre, dt, a.N, a.M, a.ldd, 64, BM);
    else tdp = td;
  } else if (a.epilogue == kEpiReduceScatter) {
    ok = ok && tmap2(&td, a.D, dt, a.N, p.rows_per_rank, a.ldd, 64, BM);
    tdp = td;
  } else {
    const uint64_t d_rows = (mode == kModeGroupK) ? static_cast<uint64_t>(a.n_groups) * a.group_stride : a.M;
    ok = ok && tmap2(&td, a.D, 2, a.N, d_rows, a.ldd, 32, BM);
    tdp = td;
  }
  if (!ok) return -5;
  PeerMaps2 pm{};
  for (int s = 0; s < kMaxPeers; ++s) pm.m[s] = ta;
  if (mode == kModeRS) {
    for (int s = 0; s < a.tp_size; ++s) {
      if (!tmap2(&pm.m[s], a.peer_buf[s], dt, a.N, static_cast<uint64_t>(a.tp_size) * p.rows_per_rank, a.N, 64, BM)) return -5;
    }
  }
  const int grid = clusters * 2;
#define G2_CASE(AM, BMJ, EPI, MODE) \
  if (a.a_mn_major == AM && a.b_mn_major == BMJ && a.epilogue == EPI && mode == MODE) \
    return launch2<AM, BMJ, EPI, MODE>(a, p, ta, tal, tb, td, tdp, pm, grid);
  G2_CASE(false, false, kEpiStore16, kModePlain)
  G2_CASE(false, true, kEpiStore16, kModePlain)
  G2_CASE(true, true, kEpiStore16, kModePlain)
  G2_CASE(true, false, kEpiStore16, kModePlain)
  G2_CASE(true, true, kEpiAccumF32, kModePlain)
  G2_CASE(true, true, kEpiStoreF32, kModePlain)
  G2_CASE(false, false, kEpiAccumF32, kModePlain)
  G2_CASE(false, false, kEpiStoreF32, kModePlain)
  // fused collectives
  G2_CASE(false, false, kEpiStore16, kModeAG)
  G2_CASE(false, true, kEpiStore16, kModeAG)
  G2_CASE(false, false, kEpiReduceScatter, kModeRS)
  G2_CASE(false, true, kEpiReduceScatter, kModeRS)
  G2_CASE(true, true, kEpiAccumF32, kModeAGK)
  G2_CASE(true, true, kEpiStoreF32, kModeAGK)
  // grouped (MoE)
  G2_CASE(false, false, kEpiStore16, kModeGroupM)
  G2_CASE(false, true, kEpiStore16, kModeGroupM)
  G2_CASE(true, true, kEpiAccumF32, kModeGroupK)
#undef G2_CASE
  g_err2 = "unsupported combination for the 2-CTA kernel";
  return -3;
}

}  // namespace b200
