// Persistent warp-specialised tcgen05 GEMM for sm_100a, with optional fused collectives.
//
//   D[M,N] = A[M,K] * B[N,K]^T      bf16/fp16 inputs, fp32 accumulation in TMEM
//
// Roles (256 threads, 1 CTA / SM):  warp 0 = TMA producer, warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> registers -> smem -> TMA store,
// or direct global/peer stores).  Three pipelines: smem full/empty ring (TMA <-> MMA),
// TMEM full/empty double buffer (MMA <-> epilogue), static persistent tile scheduler.
//
// Operand storage: K-major (row-major [rows,K]) or MN-major (stored transposed) for both A and
// B, so forward (X W^T), dgrad (dY W) and wgrad (dY^T X, fp32 accumulate into main_grad) all run
// without materialising a transpose.
//
// Fused collectives (north-star kernels, reference call sites layers.py:429-443 / :1200-1217):
//   kCommAllGather      extra "comm" CTAs push this rank's activation shard into every peer's
//                       gather buffer over NVLink (st.global on IPC-mapped peer memory) and bump
//                       a per-row-block flag (red.release.sys); GEMM producers wait on the flag
//                       of the row block they are about to load (ld.acquire.sys) - tiles are
//                       visited own-shard first, then in arrival order.
//   kCommReduceScatter  epilogue stores each partial tile straight into the owner rank's staging
//                       buffer, signals a per-tile flag; after its GEMM tiles every CTA turns
//                       reducer for the rows this rank owns (fixed src order => deterministic).
#include "b200_api.h"
#include "common.cuh"

#include <mutex>
#include <string>

namespace b200 {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 x 2B = one 128B swizzle atom
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kNumEpiThreads = 128;
constexpr int kGroupN = 8;    // n-blocks per scheduling group (L2 reuse of A)
constexpr int kBulkChunk = 32768;  // comm-CTA TMA bulk copy granule
constexpr int kBulkStages = 6;      // 6 x 32 KB in flight per comm CTA

struct DevParams {
  int M, N, K;
  int ldd;
  void* D;
  void* D_pre;
  const void* bias;
  int act;
  int is_bf16;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  // chunked tile order: chunks are visited as chunk(i) = (chunk_base + chunk_dir * i) mod n_chunks
  int n_chunks, m_blocks_per_chunk, chunk_base, chunk_dir;
  // comm
  int tp_rank, tp_size, n_comm_ctas;
  uint32_t epoch;
  int rows_per_rank;
  int col_splits;
  const void* A_local;
  int lda_local;
  void* peer_buf[kMaxPeers];
  uint32_t* peer_flags[kMaxPeers];
};

// reduce-scatter epilogue: one tensor map per owner rank over that rank's staging buffer
// [src rank][rows_per_rank][N], so partial tiles leave through asynchronous TMA stores to peer memory
struct PeerMaps {
  CUtensorMap m[kMaxPeers];
};

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.79788456f, k1 = 0.044715f;
  float inner = k0 * x * (1.f + k1 * x * x);
  return 0.5f * x * (1.f + tanhf(inner));
}

template <int BLOCK_N, int kStages>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kCBytes = BLOCK_M * 64 * 2;  // one 64-column chunk of the output tile
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAOff + kStages * kABytes;
  static constexpr int kCOff = kBOff + kStages * kBBytes;
  static constexpr int kBarOff = kCOff + 2 * kCBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;  // + alignment slack
};

// tile index -> (m_blk, n_blk) honouring the chunk order and n-grouping
__device__ __forceinline__ void tile_coords(const DevParams& p, int tile, int& m_blk, int& n_blk) {
  const int tiles_per_chunk = p.m_blocks_per_chunk * p.num_n_blocks;
  const int ci = tile / tiles_per_chunk;
  int r = tile - ci * tiles_per_chunk;
  int chunk = (p.chunk_base + p.chunk_dir * ci) % p.n_chunks;
  if (chunk < 0) chunk += p.n_chunks;
  const int group_tiles = p.m_blocks_per_chunk * kGroupN;
  const int g = r / group_tiles;
  r -= g * group_tiles;
  const int n_in_group = min(kGroupN, p.num_n_blocks - g * kGroupN);
  const int m_local = r / n_in_group;
  n_blk = g * kGroupN + (r - m_local * n_in_group);
  m_blk = chunk * p.m_blocks_per_chunk + m_local;
}

template <bool kAMn, bool kBMn, int BLOCK_N, int kStages, int kEpi, int kComm>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_local,
            const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_d,
            const __grid_constant__ CUtensorMap tmap_d_pre, const __grid_constant__ PeerMaps peer_maps,
            const DevParams p) {
  using L = SmemLayout<BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
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

  // ------------------------------------------------------------------ comm CTAs (all-gather push)
  if constexpr (kComm == kCommAllGather) {
    if (static_cast<int>(blockIdx.x) < p.n_comm_ctas) {
      // work item = (row block of 128 rows, destination, column split); all remote destinations
      // first (they gate the peers' GEMM tiles), the local copy (d == 0: complete gathered view for
      // the caller's wgrad) last.
      const int rb_per_rank = p.rows_per_rank / BLOCK_M;
      const int per_dst = rb_per_rank * p.col_splits;
      const int n_items = per_dst * p.tp_size;
      const int row_bytes = p.K * 2;
      const int split_bytes = row_bytes / p.col_splits;  // multiple of 16 (checked on host)
      const bool contiguous = (p.lda_local == p.K) && (p.col_splits == 1) && ((BLOCK_M * row_bytes) % kBulkChunk == 0);
      if (contiguous) {
        // TMA bulk pipeline: thread 0 streams gmem -> smem stages, thread 32 streams smem -> peer gmem.
        uint64_t* c_full = reinterpret_cast<uint64_t*>(smem + kBulkStages * kBulkChunk);
        uint64_t* c_empty = c_full + kBulkStages;
        if (threadIdx.x == 0) {
          for (int i = 0; i < kBulkStages; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 1); }
          mbar_fence_init();
        }
        __syncthreads();
        const int chunks_per_item = BLOCK_M * row_bytes / kBulkChunk;
        if (threadIdx.x == 0) {
          int stage = 0; uint32_t phase = 0;
          for (int item = blockIdx.x; item < n_items; item += p.n_comm_ctas) {
            const int rb = (item % per_dst);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.A_local) + static_cast<size_t>(rb) * BLOCK_M * row_bytes;
            for (int c = 0; c < chunks_per_item; ++c) {
              mbar_wait(&c_empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&c_full[stage], kBulkChunk);
              bulk_load_1d(smem + stage * kBulkChunk, src + static_cast<size_t>(c) * kBulkChunk, kBulkChunk, &c_full[stage]);
              if (++stage == kBulkStages) { stage = 0; phase ^= 1; }
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
                           (static_cast<size_t>(p.tp_rank) * p.rows_per_rank + rb * BLOCK_M) * row_bytes;
            for (int c = 0; c < chunks_per_item; ++c) {
              mbar_wait(&c_full[stage], phase);
              bulk_store_1d(dst + static_cast<size_t>(c) * kBulkChunk, smem + stage * kBulkChunk, kBulkChunk);
              tma_store_commit();
              if (prev_stage >= 0) {          // the previous store has finished *reading* its stage
                tma_store_wait_read<1>();
                mbar_arrive(&c_empty[prev_stage]);
              }
              prev_stage = stage;
              if (++stage == kBulkStages) { stage = 0; phase ^= 1; }
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
                             static_cast<size_t>(rb) * BLOCK_M * p.lda_local * 2 + cs * split_bytes;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.peer_buf[peer]) +
                       (static_cast<size_t>(p.tp_rank) * p.rows_per_rank + rb * BLOCK_M) * row_bytes +
                       cs * split_bytes;
        const int total_vec = BLOCK_M * vec_per_row;
        for (int v = threadIdx.x; v < total_vec; v += 8 * kNumThreads) {
          int4 regs[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int vv = v + u * kNumThreads;
            if (vv < total_vec) {
              const int r = vv / vec_per_row, c = vv - r * vec_per_row;
              regs[u] = ld_nc_v4(reinterpret_cast<const int4*>(
                  src + static_cast<size_t>(r) * p.lda_local * 2 + c * 16));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int vv = v + u * kNumThreads;
            if (vv < total_vec) {
              const int r = vv / vec_per_row, c = vv - r * vec_per_row;
              st_na_v4(reinterpret_cast<int4*>(dst + static_cast<size_t>(r) * row_bytes + c * 16),
                       regs[u]);
            }
          }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && d != 0) {
          // single-writer flag carrying the call epoch (monotonic per arena slot)
          st_release_sys(p.peer_flags[peer] +
                             (static_cast<size_t>(p.tp_rank) * rb_per_rank + rb) * p.col_splits + cs,
                         p.epoch);
        }
      }
      return;
    }
  }

  const int cta = (kComm == kCommAllGather) ? static_cast<int>(blockIdx.x) - p.n_comm_ctas
                                            : static_cast<int>(blockIdx.x);
  const int n_ctas = (kComm == kCommAllGather) ? static_cast<int>(gridDim.x) - p.n_comm_ctas
                                               : static_cast<int>(gridDim.x);
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  // ------------------------------------------------------------------ one-time setup
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (kEpi != kEpiReduceScatter) tma_prefetch_desc(&tmap_d);
    if constexpr (kComm == kCommAllGather) tma_prefetch_desc(&tmap_a_local);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<2 * BLOCK_N>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  constexpr uint32_t kStageBytes = L::kABytes + L::kBBytes;

  if (warp_idx == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta; tile < num_tiles; tile += n_ctas) {
        int m_blk, n_blk;
        tile_coords(p, tile, m_blk, n_blk);
        bool use_local = false;
        int m_coord = m_blk * BLOCK_M;
        if constexpr (kComm == kCommAllGather) {
          const int src = (m_blk * BLOCK_M) / p.rows_per_rank;
          if (src == p.tp_rank) {
            use_local = true;
            m_coord = m_blk * BLOCK_M - src * p.rows_per_rank;
          } else {
            const int rb_per_rank = p.rows_per_rank / BLOCK_M;
            const int rb = (m_blk * BLOCK_M - src * p.rows_per_rank) / BLOCK_M;
            const uint32_t* flag = p.peer_flags[p.tp_rank] +
                                   (static_cast<size_t>(src) * rb_per_rank + rb) * p.col_splits;
            for (int cs = 0; cs < p.col_splits; ++cs) {
              while (static_cast<int32_t>(ld_acquire_sys(flag + cs) - p.epoch) < 0) __nanosleep(64);
            }
            fence_proxy_async_all();
          }
        }
        const CUtensorMap* ta = use_local ? &tmap_a_local : &tmap_a;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          uint8_t* sa = smem_a + stage * L::kABytes;
          uint8_t* sb = smem_b + stage * L::kBBytes;
          if constexpr (!kAMn) {
            tma_load_2d(ta, &full_bar[stage], sa, kb * BLOCK_K, m_coord);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(ta, &full_bar[stage], sa + j * (BLOCK_K * 128), m_coord + j * 64,
                          kb * BLOCK_K);
          }
          if constexpr (!kBMn) {
            tma_load_2d(&tmap_b, &full_bar[stage], sb, kb * BLOCK_K, n_blk * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(&tmap_b, &full_bar[stage], sb + j * (BLOCK_K * 128),
                          n_blk * BLOCK_N + j * 64, kb * BLOCK_K);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cta; tile < num_tiles; tile += n_ctas) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, BLOCK_M, BLOCK_N, kAMn, kBMn);
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * L::kABytes);
          const uint32_t sb = smem_u32(smem_b + stage * L::kBBytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: advance 16 elements (32 B) inside the 128B swizzle atom; SBO = 8 rows.
            // MN-major: advance 16 k-rows (16*128 B); SBO = 8 k-rows, LBO = next 64-wide MN atom.
            const uint64_t da = kAMn ? make_smem_desc_sw128(sa + k * UMMA_K * 128, 1024, BLOCK_K * 128)
                                     : make_smem_desc_sw128(sa + k * UMMA_K * 2, 1024, 0);
            const uint64_t db = kBMn ? make_smem_desc_sw128(sb + k * UMMA_K * 128, 1024, BLOCK_K * 128)
                                     : make_smem_desc_sw128(sb + k * UMMA_K * 2, 1024, 0);
            umma_f16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp_idx >= 4) {
    // ================================================================ epilogue
    const int ew = warp_idx - 4;             // == warp_idx % 4 -> TMEM lane quadrant
    const int epi_tid = threadIdx.x - 128;
    const int row_in_tile = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    uint32_t* pending_flag = nullptr;   // reduce-scatter: flag of the previous tile, not yet published
    for (int tile = cta; tile < num_tiles; tile += n_ctas) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;
      const int grow = m_blk * BLOCK_M + row_in_tile;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 64; ++c) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr + c * 64, v0);
        tmem_ld_32x32(taddr + c * 64 + 32, v1);
        tmem_ld_wait();
        if (c == BLOCK_N / 64 - 1) {
          // accumulator fully drained into registers: hand TMEM back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int gcol0 = n_blk * BLOCK_N + c * 64;
        if constexpr (kEpi == kEpiStore16) {
          float f[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            f[i] = __uint_as_float(v0[i]);
            f[32 + i] = __uint_as_float(v1[i]);
          }
          if (p.bias != nullptr) {
            const uint32_t* b32 = reinterpret_cast<const uint32_t*>(
                reinterpret_cast<const uint16_t*>(p.bias) + gcol0);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (gcol0 + 2 * i < p.N) {
                uint32_t bb = __ldg(b32 + i);
                float2 bf;
                if (p.is_bf16) bf = unpack_bf16x2(bb);
                else bf = __half22float2(*reinterpret_cast<__half2*>(&bb));
                f[2 * i] += bf.x;
                f[2 * i + 1] += bf.y;
              }
            }
          }
          const int n_out = (p.D_pre != nullptr) ? 2 : 1;
          for (int o = 0; o < n_out; ++o) {
            const bool pre = (n_out == 2 && o == 0);
            // staging buffer cbuf was last read by the TMA store issued two chunks ago
            if (epi_tid == 0) tma_store_wait_read<1>();
            named_bar_sync(1, kNumEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kCBytes + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float a = f[j * 8 + 2 * q], b = f[j * 8 + 2 * q + 1];
                if (!pre && p.act == 1) { a = gelu_tanh(a); b = gelu_tanh(b); }
                if (p.is_bf16) pk[q] = pack_bf16x2(a, b);
                else {
                  __half2 h = __floats2half2_rn(a, b);
                  pk[q] = *reinterpret_cast<uint32_t*>(&h);
                }
              }
              // 128B swizzle: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, kNumEpiThreads);
            if (epi_tid == 0) {
              tma_store_2d(pre ? &tmap_d_pre : &tmap_d, smem_c + cbuf * L::kCBytes, gcol0,
                           m_blk * BLOCK_M);
              tma_store_commit();
            }
            cbuf ^= 1;
          }
        } else if constexpr (kEpi == kEpiAccumF32 || kEpi == kEpiStoreF32) {
          // fp32 tile -> smem (two 32-column halves, 128B-swizzled) -> TMA store / TMA reduce-add.
          // The "+=" happens in L2 (cp.reduce.async.bulk.tensor .add), so no main_grad read
          // round trip stalls the epilogue warps.
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            const uint32_t* v = hseg == 0 ? v0 : v1;
            if (epi_tid == 0) tma_store_wait_read<1>();
            named_bar_sync(1, kNumEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kCBytes + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) =
                  make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, kNumEpiThreads);
            if (epi_tid == 0) {
              if constexpr (kEpi == kEpiAccumF32)
                tma_reduce_add_2d(&tmap_d, smem_c + cbuf * L::kCBytes, gcol0 + hseg * 32, m_blk * BLOCK_M);
              else
                tma_store_2d(&tmap_d, smem_c + cbuf * L::kCBytes, gcol0 + hseg * 32, m_blk * BLOCK_M);
              tma_store_commit();
            }
            cbuf ^= 1;
          }
        } else {  // kEpiReduceScatter: partial tile -> owner's staging buffer [src][rows_per_rank][N]
          // Same swizzled-smem staging as the plain store epilogue, but the TMA store targets the
          // owner's arena over NVLink: the epilogue warps never wait on a remote write.
          const int owner = (m_blk * BLOCK_M) / p.rows_per_rank;
          if (epi_tid == 0) tma_store_wait_read<1>();
          named_bar_sync(1, kNumEpiThreads);
          uint8_t* sc = smem_c + cbuf * L::kCBytes + row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = j * 8 + 2 * q;
              float a = __uint_as_float(i < 32 ? v0[i] : v1[i - 32]);
              float b = __uint_as_float(i + 1 < 32 ? v0[i + 1] : v1[i + 1 - 32]);
              if (p.is_bf16) pk[q] = pack_bf16x2(a, b);
              else {
                __half2 h = __floats2half2_rn(a, b);
                pk[q] = *reinterpret_cast<uint32_t*>(&h);
              }
            }
            *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(2, kNumEpiThreads);
          if (epi_tid == 0) {
            const int row0 = p.tp_rank * p.rows_per_rank + (m_blk * BLOCK_M - owner * p.rows_per_rank);
            tma_store_2d(&peer_maps.m[owner], smem_c + cbuf * L::kCBytes, gcol0, row0);
            tma_store_commit();
          }
          cbuf ^= 1;
        }
      }
      if constexpr (kEpi == kEpiReduceScatter) {
        // The flag of a tile is published one tile later: bulk-store groups retire in order, so once at
        // most this tile's BLOCK_N/64 groups are pending the previous tile's data has landed at its owner.
        if (epi_tid == 0) {
          if (pending_flag != nullptr) {
            tma_store_wait_all<BLOCK_N / 64>();
            __threadfence_system();
            st_release_sys(pending_flag, p.epoch);
          }
          const int owner = (m_blk * BLOCK_M) / p.rows_per_rank;
          const int lm = m_blk - owner * p.m_blocks_per_chunk;
          // flags layout on the owner: [src rank][local m block][n block]
          pending_flag = p.peer_flags[owner] +
                         (static_cast<size_t>(p.tp_rank) * p.m_blocks_per_chunk + lm) * p.num_n_blocks + n_blk;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (kEpi == kEpiStore16 || kEpi == kEpiAccumF32 || kEpi == kEpiStoreF32) {
      if (epi_tid == 0) tma_store_wait_all<0>();
    }
    if constexpr (kEpi == kEpiReduceScatter) {
      if (epi_tid == 0) {
        tma_store_wait_all<0>();
        if (pending_flag != nullptr) {
          __threadfence_system();
          st_release_sys(pending_flag, p.epoch);
        }
      }
      // ============================================================== reduce phase (owned rows)
      const int own_tiles = p.m_blocks_per_chunk * p.num_n_blocks;
      const uint32_t* my_flags = p.peer_flags[p.tp_rank];
      const uint16_t* stage = reinterpret_cast<const uint16_t*>(p.peer_buf[p.tp_rank]);
      for (int t = cta; t < own_tiles; t += n_ctas) {
        const int lm = t / p.num_n_blocks, nb = t - lm * p.num_n_blocks;
        if (epi_tid < p.tp_size) {
          const uint32_t* flag = my_flags + (static_cast<size_t>(epi_tid) * p.m_blocks_per_chunk + lm) * p.num_n_blocks + nb;
          while (static_cast<int32_t>(ld_acquire_sys(flag) - p.epoch) < 0) __nanosleep(64);
        }
        __threadfence_system();
        named_bar_sync(1, kNumEpiThreads);
        const int vec_per_row = BLOCK_N / 8;
        constexpr int kIlp = 4;
        for (int v0i = epi_tid; v0i < BLOCK_M * vec_per_row; v0i += kNumEpiThreads * kIlp) {
          int4 raw[kIlp][kMaxPeers];
          bool valid[kIlp];
          int lrows[kIlp], cols[kIlp];
#pragma unroll
          for (int u = 0; u < kIlp; ++u) {
            const int v = v0i + u * kNumEpiThreads;
            const int r = v / vec_per_row, cv = v - r * vec_per_row;
            lrows[u] = lm * BLOCK_M + r;
            cols[u] = nb * BLOCK_N + cv * 8;
            valid[u] = (v < BLOCK_M * vec_per_row) && lrows[u] < p.rows_per_rank && cols[u] < p.N;
            if (valid[u]) {
              for (int s = 0; s < p.tp_size; ++s)
                raw[u][s] = ld_cg_v4(reinterpret_cast<const int4*>(
                    stage + (static_cast<size_t>(s) * p.rows_per_rank + lrows[u]) * p.N + cols[u]));
            }
          }
#pragma unroll
          for (int u = 0; u < kIlp; ++u) {
            if (!valid[u]) continue;
            float accv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < p.tp_size; ++s) {
              const uint32_t* uu = reinterpret_cast<const uint32_t*>(&raw[u][s]);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float2 f2 = p.is_bf16 ? unpack_bf16x2(uu[q])
                                      : __half22float2(*reinterpret_cast<const __half2*>(&uu[q]));
                accv[2 * q] += f2.x;
                accv[2 * q + 1] += f2.y;
              }
            }
            if (p.bias != nullptr) {
              const uint16_t* bp = reinterpret_cast<const uint16_t*>(p.bias) + cols[u];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                accv[q] += p.is_bf16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(bp + q))
                                     : __half2float(*reinterpret_cast<const __half*>(bp + q));
              }
            }
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (p.is_bf16) pk[q] = pack_bf16x2(accv[2 * q], accv[2 * q + 1]);
              else {
                __half2 h = __floats2half2_rn(accv[2 * q], accv[2 * q + 1]);
                pk[q] = *reinterpret_cast<uint32_t*>(&h);
              }
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.D) +
                                      static_cast<size_t>(lrows[u]) * p.ldd + cols[u]) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        named_bar_sync(1, kNumEpiThreads);
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<2 * BLOCK_N>(tmem_base);
  }
}

// ---------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

thread_local std::string g_last_error;

bool make_tmap_2d(CUtensorMap* map, const void* base, bool is_bf16, uint64_t inner, uint64_t outer,
                  uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer, bool is_f32 = false) {
  ensure_cuda_context();
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { g_last_error = "cuTensorMapEncodeTiled unavailable"; return false; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * (is_f32 ? 4 : 2)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_error = "cuTensorMapEncodeTiled failed: code " + std::to_string(static_cast<int>(r)) +
                   " (inner=" + std::to_string(inner) + " outer=" + std::to_string(outer) +
                   " ld=" + std::to_string(ld_elems) + ")";
    return false;
  }
  return true;
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

template <bool kAMn, bool kBMn, int BLOCK_N, int kStages, int kEpi, int kComm>
int launch(const GemmArgs& a, const DevParams& p, const CUtensorMap& ta, const CUtensorMap& tal,
           const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tdp, const PeerMaps& pm, int grid) {
  using L = SmemLayout<BLOCK_N, kStages>;
  auto kern = gemm_kernel<kAMn, kBMn, BLOCK_N, kStages, kEpi, kComm>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) { g_last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -10; }
    attr_set = true;
  }
  kern<<<grid, kNumThreads, L::kTotal, a.stream>>>(ta, tal, tb, td, tdp, pm, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_error = std::string("launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}

template <int BLOCK_N, int kStages>
int dispatch(const GemmArgs& a, const DevParams& p, const CUtensorMap& ta, const CUtensorMap& tal,
             const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tdp, const PeerMaps& pm, int grid) {
#define B200_GEMM_CASE(AM, BM, EPI, COMM)                                                   \
  if (a.a_mn_major == AM && a.b_mn_major == BM && a.epilogue == EPI && a.comm == COMM)      \
    return launch<AM, BM, BLOCK_N, kStages, EPI, COMM>(a, p, ta, tal, tb, td, tdp, pm, grid);
  // forward / generic
  B200_GEMM_CASE(false, false, kEpiStore16, kCommNone)
  B200_GEMM_CASE(false, true, kEpiStore16, kCommNone)
  B200_GEMM_CASE(true, true, kEpiStore16, kCommNone)
  B200_GEMM_CASE(true, false, kEpiStore16, kCommNone)
  // wgrad
  B200_GEMM_CASE(true, true, kEpiAccumF32, kCommNone)
  B200_GEMM_CASE(true, true, kEpiStoreF32, kCommNone)
  B200_GEMM_CASE(false, false, kEpiStoreF32, kCommNone)
  B200_GEMM_CASE(false, false, kEpiAccumF32, kCommNone)
  // fused collectives
  B200_GEMM_CASE(false, false, kEpiStore16, kCommAllGather)
  B200_GEMM_CASE(false, true, kEpiStore16, kCommAllGather)
  B200_GEMM_CASE(false, false, kEpiReduceScatter, kCommReduceScatter)
  B200_GEMM_CASE(false, true, kEpiReduceScatter, kCommReduceScatter)
#undef B200_GEMM_CASE
  g_last_error = "unsupported (a_major, b_major, epilogue, comm) combination";
  return -3;
}

}  // namespace

const char* gemm_last_error() { return g_last_error.c_str(); }

int gemm_flags_needed(int M, int N, int tp_size) {
  // worst case of the two modes: AG needs tp*rows/128 flags, RS needs tp*(rows/tp/128)*(N/128)
  (void)tp_size;
  const int mb = (M + BLOCK_M - 1) / BLOCK_M;
  const int nb = (N + 127) / 128;
  return mb * (nb > 8 ? nb : 8) + 64;
}

static int prefer_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_GEMM_2CTA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;   // default on; B200_GEMM_2CTA=0 forces the 1-CTA kernel
  }
  return v;
}

int gemm_bf16(const GemmArgs& a) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) { g_last_error = "empty problem"; return -1; }
  // block_n: 2 = force the cta_group::2 kernel, 128 / 256 = force the 1-CTA kernel, 0 = auto
  if (a.group_mode != kGroupNone && a.block_n != 2 && a.block_n != 0) { g_last_error = "grouped GEMM runs on the 2-CTA kernel only"; return -3; }
  if (a.comm == kCommAllGatherK && !gemm2_supported(a)) { g_last_error = "AG-K wgrad: unsupported shape for the 2-CTA kernel"; return -3; }
  if ((a.group_mode != kGroupNone || a.comm == kCommAllGatherK || a.block_n == 2 || (a.block_n == 0 && prefer_2cta())) && gemm2_supported(a)) {
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) ||
        (reinterpret_cast<uintptr_t>(a.D) & 15) || (a.lda % 8) || (a.ldb % 8) ||
        ((a.epilogue == kEpiStore16 || a.epilogue == kEpiReduceScatter) ? (a.ldd % 8) : ((a.ldd % 4) || (a.N % 4)))) {
      g_last_error = "alignment: pointers 16B, lda/ldb % 8, ldd % 8 (16-bit) or % 4 (fp32)";
      return -2;
    }
    int rc = gemm2_bf16(a);
    if (rc != 0) g_last_error = gemm2_last_error();
    return rc;
  }
  if (a.block_n == 2) { g_last_error = "2-CTA kernel does not support this problem"; return -3; }
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) ||
      (reinterpret_cast<uintptr_t>(a.D) & 15)) { g_last_error = "pointers must be 16B aligned"; return -2; }
  if ((a.lda % 8) || (a.ldb % 8)) { g_last_error = "lda/ldb must be multiples of 8 elements"; return -2; }
  if (a.epilogue == kEpiStore16 && (a.ldd % 8)) { g_last_error = "ldd must be a multiple of 8"; return -2; }
  if ((a.epilogue == kEpiAccumF32 || a.epilogue == kEpiStoreF32) && ((a.ldd % 4) || (a.N % 4))) {
    g_last_error = "fp32 epilogue needs N, ldd multiples of 4"; return -2;
  }

  int block_n = a.block_n;
  const int sms = a.num_sms > 0 ? a.num_sms : num_sms();
  const int mb = (a.M + BLOCK_M - 1) / BLOCK_M;
  if (block_n == 0) {
    // pick the tile width with the better wave quantisation (ties -> 256: fewer B re-reads)
    auto eff = [&](int bn) {
      const long tiles = static_cast<long>(mb) * ((a.N + bn - 1) / bn);
      const long waves = (tiles + sms - 1) / sms;
      const double work = static_cast<double>(a.N) / bn * mb;  // in tile units of useful work
      return work / (waves * sms) * (bn == 256 ? 1.0 : 0.66);  // 128-wide tiles are smem-bandwidth bound (measured 0.65x)
    };
    block_n = (eff(256) >= eff(128)) ? 256 : 128;
    if (a.N <= 128) block_n = 128;
  }
  if (block_n != 128 && block_n != 256) { g_last_error = "block_n must be 128 or 256"; return -2; }

  DevParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K; p.ldd = a.ldd; p.D = a.D; p.D_pre = a.D_pre; p.bias = a.bias;
  p.act = a.act; p.is_bf16 = a.is_bf16 ? 1 : 0;
  p.num_m_blocks = mb;
  p.num_n_blocks = (a.N + block_n - 1) / block_n;
  p.num_k_blocks = (a.K + BLOCK_K - 1) / BLOCK_K;
  p.n_chunks = 1; p.m_blocks_per_chunk = mb; p.chunk_base = 0; p.chunk_dir = 1;
  p.tp_rank = a.tp_rank; p.tp_size = a.tp_size; p.n_comm_ctas = 0; p.epoch = a.epoch;
  p.col_splits = 1;

  int grid = sms;
  if (a.comm != kCommNone) {
    if (a.tp_size < 2 || a.tp_size > kMaxPeers) { g_last_error = "fused collective needs 2 <= tp <= 8"; return -4; }
    if (a.M % (a.tp_size * BLOCK_M) != 0) { g_last_error = "fused collective needs M % (tp*128) == 0"; return -4; }
    p.rows_per_rank = a.M / a.tp_size;
    p.n_chunks = a.tp_size;
    p.m_blocks_per_chunk = p.rows_per_rank / BLOCK_M;
    for (int i = 0; i < a.tp_size; ++i) { p.peer_buf[i] = a.peer_buf[i]; p.peer_flags[i] = a.peer_flags[i]; }
    if (a.comm == kCommAllGather) {
      if (a.a_mn_major) { g_last_error = "AG fusion needs K-major A"; return -4; }
      if ((a.K * 2) % 16 != 0) { g_last_error = "AG fusion needs K % 8 == 0"; return -4; }
      p.chunk_base = a.tp_rank; p.chunk_dir = -1;        // own shard first, then arrival order
      p.A_local = a.A_local; p.lda_local = a.lda_local;
      p.n_comm_ctas = a.n_comm_ctas > 0 ? a.n_comm_ctas : 12;
      const int base_items = p.m_blocks_per_chunk * a.tp_size;
      int cs = 1;
      const bool bulk_ok = (a.lda_local == a.K) && ((BLOCK_M * a.K * 2) % kBulkChunk == 0);
      while (!bulk_ok && base_items * cs < p.n_comm_ctas && (a.K * 2) % (32 * cs) == 0 && cs < 8) cs *= 2;
      p.col_splits = cs;
      grid = sms;  // comm CTAs + GEMM CTAs together fill the machine
      if (p.n_comm_ctas >= grid) { g_last_error = "too many comm CTAs"; return -4; }
    } else {
      if (a.epilogue != kEpiReduceScatter) { g_last_error = "RS fusion needs the RS epilogue"; return -4; }
      if (a.N % 8 != 0) { g_last_error = "RS fusion needs N % 8 == 0"; return -4; }
      p.chunk_base = (a.tp_rank + 1) % a.tp_size; p.chunk_dir = 1;  // own rows last
    }
  }
  const long num_tiles = static_cast<long>(p.num_m_blocks) * p.num_n_blocks;
  const int gemm_ctas_avail = grid - p.n_comm_ctas;
  if (a.comm != kCommAllGather && num_tiles < grid) grid = static_cast<int>(num_tiles);
  if (a.comm == kCommAllGather && num_tiles < gemm_ctas_avail) grid = p.n_comm_ctas + static_cast<int>(num_tiles);
  if (a.comm == kCommReduceScatter) grid = sms;  // reducers need every CTA to reach the reduce phase

  CUtensorMap ta{}, tal{}, tb{}, td{}, tdp{};
  bool ok = true;
  if (!a.a_mn_major) ok = ok && make_tmap_2d(&ta, a.A, a.is_bf16, a.K, a.M, a.lda, BLOCK_K, BLOCK_M);
  else ok = ok && make_tmap_2d(&ta, a.A, a.is_bf16, a.M, a.K, a.lda, 64, BLOCK_K);
  if (a.comm == kCommAllGather)
    ok = ok && make_tmap_2d(&tal, a.A_local, a.is_bf16, a.K, p.rows_per_rank, a.lda_local, BLOCK_K, BLOCK_M);
  else tal = ta;
  if (!a.b_mn_major) ok = ok && make_tmap_2d(&tb, a.B, a.is_bf16, a.K, a.N, a.ldb, BLOCK_K, block_n);
  else ok = ok && make_tmap_2d(&tb, a.B, a.is_bf16, a.N, a.K, a.ldb, 64, BLOCK_K);
  if (a.epilogue == kEpiStore16) {
    ok = ok && make_tmap_2d(&td, a.D, a.is_bf16, a.N, a.M, a.ldd, 64, BLOCK_M);
    if (a.D_pre) ok = ok && make_tmap_2d(&tdp, a.D_pre, a.is_bf16, a.N, a.M, a.ldd, 64, BLOCK_M);
    else tdp = td;
  } else if (a.epilogue == kEpiAccumF32 || a.epilogue == kEpiStoreF32) {
    ok = ok && make_tmap_2d(&td, a.D, a.is_bf16, a.N, a.M, a.ldd, 32, BLOCK_M, /*is_f32=*/true);
    tdp = td;
  } else { td = ta; tdp = ta; }
  if (!ok) return -5;

  PeerMaps pm{};
  for (int s = 0; s < kMaxPeers; ++s) pm.m[s] = ta;
  if (a.comm == kCommReduceScatter) {
    for (int s = 0; s < a.tp_size; ++s) {
      if (!make_tmap_2d(&pm.m[s], a.peer_buf[s], a.is_bf16, a.N, static_cast<uint64_t>(a.tp_size) * p.rows_per_rank, a.N, 64,
                        BLOCK_M))
        return -5;
    }
  }
  if (block_n == 256) return dispatch<256, 4>(a, p, ta, tal, tb, td, tdp, pm, grid);
  return dispatch<128, 6>(a, p, ta, tal, tb, td, tdp, pm, grid);
}

}  // namespace b200
