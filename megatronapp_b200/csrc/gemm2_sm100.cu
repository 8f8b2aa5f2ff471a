// 2-CTA (cta_group::2) persistent tcgen05 GEMM for sm_100a.
//
// A CTA pair (cluster 2x1x1, same TPC) computes one 256 x 256 output tile per step:
//   * CTA r loads rows [128r, 128r+128) of the A tile and rows [128r, 128r+128) of the B tile
//     (its half of N) -> each SM stages 32 KB per k-block instead of 48 KB, and the tensor core
//     pair fetches the other half of B from the peer's smem.  Per-SM smem traffic drops from
//     ~192 B/clk (reads + TMA writes, measured limiter of the 1-CTA kernel) to ~128 B/clk.
//   * one elected thread of the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256);
//     tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs' mbarriers.
//   * each CTA drains its own 128 TMEM lanes with the same epilogues as the 1-CTA kernel
//     (bf16 TMA store with bias / GELU / pre-activation, fp32 TMA store or TMA reduce-add).
// Six 32 KB stages (192 KB) + 2 x 16 KB epilogue staging; TMEM 2 x 256 columns per CTA.
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
constexpr int kStages = 6;
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kGroupN2 = 8;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address

struct P2 {
  int M, N, K, ldd;
  void* D;
  void* D_pre;
  const void* bias;
  int act, is_bf16;
  int num_m_blocks, num_n_blocks, num_k_blocks;   // m blocks of 256
};

struct Smem2 {
  static constexpr int kA = BM * BK * 2;          // 16 KB
  static constexpr int kB = (BN / 2) * BK * 2;    // 16 KB (this CTA's half of the B tile)
  static constexpr int kC = BM * 64 * 2;          // 16 KB
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kAOff + kStages * kA;
  static constexpr int kCOff = kBOff + kStages * kB;
  static constexpr int kBarOff = kCOff + 2 * kC;
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

__device__ __forceinline__ void tile_coords2(const P2& p, int tile, int& m_blk, int& n_blk) {
  const int group_tiles = p.num_m_blocks * kGroupN2;
  const int g = tile / group_tiles;
  int r = tile - g * group_tiles;
  const int n_in_group = min(kGroupN2, p.num_n_blocks - g * kGroupN2);
  m_blk = r / n_in_group;
  n_blk = g * kGroupN2 + (r - m_blk * n_in_group);
}

template <bool kAMn, bool kBMn, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_d_pre, const P2 p) {
  using L = Smem2;
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
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
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
        int m_blk, n_blk;
        tile_coords2(p, tile, m_blk, n_blk);
        const int m0 = m_blk * 256 + cta_rank * BM;
        const int n0 = n_blk * BN + cta_rank * (BN / 2);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], kPairStageBytes);
          else mbar_arrive_remote(&full_bar[stage], 0);
          uint8_t* sa = smem_a + stage * L::kA;
          uint8_t* sb = smem_b + stage * L::kB;
          if constexpr (!kAMn) {
            tma_load_2d_2sm(&tmap_a, &full_bar[stage], sa, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_2sm(&tmap_a, &full_bar[stage], sa + j * (BK * 128), m0 + j * 64, kb * BK);
          }
          if constexpr (!kBMn) {
            tma_load_2d_2sm(&tmap_b, &full_bar[stage], sb, kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < (BN / 2) / 64; ++j)
              tma_load_2d_2sm(&tmap_b, &full_bar[stage], sb + j * (BK * 128), n0 + j * 64, kb * BK);
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
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
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
            umma2_f16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma2_commit_mc(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma2_commit_mc(&tmem_full_bar[acc]);
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
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      int m_blk, n_blk;
      tile_coords2(p, tile, m_blk, n_blk);
      const int m0 = m_blk * 256 + cta_rank * BM;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
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
        if constexpr (kEpi == kEpiStore16) {
          float f[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            f[i] = __uint_as_float(v0[i]);
            f[32 + i] = __uint_as_float(v1[i]);
          }
          if (p.bias != nullptr) {
            const uint32_t* b32 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(p.bias) + gcol0);
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
          const int n_out = (p.D_pre != nullptr) ? 2 : 1;
          for (int o = 0; o < n_out; ++o) {
            const bool pre = (n_out == 2 && o == 0);
            if (epi_tid == 0) tma_store_wait_read<1>();
            named_bar_sync(1, kEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kC + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float a = f[j * 8 + 2 * q], b = f[j * 8 + 2 * q + 1];
                if (!pre && p.act == 1) { a = gelu2(a); b = gelu2(b); }
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
              tma_store_2d(pre ? &tmap_d_pre : &tmap_d, smem_c + cbuf * L::kC, gcol0, m0);
              tma_store_commit();
            }
            cbuf ^= 1;
          }
        } else {
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            const uint32_t* v = hseg == 0 ? v0 : v1;
            if (epi_tid == 0) tma_store_wait_read<1>();
            named_bar_sync(1, kEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kC + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) =
                  make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
            fence_proxy_async_smem();
            named_bar_sync(2, kEpiThreads);
            if (epi_tid == 0) {
              if constexpr (kEpi == kEpiAccumF32) tma_reduce_add_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0 + hseg * 32, m0);
              else tma_store_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0 + hseg * 32, m0);
              tma_store_commit();
            }
            cbuf ^= 1;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (epi_tid == 0) tma_store_wait_all<0>();
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

template <bool AM, bool BMj, int EPI>
int launch2(const GemmArgs& a, const P2& p, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td,
            const CUtensorMap& tdp, int grid) {
  auto kern = gemm2_kernel<AM, BMj, EPI>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2::kTotal) != cudaSuccess) {
      g_err2 = "cudaFuncSetAttribute failed"; return -10;
    }
    set = true;
  }
  kern<<<grid, kThreads, Smem2::kTotal, a.stream>>>(ta, tb, td, tdp, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err2 = std::string("launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}

}  // namespace

const char* gemm2_last_error() { return g_err2.c_str(); }

bool gemm2_supported(const GemmArgs& a) {
  if (a.comm != kCommNone) return false;
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
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (a.num_sms > 0) sms = a.num_sms;
  int clusters = sms / 2;
  const long tiles = static_cast<long>(p.num_m_blocks) * p.num_n_blocks;
  if (tiles < clusters) clusters = static_cast<int>(tiles);
  const int dt = a.is_bf16 ? 0 : 1;
  CUtensorMap ta{}, tb{}, td{}, tdp{};
  bool ok = true;
  if (!a.a_mn_major) ok = ok && tmap2(&ta, a.A, dt, a.K, a.M, a.lda, BK, BM);
  else ok = ok && tmap2(&ta, a.A, dt, a.M, a.K, a.lda, 64, BK);
  if (!a.b_mn_major) ok = ok && tmap2(&tb, a.B, dt, a.K, a.N, a.ldb, BK, BN / 2);
  else ok = ok && tmap2(&tb, a.B, dt, a.N, a.K, a.ldb, 64, BK);
  if (a.epilogue == kEpiStore16) {
    ok = ok && tmap2(&td, a.D, dt, a.N, a.M, a.ldd, 64, BM);
    if (a.D_pre) ok = ok && tmap2(&tdp, a.D_pre, dt, a.N, a.M, a.ldd, 64, BM);
    else tdp = td;
  } else {
    ok = ok && tmap2(&td, a.D, 2, a.N, a.M, a.ldd, 32, BM);
    tdp = td;
  }
  if (!ok) return -5;
  const int grid = clusters * 2;
#define G2_CASE(AM, BMJ, EPI) \
  if (a.a_mn_major == AM && a.b_mn_major == BMJ && a.epilogue == EPI) return launch2<AM, BMJ, EPI>(a, p, ta, tb, td, tdp, grid);
  G2_CASE(false, false, kEpiStore16)
  G2_CASE(false, true, kEpiStore16)
  G2_CASE(true, true, kEpiStore16)
  G2_CASE(true, false, kEpiStore16)
  G2_CASE(true, true, kEpiAccumF32)
  G2_CASE(true, true, kEpiStoreF32)
  G2_CASE(false, false, kEpiAccumF32)
  G2_CASE(false, false, kEpiStoreF32)
#undef G2_CASE
  g_err2 = "unsupported combination for the 2-CTA kernel";
  return -3;
}

}  // namespace b200
