// MXFP8 (OCP micro-scaling: fp8 elements, one E8M0 scale per 32 elements along K) GEMM on the block-scaled tensor core
// path of sm_100a:
//     D[M,N] (bf16 / fp16, or fp32 +=)  =  ( A8[M,K] * 2^(SFA-127) ) . ( B8[N,K] * 2^(SFB-127) )^T   (+ bias)
// Both operands K-major, one byte per element, K % 128 == 0.  ``tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale``
// consumes the scale factors from TENSOR MEMORY: per k-block (128 bytes of K = four K=32 instructions) the MMA warp
// copies the 128-row x 4-scale chunks of A and B from shared memory into TMEM with ``tcgen05.cp.32x128b.warpx4``
// (UTCCP; issue order with the MMAs is the only synchronisation needed) and selects the byte of the 32-bit scale word
// with the a_sf_id / b_sf_id fields of the instruction descriptor.
//
// Scale-factor storage (produced by mx_quantize below, consumed by plain 1-D bulk copies -- the layout IS the UTCCP
// source layout, so no shuffling happens in the GEMM):
//   chunk = 128 rows x 4 scales = 512 B, byte (row r, scale s) at (r % 32) * 16 + (r / 32) * 4 + s
//   A-role tensor: chunks ordered [row block of 128][k block of 128]
//   B-role tensor: chunks ordered [row tile of 192][k block of 128][2]     (tile height = BN of this kernel; the
//                  second chunk of a tile holds rows 128..191)
// Tile 128 x 192 x 128B: TMEM holds two accumulators (2 x 192 columns) plus two sets of scale columns, so the
// epilogue of tile i overlaps the main loop of tile i+1 exactly as in the bf16 / per-tensor-fp8 kernels.
//
// Reference: fp8_utils.py:384-400 (recipe ``mxfp8`` -> TE MXFP8BlockScaling), TE's rowwise + columnwise quantisation.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <mutex>
#include <string>

#include "b200_api.h"
#include "common.cuh"

namespace b200 {
namespace {

constexpr int BM = 128, BN = 192, BKB = 128;
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kGroupN = 8;
constexpr int kSfaBytes = 512, kSfbBytes = 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSfCol = 2 * BN;          // scale columns start behind the two accumulators (384)

struct PM {
  int M, N, K;
  void* D;
  const void* bias;
  const uint8_t* sfa;
  const uint8_t* sfb;
  int out_bf16, accum_f32, a_e5m2, b_e5m2;
  int num_m_blocks, num_n_blocks, num_k_blocks;
};

struct SmemM {
  static constexpr int kA = BM * BKB;        // 16 KB
  static constexpr int kB = BN * BKB;        // 24 KB
  static constexpr int kC = BM * 128;
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kStages * kA;
  static constexpr int kSfaOff = kBOff + kStages * kB;
  static constexpr int kSfbOff = kSfaOff + kStages * kSfaBytes;
  static constexpr int kCOff = (kSfbOff + kStages * kSfbBytes + 1023) / 1024 * 1024;
  static constexpr int kBarOff = kCOff + 2 * kC;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

// block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): b_sf_id [4,6), a/b format [7,10)/[10,13)
// (0 = E4M3, 1 = E5M2), a/b major bits 15/16 (0 = K-major), N >> 3 at [17,23), scale format bit 23 (1 = UE8M0),
// M >> 4 at [24,29), a_sf_id [29,31), bit 31 = 0 (K = 32 per instruction)
__device__ __forceinline__ uint32_t make_idesc_mx(bool a_e5m2, bool b_e5m2, uint32_t M, uint32_t N, uint32_t sf_id) {
  return (sf_id << 4) | ((a_e5m2 ? 1u : 0u) << 7) | ((b_e5m2 ? 1u : 0u) << 10) | ((N >> 3) << 17) | (1u << 23) |
         ((M >> 4) << 24) | (sf_id << 29);
}
__device__ __forceinline__ void umma_mx_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa,
                                           uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// shared memory -> tensor memory: 32 rows x 128 bit, replicated to the four 32-lane quadrants
__device__ __forceinline__ void utccp_32x128b_x4(uint32_t tmem_dst, uint64_t smem_desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}
// un-swizzled K-major descriptor of a scale chunk: 8-row x 16-byte core matrices, 128 B apart along the rows
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((16u >> 4) & 0x3FFF) << 16;    // LBO: next core matrix along K (unused: one per row group)
  d |= static_cast<uint64_t>((128u >> 4) & 0x3FFF) << 32;   // SBO: next 8-row group
  d |= 1ull << 46;
  return d;                                                  // layout type 0 = no swizzle
}

__device__ __forceinline__ void tile_coords_m(const PM& p, int tile, int& m_blk, int& n_blk) {
  const int group_tiles = p.num_m_blocks * kGroupN;
  const int g = tile / group_tiles;
  int r = tile - g * group_tiles;
  const int n_in_group = min(kGroupN, p.num_n_blocks - g * kGroupN);
  m_blk = r / n_in_group;
  n_blk = g * kGroupN + (r - m_blk * n_in_group);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_d, const PM p) {
  using L = SmemM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem + L::kAOff;
  uint8_t* smem_b = smem + L::kBOff;
  uint8_t* smem_sfa = smem + L::kSfaOff;
  uint8_t* smem_sfb = smem + L::kSfbOff;
  uint8_t* smem_c = smem + L::kCOff;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x, n_ctas = gridDim.x;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_d);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
    mbar_fence_init();
  }
  if (warp_idx == 2) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta; tile < num_tiles; tile += n_ctas) {
        int m_blk, n_blk;
        tile_coords_m(p, tile, m_blk, n_blk);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::kA + L::kB + kSfaBytes + kSfbBytes);
          tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * L::kA, kb * BKB, m_blk * BM);
          tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * L::kB, kb * BKB, n_blk * BN);
          bulk_load_1d(smem_sfa + stage * kSfaBytes,
                       p.sfa + (static_cast<size_t>(m_blk) * p.num_k_blocks + kb) * kSfaBytes, kSfaBytes, &full_bar[stage]);
          bulk_load_1d(smem_sfb + stage * kSfbBytes,
                       p.sfb + (static_cast<size_t>(n_blk) * p.num_k_blocks + kb) * kSfbBytes, kSfbBytes, &full_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      uint32_t sf_set = 0;
      for (int tile = cta; tile < num_tiles; tile += n_ctas) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * L::kA), sb = smem_u32(smem_b + stage * L::kB);
          // scales of this k-block -> TMEM (two column sets alternate so a copy never lands under a running MMA)
          const uint32_t t_sfa = tmem_base + kSfCol + sf_set * 16;
          const uint32_t t_sfb = t_sfa + 4;
          utccp_32x128b_x4(t_sfa, make_sf_desc(smem_u32(smem_sfa + stage * kSfaBytes)));
          utccp_32x128b_x4(t_sfb, make_sf_desc(smem_u32(smem_sfb + stage * kSfbBytes)));
          utccp_32x128b_x4(t_sfb + 4, make_sf_desc(smem_u32(smem_sfb + stage * kSfbBytes + 512)));
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k) {
            const uint32_t idesc = make_idesc_mx(p.a_e5m2 != 0, p.b_e5m2 != 0, BM, BN, static_cast<uint32_t>(k));
            umma_mx_ss(tmem_d, make_smem_desc_sw128(sa + k * 32, 1024, 0), make_smem_desc_sw128(sb + k * 32, 1024, 0), idesc,
                       t_sfa, t_sfb, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);
          sf_set ^= 1;
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp_idx >= 4) {
    const int ew = warp_idx - 4;
    const int epi_tid = threadIdx.x - 128;
    const int row_in_tile = ew * 32 + lane;
    int acc = 0, cbuf = 0;
    uint32_t acc_phase = 0;
    for (int tile = cta; tile < num_tiles; tile += n_ctas) {
      int m_blk, n_blk;
      tile_coords_m(p, tile, m_blk, n_blk);
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
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int gcol0 = n_blk * BN + c * 64;
        if (!p.accum_f32) {
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
                float2 bf = p.out_bf16 ? unpack_bf16x2(bb) : __half22float2(*reinterpret_cast<__half2*>(&bb));
                f[2 * i] += bf.x;
                f[2 * i + 1] += bf.y;
              }
            }
          }
          if (epi_tid == 0) tma_store_wait_read<1>();
          named_bar_sync(1, kEpiThreads);
          uint8_t* sc = smem_c + cbuf * L::kC + row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float a = f[j * 8 + 2 * q], b = f[j * 8 + 2 * q + 1];
              if (p.out_bf16) pk[q] = pack_bf16x2(a, b);
              else { __half2 h = __floats2half2_rn(a, b); pk[q] = *reinterpret_cast<uint32_t*>(&h); }
            }
            *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(2, kEpiThreads);
          if (epi_tid == 0) {
            tma_store_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0, m_blk * BM);
            tma_store_commit();
          }
          cbuf ^= 1;
        } else {
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            const uint32_t* v = hseg == 0 ? v0 : v1;
            if (epi_tid == 0) tma_store_wait_read<1>();
            named_bar_sync(1, kEpiThreads);
            uint8_t* sc = smem_c + cbuf * L::kC + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) = make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
            fence_proxy_async_smem();
            named_bar_sync(2, kEpiThreads);
            if (epi_tid == 0) {
              tma_reduce_add_2d(&tmap_d, smem_c + cbuf * L::kC, gcol0 + hseg * 32, m_blk * BM);
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
  if (warp_idx == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ quantisation
// byte offset of the scale of (row r, 32-element block kb32 along K) in an A-role (role 0) / B-role (role 1) buffer
__device__ __forceinline__ size_t sf_offset(int role, int r, int kb32, int k_blocks) {
  const int kb = kb32 >> 2, ks = kb32 & 3;
  size_t chunk;
  int rr;
  if (role == 0) {
    chunk = static_cast<size_t>(r / 128) * k_blocks + kb;
    rr = r % 128;
  } else {
    const int t = r / BN, r2 = r - t * BN;
    chunk = (static_cast<size_t>(t) * k_blocks + kb) * 2 + r2 / 128;
    rr = r2 % 128;
  }
  return chunk * 512 + (rr % 32) * 16 + (rr / 32) * 4 + ks;
}

// shared exponent of a block (OCP MX v1.0 6.3): floor(log2(amax)) - emax_elem, biased by 127; 0 for an all-zero block
__device__ __forceinline__ int mx_shared_exp(float amax, int emax_elem) {
  if (!(amax > 0.f)) return -127;
  const int e = static_cast<int>((__float_as_uint(amax) >> 23) & 0xFF) - 127;     // floor(log2) of a normal float
  return max(-127, min(127, e - emax_elem));
}
template <typename F8>
__device__ __forceinline__ uint8_t f8_bits(float v) {
  F8 q(v);
  return *reinterpret_cast<uint8_t*>(&q);
}

// One 128 x 128 tile of x [R, C] per CTA.  Row-wise output: q_row [R, C] with blocks along C (scales in role_row layout);
// column-wise output: q_col [C, R] (x transposed) with blocks along R (scales in role_col layout).  Either may be null.
template <typename T, typename F8>
__global__ void __launch_bounds__(256) mx_quantize_kernel(const T* __restrict__ x, int R, int C, uint8_t* __restrict__ q_row,
                                                          uint8_t* __restrict__ sf_row, int role_row, uint8_t* __restrict__ q_col,
                                                          uint8_t* __restrict__ sf_col, int role_col, int emax_elem) {
  __shared__ uint16_t tile[128][130];          // raw 16-bit input values (33 KB); pitch 65 words: conflict-free columns
  auto ld = [&](int r, int c) -> float {
    const uint16_t u = tile[r][c];
    if (std::is_same<T, __nv_bfloat16>::value) return __uint_as_float(static_cast<uint32_t>(u) << 16);
    return __half2float(*reinterpret_cast<const __half*>(&u));
  };
  const int r0 = blockIdx.y * 128, c0 = blockIdx.x * 128;
  for (int i = threadIdx.x; i < 128 * 16; i += 256) {          // 16 vectors of 8 elements per row
    const int r = i >> 4, cv = (i & 15) * 8;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r0 + r < R && c0 + cv < C) u = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r0 + r) * C + c0 + cv);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint32_t*>(&tile[r][cv + 2 * j]) = w[j];
  }
  __syncthreads();
  // 512 (line, block) tasks per direction: line = row (row-wise) or column (column-wise), block = 32 elements
  for (int task = threadIdx.x; task < 512; task += 256) {
    const int line = task >> 2, blk = task & 3;
    if (q_row != nullptr && r0 + line < R && c0 + blk * 32 < C) {
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(ld(line, blk * 32 + j)));
      const int se = mx_shared_exp(amax, emax_elem);
      const float inv = __uint_as_float(static_cast<uint32_t>(127 - se) << 23);      // 2^-se (se in [-127, 127] -> exponent 0..254)
      uint8_t b[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) b[j] = f8_bits<F8>(ld(line, blk * 32 + j) * (se == -127 ? 0.f : inv));
      uint4* dst = reinterpret_cast<uint4*>(q_row + static_cast<size_t>(r0 + line) * C + c0 + blk * 32);
      dst[0] = *reinterpret_cast<uint4*>(b);
      dst[1] = *reinterpret_cast<uint4*>(b + 16);
      sf_row[sf_offset(role_row, r0 + line, (c0 >> 5) + blk, C / 128)] = static_cast<uint8_t>(se + 127);
    }
    if (q_col != nullptr && c0 + line < C && r0 + blk * 32 < R) {
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(ld(blk * 32 + j, line)));
      const int se = mx_shared_exp(amax, emax_elem);
      const float inv = __uint_as_float(static_cast<uint32_t>(127 - se) << 23);
      uint8_t b[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) b[j] = f8_bits<F8>(ld(blk * 32 + j, line) * (se == -127 ? 0.f : inv));
      uint4* dst = reinterpret_cast<uint4*>(q_col + static_cast<size_t>(c0 + line) * R + r0 + blk * 32);
      dst[0] = *reinterpret_cast<uint4*>(b);
      dst[1] = *reinterpret_cast<uint4*>(b + 16);
      sf_col[sf_offset(role_col, c0 + line, (r0 >> 5) + blk, R / 128)] = static_cast<uint8_t>(se + 127);
    }
  }
}

typedef CUresult (*EncodeTiledFnM)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnM encode_m() {
  static EncodeTiledFnM fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnM>(ptr);
  });
  return fn;
}
thread_local std::string g_err_m;

bool tmap_m(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int es, uint64_t inner, uint64_t outer, uint64_t ld,
            uint32_t bi, uint32_t bo) {
  ensure_cuda_context();
  auto fn = encode_m();
  if (!fn) { g_err_m = "cuTensorMapEncodeTiled unavailable"; return false; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * es};
  cuuint32_t box[2] = {bi, bo};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_err_m = "tensor map encode failed: " + std::to_string(static_cast<int>(r)); return false; }
  return true;
}

}  // namespace

const char* gemm_mxfp8_last_error() { return g_err_m.c_str(); }

int64_t mx_sf_bytes(int role, int rows, int K) {
  const int64_t kb = K / 128;
  return role == 0 ? static_cast<int64_t>((rows + 127) / 128) * kb * 512 : static_cast<int64_t>((rows + BN - 1) / BN) * kb * 1024;
}

void mx_quantize(const void* x, int R, int C, bool is_bf16, bool e5m2, void* q_row, void* sf_row, int role_row, void* q_col,
                 void* sf_col, int role_col, cudaStream_t s) {
  dim3 grid((C + 127) / 128, (R + 127) / 128);
  const int emax = e5m2 ? 15 : 8;
  auto qr = reinterpret_cast<uint8_t*>(q_row), sr = reinterpret_cast<uint8_t*>(sf_row);
  auto qc = reinterpret_cast<uint8_t*>(q_col), sc = reinterpret_cast<uint8_t*>(sf_col);
#define MXQ(T, F8) mx_quantize_kernel<T, F8><<<grid, 256, 0, s>>>(reinterpret_cast<const T*>(x), R, C, qr, sr, role_row, qc, sc, role_col, emax)
  if (is_bf16) { if (e5m2) MXQ(__nv_bfloat16, __nv_fp8_e5m2); else MXQ(__nv_bfloat16, __nv_fp8_e4m3); }
  else { if (e5m2) MXQ(__half, __nv_fp8_e5m2); else MXQ(__half, __nv_fp8_e4m3); }
#undef MXQ
}

int gemm_mxfp8(const MxGemmArgs& a) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % 128 != 0) { g_err_m = "mxfp8 GEMM needs K % 128 == 0"; return -1; }
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) || (reinterpret_cast<uintptr_t>(a.D) & 15) ||
      (a.lda % 16) || (a.ldb % 16)) { g_err_m = "mxfp8 GEMM: 16-byte aligned operands"; return -2; }
  PM p{};
  p.M = a.M; p.N = a.N; p.K = a.K; p.D = a.D; p.bias = a.bias;
  p.sfa = reinterpret_cast<const uint8_t*>(a.sfa); p.sfb = reinterpret_cast<const uint8_t*>(a.sfb);
  p.out_bf16 = a.out_bf16 ? 1 : 0; p.accum_f32 = a.accum_f32 ? 1 : 0; p.a_e5m2 = a.a_e5m2 ? 1 : 0; p.b_e5m2 = a.b_e5m2 ? 1 : 0;
  p.num_m_blocks = (a.M + BM - 1) / BM; p.num_n_blocks = (a.N + BN - 1) / BN; p.num_k_blocks = a.K / BKB;
  CUtensorMap ta{}, tb{}, td{};
  bool ok = tmap_m(&ta, a.A, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.K, a.M, a.lda, BKB, BM) &&
            tmap_m(&tb, a.B, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.K, a.N, a.ldb, BKB, BN);
  if (a.accum_f32) ok = ok && tmap_m(&td, a.D, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.N, a.M, a.ldd, 32, BM);
  else ok = ok && tmap_m(&td, a.D, a.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.N, a.M, a.ldd, 64, BM);
  if (!ok) return -5;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemM::kTotal) != cudaSuccess) {
      g_err_m = "cudaFuncSetAttribute failed"; return -10;
    }
    set = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long tiles = static_cast<long>(p.num_m_blocks) * p.num_n_blocks;
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  gemm_mxfp8_kernel<<<grid, kThreads, SmemM::kTotal, a.stream>>>(ta, tb, td, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err_m = std::string("launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}

}  // namespace b200
