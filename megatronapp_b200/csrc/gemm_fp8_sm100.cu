// FP8 (E4M3 / E5M2) tcgen05 GEMM with per-tensor scaling:
//     D[M,N] (bf16 / fp16, or fp32 +=)  =  alpha_a * alpha_b * ( A8[M,K] . B8[N,K]^T )  (+ bias)
// Both operands K-major, one byte per element (tcgen05.mma.kind::f8f6f4, K = 32 per instruction,
// 128 elements per 128-byte swizzle row).  Persistent, warp-specialised like gemm_sm100.cu: warp 0 TMA
// producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue (TMEM -> registers ->
// dequant scale -> swizzled smem -> TMA store / TMA reduce-add).  The dequantisation scales are
// device pointers so that the amax -> scale -> cast -> GEMM chain never synchronises with the host.
//
// Replaces the reference's Transformer-Engine fp8 path (core/fp8_utils.py:384-400 recipes
// "tensorwise"/"delayed"; extensions/transformer_engine.py TELinear fp8 GEMMs).
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

constexpr int BM = 128, BN = 256, BKB = 128;   // tile (BKB = bytes = fp8 elements along K per stage)
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kGroupN = 8;

struct P8 {
  int M, N, K;
  void* D;
  const void* bias;            // optional 16-bit bias [N]
  const float* scale_a_inv;    // device scalars (dequantisation factors)
  const float* scale_b_inv;
  int out_bf16;                // 16-bit output type
  int accum_f32;               // 1: D is fp32 and receives "+="
  int a_e5m2, b_e5m2;
  int num_m_blocks, num_n_blocks, num_k_blocks;
};

struct Smem8 {
  static constexpr int kA = BM * BKB;        // 16 KB
  static constexpr int kB = BN * BKB;        // 32 KB
  static constexpr int kC = BM * 128;        // one 128-byte-wide column slab of the output tile
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kStages * kA;
  static constexpr int kCOff = kBOff + kStages * kB;
  static constexpr int kBarOff = kCOff + 2 * kC;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

__device__ __forceinline__ void umma_f8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 instruction descriptor: c_format F32 (bit 4), a/b format at [7,10)/[10,13): 0 = E4M3, 1 = E5M2
__host__ __device__ constexpr uint32_t make_idesc_f8(bool a_e5m2, bool b_e5m2, uint32_t M, uint32_t N) {
  return (1u << 4) | ((a_e5m2 ? 1u : 0u) << 7) | ((b_e5m2 ? 1u : 0u) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void tile_coords8(const P8& p, int tile, int& m_blk, int& n_blk) {
  const int group_tiles = p.num_m_blocks * kGroupN;
  const int g = tile / group_tiles;
  int r = tile - g * group_tiles;
  const int n_in_group = min(kGroupN, p.num_n_blocks - g * kGroupN);
  m_blk = r / n_in_group;
  n_blk = g * kGroupN + (r - m_blk * n_in_group);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_d, const P8 p) {
  using L = Smem8;
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
  if (warp_idx == 2) tmem_alloc<2 * BN>(tmem_ptr_smem);
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
        tile_coords8(p, tile, m_blk, n_blk);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::kA + L::kB);
          tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * L::kA, kb * BKB, m_blk * BM);
          tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * L::kB, kb * BKB, n_blk * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = make_idesc_f8(p.a_e5m2 != 0, p.b_e5m2 != 0, BM, BN);
      for (int tile = cta; tile < num_tiles; tile += n_ctas) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * L::kA), sb = smem_u32(smem_b + stage * L::kB);
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k) {
            // K-major, 32 one-byte elements per MMA: advance 32 B inside the 128 B swizzle atom
            umma_f8_ss(tmem_d, make_smem_desc_sw128(sa + k * 32, 1024, 0), make_smem_desc_sw128(sb + k * 32, 1024, 0),
                       idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp_idx >= 4) {
    const int ew = warp_idx - 4;
    const int epi_tid = threadIdx.x - 128;
    const int row_in_tile = ew * 32 + lane;
    const float alpha = __ldg(p.scale_a_inv) * __ldg(p.scale_b_inv);
    int acc = 0, cbuf = 0;
    uint32_t acc_phase = 0;
    for (int tile = cta; tile < num_tiles; tile += n_ctas) {
      int m_blk, n_blk;
      tile_coords8(p, tile, m_blk, n_blk);
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
            f[i] = __uint_as_float(v0[i]) * alpha;
            f[32 + i] = __uint_as_float(v1[i]) * alpha;
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
            for (int j = 0; j < 8; ++j) {
              *reinterpret_cast<uint4*>(sc + ((j ^ (row_in_tile & 7)) << 4)) = make_uint4(
                  __float_as_uint(__uint_as_float(v[j * 4]) * alpha), __float_as_uint(__uint_as_float(v[j * 4 + 1]) * alpha),
                  __float_as_uint(__uint_as_float(v[j * 4 + 2]) * alpha), __float_as_uint(__uint_as_float(v[j * 4 + 3]) * alpha));
            }
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
  if (warp_idx == 2) tmem_dealloc<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------ quantisation helpers
// amax over a 16-bit tensor (atomicMax on the bit pattern of a non-negative float)
template <typename T>
__global__ void amax_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ amax) {
  float m = 0.f;
  const int64_t nv = n / 8;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nv;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + v * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f;
      if (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) f = unpack_bf16x2(w[i]);
      else f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

// scale = fp8_max / amax (1 when amax == 0 or not finite); scale_inv = 1 / scale
__global__ void scale_from_amax_kernel(const float* amax, float fp8_max, float margin_pow2, float* scale, float* scale_inv) {
  float a = *amax, s = 1.f;
  if (a > 0.f && isfinite(a)) s = fp8_max / a * margin_pow2;
  if (!isfinite(s)) s = 1.f;
  *scale = s;
  *scale_inv = 1.f / s;
}

template <typename T, typename F8>
__device__ __forceinline__ uint8_t to_f8(float v) {
  F8 q(v);
  return *reinterpret_cast<uint8_t*>(&q);
}

// x [rows, cols] 16-bit -> q [rows, cols] fp8 and (optionally) qt [cols, rows] fp8, one 64x64 tile per CTA
template <typename T, typename F8>
__global__ void __launch_bounds__(256) cast_transpose_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                             uint8_t* __restrict__ q, uint8_t* __restrict__ qt,
                                                             int rows, int cols) {
  __shared__ uint8_t tile[64][64 + 4];
  const float s = __ldg(scale);
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  // 256 threads: thread -> (row = tid / 8 + 32 * pass, 8 consecutive columns)
  const int tc = (threadIdx.x & 7) * 8, tr = threadIdx.x >> 3;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = r0 + tr + pass * 32, c = c0 + tc;
    uint8_t b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < rows && c < cols) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * cols + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f;
        if (std::is_same<T, __nv_bfloat16>::value) f = unpack_bf16x2(w[i]);
        else f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        b[2 * i] = to_f8<T, F8>(f.x * s);
        b[2 * i + 1] = to_f8<T, F8>(f.y * s);
      }
      *reinterpret_cast<uint2*>(q + static_cast<size_t>(r) * cols + c) = *reinterpret_cast<uint2*>(b);
    }
    if (qt != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[tr + pass * 32][tc + i] = b[i];
    }
  }
  if (qt == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int oc = tr + pass * 32;            // column of x = row of x^T
    const int c = c0 + oc, r = r0 + tc;
    if (c < cols && r < rows) {
      uint8_t b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = tile[tc + i][oc];
      *reinterpret_cast<uint2*>(qt + static_cast<size_t>(c) * rows + r) = *reinterpret_cast<uint2*>(b);
    }
  }
}

typedef CUresult (*EncodeFn8)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn8 encode8() {
  static EncodeFn8 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn8>(ptr);
  });
  return fn;
}
thread_local std::string g_err8;

bool tmap8(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem, uint64_t inner, uint64_t outer,
           uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  ensure_cuda_context();
  EncodeFn8 fn = encode8();
  if (!fn) { g_err8 = "cuTensorMapEncodeTiled unavailable"; return false; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * elem};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_err8 = "cuTensorMapEncodeTiled failed: code " + std::to_string(static_cast<int>(r)); return false; }
  return true;
}

}  // namespace

const char* gemm_fp8_last_error() { return g_err8.c_str(); }

int gemm_fp8(const Fp8GemmArgs& a) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) { g_err8 = "empty problem"; return -1; }
  if (a.K % 16 || a.lda % 16 || a.ldb % 16) { g_err8 = "fp8 operands need K and leading dims multiples of 16 bytes"; return -2; }
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B) | reinterpret_cast<uintptr_t>(a.D)) % 16) {
    g_err8 = "operands must be 16-byte aligned"; return -2;
  }
  if (a.accum_f32 ? (a.ldd % 4) : (a.ldd % 8)) { g_err8 = "output leading dim must be 16-byte aligned"; return -2; }
  P8 p{};
  p.M = a.M; p.N = a.N; p.K = a.K; p.D = a.D; p.bias = a.bias;
  p.scale_a_inv = a.scale_a_inv; p.scale_b_inv = a.scale_b_inv;
  p.out_bf16 = a.out_bf16 ? 1 : 0; p.accum_f32 = a.accum_f32 ? 1 : 0;
  p.a_e5m2 = a.a_e5m2 ? 1 : 0; p.b_e5m2 = a.b_e5m2 ? 1 : 0;
  p.num_m_blocks = (a.M + BM - 1) / BM;
  p.num_n_blocks = (a.N + BN - 1) / BN;
  p.num_k_blocks = (a.K + BKB - 1) / BKB;
  CUtensorMap ta{}, tb{}, td{};
  bool ok = tmap8(&ta, a.A, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.K, a.M, a.lda, BKB, BM);
  ok = ok && tmap8(&tb, a.B, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.K, a.N, a.ldb, BKB, BN);
  if (a.accum_f32) ok = ok && tmap8(&td, a.D, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.N, a.M, a.ldd, 32, BM);
  else ok = ok && tmap8(&td, a.D, a.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a.N, a.M,
                        a.ldd, 64, BM);
  if (!ok) return -5;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_fp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem8::kTotal);
    if (e != cudaSuccess) { g_err8 = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -10; }
    attr_set = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < sms ? tiles : sms;
  gemm_fp8_kernel<<<grid, kThreads, Smem8::kTotal, a.stream>>>(ta, tb, td, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err8 = std::string("gemm_fp8 launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}

void fp8_amax(const void* x, int64_t n, bool is_bf16, float* amax, cudaStream_t s) {
  const int grid = 148 * 4;
  if (is_bf16) amax_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, amax);
  else amax_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(x), n, amax);
}

void fp8_scale_from_amax(const float* amax, bool e5m2, int margin, float* scale, float* scale_inv, cudaStream_t s) {
  scale_from_amax_kernel<<<1, 1, 0, s>>>(amax, e5m2 ? 57344.f : 448.f, exp2f(-static_cast<float>(margin)), scale, scale_inv);
}

void fp8_cast_transpose(const void* x, const float* scale, void* q, void* qt, int rows, int cols, bool is_bf16, bool e5m2,
                        cudaStream_t s) {
  dim3 grid((cols + 63) / 64, (rows + 63) / 64);
  uint8_t* q8 = reinterpret_cast<uint8_t*>(q);
  uint8_t* qt8 = reinterpret_cast<uint8_t*>(qt);
#define CT(T, F) cast_transpose_kernel<T, F><<<grid, 256, 0, s>>>(reinterpret_cast<const T*>(x), scale, q8, qt8, rows, cols)
  if (is_bf16) { if (e5m2) CT(__nv_bfloat16, __nv_fp8_e5m2); else CT(__nv_bfloat16, __nv_fp8_e4m3); }
  else { if (e5m2) CT(__half, __nv_fp8_e5m2); else CT(__half, __nv_fp8_e4m3); }
#undef CT
}

}  // namespace b200
