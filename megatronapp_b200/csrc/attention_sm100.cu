// Flash attention for sm_100a: tcgen05 MMAs with S / O accumulators in TMEM, TMA-fed K/V ring,
// two 128-row query tiles per CTA processed ping-pong by two softmax warpgroups.
//
// Forward  (attn_fwd_kernel):  S = Q K^T -> online softmax (exp2 domain, lazy rescale) ->
//                              P (bf16, swizzled smem) -> O += P V (O stays in TMEM).
// Backward (two kernels, deterministic, no atomics):
//   attn_bwd_dkdv_kernel : CTA per 128-key block, loops over 64-query blocks
//                          S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
//   attn_bwd_dq_kernel   : CTA per 128-query block, loops over 64-key blocks
//                          S = Q K^T, dP = dO V^T, dQ += dS K
// Operands are strided [b, h, s, d] views (last dim contiguous) so the Megatron [s, b, h, d]
// activations are consumed in place (4-D tensor maps); outputs are written in [s, b, h, d].
//
// Replaces reference dot_product_attention.py:106-208 (baddbmm -> fused softmax -> bmm) and the
// TE/cuDNN path extensions/transformer_engine.py:609-.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>

#include "b200_api.h"
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kBlockQ = 128;    // query rows per tile
constexpr int kBlockKV = 128;   // keys per block (forward)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // lazy O rescale: only when the row max grows by > 2^8

__device__ __forceinline__ void tma_load_4d(const void* tmap, uint64_t* bar, void* smem_dst, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool bf16) {
  if (bf16) return pack_bf16x2(lo, hi);
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte unit `u16` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int u16) {
  return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(u16 ^ (r & 7)) << 4);
}

struct FwdParams {
  float* lse;            // [b, h, sq]
  void* o;               // [sq, b, h, d] strides below
  int64_t o_sb, o_sh, o_ss;
  int b, hq, hkv, sq, sk;
  float scale_log2;      // softmax scale * log2(e)
  int causal;
  int is_bf16;
};

// kPTmem: P (bf16) overwrites the S columns in TMEM and feeds the PV product as the TMEM A operand
// (halves the shared-memory traffic of a key block and frees room for a second V stage);
// otherwise P goes through swizzled shared memory.
template <int D, bool kPTmem>
struct FwdSmem {
  static constexpr int kDC = D / 64;                    // 64-element (128 B) chunks along d
  static constexpr int kTile = kBlockQ * 128 * kDC;     // bytes of a [128 x D] 16-bit tile
  static constexpr int kKStages = 2;
  static constexpr int kVStages = (D == 64 || kPTmem) ? 2 : 1;
  static constexpr int kP = kPTmem ? 0 : kBlockQ * kBlockKV * 2;     // 32 KB
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + 2 * kTile;
  static constexpr int kVOff = kKOff + kKStages * kTile;
  static constexpr int kPOff = kVOff + kVStages * kTile;
  static constexpr int kBarOff = kPOff + 2 * kP;
  static constexpr int kTotal = kBarOff + 256 + 1024;   // + barriers + alignment slack
};

// ------------------------------------------------------------------------------------ forward
template <int D, bool kPTmem>
__global__ void __launch_bounds__(384, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const FwdParams p) {
  using L = FwdSmem<D, kPTmem>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + L::kQOff;
  uint8_t* sK = smem + L::kKOff;
  uint8_t* sV = smem + L::kVOff;
  uint8_t* sP = smem + L::kPOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 2
  uint64_t* k_empty = bars + 3;       // 2
  uint64_t* v_full = bars + 5;        // 2
  uint64_t* v_empty = bars + 7;       // 2
  uint64_t* s_full = bars + 9;        // 2 (per q tile)
  uint64_t* p_ready = bars + 11;      // 2
  uint64_t* pv_done = bars + 13;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // heaviest (largest q index) CTAs first for causal load balance
  const int qblk2 = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);
  const int h = blockIdx.y, bi = blockIdx.z;
  const int hk = h / (p.hq / p.hkv);
  const int q0_0 = qblk2 * 2 * kBlockQ, q0_1 = q0_0 + kBlockQ;
  const int nkv_all = (p.sk + kBlockKV - 1) / kBlockKV;
  auto blocks_for = [&](int q0) {
    if (q0 >= p.sq) return 0;
    return p.causal ? min(nkv_all, q0 / kBlockKV + 1) : nkv_all;
  };
  const int nkv_0 = blocks_for(q0_0), nkv_1 = blocks_for(q0_1);
  const int nmax = max(nkv_0, nkv_1);

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&pv_done[i], 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // S[t] at col t*128
  const uint32_t tmem_O = tmem_base + 256;    // O[t] at col 256 + t*D

  if (warp_idx == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      const int ntiles = (nkv_1 > 0) ? 2 : 1;
      mbar_arrive_expect_tx(q_full, ntiles * L::kTile);
      for (int t = 0; t < ntiles; ++t)
        for (int c = 0; c < L::kDC; ++c)
          tma_load_4d(&tmap_q, q_full, sQ + t * L::kTile + c * (kBlockQ * 128), c * 64, h, bi, t == 0 ? q0_0 : q0_1);
      for (int j = 0; j < nmax; ++j) {
        const int ks = j % L::kKStages, vs = j % L::kVStages;
        mbar_wait(&k_empty[ks], ((j / L::kKStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], L::kTile);
        for (int c = 0; c < L::kDC; ++c)
          tma_load_4d(&tmap_k, &k_full[ks], sK + ks * L::kTile + c * (kBlockKV * 128), c * 64, hk, bi, j * kBlockKV);
        mbar_wait(&v_empty[vs], ((j / L::kVStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], L::kTile);
        for (int c = 0; c < L::kDC; ++c)
          tma_load_4d(&tmap_v, &v_full[vs], sV + vs * L::kTile + c * (kBlockKV * 128), c * 64, hk, bi, j * kBlockKV);
      }
    }
  } else if (warp_idx == 1) {
    // ============================================================ MMA issuer
    if (lane == 0 && nmax > 0) {
      const uint32_t idesc_s = make_idesc_f16(p.is_bf16 != 0, kBlockQ, kBlockKV, false, false);
      const uint32_t idesc_o = make_idesc_f16(p.is_bf16 != 0, kBlockQ, D, false, true);
      auto issue_S = [&](int t, int j) {
        const int ks = j % L::kKStages;
        const uint32_t a0 = smem_u32(sQ + t * L::kTile), b0 = smem_u32(sK + ks * L::kTile);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * (kBlockQ * 128) + (k & 3) * 32;
          umma_f16_ss(tmem_S + t * 128, make_smem_desc_sw128(a0 + off, 1024, 0),
                      make_smem_desc_sw128(b0 + off, 1024, 0), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_PV = [&](int t, int j) {
        const int vs = j % L::kVStages;
        const uint32_t a0 = smem_u32(sP + t * L::kP), b0 = smem_u32(sV + vs * L::kTile);
#pragma unroll
        for (int k = 0; k < kBlockKV / 16; ++k) {
          // V is MN-major for this product (N = d contiguous): 16 key rows per k-step,
          // LBO = next 64-wide d chunk
          const uint64_t bdesc = make_smem_desc_sw128(b0 + k * 16 * 128, 1024, kBlockKV * 128);
          if constexpr (kPTmem) {
            // A = P from TMEM: row = lane, 16 keys = 8 packed 32-bit columns per k-step
            umma_f16_ts(tmem_O + t * D, tmem_S + t * 128 + k * 8, bdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
          } else {
            const uint32_t aoff = (k >> 2) * (kBlockQ * 128) + (k & 3) * 32;
            umma_f16_ss(tmem_O + t * D, make_smem_desc_sw128(a0 + aoff, 1024, 0), bdesc, idesc_o,
                        (j | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&pv_done[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (nkv_0 > 0) issue_S(0, 0);
      if (nkv_1 > 0) issue_S(1, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nmax; ++j) {
        const int vs = j % L::kVStages;
        mbar_wait(&v_full[vs], (j / L::kVStages) & 1);
        if (j + 1 < nmax) mbar_wait(&k_full[(j + 1) % L::kKStages], ((j + 1) / L::kKStages) & 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int nt = t == 0 ? nkv_0 : nkv_1;
          if (j >= nt) continue;
          mbar_wait(&p_ready[t], j & 1);
          tc_fence_after();
          issue_PV(t, j);
          if (j + 1 < nt) issue_S(t, j + 1);
        }
        umma_commit(&v_empty[vs]);
        if (j + 1 < nmax) umma_commit(&k_empty[(j + 1) % L::kKStages]);
      }
    }
  } else if (warp_idx >= 4) {
    // ============================================================ softmax / correction / epilogue
    const int t = (warp_idx - 4) >> 2;            // q tile of this warpgroup
    const int quad = warp_idx & 3;                // TMEM lane quadrant
    const int r = quad * 32 + lane;               // row inside the tile
    const int q0 = t == 0 ? q0_0 : q0_1;
    const int qrow = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_S + t * 128 + lane_addr;
    const uint32_t tO = tmem_O + t * D + lane_addr;
    uint8_t* sPt = sP + t * L::kP;
    float m = -INFINITY, l = 0.f;   // running max (exp2 domain, already scaled) and row sum
    const int n = t == 0 ? nkv_0 : nkv_1;
    const float sc = p.scale_log2;
    const bool bf = p.is_bf16 != 0;

    // one key block: two sweeps over the S tile in TMEM (row max, then exp2 / pack / store P)
    auto block = [&](int j, auto masked_tag) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      const int kbase = j * kBlockKV;
      const int limit = (kMasked ? min(p.sk, p.causal ? qrow + 1 : p.sk) : 0) - kbase;   // valid: col < limit
      float mx = -INFINITY;
      {
        uint32_t v[2][32];
        tmem_ld_32x32(tS, v[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32(tS + (c + 1) * 32, v[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[c & 1][i]);
            if (kMasked && c * 32 + i >= limit) x = -INFINITY;
            mx = fmaxf(mx, x);
          }
        }
      }
      float m_new = fmaxf(m, mx * sc);             // softmax scale is positive
      if (m_new == -INFINITY) m_new = 0.f;         // fully masked so far: keep everything finite
      const bool grow = (m_new - m) > kRescaleThreshold;   // also true on the first block (m = -inf)
      const float m_cur = grow ? m_new : m;
      const float alpha = grow ? fast_exp2(m - m_cur) : 1.f;
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tO + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32(tO + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      l *= alpha;
      float rs0 = 0.f, rs1 = 0.f;
      {
        uint32_t v[2][32];
        tmem_ld_32x32(tS, v[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32(tS + (c + 1) * 32, v[(c + 1) & 1]);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(v[c & 1][i]), sc, -m_cur));
            float p1 = fast_exp2(fmaf(__uint_as_float(v[c & 1][i + 1]), sc, -m_cur));
            if (kMasked) {
              if (c * 32 + i >= limit) p0 = 0.f;
              if (c * 32 + i + 1 >= limit) p1 = 0.f;
            }
            rs0 += p0;
            rs1 += p1;
            pk[i >> 1] = pack16(p0, p1, bf);
          }
          if constexpr (kPTmem) {
            // packed P chunk c -> TMEM columns [16c, 16c+16) of this tile's S region; the fp32 S
            // columns it overwrites ([32c', ...) with c' < c, or chunk 0 itself) are already consumed
            tmem_st_32x16(tS + c * 16, pk);
          } else {
            // keys c*32 .. c*32+31 -> chunk (c>>1), 16-byte units (c&1)*4 .. +3 of row r
            uint8_t* chunk = sPt + (c >> 1) * (kBlockQ * 128);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(chunk + sw128_off(r, (c & 1) * 4 + u)) =
                  make_uint4(pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
          }
        }
      }
      l += rs0 + rs1;
      m = m_cur;
      if constexpr (kPTmem) tmem_st_wait();
      tc_fence_before();
      if constexpr (!kPTmem) fence_proxy_async_smem();
      mbar_arrive(&p_ready[t]);
    };

    for (int j = 0; j < n; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const int kbase = j * kBlockKV;
      // diagonal block of a causal tile, or the ragged tail of the key sequence
      const bool need_mask = (p.causal && kbase + kBlockKV - 1 > q0) || (kbase + kBlockKV > p.sk);
      if (need_mask) block(j, std::true_type{});
      else block(j, std::false_type{});
    }
    if (n > 0) {
      mbar_wait(&pv_done[t], (n - 1) & 1);
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      const bool valid = qrow < p.sq;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                      (static_cast<int64_t>(qrow) * p.o_ss + static_cast<int64_t>(bi) * p.o_sb +
                       static_cast<int64_t>(h) * p.o_sh) * 2;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tO + c * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack16(__uint_as_float(v[u * 8 + 0]) * inv_l, __uint_as_float(v[u * 8 + 1]) * inv_l, bf);
            w.y = pack16(__uint_as_float(v[u * 8 + 2]) * inv_l, __uint_as_float(v[u * 8 + 3]) * inv_l, bf);
            w.z = pack16(__uint_as_float(v[u * 8 + 4]) * inv_l, __uint_as_float(v[u * 8 + 5]) * inv_l, bf);
            w.w = pack16(__uint_as_float(v[u * 8 + 6]) * inv_l, __uint_as_float(v[u * 8 + 7]) * inv_l, bf);
            *reinterpret_cast<uint4*>(orow + c * 64 + u * 16) = w;
          }
        }
      }
      if (valid)
        p.lse[(static_cast<int64_t>(bi) * p.hq + h) * p.sq + qrow] = l > 0.f ? (m + log2f(l)) * kLn2 : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc<512>(tmem_base);
}


// ------------------------------------------------------------------------------------ backward
// delta[b, h, s] = sum_d O * dO   (one warp per row)
__global__ void attn_delta_kernel(const void* __restrict__ o, const void* __restrict__ d_o, float* __restrict__ delta,
                                  const float* __restrict__ lse, float* __restrict__ lse2, int b, int h, int s, int d, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                                  int64_t do_sb, int64_t do_sh, int64_t do_ss, int is_bf16) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t total = static_cast<int64_t>(b) * h * s;
  if (row >= total) return;
  const int si = static_cast<int>(row % s);
  const int hi = static_cast<int>((row / s) % h);
  const int bi = static_cast<int>(row / (static_cast<int64_t>(s) * h));
  const uint32_t* po = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(o) + bi * o_sb + hi * o_sh + si * o_ss);
  const uint32_t* pd = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(d_o) + bi * do_sb + hi * do_sh + si * do_ss);
  float acc = 0.f;
  for (int i = lane_id(); i < d / 2; i += 32) {
    const uint32_t a = po[i], c = pd[i];
    float2 fa, fc;
    if (is_bf16) { fa = unpack_bf16x2(a); fc = unpack_bf16x2(c); }
    else { fa = __half22float2(*reinterpret_cast<const __half2*>(&a)); fc = __half22float2(*reinterpret_cast<const __half2*>(&c)); }
    acc += fa.x * fc.x + fa.y * fc.y;
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) {
    delta[row] = acc;
    const float lv = lse[row];
    lse2[row] = (lv == -INFINITY) ? INFINITY : lv * kLog2e;   // exp2-domain LSE; fully masked rows give P = 0
  }
}

struct BwdParams {
  const float* lse;     // [b, hq, sq] LSE in the exp2 domain (lse * log2e)
  const float* delta;   // [b, hq, sq]
  void* out1;           // dq (kDQ) or dv
  void* out2;           // unused (kDQ) or dk
  int64_t o1_sb, o1_sh, o1_ss, o2_sb, o2_sh, o2_ss;
  int b, hq, hkv, sq, sk;
  float scale_log2, scale;
  int causal, is_bf16;
};

constexpr int kBwdRows = 128;   // resident rows per CTA (keys for dK/dV, queries for dQ)
constexpr int kBwdCols = 64;    // streamed rows per iteration (queries for dK/dV, keys for dQ)

template <int D, bool kDQ>
struct BwdSmem {
  static constexpr int kDC = D / 64;
  static constexpr int kStages = 4;                     // streamed-tile ring depth (hides TMA latency)
  static constexpr int kX = kBwdRows * 128 * kDC;       // resident tile bytes
  static constexpr int kY = kBwdCols * 128 * kDC;       // streamed tile bytes
  static constexpr int kX1Off = 0;
  static constexpr int kX2Off = kX;
  static constexpr int kStageOff = 2 * kX;              // stage s: Y1 at s*2*kY, Y2 at s*2*kY + kY
  static constexpr int kStatOff = kStageOff + kStages * 2 * kY;   // stage s: lse2[64], delta[64] floats
  static constexpr int kBarOff = kStatOff + kStages * 512;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

// kDQ = false: rows are keys   (X1 = K, X2 = V, Y1 = Q_i, Y2 = dO_i): acc1 = dV, acc2 = dK
// kDQ = true : rows are queries (X1 = Q, X2 = dO, Y1 = K_j, Y2 = V_j): acc1 = dQ
// P / dS (16-bit) overwrite the T1 / T2 accumulator columns in TMEM and feed the accumulate MMAs as
// the TMEM A operand, so the only shared-memory traffic is the operand tiles themselves.
template <int D, bool kDQ>
__global__ void __launch_bounds__(256, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_x2,
                const __grid_constant__ CUtensorMap tmap_y1, const __grid_constant__ CUtensorMap tmap_y2,
                const BwdParams p) {
  using L = BwdSmem<D, kDQ>;
  constexpr int kS = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX1 = smem + L::kX1Off;
  uint8_t* sX2 = smem + L::kX2Off;
  uint8_t* sY = smem + L::kStageOff;
  float* sStat = reinterpret_cast<float*>(smem + L::kStatOff);   // stage s: lse2 at s*128, delta at s*128+64
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* x_full = bars;               // 1
  uint64_t* y_full = bars + 1;           // kS
  uint64_t* y_empty = bars + 1 + kS;     // kS
  uint64_t* t_full = bars + 1 + 2 * kS;  // 2
  uint64_t* pds_ready = t_full + 2;      // 2
  uint64_t* acc_done = pds_ready + 2;    // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bi = blockIdx.z;
  const int G = p.hq / p.hkv;
  const int r0 = blockIdx.x * kBwdRows;
  // iteration space
  int col_blk0, n_col, n_heads, head0;
  if (kDQ) {
    head0 = blockIdx.y; n_heads = 1;
    col_blk0 = 0;
    const int kmax = p.causal ? min(p.sk, r0 + kBwdRows) : p.sk;
    n_col = (kmax + kBwdCols - 1) / kBwdCols;
  } else {
    head0 = blockIdx.y * G; n_heads = G;
    col_blk0 = p.causal ? r0 / kBwdCols : 0;
    n_col = (p.sq + kBwdCols - 1) / kBwdCols - col_blk0;
    if (n_col < 0) n_col = 0;
  }
  const int n_iter = n_col * n_heads;

  if (threadIdx.x == 0) {
    mbar_init(x_full, 1);
    for (int i = 0; i < kS; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&pds_ready[i], 128); }
    mbar_init(acc_done, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x1); tma_prefetch_desc(&tmap_x2); tma_prefetch_desc(&tmap_y1); tma_prefetch_desc(&tmap_y2);
  }
  if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_T1 = tmem_base;          // T1[buf] at col buf*64  (P overwrites its first 32 cols)
  const uint32_t tmem_T2 = tmem_base + 128;    // T2[buf] at col 128 + buf*64 (dS overwrites its first 32 cols)
  const uint32_t tmem_A1 = tmem_base + 256;    // dV or dQ
  const uint32_t tmem_A2 = tmem_base + 256 + D;  // dK

  if (warp_idx == 0) {
    // ============================================================ TMA producer
    if (lane == 0 && n_iter > 0) {
      const int hx = blockIdx.y;   // head of the resident tiles (kv head for dK/dV, q head for dQ)
      mbar_arrive_expect_tx(x_full, 2 * L::kX);
      for (int c = 0; c < L::kDC; ++c) {
        tma_load_4d(&tmap_x1, x_full, sX1 + c * (kBwdRows * 128), c * 64, hx, bi, r0);
        tma_load_4d(&tmap_x2, x_full, sX2 + c * (kBwdRows * 128), c * 64, hx, bi, r0);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int ys = it % kS;
        const int head = kDQ ? blockIdx.y / G : head0 + it / n_col;   // head of the streamed tiles
        const int c0 = (col_blk0 + it % n_col) * kBwdCols;
        mbar_wait(&y_empty[ys], ((it / kS) & 1) ^ 1);
        mbar_arrive_expect_tx(&y_full[ys], 2 * L::kY + (kDQ ? 0 : 512));
        uint8_t* y1 = sY + ys * 2 * L::kY;
        uint8_t* y2 = y1 + L::kY;
        for (int c = 0; c < L::kDC; ++c) {
          tma_load_4d(&tmap_y1, &y_full[ys], y1 + c * (kBwdCols * 128), c * 64, head, bi, c0);
          tma_load_4d(&tmap_y2, &y_full[ys], y2 + c * (kBwdCols * 128), c * 64, head, bi, c0);
        }
        if (!kDQ) {
          const int64_t base = (static_cast<int64_t>(bi) * p.hq + head) * p.sq + c0;
          bulk_load_1d(sStat + ys * 128, p.lse + base, 256, &y_full[ys]);
          bulk_load_1d(sStat + ys * 128 + 64, p.delta + base, 256, &y_full[ys]);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================================================ MMA issuer
    if (lane == 0 && n_iter > 0) {
      const uint32_t idesc_t = make_idesc_f16(p.is_bf16 != 0, kBwdRows, kBwdCols, false, false);
      const uint32_t idesc_a = make_idesc_f16(p.is_bf16 != 0, kBwdRows, D, false, true);
      auto issue_T = [&](int it) {
        const int ys = it % kS, tb = it & 1;
        const uint32_t x1 = smem_u32(sX1), x2 = smem_u32(sX2);
        const uint32_t y1 = smem_u32(sY + ys * 2 * L::kY), y2 = y1 + L::kY;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t xo = (k >> 2) * (kBwdRows * 128) + (k & 3) * 32;
          const uint32_t yo = (k >> 2) * (kBwdCols * 128) + (k & 3) * 32;
          umma_f16_ss(tmem_T1 + tb * 64, make_smem_desc_sw128(x1 + xo, 1024, 0), make_smem_desc_sw128(y1 + yo, 1024, 0),
                      idesc_t, k != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t xo = (k >> 2) * (kBwdRows * 128) + (k & 3) * 32;
          const uint32_t yo = (k >> 2) * (kBwdCols * 128) + (k & 3) * 32;
          umma_f16_ss(tmem_T2 + tb * 64, make_smem_desc_sw128(x2 + xo, 1024, 0), make_smem_desc_sw128(y2 + yo, 1024, 0),
                      idesc_t, k != 0 ? 1u : 0u);
        }
        umma_commit(&t_full[tb]);
      };
      mbar_wait(x_full, 0);
      mbar_wait(&y_full[0], 0);
      tc_fence_after();
      issue_T(0);
      for (int it = 0; it < n_iter; ++it) {
        const int ys = it % kS, tb = it & 1;
        if (it + 1 < n_iter) {
          mbar_wait(&y_full[(it + 1) % kS], ((it + 1) / kS) & 1);
          tc_fence_after();
          issue_T(it + 1);     // overwrites T[(it+1)&1]: its P/dS were consumed by the accumulate MMAs of it-1
        }
        mbar_wait(&pds_ready[tb], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t y1 = smem_u32(sY + ys * 2 * L::kY), y2 = y1 + L::kY;
#pragma unroll
        for (int k = 0; k < kBwdCols / 16; ++k) {
          // streamed tile as MN-major B (N = d contiguous, K = its 64 rows); A = P / dS from TMEM
          // (row = lane, 16 k-values = 8 packed 32-bit columns per step)
          const uint64_t bdesc1 = make_smem_desc_sw128(y1 + k * 16 * 128, 1024, kBwdCols * 128);
          if (kDQ) {
            umma_f16_ts(tmem_A1, tmem_T2 + tb * 64 + k * 8, bdesc1, idesc_a, (it | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t bdesc2 = make_smem_desc_sw128(y2 + k * 16 * 128, 1024, kBwdCols * 128);
            umma_f16_ts(tmem_A1, tmem_T1 + tb * 64 + k * 8, bdesc2, idesc_a, (it | k) != 0 ? 1u : 0u);
            umma_f16_ts(tmem_A2, tmem_T2 + tb * 64 + k * 8, bdesc1, idesc_a, (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&y_empty[ys]);
      }
      umma_commit(acc_done);
    }
  } else if (warp_idx >= 4) {
    // ============================================================ P / dS math + epilogue
    const int quad = warp_idx & 3;
    const int r = quad * 32 + lane;
    const int grow = r0 + r;                       // global row (key for dK/dV, query for dQ)
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const bool bf = p.is_bf16 != 0;
    const float sc = p.scale_log2;
    float row_lse2 = 0.f, row_delta = 0.f;         // dQ mode: statistics of this thread's query row
    if (kDQ) {
      const int64_t idx = (static_cast<int64_t>(bi) * p.hq + blockIdx.y) * p.sq + min(grow, p.sq - 1);
      row_lse2 = p.lse[idx];                       // already in the exp2 domain (-inf mapped to +inf)
      row_delta = p.delta[idx];
    }
    // one streamed block: P = exp2(T1 * scale - lse2), dS = P * (T2 - delta), written back to TMEM
    auto block = [&](int it, auto masked_tag) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      const int tb = it & 1, ys = it % kS;
      const int c0 = (col_blk0 + it % n_col) * kBwdCols;
      const float4* s_lse4 = reinterpret_cast<const float4*>(sStat + ys * 128);
      const float4* s_delta4 = s_lse4 + 16;
      const uint32_t a1 = tmem_T1 + tb * 64 + lane_addr, a2 = tmem_T2 + tb * 64 + lane_addr;
      uint32_t t1[2][32], t2[2][32];
      tmem_ld_32x32(a1, t1[0]);
      tmem_ld_32x32(a2, t2[0]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_wait();
        if (c == 0) {
          tmem_ld_32x32(a1 + 32, t1[1]);
          tmem_ld_32x32(a2 + 32, t2[1]);
        }
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float l2[4], dl[4];
          if (kDQ) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { l2[e] = row_lse2; dl[e] = row_delta; }
          } else {
            const float4 a = s_lse4[c * 8 + i4], d4 = s_delta4[c * 8 + i4];
            l2[0] = a.x; l2[1] = a.y; l2[2] = a.z; l2[3] = a.w;
            dl[0] = d4.x; dl[1] = d4.y; dl[2] = d4.z; dl[3] = d4.w;
          }
          float pv[4], dv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i4 * 4 + e;
            float pe = fast_exp2(fmaf(__uint_as_float(t1[c][i]), sc, -l2[e]));
            if (kMasked) {
              const int col = c0 + c * 32 + i;     // global column (query for dK/dV, key for dQ)
              const int key = kDQ ? col : grow;
              const int qry = kDQ ? grow : col;
              if ((p.causal && key > qry) || key >= p.sk || qry >= p.sq) pe = 0.f;
            }
            pv[e] = pe;
            dv[e] = pe * (__uint_as_float(t2[c][i]) - dl[e]);
          }
          pk[i4 * 2] = pack16(pv[0], pv[1], bf);
          pk[i4 * 2 + 1] = pack16(pv[2], pv[3], bf);
          dk[i4 * 2] = pack16(dv[0], dv[1], bf);
          dk[i4 * 2 + 1] = pack16(dv[2], dv[3], bf);
        }
        // packed chunk c -> columns [16c, 16c+16) of the same accumulator buffer; only fp32 columns that
        // are already in registers are overwritten (chunk 1 was loaded before chunk 0 is stored)
        tmem_st_32x16(a2 + c * 16, dk);
        if (!kDQ) tmem_st_32x16(a1 + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pds_ready[tb]);
    };
    for (int it = 0; it < n_iter; ++it) {
      const int tb = it & 1;
      const int c0 = (col_blk0 + it % n_col) * kBwdCols;
      mbar_wait(&t_full[tb], (it >> 1) & 1);
      tc_fence_after();
      // masking: key index > query index is masked (causal); also ragged tails
      bool need_mask;
      if (kDQ) need_mask = (p.causal && c0 + kBwdCols - 1 > r0) || (c0 + kBwdCols > p.sk) || (r0 + kBwdRows > p.sq);
      else need_mask = (p.causal && c0 < r0 + kBwdRows - 1) || (c0 + kBwdCols > p.sq) || (r0 + kBwdRows > p.sk);
      if (need_mask) block(it, std::true_type{});
      else block(it, std::false_type{});
    }
    if (n_iter > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
    }
    const int limit = kDQ ? p.sq : p.sk;
    const bool valid = grow < limit;
    const int oh = blockIdx.y;
    auto store_acc = [&](uint32_t tacc, void* out, int64_t sb, int64_t sh, int64_t ss, float mul) {
      uint8_t* orow = reinterpret_cast<uint8_t*>(out) +
                      (static_cast<int64_t>(grow) * ss + static_cast<int64_t>(bi) * sb + static_cast<int64_t>(oh) * sh) * 2;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t v[32];
        if (n_iter > 0) {
          tmem_ld_32x32(tacc + c * 32 + lane_addr, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (valid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack16(__uint_as_float(v[u * 8 + 0]) * mul, __uint_as_float(v[u * 8 + 1]) * mul, bf);
            w.y = pack16(__uint_as_float(v[u * 8 + 2]) * mul, __uint_as_float(v[u * 8 + 3]) * mul, bf);
            w.z = pack16(__uint_as_float(v[u * 8 + 4]) * mul, __uint_as_float(v[u * 8 + 5]) * mul, bf);
            w.w = pack16(__uint_as_float(v[u * 8 + 6]) * mul, __uint_as_float(v[u * 8 + 7]) * mul, bf);
            *reinterpret_cast<uint4*>(orow + c * 64 + u * 16) = w;
          }
        }
      }
    };
    if (kDQ) {
      store_acc(tmem_A1, p.out1, p.o1_sb, p.o1_sh, p.o1_ss, p.scale);
    } else {
      store_acc(tmem_A1, p.out1, p.o1_sb, p.o1_sh, p.o1_ss, 1.f);
      store_acc(tmem_A2, p.out2, p.o2_sb, p.o2_sh, p.o2_ss, p.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  });
  return fn;
}
thread_local std::string g_attn_error;

// 4-D map over a strided [b, h, s, d] view: dims (d, h, b, s); box (64, 1, 1, rows)
bool make_tmap_bhsd(CUtensorMap* map, const void* base, bool is_bf16, int d, int h, int b, int s, int64_t sh,
                    int64_t sb, int64_t ss, int rows) {
  ensure_cuda_context();
  EncodeFn fn = encode_fn();
  if (!fn) { g_attn_error = "cuTensorMapEncodeTiled unavailable"; return false; }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(b),
                        static_cast<cuuint64_t>(s)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(sh) * 2, static_cast<cuuint64_t>(sb) * 2,
                           static_cast<cuuint64_t>(ss) * 2};
  cuuint32_t box[4] = {64, 1, 1, static_cast<cuuint32_t>(rows)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_attn_error = "cuTensorMapEncodeTiled(4d) failed: code " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

bool strides_ok(int64_t sb, int64_t sh, int64_t ss, const void* ptr) {
  return (sb % 8 == 0) && (sh % 8 == 0) && (ss % 8 == 0) && (reinterpret_cast<uintptr_t>(ptr) % 16 == 0);
}

template <int D, bool kPTmem>
int launch_fwd(const AttnArgs& a, cudaStream_t s) {
  using L = FwdSmem<D, kPTmem>;
  CUtensorMap tq{}, tk{}, tv{};
  if (!make_tmap_bhsd(&tq, a.q, a.is_bf16, D, a.hq, a.b, a.sq, a.q_sh, a.q_sb, a.q_ss, kBlockQ)) return -5;
  if (!make_tmap_bhsd(&tk, a.k, a.is_bf16, D, a.hkv, a.b, a.sk, a.k_sh, a.k_sb, a.k_ss, kBlockKV)) return -5;
  if (!make_tmap_bhsd(&tv, a.v, a.is_bf16, D, a.hkv, a.b, a.sk, a.v_sh, a.v_sb, a.v_ss, kBlockKV)) return -5;
  auto kern = attn_fwd_kernel<D, kPTmem>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) { g_attn_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -10; }
    attr_set = true;
  }
  FwdParams p{};
  p.lse = a.lse; p.o = a.o; p.o_sb = a.o_sb; p.o_sh = a.o_sh; p.o_ss = a.o_ss;
  p.b = a.b; p.hq = a.hq; p.hkv = a.hkv; p.sq = a.sq; p.sk = a.sk;
  p.scale_log2 = a.scale * kLog2e; p.causal = a.causal ? 1 : 0; p.is_bf16 = a.is_bf16 ? 1 : 0;
  dim3 grid((a.sq + 2 * kBlockQ - 1) / (2 * kBlockQ), a.hq, a.b);
  kern<<<grid, 384, L::kTotal, s>>>(tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_attn_error = std::string("attn_fwd launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}


template <int D, bool kDQ>
int launch_bwd_one(const AttnArgs& a, cudaStream_t s) {
  using L = BwdSmem<D, kDQ>;
  CUtensorMap x1{}, x2{}, y1{}, y2{};
  bool ok = true;
  if (kDQ) {
    ok = ok && make_tmap_bhsd(&x1, a.q, a.is_bf16, D, a.hq, a.b, a.sq, a.q_sh, a.q_sb, a.q_ss, kBwdRows);
    ok = ok && make_tmap_bhsd(&x2, a.d_o, a.is_bf16, D, a.hq, a.b, a.sq, a.do_sh, a.do_sb, a.do_ss, kBwdRows);
    ok = ok && make_tmap_bhsd(&y1, a.k, a.is_bf16, D, a.hkv, a.b, a.sk, a.k_sh, a.k_sb, a.k_ss, kBwdCols);
    ok = ok && make_tmap_bhsd(&y2, a.v, a.is_bf16, D, a.hkv, a.b, a.sk, a.v_sh, a.v_sb, a.v_ss, kBwdCols);
  } else {
    ok = ok && make_tmap_bhsd(&x1, a.k, a.is_bf16, D, a.hkv, a.b, a.sk, a.k_sh, a.k_sb, a.k_ss, kBwdRows);
    ok = ok && make_tmap_bhsd(&x2, a.v, a.is_bf16, D, a.hkv, a.b, a.sk, a.v_sh, a.v_sb, a.v_ss, kBwdRows);
    ok = ok && make_tmap_bhsd(&y1, a.q, a.is_bf16, D, a.hq, a.b, a.sq, a.q_sh, a.q_sb, a.q_ss, kBwdCols);
    ok = ok && make_tmap_bhsd(&y2, a.d_o, a.is_bf16, D, a.hq, a.b, a.sq, a.do_sh, a.do_sb, a.do_ss, kBwdCols);
  }
  if (!ok) return -5;
  auto kern = attn_bwd_kernel<D, kDQ>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) { g_attn_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -10; }
    attr_set = true;
  }
  BwdParams p{};
  p.lse = a.delta + static_cast<int64_t>(a.b) * a.hq * a.sq;   // exp2-domain LSE written by attn_delta_kernel
  p.delta = a.delta;
  if (kDQ) { p.out1 = a.dq; p.o1_sb = a.dq_sb; p.o1_sh = a.dq_sh; p.o1_ss = a.dq_ss; }
  else {
    p.out1 = a.dv; p.o1_sb = a.dv_sb; p.o1_sh = a.dv_sh; p.o1_ss = a.dv_ss;
    p.out2 = a.dk; p.o2_sb = a.dk_sb; p.o2_sh = a.dk_sh; p.o2_ss = a.dk_ss;
  }
  p.b = a.b; p.hq = a.hq; p.hkv = a.hkv; p.sq = a.sq; p.sk = a.sk;
  p.scale_log2 = a.scale * kLog2e; p.scale = a.scale; p.causal = a.causal ? 1 : 0; p.is_bf16 = a.is_bf16 ? 1 : 0;
  dim3 grid(kDQ ? (a.sq + kBwdRows - 1) / kBwdRows : (a.sk + kBwdRows - 1) / kBwdRows, kDQ ? a.hq : a.hkv, a.b);
  kern<<<grid, 256, L::kTotal, s>>>(x1, x2, y1, y2, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_attn_error = std::string("attn_bwd launch: ") + cudaGetErrorString(e); return -11; }
  return 0;
}

template <int D>
int launch_bwd(const AttnArgs& a, cudaStream_t s) {
  const int64_t rows = static_cast<int64_t>(a.b) * a.hq * a.sq;
  attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(
      a.o, a.d_o, a.delta, a.lse, a.delta + rows, a.b, a.hq, a.sq, D, a.o_sb, a.o_sh, a.o_ss, a.do_sb, a.do_sh, a.do_ss, a.is_bf16 ? 1 : 0);
  int rc = launch_bwd_one<D, false>(a, s);
  if (rc != 0) return rc;
  return launch_bwd_one<D, true>(a, s);
}

}  // namespace

const char* attn_last_error() { return g_attn_error.c_str(); }

int attn_fwd(const AttnArgs& a, cudaStream_t s) {
  if (a.d != 64 && a.d != 128) { g_attn_error = "head dim must be 64 or 128"; return -1; }
  if (a.hq % a.hkv != 0) { g_attn_error = "hq must be a multiple of hkv"; return -1; }
  if (a.causal && a.sq != a.sk) { g_attn_error = "causal attention needs sq == sk"; return -2; }
  if (a.probs_out) { g_attn_error = "probability export uses the materialising path"; return -3; }
  if (!strides_ok(a.q_sb, a.q_sh, a.q_ss, a.q) || !strides_ok(a.k_sb, a.k_sh, a.k_ss, a.k) ||
      !strides_ok(a.v_sb, a.v_sh, a.v_ss, a.v) || !strides_ok(a.o_sb, a.o_sh, a.o_ss, a.o)) {
    g_attn_error = "q/k/v/o need 16-byte aligned base and strides";
    return -4;
  }
  static const bool p_tmem = [] {
    const char* e = getenv("B200_ATTN_P_TMEM");
    return e == nullptr || e[0] != '0';
  }();
  if (p_tmem) return a.d == 128 ? launch_fwd<128, true>(a, s) : launch_fwd<64, true>(a, s);
  return a.d == 128 ? launch_fwd<128, false>(a, s) : launch_fwd<64, false>(a, s);
}

int attn_bwd(const AttnArgs& a, cudaStream_t s) {
  if (a.d != 64 && a.d != 128) { g_attn_error = "head dim must be 64 or 128"; return -1; }
  if (a.hq % a.hkv != 0) { g_attn_error = "hq must be a multiple of hkv"; return -1; }
  if (a.causal && a.sq != a.sk) { g_attn_error = "causal attention needs sq == sk"; return -2; }
  if (a.sq % kBwdCols != 0 || a.sk % kBwdCols != 0) { g_attn_error = "backward needs sq, sk multiples of 64"; return -3; }
  if (!strides_ok(a.q_sb, a.q_sh, a.q_ss, a.q) || !strides_ok(a.k_sb, a.k_sh, a.k_ss, a.k) ||
      !strides_ok(a.v_sb, a.v_sh, a.v_ss, a.v) || !strides_ok(a.o_sb, a.o_sh, a.o_ss, a.o) ||
      !strides_ok(a.do_sb, a.do_sh, a.do_ss, a.d_o) || !strides_ok(a.dq_sb, a.dq_sh, a.dq_ss, a.dq) ||
      !strides_ok(a.dk_sb, a.dk_sh, a.dk_ss, a.dk) || !strides_ok(a.dv_sb, a.dv_sh, a.dv_ss, a.dv)) {
    g_attn_error = "attention operands need 16-byte aligned base and strides";
    return -4;
  }
  return a.d == 128 ? launch_bwd<128>(a, s) : launch_bwd<64>(a, s);
}

}  // namespace b200
