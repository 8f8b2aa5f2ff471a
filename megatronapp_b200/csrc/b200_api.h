// Host-callable launchers for every sm_100a kernel in the framework.
// Plain C++ (no torch headers) so kernel TUs compile fast; bindings.cpp adapts tensors.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace b200 {

// cuTensorMapEncodeTiled is a driver-API call and needs a current context on the calling thread.  The
// autograd worker thread may reach a launcher before any runtime call bound the primary context
// (notably under CUDA-graph capture, where allocations come from a pre-built pool): bind it once.
inline void ensure_cuda_context() {
  thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(dev);
    bound = true;
  }
}

constexpr int kMaxPeers = 8;

enum GemmEpilogue : int {
  kEpiStore16 = 0,    // D (bf16/fp16) = act(acc + bias)           (TMA store)
  kEpiAccumF32 = 1,   // D32 += acc                                (wgrad into fp32 main_grad)
  kEpiStoreF32 = 2,   // D32  = acc
  kEpiReduceScatter = 3,  // GEMM -> reduce-scatter over peer memory (row-parallel linear)
};

enum GemmComm : int {
  kCommNone = 0,
  kCommAllGather = 1,      // all-gather(A rows) fused in front of the GEMM (column-parallel linear)
  kCommReduceScatter = 2,  // reduce-scatter(D rows) fused behind the GEMM
  kCommAllGatherK = 3,     // all-gather along the reduction (token) axis of B fused into the wgrad GEMM (2-CTA kernel only)
};

enum GemmGroupMode : int {
  kGroupNone = 0,
  kGroupM = 1,   // rows of A / D belong to groups (MoE forward + dgrad); B and bias are stacked per group
  kGroupK = 2,   // the reduction dimension is split by group (MoE wgrad); D is stacked per group
};

struct GemmArgs {
  // D[M,N] = A[M,K] * B[N,K]^T  (logical).  *_mn_major: operand stored transposed
  // (A stored [K,M] / B stored [K,N], leading dimension = lda/ldb elements).
  const void* A = nullptr;
  const void* B = nullptr;
  void* D = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0, ldd = 0;
  bool a_mn_major = false;
  bool b_mn_major = false;
  bool is_bf16 = true;
  int epilogue = kEpiStore16;
  const void* bias = nullptr;  // optional [N], same 16-bit dtype as D
  int act = 0;                 // 0 none, 1 tanh-GELU
  void* D_pre = nullptr;       // optional pre-activation output (same shape/ld as D)
  int block_n = 0;             // 0 = auto (128 or 256)
  int num_sms = 0;             // 0 = all
  // ---- fused collective (kCommAllGather / kCommReduceScatter) ----------------
  int comm = kCommNone;
  int tp_rank = 0, tp_size = 1;
  int n_comm_ctas = 0;
  uint32_t epoch = 0;                 // monotonically increasing per arena slot
  // all-gather: A is this rank's gather buffer [M,K] in the symmetric arena,
  // A_local is the local shard [M/tp,K]; peers' gather buffers / flag arrays below.
  const void* A_local = nullptr;
  int lda_local = 0;
  void* peer_buf[kMaxPeers] = {nullptr};       // AG: peers' gather buffers; RS: peers' staging buffers
  uint32_t* peer_flags[kMaxPeers] = {nullptr};  // per-peer flag arrays (index tp_rank = local)
  // reduce-scatter: D is the local output shard [M/tp, N]; staging in peer_buf[tp_rank]
  // ---- grouped GEMM (cta_group::2 kernel only) -------------------------------------------
  // group_offsets: device int32 [n_groups + 1], multiples of 256 (the padded ragged row layout written by the MoE
  // permute kernel).  kGroupM: M = row capacity, B is the stacked weight [n_groups * group_stride, ...].
  // kGroupK: K = row capacity, D is the stacked fp32 gradient [n_groups * group_stride, N].
  int group_mode = kGroupNone;
  const int32_t* group_offsets = nullptr;
  int n_groups = 0;
  int group_stride = 0;
  const float* row_scale = nullptr;   // kGroupM: optional per-row multiplier (router probability)
  cudaStream_t stream = nullptr;
};

// returns 0 on success, otherwise a negative error code (bad shape / alignment)
int gemm_bf16(const GemmArgs& args);
const char* gemm_last_error();
int gemm_flags_needed(int M, int N, int tp_size);
// cta_group::2 variant (gemm2_sm100.cu); gemm_bf16 dispatches to it when profitable
bool gemm2_supported(const GemmArgs& args);
int gemm2_bf16(const GemmArgs& args);
const char* gemm2_last_error();

// ---------------------------------------------------------------- fp8 (gemm_fp8_sm100.cu)
struct Fp8GemmArgs {
  // D[M,N] = scale_a_inv * scale_b_inv * (A8[M,K] . B8[N,K]^T) (+ bias); operands one byte / element
  const void* A = nullptr;
  const void* B = nullptr;
  void* D = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0, ldd = 0;
  bool a_e5m2 = false, b_e5m2 = false;
  const float* scale_a_inv = nullptr;   // device scalars
  const float* scale_b_inv = nullptr;
  const void* bias = nullptr;           // optional 16-bit [N] (store epilogue only)
  bool out_bf16 = true;                 // 16-bit output dtype (store epilogue)
  bool accum_f32 = false;               // D is fp32 and receives "+=" (wgrad into main_grad)
  cudaStream_t stream = nullptr;
};
int gemm_fp8(const Fp8GemmArgs& a);
const char* gemm_fp8_last_error();
void fp8_amax(const void* x, int64_t n, bool is_bf16, float* amax, cudaStream_t s);
void fp8_scale_from_amax(const float* amax, bool e5m2, int margin, float* scale, float* scale_inv, cudaStream_t s);
// q [rows, cols] and (optionally) qt [cols, rows] fp8 copies of a 16-bit matrix, x * scale
void fp8_cast_transpose(const void* x, const float* scale, void* q, void* qt, int rows, int cols, bool is_bf16,
                        bool e5m2, cudaStream_t s);

// ---------------------------------------------------------------- mxfp8 (gemm_mxfp8_sm100.cu)
struct MxGemmArgs {
  // D[M,N] = (A8 * 2^(sfa-127)) . (B8 * 2^(sfb-127))^T (+ bias); K % 128 == 0; sfa in A-role, sfb in B-role chunk layout
  const void* A = nullptr;
  const void* B = nullptr;
  void* D = nullptr;
  const void* sfa = nullptr;
  const void* sfb = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0, ldd = 0;
  bool a_e5m2 = false, b_e5m2 = false;
  const void* bias = nullptr;
  bool out_bf16 = true;
  bool accum_f32 = false;
  cudaStream_t stream = nullptr;
};
int gemm_mxfp8(const MxGemmArgs& a);
const char* gemm_mxfp8_last_error();
int64_t mx_sf_bytes(int role, int rows, int K);     // role 0: A operand (128-row chunks), 1: B operand (192-row tiles)
// x [R, C] 16-bit -> row-wise (blocks along C) and / or column-wise (x^T, blocks along R) fp8 + E8M0 scales
void mx_quantize(const void* x, int R, int C, bool is_bf16, bool e5m2, void* q_row, void* sf_row, int role_row, void* q_col,
                 void* sf_col, int role_col, cudaStream_t s);

// ---------------------------------------------------------------- elementwise / norm
void layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                   float* rstd, int rows, int cols, float eps, bool rms, bool zero_centered,
                   bool is_bf16, cudaStream_t s);
void layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                   const float* rstd, void* dx, float* dgamma_partial, float* dbeta_partial,
                   int rows, int cols, bool rms, bool zero_centered, bool is_bf16, int n_partial,
                   cudaStream_t s);
void layernorm_bwd_finalize(const float* dgamma_partial, const float* dbeta_partial,
                            float* dgamma, float* dbeta, int n_partial, int cols, bool accumulate,
                            cudaStream_t s);
int layernorm_bwd_partials(int rows);

void bias_gelu_fwd(const void* x, const void* bias, void* y, int64_t rows, int cols, bool is_bf16,
                   cudaStream_t s);
void bias_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, int64_t rows,
                   int cols, bool is_bf16, cudaStream_t s);
void swiglu_fwd(const void* x, const void* bias, void* y, int64_t rows, int cols_out, bool is_bf16,
                cudaStream_t s);
void swiglu_bwd(const void* dy, const void* x, const void* bias, void* dx, int64_t rows,
                int cols_out, bool is_bf16, cudaStream_t s);
void bias_dropout_add(const void* x, const void* bias, const void* residual, void* out,
                      uint8_t* mask_out, int64_t rows, int cols, float p, uint64_t seed,
                      uint64_t offset, bool is_bf16, cudaStream_t s);
void dropout_bwd(const void* dy, const uint8_t* mask, void* dx, int64_t n, float p, bool is_bf16,
                 cudaStream_t s);
void rope_apply(const void* x, const float* cos_sin, void* y, int s, int b, int h, int d,
                int rot_dim, bool interleaved, bool inverse, bool is_bf16, cudaStream_t st);
void bias_grad_reduce(const void* dy, float* dbias, int64_t rows, int cols, bool is_bf16,
                      bool accumulate, cudaStream_t s);

// vocab-parallel cross entropy pieces (fp32 statistics; logits 16-bit or fp32)
void ce_local_stats(const void* logits, const int64_t* labels, float* row_max, float* row_sumexp,
                    float* tgt_logit, int64_t rows, int vocab_local, int vocab_start, int ldl,
                    int logits_dtype, cudaStream_t s);
void ce_rescale_sumexp(float* row_sumexp, const float* local_max, const float* global_max,
                       int64_t rows, cudaStream_t s);
void ce_grad(const void* logits, const int64_t* labels, const float* global_max,
             const float* global_sumexp, const float* grad_loss, void* dlogits, int64_t rows,
             int vocab_local, int vocab_start, int ldl, int logits_dtype, float label_smoothing,
             int64_t vocab_total, cudaStream_t s);

// optimizer / multi-tensor
struct AdamArgs {
  float* master;       // fp32 master params (flat)
  float* exp_avg;
  float* exp_avg_sq;
  const float* grad;   // fp32 grads (flat), already scaled by grad_scale inside the kernel
  void* param16;       // optional 16-bit model params (flat), may be nullptr
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, grad_scale;
  float bias_correction1, bias_correction2;
  bool is_bf16;
  const float* found_inf;  // optional device flag: skip update when != 0
};
void adam_step(const AdamArgs& a, cudaStream_t s);
void l2norm_sq(const void* x, int64_t n, int dtype, float* out_accum, cudaStream_t s);
void scale_inplace(void* x, int64_t n, int dtype, float scale, const float* scale_dev,
                   cudaStream_t s);
void cast_copy(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, float scale,
               cudaStream_t s);
void add_into_f32(float* dst, const void* src, int src_dtype, int64_t n, cudaStream_t s);
void check_finite(const void* x, int64_t n, int dtype, float* found_inf, cudaStream_t s);
void chunk_mean(const void* x, float* out, int64_t rows, int cols, int pixels, bool is_bf16,
                cudaStream_t s);

// flash attention (tcgen05).  Every operand is a strided logical [b, h, s, d] view with a
// contiguous last dim (strides in elements), d in {64, 128}; lse / delta are [b, h, s] fp32.
struct AttnArgs {
  const void* q; const void* k; const void* v; void* o; float* lse;
  const void* d_o; void* dq; void* dk; void* dv; float* delta; float* dq_accum;
  int b, hq, hkv, sq, sk, d;
  float scale;
  bool causal;
  bool is_bf16;
  void* probs_out;  // optional MegaScope export [b,hq,sq,sk] (fwd only)
  int64_t q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss, o_sb, o_sh, o_ss;
  int64_t do_sb, do_sh, do_ss, dq_sb, dq_sh, dq_ss, dk_sb, dk_sh, dk_ss, dv_sb, dv_sh, dv_ss;
};
int attn_fwd(const AttnArgs& a, cudaStream_t s);
int attn_bwd(const AttnArgs& a, cudaStream_t s);
const char* attn_last_error();

// ---------------------------------------------------------------- comm
void push_flags_barrier(uint32_t* const* peer_flags, int rank, int size, int slot, uint32_t epoch,
                        cudaStream_t s);
struct RsArgs {  // fused scale + cast + reduce-scatter of a gradient bucket
  const float* src;            // local fp32 bucket [n]
  void* peer_stage[kMaxPeers]; // staging [size][n/size] in each peer's arena (16-bit or fp32)
  uint32_t* peer_flags[kMaxPeers];
  float* dst;                  // local fp32 shard [n/size]
  int64_t n; int rank; int size; float scale; uint32_t epoch; int wire_dtype; int n_ctas;
};
void bucket_reduce_scatter(const RsArgs& a, cudaStream_t s);
struct AgArgs {  // param all-gather with fused fp32->16-bit cast
  const float* src_shard; void* peer_dst[kMaxPeers]; uint32_t* peer_flags[kMaxPeers];
  int64_t shard_n; int rank; int size; uint32_t epoch; int dst_dtype; int n_ctas;
};
void shard_all_gather_cast(const AgArgs& a, cudaStream_t s);
struct A2aArgs {  // MoE dispatch: scatter routed rows into expert-owner buffers
  const void* src; const int32_t* dst_rank; const int32_t* dst_row; void* peer_buf[kMaxPeers];
  uint32_t* peer_flags[kMaxPeers]; int64_t rows; int cols; int rank; int size; uint32_t epoch;
  bool is_bf16; int n_ctas;
};
void a2a_scatter_rows(const A2aArgs& a, cudaStream_t s);
struct EmbedArgs {  // K14: vocab-parallel embedding + sequence-parallel scatter
  const int64_t* tokens; const void* weight; void* peer_out[kMaxPeers]; uint32_t* peer_flags[kMaxPeers];
  int b; int s; int h; int64_t vocab_start; int64_t vocab_end; int rank; int size; uint32_t epoch; int n_ctas;
};
void embed_scatter(const EmbedArgs& a, cudaStream_t s);
void wait_flags(const uint32_t* flags, int n, uint32_t target, cudaStream_t s);
void spin_kernel(int64_t cycles, cudaStream_t s);  // MegaScan software slow-rank injector
void write_flag(uint32_t* flag, uint32_t value, cudaStream_t s);

// ---------------------------------------------------------------- paged-KV attention (paged_attn.cu)
struct PagedAttnArgs {
  const void* q;            // [T, ng * group, d]
  const void* k_cache;      // [n_slots, ng, d]   slot = block * block_size + offset
  const void* v_cache;
  void* out;                // [T, ng * group, d]
  const int32_t* block_table;  // [R, max_blocks]
  const int32_t* tok_req;   // [T] row of block_table
  const int32_t* tok_len;   // [T] keys visible to the token (its position + 1)
  float* ws_o;              // [T * heads * splits, d]   (splits > 1)
  float* ws_ml;             // [T * heads * splits, 2]
  int n_tokens, ng, group, head_dim, block_size, max_blocks;
  float scale_log2;
  bool is_bf16;
};
bool paged_attention_supported(int head_dim, int group);
void paged_attention(const PagedAttnArgs& a, int splits, cudaStream_t s);

// ---------------------------------------------------------------- MoE (moe.cu)
void moe_router_topk_fwd(const float* logits, float* probs, uint8_t* map, int32_t* tokens_per_expert, int n, int E, int k,
                         bool pre_softmax, float scaling, cudaStream_t s);
void moe_router_topk_bwd(const float* logits, const float* probs, const uint8_t* map, const float* dprobs, float* dlogits,
                         int n, int E, bool pre_softmax, float scaling, cudaStream_t s);
// ragged layout: counts[E_local], offsets[E_local+1] (segments padded to `pad` rows), row_slot[cap] = token*E_total+expert
// (or -1), row_prob[cap], dest_row[n*E_local] = row (or -1); pos_scratch[n*E_local]
void moe_build_layout(const uint8_t* map, const float* probs, int n, int E_total, int e_begin, int E_local, int pad,
                      int32_t* pos_scratch, int32_t* counts, int32_t* offsets, int32_t* row_slot, float* row_prob,
                      int32_t* dest_row, int cap, cudaStream_t s);
void moe_gather_rows(const void* in, const int32_t* row_slot, int slot_div, const float* row_scale, const int32_t* offsets,
                     int n_groups, int cap, int cols, void* out, const void* other, float* dot_out, bool is_bf16, cudaStream_t s);
void moe_gather_sum(const void* y, const int32_t* dest_row, int n, int E_local, int cols, void* out, bool is_bf16, cudaStream_t s);

// expert parallelism over the EP symmetric arena (pointer tables are indexed by EP rank)
void moe_ep_dispatch(const void* x, const int32_t* send_offsets, const int32_t* send_row, const int32_t* counts_send, const float* probs,
                     void* const* stage_x, int32_t* const* stage_idx, float* const* stage_prob, int32_t* const* recv_counts,
                     uint32_t* const* flags, int n, int E_total, int E_local, int me, int ep, int Rs, int cols, uint32_t epoch,
                     int32_t* done_counter, cudaStream_t s);
void moe_ep_recv_layout(const uint32_t* flags, const int32_t* recv_counts, int ep, int E_local, int Rs, int pad, int cap,
                        uint32_t epoch, int32_t* offsets, int32_t* counts, int32_t* row_stage, int32_t* overflow, cudaStream_t s);
void moe_ep_combine_push(const void* z, const int32_t* row_stage, const int32_t* row_idx, const float* side, const int32_t* offsets,
                         int n_groups, int cap, int Rs, int cols, void* const* ret_x, float* const* ret_f, uint32_t* const* flags,
                         int me, int ep, uint32_t epoch, int32_t* done_counter, cudaStream_t s);
void moe_ep_wait(const uint32_t* flags, int ep, uint32_t epoch, cudaStream_t s);
void moe_scatter_dprobs(const float* ret_f, const int32_t* send_row, float* dprobs, int64_t pairs, cudaStream_t s);

// MegaDPP mailbox (mailbox.cpp)
int mailbox_create(int n_slots, size_t slot_bytes, const char* shm_name, void* ipc_handle_out64);
int mailbox_connect(int id, const void* peer_ipc_handle64, const char* peer_shm_name);
int mailbox_put(int id, int slot, const void* src, size_t nbytes, uint32_t epoch, cudaStream_t s);
bool mailbox_try_get(int id, int slot, uint32_t epoch);
void* mailbox_slot_ptr(int id, int slot);
long mailbox_sent(int id);
void mailbox_destroy(int id);
const char* mailbox_last_error();

}  // namespace b200
