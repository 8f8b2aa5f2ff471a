// PyTorch bindings for the sm_100a kernels (pybind11 module megatronapp_b200._C).
#include <torch/extension.h>
#include <mutex>
#include <vector>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "b200_api.h"

namespace py = pybind11;
using torch::Tensor;

namespace {

std::atomic<int64_t> g_launches{0};

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
inline void count(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int dtype_code(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kBFloat16: return 0;
    case at::kHalf: return 1;
    case at::kFloat: return 2;
    default: TORCH_CHECK(false, "unsupported dtype ", t.scalar_type());
  }
}
bool is_bf16(const Tensor& t) {
  TORCH_CHECK(t.scalar_type() == at::kBFloat16 || t.scalar_type() == at::kHalf, "expected bf16/fp16, got ", t.scalar_type());
  return t.scalar_type() == at::kBFloat16;
}
const void* optptr(const c10::optional<Tensor>& t) { return t.has_value() && t->defined() ? t->data_ptr() : nullptr; }

void check_2d(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.dim() == 2, name, " must be 2-D");
  TORCH_CHECK(t.stride(1) == 1, name, " must be row-major (inner stride 1)");
}

void fill_common(b200::GemmArgs& g, const Tensor& A, const Tensor& B, const Tensor& D, bool a_mn,
                 bool b_mn, int epilogue) {
  check_2d(A, "A"); check_2d(B, "B"); check_2d(D, "D");
  g.A = A.data_ptr(); g.B = B.data_ptr(); g.D = D.data_ptr();
  g.M = static_cast<int>(D.size(0)); g.N = static_cast<int>(D.size(1));
  g.K = static_cast<int>(a_mn ? A.size(0) : A.size(1));
  TORCH_CHECK((a_mn ? A.size(1) : A.size(0)) == g.M, "A shape does not match D rows");
  TORCH_CHECK((b_mn ? B.size(1) : B.size(0)) == g.N, "B shape does not match D cols");
  TORCH_CHECK((b_mn ? B.size(0) : B.size(1)) == g.K, "A/B reduction dims differ");
  g.lda = static_cast<int>(A.stride(0)); g.ldb = static_cast<int>(B.stride(0)); g.ldd = static_cast<int>(D.stride(0));
  g.a_mn_major = a_mn; g.b_mn_major = b_mn;
  g.is_bf16 = is_bf16(A);
  TORCH_CHECK(A.scalar_type() == B.scalar_type(), "A and B dtypes differ");
  g.epilogue = epilogue;
  if (epilogue == b200::kEpiStore16 || epilogue == b200::kEpiReduceScatter) {
    TORCH_CHECK(D.scalar_type() == A.scalar_type(), "D dtype must match inputs");
  } else {
    TORCH_CHECK(D.scalar_type() == at::kFloat, "fp32 epilogue needs an fp32 D");
  }
  g.stream = cur_stream();
}

void gemm(const Tensor& A, const Tensor& B, Tensor& D, bool a_mn, bool b_mn, int epilogue,
          const c10::optional<Tensor>& bias, int act, const c10::optional<Tensor>& D_pre, int block_n) {
  c10::cuda::CUDAGuard guard(A.device());
  b200::GemmArgs g;
  fill_common(g, A, B, D, a_mn, b_mn, epilogue);
  g.bias = optptr(bias); g.act = act;
  g.D_pre = const_cast<void*>(optptr(D_pre));
  g.block_n = block_n;
  int rc = b200::gemm_bf16(g);
  TORCH_CHECK(rc == 0, "b200 gemm failed (", rc, "): ", b200::gemm_last_error());
  count();
}

// Grouped GEMM over the 256-row-padded ragged layout written by moe_permute (MoE experts).
//   mode 1 (GroupM): D[cap, N] = A[cap, K] . W_g^T, W stacked [G * stride, .]; b_mn selects dgrad (W stored [K-rows, N]).
//   mode 2 (GroupK): D32[G * stride, N] += A[cap, M]^T . B[cap, N] restricted to the rows of each group.
void gemm_grouped(const Tensor& A, const Tensor& B, Tensor& D, const Tensor& offsets, int mode, bool b_mn,
                  const c10::optional<Tensor>& bias, int act, const c10::optional<Tensor>& D_pre,
                  const c10::optional<Tensor>& row_scale, int n_groups, int group_stride) {
  c10::cuda::CUDAGuard guard(A.device());
  check_2d(A, "A"); check_2d(B, "B"); check_2d(D, "D");
  TORCH_CHECK(offsets.is_cuda() && offsets.scalar_type() == at::kInt && offsets.numel() >= n_groups + 1, "offsets: int32 [G+1]");
  b200::GemmArgs g;
  g.A = A.data_ptr(); g.B = B.data_ptr(); g.D = D.data_ptr();
  g.lda = static_cast<int>(A.stride(0)); g.ldb = static_cast<int>(B.stride(0)); g.ldd = static_cast<int>(D.stride(0));
  g.is_bf16 = is_bf16(A);
  g.group_offsets = offsets.data_ptr<int32_t>(); g.n_groups = n_groups; g.group_stride = group_stride;
  g.bias = optptr(bias); g.act = act; g.D_pre = const_cast<void*>(optptr(D_pre));
  g.row_scale = row_scale.has_value() && row_scale->defined() ? row_scale->data_ptr<float>() : nullptr;
  if (mode == 1) {
    g.group_mode = b200::kGroupM; g.epilogue = b200::kEpiStore16;
    g.a_mn_major = false; g.b_mn_major = b_mn;
    g.M = static_cast<int>(A.size(0)); g.K = static_cast<int>(A.size(1)); g.N = static_cast<int>(D.size(1));
    TORCH_CHECK(D.size(0) == A.size(0) && A.size(0) % 256 == 0, "GroupM: row capacity must be a multiple of 256");
    TORCH_CHECK(B.size(0) == static_cast<int64_t>(n_groups) * group_stride, "GroupM: stacked weight rows != G * stride");
    TORCH_CHECK(b_mn ? (group_stride == g.K && B.size(1) == g.N) : (group_stride == g.N && B.size(1) == g.K),
                "GroupM: weight shape does not match A/D");
  } else {
    g.group_mode = b200::kGroupK; g.epilogue = b200::kEpiAccumF32;
    g.a_mn_major = true; g.b_mn_major = true;
    g.M = static_cast<int>(A.size(1)); g.K = static_cast<int>(A.size(0)); g.N = static_cast<int>(B.size(1));
    TORCH_CHECK(A.size(0) == B.size(0), "GroupK: A and B must have the same row capacity");
    TORCH_CHECK(D.scalar_type() == at::kFloat && D.size(0) == static_cast<int64_t>(n_groups) * group_stride && D.size(1) == g.N,
                "GroupK: D must be fp32 [G * stride, N]");
    TORCH_CHECK(group_stride == g.M, "GroupK: stride must equal the per-group output rows");
  }
  g.stream = cur_stream();
  int rc = b200::gemm_bf16(g);
  TORCH_CHECK(rc == 0, "b200 grouped gemm failed (", rc, "): ", b200::gemm_last_error());
  count();
}

void gemm_ag(const Tensor& A_local, int64_t gather_ptr, const Tensor& B, Tensor& D, bool b_mn,
             const c10::optional<Tensor>& bias, int act, std::vector<int64_t> peer_bufs,
             std::vector<int64_t> peer_flags, int rank, int size, int64_t epoch, int n_comm_ctas,
             int block_n, const c10::optional<Tensor>& D_pre) {
  c10::cuda::CUDAGuard guard(A_local.device());
  check_2d(A_local, "A_local"); check_2d(B, "B"); check_2d(D, "D");
  b200::GemmArgs g;
  g.A = reinterpret_cast<const void*>(gather_ptr);
  g.A_local = A_local.data_ptr(); g.lda_local = static_cast<int>(A_local.stride(0));
  g.B = B.data_ptr(); g.D = D.data_ptr();
  g.M = static_cast<int>(D.size(0)); g.N = static_cast<int>(D.size(1)); g.K = static_cast<int>(A_local.size(1));
  TORCH_CHECK(A_local.size(0) * size == g.M, "A_local rows * tp must equal D rows");
  g.lda = g.K; g.ldb = static_cast<int>(B.stride(0)); g.ldd = static_cast<int>(D.stride(0));
  g.a_mn_major = false; g.b_mn_major = b_mn; g.is_bf16 = is_bf16(A_local);
  g.epilogue = b200::kEpiStore16; g.bias = optptr(bias); g.act = act; g.block_n = block_n;
  g.D_pre = const_cast<void*>(optptr(D_pre));
  g.comm = b200::kCommAllGather; g.tp_rank = rank; g.tp_size = size; g.epoch = static_cast<uint32_t>(epoch);
  g.n_comm_ctas = n_comm_ctas;
  for (int i = 0; i < size; ++i) {
    g.peer_buf[i] = reinterpret_cast<void*>(peer_bufs[i]);
    g.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]);
  }
  g.stream = cur_stream();
  int rc = b200::gemm_bf16(g);
  TORCH_CHECK(rc == 0, "b200 gemm_ag failed (", rc, "): ", b200::gemm_last_error());
  count();
}

// column-parallel wgrad under sequence parallelism: D32[N_out, K_in] (+)= dY[tokens, N_out]^T . all_gather(X_local)[tokens, K_in]
void gemm_ag_wgrad(const Tensor& dY, const Tensor& X_local, int64_t gather_ptr, Tensor& D, bool accumulate,
                   std::vector<int64_t> peer_bufs, std::vector<int64_t> peer_flags, int rank, int size, int64_t epoch,
                   int n_comm_ctas) {
  c10::cuda::CUDAGuard guard(dY.device());
  check_2d(dY, "dY"); check_2d(X_local, "X_local"); check_2d(D, "D");
  TORCH_CHECK(D.scalar_type() == at::kFloat, "AG-wgrad accumulates into an fp32 gradient");
  b200::GemmArgs g;
  g.A = dY.data_ptr(); g.B = reinterpret_cast<const void*>(gather_ptr); g.D = D.data_ptr();
  g.A_local = X_local.data_ptr(); g.lda_local = static_cast<int>(X_local.stride(0));
  g.M = static_cast<int>(dY.size(1)); g.N = static_cast<int>(X_local.size(1)); g.K = static_cast<int>(dY.size(0));
  TORCH_CHECK(X_local.size(0) * size == g.K && D.size(0) == g.M && D.size(1) == g.N, "AG-wgrad shapes");
  g.lda = static_cast<int>(dY.stride(0)); g.ldb = g.N; g.ldd = static_cast<int>(D.stride(0));
  g.a_mn_major = true; g.b_mn_major = true; g.is_bf16 = is_bf16(dY);
  g.epilogue = accumulate ? b200::kEpiAccumF32 : b200::kEpiStoreF32;
  g.comm = b200::kCommAllGatherK; g.tp_rank = rank; g.tp_size = size; g.epoch = static_cast<uint32_t>(epoch);
  g.n_comm_ctas = n_comm_ctas;
  for (int i = 0; i < size; ++i) {
    g.peer_buf[i] = reinterpret_cast<void*>(peer_bufs[i]);
    g.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]);
  }
  g.stream = cur_stream();
  int rc = b200::gemm_bf16(g);
  TORCH_CHECK(rc == 0, "b200 gemm_ag_wgrad failed (", rc, "): ", b200::gemm_last_error());
  count();
}

void gemm_rs(const Tensor& A, const Tensor& B, Tensor& D_out, bool b_mn, const c10::optional<Tensor>& bias,
             std::vector<int64_t> peer_stage, std::vector<int64_t> peer_flags, int rank, int size,
             int64_t epoch, int block_n) {
  c10::cuda::CUDAGuard guard(A.device());
  check_2d(A, "A"); check_2d(B, "B"); check_2d(D_out, "D_out");
  b200::GemmArgs g;
  g.A = A.data_ptr(); g.B = B.data_ptr(); g.D = D_out.data_ptr();
  g.M = static_cast<int>(A.size(0)); g.K = static_cast<int>(A.size(1));
  g.N = static_cast<int>(D_out.size(1));
  TORCH_CHECK(D_out.size(0) * size == g.M, "D_out rows * tp must equal A rows");
  g.lda = static_cast<int>(A.stride(0)); g.ldb = static_cast<int>(B.stride(0)); g.ldd = static_cast<int>(D_out.stride(0));
  g.a_mn_major = false; g.b_mn_major = b_mn; g.is_bf16 = is_bf16(A);
  g.epilogue = b200::kEpiReduceScatter; g.bias = optptr(bias); g.block_n = block_n;
  g.comm = b200::kCommReduceScatter; g.tp_rank = rank; g.tp_size = size; g.epoch = static_cast<uint32_t>(epoch);
  for (int i = 0; i < size; ++i) {
    g.peer_buf[i] = reinterpret_cast<void*>(peer_stage[i]);
    g.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]);
  }
  g.stream = cur_stream();
  int rc = b200::gemm_bf16(g);
  TORCH_CHECK(rc == 0, "b200 gemm_rs failed (", rc, "): ", b200::gemm_last_error());
  count();
}

// ---------------------------------------------------------------- mxfp8
// -> [q_row, sf_row, q_col, sf_col] (undefined tensors for the directions that were not asked for)
std::vector<Tensor> mx_quantize(const Tensor& x, bool e5m2, int role_row, int role_col) {
  c10::cuda::CUDAGuard guard(x.device());
  check_2d(x, "x");
  TORCH_CHECK(x.is_contiguous(), "mx_quantize: contiguous input");
  const int R = static_cast<int>(x.size(0)), C = static_cast<int>(x.size(1));
  auto bo = x.options().dtype(at::kByte);
  Tensor q_row, sf_row, q_col, sf_col;
  if (role_row >= 0) {
    TORCH_CHECK(C % 128 == 0, "row-wise mxfp8 needs cols % 128 == 0");
    q_row = torch::empty({R, C}, bo);
    sf_row = torch::zeros({b200::mx_sf_bytes(role_row, R, C)}, bo);
  }
  if (role_col >= 0) {
    TORCH_CHECK(R % 128 == 0 && C % 8 == 0, "column-wise mxfp8 needs rows % 128 == 0");
    q_col = torch::empty({C, R}, bo);
    sf_col = torch::zeros({b200::mx_sf_bytes(role_col, C, R)}, bo);
  }
  b200::mx_quantize(x.data_ptr(), R, C, is_bf16(x), e5m2, role_row >= 0 ? q_row.data_ptr() : nullptr,
                    role_row >= 0 ? sf_row.data_ptr() : nullptr, role_row, role_col >= 0 ? q_col.data_ptr() : nullptr,
                    role_col >= 0 ? sf_col.data_ptr() : nullptr, role_col, cur_stream());
  count();
  return {q_row, sf_row, q_col, sf_col};
}

void gemm_mxfp8(const Tensor& A, const Tensor& sfa, const Tensor& B, const Tensor& sfb, Tensor& D, bool a_e5m2, bool b_e5m2,
                const c10::optional<Tensor>& bias, bool accumulate) {
  c10::cuda::CUDAGuard guard(A.device());
  check_2d(A, "A"); check_2d(B, "B"); check_2d(D, "D");
  TORCH_CHECK(A.scalar_type() == at::kByte && B.scalar_type() == at::kByte && A.size(1) == B.size(1), "mxfp8 operands: uint8 [rows, K]");
  b200::MxGemmArgs g;
  g.A = A.data_ptr(); g.B = B.data_ptr(); g.D = D.data_ptr(); g.sfa = sfa.data_ptr(); g.sfb = sfb.data_ptr();
  g.M = static_cast<int>(A.size(0)); g.N = static_cast<int>(B.size(0)); g.K = static_cast<int>(A.size(1));
  TORCH_CHECK(D.size(0) == g.M && D.size(1) == g.N, "D shape");
  TORCH_CHECK(sfa.numel() >= b200::mx_sf_bytes(0, g.M, g.K) && sfb.numel() >= b200::mx_sf_bytes(1, g.N, g.K), "scale buffers too small");
  g.lda = static_cast<int>(A.stride(0)); g.ldb = static_cast<int>(B.stride(0)); g.ldd = static_cast<int>(D.stride(0));
  g.a_e5m2 = a_e5m2; g.b_e5m2 = b_e5m2; g.bias = optptr(bias); g.accum_f32 = accumulate;
  if (accumulate) {
    TORCH_CHECK(D.scalar_type() == at::kFloat, "accumulate needs fp32 D");
  } else {
    g.out_bf16 = is_bf16(D);
  }
  g.stream = cur_stream();
  int rc = b200::gemm_mxfp8(g);
  TORCH_CHECK(rc == 0, "b200 gemm_mxfp8 failed (", rc, "): ", b200::gemm_mxfp8_last_error());
  count();
}

// ---------------------------------------------------------------- MoE
std::vector<Tensor> moe_router_topk_fwd(const Tensor& logits, int k, bool pre_softmax, double scaling) {
  c10::cuda::CUDAGuard guard(logits.device());
  TORCH_CHECK(logits.dim() == 2 && logits.scalar_type() == at::kFloat && logits.is_contiguous(), "router logits: fp32 [n, E]");
  const int n = static_cast<int>(logits.size(0)), E = static_cast<int>(logits.size(1));
  TORCH_CHECK(E <= 256 && k >= 1 && k <= 8 && k <= E, "router kernel: E <= 256, 1 <= topk <= 8");
  auto probs = torch::empty_like(logits);
  auto map = torch::empty({n, E}, logits.options().dtype(at::kBool));
  auto tpe = torch::zeros({E}, logits.options().dtype(at::kInt));
  b200::moe_router_topk_fwd(logits.data_ptr<float>(), probs.data_ptr<float>(), reinterpret_cast<uint8_t*>(map.data_ptr()),
                            tpe.data_ptr<int32_t>(), n, E, k, pre_softmax, static_cast<float>(scaling), cur_stream());
  count();
  return {probs, map, tpe};
}

Tensor moe_router_topk_bwd(const Tensor& logits, const Tensor& probs, const Tensor& map, const Tensor& dprobs, bool pre_softmax,
                           double scaling) {
  c10::cuda::CUDAGuard guard(logits.device());
  auto dl = torch::empty_like(logits);
  auto dp = dprobs.contiguous();
  b200::moe_router_topk_bwd(logits.data_ptr<float>(), probs.data_ptr<float>(), reinterpret_cast<const uint8_t*>(map.data_ptr()),
                            dp.data_ptr<float>(), dl.data_ptr<float>(), static_cast<int>(logits.size(0)),
                            static_cast<int>(logits.size(1)), pre_softmax, static_cast<float>(scaling), cur_stream());
  count();
  return dl;
}

// -> counts [E_local], offsets [E_local+1], row_slot [cap], row_prob [cap], dest_row [n, E_local]   (all on the device)
std::vector<Tensor> moe_build_layout(const Tensor& map, const c10::optional<Tensor>& probs, int e_begin, int E_local, int pad, int cap) {
  c10::cuda::CUDAGuard guard(map.device());
  TORCH_CHECK(map.dim() == 2 && map.is_contiguous() && (map.scalar_type() == at::kBool || map.scalar_type() == at::kByte),
              "routing map: bool [n, E]");
  const int n = static_cast<int>(map.size(0)), E = static_cast<int>(map.size(1));
  TORCH_CHECK(e_begin >= 0 && e_begin + E_local <= E && cap % pad == 0, "bad expert range / capacity");
  auto io = map.options().dtype(at::kInt);
  auto pos = torch::empty({n, E_local}, io);
  auto counts = torch::empty({E_local}, io);
  auto offsets = torch::empty({E_local + 1}, io);
  auto row_slot = torch::empty({cap}, io);
  auto row_prob = torch::zeros({cap}, map.options().dtype(at::kFloat));
  auto dest_row = torch::empty({n, E_local}, io);
  const float* pp = nullptr;
  if (probs.has_value() && probs->defined()) {
    TORCH_CHECK(probs->scalar_type() == at::kFloat && probs->is_contiguous() && probs->sizes() == map.sizes(), "probs: fp32 [n, E]");
    pp = probs->data_ptr<float>();
  }
  b200::moe_build_layout(reinterpret_cast<const uint8_t*>(map.data_ptr()), pp, n, E, e_begin, E_local, pad, pos.data_ptr<int32_t>(),
                         counts.data_ptr<int32_t>(), offsets.data_ptr<int32_t>(), row_slot.data_ptr<int32_t>(),
                         row_prob.data_ptr<float>(), dest_row.data_ptr<int32_t>(), cap, cur_stream());
  count(3);
  return {counts, offsets, row_slot, row_prob, dest_row};
}

// out[r] = scale[r] * in[row_slot[r] / slot_div]; optional dot_out[row_slot[r]] = <in_row, other[r]> / scale[r]
Tensor moe_gather_rows(const Tensor& in, const Tensor& row_slot, int slot_div, const c10::optional<Tensor>& row_scale,
                       const Tensor& offsets, int n_groups, const c10::optional<Tensor>& other, c10::optional<Tensor> dot_out) {
  c10::cuda::CUDAGuard guard(in.device());
  check_2d(in, "in");
  TORCH_CHECK(in.is_contiguous() && in.size(1) % 8 == 0, "gather_rows: contiguous rows, cols % 8 == 0");
  const int cap = static_cast<int>(row_slot.numel()), cols = static_cast<int>(in.size(1));
  auto out = torch::empty({cap, cols}, in.options());
  const float* sc = row_scale.has_value() && row_scale->defined() ? row_scale->data_ptr<float>() : nullptr;
  float* dp = dot_out.has_value() && dot_out->defined() ? dot_out->data_ptr<float>() : nullptr;
  TORCH_CHECK(dp == nullptr || (sc != nullptr && other.has_value()), "dot output needs row_scale and other");
  b200::moe_gather_rows(in.data_ptr(), row_slot.data_ptr<int32_t>(), slot_div, sc, offsets.data_ptr<int32_t>(), n_groups, cap, cols,
                        out.data_ptr(), optptr(other), dp, is_bf16(in), cur_stream());
  count();
  return out;
}

Tensor moe_gather_sum(const Tensor& y, const Tensor& dest_row) {
  c10::cuda::CUDAGuard guard(y.device());
  check_2d(y, "y");
  TORCH_CHECK(y.is_contiguous() && y.size(1) % 8 == 0 && dest_row.dim() == 2, "gather_sum: contiguous y, dest_row [n, E_local]");
  const int n = static_cast<int>(dest_row.size(0)), el = static_cast<int>(dest_row.size(1));
  auto out = torch::empty({n, y.size(1)}, y.options());
  b200::moe_gather_sum(y.data_ptr(), dest_row.data_ptr<int32_t>(), n, el, static_cast<int>(y.size(1)), out.data_ptr(), is_bf16(y),
                       cur_stream());
  count();
  return out;
}

// ---- expert parallel over the EP arena: pointer tables arrive as lists of device addresses (one per EP rank)
template <typename T>
std::vector<T*> ptrs_of(const std::vector<int64_t>& v) {
  std::vector<T*> o(b200::kMaxPeers, nullptr);
  TORCH_CHECK(v.size() <= static_cast<size_t>(b200::kMaxPeers), "at most 8 expert-parallel peers");
  for (size_t i = 0; i < v.size(); ++i) o[i] = reinterpret_cast<T*>(v[i]);
  return o;
}

void moe_ep_dispatch(const Tensor& x, const Tensor& send_offsets, const Tensor& send_row, const Tensor& counts_send,
                     const c10::optional<Tensor>& probs, std::vector<int64_t> stage_x, std::vector<int64_t> stage_idx,
                     std::vector<int64_t> stage_prob, std::vector<int64_t> recv_counts, std::vector<int64_t> flags, int E_local,
                     int me, int Rs, int64_t epoch, Tensor& done_counter) {
  c10::cuda::CUDAGuard guard(x.device());
  check_2d(x, "x");
  TORCH_CHECK(x.is_contiguous() && x.size(1) % 8 == 0 && send_row.dim() == 2 && send_row.size(0) == x.size(0), "ep_dispatch shapes");
  const int ep = static_cast<int>(stage_x.size());
  const int E_total = static_cast<int>(send_row.size(1));
  TORCH_CHECK(E_total == ep * E_local && E_local <= 64, "E_total == ep * E_local, E_local <= 64");
  is_bf16(x);
  const float* pp = probs.has_value() && probs->defined() ? probs->data_ptr<float>() : nullptr;
  auto sx = ptrs_of<void>(stage_x); auto si = ptrs_of<int32_t>(stage_idx); auto sp = ptrs_of<float>(stage_prob);
  auto rc = ptrs_of<int32_t>(recv_counts); auto fl = ptrs_of<uint32_t>(flags);
  b200::moe_ep_dispatch(x.data_ptr(), send_offsets.data_ptr<int32_t>(), send_row.data_ptr<int32_t>(), counts_send.data_ptr<int32_t>(),
                        pp, sx.data(), si.data(), sp.data(), rc.data(), fl.data(), static_cast<int>(x.size(0)), E_total, E_local, me,
                        ep, Rs, static_cast<int>(x.size(1)), static_cast<uint32_t>(epoch), done_counter.data_ptr<int32_t>(), cur_stream());
  count();
}

// -> offsets [E_local+1], counts [E_local], row_stage [cap]; *overflow is set when the compact capacity was exceeded
std::vector<Tensor> moe_ep_recv_layout(int64_t flags_ptr, int64_t recv_counts_ptr, int ep, int E_local, int Rs, int pad, int cap,
                                       int64_t epoch, Tensor& overflow) {
  c10::cuda::CUDAGuard guard(overflow.device());
  TORCH_CHECK(E_local <= 64 && ep <= b200::kMaxPeers && cap % pad == 0, "ep_recv_layout limits");
  auto io = overflow.options().dtype(at::kInt);
  auto offsets = torch::empty({E_local + 1}, io);
  auto counts = torch::empty({E_local}, io);
  auto row_stage = torch::empty({cap}, io);
  b200::moe_ep_recv_layout(reinterpret_cast<const uint32_t*>(flags_ptr), reinterpret_cast<const int32_t*>(recv_counts_ptr), ep, E_local, Rs,
                           pad, cap, static_cast<uint32_t>(epoch), offsets.data_ptr<int32_t>(), counts.data_ptr<int32_t>(),
                           row_stage.data_ptr<int32_t>(), overflow.data_ptr<int32_t>(), cur_stream());
  count();
  return {offsets, counts, row_stage};
}

void moe_ep_combine_push(const Tensor& z, const Tensor& row_stage, const Tensor& row_idx, const c10::optional<Tensor>& side,
                         const Tensor& offsets, int n_groups, int Rs, std::vector<int64_t> ret_x, std::vector<int64_t> ret_f,
                         std::vector<int64_t> flags, int me, int64_t epoch, Tensor& done_counter) {
  c10::cuda::CUDAGuard guard(z.device());
  check_2d(z, "z");
  TORCH_CHECK(z.is_contiguous() && z.size(1) % 8 == 0 && z.size(0) == row_stage.numel(), "ep_combine_push shapes");
  is_bf16(z);
  const int ep = static_cast<int>(ret_x.size());
  auto rx = ptrs_of<void>(ret_x); auto rf = ptrs_of<float>(ret_f); auto fl = ptrs_of<uint32_t>(flags);
  const float* sd = side.has_value() && side->defined() ? side->data_ptr<float>() : nullptr;
  b200::moe_ep_combine_push(z.data_ptr(), row_stage.data_ptr<int32_t>(), row_idx.data_ptr<int32_t>(), sd,
                            offsets.data_ptr<int32_t>(), n_groups, static_cast<int>(z.size(0)), Rs, static_cast<int>(z.size(1)),
                            rx.data(), rf.data(), fl.data(), me, ep, static_cast<uint32_t>(epoch), done_counter.data_ptr<int32_t>(),
                            cur_stream());
  count();
}

void moe_ep_wait(int64_t flags_ptr, int ep, int64_t epoch) {
  b200::moe_ep_wait(reinterpret_cast<const uint32_t*>(flags_ptr), ep, static_cast<uint32_t>(epoch), cur_stream());
  count();
}

Tensor moe_scatter_dprobs(const Tensor& ret_f, const Tensor& send_row) {
  c10::cuda::CUDAGuard guard(ret_f.device());
  auto out = torch::empty(send_row.sizes(), ret_f.options());
  b200::moe_scatter_dprobs(ret_f.data_ptr<float>(), send_row.data_ptr<int32_t>(), out.data_ptr<float>(), send_row.numel(), cur_stream());
  count();
  return out;
}

// ---------------------------------------------------------------- norms etc.
std::vector<Tensor> layernorm_fwd(const Tensor& x, const Tensor& gamma, const c10::optional<Tensor>& beta,
                                  double eps, bool rms, bool zero_centered) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.size(-1) % 8 == 0, "layernorm: contiguous input with cols % 8 == 0");
  const int cols = static_cast<int>(x.size(-1));
  const int rows = static_cast<int>(x.numel() / cols);
  auto y = torch::empty_like(x);
  auto fo = x.options().dtype(at::kFloat);
  auto mean = rms ? torch::empty({0}, fo) : torch::empty({rows}, fo);
  auto rstd = torch::empty({rows}, fo);
  b200::layernorm_fwd(x.data_ptr(), gamma.data_ptr(), optptr(beta), y.data_ptr(),
                      rms ? nullptr : mean.data_ptr<float>(), rstd.data_ptr<float>(), rows, cols,
                      static_cast<float>(eps), rms, zero_centered, is_bf16(x), cur_stream());
  count();
  return {y, mean, rstd};
}

std::vector<Tensor> layernorm_bwd(const Tensor& dy, const Tensor& x, const Tensor& gamma, const Tensor& mean,
                                  const Tensor& rstd, bool rms, bool zero_centered, bool has_beta) {
  c10::cuda::CUDAGuard guard(x.device());
  const int cols = static_cast<int>(x.size(-1));
  const int rows = static_cast<int>(x.numel() / cols);
  TORCH_CHECK(cols <= 8192, "layernorm_bwd supports cols <= 8192");
  auto dx = torch::empty_like(x);
  const int np = b200::layernorm_bwd_partials(rows);
  auto fo = x.options().dtype(at::kFloat);
  auto pg = torch::empty({np, cols}, fo);
  auto pb = has_beta ? torch::empty({np, cols}, fo) : torch::empty({0}, fo);
  auto dgamma = torch::empty({cols}, fo);
  auto dbeta = has_beta ? torch::empty({cols}, fo) : torch::empty({0}, fo);
  auto dyc = dy.contiguous();
  b200::layernorm_bwd(dyc.data_ptr(), x.data_ptr(), gamma.data_ptr(), rms ? nullptr : mean.data_ptr<float>(),
                      rstd.data_ptr<float>(), dx.data_ptr(), pg.data_ptr<float>(),
                      has_beta ? pb.data_ptr<float>() : nullptr, rows, cols, rms, zero_centered,
                      is_bf16(x), np, cur_stream());
  b200::layernorm_bwd_finalize(pg.data_ptr<float>(), has_beta ? pb.data_ptr<float>() : nullptr,
                               dgamma.data_ptr<float>(), has_beta ? dbeta.data_ptr<float>() : nullptr, np,
                               cols, false, cur_stream());
  count(2);
  return {dx, dgamma, dbeta};
}

Tensor bias_gelu_fwd(const Tensor& x, const c10::optional<Tensor>& bias) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.size(-1) % 8 == 0);
  auto y = torch::empty_like(x);
  const int cols = static_cast<int>(x.size(-1));
  b200::bias_gelu_fwd(x.data_ptr(), optptr(bias), y.data_ptr(), x.numel() / cols, cols, is_bf16(x), cur_stream());
  count();
  return y;
}
Tensor bias_gelu_bwd(const Tensor& dy, const Tensor& x, const c10::optional<Tensor>& bias) {
  c10::cuda::CUDAGuard guard(x.device());
  auto dyc = dy.contiguous();
  auto dx = torch::empty_like(x);
  const int cols = static_cast<int>(x.size(-1));
  b200::bias_gelu_bwd(dyc.data_ptr(), x.data_ptr(), optptr(bias), dx.data_ptr(), x.numel() / cols, cols, is_bf16(x), cur_stream());
  count();
  return dx;
}
Tensor swiglu_fwd(const Tensor& x, const c10::optional<Tensor>& bias) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.size(-1) % 16 == 0);
  auto sizes = x.sizes().vec();
  const int cols_out = static_cast<int>(sizes.back() / 2);
  sizes.back() = cols_out;
  auto y = torch::empty(sizes, x.options());
  b200::swiglu_fwd(x.data_ptr(), optptr(bias), y.data_ptr(), x.numel() / (2 * cols_out), cols_out, is_bf16(x), cur_stream());
  count();
  return y;
}
Tensor swiglu_bwd(const Tensor& dy, const Tensor& x, const c10::optional<Tensor>& bias) {
  c10::cuda::CUDAGuard guard(x.device());
  auto dyc = dy.contiguous();
  auto dx = torch::empty_like(x);
  const int cols_out = static_cast<int>(x.size(-1) / 2);
  b200::swiglu_bwd(dyc.data_ptr(), x.data_ptr(), optptr(bias), dx.data_ptr(), x.numel() / (2 * cols_out), cols_out, is_bf16(x), cur_stream());
  count();
  return dx;
}
std::vector<Tensor> bias_dropout_add(const Tensor& x, const c10::optional<Tensor>& bias,
                                     const c10::optional<Tensor>& residual, double p, int64_t seed, int64_t offset) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.size(-1) % 8 == 0);
  auto out = torch::empty_like(x);
  Tensor mask = p > 0 ? torch::empty(x.sizes(), x.options().dtype(at::kByte)) : torch::empty({0}, x.options().dtype(at::kByte));
  const int cols = static_cast<int>(x.size(-1));
  const void* res = nullptr;
  Tensor resc;
  if (residual.has_value() && residual->defined()) { resc = residual->contiguous(); res = resc.data_ptr(); }
  b200::bias_dropout_add(x.data_ptr(), optptr(bias), res, out.data_ptr(), p > 0 ? mask.data_ptr<uint8_t>() : nullptr,
                         x.numel() / cols, cols, static_cast<float>(p), static_cast<uint64_t>(seed),
                         static_cast<uint64_t>(offset), is_bf16(x), cur_stream());
  count();
  return {out, mask};
}
Tensor dropout_bwd(const Tensor& dy, const Tensor& mask, double p) {
  c10::cuda::CUDAGuard guard(dy.device());
  auto dyc = dy.contiguous();
  auto dx = torch::empty_like(dyc);
  b200::dropout_bwd(dyc.data_ptr(), mask.data_ptr<uint8_t>(), dx.data_ptr(), dyc.numel(), static_cast<float>(p), is_bf16(dyc), cur_stream());
  count();
  return dx;
}
Tensor rope_apply(const Tensor& x, const Tensor& cos_sin, bool interleaved, bool inverse) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.dim() == 4 && x.is_contiguous(), "rope expects contiguous [s, b, h, d]");
  TORCH_CHECK(cos_sin.scalar_type() == at::kFloat && cos_sin.is_contiguous() && cos_sin.size(0) >= x.size(0));
  auto y = torch::empty_like(x);
  b200::rope_apply(x.data_ptr(), cos_sin.data_ptr<float>(), y.data_ptr(), static_cast<int>(x.size(0)),
                   static_cast<int>(x.size(1)), static_cast<int>(x.size(2)), static_cast<int>(x.size(3)),
                   static_cast<int>(cos_sin.size(-1)), interleaved, inverse, is_bf16(x), cur_stream());
  count();
  return y;
}
void bias_grad(const Tensor& dy, Tensor& dbias, bool accumulate) {
  c10::cuda::CUDAGuard guard(dy.device());
  TORCH_CHECK(dy.is_contiguous() && dbias.scalar_type() == at::kFloat);
  const int cols = static_cast<int>(dy.size(-1));
  TORCH_CHECK(cols % 8 == 0);
  b200::bias_grad_reduce(dy.data_ptr(), dbias.data_ptr<float>(), dy.numel() / cols, cols, is_bf16(dy), accumulate, cur_stream());
  count();
}

std::vector<Tensor> ce_local_stats(const Tensor& logits, const Tensor& labels, int64_t vocab_start) {
  c10::cuda::CUDAGuard guard(logits.device());
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1 && labels.scalar_type() == at::kLong && labels.is_contiguous());
  const int64_t rows = logits.size(0);
  auto fo = logits.options().dtype(at::kFloat);
  auto mx = torch::empty({rows}, fo), se = torch::empty({rows}, fo), tg = torch::empty({rows}, fo);
  b200::ce_local_stats(logits.data_ptr(), labels.data_ptr<int64_t>(), mx.data_ptr<float>(), se.data_ptr<float>(),
                       tg.data_ptr<float>(), rows, static_cast<int>(logits.size(1)), static_cast<int>(vocab_start),
                       static_cast<int>(logits.stride(0)), dtype_code(logits), cur_stream());
  count();
  return {mx, se, tg};
}
void ce_rescale(Tensor& sumexp, const Tensor& local_max, const Tensor& global_max) {
  c10::cuda::CUDAGuard guard(sumexp.device());
  b200::ce_rescale_sumexp(sumexp.data_ptr<float>(), local_max.data_ptr<float>(), global_max.data_ptr<float>(), sumexp.numel(), cur_stream());
  count();
}
void ce_grad(const Tensor& logits, const Tensor& labels, const Tensor& gmax, const Tensor& gsum,
             const Tensor& grad_loss, Tensor& dlogits, int64_t vocab_start, double label_smoothing, int64_t vocab_total) {
  c10::cuda::CUDAGuard guard(logits.device());
  TORCH_CHECK(grad_loss.scalar_type() == at::kFloat && grad_loss.is_contiguous());
  b200::ce_grad(logits.data_ptr(), labels.data_ptr<int64_t>(), gmax.data_ptr<float>(), gsum.data_ptr<float>(),
                grad_loss.data_ptr<float>(), dlogits.data_ptr(), logits.size(0), static_cast<int>(logits.size(1)),
                static_cast<int>(vocab_start), static_cast<int>(logits.stride(0)), dtype_code(logits),
                static_cast<float>(label_smoothing), vocab_total, cur_stream());
  count();
}

void adam_step(Tensor& master, Tensor& exp_avg, Tensor& exp_avg_sq, const Tensor& grad,
               const c10::optional<Tensor>& param16, double lr, double beta1, double beta2, double eps,
               double weight_decay, double grad_scale, int64_t step, const c10::optional<Tensor>& found_inf) {
  c10::cuda::CUDAGuard guard(master.device());
  b200::AdamArgs a{};
  a.master = master.data_ptr<float>(); a.exp_avg = exp_avg.data_ptr<float>(); a.exp_avg_sq = exp_avg_sq.data_ptr<float>();
  a.grad = grad.data_ptr<float>(); a.param16 = const_cast<void*>(optptr(param16)); a.n = master.numel();
  TORCH_CHECK(grad.numel() == a.n && exp_avg.numel() == a.n);
  a.lr = static_cast<float>(lr); a.beta1 = static_cast<float>(beta1); a.beta2 = static_cast<float>(beta2);
  a.eps = static_cast<float>(eps); a.weight_decay = static_cast<float>(weight_decay); a.grad_scale = static_cast<float>(grad_scale);
  a.bias_correction1 = static_cast<float>(1.0 - std::pow(beta1, static_cast<double>(step)));
  a.bias_correction2 = static_cast<float>(1.0 - std::pow(beta2, static_cast<double>(step)));
  a.is_bf16 = param16.has_value() && param16->defined() ? param16->scalar_type() == at::kBFloat16 : true;
  a.found_inf = found_inf.has_value() && found_inf->defined() ? found_inf->data_ptr<float>() : nullptr;
  b200::adam_step(a, cur_stream());
  count();
}
void l2norm_sq(const Tensor& x, Tensor& out) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && out.scalar_type() == at::kFloat);
  b200::l2norm_sq(x.data_ptr(), x.numel(), dtype_code(x), out.data_ptr<float>(), cur_stream());
  count();
}
void scale_(Tensor& x, double scale, const c10::optional<Tensor>& scale_dev) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous());
  b200::scale_inplace(x.data_ptr(), x.numel(), dtype_code(x), static_cast<float>(scale),
                      scale_dev.has_value() && scale_dev->defined() ? scale_dev->data_ptr<float>() : nullptr, cur_stream());
  count();
}
void cast_copy(const Tensor& src, Tensor& dst, double scale) {
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.is_contiguous() && dst.is_contiguous() && src.numel() == dst.numel());
  b200::cast_copy(src.data_ptr(), dtype_code(src), dst.data_ptr(), dtype_code(dst), src.numel(), static_cast<float>(scale), cur_stream());
  count();
}
void add_into_f32(Tensor& dst, const Tensor& src) {
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.is_contiguous() && dst.is_contiguous() && dst.scalar_type() == at::kFloat && src.numel() == dst.numel());
  b200::add_into_f32(dst.data_ptr<float>(), src.data_ptr(), dtype_code(src), src.numel(), cur_stream());
  count();
}
void check_finite(const Tensor& x, Tensor& found) {
  c10::cuda::CUDAGuard guard(x.device());
  b200::check_finite(x.data_ptr(), x.numel(), dtype_code(x), found.data_ptr<float>(), cur_stream());
  count();
}
Tensor chunk_mean(const Tensor& x, int64_t pixels) {
  c10::cuda::CUDAGuard guard(x.device());
  auto xc = x.contiguous();
  const int cols = static_cast<int>(xc.size(-1));
  const int64_t rows = xc.numel() / cols;
  auto sizes = xc.sizes().vec();
  sizes.back() = pixels;
  auto out = torch::empty(sizes, xc.options().dtype(at::kFloat));
  b200::chunk_mean(xc.data_ptr(), out.data_ptr<float>(), rows, cols, static_cast<int>(pixels), is_bf16(xc), cur_stream());
  count();
  return out;
}

// ---------------------------------------------------------------- fp8
// Quantise a 16-bit matrix with current (tensor-wise) scaling entirely on the device:
// amax -> scale -> cast (+ transposed copy).  Returns {q, qt (or empty), scale_inv, amax}.
std::vector<Tensor> fp8_quantize(const Tensor& x, bool e5m2, bool want_transpose, int64_t margin,
                                 const c10::optional<Tensor>& amax_in) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.dim() == 2 && x.is_contiguous() && x.size(1) % 8 == 0, "fp8_quantize expects contiguous [rows, cols], cols % 8 == 0");
  const int rows = static_cast<int>(x.size(0)), cols = static_cast<int>(x.size(1));
  auto fopt = x.options().dtype(at::kFloat);
  Tensor amax = amax_in.has_value() ? amax_in.value() : torch::zeros({1}, fopt);
  if (!amax_in.has_value()) b200::fp8_amax(x.data_ptr(), x.numel(), is_bf16(x), amax.data_ptr<float>(), cur_stream());
  auto scales = torch::empty({2}, fopt);
  b200::fp8_scale_from_amax(amax.data_ptr<float>(), e5m2, static_cast<int>(margin), scales.data_ptr<float>(),
                            scales.data_ptr<float>() + 1, cur_stream());
  auto bopt = x.options().dtype(at::kByte);
  auto q = torch::empty({rows, cols}, bopt);
  Tensor qt = want_transpose ? torch::empty({cols, rows}, bopt) : torch::empty({0}, bopt);
  b200::fp8_cast_transpose(x.data_ptr(), scales.data_ptr<float>(), q.data_ptr(), want_transpose ? qt.data_ptr() : nullptr,
                           rows, cols, is_bf16(x), e5m2, cur_stream());
  count(3);
  return {q, qt, scales.slice(0, 1, 2), amax};
}

// D = sa_inv * sb_inv * A8 @ B8^T (+bias).  ``D`` 16-bit (store) or fp32 (accumulate=true: "+=").
void gemm_fp8(const Tensor& A8, const Tensor& B8, Tensor& D, const Tensor& sa_inv, const Tensor& sb_inv, bool a_e5m2,
              bool b_e5m2, const c10::optional<Tensor>& bias, bool accumulate) {
  c10::cuda::CUDAGuard guard(A8.device());
  TORCH_CHECK(A8.dim() == 2 && B8.dim() == 2 && A8.size(1) == B8.size(1), "fp8 gemm: A[M,K], B[N,K]");
  TORCH_CHECK(A8.scalar_type() == at::kByte && B8.scalar_type() == at::kByte, "fp8 operands are uint8 storage");
  b200::Fp8GemmArgs a;
  a.A = A8.data_ptr(); a.B = B8.data_ptr(); a.D = D.data_ptr();
  a.M = static_cast<int>(A8.size(0)); a.N = static_cast<int>(B8.size(0)); a.K = static_cast<int>(A8.size(1));
  a.lda = static_cast<int>(A8.stride(0)); a.ldb = static_cast<int>(B8.stride(0)); a.ldd = static_cast<int>(D.stride(0));
  a.a_e5m2 = a_e5m2; a.b_e5m2 = b_e5m2;
  a.scale_a_inv = sa_inv.data_ptr<float>(); a.scale_b_inv = sb_inv.data_ptr<float>();
  a.bias = optptr(bias);
  a.accum_f32 = accumulate;
  if (accumulate) {
    TORCH_CHECK(D.scalar_type() == at::kFloat, "accumulate needs an fp32 output");
  } else {
    a.out_bf16 = is_bf16(D);
  }
  a.stream = cur_stream();
  int rc = b200::gemm_fp8(a);
  TORCH_CHECK(rc == 0, "b200 gemm_fp8 failed: ", rc, " ", b200::gemm_fp8_last_error());
  count();
}

// ---------------------------------------------------------------- attention
namespace {
void check_bhsd(const Tensor& t, const char* name) {
  TORCH_CHECK(t.dim() == 4 && t.stride(3) == 1, name, " must be a [b,h,s,d] view with contiguous last dim");
}
}  // namespace
// q/k/v: strided logical [b,h,s,d] views (e.g. permuted [s,b,h,d] activations).  The output is
// allocated [s,b,h,d]-contiguous and returned as its [b,h,s,d] view.
std::vector<Tensor> attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, double scale, bool causal, bool want_probs) {
  c10::cuda::CUDAGuard guard(q.device());
  check_bhsd(q, "q"); check_bhsd(k, "k"); check_bhsd(v, "v");
  b200::AttnArgs a{};
  a.b = static_cast<int>(q.size(0)); a.hq = static_cast<int>(q.size(1)); a.sq = static_cast<int>(q.size(2)); a.d = static_cast<int>(q.size(3));
  a.hkv = static_cast<int>(k.size(1)); a.sk = static_cast<int>(k.size(2));
  auto o_sbhd = torch::empty({a.sq, a.b, a.hq, a.d}, q.options());
  auto o = o_sbhd.permute({1, 2, 0, 3});
  auto lse = torch::empty({a.b, a.hq, a.sq}, q.options().dtype(at::kFloat));
  Tensor probs = torch::empty({0}, q.options());
  a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.o = o_sbhd.data_ptr(); a.lse = lse.data_ptr<float>();
  a.q_sb = q.stride(0); a.q_sh = q.stride(1); a.q_ss = q.stride(2);
  a.k_sb = k.stride(0); a.k_sh = k.stride(1); a.k_ss = k.stride(2);
  a.v_sb = v.stride(0); a.v_sh = v.stride(1); a.v_ss = v.stride(2);
  a.o_sb = o.stride(0); a.o_sh = o.stride(1); a.o_ss = o.stride(2);
  a.scale = static_cast<float>(scale); a.causal = causal; a.is_bf16 = is_bf16(q);
  a.probs_out = want_probs ? reinterpret_cast<void*>(1) : nullptr;
  int rc = b200::attn_fwd(a, cur_stream());
  TORCH_CHECK(rc == 0, "b200 attn_fwd failed: ", rc, " ", b200::attn_last_error());
  count();
  return {o, lse, probs};
}
std::vector<Tensor> attn_bwd(const Tensor& d_o, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& o,
                             const Tensor& lse, double scale, bool causal, const c10::optional<Tensor>& dq_out,
                             const c10::optional<Tensor>& dk_out, const c10::optional<Tensor>& dv_out) {
  c10::cuda::CUDAGuard guard(q.device());
  check_bhsd(q, "q"); check_bhsd(k, "k"); check_bhsd(v, "v"); check_bhsd(o, "o");
  b200::AttnArgs a{};
  a.b = static_cast<int>(q.size(0)); a.hq = static_cast<int>(q.size(1)); a.sq = static_cast<int>(q.size(2)); a.d = static_cast<int>(q.size(3));
  a.hkv = static_cast<int>(k.size(1)); a.sk = static_cast<int>(k.size(2));
  Tensor doc = d_o.stride(3) == 1 ? d_o : d_o.contiguous();
  // gradients go either to fresh [s,b,h,d] buffers or straight into caller-provided strided views
  // (e.g. the three slices of one packed d(qkv) tensor: no concatenation afterwards)
  Tensor dq = dq_out.has_value() ? dq_out.value() : torch::empty({a.sq, a.b, a.hq, a.d}, q.options()).permute({1, 2, 0, 3});
  Tensor dk = dk_out.has_value() ? dk_out.value() : torch::empty({a.sk, a.b, a.hkv, a.d}, q.options()).permute({1, 2, 0, 3});
  Tensor dv = dv_out.has_value() ? dv_out.value() : torch::empty({a.sk, a.b, a.hkv, a.d}, q.options()).permute({1, 2, 0, 3});
  check_bhsd(dq, "dq"); check_bhsd(dk, "dk"); check_bhsd(dv, "dv");
  auto delta = torch::empty({2, a.b, a.hq, a.sq}, q.options().dtype(at::kFloat));   // [delta ; lse * log2e]
  a.q = q.data_ptr(); a.k = k.data_ptr(); a.v = v.data_ptr(); a.o = const_cast<void*>(o.data_ptr()); a.lse = lse.data_ptr<float>();
  a.d_o = doc.data_ptr(); a.dq = dq.data_ptr(); a.dk = dk.data_ptr(); a.dv = dv.data_ptr();
  a.delta = delta.data_ptr<float>(); a.dq_accum = nullptr;
  a.q_sb = q.stride(0); a.q_sh = q.stride(1); a.q_ss = q.stride(2);
  a.k_sb = k.stride(0); a.k_sh = k.stride(1); a.k_ss = k.stride(2);
  a.v_sb = v.stride(0); a.v_sh = v.stride(1); a.v_ss = v.stride(2);
  a.o_sb = o.stride(0); a.o_sh = o.stride(1); a.o_ss = o.stride(2);
  a.do_sb = doc.stride(0); a.do_sh = doc.stride(1); a.do_ss = doc.stride(2);
  a.dq_sb = dq.stride(0); a.dq_sh = dq.stride(1); a.dq_ss = dq.stride(2);
  a.dk_sb = dk.stride(0); a.dk_sh = dk.stride(1); a.dk_ss = dk.stride(2);
  a.dv_sb = dv.stride(0); a.dv_sh = dv.stride(1); a.dv_ss = dv.stride(2);
  a.scale = static_cast<float>(scale); a.causal = causal; a.is_bf16 = is_bf16(q);
  int rc = b200::attn_bwd(a, cur_stream());
  TORCH_CHECK(rc == 0, "b200 attn_bwd failed: ", rc, " ", b200::attn_last_error());
  count(3);
  return {dq, dk, dv};
}

// ---------------------------------------------------------------- arena / comm
py::tuple arena_alloc(int64_t nbytes) {
  void* p = nullptr;
  auto e = cudaMalloc(&p, static_cast<size_t>(nbytes));
  TORCH_CHECK(e == cudaSuccess, "arena cudaMalloc failed: ", cudaGetErrorString(e));
  cudaMemset(p, 0, static_cast<size_t>(nbytes));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  TORCH_CHECK(e == cudaSuccess, "cudaIpcGetMemHandle failed: ", cudaGetErrorString(e));
  cudaDeviceSynchronize();
  return py::make_tuple(reinterpret_cast<int64_t>(p), py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
}
int64_t arena_open(py::bytes handle) {
  std::string s = handle;
  TORCH_CHECK(s.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  memcpy(&h, s.data(), sizeof(h));
  void* p = nullptr;
  auto e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  TORCH_CHECK(e == cudaSuccess, "cudaIpcOpenMemHandle failed: ", cudaGetErrorString(e));
  return reinterpret_cast<int64_t>(p);
}
void arena_close(int64_t ptr) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)); }
void arena_free(int64_t ptr) { cudaFree(reinterpret_cast<void*>(ptr)); }
Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> shape, py::object dtype, int64_t device) {
  auto st = torch::python::detail::py_object_to_dtype(dtype);
  auto opts = torch::TensorOptions().dtype(st).device(torch::kCUDA, static_cast<int>(device));
  return torch::from_blob(reinterpret_cast<void*>(ptr), shape, [](void*) {}, opts);
}

void flags_barrier(std::vector<int64_t> peer_flags, int rank, int size, int slot, int64_t epoch) {
  uint32_t* pf[b200::kMaxPeers] = {nullptr};
  for (int i = 0; i < size; ++i) pf[i] = reinterpret_cast<uint32_t*>(peer_flags[i]);
  b200::push_flags_barrier(pf, rank, size, slot, static_cast<uint32_t>(epoch), cur_stream());
  count();
}
void bucket_reduce_scatter(const Tensor& src, Tensor& dst, std::vector<int64_t> peer_stage,
                           std::vector<int64_t> peer_flags, int rank, int size, double scale, int64_t epoch,
                           int wire_dtype, int n_ctas) {
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.scalar_type() == at::kFloat && dst.scalar_type() == at::kFloat && src.is_contiguous() && dst.is_contiguous());
  TORCH_CHECK(src.numel() % (size * 4) == 0, "bucket size must be a multiple of 4*dp");
  TORCH_CHECK(dst.numel() == src.numel() / size);
  b200::RsArgs a{};
  a.src = src.data_ptr<float>(); a.dst = dst.data_ptr<float>(); a.n = src.numel(); a.rank = rank; a.size = size;
  a.scale = static_cast<float>(scale); a.epoch = static_cast<uint32_t>(epoch); a.wire_dtype = wire_dtype; a.n_ctas = n_ctas;
  for (int i = 0; i < size; ++i) { a.peer_stage[i] = reinterpret_cast<void*>(peer_stage[i]); a.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]); }
  b200::bucket_reduce_scatter(a, cur_stream());
  count();
}
void shard_all_gather_cast(const Tensor& src_shard, std::vector<int64_t> peer_dst, std::vector<int64_t> peer_flags,
                           int rank, int size, int64_t epoch, int dst_dtype, int n_ctas) {
  c10::cuda::CUDAGuard guard(src_shard.device());
  TORCH_CHECK(src_shard.scalar_type() == at::kFloat && src_shard.is_contiguous() && src_shard.numel() % 4 == 0);
  b200::AgArgs a{};
  a.src_shard = src_shard.data_ptr<float>(); a.shard_n = src_shard.numel(); a.rank = rank; a.size = size;
  a.epoch = static_cast<uint32_t>(epoch); a.dst_dtype = dst_dtype; a.n_ctas = n_ctas;
  for (int i = 0; i < size; ++i) { a.peer_dst[i] = reinterpret_cast<void*>(peer_dst[i]); a.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]); }
  b200::shard_all_gather_cast(a, cur_stream());
  count();
}
void a2a_scatter_rows(const Tensor& src, const Tensor& dst_rank, const Tensor& dst_row, std::vector<int64_t> peer_buf,
                      std::vector<int64_t> peer_flags, int rank, int size, int64_t epoch, int n_ctas) {
  c10::cuda::CUDAGuard guard(src.device());
  TORCH_CHECK(src.dim() == 2 && src.is_contiguous() && src.size(1) % 8 == 0);
  TORCH_CHECK(dst_rank.scalar_type() == at::kInt && dst_row.scalar_type() == at::kInt);
  b200::A2aArgs a{};
  a.src = src.data_ptr(); a.dst_rank = dst_rank.data_ptr<int32_t>(); a.dst_row = dst_row.data_ptr<int32_t>();
  a.rows = src.size(0); a.cols = static_cast<int>(src.size(1)); a.rank = rank; a.size = size;
  a.epoch = static_cast<uint32_t>(epoch); a.is_bf16 = is_bf16(src); a.n_ctas = n_ctas;
  for (int i = 0; i < size; ++i) { a.peer_buf[i] = reinterpret_cast<void*>(peer_buf[i]); a.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]); }
  b200::a2a_scatter_rows(a, cur_stream());
  count();
}
void embed_scatter(const Tensor& tokens, const Tensor& weight, std::vector<int64_t> peer_out, std::vector<int64_t> peer_flags,
                   int64_t vocab_start, int rank, int size, int64_t epoch, int n_ctas) {
  c10::cuda::CUDAGuard guard(weight.device());
  TORCH_CHECK(tokens.dim() == 2 && tokens.is_contiguous() && tokens.scalar_type() == at::kLong, "tokens must be [b, s] int64");
  TORCH_CHECK(weight.dim() == 2 && weight.is_contiguous() && weight.element_size() == 2 && weight.size(1) % 8 == 0);
  TORCH_CHECK(tokens.size(1) % size == 0, "sequence length must divide over the group");
  b200::EmbedArgs a{};
  a.tokens = tokens.data_ptr<int64_t>(); a.weight = weight.data_ptr();
  a.b = static_cast<int>(tokens.size(0)); a.s = static_cast<int>(tokens.size(1)); a.h = static_cast<int>(weight.size(1));
  a.vocab_start = vocab_start; a.vocab_end = vocab_start + weight.size(0); a.rank = rank; a.size = size;
  a.epoch = static_cast<uint32_t>(epoch); a.n_ctas = n_ctas;
  for (int i = 0; i < size; ++i) { a.peer_out[i] = reinterpret_cast<void*>(peer_out[i]); a.peer_flags[i] = reinterpret_cast<uint32_t*>(peer_flags[i]); }
  b200::embed_scatter(a, cur_stream());
  count();
}
Tensor paged_attention(const Tensor& q, const Tensor& k_cache, const Tensor& v_cache, const Tensor& block_table,
                       const Tensor& tok_req, const Tensor& tok_len, int64_t block_size, double scale, int64_t splits) {
  c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(q.dim() == 3 && q.is_contiguous() && k_cache.dim() == 3 && k_cache.is_contiguous() && v_cache.is_contiguous());
  TORCH_CHECK(q.scalar_type() == k_cache.scalar_type() && (is_bf16(q) || q.scalar_type() == at::kHalf));
  TORCH_CHECK(block_table.scalar_type() == at::kInt && tok_req.scalar_type() == at::kInt && tok_len.scalar_type() == at::kInt);
  TORCH_CHECK(block_table.dim() == 2 && block_table.is_contiguous() && v_cache.sizes() == k_cache.sizes());
  const int T = static_cast<int>(q.size(0)), np = static_cast<int>(q.size(1)), d = static_cast<int>(q.size(2));
  const int ng = static_cast<int>(k_cache.size(1));
  TORCH_CHECK(np % ng == 0 && k_cache.size(2) == d && b200::paged_attention_supported(d, np / ng), "unsupported paged attention shape");
  auto out = at::empty_like(q);
  b200::PagedAttnArgs a{};
  a.q = q.data_ptr(); a.k_cache = k_cache.data_ptr(); a.v_cache = v_cache.data_ptr(); a.out = out.data_ptr();
  a.block_table = block_table.data_ptr<int32_t>(); a.tok_req = tok_req.data_ptr<int32_t>(); a.tok_len = tok_len.data_ptr<int32_t>();
  a.n_tokens = T; a.ng = ng; a.group = np / ng; a.head_dim = d; a.block_size = static_cast<int>(block_size);
  a.max_blocks = static_cast<int>(block_table.size(1)); a.scale_log2 = static_cast<float>(scale * 1.4426950408889634);
  a.is_bf16 = is_bf16(q);
  Tensor ws_o, ws_ml;
  if (splits > 1) {
    ws_o = at::empty({static_cast<int64_t>(T) * np * splits, d}, q.options().dtype(at::kFloat));
    ws_ml = at::empty({static_cast<int64_t>(T) * np * splits, 2}, q.options().dtype(at::kFloat));
    a.ws_o = ws_o.data_ptr<float>(); a.ws_ml = ws_ml.data_ptr<float>();
  }
  if (T > 0) b200::paged_attention(a, static_cast<int>(splits), cur_stream());
  count();
  return out;
}
bool paged_attention_supported(int64_t head_dim, int64_t group) { return b200::paged_attention_supported(static_cast<int>(head_dim), static_cast<int>(group)); }
void spin(int64_t cycles) { b200::spin_kernel(cycles, cur_stream()); count(); }
void write_flag(int64_t flag_ptr, int64_t value) { b200::write_flag(reinterpret_cast<uint32_t*>(flag_ptr), static_cast<uint32_t>(value), cur_stream()); count(); }
void peer_copy(int64_t dst_ptr, const Tensor& src) {
  auto e = cudaMemcpyAsync(reinterpret_cast<void*>(dst_ptr), src.data_ptr(), src.nbytes(), cudaMemcpyDeviceToDevice, cur_stream());
  TORCH_CHECK(e == cudaSuccess, "peer copy failed: ", cudaGetErrorString(e));
}
int64_t host_register(int64_t host_ptr, int64_t nbytes) {
  auto e = cudaHostRegister(reinterpret_cast<void*>(host_ptr), static_cast<size_t>(nbytes), cudaHostRegisterMapped | cudaHostRegisterPortable);
  TORCH_CHECK(e == cudaSuccess, "cudaHostRegister failed: ", cudaGetErrorString(e));
  void* dptr = nullptr;
  e = cudaHostGetDevicePointer(&dptr, reinterpret_cast<void*>(host_ptr), 0);
  TORCH_CHECK(e == cudaSuccess, "cudaHostGetDevicePointer failed: ", cudaGetErrorString(e));
  return reinterpret_cast<int64_t>(dptr);
}
void host_unregister(int64_t host_ptr) { cudaHostUnregister(reinterpret_cast<void*>(host_ptr)); }

// ---------------------------------------------------------------- MegaDPP mailbox
py::tuple mailbox_create(int n_slots, int64_t slot_bytes, const std::string& shm_name) {
  char h[64];
  int id = b200::mailbox_create(n_slots, static_cast<size_t>(slot_bytes), shm_name.c_str(), h);
  TORCH_CHECK(id > 0, "mailbox_create failed: ", b200::mailbox_last_error());
  return py::make_tuple(id, py::bytes(h, 64));
}
void mailbox_connect(int id, py::bytes peer_handle, const std::string& peer_shm) {
  std::string s = peer_handle;
  TORCH_CHECK(s.size() == 64, "bad IPC handle");
  TORCH_CHECK(b200::mailbox_connect(id, s.data(), peer_shm.c_str()) == 0, "mailbox_connect failed: ", b200::mailbox_last_error());
}
void mailbox_put(int id, int slot, const Tensor& t, int64_t epoch) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous());
  c10::cuda::CUDAGuard guard(t.device());
  TORCH_CHECK(b200::mailbox_put(id, slot, t.data_ptr(), t.nbytes(), static_cast<uint32_t>(epoch), cur_stream()) == 0,
              "mailbox_put failed: ", b200::mailbox_last_error());
  count();
}
bool mailbox_try_get(int id, int slot, int64_t epoch) { return b200::mailbox_try_get(id, slot, static_cast<uint32_t>(epoch)); }
Tensor mailbox_slot_tensor(int id, int slot, std::vector<int64_t> shape, py::object dtype, int64_t device) {
  auto st = torch::python::detail::py_object_to_dtype(dtype);
  auto opts = torch::TensorOptions().dtype(st).device(torch::kCUDA, static_cast<int>(device));
  return torch::from_blob(b200::mailbox_slot_ptr(id, slot), shape, [](void*) {}, opts);
}
int64_t mailbox_sent(int id) { return b200::mailbox_sent(id); }
void mailbox_destroy(int id) { b200::mailbox_destroy(id); }

int64_t launch_count() { return g_launches.load(); }
void reset_launch_count() { g_launches.store(0); }

}  // namespace

// ---- MegaScan native event pool ------------------------------------------------------------------------------------------
// The tracer's hot path is two event records per scope.  Events live in a process-wide pool addressed by index, so Python
// holds plain ints; resolving a whole iteration (thousands of cudaEventElapsedTime calls) runs with the GIL released.
namespace {
struct TraceEventPool {
  std::vector<cudaEvent_t> evs;
  std::vector<int64_t> free_;
  std::mutex mu;
};
TraceEventPool& trace_pool() {
  static TraceEventPool p;
  return p;
}
}  // namespace

static int64_t trace_tick() {
  auto& p = trace_pool();
  int64_t idx;
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> g(p.mu);
    if (p.free_.empty()) {
      cudaEvent_t e;
      C10_CUDA_CHECK(cudaEventCreate(&e));
      p.evs.push_back(e);
      idx = (int64_t)p.evs.size() - 1;
    } else {
      idx = p.free_.back();
      p.free_.pop_back();
    }
    ev = p.evs[idx];
  }
  C10_CUDA_CHECK(cudaEventRecord(ev, at::cuda::getCurrentCUDAStream()));
  return idx;
}

static void trace_wait(int64_t idx) {
  cudaEvent_t ev;
  {
    auto& p = trace_pool();
    std::lock_guard<std::mutex> g(p.mu);
    ev = p.evs.at(idx);
  }
  py::gil_scoped_release nogil;
  C10_CUDA_CHECK(cudaEventSynchronize(ev));
}

// ns offsets of `idx[i]` from `base`; every event in `idx` (and `base` when `recycle_base`) returns to the pool
static std::vector<int64_t> trace_resolve(int64_t base, const std::vector<int64_t>& idx, bool recycle_base) {
  auto& p = trace_pool();
  std::vector<cudaEvent_t> evs(idx.size());
  cudaEvent_t b;
  {
    std::lock_guard<std::mutex> g(p.mu);
    b = p.evs.at(base);
    for (size_t i = 0; i < idx.size(); ++i) evs[i] = p.evs.at(idx[i]);
  }
  std::vector<int64_t> out(idx.size());
  {
    py::gil_scoped_release nogil;
    for (size_t i = 0; i < idx.size(); ++i) {
      float ms = 0.f;
      C10_CUDA_CHECK(cudaEventElapsedTime(&ms, b, evs[i]));
      out[i] = (int64_t)((double)ms * 1e6);
    }
    std::lock_guard<std::mutex> g(p.mu);
    for (auto i : idx) p.free_.push_back(i);
    if (recycle_base) p.free_.push_back(base);
  }
  return out;
}

static int64_t trace_pool_size() {
  auto& p = trace_pool();
  std::lock_guard<std::mutex> g(p.mu);
  return (int64_t)p.evs.size();
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm", &gemm);
  m.def("gemm_fp8", &gemm_fp8);
  m.def("fp8_quantize", &fp8_quantize);
  m.def("gemm_ag", &gemm_ag, py::arg("A_local"), py::arg("gather_ptr"), py::arg("B"), py::arg("D"), py::arg("b_mn"),
        py::arg("bias"), py::arg("act"), py::arg("peer_bufs"), py::arg("peer_flags"), py::arg("rank"), py::arg("size"),
        py::arg("epoch"), py::arg("n_comm_ctas"), py::arg("block_n"), py::arg("D_pre") = c10::optional<Tensor>());
  m.def("gemm_grouped", &gemm_grouped);
  m.def("mx_quantize", &mx_quantize);
  m.def("gemm_mxfp8", &gemm_mxfp8);
  m.def("mx_sf_bytes", &b200::mx_sf_bytes);
  m.def("moe_router_topk_fwd", &moe_router_topk_fwd);
  m.def("moe_router_topk_bwd", &moe_router_topk_bwd);
  m.def("moe_build_layout", &moe_build_layout);
  m.def("moe_gather_rows", &moe_gather_rows);
  m.def("moe_gather_sum", &moe_gather_sum);
  m.def("moe_ep_dispatch", &moe_ep_dispatch);
  m.def("moe_ep_recv_layout", &moe_ep_recv_layout);
  m.def("moe_ep_combine_push", &moe_ep_combine_push);
  m.def("moe_ep_wait", &moe_ep_wait);
  m.def("moe_scatter_dprobs", &moe_scatter_dprobs);
  m.def("gemm_rs", &gemm_rs);
  m.def("gemm_ag_wgrad", &gemm_ag_wgrad);
  m.def("gemm_flags_needed", &b200::gemm_flags_needed);
  m.def("layernorm_fwd", &layernorm_fwd);
  m.def("layernorm_bwd", &layernorm_bwd);
  m.def("bias_gelu_fwd", &bias_gelu_fwd);
  m.def("bias_gelu_bwd", &bias_gelu_bwd);
  m.def("swiglu_fwd", &swiglu_fwd);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("bias_dropout_add", &bias_dropout_add);
  m.def("dropout_bwd", &dropout_bwd);
  m.def("rope_apply", &rope_apply);
  m.def("bias_grad", &bias_grad);
  m.def("ce_local_stats", &ce_local_stats);
  m.def("ce_rescale", &ce_rescale);
  m.def("ce_grad", &ce_grad);
  m.def("adam_step", &adam_step);
  m.def("l2norm_sq", &l2norm_sq);
  m.def("scale_", &scale_);
  m.def("cast_copy", &cast_copy);
  m.def("add_into_f32", &add_into_f32);
  m.def("check_finite", &check_finite);
  m.def("chunk_mean", &chunk_mean);
  m.def("attn_fwd", &attn_fwd);
  m.def("attn_bwd", &attn_bwd);
  m.def("arena_alloc", &arena_alloc);
  m.def("arena_open", &arena_open);
  m.def("arena_close", &arena_close);
  m.def("arena_free", &arena_free);
  m.def("tensor_from_ptr", &tensor_from_ptr);
  m.def("flags_barrier", &flags_barrier);
  m.def("bucket_reduce_scatter", &bucket_reduce_scatter);
  m.def("shard_all_gather_cast", &shard_all_gather_cast);
  m.def("a2a_scatter_rows", &a2a_scatter_rows);
  m.def("embed_scatter", &embed_scatter);
  m.def("paged_attention", &paged_attention);
  m.def("paged_attention_supported", &paged_attention_supported);
  m.def("spin", &spin);
  m.def("write_flag", &write_flag);
  m.def("peer_copy", &peer_copy);
  m.def("host_register", &host_register);
  m.def("host_unregister", &host_unregister);
  m.def("mailbox_create", &mailbox_create);
  m.def("mailbox_connect", &mailbox_connect);
  m.def("mailbox_put", &mailbox_put);
  m.def("mailbox_try_get", &mailbox_try_get);
  m.def("mailbox_slot_tensor", &mailbox_slot_tensor);
  m.def("mailbox_sent", &mailbox_sent);
  m.def("mailbox_destroy", &mailbox_destroy);
  m.def("trace_tick", &trace_tick);
  m.def("trace_wait", &trace_wait);
  m.def("trace_resolve", &trace_resolve);
  m.def("trace_pool_size", &trace_pool_size);
  m.def("launch_count", &launch_count);
  m.def("reset_launch_count", &reset_launch_count);
}
