// Dataset index builders (pure CPU, pybind11).  Built at install time by build.py into
// megatronapp_b200/data/_helpers.so -- no runtime `make` on rank 0 as in the reference
// (datasets/utils.py:21-31).  Same entry points as reference core/datasets/helpers.cpp:
// build_sample_idx (GPT doc -> fixed-length sample packing), build_blending_indices,
// build_exhaustive_blending_indices, build_mapping (BERT sentence spans), build_blocks_mapping.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

namespace py = pybind11;

template <typename DocIdx>
py::array_t<int64_t> build_sample_idx_t(const py::array_t<int32_t>& sizes_, const py::array_t<DocIdx>& doc_idx_,
                                        int32_t seq_length, int32_t num_epochs, int64_t tokens_per_epoch,
                                        bool drop_last_partial_sequence, int add_extra_token) {
  if (seq_length <= 1 || num_epochs <= 0 || tokens_per_epoch <= 1) throw std::invalid_argument("bad arguments");
  auto sizes = sizes_.unchecked<1>();
  auto doc_idx = doc_idx_.template unchecked<1>();
  const int64_t total = static_cast<int64_t>(num_epochs) * tokens_per_epoch - add_extra_token;
  int64_t num_samples = drop_last_partial_sequence ? total / seq_length
                                                   : static_cast<int64_t>(std::ceil(static_cast<double>(total) / seq_length));
  py::array_t<int64_t> out({num_samples + 1, static_cast<int64_t>(2)});
  auto s = out.mutable_unchecked<2>();
  int64_t di = 0;        // position in doc_idx
  int64_t off = 0;       // token offset inside the current document
  const int64_t n_docs = doc_idx.shape(0);
  s(0, 0) = 0; s(0, 1) = 0;
  for (int64_t i = 1; i <= num_samples; ++i) {
    int64_t remaining = seq_length + add_extra_token;
    while (remaining != 0) {
      const int64_t doc_len = static_cast<int64_t>(sizes(doc_idx(di))) - off;
      remaining -= doc_len;
      if (remaining <= 0) {
        // sample ends inside this document; the next one re-reads the last token (label shift)
        off += remaining + doc_len - add_extra_token;
        remaining = 0;
      } else {
        if (di == n_docs - 1) {  // ran out of documents: partial last sample
          off = sizes(doc_idx(di)) - add_extra_token;
          break;
        }
        ++di;
        off = 0;
      }
    }
    s(i, 0) = di;
    s(i, 1) = off;
  }
  return out;
}

void build_blending_indices(py::array_t<int16_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index,
                            const py::array_t<double>& weights, int32_t num_datasets, int64_t size, bool verbose) {
  auto di = dataset_index.mutable_unchecked<1>();
  auto dsi = dataset_sample_index.mutable_unchecked<1>();
  auto w = weights.unchecked<1>();
  std::vector<int64_t> cur(num_datasets, 0);
  for (int64_t i = 0; i < size; ++i) {
    const double denom = std::max(static_cast<double>(i), 1.0);
    int best = 0;
    double best_err = w(0) * denom - static_cast<double>(cur[0]);
    for (int d = 1; d < num_datasets; ++d) {
      const double err = w(d) * denom - static_cast<double>(cur[d]);
      if (err > best_err) { best_err = err; best = d; }
    }
    di(i) = static_cast<int16_t>(best);
    dsi(i) = cur[best];
    ++cur[best];
  }
  if (verbose) {
    for (int d = 0; d < num_datasets; ++d)
      py::print("dataset", d, "weight", w(d), "achieved", static_cast<double>(cur[d]) / static_cast<double>(size));
  }
}

void build_exhaustive_blending_indices(py::array_t<int16_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index,
                                       const py::array_t<int64_t>& sizes, int32_t num_datasets) {
  auto di = dataset_index.mutable_unchecked<1>();
  auto dsi = dataset_sample_index.mutable_unchecked<1>();
  auto sz = sizes.unchecked<1>();
  int64_t total = 0;
  for (int d = 0; d < num_datasets; ++d) total += sz(d);
  std::vector<double> w(num_datasets);
  for (int d = 0; d < num_datasets; ++d) w[d] = static_cast<double>(sz(d)) / static_cast<double>(total);
  std::vector<int64_t> cur(num_datasets, 0);
  std::vector<bool> done(num_datasets, false);
  for (int64_t i = 0; i < total; ++i) {
    const double denom = std::max(static_cast<double>(i), 1.0);
    int best = -1;
    double best_err = -std::numeric_limits<double>::infinity();
    for (int d = 0; d < num_datasets; ++d) {
      if (done[d]) continue;
      const double err = w[d] * denom - static_cast<double>(cur[d]);
      if (err > best_err) { best_err = err; best = d; }
    }
    di(i) = static_cast<int16_t>(best);
    dsi(i) = cur[best];
    if (++cur[best] == sz(best)) done[best] = true;
  }
}

// BERT-style sentence-span mapping: rows of (first sentence, one-past-last sentence, target length)
py::array build_mapping(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, int32_t num_epochs,
                        uint64_t max_num_samples, int32_t max_seq_length, double short_seq_prob, int32_t seed,
                        bool verbose, int32_t min_num_sent) {
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  std::vector<int64_t> rows;
  const int32_t short_ratio = short_seq_prob > 0 ? static_cast<int32_t>(std::round(1.0 / short_seq_prob)) : 0;
  std::mt19937 rng(seed);
  auto target_len = [&](void) -> int32_t {
    if (short_ratio == 0) return max_seq_length;
    const uint32_t r = rng();
    if ((r % short_ratio) == 0) return 2 + static_cast<int32_t>(r % (max_seq_length - 1));
    return max_seq_length;
  };
  uint64_t n = 0;
  for (int32_t epoch = 0; epoch < num_epochs && n < max_num_samples; ++epoch) {
    for (int64_t d = 0; d + 1 < docs.shape(0) && n < max_num_samples; ++d) {
      const int64_t first = docs(d), last = docs(d + 1);
      if (last - first < min_num_sent) continue;
      bool has_long = false;
      for (int64_t sidx = first; sidx < last; ++sidx)
        if (sizes(sidx) > 512) { has_long = true; break; }   // same filter constant as upstream
      if (has_long) continue;
      int64_t start = first;
      int32_t seq_len = 0, num_sent = 0;
      int32_t tgt = target_len();
      for (int64_t sidx = first; sidx < last; ++sidx) {
        seq_len += sizes(sidx);
        ++num_sent;
        const int64_t remain = last - sidx - 1;
        if ((seq_len >= tgt && remain >= min_num_sent && num_sent >= min_num_sent) || remain == 0) {
          if (num_sent >= min_num_sent) {
            rows.push_back(start); rows.push_back(sidx + 1); rows.push_back(tgt);
            if (++n >= max_num_samples) break;
          }
          start = sidx + 1;
          seq_len = 0; num_sent = 0;
          tgt = target_len();
        }
      }
    }
  }
  const int64_t count = static_cast<int64_t>(rows.size() / 3);
  // deterministic shuffle of rows
  std::mt19937_64 rng64(seed + 1);
  for (int64_t i = count - 1; i > 0; --i) {
    const int64_t j = static_cast<int64_t>(rng64() % static_cast<uint64_t>(i + 1));
    for (int c = 0; c < 3; ++c) std::swap(rows[3 * i + c], rows[3 * j + c]);
  }
  py::array_t<int64_t> out({count, static_cast<int64_t>(3)});
  std::copy(rows.begin(), rows.end(), out.mutable_data());
  if (verbose) py::print("build_mapping:", count, "samples");
  return std::move(out);
}

// ICT / REALM blocks: rows of (first sentence, one-past-last, document, block id)
py::array build_blocks_mapping(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_,
                               const py::array_t<int32_t>& titles_sizes_, int32_t num_epochs,
                               uint64_t max_num_samples, int32_t max_seq_length, int32_t seed, bool verbose,
                               bool use_one_sent_blocks) {
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  auto titles = titles_sizes_.unchecked<1>();
  std::vector<int64_t> rows;
  const int32_t min_sent = use_one_sent_blocks ? 1 : 2;
  uint64_t n = 0;
  for (int32_t epoch = 0; epoch < num_epochs && n < max_num_samples; ++epoch) {
    int64_t block_id = 0;
    for (int64_t d = 0; d + 1 < docs.shape(0) && n < max_num_samples; ++d) {
      const int64_t first = docs(d), last = docs(d + 1);
      const int32_t budget = max_seq_length - titles(d);
      if (last - first < min_sent) continue;
      int64_t start = first;
      int32_t seq_len = 0, num_sent = 0;
      for (int64_t sidx = first; sidx < last; ++sidx) {
        seq_len += sizes(sidx);
        ++num_sent;
        const int64_t remain = last - sidx - 1;
        if ((seq_len >= budget && remain >= min_sent && num_sent >= min_sent) || remain == 0) {
          rows.push_back(start); rows.push_back(sidx + 1); rows.push_back(d); rows.push_back(block_id++);
          if (++n >= max_num_samples) break;
          start = sidx + 1; seq_len = 0; num_sent = 0;
        }
      }
    }
  }
  const int64_t count = static_cast<int64_t>(rows.size() / 4);
  std::mt19937_64 rng64(seed + 1);
  for (int64_t i = count - 1; i > 0; --i) {
    const int64_t j = static_cast<int64_t>(rng64() % static_cast<uint64_t>(i + 1));
    for (int c = 0; c < 4; ++c) std::swap(rows[4 * i + c], rows[4 * j + c]);
  }
  py::array_t<int64_t> out({count, static_cast<int64_t>(4)});
  std::copy(rows.begin(), rows.end(), out.mutable_data());
  if (verbose) py::print("build_blocks_mapping:", count, "blocks");
  return std::move(out);
}

PYBIND11_MODULE(_helpers, m) {
  m.def("build_sample_idx_int32", &build_sample_idx_t<int32_t>);
  m.def("build_sample_idx_int64", &build_sample_idx_t<int64_t>);
  m.def("build_blending_indices", &build_blending_indices);
  m.def("build_exhaustive_blending_indices", &build_exhaustive_blending_indices);
  m.def("build_mapping", &build_mapping);
  m.def("build_blocks_mapping", &build_blocks_mapping);
}
