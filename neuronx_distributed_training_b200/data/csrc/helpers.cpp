// Index builders for the Megatron-style GPT dataset (pybind11).
//
//   build_sample_idx      : for fixed-length samples of `seq_length` (+1 label token) drawn from a stream of
//                           documents (in `doc_idx` order, repeated over epochs), record for every sample the
//                           (index into doc_idx, token offset inside that document) where it starts.  O(#docs) scan.
//   build_blending_indices: greedy error-minimising interleave of several datasets to hit target weights.
//   build_shuffle_idx / build_doc_idx helpers live in Python (numpy RNG semantics).
//
// Same capability as the compiled helpers the reference's dataset code imports
// (reference datasets/gpt_dataset_patch.py:516-530); written from the algorithm description, OpenMP-free so it builds
// anywhere, with 64-bit indices throughout.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace py = pybind11;

py::array_t<int64_t> build_sample_idx(const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& sizes_,
                                      const py::array_t<int64_t, py::array::c_style | py::array::forcecast>& doc_idx_,
                                      int64_t seq_length, int64_t num_epochs, int64_t tokens_per_epoch,
                                      bool drop_last, int64_t add_extra_token) {
  if (seq_length <= 1) throw std::invalid_argument("seq_length must be > 1");
  auto sizes = sizes_.unchecked<1>();
  auto doc_idx = doc_idx_.unchecked<1>();
  const int64_t n_doc_idx = doc_idx.shape(0);
  int64_t num_samples;
  const int64_t total = num_epochs * tokens_per_epoch - add_extra_token;
  num_samples = drop_last ? total / seq_length : (total + seq_length - 1) / seq_length;
  if (num_samples < 0) num_samples = 0;
  py::array_t<int64_t> out({num_samples + 1, (int64_t)2});
  auto o = out.mutable_unchecked<2>();
  int64_t sample = 0, di = 0, offset = 0;
  o(0, 0) = 0;
  o(0, 1) = 0;
  ++sample;
  while (sample <= num_samples) {
    int64_t remaining = seq_length + add_extra_token;
    while (remaining != 0) {
      if (di >= n_doc_idx) {           // ran out of tokens (only possible when !drop_last): clamp to the end
        di = n_doc_idx - 1;
        offset = sizes(doc_idx(di)) - add_extra_token;
        break;
      }
      const int64_t doc_len = (int64_t)sizes(doc_idx(di)) - offset;
      remaining -= doc_len;
      if (remaining <= 0) {
        // the sample ends inside this document; the next one re-reads the last token as its first input
        offset += remaining + doc_len - add_extra_token;
        remaining = 0;
      } else {
        ++di;
        offset = 0;
      }
    }
    o(sample, 0) = di;
    o(sample, 1) = offset;
    ++sample;
  }
  return out;
}

void build_blending_indices(py::array_t<uint8_t>& dataset_index_, py::array_t<int64_t>& dataset_sample_index_,
                            const py::array_t<double, py::array::c_style | py::array::forcecast>& weights_,
                            int32_t num_datasets, int64_t size, bool verbose) {
  auto dataset_index = dataset_index_.mutable_unchecked<1>();
  auto dataset_sample_index = dataset_sample_index_.mutable_unchecked<1>();
  auto weights = weights_.unchecked<1>();
  std::vector<int64_t> current(num_datasets, 0);
  for (int64_t i = 0; i < size; ++i) {
    const double denom = std::max<double>(static_cast<double>(i), 1.0);
    int64_t best = 0;
    double best_err = weights(0) * denom - static_cast<double>(current[0]);
    for (int32_t d = 1; d < num_datasets; ++d) {
      const double err = weights(d) * denom - static_cast<double>(current[d]);
      if (err > best_err) {
        best_err = err;
        best = d;
      }
    }
    dataset_index(i) = static_cast<uint8_t>(best);
    dataset_sample_index(i) = current[best];
    ++current[best];
  }
  (void)verbose;
}

PYBIND11_MODULE(data_helpers, m) {
  m.def("build_sample_idx", &build_sample_idx, py::arg("sizes"), py::arg("doc_idx"), py::arg("seq_length"),
        py::arg("num_epochs"), py::arg("tokens_per_epoch"), py::arg("drop_last") = true, py::arg("add_extra_token") = 1);
  m.def("build_blending_indices", &build_blending_indices);
}
