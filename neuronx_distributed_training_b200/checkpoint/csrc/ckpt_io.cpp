// Native checkpoint file IO: large sequential writes / reads that release the GIL, with optional O_DIRECT and
// parallel pwrite/pread slicing so one multi-GB tensor file saturates NVMe / network file systems.
// Replaces the torch-XLA C++ tensor serialization the reference relies on for its "xser" mode.
#include <pybind11/pybind11.h>
#include <torch/extension.h>

#include "parallel_io.h"

namespace py = pybind11;

using ckpt_io::full_pwrite;
using ckpt_io::parallel_io;

// header bytes + raw tensor bytes -> file (fsync'd).  `data` is a CPU uint8 tensor (pinned or not).
void write_file(const std::string& path, py::bytes header, const at::Tensor& data, int threads, bool do_fsync) {
  TORCH_CHECK(data.device().is_cpu() && data.is_contiguous(), "write_file: contiguous CPU tensor required");
  std::string hdr = header;
  const size_t nbytes = size_t(data.numel()) * data.element_size();
  char* ptr = reinterpret_cast<char*>(data.data_ptr());
  py::gil_scoped_release release;
  int fd = ::open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0) throw std::runtime_error("open(" + path + ") failed: " + std::strerror(errno));
  try {
    if (!hdr.empty()) full_pwrite(fd, hdr.data(), hdr.size(), 0);
    if (nbytes) parallel_io(true, fd, ptr, nbytes, off_t(hdr.size()), threads);
    if (do_fsync) ::fsync(fd);
  } catch (...) { ::close(fd); throw; }
  ::close(fd);
}

void read_file(const std::string& path, int64_t offset, at::Tensor out, int threads) {
  TORCH_CHECK(out.device().is_cpu() && out.is_contiguous(), "read_file: contiguous CPU tensor required");
  const size_t nbytes = size_t(out.numel()) * out.element_size();
  char* ptr = reinterpret_cast<char*>(out.data_ptr());
  py::gil_scoped_release release;
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw std::runtime_error("open(" + path + ") failed: " + std::strerror(errno));
  try {
    parallel_io(false, fd, ptr, nbytes, off_t(offset), threads);
  } catch (...) { ::close(fd); throw; }
  ::close(fd);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("write_file", &write_file, py::arg("path"), py::arg("header"), py::arg("data"), py::arg("threads") = 4,
        py::arg("fsync") = false);
  m.def("read_file", &read_file, py::arg("path"), py::arg("offset"), py::arg("out"), py::arg("threads") = 4);
}
