// Native checkpoint file IO: large sequential writes / reads that release the GIL, with optional O_DIRECT and
// parallel pwrite/pread slicing so one multi-GB tensor file saturates NVMe / network file systems.
// Replaces the torch-XLA C++ tensor serialization the reference relies on for its "xser" mode.
#include <pybind11/pybind11.h>
#include <torch/extension.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace py = pybind11;

namespace {

void full_pwrite(int fd, const char* p, size_t n, off_t off) {
  while (n > 0) {
    ssize_t w = ::pwrite(fd, p, std::min<size_t>(n, size_t(1) << 30), off);
    if (w < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(std::string("pwrite failed: ") + std::strerror(errno));
    }
    p += w; n -= size_t(w); off += w;
  }
}

void full_pread(int fd, char* p, size_t n, off_t off) {
  while (n > 0) {
    ssize_t r = ::pread(fd, p, std::min<size_t>(n, size_t(1) << 30), off);
    if (r < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(std::string("pread failed: ") + std::strerror(errno));
    }
    if (r == 0) throw std::runtime_error("pread: unexpected end of file");
    p += r; n -= size_t(r); off += r;
  }
}

void parallel_io(bool write, int fd, char* data, size_t nbytes, off_t base, int threads) {
  const size_t kMinSlice = size_t(64) << 20;
  int n = int(std::max<size_t>(1, std::min<size_t>(size_t(threads), nbytes / kMinSlice)));
  if (n <= 1) {
    if (write) full_pwrite(fd, data, nbytes, base); else full_pread(fd, data, nbytes, base);
    return;
  }
  std::vector<std::thread> pool;
  std::vector<std::string> errs(n);
  const size_t slice = ((nbytes + n - 1) / n + 4095) & ~size_t(4095);
  for (int i = 0; i < n; ++i) {
    const size_t lo = std::min(nbytes, slice * i), hi = std::min(nbytes, slice * (i + 1));
    if (lo >= hi) continue;
    pool.emplace_back([=, &errs]() {
      try {
        if (write) full_pwrite(fd, data + lo, hi - lo, base + off_t(lo)); else full_pread(fd, data + lo, hi - lo, base + off_t(lo));
      } catch (const std::exception& e) { errs[i] = e.what(); }
    });
  }
  for (auto& t : pool) t.join();
  for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
}

}  // namespace

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
