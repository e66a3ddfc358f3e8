// Torch-free core of the native checkpoint IO: full pread / pwrite loops and the sliced multi-threaded transfer.
// Kept separate from the Python binding (ckpt_io.cpp) so that tests/native/tsan_parallel_io.cpp can build it with
// -fsanitize=thread without Python / torch in the process.
#pragma once

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

namespace ckpt_io {

inline void full_pwrite(int fd, const char* p, size_t n, off_t off) {
  while (n > 0) {
    ssize_t w = ::pwrite(fd, p, std::min<size_t>(n, size_t(1) << 30), off);
    if (w < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error(std::string("pwrite failed: ") + std::strerror(errno));
    }
    p += w; n -= size_t(w); off += w;
  }
}

inline void full_pread(int fd, char* p, size_t n, off_t off) {
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

// `min_slice`: a transfer is only split when every thread gets at least this many bytes
inline void parallel_io(bool write, int fd, char* data, size_t nbytes, off_t base, int threads,
                        size_t min_slice = size_t(64) << 20) {
  const size_t kMinSlice = std::max<size_t>(min_slice, 4096);
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

}  // namespace ckpt_io
