// ThreadSanitizer harness for the native checkpoint IO core (checkpoint/csrc/parallel_io.h).
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -I neuronx_distributed_training_b200/checkpoint/csrc \
//       tests/native/tsan_parallel_io.cpp -o /tmp/tsan_parallel_io && /tmp/tsan_parallel_io /tmp/scratch.bin
// Writes a buffer with 1..8 threads (tiny slices so every thread is used), reads it back with a different thread
// count, checks the bytes, and exercises the error path (reading past the end of the file from several threads).
#include "parallel_io.h"

#include <cstdint>
#include <cstdio>

int main(int argc, char** argv) {
  const std::string path = argc > 1 ? argv[1] : "/tmp/tsan_parallel_io.bin";
  const size_t n = (size_t(3) << 20) + 12345;             // not a multiple of the 4 KiB slice alignment
  std::vector<char> src(n), dst(n);
  uint32_t x = 12345u;
  for (size_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; src[i] = char(x >> 24); }
  for (int wt = 1; wt <= 8; wt *= 2) {
    for (int rt = 1; rt <= 8; rt *= 2) {
      int fd = ::open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
      if (fd < 0) { std::perror("open"); return 2; }
      const char hdr[7] = {'h', 'e', 'a', 'd', 'e', 'r', '!'};
      ckpt_io::full_pwrite(fd, hdr, sizeof(hdr), 0);
      ckpt_io::parallel_io(true, fd, src.data(), n, off_t(sizeof(hdr)), wt, /*min_slice=*/4096);
      std::fill(dst.begin(), dst.end(), 0);
      ckpt_io::parallel_io(false, fd, dst.data(), n, off_t(sizeof(hdr)), rt, 4096);
      if (std::memcmp(src.data(), dst.data(), n) != 0) { std::fprintf(stderr, "mismatch wt=%d rt=%d\n", wt, rt); return 1; }
      bool threw = false;
      try {
        ckpt_io::parallel_io(false, fd, dst.data(), n, off_t(sizeof(hdr) + 4096), rt, 4096);   // runs past EOF
      } catch (const std::exception&) { threw = true; }
      ::close(fd);
      if (!threw) { std::fprintf(stderr, "short read was not reported (rt=%d)\n", rt); return 1; }
    }
  }
  ::unlink(path.c_str());
  std::puts("TSAN_PARALLEL_IO_OK");
  return 0;
}
