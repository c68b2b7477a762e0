// Host-side check of the division-free tile decode (vpd_b200/csrc/conv_igemm.cuh: fd_make /
// fd_div): every divisor a launch can have against every dividend a tile index can be, plus
// the 32-bit boundaries. Built and run by tests/test_native_host_cpu.py with nvcc (no GPU).
#include <cstdio>
#include <cstdint>
#include "conv_igemm.cuh"

int main() {
  using namespace vpd;
  unsigned long long checked = 0;
  for (uint32_t d = 1; d <= 4096; ++d) {
    const FastDiv f = fd_make(d);
    for (uint32_t n = 0; n < 70000; ++n, ++checked)
      if (fd_div(f, n) != n / d) { std::printf("FAIL d=%u n=%u\n", d, n); return 1; }
    const uint32_t edge[] = {0x7FFFFFFFu, 0x80000000u, 0xFFFFFFFEu, 0xFFFFFFFFu, d * 65537u, d * 65537u - 1};
    for (uint32_t n : edge) { ++checked; if (fd_div(f, n) != n / d) { std::printf("FAIL d=%u n=%u\n", d, n); return 1; } }
  }
  const uint32_t big[] = {65535u, 65536u, 1000003u, 0x7FFFFFFFu, 0x80000001u, 0xFFFFFFFFu};
  for (uint32_t d : big) {
    const FastDiv f = fd_make(d);
    for (uint64_t n = 0; n <= 0xFFFFFFFFull; n += 65521, ++checked)
      if (fd_div(f, (uint32_t)n) != (uint32_t)n / d) { std::printf("FAIL d=%u n=%llu\n", d, (unsigned long long)n); return 1; }
  }
  std::printf("ok %llu\n", checked);
  return 0;
}
