// gemm_multi_b200.cpp -- one SGEMM over all B200s of the box through blas::multi (include/interface/blas3_interface_multi.h):
// BASELINE configs[2] "SGEMM 16384^3 fp32 M-block sharded across 1/2/4/8 B200" from a plain C++ caller.
//
//   gemm_multi_b200 [n = 4096] [devices = all] [repeat-device = 0]
//
// 1. blas::multi::_gemm_host: host operands in, host C out (upload of the M-blocks of A and of 1/G of B per device, NVLink
//    exchange of the B panels, compute, download), checked against a double-precision host product on sampled entries;
// 2. blas::multi::_gemm on device-resident shards with the fused gather: every device ends with the WHOLE C; device 0's
//    and the last device's copies are downloaded and compared, and the step is timed.
// "repeat-device 1" builds the group from `devices` shards that all live on device 0 (the partition, the peer stores
// and the exchange then run on a one-GPU box).  Exit code 0 == all checks passed.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "interface/blas3_interface_multi.h"

int main(int argc, char** argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 4096;
  const int want_dev = argc > 2 ? std::atoi(argv[2]) : 0;
  const bool repeat = argc > 3 && std::atoi(argv[3]) != 0;
  std::vector<int> ordinals(want_dev > 0 ? want_dev : 1, 0);
  blas::multi::SB_Handle_Group group(repeat ? want_dev : want_dev, repeat ? ordinals.data() : nullptr);
  const int G = group.device_count();
  std::printf("group of %d device shard(s), SGEMM %d x %d x %d NN\n", G, n, n, n);

  const size_t nn = (size_t)n * n;
  std::vector<float> a(nn), b(nn), c(nn, 0.0f);
  std::mt19937 gen(12345);
  std::uniform_real_distribution<float> dis(-2.0f, 5.0f);
  for (auto& x : a) x = dis(gen);
  for (auto& x : b) x = dis(gen);
  bool ok = true;
  auto check = [&](const std::vector<float>& got, const char* what) {
    std::mt19937 pick(7);
    double worst = 0;
    for (int s = 0; s < 256; ++s) {
      const int i = (s == 0) ? 0 : (s == 1 ? n - 1 : (int)(pick() % n)), j = (s < 2) ? (s ? n - 1 : 0) : (int)(pick() % n);
      double acc = 0, bound = 0;
      for (int l = 0; l < n; ++l) {
        const double p = (double)a[i + (size_t)l * n] * (double)b[l + (size_t)j * n];
        acc += p; bound += std::fabs(p);
      }
      worst = std::max(worst, std::fabs(1.5 * acc - got[i + (size_t)j * n]) / (1.5 * bound));
    }
    const bool good = worst <= 1e-5;   // north-star fp32 bound, relative to |alpha||A||B|
    std::printf("%-58s max err / bound %.2e  %s\n", what, worst, good ? "PASS" : "FAIL");
    ok = ok && good;
  };

  // ---- 1. host operands ----
  auto t0 = std::chrono::steady_clock::now();
  blas::multi::_gemm_host(group, 'n', 'n', n, n, n, 1.5f, a.data(), n, b.data(), n, 0.0f, c.data(), n);
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("_gemm_host (first call, pageable host memory): %.1f ms\n", dt * 1e3);
  check(c, "blas::multi::_gemm_host vs host double product");

  // ---- 2. device-resident shards, fused gather ----
  std::vector<const float*> a_blocks(G), b_full(G);
  std::vector<float*> a_dev(G), b_dev(G), c_full(G);
  for (int g = 0; g < G; ++g) {
    const auto blk = group.mblock(n, g);
    const size_t rows = (size_t)blk.second;
    a_dev[g] = group.allocate<float>(g, std::max<size_t>(rows * n, 1));
    b_dev[g] = group.allocate<float>(g, nn);
    c_full[g] = group.allocate<float>(g, nn);
    // the M-block of A, compacted to leading dimension `rows` on the host side of the copy
    std::vector<float> a_blk(rows * n);
    for (int l = 0; l < n; ++l)
      for (size_t r = 0; r < rows; ++r) a_blk[r + (size_t)l * rows] = a[blk.first + r + (size_t)l * n];
    group.copy_to_device(g, a_blk.data(), a_dev[g], rows * n);
    group.copy_to_device(g, b.data(), b_dev[g], nn);
    group.wait();
    a_blocks[g] = a_dev[g]; b_full[g] = b_dev[g];
  }
  // lda differs per device when the blocks are compact: equal blocks here (n is a multiple of 256 * G) keep one lda
  const auto blk0 = group.mblock(n, 0);
  bool equal = true;
  for (int g = 0; g < G; ++g) equal = equal && group.mblock(n, g).second == blk0.second;
  if (!equal) { std::printf("n must be a multiple of 256 * devices for the resident-shard part\n"); return 2; }
  const int lda = (int)blk0.second;
  for (int it = 0; it < 3; ++it)   // warm-up
    blas::multi::_gemm(group, 'n', 'n', n, n, n, 1.5f, a_blocks, lda, b_full, n, 0.0f, c_full, n, true);
  group.wait();
  const int steps = 10;
  t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < steps; ++it)
    blas::multi::_gemm(group, 'n', 'n', n, n, n, 1.5f, a_blocks, lda, b_full, n, 0.0f, c_full, n, true);
  group.wait();
  dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;
  std::printf("blas::multi::_gemm (resident shards, fused gather): %.3f ms per step, %.1f TFLOP/s\n", dt * 1e3,
              2.0 * n * (double)n * n / dt / 1e12);
  std::vector<float> c0(nn), c1(nn);
  group.copy_to_host(0, c_full[0], c0.data(), nn);
  group.copy_to_host(G - 1, c_full[G - 1], c1.data(), nn);
  group.wait();
  check(c0, "device 0's full C after the fused gather");
  bool same = true;
  for (size_t i = 0; i < nn && same; ++i) same = (c0[i] == c1[i]);
  std::printf("%-58s %s\n", "first and last device hold bitwise identical C", same ? "PASS" : "FAIL");
  ok = ok && same;
  for (int g = 0; g < G; ++g) { group.deallocate(g, a_dev[g]); group.deallocate(g, b_dev[g]); group.deallocate(g, c_full[g]); }
  std::printf(ok ? "ALL PASS\n" : "FAILED\n");
  return ok ? 0 : 1;
}
