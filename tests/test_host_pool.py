"""The persistent host thread pool (sqrtlm-slam_b200/host/host_pool.h) that set_problem's preprocessing and the map
adapter share: a stress run in a small C++ program -- thousands of back-to-back loops with changing task and thread
counts, every task executed exactly once, helpers reused, no deadlock (the run is bounded by a timeout)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "host_pool.h"
int main() {
  sqrtba::HostPool pool;
  unsigned s = 12345;
  auto rnd = [&](unsigned m) { s = s * 1664525u + 1013904223u; return (s >> 8) % m; };
  long long loops = 0, tasks = 0;
  for (int it = 0; it < 4000; it++) {
    const int n = (int)rnd(40) + (it % 97 == 0 ? 2000 : 0);
    const int thr = 1 + (int)rnd(12);
    std::vector<std::atomic<int>> hit(n > 0 ? n : 1);
    for (auto& h : hit) h.store(0);
    std::atomic<long long> sum(0);
    pool.run(n, thr, [&](int t) { hit[t].fetch_add(1); sum.fetch_add(t); });
    for (int t = 0; t < n; t++) if (hit[t].load() != 1) { std::printf("task %d ran %d times (loop %d)\n", t, hit[t].load(), it); return 1; }
    if (sum.load() != (long long)n * (n - 1) / 2) { std::printf("bad sum in loop %d\n", it); return 1; }
    loops++; tasks += n;
  }
  // the range / chunk helpers cover [0, n) exactly once, in order-independent pieces
  for (int it = 0; it < 300; it++) {
    const size_t n = rnd(5000);
    std::vector<int> cover(n, 0);
    pool.ranges(n, 1 + (int)rnd(9), [&](int, size_t a, size_t b) { for (size_t i = a; i < b; i++) cover[i]++; });
    pool.chunks((long long)n, 1 + (long long)rnd(700), 1 + (int)rnd(9), [&](long long a, long long b) { for (long long i = a; i < b; i++) cover[i]++; });
    for (size_t i = 0; i < n; i++) if (cover[i] != 2) { std::printf("element %zu covered %d times\n", i, cover[i]); return 1; }
  }
  std::printf("ok %lld loops %lld tasks\n", loops, tasks);
  return 0;
}
'''


def test_host_pool_stress(tmp_path):
    src = tmp_path / "pool.cc"
    src.write_text(SRC)
    exe = tmp_path / "pool"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-Werror", "-I",
                    os.path.join(ROOT, "sqrtlm-slam_b200", "host"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=120).stdout
    assert out.startswith("ok 4000 loops"), out
