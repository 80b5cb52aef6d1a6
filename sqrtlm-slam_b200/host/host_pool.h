// A small persistent pool of host helper threads, shared by sqrtba_set_problem's preprocessing (csrc/sqrtba_solver.cu)
// and the map adapter (host/sqrtbaOptimizer.cc).  Both run a handful of short parallel loops per bundle-adjustment call;
// creating threads for each loop costs more than the loops themselves on a 70 k-observation window (and far more inside
// sandboxed containers, where thread creation takes ~1 ms), so the helpers are created once, on demand, and parked on a
// condition variable between loops.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace sqrtba {

class HostPool {
 public:
  HostPool() = default;
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_start_.notify_all();
    for (auto& t : th_) t.join();
  }

  // fn(task) for every task in [0, n_tasks), on the calling thread plus up to max_threads - 1 helpers; tasks are handed
  // out dynamically (an atomic counter), so uneven tasks balance.  Returns when every task has finished.  One loop at a
  // time per pool (callers own their pool or serialise).
  void run(int n_tasks, int max_threads, const std::function<void(int)>& fn) {
    if (n_tasks <= 0) return;
    const int helpers = std::max(0, std::min(max_threads, n_tasks) - 1);
    if (helpers == 0) {
      for (int t = 0; t < n_tasks; t++) fn(t);
      return;
    }
    {
      std::unique_lock<std::mutex> lk(m_);
      while ((int)th_.size() < helpers) {
        const int id = (int)th_.size();
        th_.emplace_back([this, id] { worker(id); });
      }
      fn_ = &fn;
      n_tasks_ = n_tasks;
      next_.store(0, std::memory_order_relaxed);
      want_ = helpers;
      pending_ = helpers;
      gen_++;
    }
    cv_start_.notify_all();
    drain(fn, n_tasks);
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

  // fn(begin, end) over [0, n) in chunks of `chunk` elements
  template <class F>
  void chunks(long long n, long long chunk, int max_threads, F fn) {
    const long long n_chunk = (n + chunk - 1) / chunk;
    if (n_chunk <= 1 || max_threads <= 1) {
      if (n > 0) fn((long long)0, n);
      return;
    }
    run((int)n_chunk, max_threads, [&](int c) { fn((long long)c * chunk, std::min(n, ((long long)c + 1) * chunk)); });
  }

  // fn(part, begin, end) over [0, n) split into `parts` contiguous ranges (part index = position in the order)
  template <class F>
  void ranges(size_t n, int parts, F fn) {
    if (parts <= 1) {
      fn(0, (size_t)0, n);
      return;
    }
    run(parts, parts, [&](int t) { fn(t, n * (size_t)t / (size_t)parts, n * ((size_t)t + 1) / (size_t)parts); });
  }

 private:
  void drain(const std::function<void(int)>& fn, int n_tasks) {
    for (;;) {
      const int t = next_.fetch_add(1, std::memory_order_relaxed);
      if (t >= n_tasks) break;
      fn(t);
    }
  }
  void worker(int id) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      int n_tasks;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_start_.wait(lk, [&] { return stop_ || (gen_ != seen && id < want_); });
        if (stop_) return;
        seen = gen_;
        fn = fn_;
        n_tasks = n_tasks_;
      }
      drain(*fn, n_tasks);
      {
        std::lock_guard<std::mutex> lk(m_);
        pending_--;
      }
      cv_done_.notify_one();
    }
  }

  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_tasks_ = 0, want_ = 0, pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

}  // namespace sqrtba
