// sb_handle/portblas_handle.h -- blas::SB_Handle for the B200 build.
// Public surface of reference include/sb_handle/portblas_handle.h:46-200 (ctor from sycl::queue or
// Temp_Mem_Pool*, get_queue, wait overloads, get_num_compute_units, get_work_group_size,
// has_local_memory, event_t).  The expression-tree execute() overloads are gone: GEMM is launched
// through the C-ABI (include/pbx_gemm.h) on the queue's CUDA stream, and the split-K temporary
// memory the reference takes from Temp_Mem_Pool (include/sb_handle/temp_memory_pool.h:33-114)
// is pooled inside the pbx handle.
#pragma once
#include <sycl/sycl.hpp>

#include <vector>

#include "../blas_meta.h"

namespace blas {

class Temp_Mem_Pool {
 public:
  explicit Temp_Mem_Pool(sycl::queue q) : q_(q) {}
  sycl::queue get_queue() const { return q_; }
 private:
  sycl::queue q_;
};

class SB_Handle {
 public:
  using event_t = std::vector<sycl::event>;

  explicit SB_Handle(sycl::queue q) : q_(q), tempMemPool_(nullptr) { cache(); }
  explicit SB_Handle(Temp_Mem_Pool* pool) : q_(pool->get_queue()), tempMemPool_(pool) { cache(); }

  sycl::queue get_queue() const { return q_; }
  pbx_handle_t pbx() const { return q_.pbx(); }

  void wait() { q_.wait(); }
  void wait(sycl::event evs) { evs.wait(); }
  void wait(std::vector<sycl::event> evs) { sycl::event::wait(evs); }
  template <typename first_event_t, typename... next_events_t>
  void wait(first_event_t first, next_events_t... rest) {
    wait(first);
    wait(rest...);
  }

  bool has_local_memory() const { return true; }
  size_t get_work_group_size() const { return workGroupSize_; }
  size_t get_num_compute_units() const { return computeUnits_; }

 private:
  void cache() {
    workGroupSize_ = 256;  // the reference caps it at 256 (include/portblas_helper.h:121-126)
    computeUnits_ = static_cast<size_t>(pbx_get_num_compute_units(q_.pbx()));
  }
  sycl::queue q_;
  Temp_Mem_Pool* tempMemPool_;
  size_t workGroupSize_ = 0;
  size_t computeUnits_ = 0;
};

}  // namespace blas
