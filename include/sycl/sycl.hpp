// sycl/sycl.hpp -- MINIMAL SYCL-COMPATIBILITY SHIM (not a SYCL implementation).
//
// The reference's GEMM callers (samples/gemm.cpp, benchmark/portblas/blas3/gemm*.cpp,
// test/unittest/blas3/*gemm*) are written against <sycl/sycl.hpp>.  This header maps the small
// SYCL surface those callers touch onto the CUDA stream / event /
// allocation calls of the C-ABI (include/pbx_gemm.h), so they compile unchanged against the
// B200 GEMM path.  There are no kernels, accessors-on-device or parallel_for here: device work
// happens only inside libpbx_gemm.so.
//
//   sycl::queue   -> one pbx handle (device + CUDA stream), shared by copies of the queue
//   sycl::event   -> CUDA event pair (start/end) for wait() and profiling
//   sycl::buffer  -> ref-counted device allocation (+ write-back to a host pointer on destruction)
//   malloc_device -> pbx_malloc
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include <complex>
#include <ostream>

#include "../pbx_gemm.h"

// sycl::half streams like a float (callers print mismatching elements, common/include/common/float_comparison.hpp:211)
inline std::ostream& operator<<(std::ostream& os, const __half& v) { return os << __half2float(v); }

namespace sycl {

using half = __half;

class exception : public std::runtime_error {
 public:
  explicit exception(const std::string& what) : std::runtime_error(what) {}
};
using exception_list = std::vector<std::exception_ptr>;
using async_handler = std::function<void(exception_list)>;

enum class access_mode { read, write, read_write, discard_write, discard_read_write };
namespace access { using mode = access_mode; enum class placeholder { false_t, true_t }; }
enum class target { device, host_task, global_buffer = device };
enum class aspect { fp16, fp64, gpu, cpu };
namespace usm { enum class alloc { host, device, shared, unknown }; }

namespace info {
enum class device_type { cpu, gpu, accelerator, custom, automatic, host, all };
enum class local_mem_type { none, local, global };
namespace device {
struct name { using return_type = std::string; };
struct vendor { using return_type = std::string; };
struct version { using return_type = std::string; };
struct driver_version { using return_type = std::string; };
struct device_type { using return_type = info::device_type; };
struct max_work_group_size { using return_type = size_t; };
struct max_compute_units { using return_type = uint32_t; };
struct local_mem_type { using return_type = info::local_mem_type; };
}  // namespace device
namespace platform { struct name { using return_type = std::string; }; }
namespace event_profiling {
struct command_submit { using return_type = uint64_t; };
struct command_start { using return_type = uint64_t; };
struct command_end { using return_type = uint64_t; };
}  // namespace event_profiling
}  // namespace info

namespace property { namespace queue { struct enable_profiling {}; struct in_order {}; } }
struct property_list {
  template <typename... Ts> property_list(Ts...) {}
};

class device;
// selectors are callables scoring a device (SYCL 2020): everything here is the one CUDA device of the queue
struct default_selector_t { int operator()(const device&) const { return 1; } };
struct gpu_selector_t { int operator()(const device&) const { return 1; } };
inline constexpr default_selector_t default_selector_v{};
inline constexpr gpu_selector_t gpu_selector_v{};

template <int D = 1> struct range {
  size_t v;
  range(size_t n = 0) : v(n) {}
  size_t operator[](int) const { return v; }
  size_t size() const { return v; }
};
template <int D = 1> struct id {
  size_t v;
  id(size_t n = 0) : v(n) {}
  size_t operator[](int) const { return v; }
};

namespace detail {
inline void check(pbx_handle_t h, int st, const char* what) {
  if (st != PBX_OK) {
    std::string msg = std::string(what) + ": " + pbx_status_string(st);
    if (h) msg += std::string(" (") + pbx_last_error(h) + ")";
    throw sycl::exception(msg);
  }
}
struct queue_impl {
  pbx_handle_t h = nullptr;
  void* epoch = nullptr;  // profiling time origin
  async_handler handler;
  explicit queue_impl(int device) {
    check(nullptr, pbx_create(&h, device, nullptr), "sycl::queue: pbx_create");
    check(h, pbx_event_create(h, &epoch), "pbx_event_create");
    check(h, pbx_event_record(h, epoch), "pbx_event_record");
  }
  ~queue_impl() {
    if (h) {
      pbx_synchronize(h);
      if (epoch) pbx_event_destroy(h, epoch);
      pbx_destroy(h);
    }
  }
};
struct event_impl {
  std::shared_ptr<queue_impl> q;
  void* start = nullptr;
  void* end = nullptr;
  ~event_impl() {
    if (q && q->h) {
      if (start) pbx_event_destroy(q->h, start);
      if (end) pbx_event_destroy(q->h, end);
    }
  }
};
// device-pointer registry so sycl::get_pointer_type can answer for malloc_device memory
inline std::unordered_map<const void*, size_t>& usm_registry() {
  static std::unordered_map<const void*, size_t> r;
  return r;
}
inline std::mutex& usm_mutex() { static std::mutex m; return m; }
// true when p lies inside a registered device allocation (base pointers and interior pointers alike)
inline bool usm_contains(const void* p) {
  std::lock_guard<std::mutex> g(usm_mutex());
  for (auto& kv : usm_registry()) {
    const char* b = static_cast<const char*>(kv.first);
    if (static_cast<const char*>(p) >= b && static_cast<const char*>(p) < b + kv.second) return true;
  }
  return false;
}
}  // namespace detail

class platform {
 public:
  template <typename P> typename P::return_type get_info() const { return "NVIDIA CUDA (pbx_gemm shim)"; }
};

class device {
  std::shared_ptr<detail::queue_impl> q_;
 public:
  device() = default;
  explicit device(std::shared_ptr<detail::queue_impl> q) : q_(std::move(q)) {}
  bool is_gpu() const { return true; }
  bool is_cpu() const { return false; }
  bool has(aspect a) const { return a == aspect::fp16 || a == aspect::fp64 || a == aspect::gpu; }
  platform get_platform() const { return platform(); }
  template <typename P> typename P::return_type get_info() const {
    if constexpr (std::is_same_v<P, info::device::name>) {
      char buf[256] = "unknown";
      if (q_) pbx_device_name(q_->h, buf, sizeof(buf));
      return std::string(buf);
    } else if constexpr (std::is_same_v<P, info::device::vendor>) {
      return std::string("NVIDIA Corporation");
    } else if constexpr (std::is_same_v<P, info::device::version> ||
                         std::is_same_v<P, info::device::driver_version>) {
      return std::string("CUDA sm_100a");
    } else if constexpr (std::is_same_v<P, info::device::device_type>) {
      return info::device_type::gpu;
    } else if constexpr (std::is_same_v<P, info::device::max_work_group_size>) {
      return size_t(1024);
    } else if constexpr (std::is_same_v<P, info::device::max_compute_units>) {
      return uint32_t(q_ ? pbx_get_num_compute_units(q_->h) : 0);
    } else {
      return info::local_mem_type::local;
    }
  }
};

class context {};

class event {
  std::shared_ptr<detail::event_impl> impl_;
 public:
  event() = default;
  explicit event(std::shared_ptr<detail::event_impl> i) : impl_(std::move(i)) {}
  void wait() const {
    if (impl_ && impl_->end) detail::check(impl_->q->h, pbx_event_synchronize(impl_->q->h, impl_->end), "event::wait");
  }
  void wait_and_throw() const { wait(); }
  static void wait(const std::vector<event>& evs) { for (auto& e : evs) e.wait(); }
  // command_start is placed on the queue's time line (float milliseconds since the queue's epoch: microsecond-scale
  // quantisation after minutes of process lifetime); command_end = command_start + cudaEventElapsedTime(start, end),
  // so end - start -- what the reference's benchmarks compute (benchmark/portblas/utils.hpp:52-63) -- keeps the
  // resolution of the event pair itself however long the process has run.
  template <typename P> uint64_t get_profiling_info() const {
    if (!impl_ || !impl_->end) return 0;
    wait();
    float ms = 0.f;
    detail::check(impl_->q->h, pbx_event_elapsed_ms(impl_->q->h, impl_->q->epoch, impl_->start, &ms), "event profiling");
    const uint64_t start_ns = static_cast<uint64_t>(static_cast<double>(ms) * 1.0e6);
    if constexpr (!std::is_same_v<P, info::event_profiling::command_end>) return start_ns;
    float dur = 0.f;
    detail::check(impl_->q->h, pbx_event_elapsed_ms(impl_->q->h, impl_->start, impl_->end, &dur), "event profiling");
    return start_ns + static_cast<uint64_t>(static_cast<double>(dur) * 1.0e6);
  }
  const std::shared_ptr<detail::event_impl>& impl() const { return impl_; }
};

class queue;

// Command-group handler: only the host-side operations the GEMM callers use.
class handler {
  friend class queue;
  std::shared_ptr<detail::queue_impl> q_;
  explicit handler(std::shared_ptr<detail::queue_impl> q) : q_(std::move(q)) {}
 public:
  pbx_handle_t pbx() const { return q_->h; }
  void depends_on(const event& e) {
    // same in-order stream: already ordered; a foreign queue's event needs a stream wait
    if (e.impl() && e.impl()->end && e.impl()->q != q_)
      detail::check(q_->h, pbx_stream_wait_event(q_->h, e.impl()->end), "depends_on");
  }
  void depends_on(const std::vector<event>& evs) { for (auto& e : evs) depends_on(e); }
  template <typename T> void memcpy_h2d(T* dst, const T* src, size_t n) {
    detail::check(q_->h, pbx_copy_to_device(q_->h, src, dst, (int64_t)(n * sizeof(T))), "copy_to_device");
  }
  template <typename F> void host_task(F&& f) {
    detail::check(q_->h, pbx_synchronize(q_->h), "host_task");
    f();
  }
};

class queue {
  std::shared_ptr<detail::queue_impl> impl_;
  std::shared_ptr<detail::event_impl> begin_event() const {
    auto ev = std::make_shared<detail::event_impl>();
    ev->q = impl_;
    detail::check(impl_->h, pbx_event_create(impl_->h, &ev->start), "event create");
    detail::check(impl_->h, pbx_event_create(impl_->h, &ev->end), "event create");
    detail::check(impl_->h, pbx_event_record(impl_->h, ev->start), "event record");
    return ev;
  }
  event end_event(std::shared_ptr<detail::event_impl> ev) const {
    detail::check(impl_->h, pbx_event_record(impl_->h, ev->end), "event record");
    return event(std::move(ev));
  }
 public:
  queue() : impl_(std::make_shared<detail::queue_impl>(0)) {}
  template <typename Selector, typename = std::enable_if_t<!std::is_same_v<std::decay_t<Selector>, queue>>>
  explicit queue(const Selector&, const property_list& = {}) : queue() {}
  template <typename Selector, typename Handler,
            typename = std::enable_if_t<std::is_invocable_v<Handler, exception_list>>>
  queue(const Selector&, Handler h, const property_list& = {}) : queue() { impl_->handler = async_handler(std::move(h)); }
  queue(const device&, const property_list& = {}) : queue() {}

  pbx_handle_t pbx() const { return impl_->h; }
  const std::shared_ptr<detail::queue_impl>& impl() const { return impl_; }
  device get_device() const { return device(impl_); }
  context get_context() const { return context(); }
  void wait() const { detail::check(impl_->h, pbx_synchronize(impl_->h), "queue::wait"); }
  void wait_and_throw() const { wait(); }

  // Runs the command group immediately (in-order stream) and returns an event bracketing it.
  template <typename F> event submit(F&& cgf) {
    auto ev = begin_event();
    handler h(impl_);
    cgf(h);
    return end_event(std::move(ev));
  }
  // Bracket an arbitrary stream operation with a profiling event pair.
  template <typename F> event enqueue(F&& op) const {
    auto ev = begin_event();
    op(impl_->h);
    return end_event(std::move(ev));
  }
  event memcpy(void* dst, const void* src, size_t bytes, const std::vector<event>& deps = {}) {
    for (auto& d : deps) d.wait();
    // interior pointers (allocation + offset) are device memory too: range lookup, as get_pointer_type does
    const bool dst_dev = detail::usm_contains(dst), src_dev = detail::usm_contains(src);
    return enqueue([&](pbx_handle_t h) {
      int st;
      if (dst_dev && src_dev) st = pbx_copy_device_to_device(h, src, dst, (int64_t)bytes);
      else if (dst_dev) st = pbx_copy_to_device(h, src, dst, (int64_t)bytes);
      else st = pbx_copy_to_host(h, src, dst, (int64_t)bytes);
      detail::check(h, st, "queue::memcpy");
      // host buffers are pageable in the callers: keep the copy's source/destination valid
      if (!(dst_dev && src_dev)) detail::check(h, pbx_synchronize(h), "queue::memcpy sync");
    });
  }
  template <typename T> event fill(T* dst, const T& value, size_t count, const std::vector<event>& deps = {}) {
    for (auto& d : deps) d.wait();
    return enqueue([&](pbx_handle_t h) {
      detail::check(h, pbx_fill(h, dst, &value, (int)sizeof(T), (int64_t)count), "queue::fill");
    });
  }
  bool operator==(const queue& o) const { return impl_ == o.impl_; }
};

// ---- USM -----------------------------------------------------------------------------------
template <typename T> T* malloc_device(size_t count, const queue& q) {
  void* p = nullptr;
  detail::check(q.pbx(), pbx_malloc(q.pbx(), &p, (int64_t)(count * sizeof(T))), "malloc_device");
  std::lock_guard<std::mutex> g(detail::usm_mutex());
  detail::usm_registry()[p] = count * sizeof(T);
  return static_cast<T*>(p);
}
inline void free(void* p, const queue& q) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> g(detail::usm_mutex());
    detail::usm_registry().erase(p);
  }
  detail::check(q.pbx(), pbx_free(q.pbx(), p), "sycl::free");
}
inline usm::alloc get_pointer_type(const void* p, const context&) {
  return detail::usm_contains(p) ? usm::alloc::device : usm::alloc::unknown;
}

// ---- buffer ----------------------------------------------------------------------------------
namespace detail {
struct buffer_impl {
  std::shared_ptr<queue_impl> q;  // owning context (lazily the default queue)
  void* dptr = nullptr;
  size_t bytes = 0;
  void* host_writeback = nullptr;  // SYCL host-pointer buffers copy back on destruction
  ~buffer_impl() {
    if (!q || !q->h || !dptr) return;
    if (host_writeback) {
      pbx_copy_to_host(q->h, dptr, host_writeback, (int64_t)bytes);
      pbx_synchronize(q->h);
    }
    {
      std::lock_guard<std::mutex> g(usm_mutex());
      usm_registry().erase(dptr);
    }
    pbx_free(q->h, dptr);
  }
};
inline std::shared_ptr<queue_impl>& default_queue_impl() {
  static std::shared_ptr<queue_impl> q = std::make_shared<queue_impl>(0);
  return q;
}
}  // namespace detail

template <typename T, int D = 1>
class buffer {
  std::shared_ptr<detail::buffer_impl> impl_;
  size_t count_ = 0;
  template <typename U, int E> friend class buffer;
  void allocate(size_t count) {
    impl_ = std::make_shared<detail::buffer_impl>();
    impl_->q = detail::default_queue_impl();
    impl_->bytes = count * sizeof(T);
    count_ = count;
    detail::check(impl_->q->h, pbx_malloc(impl_->q->h, &impl_->dptr, (int64_t)impl_->bytes), "sycl::buffer");
    std::lock_guard<std::mutex> g(detail::usm_mutex());
    detail::usm_registry()[impl_->dptr] = impl_->bytes ? impl_->bytes : 1;
  }
 public:
  using value_type = T;
  buffer() = default;
  explicit buffer(range<1> r) { allocate(r.size()); }
  buffer(T* host, range<1> r) {
    allocate(r.size());
    using NC = std::remove_const_t<T>;
    detail::check(impl_->q->h, pbx_copy_to_device(impl_->q->h, host, impl_->dptr, (int64_t)impl_->bytes), "buffer init");
    detail::check(impl_->q->h, pbx_synchronize(impl_->q->h), "buffer init");
    if constexpr (!std::is_const_v<T>) impl_->host_writeback = const_cast<NC*>(host);
  }
  size_t size() const { return count_; }
  size_t get_count() const { return count_; }
  range<1> get_range() const { return range<1>(count_); }
  T* device_ptr() const { return impl_ ? static_cast<T*>(impl_->dptr) : nullptr; }
  template <typename U, int E = 1> buffer<U, E> reinterpret(range<1> r) const {
    buffer<U, E> b;
    b.impl_ = impl_;
    b.count_ = r.size();
    return b;
  }
  template <typename U, int E = 1> buffer<U, E> reinterpret() const {
    return reinterpret<U, E>(range<1>(count_ * sizeof(T) / sizeof(U)));
  }
  bool operator==(const buffer& o) const { return impl_ == o.impl_; }
};

// sycl::ext::oneapi::experimental::complex<T> (the reference's complex_sycl, include/blas_meta.h:207-209): same layout
// as std::complex<T>, which is all the GEMM path needs
namespace ext { namespace oneapi { namespace experimental {
template <typename T> using complex = std::complex<T>;
// joint_matrix fragment precision tag named by the reference's tf32 launcher call sites
// (test/unittest/joint_matrix/joint_matrix_common.hpp); a type name only
namespace matrix { namespace precision { class tf32 {}; } }
} } }  // namespace ext::oneapi::experimental
// sycl::ext::oneapi::bfloat16: the storage type of the bf16 GEMMs of this library
namespace ext { namespace oneapi { using bfloat16 = __nv_bfloat16; } }

}  // namespace sycl
