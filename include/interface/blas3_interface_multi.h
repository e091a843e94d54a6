// blas3_interface_multi.h -- opt-in multi-GPU mode of the GEMM interface: blas::multi::SB_Handle_Group + blas::multi::_gemm*.
//
// The reference has nothing to cite here: a blas::SB_Handle wraps ONE sycl::queue (include/sb_handle/portblas_handle.h:51-60)
// and blas::_gemm (include/interface/blas3_interface.h:88-95) runs on that queue's device.  This header keeps _gemm's
// argument list (trans chars, post-transpose M N K, host scalars by value, column-major operands with leading
// dimensions) and spreads the call over the B200s of one box:
//   * _gemm(group, ..., a_blocks, lda, b_full, ldb, beta, c_full, ldc, gather): operands resident on the devices, device g
//     owning rows SB_Handle_Group::mblock(M, g) of op(A) and computing the same rows of C; with gather every device's
//     epilogue stores its tiles into all devices' C over NVLink (pbx_gemm_sharded);
//   * _gemm_strided_batched(group, ...): batch ranges SB_Handle_Group::batch_range(batch, g) per device;
//   * _gemm_host(group, ...): HOST operands in, host result out (the copy_to_device + _gemm + copy_to_host of
//     samples/gemm.cpp:50-66 across all devices; B crosses PCIe once and is exchanged over NVLink).
// Errors follow blas::_gemm: the reference's std::invalid_argument texts (gemm_interface.hpp:144-165).
#pragma once
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../pbx_gemm.h"
#include "blas3_interface.h"

namespace blas {
namespace multi {

class SB_Handle_Group {
  pbx_multi_t mh_ = nullptr;

 public:
  // n_devices == 0: every visible device.  `ordinals` may repeat a device (two shards on one GPU; used by the tests).
  explicit SB_Handle_Group(int n_devices = 0, const int* ordinals = nullptr) {
    if (n_devices <= 0) {
      // probe: the largest group that can be built
      for (int n = 8; n >= 1 && !mh_; --n)
        if (pbx_multi_create(&mh_, n, nullptr) != PBX_OK) mh_ = nullptr;
    } else if (pbx_multi_create(&mh_, n_devices, ordinals) != PBX_OK) {
      mh_ = nullptr;
    }
    if (!mh_) throw std::runtime_error("SB_Handle_Group: no group of sm_100 devices with peer access could be created");
  }
  ~SB_Handle_Group() { if (mh_) pbx_multi_destroy(mh_); }
  SB_Handle_Group(const SB_Handle_Group&) = delete;
  SB_Handle_Group& operator=(const SB_Handle_Group&) = delete;

  pbx_multi_t pbx() const { return mh_; }
  int device_count() const { return pbx_multi_device_count(mh_); }
  pbx_handle_t handle(int g) const { return pbx_multi_handle(mh_, g); }
  void wait() { check(pbx_multi_synchronize(mh_)); }

  // [first row, rows) of device g's M-block; [first entry, entries) of its batch range
  std::pair<int64_t, int64_t> mblock(int64_t m, int g) const { return range(m, g, 256); }
  std::pair<int64_t, int64_t> batch_range(int64_t batch, int g) const { return range(batch, g, 1); }

  template <typename T> T* allocate(int g, size_t count) {
    void* p = nullptr;
    check(pbx_malloc(handle(g), &p, (int64_t)(count * sizeof(T))));
    return static_cast<T*>(p);
  }
  void deallocate(int g, void* p) { check(pbx_free(handle(g), p)); }
  template <typename T> void copy_to_device(int g, const T* host_src, T* dev_dst, size_t count) {
    check(pbx_copy_to_device(handle(g), host_src, dev_dst, (int64_t)(count * sizeof(T))));
  }
  template <typename T> void copy_to_host(int g, const T* dev_src, T* host_dst, size_t count) {
    check(pbx_copy_to_host(handle(g), dev_src, host_dst, (int64_t)(count * sizeof(T))));
  }

  void check(int st) const {
    if (st == PBX_OK) return;
    if (st >= PBX_ERR_INVALID_TRANSA && st <= PBX_ERR_INVALID_STRIDEB) throw std::invalid_argument(pbx_status_string(st));
    throw std::runtime_error(std::string(pbx_status_string(st)) + ": " + pbx_multi_last_error(mh_));
  }

 private:
  std::pair<int64_t, int64_t> range(int64_t total, int g, int64_t align) const {
    int64_t s = 0, c = 0;
    if (pbx_shard_range(total, device_count(), g, align, &s, &c) != PBX_OK) throw std::invalid_argument("shard range");
    return {s, c};
  }
};

namespace detail {
template <typename in_t, typename out_t> constexpr int dtype_of() { return blas::internal::pbx_dtype_of<in_t, out_t>::value; }
template <typename out_t> using scalar_abi_t = std::conditional_t<std::is_same_v<out_t, double>, double, float>;
}  // namespace detail

// C <- alpha*op(A)*op(B) + beta*C, M-block sharded over the group; asynchronous (group.wait()).
template <typename in_t, typename out_t, typename element_t, typename index_t>
void _gemm(SB_Handle_Group& group, char _TransA, char _TransB, index_t _M, index_t _N, index_t _K, element_t _alpha,
           const std::vector<const in_t*>& a_blocks, index_t _lda, const std::vector<const in_t*>& b_full, index_t _ldb,
           element_t _beta, const std::vector<out_t*>& c_full, index_t _ldc, bool gather = true) {
  const int G = group.device_count();
  if ((int)a_blocks.size() != G || (int)b_full.size() != G || (int)c_full.size() != G)
    throw std::invalid_argument("one operand pointer per device of the group is required");
  const detail::scalar_abi_t<out_t> alpha = static_cast<detail::scalar_abi_t<out_t>>(_alpha),
                                    beta = static_cast<detail::scalar_abi_t<out_t>>(_beta);
  std::vector<const void*> a(a_blocks.begin(), a_blocks.end()), b(b_full.begin(), b_full.end());
  std::vector<void*> c(c_full.begin(), c_full.end());
  group.check(pbx_gemm_sharded(group.pbx(), detail::dtype_of<in_t, out_t>(), _TransA, _TransB, (int64_t)_M, (int64_t)_N,
                               (int64_t)_K, &alpha, a.data(), (int64_t)_lda, b.data(), (int64_t)_ldb, &beta, c.data(),
                               (int64_t)_ldc, gather ? 1 : 0));
}

// strided batches, batch-range sharded: X_shards[g] points at the first entry device g owns.
template <typename in_t, typename out_t, typename element_t, typename index_t>
void _gemm_strided_batched(SB_Handle_Group& group, char _TransA, char _TransB, index_t _M, index_t _N, index_t _K,
                           element_t _alpha, const std::vector<const in_t*>& a_shards, index_t _lda, index_t _stridea,
                           const std::vector<const in_t*>& b_shards, index_t _ldb, index_t _strideb, element_t _beta,
                           const std::vector<out_t*>& c_shards, index_t _ldc, index_t _stridec, index_t batch_size) {
  const int G = group.device_count();
  if ((int)a_shards.size() != G || (int)b_shards.size() != G || (int)c_shards.size() != G)
    throw std::invalid_argument("one operand pointer per device of the group is required");
  const detail::scalar_abi_t<out_t> alpha = static_cast<detail::scalar_abi_t<out_t>>(_alpha),
                                    beta = static_cast<detail::scalar_abi_t<out_t>>(_beta);
  std::vector<const void*> a(a_shards.begin(), a_shards.end()), b(b_shards.begin(), b_shards.end());
  std::vector<void*> c(c_shards.begin(), c_shards.end());
  group.check(pbx_gemm_strided_batched_sharded(group.pbx(), detail::dtype_of<in_t, out_t>(), _TransA, _TransB, (int64_t)_M,
                                               (int64_t)_N, (int64_t)_K, &alpha, a.data(), (int64_t)_lda, (int64_t)_stridea,
                                               b.data(), (int64_t)_ldb, (int64_t)_strideb, &beta, c.data(), (int64_t)_ldc,
                                               (int64_t)_stridec, (int64_t)batch_size));
}

// HOST operands, synchronous: returns when C_host holds the result.
template <typename in_t, typename out_t, typename element_t, typename index_t>
void _gemm_host(SB_Handle_Group& group, char _TransA, char _TransB, index_t _M, index_t _N, index_t _K, element_t _alpha,
                const in_t* a_host, index_t _lda, const in_t* b_host, index_t _ldb, element_t _beta, out_t* c_host,
                index_t _ldc) {
  const detail::scalar_abi_t<out_t> alpha = static_cast<detail::scalar_abi_t<out_t>>(_alpha),
                                    beta = static_cast<detail::scalar_abi_t<out_t>>(_beta);
  group.check(pbx_gemm_sharded_host(group.pbx(), detail::dtype_of<in_t, out_t>(), _TransA, _TransB, (int64_t)_M, (int64_t)_N,
                                    (int64_t)_K, &alpha, a_host, (int64_t)_lda, b_host, (int64_t)_ldb, &beta, c_host,
                                    (int64_t)_ldc));
}

}  // namespace multi
}  // namespace blas
