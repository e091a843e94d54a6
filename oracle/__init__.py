"""CPU oracle for the GEMM path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__ (build + smoke check) and bench.py's reference / cpu_baseline leg may import, link or execute
anything in this directory; nothing under portblas_b200/ does, and the product fails loudly without its CUDA library.

  gemm_oracle.c, oracle.py   C restatement of portBLAS's GEMM semantics (file:line citations in the C header) + CBLAS front
  blas3_ext.py               numpy restatement of _symm, _trsm and complex GEMM
  sycl_host/                 host stand-in for the SYCL runtime (sycl/sycl.hpp, ext/oneapi/experimental/sycl_complex.hpp)
  ref_host_driver.cpp        C entry points over the REFERENCE's own blas::_gemm / _symm / _trsm, compiled from
  ref_host.py                /root/reference over sycl_host/ into _ref/ (`make ref`); ctypes front for them
  _ref/                      built artefacts only (git-ignored, travels to the GPU box): the three backend libraries, the
                             reference's own unit tests / bench_gemm / sample linked with its own library on the stand-in
  _build/                    liboracle.so
"""
