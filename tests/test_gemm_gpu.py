"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle.

The parameter grids restate the reference's own GEMM test suites
(test/unittest/blas3/blas3_gemm_test.cpp:30-141, blas3_gemm_batched_test.cpp:30-147,
blas3_gemm_tall_skinny_test.cpp:30-103, test/unittest/joint_matrix/*.cpp) -- same shapes,
transposes, scalars, ld multipliers, offsets, batch sizes and stride multipliers -- and run
them for every (in,out) element-type pair of the C-ABI.

Bars: the reference's almost_equal margins (float 5e-3/1e-3, double 1e-10, half 5e-2/1.0;
common/include/common/float_comparison.hpp:101-158) AND the north-star bounds against the
long-double truth: fp64 1e-12, fp32 (3xTF32) 1e-5, relative to |alpha||A||B| + |beta||C|.
"""
from __future__ import annotations

import itertools

import pytest

from gemm_case import Case, run_case

pytestmark = pytest.mark.gpu

ALL_DTYPES = ["f32", "f64", "f16", "f16f32", "bf16", "bf16f32"]
TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
SIMT, TCGEN05, DMMA = 1, 2, 3


def _run_all(handle, cases):
    failures = []
    for cs in cases:
        r = run_case(handle, cs)
        if not r.ok:
            failures.append(f"{cs.ident()} kernel={r.kernel} sk={r.split_k} ref_mismatch={r.ref_mismatch} "
                            f"bound_viol={r.bound_violations} max_rel={r.max_rel_bound:.3e} {r.detail}")
    assert not failures, f"{len(failures)}/{len(cases)} cases failed:\n" + "\n".join(failures[:20])


# ---- Gemm suites (blas3_gemm_test.cpp) -------------------------------------------------------
def _small_beta_nonzero(dt):
    return [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=1.5)
            for m, n, k, (ta, tb) in itertools.product([11, 16, 32], [11, 16, 32], [16, 17], TRANS)]


def _small_beta_zero(dt, lds):
    return [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=17, alpha=1.5, beta=0.0, lda_mul=lds[0],
                 ldb_mul=lds[1], ldc_mul=lds[2])
            for m, n, (ta, tb) in itertools.product([11, 32], [11, 32], TRANS)]


def _alpha_zero(dt):
    return [Case(dtype=dt, m=16, n=16, k=17, alpha=0.0, beta=b, offset=off, lda_mul=la, ldb_mul=lb, ldc_mul=lc)
            for off, b, la, lb, lc in itertools.product([0, 10], [0.0, 1.0], [1, 2], [1, 2], [1, 2])]


def _offset_nonzero(dt):
    return [Case(dtype=dt, m=m, n=n, k=k, alpha=1.0, beta=1.0, offset=off, lda_mul=la, ldb_mul=lb, ldc_mul=lc)
            for off, m, n, k, la, lb, lc in itertools.product([1, 10], [16, 63], [16, 63], [17, 63], [1, 2], [1, 2],
                                                              [1, 2])]


def _large(dt, ms):
    return [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.0, beta=1.0)
            for m, n, k, (ta, tb) in itertools.product(ms, [257, 511], [253, 511], TRANS)]


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_gemm_small_beta_nonzero_ld_match(handle, dt):
    _run_all(handle, _small_beta_nonzero(dt))


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_gemm_small_beta_zero(handle, dt):
    _run_all(handle, _small_beta_zero(dt, (1, 1, 1)) + _small_beta_zero(dt, (2, 3, 4)))


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_gemm_alpha_zero(handle, dt):
    _run_all(handle, _alpha_zero(dt))


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_gemm_offset_nonzero(handle, dt):
    _run_all(handle, _offset_nonzero(dt))


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_gemm_large_beta_nonzero(handle, dt):
    ms = [253, 511, 1024, 2048, 2200] if dt in ("f32", "f64") else [253, 1024, 2200]
    _run_all(handle, _large(dt, ms))


# ---- every kernel family on the same shapes (forced) -------------------------------------------
@pytest.mark.parametrize("dt,kernel", [("f32", SIMT), ("f32", TCGEN05), ("f64", SIMT), ("f64", DMMA),
                                       ("f16", SIMT), ("f16", TCGEN05), ("bf16", TCGEN05),
                                       ("f16f32", TCGEN05), ("bf16f32", TCGEN05)])
def test_gemm_forced_kernel(handle, dt, kernel):
    # ld multiples of 8 elements keep TMA eligibility for every type; 2200/264 are ragged vs 128/256 tiles
    cases = [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=b, kernel=kernel)
             for m, n, k, b, (ta, tb) in itertools.product([8, 136, 2200], [24, 264], [40, 520], [0.0, 0.5], TRANS)]
    _run_all(handle, cases)
    r = run_case(handle, cases[-1])
    want = {SIMT: "simt", TCGEN05: "tcgen05", DMMA: "dmma"}[kernel]
    assert r.kernel == want, f"forced kernel {want} not used (got {r.kernel})"


# ---- BatchGemm suites (blas3_gemm_batched_test.cpp) ----------------------------------------------
@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_batched_beta_nonzero_ld_match(handle, dt):
    cases = [Case(dtype=dt, api="batched", transa=ta, transb=tb, m=m, n=n, k=k, alpha=3.0, beta=7.0, offset=off,
                  batch=5)
             for off, m, n, k, (ta, tb) in itertools.product([0, 33], [63, 128], [63, 128], [63, 128], TRANS)]
    _run_all(handle, cases)


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_batched_beta_nonzero_ld_multiplied(handle, dt):
    sizes = [63, 128, 129] if dt in ("f32", "f64") else [63, 129]
    cases = [Case(dtype=dt, api="batched", transa=ta, transb=tb, m=m, n=n, k=k, alpha=3.0, beta=7.0, offset=off,
                  batch=bs, batch_type=bt, lda_mul=2, ldb_mul=3, ldc_mul=4)
             for off, bs, m, n, k, (ta, tb), bt in itertools.product([0, 33], [1, 5], sizes, sizes, sizes, TRANS,
                                                                     [0, 1])]
    _run_all(handle, cases)


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_batched_alpha_zero(handle, dt):
    cases = [Case(dtype=dt, api="batched", transa=ta, transb=tb, m=128, n=128, k=128, alpha=0.0, beta=7.0, batch=5)
             for ta, tb in TRANS]
    cases += [Case(dtype=dt, api="batched", transa=ta, transb=tb, m=63, n=63, k=63, alpha=0.0, beta=7.0, batch=5,
                   offset=off, lda_mul=2, ldb_mul=3, ldc_mul=4) for off, (ta, tb) in itertools.product([0, 33], TRANS)]
    _run_all(handle, cases)


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_strided_batched_default(handle, dt):
    cases = [Case(dtype=dt, api="strided", transa=ta, transb=tb, m=m, n=n, k=k, alpha=3.0, beta=7.0, offset=off,
                  batch=bs)
             for off, bs, m, n, k, (ta, tb) in itertools.product([0, 33], [1, 5], [63, 128], [63, 128], [63, 128],
                                                                 TRANS)]
    _run_all(handle, cases)


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_strided_batched_all_strides(handle, dt):
    cases = [Case(dtype=dt, api="strided", transa=ta, transb=tb, m=63, n=63, k=128, alpha=al, beta=be, offset=off,
                  batch=5, lda_mul=2, ldb_mul=3, ldc_mul=4, stride_a_mul=sa, stride_b_mul=sb, stride_c_mul=sc)
             for off, (ta, tb), al, be, sa, sb, sc in itertools.product([0, 33], TRANS, [3.0, 0.0], [7.0, 1.0, 0.0],
                                                                        [0, 1, 2], [0, 1, 2], [1, 3])]
    _run_all(handle, cases)


def test_strided_batched_tma_aligned(handle):
    """Aligned batches (the cfg-4 shape family at reduced batch) must take the tcgen05 path,
    including stride 0 broadcast of A or B."""
    cases = []
    for dt in ("f16", "bf16", "f16f32", "f32"):
        for (ta, tb), sa, sb in itertools.product(TRANS, [0, 1], [0, 1]):
            cases.append(Case(dtype=dt, api="strided", transa=ta, transb=tb, m=256, n=256, k=256, alpha=1.0, beta=0.0,
                              batch=6, stride_a_mul=sa, stride_b_mul=sb, stride_c_mul=1))
    _run_all(handle, cases)
    assert run_case(handle, cases[0]).kernel == "tcgen05"


def test_tma_store_and_direct_epilogues(handle):
    """16-bit outputs with beta == 0 leave through shared memory + TMA stores when C is 16-byte legal;
    the direct-store epilogue (PBX_TMA_STORE=0, or an odd ldc) must agree with it on ragged shapes."""
    cases = []
    for dt, env in itertools.product(("f16", "bf16"), ((("PBX_TMA_STORE", "1"),), (("PBX_TMA_STORE", "0"),))):
        for m, n, ldc_mul, (ta, tb) in itertools.product([72, 200, 264], [40, 136, 300], [1, 2], TRANS):
            cases.append(Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=72, alpha=1.5, beta=0.0, ldc_mul=ldc_mul,
                              kernel=TCGEN05, env=env))
        cases.append(Case(dtype=dt, api="strided", m=136, n=72, k=64, alpha=1.0, beta=0.0, batch=7, stride_c_mul=3,
                          kernel=TCGEN05, env=env))
    _run_all(handle, cases)


# ---- invalid arguments (gemm_interface.hpp:144-165) ------------------------------------------------
def test_invalid_arguments(handle):
    cases = [Case(transa="x"), Case(transb="q"), Case(transa="c", transb="C"),
             Case(api="strided", batch=3, stride_c_mul=0), Case(transa="x", alpha=0.0, beta=2.0)]
    _run_all(handle, cases)


# ---- TallSkinnyGemm suites (blas3_gemm_tall_skinny_test.cpp) + forced split-K ----------------------
@pytest.mark.parametrize("dt", ["f32", "f64", "f16f32"])
def test_tall_skinny(handle, dt):
    cases = []
    for off, m, n, k, (ta, tb), be, lds in itertools.product([0, 10], [7, 65], [9, 126], [2049, 1026], TRANS,
                                                             [0.5, 0.0], [(1, 1, 1), (2, 3, 4)]):
        cases.append(Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=be, offset=off,
                          lda_mul=lds[0], ldb_mul=lds[1], ldc_mul=lds[2]))
    _run_all(handle, cases)


@pytest.mark.parametrize("dt,kernel", [("f32", TCGEN05), ("f32", SIMT), ("f64", DMMA), ("bf16f32", TCGEN05),
                                       ("f16", TCGEN05)])
def test_split_k_forced(handle, dt, kernel):
    cases = [Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=be, kernel=kernel, split_k=sk)
             for m, n, k, be, sk, (ta, tb) in itertools.product([72, 256], [136], [4104, 16384], [0.0, 0.5], [3, 7],
                                                                TRANS)]
    _run_all(handle, cases)
    r = run_case(handle, cases[-1])
    assert r.split_k == 7


def test_split_k_auto_selected(handle):
    """cfg-5 family at reduced K: M=N=512 gives 16 tiles on 148 SMs -> the selector must split K."""
    r = run_case(handle, Case(dtype="f32", m=512, n=512, k=65536, alpha=1.0, beta=0.0))
    assert r.ok, r
    assert r.kernel == "tcgen05" and r.split_k > 1, r


def test_split_k_mid_k_few_tiles(handle):
    """CNN-style rows of the reference sweeps (benchmark/config_csv/blas3/gemm/*im2col*): one or two output tiles
    with K of a few thousand must not leave 140 SMs idle behind a single CTA's latency-bound K loop."""
    for dt, m, n, k in [("f32", 128, 128, 3136), ("f32", 64, 64, 784), ("bf16", 256, 196, 2304), ("f16f32", 64, 576, 3072)]:
        r = run_case(handle, Case(dtype=dt, m=m, n=n, k=k, alpha=1.0, beta=0.0))
        assert r.ok, r
        assert r.kernel == "tcgen05" and r.split_k > 1, (dt, m, n, k, r)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16", "f16f32", "bf16f32"])
def test_skinny_m_swapped_operands(handle, dt):
    """M <= 64 < N runs as C^T = op(B)^T op(A)^T with a transposed-store epilogue (vector stores, TMA store for
    16-bit C, split-K partials); it must agree with the un-swapped plan (PBX_TC_SWAP=0) on ragged shapes."""
    cases = []
    for env in ((), (("PBX_TC_SWAP", "0"),)):
        for m, n, k, be, (ta, tb) in itertools.product([8, 40, 64], [136, 1000], [72, 520], [0.0, 0.5], TRANS):
            cases.append(Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=be, kernel=TCGEN05, env=env))
        cases.append(Case(dtype=dt, m=24, n=264, k=4104, alpha=1.0, beta=0.5, kernel=TCGEN05, split_k=5, env=env))
        cases.append(Case(dtype=dt, m=56, n=200, k=136, alpha=1.0, beta=0.0, ldc_mul=3, kernel=TCGEN05, env=env))
        cases.append(Case(dtype=dt, api="strided", transa="t", m=64, n=392, k=72, alpha=1.0, beta=0.0, batch=6,
                          stride_a_mul=0, kernel=TCGEN05, env=env))
    _run_all(handle, cases)


def test_unaligned_operands_repacked_to_tensor_cores(handle):
    """Odd leading dimensions / element offsets (the reference's OffsetNonZero and LD-multiplied grids, and e.g.
    64x147x423200 of its ResNet sweep) cannot be addressed by TMA: the operand is re-laid out once and the call
    still runs on tcgen05."""
    cases = []
    for dt in ("f32", "f16", "bf16f32"):
        for (ta, tb), off in itertools.product(TRANS, [1, 33]):
            cases.append(Case(dtype=dt, transa=ta, transb=tb, m=261, n=259, k=517, alpha=1.5, beta=0.5, offset=off))
        cases.append(Case(dtype=dt, api="strided", m=131, n=67, k=129, batch=5, stride_a_mul=0, stride_b_mul=2,
                          offset=3, alpha=1.0, beta=0.0))
        cases.append(Case(dtype=dt, api="strided", transa="t", transb="t", m=131, n=67, k=129, batch=5, stride_a_mul=2,
                          stride_b_mul=0, offset=3, alpha=1.0, beta=1.0))
    _run_all(handle, cases)
    r = run_case(handle, cases[0])
    assert r.kernel == "tcgen05", r
    assert handle.last_repack == 3


# ---- JointMatrix-style grid (test/unittest/joint_matrix/*.cpp): narrow-compute numerics ---------------
@pytest.mark.parametrize("dt", ["f16f32", "bf16f32", "f16"])
def test_joint_matrix_grid(handle, dt):
    cases = []
    for m, n, k in [(11, 11, 17), (33, 63, 64), (65, 127, 65), (255, 511, 127), (1024, 1535, 1536)]:
        for (ta, tb), be in itertools.product(TRANS, [0.0, 1.5]):
            cases.append(Case(dtype=dt, transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=be, offset=33))
    _run_all(handle, cases)


def test_joint_matrix_tf32_mode(handle):
    """SB_ENABLE_JOINT_MATRIX=1 (the reference's per-call switch, nvidia_gpu.hpp:68-69) runs float GEMMs with
    10-bit-mantissa fragments and fp32 accumulation; here: one tf32 MMA per product.  As in
    test/unittest/joint_matrix/tf32_float_16_16_8.cpp the inputs carry 13 zeroed mantissa bits, so the narrow
    fragments are exact and the float tolerances apply.  Without the switch the same call is a 3xTF32 GEMM."""
    env = (("SB_ENABLE_JOINT_MATRIX", "1"),)
    cases = []
    for m, n, k in [(11, 11, 17), (33, 63, 64), (65, 127, 65), (255, 511, 127), (1024, 1535, 1536)]:
        for (ta, tb), be in itertools.product(TRANS, [0.0, 1.5]):
            cases.append(Case(dtype="f32", transa=ta, transb=tb, m=m, n=n, k=k, alpha=1.5, beta=be, offset=32,
                              kernel=TCGEN05, env=env, zero_low_bits=13))
    for cfg in ("1,128", "2,128", "2,256"):
        cases.append(Case(dtype="f32", transa="t", m=520, n=392, k=1096, alpha=1.5, beta=0.5, kernel=TCGEN05,
                          env=env + (("PBX_TC_CONFIG", cfg),), zero_low_bits=13))
        cases.append(Case(dtype="f32", api="strided", m=264, n=392, k=200, alpha=1.0, beta=0.0, batch=5,
                          kernel=TCGEN05, env=env + (("PBX_TC_CONFIG", cfg),), zero_low_bits=13))
    cases.append(Case(dtype="f32", m=40, n=1000, k=520, alpha=1.5, beta=0.5, kernel=TCGEN05, env=env, zero_low_bits=13))
    cases.append(Case(dtype="f32", m=300, n=260, k=4104, alpha=1.0, beta=0.5, kernel=TCGEN05, split_k=3, env=env,
                      zero_low_bits=13))
    _run_all(handle, cases)
    run_case(handle, cases[0])
    assert handle.last_presplit == 2
    # general inputs: the single-tf32 product is only tf32-accurate, and the default path must not be affected
    r = run_case(handle, Case(dtype="f32", m=256, n=256, k=512, alpha=1.0, beta=0.0, kernel=TCGEN05, env=env))
    assert not r.ok and r.max_rel_bound < 2e-3, r          # ~2^-11 relative to |A||B|, beyond the fp32 bar
    r = run_case(handle, Case(dtype="f32", m=256, n=256, k=512, alpha=1.0, beta=0.0, kernel=TCGEN05))
    assert r.ok and handle.last_presplit != 2, r


# ---- committed golden vectors (tests/golden/make_golden.py) through the CUDA path ----------------------
def test_golden_fixtures_gpu(handle):
    from pathlib import Path

    import numpy as np
    import torch

    from oracle import oracle
    from portblas_b200 import blas

    g = np.load(Path(__file__).parent / "golden" / "gemm_golden.npz", allow_pickle=False)
    n_cases = len([k for k in g.files if k.endswith("_meta")])
    tmap = {"f32": (torch.float32, "float"), "f64": (torch.float64, "double"), "f16": (torch.float16, "half"),
            "bf16": (torch.bfloat16, "half")}
    for i in range(n_cases):
        dt, ta, tb, m, n, k, al, be, la, lb, lc, batch = g[f"case{i}_meta"]
        m, n, k, la, lb, lc, batch = (int(x) for x in (m, n, k, la, lb, lc, batch))
        tdt, kind = tmap[str(dt)]
        lda, ldb, ldc = (k if ta == "t" else m) * la, (n if tb == "t" else k) * lb, m * lc
        a = torch.from_numpy(g[f"case{i}_A"]).cuda().to(tdt)
        b = torch.from_numpy(g[f"case{i}_B"]).cuda().to(tdt)
        c = torch.from_numpy(g[f"case{i}_C"]).cuda().to(tdt)
        if batch == 1:
            blas._gemm(handle, str(ta), str(tb), m, n, k, float(al), a, lda, b, ldb, float(be), c, ldc)
        else:
            blas._gemm_strided_batched(handle, str(ta), str(tb), m, n, k, float(al), a, lda, m * k * la, b, ldb,
                                       k * n * lb, float(be), c, ldc, m * n * lc, batch)
        handle.wait()
        got = c.to(torch.float64 if dt == "f64" else torch.float32).cpu().numpy()
        assert oracle.compare(got, g[f"case{i}_out"], kind) == 0, f"golden case {i} ({dt} {ta}{tb} {m}x{n}x{k})"


def test_fp32_presplit_operands(handle):
    """fp32 with the tf32 lo halves of A and B pre-split in global memory (PBX_TF32_PRESPLIT=1 with PBX_F32_SPLIT16=0:
    the 3xTF32 form of the pre-split; the default for compute-bound shapes is the tf32 + 2 x bf16 form, covered by
    tests/test_split16_gpu.py) must agree with the in-kernel split on every tile configuration, transpose, ragged
    edge, strided batch (incl. stride-0 broadcast), K slice count, the skinny-M operand swap and repacked
    (odd-ld) operands; and a large square shape must pick a pre-split mode on its own."""
    pre = ("PBX_TF32_PRESPLIT", "1")
    no16 = ("PBX_F32_SPLIT16", "0")
    cases = []
    for cfg in ("1,128", "2,128", "2,256"):
        env = (pre, no16, ("PBX_TC_CONFIG", cfg))
        for (ta, tb), be in itertools.product(TRANS, [0.0, 0.5]):
            cases.append(Case(dtype="f32", transa=ta, transb=tb, m=392, n=520, k=1096, alpha=1.5, beta=be,
                              kernel=TCGEN05, env=env))
        cases.append(Case(dtype="f32", m=300, n=260, k=4104, alpha=1.0, beta=0.5, kernel=TCGEN05, split_k=3, env=env))
        cases.append(Case(dtype="f32", api="strided", transa="t", m=264, n=392, k=200, alpha=1.0, beta=0.0, batch=5,
                          stride_a_mul=0, kernel=TCGEN05, env=env))
        cases.append(Case(dtype="f32", api="strided", transb="t", m=264, n=136, k=328, alpha=-1.0, beta=1.0, batch=4,
                          stride_b_mul=2, stride_c_mul=2, kernel=TCGEN05, env=env))
    for (ta, tb) in TRANS:   # skinny-M swap and odd-ld repack under the pre-split
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=40, n=1000, k=520, alpha=1.5, beta=0.5, kernel=TCGEN05,
                          env=(pre, no16)))
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=263, n=131, k=517, alpha=1.5, beta=0.5, offset=1,
                          kernel=TCGEN05, env=(pre, no16)))
    cases.append(Case(dtype="f32", m=512, n=512, k=65536, alpha=1.0, beta=0.0, env=(pre, no16)))
    _run_all(handle, cases)
    # auto selection on a compute-bound shape, checked on the device against an fp64 product
    import torch
    from portblas_b200 import blas
    n = 2048
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.rand(n * n, device="cuda", generator=g) * 7 - 2
    b = torch.rand(n * n, device="cuda", generator=g) * 7 - 2
    c = torch.zeros(n * n, device="cuda")
    blas._gemm(handle, "n", "t", n, n, n, 1.0, a, n, b, n, 0.0, c, n)
    handle.wait()
    assert handle.last_kernel == "tcgen05" and handle.last_presplit == 3   # tf32 + 2 x bf16
    A, B = a.view(n, n).T.double(), b.view(n, n).T.double()          # column-major -> (rows, cols)
    want = A @ B.T
    bound = A.abs() @ B.abs().T
    assert float(((c.view(n, n).T.double() - want).abs() / bound).max()) <= 1e-5
    r = run_case(handle, Case(dtype="f32", m=512, n=512, k=65536, alpha=1.0, beta=0.0))   # AI 127: no pre-pass, the
    assert r.ok and handle.last_presplit == 4                                             # kernel's splitters make bf16 tiles
    r = run_case(handle, Case(dtype="f32", m=512, n=512, k=65536, alpha=1.0, beta=0.0, env=(("PBX_F32_SPLIT16", "0"),)))
    assert r.ok and handle.last_presplit == 0                                             # 3xTF32 in-kernel split on request


@pytest.mark.parametrize("dt", ["f32", "f64", "f16", "bf16f32"])
def test_interleaved_through_the_strided_path(handle, dt):
    """Large interleaved batches are re-laid out as strided batches (one pass each for A, B and -- when beta != 0 -- C),
    run on the tensor-core path and written back interleaved (pbx_api.cu: interleaved_via_strided).  Same grid style as
    the reference's interleaved suite (blas3_gemm_batched_test.cpp: all transposes, ld multipliers, offsets), batch
    counts that are not multiples of the 32-entry transpose tiles, beta == 0 and != 0; the dedicated kernel must give
    the same numbers when PBX_ILV_VIA_STRIDED=0 selects it."""
    cases = []
    for (ta, tb), (al, be) in itertools.product(TRANS, [(1.5, 0.0), (3.0, 7.0)]):
        cases.append(Case(dtype=dt, api="batched", batch_type=1, transa=ta, transb=tb, m=136, n=72, k=200, alpha=al, beta=be,
                          batch=70, env=(("PBX_ILV_VIA_STRIDED", "1"),)))
    cases.append(Case(dtype=dt, api="batched", batch_type=1, m=63, n=129, k=65, alpha=3.0, beta=7.0, batch=33, offset=33,
                      lda_mul=2, ldb_mul=3, ldc_mul=4, env=(("PBX_ILV_VIA_STRIDED", "1"),)))
    cases.append(Case(dtype=dt, api="batched", batch_type=1, m=230, n=49, k=230, alpha=1.0, beta=0.0, batch=100))   # auto
    _run_all(handle, cases)
    r = run_case(handle, cases[-1])
    assert r.ok and r.kernel != "interleaved", r
    r = run_case(handle, Case(dtype=dt, api="batched", batch_type=1, m=230, n=49, k=230, alpha=1.0, beta=0.0, batch=100,
                              env=(("PBX_ILV_VIA_STRIDED", "0"),)))
    assert r.ok and r.kernel == "interleaved", r
