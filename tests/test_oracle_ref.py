"""The oracle pinned against REFERENCE OUTPUTS: oracle/_ref/libportblas_ref_<backend>.so is portBLAS's own GEMM path
(blas::_gemm ... the Gemm<> kernels), compiled unchanged from /root/reference over a host stand-in for the SYCL runtime
(oracle/ref_host_driver.cpp, oracle/sycl_host/sycl/sycl.hpp, `make -C oracle ref`).  These CPU tests run the reference
itself and check, on the reference's own grids and U(-2,5) inputs:

  * the C restatement (oracle/gemm_oracle.c, "gemm_local ordering") is BIT-EXACT with the reference's production kernels
    -- the no-local kernels of default.hpp and the local-memory kernels of nvidia_gpu.hpp -- for every transpose, ragged
    shape, ld multiplier and (alpha, beta), single and strided-batched;
  * the restated DEFAULT-backend CPU kernel (the timed "port" of bench.py) is bit-exact with the reference's;
  * the front-end rules (alpha == 0 before validation, exact zeros vs 0*C, error strings) are the reference's;
  * the interleaved layout, the tall-skinny GemmPartial + Reduction path (intel_gpu.hpp with GEMM_TALL_SKINNY_SUPPORT)
    and the half instantiations agree with the restatement / CBLAS within the reference's own tolerance.

The libraries need /root/reference only at BUILD time; where they were not built (and cannot be) the tests skip, and the
committed outputs of the same libraries (tests/golden/ref_host_golden.npz) still pin the oracle: test_ref_host_golden."""
import itertools
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle, ref_host

TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
GOLDEN = Path(__file__).parent / "golden" / "ref_host_golden.npz"


@pytest.fixture(scope="module")
def ref_libs():
    built = ref_host.build()
    if len(built) < len(ref_host.BACKENDS):
        pytest.skip("oracle/_ref is not built and /root/reference is not present to build it")
    return built


def _case(rng, npdt, ta, tb, m, n, k, lam=1, lbm=1, lcm=1, batch=1):
    lda = (k if ta == "t" else m) * lam
    ldb = (n if tb == "t" else k) * lbm
    ldc = m * lcm
    sa, sb, sc = m * k * lam, k * n * lbm, m * n * lcm
    A = oracle.random_uniform(rng, sa * batch, npdt)
    B = oracle.random_uniform(rng, sb * batch, npdt)
    C = oracle.random_uniform(rng, sc * batch, npdt)
    return A, B, C, lda, ldb, ldc, sa, sb, sc


def test_libraries_are_the_reference_sources(ref_libs):
    for backend in ref_host.BACKENDS:
        L = ref_host.lib(backend)
        assert L.ref_backend().decode() == backend
        assert ref_host.compute_units(backend) >= 1
    # nothing of the reference is copied into the repository: the driver only #includes it
    drv = (Path(ref_host.HERE) / "ref_host_driver.cpp").read_text()
    assert '#include "interface/gemm_interface.hpp"' in drv and "blas::_gemm(" in drv


@pytest.mark.parametrize("backend", ["default", "nvidia_gpu"])
@pytest.mark.parametrize("npdt", [np.float32, np.float64])
def test_restatement_is_bit_exact_with_reference_kernels(ref_libs, backend, npdt):
    """Gemm/Small*, OffsetNonZero-like and LDMultiplied grids (blas3_gemm_test.cpp:30-122) through the reference's own
    backend heuristics: default.hpp picks Tile<2,2,2,2> full-vec (M,N <= 128, K <= 256), Tile<4,4,8,8> full-vec or
    Tile<4,4,4,4> partial-vec (M*N >= 524288), all without local memory; nvidia_gpu.hpp picks its local-memory tiles
    (barriers run as fibers).  The restatement's MODE_LOCAL must reproduce every bit, padding included."""
    rng = np.random.default_rng(2024)
    n_cases = 0
    grid = itertools.product(TRANS, [11, 16, 63], [11, 33], [16, 17, 63], [(1.5, 1.5), (1.5, 0.0), (1.0, 1.0)],
                             [(1, 1, 1), (2, 3, 4)])
    big = [(("n", "n"), 255, 129, 300, (2.0, 3.0), (3, 1, 2)), (("t", "t"), 129, 255, 257, (1.0, 1.0), (1, 1, 1)),
           (("n", "t"), 1024, 520, 24, (1.5, 0.5), (1, 1, 1)), (("t", "n"), 700, 750, 31, (1.5, 0.0), (1, 2, 1)),
           (("n", "n"), 63, 63, 2500, (1.5, 0.5), (1, 1, 1))]
    for (ta, tb), m, n, k, (al, be), lds in itertools.chain(grid, big):
        A, B, C, lda, ldb, ldc, *_ = _case(rng, npdt, ta, tb, m, n, k, *lds)
        got, want = C.copy(), C.copy()
        ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, want, ldc, backend=backend)
        assert oracle.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, got, ldc, mode=oracle.MODE_LOCAL) == 0
        assert np.array_equal(got, want), (backend, ta, tb, m, n, k, al, be, lds, float(np.abs(got - want).max()))
        n_cases += 1
    assert n_cases > 400


def test_reference_passes_its_own_test_criterion_on_the_host(ref_libs):
    """What test/unittest/blas3/blas3_gemm_common.hpp:164-225 checks: the library result against CBLAS on the whole
    buffer with utils::compare_vectors -- here for the reference itself on the stand-in executor (all three backends)."""
    rng = np.random.default_rng(5)
    for backend, npdt, kind in itertools.product(ref_host.BACKENDS, [np.float32, np.float64], ["x"]):
        kind = "float" if npdt == np.float32 else "double"
        for (ta, tb), (m, n, k) in itertools.product(TRANS, [(11, 16, 17), (63, 33, 63), (253, 257, 253)]):
            A, B, C, lda, ldb, ldc, *_ = _case(rng, npdt, ta, tb, m, n, k, 2, 3, 4)
            got, want = C.copy(), C.copy()
            ref_host.gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 1.5, got, ldc, backend=backend)
            oracle.cblas_gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 1.5, want, ldc)
            assert oracle.compare(got, want, kind) == 0, (backend, ta, tb, m, n, k)


@pytest.mark.parametrize("npdt", [np.float32, np.float64])
def test_default_cpu_port_is_bit_exact_with_reference(ref_libs, npdt):
    """oracle_gemm_default_cpu_* restates the kernel default.hpp:98-112 selects for M*N >= 524288 (BASELINE configs[0]:
    1024^3): it is bench.py's timed "port" when oracle/_ref is absent, so it must be the same computation."""
    rng = np.random.default_rng(9)
    for (ta, tb), (m, n, k), (al, be) in itertools.product(TRANS, [(1024, 512, 40), (733, 719, 33)], [(1.0, 0.0), (1.5, 0.5)]):
        A, B, C, lda, ldb, ldc, *_ = _case(rng, npdt, ta, tb, m, n, k)
        got, want = C.copy(), C.copy()
        ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, want, ldc, backend="default")
        oracle.gemm_default_cpu(ta == "t", tb == "t", m, n, k, al, A, lda, B, ldb, be, got, ldc)
        assert np.array_equal(got, want), (ta, tb, m, n, k, al, be)


def test_fibers_off_changes_nothing(ref_libs):
    """bench.py times the default backend with work-items as plain loop iterations (its GEMMs never reach a barrier)."""
    rng = np.random.default_rng(3)
    A, B, C, lda, ldb, ldc, *_ = _case(rng, np.float32, "n", "t", 300, 200, 77)
    a, b = C.copy(), C.copy()
    ref_host.gemm("n", "t", 300, 200, 77, 1.5, A, lda, B, ldb, 0.5, a, ldc)
    ref_host.set_fibers(False)
    try:
        ref_host.gemm("n", "t", 300, 200, 77, 1.5, A, lda, B, ldb, 0.5, b, ldc)
    finally:
        ref_host.set_fibers(True)
    assert np.array_equal(a, b)


def test_front_end_rules_are_the_reference_s(ref_libs):
    """gemm_interface.hpp:105-185 run for real: alpha == 0 short-cuts BEFORE validation; _scal writes exact zeros,
    _scal_matrix computes beta*C (NaN survives beta == 0) and is a no-op for beta == 1; std::invalid_argument texts."""
    rng = np.random.default_rng(11)
    m, n, k = 8, 6, 5
    A, B, C, lda, ldb, ldc, sa, sb, sc = _case(rng, np.float32, "n", "n", m, n, k, 1, 1, 2, batch=3)
    for bad in [("x", "n", "invalid _TransA"), ("n", "y", "invalid _TransB")]:
        with pytest.raises(ref_host.ReferenceError_, match=bad[2]):
            ref_host.gemm(bad[0], bad[1], m, n, k, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc)
        st = oracle.gemm(bad[0], bad[1], m, n, k, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc)
        assert oracle.STATUS_TEXT[st] == bad[2]
        # alpha == 0: not rejected, C <- beta*C
        got, want = C.copy(), C.copy()
        ref_host.gemm(bad[0], bad[1], m, n, k, 0.0, A, lda, B, ldb, 2.0, want, ldc)
        assert oracle.gemm(bad[0], bad[1], m, n, k, 0.0, A, lda, B, ldb, 2.0, got, ldc) == 0
        assert np.array_equal(got, want)
    # stride validation (batch > 1 only): stride_c < ldc*n, negative stride_a / stride_b
    for sa_, sb_, sc_, msg in [(sa, sb, ldc * n - 1, "invalid _stridec"), (-1, sb, sc, "invalid _stridea"),
                               (sa, -1, sc, "invalid _strideb")]:
        with pytest.raises(ref_host.ReferenceError_, match=msg):
            ref_host.gemm_strided_batched("n", "n", m, n, k, 1.0, A, lda, sa_, B, ldb, sb_, 0.0, C.copy(), ldc, sc_, 3)
        st = oracle.gemm("n", "n", m, n, k, 1.0, A, lda, B, ldb, 0.0, C.copy(), ldc, stridea=sa_, strideb=sb_,
                         stridec=sc_, batch=3)
        assert oracle.STATUS_TEXT[st] == msg
    # alpha == 0 with NaN in C: plain _gemm takes _scal_matrix (0*NaN = NaN); beta == 1 leaves C alone
    Cn = C.copy()
    Cn[3] = np.nan
    for beta in (0.0, 1.0, 2.5):
        got, want = Cn.copy(), Cn.copy()
        ref_host.gemm("n", "n", m, n, k, 0.0, A, lda, B, ldb, beta, want, ldc)
        oracle.gemm("n", "n", m, n, k, 0.0, A, lda, B, ldb, beta, got, ldc)
        assert np.array_equal(got, want, equal_nan=True), beta
        assert np.isnan(want[3])
    # contiguous batched C (ldc == m, stride_c == ldc*n): _scal -> exact zeros even over NaN
    A2, B2, C2, lda2, ldb2, ldc2, sa2, sb2, sc2 = _case(rng, np.float32, "n", "n", m, n, k, batch=3)
    C2[5] = np.nan
    got, want = C2.copy(), C2.copy()
    ref_host.gemm_strided_batched("n", "n", m, n, k, 0.0, A2, lda2, sa2, B2, ldb2, sb2, 0.0, want, ldc2, sc2, 3)
    oracle.gemm("n", "n", m, n, k, 0.0, A2, lda2, B2, ldb2, 0.0, got, ldc2, stridea=sa2, strideb=sb2, stridec=sc2, batch=3)
    assert np.array_equal(got, want) and not np.isnan(want).any() and np.all(want[:sc2 * 3] == 0)
    # beta == 0 never reads C (NaN-safe) when alpha != 0
    got = Cn.copy()
    ref_host.gemm("n", "n", m, n, k, 1.5, A, lda, B, ldb, 0.0, got, ldc)
    assert not np.isnan(got.reshape(-1, ldc)[:n, :m]).any()


@pytest.mark.parametrize("backend", ["default", "nvidia_gpu"])
def test_strided_batched_is_bit_exact(ref_libs, backend):
    """BatchStridedGemm grid (blas3_gemm_batched_test.cpp:100-147): stride multipliers 0 (broadcast) / 1 / 2 for A and B,
    1 / 3 for C, ld multipliers 2/3/4."""
    rng = np.random.default_rng(17)
    for (ta, tb), sam, sbm, scm in itertools.product(TRANS, [0, 1, 2], [0, 1, 2], [1, 3]):
        m, n, k, batch = 31, 33, 40, 5
        A, B, C, lda, ldb, ldc, sa, sb, sc = _case(rng, np.float64, ta, tb, m, n, k, 2, 3, 4, batch=3 * batch)
        got, want = C.copy(), C.copy()
        ref_host.gemm_strided_batched(ta, tb, m, n, k, 3.0, A, lda, sa * sam, B, ldb, sb * sbm, 7.0, want, ldc, sc * scm,
                                      batch, backend=backend)
        assert oracle.gemm(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, got, ldc, stridea=sa * sam, strideb=sb * sbm,
                           stridec=sc * scm, batch=batch, mode=oracle.MODE_LOCAL) == 0
        assert np.array_equal(got, want), (backend, ta, tb, sam, sbm, scm)


@pytest.mark.parametrize("backend", ["default", "nvidia_gpu"])
def test_batched_default_strides_and_interleaved(ref_libs, backend):
    """_gemm_batched: strided with the default strides (matrix footprints), and the interleaved kernel
    (gemm_interleaved.hpp:219-312, element (r,c,b) at (c*ld+r)*batch+b; batch sizes that are not a multiple of its
    4-wide batch vectors included) on the same buffers read as interleaved: both bit-exact with the restatement."""
    rng = np.random.default_rng(23)
    for (ta, tb), (m, n, k, batch) in itertools.product(TRANS, [(15, 32, 17, 3), (63, 16, 33, 5)]):
        A, B, C, lda, ldb, ldc, sa, sb, sc = _case(rng, np.float32, ta, tb, m, n, k, batch=batch)
        got, want = C.copy(), C.copy()
        ref_host.gemm_batched(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, want, ldc, batch, backend=backend)
        oracle.gemm(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, got, ldc, stridea=sa, strideb=sb, stridec=sc, batch=batch,
                    mode=oracle.MODE_LOCAL)
        assert np.array_equal(got, want), (backend, ta, tb, m, n, k)
        got, want = C.copy(), C.copy()
        ref_host.gemm_batched(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, want, ldc, batch, True, backend=backend)
        oracle.gemm(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, got, ldc, batch=batch, interleaved=True,
                    mode=oracle.MODE_LOCAL)
        assert np.array_equal(got, want), (backend, "interleaved", ta, tb, m, n, k)


def test_tall_skinny_path_of_the_reference(ref_libs):
    """intel_gpu.hpp:67-140 with GEMM_TALL_SKINNY_SUPPORT sends K >= 4096 && M*N <= 16384 to GemmPartial + Reduction
    (portblas_handle.hpp:302-403): K is cut into `depth` slices, partial products land in a cube and are reduced.  The
    order of additions differs from a single k-ascending sum, so this is a tolerance check -- the reference's own --
    against the long-double truth, for the reference and for the restatement alike (TallSkinnyGemm grid,
    blas3_gemm_tall_skinny_test.cpp, with k reduced to seconds of CPU time)."""
    rng = np.random.default_rng(29)
    for (ta, tb), (m, n, k), (al, be), lcm in itertools.product(TRANS, [(7, 9, 4099), (64, 33, 8200), (16, 255, 4100)],
                                                                [(1.5, 0.0), (1.5, 0.5)], [1, 2]):
        A, B, C, lda, ldb, ldc, *_ = _case(rng, np.float32, ta, tb, m, n, k, 1, 1, lcm)
        ref, truth = C.copy(), C.copy()
        ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, ref, ldc, backend="intel_gpu")
        oracle.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, truth, ldc, mode=oracle.MODE_TRUTH)
        assert oracle.compare(ref, truth, "float") == 0, (ta, tb, m, n, k, al, be, lcm)


def test_half_instantiations(ref_libs):
    """(half, half) and (half, float) (gemm.cpp.in:33-137; default.hpp:150-200 Tile<4,4,8,8>): against the reference's own
    oracle for half -- up-cast, sgemm, down-cast (system_reference_blas.hpp:410-430) -- with its half margins."""
    rng = np.random.default_rng(31)
    for (ta, tb), (m, n, k) in itertools.product(TRANS, [(16, 16, 17), (63, 33, 31)]):
        A, B, C, lda, ldb, ldc, *_ = _case(rng, np.float32, ta, tb, m, n, k)
        Ah, Bh, Ch = (x.astype(np.float16) for x in (A, B, C))
        want = Ch.astype(np.float32)
        oracle.cblas_gemm(ta, tb, m, n, k, 1.5, Ah.astype(np.float32), lda, Bh.astype(np.float32), ldb, 1.5, want, ldc)
        got = Ch.copy()
        ref_host.gemm(ta, tb, m, n, k, 1.5, Ah, lda, Bh, ldb, 1.5, got, ldc)
        assert oracle.compare(got.astype(np.float32), want.astype(np.float16).astype(np.float32), "half") == 0
        got32 = Ch.astype(np.float32)
        ref_host.gemm(ta, tb, m, n, k, 1.5, Ah, lda, Bh, ldb, 1.5, got32, ldc)
        assert oracle.compare(got32, want, "float") == 0


def test_ref_host_golden():
    """Outputs of the reference itself (oracle/_ref, generated here by tests/golden/make_ref_host_golden.py) committed
    as fixtures: they pin the restatement bit for bit wherever oracle/_ref is absent, e.g. a box without /root/reference."""
    g = np.load(GOLDEN, allow_pickle=False)
    n = len([k for k in g.files if k.endswith("_meta")])
    assert n >= 12
    for i in range(n):
        backend, dt, ta, tb, m, nn, k, al, be, la, lb, lc, batch = g[f"case{i}_meta"]
        m, nn, k, la, lb, lc, batch = (int(x) for x in (m, nn, k, la, lb, lc, batch))
        al, be = float(al), float(be)
        A, B, C, want = g[f"case{i}_A"], g[f"case{i}_B"], g[f"case{i}_C"], g[f"case{i}_out"]
        lda = (k if ta == "t" else m) * la
        ldb = (nn if tb == "t" else k) * lb
        ldc = m * lc
        got = C.copy()
        assert oracle.gemm(ta, tb, m, nn, k, al, A, lda, B, ldb, be, got, ldc, stridea=m * k * la, strideb=k * nn * lb,
                           stridec=m * nn * lc, batch=batch, mode=oracle.MODE_LOCAL) == 0
        assert np.array_equal(got, want), f"reference-output fixture {i} ({backend})"
        if ref_host.usable(backend):  # and the library still produces what was committed
            again = C.copy()
            if batch > 1:
                ref_host.gemm_strided_batched(ta, tb, m, nn, k, al, A, lda, m * k * la, B, ldb, k * nn * lb, be, again, ldc,
                                              m * nn * lc, batch, backend=backend)
            else:
                ref_host.gemm(ta, tb, m, nn, k, al, A, lda, B, ldb, be, again, ldc, backend=backend)
            assert np.array_equal(again, want)


# ---- the two callers of the path that the reference implements on top of _gemm: _symm and _trsm -------
def test_symm_of_the_reference_is_the_restated_gemm_on_the_mirrored_matrix(ref_libs):
    """blas::_symm = _gemm_backend with a mirroring operand loader (symm_interface.hpp:35-71, gemm_local.hpp:813-873).
    Grid of blas3_symm_test.cpp:155-209 (reduced): the reference's result must equal, bit for bit, the restated GEMM
    (production ordering) applied to the explicitly mirrored matrix -- which is how pbx_symm is built -- and agree with
    the ext oracle (long-double) and CBLAS under the reference's predicate."""
    from oracle import blas3_ext as ox
    rng = np.random.default_rng(41)
    for npdt, kind in ((np.float32, "float"), (np.float64, "double")):
        for side, uplo, (m, n), (al, be), (la, lb, lc) in itertools.product("lr", "ul", [(14, 9), (63, 40), (127, 130)],
                                                                            [(1.5, 0.5), (3.0, 0.0)], [(1, 1, 1), (2, 3, 4)]):
            k = m if side == "l" else n
            lda, ldb, ldc = k * la, m * lb, m * lc
            A = oracle.random_uniform(rng, k * lda, npdt)
            B = oracle.random_uniform(rng, n * ldb, npdt)
            C = oracle.random_uniform(rng, n * ldc, npdt)
            want = C.copy()
            ref_host.symm(side, uplo, m, n, al, A, lda, B, ldb, be, want, ldc)
            full = np.asfortranarray(ox.symm_full(uplo, k, A, lda)).ravel(order="F").copy()
            got = C.copy()
            if side == "l":
                oracle.gemm("n", "n", m, n, k, al, full, k, B, ldb, be, got, ldc, mode=oracle.MODE_LOCAL)
            else:
                oracle.gemm("n", "n", m, n, k, al, B, ldb, full, k, be, got, ldc, mode=oracle.MODE_LOCAL)
            assert np.array_equal(got, want), (npdt.__name__, side, uplo, m, n, al, be, la)
            truth = C.copy()
            assert ox.symm(side, uplo, m, n, al, A, lda, B, ldb, be, truth, ldc) == 0
            assert oracle.compare(want, truth, kind) == 0
    A = oracle.random_uniform(rng, 64, np.float32)
    for side, uplo, msg in (("l", "x", "invalid _uplo"), ("q", "u", "invalid _side")):
        with pytest.raises(ref_host.ReferenceError_, match=msg):
            ref_host.symm(side, uplo, 8, 8, 1.0, A, 8, A, 8, 0.0, A.copy(), 8)
        assert ox.STATUS_TEXT[ox.symm_status(side, uplo)] == msg


def test_trsm_of_the_reference_against_the_restated_scheme(ref_libs):
    """blas::_trsm (trsm_interface.hpp:150-387: 16-wide DiagonalBlocksInverter blocks, then the GEMM loop) on the grid of
    blas3_trsm_test.cpp:130-163 (reduced; NaN in the unused triangle included): the ext oracle's restatement of that
    scheme agrees with the reference's output far inside the reference's own margins, both agree with CBLAS under
    almost_equal, and the validation messages are the reference's."""
    from oracle import blas3_ext as ox
    rng = np.random.default_rng(43)
    for npdt, kind, tight in ((np.float32, "float", 2e-5), (np.float64, "double", 1e-13)):
        for side, uplo, trans, diag, (m, n), unused in itertools.product("lr", "ul", "nt", "un", [(7, 9), (33, 17), (70, 45)],
                                                                         [0.0, np.nan]):
            k = m if side == "l" else n
            lda, ldb = k * 2, m * 3
            A = ox.fill_trsm_matrix(rng, k, lda, uplo, diag, 5.0, unused, npdt)
            B = oracle.random_uniform(rng, n * ldb, npdt)
            want, model, cb = B.copy(), B.copy(), B.copy()
            ref_host.trsm(side, uplo, trans, diag, m, n, 2.0, A, lda, want, ldb)
            assert ox.trsm_ref_algorithm(side, uplo, trans, diag, m, n, 2.0, A, lda, model, ldb) == 0
            ox.cblas_trsm(side, uplo, trans, diag, m, n, 2.0, np.nan_to_num(A), lda, cb, ldb)
            what = (npdt.__name__, side, uplo, trans, diag, m, n, unused)
            assert not np.isnan(want).any(), what
            assert oracle.compare(want, cb, kind) == 0, what
            scale = np.abs(cb).max()
            assert np.abs(want.astype(np.float64) - model.astype(np.float64)).max() <= tight * scale, what
    A = oracle.random_uniform(rng, 64, np.float32)
    for args, msg in ((("x", "u", "n", "n"), "invalid Side argument"), (("l", "x", "n", "n"), "invalid Triangle argument"),
                      (("l", "u", "x", "n"), "invalid Transpose argument"), (("l", "u", "n", "x"), "invalid Diagonal argument")):
        with pytest.raises(ref_host.ReferenceError_, match=msg):
            ref_host.trsm(*args, 8, 8, 1.0, A, 8, A.copy(), 8)
        assert ox.STATUS_TEXT[ox.trsm_status(*args, 8, 8, 8, 8)] == msg
    with pytest.raises(ref_host.ReferenceError_, match="invalid matrix size argument"):
        ref_host.trsm("l", "u", "n", "n", 0, 8, 1.0, A, 8, A.copy(), 8)


def test_complex_gemm_of_the_reference(ref_libs):
    """complex<float> / complex<double> (BLAS_ENABLE_COMPLEX; grids of blas3_gemm_test.cpp:143-259, reduced) through the
    reference's own complex kernels (default.hpp:202-246: no-local Tile<2,2,4,4> / Tile<8,8,4,4>; nvidia_gpu.hpp:237-260):
    the ext oracle's cgemm -- INCLUDING its restatement of the reference's quirk that 'c' is handled as 't', without
    conjugation -- agrees under the reference's predicate and to 1e-5 / 1e-12 of |alpha||A||B| + |beta||C|; the BLAS
    meaning of 'c' (conjugate) demonstrably does not."""
    from oracle import blas3_ext as ox
    rng = np.random.default_rng(47)

    def rnd(count, cdt):
        return (rng.uniform(-2, 5, count) + 1j * rng.uniform(-2, 5, count)).astype(cdt)

    for backend, (cdt, kind, rel) in itertools.product(["default", "nvidia_gpu"],
                                                       [(np.complex64, "float", 1e-5), (np.complex128, "double", 1e-12)]):
        for ta, tb, (m, n, k), (al, be), lm in itertools.product("ntc", "ntc", [(11, 16, 17), (63, 33, 40), (260, 40, 300)],
                                                                 [(1.5 + 0.5j, 0.5 - 1j), (1 + 0j, 0j)], [1, 2]):
            lda, ldb, ldc = (k if ta != "n" else m) * lm, (n if tb != "n" else k) * lm, m * lm
            A, B, C = rnd(m * k * lm, cdt), rnd(k * n * lm, cdt), rnd(m * n * lm, cdt)
            want, got = C.copy(), C.copy()
            ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, want, ldc, backend=backend)
            assert ox.cgemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, got, ldc) == 0
            what = (backend, cdt.__name__, ta, tb, m, n, k, al, be, lm)
            assert oracle.compare(got.view(got.real.dtype), want.view(want.real.dtype), kind) == 0, what
            bound = np.abs(C).astype(np.float64)
            oracle.gemm("t" if ta != "n" else "n", "t" if tb != "n" else "n", m, n, k, abs(al), np.abs(A).astype(np.float64),
                        lda, np.abs(B).astype(np.float64), ldb, abs(be), bound, ldc)
            assert (np.abs(got.astype(np.complex128) - want) <= rel * 2 * bound + 1e-300).all(), what
    # 'c' really is 't' in the reference
    m, n, k = 9, 7, 8
    A, B, C = rnd(m * k, np.complex128), rnd(k * n, np.complex128), rnd(m * n, np.complex128)
    as_c, as_t, blas_c = C.copy(), C.copy(), C.copy()
    ref_host.gemm("c", "n", m, n, k, 1.0, A, k, B, k, 0.0, as_c, m)
    ref_host.gemm("t", "n", m, n, k, 1.0, A, k, B, k, 0.0, as_t, m)
    ox.cgemm("c", "n", m, n, k, 1.0, A, k, B, k, 0.0, blas_c, m, conj=True)
    assert np.array_equal(as_c, as_t) and np.abs(as_c - blas_c).max() > 1.0


# binary -> (--gtest_filter, minimum number of tests): a bounded subset; the full runs (11 392 tests, all passing) are
# recorded in profiles/r01/ref_host_unittests/
HOST_UNITTESTS = {
    "blas3_gemm_tall_skinny_test": ("*", 160),
    "blas3_symm_test": ("*alloc_usm*", 480),
    "blas3_gemm_test": ("*Small*:*AlphaZero*:*OffsetNonZero*alloc_usm*", 700),
    "blas3_gemm_batched_test": ("*BetaNonZeroLDMatchFloat*", 100),
    "blas3_trsm_test": ("*m_7__n_7_*:*m_16__n_16_*", 100),
}


@pytest.mark.parametrize("name", list(HOST_UNITTESTS))
def test_reference_unit_tests_pass_on_the_host_stand_in(name):
    """The executor behind "reference outputs" is checked with the reference's OWN tests: test/unittest/blas3/<name>.cpp
    (unchanged) linked with the reference's own header-only library over oracle/sycl_host (`make -C oracle ref_tests`),
    comparing with CBLAS through the reference's verifier."""
    import subprocess
    built = {p.name: p for p in ref_host.build_unittests()}
    exe = built.get(f"ref_unittest_host_{name}")
    if exe is None:
        pytest.skip("oracle/_ref unit tests are not built and /root/reference is not present to build them")
    flt, at_least = HOST_UNITTESTS[name]
    r = subprocess.run([str(exe), f"--gtest_filter={flt}"], capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-10:])
    ran = [ln for ln in r.stdout.splitlines() if ln.startswith("[==========]")]
    assert ran, tail + r.stderr[-2000:]
    assert "[  FAILED  ]" not in r.stdout, tail
    assert r.returncode == 0 and int(ran[-1].split()[1]) >= at_least, (r.returncode, ran[-1], tail)


def test_reference_sample_and_benchmark_run_on_the_host_stand_in(tmp_path):
    """BASELINE configs[0] names samples/gemm.cpp on the SYCL host CPU device: here the reference's sample and its
    bench_gemm (BLAS_VERIFY_BENCHMARK on: each benchmark first checks its result against CBLAS) run on the reference's own
    library over the host stand-in (`make -C oracle ref_bench`; cfg1 numbers: profiles/r01/ref_host_bench_gemm_cfg1.txt)."""
    import json
    import subprocess
    built = {p.name: p for p in ref_host.build_bench()}
    if len(built) < 2:
        pytest.skip("oracle/_ref benchmark binaries are not built and /root/reference is not present to build them")
    r = subprocess.run([str(built["ref_sample_gemm_host"])], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C (after):" in r.stdout

    def block(name, rows):
        return np.array([[float(x) for x in ln.split()] for ln in r.stdout.split(name)[1].strip().splitlines()[:rows]])
    A, B, C0, C1 = block("A:", 7), block("B:", 9), block("C (before):", 7), block("C (after):", 7)
    assert np.allclose(C1, 1.5 * A @ B + 0.5 * C0, rtol=2e-3, atol=5e-2)   # the sample prints ~5 significant digits
    csv = tmp_path / "p.csv"
    csv.write_text("n,n,256,256,256,1,0\nt,n,127,65,33,1.5,0.5\n")
    out = tmp_path / "r.json"
    r = subprocess.run([str(built["ref_bench_host_gemm"]), "--csv-param", str(csv), "--benchmark_min_time=0.05",
                        f"--benchmark_out={out}"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ERROR OCCURRED" not in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(out.read_text())["benchmarks"]
    assert len(rep) == 8 and all("error_occurred" not in b and b["n_fl_ops"] > 0 for b in rep)   # 2 rows x {float, double} x {buffer, usm}


def test_fuzz_restatement_against_reference(ref_libs):
    """Seeded random shapes (1 ... 257, occasionally K up to 2049), 'n' / 't' / 'c' in either case, (alpha, beta) incl. 0 and
    negative values, ld multipliers, single and strided-batched, both backends, both real types: every bit equal.
    (A 150 s run of the same generator: 20 229 cases, 0 mismatches.)"""
    import time
    rng = np.random.default_rng(2025)
    t0, n_cases = time.time(), 0
    dims = [1, 2, 3, 5, 8, 15, 16, 17, 31, 33, 64, 65, 100, 129, 200, 257]
    while n_cases < 1500 and time.time() - t0 < 60:
        backend = str(rng.choice(["default", "nvidia_gpu"]))
        npdt = [np.float32, np.float64][int(rng.integers(2))]
        ta, tb = str(rng.choice(list("ntcNT"))), str(rng.choice(list("ntcNT")))
        m, n, k = (int(rng.choice(dims)) for _ in range(3))
        if rng.random() < 0.1:
            k = int(rng.choice([513, 1000, 2049]))
        al, be = float(rng.choice([1.0, 1.5, -2.0, 0.0])), float(rng.choice([0.0, 1.0, 0.5, -1.5]))
        lam, lbm, lcm = (int(rng.choice([1, 1, 2, 3])) for _ in range(3))
        lda, ldb, ldc = (k if ta.lower() != "n" else m) * lam, (n if tb.lower() != "n" else k) * lbm, m * lcm
        batch = int(rng.choice([1, 1, 1, 3]))
        sa, sb, sc = m * k * lam, k * n * lbm, m * n * lcm
        A, B, C = (oracle.random_uniform(rng, s * batch, npdt) for s in (sa, sb, sc))
        got, want = C.copy(), C.copy()
        if batch > 1:
            ref_host.gemm_strided_batched(ta, tb, m, n, k, al, A, lda, sa, B, ldb, sb, be, want, ldc, sc, batch, backend=backend)
            st = oracle.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, got, ldc, stridea=sa, strideb=sb, stridec=sc, batch=batch,
                             mode=oracle.MODE_LOCAL)
        else:
            ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, want, ldc, backend=backend)
            st = oracle.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, got, ldc, mode=oracle.MODE_LOCAL)
        assert st == 0 and np.array_equal(got, want), (backend, npdt.__name__, ta, tb, m, n, k, al, be, lam, lbm, lcm, batch)
        n_cases += 1
    assert n_cases >= 300
