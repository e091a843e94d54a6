"""The CUDA path against OUTPUTS OF THE REFERENCE ITSELF.

oracle/_ref is portBLAS's own GEMM (blas::_gemm down to its Gemm<> kernels) compiled from /root/reference over a host
stand-in for the SYCL runtime (oracle/ref_host_driver.cpp).  Two checks through the C-ABI (portblas_b200.blas -> pbx_gemm):

  * the committed reference outputs (tests/golden/ref_host_golden.npz, made by tests/golden/make_ref_host_golden.py):
    always available, no /root/reference needed on the box;
  * the prebuilt libraries, live, on seeded U(-2,5) inputs over the reference's grids -- strided, batched, interleaved,
    tall-skinny (the reference's GemmPartial + Reduction path), half -- when oracle/_ref travelled to the box.

Tolerances are the north star's: fp64 1e-12 and fp32 1e-5 relative to |alpha||A||B| + |beta||C| (the scale of the terms
that were summed), plus the reference's own predicate (utils::compare_vectors, float_comparison.hpp:163-188) that its unit
tests apply between the library and CBLAS; half uses that predicate with its half margins."""
import itertools
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle, ref_host

pytestmark = pytest.mark.gpu
TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
GOLDEN = Path(__file__).parent / "golden" / "ref_host_golden.npz"
REL = {"f32": 1e-5, "f64": 1e-12}


def _bound(ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc, sa=0, sb=0, sc=0, batch=1, interleaved=False):
    """|alpha| |op(A)| |op(B)| + |beta| |C| on the same buffers (what the error of any summation order scales with)."""
    out = np.abs(C).astype(np.float64)
    oracle.gemm(ta, tb, m, n, k, abs(al), np.abs(A).astype(np.float64), lda, np.abs(B).astype(np.float64), ldb, abs(be),
                out, ldc, stridea=sa, strideb=sb, stridec=sc, batch=batch, interleaved=interleaved, mode=oracle.MODE_REF)
    return out


def _cuda(handle, tdt, ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc, sa=0, sb=0, sc=0, batch=1, interleaved=False,
          out_dt=None):
    import torch
    from portblas_b200 import blas
    a = torch.from_numpy(A).cuda().to(tdt)
    b = torch.from_numpy(B).cuda().to(tdt)
    c = torch.from_numpy(C).cuda().to(out_dt or tdt)
    if interleaved:
        blas._gemm_batched(handle, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc, batch, blas.gemm_batch_type_t.interleaved)
    elif batch > 1:
        blas._gemm_strided_batched(handle, ta, tb, m, n, k, al, a, lda, sa, b, ldb, sb, be, c, ldc, sc, batch)
    else:
        blas._gemm(handle, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc)
    handle.wait()
    return c.cpu().numpy()


def _close(got, want, bound, rel, what):
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    worst = float((err / (bound + 1e-300)).max())
    assert worst <= rel, (what, worst)


def test_reference_output_fixtures_through_the_cuda_path(handle):
    import torch
    g = np.load(GOLDEN, allow_pickle=False)
    n_cases = len([k for k in g.files if k.endswith("_meta")])
    assert n_cases >= 12
    for i in range(n_cases):
        backend, dt, ta, tb, m, n, k, al, be, la, lb, lc, batch = (str(x) for x in g[f"case{i}_meta"])
        m, n, k, la, lb, lc, batch = (int(x) for x in (m, n, k, la, lb, lc, batch))
        al, be = float(al), float(be)
        A, B, C, want = g[f"case{i}_A"], g[f"case{i}_B"], g[f"case{i}_C"], g[f"case{i}_out"]
        lda, ldb, ldc = (k if ta == "t" else m) * la, (n if tb == "t" else k) * lb, m * lc
        sa, sb, sc = m * k * la, k * n * lb, m * n * lc
        tdt = torch.float64 if dt == "f64" else torch.float32
        got = _cuda(handle, tdt, ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc, sa, sb, sc, batch)
        bound = _bound(ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc, sa, sb, sc, batch)
        what = f"fixture {i} ({backend} {dt} {ta}{tb} {m}x{n}x{k} batch {batch})"
        _close(got, want, bound, REL[dt], what)
        assert oracle.compare(got, want, "double" if dt == "f64" else "float") == 0, what   # padding untouched too


@pytest.fixture(scope="module")
def ref_libs():
    if not all(ref_host.usable(b) for b in ref_host.BACKENDS):   # probed in a child process
        pytest.skip("oracle/_ref did not travel to this box or cannot run on it (it is built where /root/reference exists)")
    return True


@pytest.mark.parametrize("backend", ["default", "nvidia_gpu"])
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_cuda_path_matches_the_reference_live(handle, ref_libs, backend, dt):
    import torch
    npdt, tdt = (np.float64, torch.float64) if dt == "f64" else (np.float32, torch.float32)
    rng = np.random.default_rng(77)
    shapes = [(11, 16, 17), (63, 33, 63), (128, 127, 300), (264, 136, 1032), (517, 260, 129)]
    for (ta, tb), (m, n, k), (al, be), (la, lb, lc) in itertools.product(TRANS, shapes, [(1.5, 0.5), (1.0, 0.0)],
                                                                         [(1, 1, 1), (2, 3, 4)]):
        lda, ldb, ldc = (k if ta == "t" else m) * la, (n if tb == "t" else k) * lb, m * lc
        A = oracle.random_uniform(rng, m * k * la, npdt)
        B = oracle.random_uniform(rng, k * n * lb, npdt)
        C = oracle.random_uniform(rng, m * n * lc, npdt)
        want = C.copy()
        ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, want, ldc, backend=backend)
        got = _cuda(handle, tdt, ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc)
        _close(got, want, _bound(ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc), REL[dt],
               (backend, dt, ta, tb, m, n, k, al, be, la))
        assert oracle.compare(got, want, "double" if dt == "f64" else "float") == 0


def test_batched_and_interleaved_match_the_reference_live(handle, ref_libs):
    import torch
    rng = np.random.default_rng(78)
    for (ta, tb), sam, sbm, scm in itertools.product(TRANS, [0, 1, 2], [0, 1], [1, 3]):
        m, n, k, batch = 63, 40, 72, 5
        lda, ldb, ldc = (k if ta == "t" else m) * 2, (n if tb == "t" else k) * 3, m * 4
        sa, sb, sc = m * k * 2, k * n * 3, m * n * 4
        A = oracle.random_uniform(rng, sa * 3 * batch, np.float32)
        B = oracle.random_uniform(rng, sb * 3 * batch, np.float32)
        C = oracle.random_uniform(rng, sc * 3 * batch, np.float32)
        want = C.copy()
        ref_host.gemm_strided_batched(ta, tb, m, n, k, 3.0, A, lda, sa * sam, B, ldb, sb * sbm, 7.0, want, ldc, sc * scm,
                                      batch, backend="nvidia_gpu")
        got = _cuda(handle, torch.float32, ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, C, ldc, sa * sam, sb * sbm, sc * scm,
                    batch)
        bound = _bound(ta, tb, m, n, k, 3.0, A, lda, B, ldb, 7.0, C, ldc, sa * sam, sb * sbm, sc * scm, batch)
        _close(got, want, bound, REL["f32"], ("strided", ta, tb, sam, sbm, scm))
    for (ta, tb), (m, n, k, batch) in itertools.product(TRANS, [(15, 32, 17, 3), (63, 16, 33, 5), (49, 65, 3, 32)]):
        lda, ldb, ldc = (k if ta == "t" else m), (n if tb == "t" else k), m
        A = oracle.random_uniform(rng, m * k * batch, np.float32)
        B = oracle.random_uniform(rng, k * n * batch, np.float32)
        C = oracle.random_uniform(rng, m * n * batch, np.float32)
        want = C.copy()
        ref_host.gemm_batched(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 0.5, want, ldc, batch, True, backend="default")
        got = _cuda(handle, torch.float32, ta, tb, m, n, k, 1.5, A, lda, B, ldb, 0.5, C, ldc, batch=batch, interleaved=True)
        bound = _bound(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 0.5, C, ldc, batch=batch, interleaved=True)
        _close(got, want, bound, REL["f32"], ("interleaved", ta, tb, m, n, k, batch))


def test_tall_skinny_matches_the_reference_s_split_k_live(handle, ref_libs):
    """The reference's tall-skinny route (GemmPartial + Reduction, intel_gpu.hpp with GEMM_TALL_SKINNY_SUPPORT) beside
    this library's split-K on the same inputs: two different summation orders, both within 1e-5 of the scale bound."""
    import torch
    rng = np.random.default_rng(79)
    for (ta, tb), (m, n, k), be in itertools.product(TRANS, [(64, 33, 8200), (16, 255, 4100), (128, 128, 8192)], [0.0, 0.5]):
        lda, ldb, ldc = (k if ta == "t" else m), (n if tb == "t" else k), m
        A = oracle.random_uniform(rng, m * k, np.float32)
        B = oracle.random_uniform(rng, k * n, np.float32)
        C = oracle.random_uniform(rng, m * n, np.float32)
        want = C.copy()
        ref_host.gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, be, want, ldc, backend="intel_gpu")
        got = _cuda(handle, torch.float32, ta, tb, m, n, k, 1.5, A, lda, B, ldb, be, C, ldc)
        _close(got, want, _bound(ta, tb, m, n, k, 1.5, A, lda, B, ldb, be, C, ldc), REL["f32"], ("tall", ta, tb, m, n, k, be))


def test_half_matches_the_reference_live(handle, ref_libs):
    """(half, half) and (half, float): the reference's kernel accumulates in the OUTPUT type (gemm_common.hpp:53-59), this
    library always in fp32 -- so (half, float) is held to the fp32 bound and (half, half) to the reference's half margins."""
    import torch
    rng = np.random.default_rng(80)
    for (ta, tb), (m, n, k) in itertools.product(TRANS, [(16, 16, 17), (63, 33, 31), (128, 72, 64)]):
        lda, ldb, ldc = (k if ta == "t" else m), (n if tb == "t" else k), m
        A = oracle.random_uniform(rng, m * k, np.float32).astype(np.float16)
        B = oracle.random_uniform(rng, k * n, np.float32).astype(np.float16)
        C = oracle.random_uniform(rng, m * n, np.float32).astype(np.float16)
        want32 = C.astype(np.float32)
        ref_host.gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 1.5, want32, ldc)
        got32 = _cuda(handle, torch.float16, ta, tb, m, n, k, 1.5, A, lda, B, ldb, 1.5, C.astype(np.float32), ldc,
                      out_dt=torch.float32)
        bound = _bound(ta, tb, m, n, k, 1.5, A.astype(np.float32), lda, B.astype(np.float32), ldb, 1.5,
                       C.astype(np.float32), ldc)
        _close(got32, want32, bound, REL["f32"], ("f16f32", ta, tb, m, n, k))
        want16 = C.copy()
        ref_host.gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, 1.5, want16, ldc)
        got16 = _cuda(handle, torch.float16, ta, tb, m, n, k, 1.5, A, lda, B, ldb, 1.5, C, ldc)
        assert oracle.compare(got16.astype(np.float32), want16.astype(np.float32), "half") == 0, ("f16", ta, tb, m, n, k)


def test_symm_and_trsm_match_the_reference_live(handle, ref_libs):
    """pbx_symm / pbx_trsm beside the reference's own blas::_symm / blas::_trsm (symm_interface.hpp:35-71,
    trsm_interface.hpp:150-387) on the reference's grids (blas3_symm_test.cpp:155-209, blas3_trsm_test.cpp:130-163 with NaN
    in the unused triangle): _symm to the GEMM bound, _trsm (two different blockings of the same substitution, 16-wide
    inverses there, 128 / 64-wide here) under the reference's predicate and 1e-4 / 1e-11 of the solution's scale."""
    import torch
    from oracle import blas3_ext as ox
    from portblas_b200 import blas
    rng = np.random.default_rng(81)
    for (npdt, tdt, dt, kind, ttol) in ((np.float32, torch.float32, "f32", "float", 1e-4),
                                        (np.float64, torch.float64, "f64", "double", 1e-11)):
        for side, uplo, (m, n), (al, be) in itertools.product("lr", "ul", [(14, 9), (127, 130), (300, 264)],
                                                              [(1.5, 0.5), (3.0, 0.0)]):
            k = m if side == "l" else n
            lda, ldb, ldc = k * 2, m * 3, m * 4
            A = oracle.random_uniform(rng, k * lda, npdt)
            B = oracle.random_uniform(rng, n * ldb, npdt)
            C = oracle.random_uniform(rng, n * ldc, npdt)
            want = C.copy()
            ref_host.symm(side, uplo, m, n, al, A, lda, B, ldb, be, want, ldc)
            a, b, c = (torch.from_numpy(x).cuda() for x in (A, B, C))
            blas._symm(handle, side, uplo, m, n, al, a, lda, b, ldb, be, c, ldc)
            handle.wait()
            got = c.cpu().numpy()
            full = np.abs(np.asfortranarray(ox.symm_full(uplo, k, A, lda)).ravel(order="F")).astype(np.float64)
            bound = np.abs(C).astype(np.float64)
            absB = np.abs(B).astype(np.float64)
            if side == "l":
                oracle.gemm("n", "n", m, n, k, abs(al), full, k, absB, ldb, abs(be), bound, ldc)
            else:
                oracle.gemm("n", "n", m, n, k, abs(al), absB, ldb, full, k, abs(be), bound, ldc)
            _close(got, want, bound, REL[dt], ("symm", dt, side, uplo, m, n, al, be))
            assert oracle.compare(got, want, kind) == 0
        for side, uplo, trans, diag, (m, n), unused in itertools.product("lr", "ul", "nt", "un", [(33, 17), (200, 136)],
                                                                         [0.0, np.nan]):
            k = m if side == "l" else n
            lda, ldb = k * 2, m * 3
            A = ox.fill_trsm_matrix(rng, k, lda, uplo, diag, 5.0, unused, npdt)
            B = oracle.random_uniform(rng, n * ldb, npdt)
            want = B.copy()
            ref_host.trsm(side, uplo, trans, diag, m, n, 2.0, A, lda, want, ldb)
            a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
            blas._trsm(handle, side, uplo, trans, diag, m, n, 2.0, a, lda, b, ldb)
            handle.wait()
            got = b.cpu().numpy()
            what = ("trsm", dt, side, uplo, trans, diag, m, n, unused)
            assert oracle.compare(got, want, kind) == 0, what
            assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= ttol * np.abs(want).max(), what


def test_complex_gemm_matches_the_reference_live(handle, ref_libs):
    """pbx_cgemm / pbx_zgemm beside the reference's own complex GEMM (BLAS_ENABLE_COMPLEX; 'c' handled as 't', as the
    reference does) on the grid of blas3_gemm_test.cpp:143-259 (reduced): 1e-5 / 1e-12 of |alpha||A||B| + |beta||C|."""
    import torch
    from portblas_b200 import blas
    rng = np.random.default_rng(82)

    def rnd(count, cdt):
        return (rng.uniform(-2, 5, count) + 1j * rng.uniform(-2, 5, count)).astype(cdt)

    for cdt, rel in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        for ta, tb, (m, n, k), (al, be) in itertools.product("ntc", "ntc", [(11, 16, 17), (63, 33, 40), (260, 136, 300)],
                                                             [(1.5 + 0.5j, 0.5 - 1j), (1 + 0j, 0j)]):
            lda, ldb, ldc = (k if ta != "n" else m) * 2, (n if tb != "n" else k) * 2, m * 2
            A, B, C = rnd(m * k * 2, cdt), rnd(k * n * 2, cdt), rnd(m * n * 2, cdt)
            want = C.copy()
            ref_host.gemm(ta, tb, m, n, k, al, A, lda, B, ldb, be, want, ldc, backend="nvidia_gpu")
            a, b, c = (torch.from_numpy(x).cuda() for x in (A, B, C))
            blas._gemm(handle, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc)
            handle.wait()
            got = c.cpu().numpy()
            bound = np.abs(C).astype(np.float64)
            oracle.gemm("t" if ta != "n" else "n", "t" if tb != "n" else "n", m, n, k, abs(al), np.abs(A).astype(np.float64),
                        lda, np.abs(B).astype(np.float64), ldb, abs(be), bound, ldc)
            what = ("cgemm", cdt.__name__, ta, tb, m, n, k, al, be)
            assert (np.abs(got.astype(np.complex128) - want) <= rel * 2 * bound + 1e-300).all(), what
            pad = np.ones(C.shape, bool)
            pad.reshape(n, ldc)[:, :m] = False
            assert np.array_equal(got[pad], C[pad]), what
