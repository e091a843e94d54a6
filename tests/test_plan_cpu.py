"""CPU tests of the tensor-core tile selector (pbx_plan_query -> gemm_tcgen05.cu:make_plan), the replacement of the
reference's NVIDIA shape heuristics (src/interface/blas3/backend/nvidia_gpu.hpp:116-171) and of its tall-skinny rule
(gemm_partial_local.hpp:191-199, portblas_handle.hpp:323).  No device is needed: the plan is a pure function of the
shape and the SM count."""
import ctypes

import pytest

from portblas_b200 import _lib

F32, F16, BF16 = _lib.F32, _lib.F16, _lib.BF16


def plan(pbx_lib, dt, m, n, k, batch=1, sms=148):
    cg, bn, sl, sw = (ctypes.c_int() for _ in range(4))
    st = pbx_lib.pbx_plan_query(sms, dt, m, n, k, batch, ctypes.byref(cg), ctypes.byref(bn), ctypes.byref(sl),
                                ctypes.byref(sw))
    assert st == 0
    return cg.value, bn.value, sl.value, bool(sw.value)


def test_baseline_configs(pbx_lib):
    # cfg3 SGEMM 16384^3 and its 8-GPU M-block shard: CTA pairs on 256x256 tiles, no split
    assert plan(pbx_lib, F32, 16384, 16384, 16384) == (2, 256, 1, False)
    assert plan(pbx_lib, F32, 2048, 16384, 16384) == (2, 256, 1, False)
    # cfg4: HBM-bound 16-bit batches -> one 256x256 pair tile per batch entry (every operand byte loaded once), handed
    # out dynamically (round 2: 548 vs 490 TFLOP/s for the 128x128 single-CTA tiles the static schedule preferred)
    assert plan(pbx_lib, BF16, 256, 256, 256, 4096) == (2, 256, 1, False)
    assert plan(pbx_lib, F16, 256, 256, 256, 512) == (2, 256, 1, False)
    # compute-bound 16-bit: CTA pairs
    assert plan(pbx_lib, BF16, 8192, 8192, 8192) == (2, 256, 1, False)
    # cfg5 tall-skinny: 4 pair tiles x 18 K slices = 72 of the 74 pairs in one round (round 1: 37 slices in two rounds;
    # measured equal: 1.79 vs 1.88 ms, profiles/r02/plan_probe_f32.jsonl)
    cg, bn, slices, swapped = plan(pbx_lib, F32, 512, 512, 1 << 20)
    assert (cg, bn, swapped) == (2, 256, False) and slices == 18
    # cfg1 shape: 16 pair tiles x 4 K slices fill 64 of the 74 pairs (measured best of every configuration: 19.2 us)
    assert plan(pbx_lib, F32, 1024, 1024, 1024) == (2, 256, 4, False)
    # the shape that exposed round 1's ">= 0.6 of a wave" rule (86 half-tiles in two rounds, 53 TFLOP/s): 44 pair tiles
    # x 3 slices in two rounds, 187 TFLOP/s
    assert plan(pbx_lib, F32, 384, 5408, 3456) == (2, 256, 3, False)
    # 16-bit mid-size shapes: one round of single-CTA tiles beats 0.6 rounds of pairs
    assert plan(pbx_lib, BF16, 384, 5408, 3456) == (1, 128, 1, False)


def test_skinny_m_swaps_operands(pbx_lib):
    cg, bn, slices, swapped = plan(pbx_lib, F32, 40, 1000, 520)
    assert (cg, bn, swapped) == (1, 64, True)
    assert plan(pbx_lib, F32, 64, 401408, 1152)[3] is True
    assert plan(pbx_lib, F32, 65, 401408, 1152)[3] is False          # only M <= 64
    assert plan(pbx_lib, F32, 40, 30, 520)[3] is False               # and only when N > M


def test_split_k_rules(pbx_lib):
    # never more slices than K blocks / 4, never for short K loops, never more than two rounds' worth of units
    for dt, kblock in ((F32, 32), (BF16, 64)):
        for m, n, k in [(128, 128, 3136), (64, 64, 784), (256, 196, 2304), (512, 512, 65536), (4096, 4096, 4096),
                        (128, 128, 256), (300, 260, 4104)]:
            cg, bn, slices, _ = plan(pbx_lib, dt, m, n, k)
            kb = -(-k // kblock)
            assert 1 <= slices <= max(1, kb // 4) or slices == 1
            tiles = -(-m // (128 * cg)) * -(-n // bn)
            assert tiles * slices <= 2 * (148 // cg) or slices == 1, (dt, m, n, k, cg, bn, slices)
            if kb < 16:
                assert slices == 1, (dt, m, n, k, slices)
    assert plan(pbx_lib, F32, 128, 128, 3136)[2] > 1
    assert plan(pbx_lib, F32, 4096, 4096, 4096)[2] == 1
    assert plan(pbx_lib, BF16, 8192, 8192, 8192)[2] == 1


def test_plan_scales_with_sm_count(pbx_lib):
    # a smaller part needs fewer tiles to be "full": the same shape stops splitting
    assert plan(pbx_lib, F32, 1024, 1024, 1024, sms=148)[2] > 1
    assert plan(pbx_lib, F32, 1024, 1024, 1024, sms=32)[2] == 1


def test_invalid_queries(pbx_lib):
    z = ctypes.c_int()
    args = [ctypes.byref(z)] * 4
    assert pbx_lib.pbx_plan_query(0, F32, 8, 8, 8, 1, *args) == 6
    assert pbx_lib.pbx_plan_query(148, _lib.F64, 8, 8, 8, 1, *args) == 6      # fp64 runs on the DMMA kernel
    assert pbx_lib.pbx_plan_query(148, F32, 0, 8, 8, 1, *args) == 6


def test_plan_invariants_over_random_shapes(pbx_lib):
    """Whatever the shape, the plan is one the kernels can run: a pair tile only where M (and N for 256-wide tiles) can
    feed it, the operand swap exactly for skinny M, K slices that all own at least one K block (an empty slice would add
    an uninitialised partial in the reduce), at least 4 K blocks per slice when the planner chose to split, and not more
    work items than two rounds of the machine unless the tiles alone exceed that."""
    from hypothesis import given, settings, strategies as st
    dims = st.one_of(st.integers(1, 600), st.integers(1, 20000))

    @settings(max_examples=600, deadline=None)
    @given(st.sampled_from([F32, F16, BF16]), dims, dims, st.one_of(st.integers(1, 5000), st.integers(1, 1 << 20)),
           st.sampled_from([1, 1, 1, 3, 32, 1000]), st.sampled_from([148, 132, 74]))
    def check(dt, m, n, k, batch, sms):
        cg, bn, slices, swapped = plan(pbx_lib, dt, m, n, k, batch, sms)
        assert cg in (1, 2) and bn in (64, 128, 256) and slices >= 1
        assert swapped == (m <= 64 and n > m)
        if swapped:
            assert (cg, bn) == (1, 64)
        else:
            assert bn != 64 and not (cg == 2 and m <= 128) and not (bn == 256 and (n <= 128 or cg == 1))
        k_block = 32 if dt == F32 else 64
        kb = -(-k // k_block)
        assert slices <= kb
        per = -(-kb // slices)
        assert per * (slices - 1) < kb                       # no empty trailing slice
        if slices > 1:
            assert kb >= 16 and per >= 4
            units = sms // cg
            tiles = (-(-n // 128) if swapped else -(-m // (128 * cg)) * -(-n // bn)) * batch
            assert tiles * slices <= 2 * units
    check()
