"""fp32 as tf32 + 2 x bf16 (gemm_tcgen05_kernel.cuh PRE == 3, gemm_simt.cu split16_kernel): the default for compute-bound fp32
shapes since round 2 (PBX_F32_SPLIT16=0 switches back to the 3xTF32 pre-split).

A_hi*B_hi stays one tf32 MMA on the raw tiles; the two cross terms run as kind::f16 MMAs on bf16 copies made by a
pre-pass (bf16(a) and bf16(a - trunc_tf32(a))) -- two tf32-MMA times per k-step instead of three, and the split itself
stays two orders of magnitude inside the 1e-5 budget (tools/split_emulation.py, profiles/r01/split_emulation.txt).
First box run (round 2): every case below green; SGEMM 8192^3 270 -> 329 TFLOP/s, 16384^3 197 -> 257 (power-capped).
Bars are gemm_case's: the reference's float margins and <= 1e-5 of |alpha||A||B| + |beta||C| against the long-double truth."""
import itertools
import os

import pytest

from gemm_case import Case, run_case

pytestmark = pytest.mark.gpu
TRANS = [("n", "n"), ("n", "t"), ("t", "n"), ("t", "t")]
S16 = (("PBX_F32_SPLIT16", "1"), ("PBX_TF32_PRESPLIT", "1"))
TCGEN05 = 2   # PBX_KERNEL_TCGEN05 (include/pbx_gemm.h)


def test_smallest_case_first(handle):
    """One single-CTA tile, one k-block per operand layout: if a descriptor or byte count is wrong it shows here."""
    for ta, tb in TRANS:
        cs = Case(dtype="f32", transa=ta, transb=tb, m=128, n=128, k=32, alpha=1.0, beta=0.0, kernel=TCGEN05,
                  env=S16 + (("PBX_TC_CONFIG", "1,128"),))
        r = run_case(handle, cs)
        assert r.ok and handle.last_presplit == 3, (cs.ident(), r)


def test_every_tile_configuration_and_layout(handle):
    cases = []
    for cfg in ("1,128", "2,128", "2,256"):
        env = S16 + (("PBX_TC_CONFIG", cfg),)
        for (ta, tb), be in itertools.product(TRANS, [0.0, 0.5]):
            cases.append(Case(dtype="f32", transa=ta, transb=tb, m=392, n=520, k=1096, alpha=1.5, beta=be, kernel=TCGEN05,
                              env=env))
        cases.append(Case(dtype="f32", m=300, n=260, k=4104, alpha=1.0, beta=0.5, kernel=TCGEN05, split_k=3, env=env))
        cases.append(Case(dtype="f32", api="strided", transa="t", m=264, n=392, k=200, alpha=1.0, beta=0.0, batch=5,
                          stride_a_mul=0, kernel=TCGEN05, env=env))
        cases.append(Case(dtype="f32", api="strided", transb="t", m=264, n=136, k=328, alpha=-1.0, beta=1.0, batch=4,
                          stride_b_mul=2, stride_c_mul=2, kernel=TCGEN05, env=env))
    for ta, tb in TRANS:   # skinny-M swap, odd-ld repack, ld multipliers
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=40, n=1000, k=520, alpha=1.5, beta=0.5, kernel=TCGEN05, env=S16))
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=263, n=131, k=517, alpha=1.5, beta=0.5, offset=1,
                          kernel=TCGEN05, env=S16))
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=200, n=136, k=264, lda_mul=2, ldb_mul=3, ldc_mul=4,
                          kernel=TCGEN05, env=S16))
    cases.append(Case(dtype="f32", m=512, n=512, k=65536, alpha=1.0, beta=0.0, env=S16))
    for cs in cases:
        r = run_case(handle, cs)
        assert r.ok, (cs.ident(), r)


IN_KERNEL = (("PBX_TF32_PRESPLIT", "0"),)


def test_in_kernel_bf16_split_every_tile_configuration_and_layout(handle):
    """PRE == 4: no pre-pass, the kernel's splitter warps convert every staged fp32 tile into the bf16 hi / lo tiles
    (a re-layout from 128-byte-swizzled fp32 rows to 64-byte-swizzled bf16 rows that differs for K-major and MN-major
    operands): all four transposes on every tile configuration, ragged edges, batches, split-K, skinny-M swap."""
    cases = []
    for cfg in ("1,128", "2,128", "2,256"):
        env = IN_KERNEL + (("PBX_TC_CONFIG", cfg),)
        for (ta, tb), be in itertools.product(TRANS, [0.0, 0.5]):
            cases.append(Case(dtype="f32", transa=ta, transb=tb, m=392, n=520, k=1096, alpha=1.5, beta=be, kernel=TCGEN05,
                              env=env))
        cases.append(Case(dtype="f32", m=300, n=260, k=4104, alpha=1.0, beta=0.5, kernel=TCGEN05, split_k=3, env=env))
        cases.append(Case(dtype="f32", api="strided", transa="t", m=264, n=392, k=200, alpha=1.0, beta=0.0, batch=5,
                          stride_a_mul=0, kernel=TCGEN05, env=env))
    for ta, tb in TRANS:
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=40, n=1000, k=520, alpha=1.5, beta=0.5, kernel=TCGEN05,
                          env=IN_KERNEL))
        cases.append(Case(dtype="f32", transa=ta, transb=tb, m=128, n=128, k=32, alpha=1.0, beta=0.0, kernel=TCGEN05,
                          env=IN_KERNEL + (("PBX_TC_CONFIG", "1,128"),)))
    cases.append(Case(dtype="f32", m=512, n=512, k=1 << 17, alpha=1.0, beta=0.0, env=IN_KERNEL))
    for cs in cases:
        r = run_case(handle, cs)
        assert r.ok and handle.last_presplit == 4, (cs.ident(), r, handle.last_presplit)


def test_3xtf32_in_kernel_split_still_available(handle):
    """PBX_F32_SPLIT16=0: the round-1 form (three tf32 MMAs per k-step, lo tiles made by the splitter warps)."""
    env = (("PBX_F32_SPLIT16", "0"), ("PBX_TF32_PRESPLIT", "0"))
    for (ta, tb), cfg in itertools.product(TRANS, ("1,128", "2,256")):
        cs = Case(dtype="f32", transa=ta, transb=tb, m=392, n=520, k=1096, alpha=1.5, beta=0.5, kernel=TCGEN05,
                  env=env + (("PBX_TC_CONFIG", cfg),))
        r = run_case(handle, cs)
        assert r.ok and handle.last_presplit == 0, (cs.ident(), r)
