"""CPU emulation of fp32-product splittings for the tensor cores: how much of the 1e-5 error budget does the operand
splitting itself use?  (Accumulation is exact here -- float64 -- so only the split is measured.)

  cur        what gemm_tcgen05.cu does today: a = hi + lo, hi = the top 10 mantissa bits (kind::tf32 ignores the rest),
             lo = rn_tf32(a - hi); D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi           -> three tf32 MMAs per k-step
  tf32+2bf16 hi*hi as one tf32 MMA, the two cross terms as bf16 MMAs on bf16(lo) x bf16(a): the lo halves only need ~8
             bits (they are 2^-11 of the operand), and kind::f16 runs at twice the tf32 rate
                                                                                  -> 1 + 1/2 + 1/2 = two tf32-MMA times
Errors are relative to sum |a||b| (the bound the parity tests use).

    python tools/split_emulation.py
"""
import numpy as np


def trunc_bits(x, keep):
    u = x.astype(np.float32).view(np.uint32)
    return (u & (np.uint32(0xFFFFFFFF) << np.uint32(23 - keep))).view(np.float32)


def rn_bits(x, keep):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    sh = 23 - keep
    u = u + ((1 << (sh - 1)) - 1) + ((u >> sh) & 1)
    return ((u >> sh) << sh).astype(np.uint32).view(np.float32)


def main():
    rng = np.random.default_rng(0)
    tf32_t, tf32_r, bf16_r = (lambda x: trunc_bits(x, 10)), (lambda x: rn_bits(x, 10)), (lambda x: rn_bits(x, 7))
    f8 = np.float64
    print(f"{'inputs':8} {'k':>8} {'cur (3 x tf32)':>16} {'tf32 + 2 x bf16':>16}")
    for dist in ("U(-2,5)", "U(0,1)", "N(0,1)"):
        for k in (1024, 65536, 1048576):
            worst = [0.0, 0.0]
            for _ in range(10):
                if dist == "U(-2,5)":
                    a, b = (rng.uniform(-2, 5, k).astype(np.float32) for _ in range(2))
                elif dist == "U(0,1)":
                    a, b = (rng.uniform(0, 1, k).astype(np.float32) for _ in range(2))
                else:
                    a, b = (rng.standard_normal(k).astype(np.float32) for _ in range(2))
                exact = (a.astype(f8) * b.astype(f8)).sum()
                bound = (np.abs(a).astype(f8) * np.abs(b)).sum()
                ah, bh = tf32_t(a), tf32_t(b)
                al, bl = a - ah, b - bh            # exact in fp32
                cur = (ah.astype(f8) * bh + tf32_r(al).astype(f8) * bh + ah.astype(f8) * tf32_r(bl)).sum()
                new = (ah.astype(f8) * bh + bf16_r(al).astype(f8) * bf16_r(b) + bf16_r(a).astype(f8) * bf16_r(bl)).sum()
                worst[0] = max(worst[0], abs(cur - exact) / bound)
                worst[1] = max(worst[1], abs(new - exact) / bound)
            print(f"{dist:8} {k:8d} {worst[0]:16.2e} {worst[1]:16.2e}")


if __name__ == "__main__":
    main()
