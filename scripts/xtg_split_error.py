"""CPU experiment (numpy, no GPU): accuracy of the two-term bf16 operand split of the XtG contraction.

The kernel (tc_xtg.cu) splits both fp32 operands as x = hi + mid and issues the products hh + hm + mh + mm.
This compares truncation (what the kernel does today: one PRMT per packed pair) with round-to-nearest
(cvt.rn.bf16x2: the same instruction count) and 4 against 3 products, on a K = 200 000 contraction.
Result (seed 0): trunc 4 products 1.7e-5, trunc 3 products 2.7e-5, rn 4 products 3.5e-6, rn 3 products 4.4e-6
(rms error / rms value) -> round-to-nearest with THREE products is 4x more accurate than today's four.
"""
import numpy as np


def trunc_bf16(a):
    return (a.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def rn_bf16(a):
    u = a.view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(np.float32)


def main():
    rng = np.random.default_rng(0)
    P, M, N = 200000, 64, 64
    X = rng.standard_normal((P, M)).astype(np.float32) * rng.lognormal(0, 1, (P, 1)).astype(np.float32)
    G = rng.standard_normal((P, N)).astype(np.float32) * 1e-3
    ref = X.astype(np.float64).T @ G.astype(np.float64)
    f = lambda a, b: a.astype(np.float64).T @ b.astype(np.float64)
    rms = lambda d: np.sqrt(((d - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean())
    for name, fn in (("trunc", trunc_bf16), ("rn", rn_bf16)):
        xh = fn(X); xm = fn(X - xh); gh = fn(G); gm = fn(G - gh)
        d3 = f(xh, gh) + f(xh, gm) + f(xm, gh)
        print(f"{name:6s} 4 products: rms err / rms = {rms(d3 + f(xm, gm)):.2e}   3 products: {rms(d3):.2e}")
    print(f"fp32 matmul           : rms err / rms = {rms((X.T @ G).astype(np.float64)):.2e}")


if __name__ == "__main__":
    main()
