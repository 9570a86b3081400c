"""Probe for the tensor-core question of the north star ("tensor cores only if an Ozaki-style int8 split beats IMAD in ncu"),
for the one stage with a table-independent multiplier: k_horner, s <- RN(s * rho + a_k[n]).

With 8-bit digits, s * rho for all chains of a chunk is ONE integer GEMM per Horner step:
    (chains x 128 digits of s)  x  (128 x 256 Toeplitz matrix of rho's digits)  ->  chains x 256 int32 columns
(exact: 128 products of at most 255^2 fit int32).  This script measures what the library path delivers for that shape
(torch._int_mm -> cuBLASLt IGEMM), i.e. the UNFUSED variant, whose int32 result goes through HBM before the carry
normalisation / rounding / re-splitting that every step needs.  Printed beside it: the time k_horner needs for the same
chains and step on the IMAD path (13.2 ms per 11 364-table chunk, 401 steps: profiles/r2_ncu_full_summary.txt)."""
import time

import torch

chains = 11364 * 20
a = torch.randint(-128, 127, (chains, 128), dtype=torch.int8, device="cuda")
b = torch.randint(-128, 127, (128, 256), dtype=torch.int8, device="cuda")
for _ in range(3):
    c = torch._int_mm(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
e0.record()
for _ in range(reps):
    c = torch._int_mm(a, b)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
macs = chains * 128 * 256
print(f"int8 GEMM {chains} x 128 x 256 -> int32: {us:.1f} us per Horner step, {macs / us / 1e6:.1f} T int8-MAC/s, "
      f"{(a.numel() + c.numel() * 4) / us / 1e3:.0f} GB/s of A-read + C-write")
print(f"the same step on the IMAD path (k_horner, product + alignment + rounding, nothing through HBM): {13.2e3 / 401:.1f} us")
print(f"401 steps: {us * 401 / 1e3:.1f} ms for the products alone vs 13.2 ms for the whole kernel")
