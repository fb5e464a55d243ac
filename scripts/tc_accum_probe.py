"""How exact is the tensor core's fp32 accumulation?  (GPU probe for the next round.)

The error bound of the tensor-core scan (csrc/tc.h: tc_gamma) has three terms; two are arithmetic
facts checked on the CPU (tests/test_tc_bound_cpu.py), the third -- "<= 2^-21 of the running
magnitude per MMA instruction" -- is an ASSUMPTION about tcgen05's accumulator, doubled for
safety.  At K = 800 (ComplEx d=400) it is the largest term, so a measured, smaller constant would
shrink the near-tie band there by up to a third.

This script isolates that error.  For a DistMult-shaped model the dumped approximate score
(kge_rank_args_t.tc_dump) is  fl_tc( sum_k a_hi b_hi + a_lo b_hi + a_hi b_lo ): each product of two
bf16 values is exact in fp32, so the difference to the float64 sum of the SAME products is purely
the accumulation error of the 3 * ceil(K/16) MMA instructions.  It is reported in units of
2^-24 * sum_k |products| per instruction (tc_gamma assumes 8 of them: 2^-21) for random, cancelling
and wide-dynamic-range operands.

    python scripts/tc_accum_probe.py            ->  gpurun_out/tc_accum_probe.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torchkge_b200 import _lib  # noqa: E402
from torchkge_b200.engine import CudaEngine, ModelSpec  # noqa: E402


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def operands(kind, n_q, n_c, d, g):
    a = torch.randn(n_q, d, generator=g)
    b = torch.randn(n_c, d, generator=g)
    if kind == "cancelling":      # large products of alternating sign: running sums far below sum |terms|
        sign = torch.where(torch.arange(d) % 2 == 0, 1.0, -1.0)
        a, b = a.abs(), b.abs() * sign
    elif kind == "wide":          # 2^-12 .. 2^12 per element
        a = a * torch.exp2(torch.randint(-12, 13, (n_q, d), generator=g).float())
        b = b * torch.exp2(torch.randint(-12, 13, (n_c, d), generator=g).float())
    elif kind == "normalised":
        a = torch.nn.functional.normalize(a, dim=1)
        b = torch.nn.functional.normalize(b, dim=1)
    return a.contiguous(), b.contiguous()


def probe(kind, d, dev, eng):
    g = torch.Generator().manual_seed(d)
    n_q, n_c = 256, 2048
    a, b = operands(kind, n_q, n_c, d, g)
    # DistMult tail side with relation rows of ones: query vector = h * 1 = h exactly
    ent = torch.cat([b, a]).to(dev)                       # candidates first, then the query rows
    rel = torch.ones(1, d, device=dev)
    spec = ModelSpec(_lib.DISTMULT, d, n_c, 1, ent[:n_c].contiguous(), None, rel, None)
    hrows = ent[n_c:].contiguous().view(n_q, 1, d)
    r_idx = torch.zeros(n_q, dtype=torch.int64, device=dev)
    true_idx = torch.zeros(n_q, dtype=torch.int64, device=dev)
    tcp = eng.pack_tc(spec)
    dump = torch.full((n_q, n_c), float("nan"), device=dev)
    raw = torch.zeros(n_q, dtype=torch.int32, device=dev)
    sub = torch.zeros_like(raw)
    eng.rank_side(spec, None, _lib.SIDE_TAIL, hrows, hrows, r_idx, true_idx, None, raw, sub,
                  tc_packed=tcp, tc_dump=dump)
    torch.cuda.synchronize()
    got = dump.cpu().double()
    a_hi, b_hi = bf16(a), bf16(b)
    a_lo, b_lo = bf16(a - a_hi), bf16(b - b_hi)
    A_hi, A_lo, B_hi, B_lo = (x.double() for x in (a_hi, a_lo, b_hi, b_lo))
    exact = A_hi @ B_hi.T + A_lo @ B_hi.T + A_hi @ B_lo.T
    mag = A_hi.abs() @ B_hi.abs().T + A_lo.abs() @ B_hi.abs().T + A_hi.abs() @ B_lo.abs().T
    n_instr = 3 * ((d + 15) // 16)
    per_instr = ((got - exact).abs() / (mag * 2.0 ** -24 * n_instr))
    return {"kind": kind, "k": d, "mma_instructions": n_instr,
            "max_error_in_u_sum_abs_per_instruction": float(per_instr.max()),
            "p999": float(per_instr.flatten().kthvalue(int(0.999 * per_instr.numel())).values),
            "assumed_by_tc_gamma": 8.0,
            "max_error_over_u_sum_abs": float(((got - exact).abs() / (mag * 2.0 ** -24)).max())}


def main():
    dev = torch.device("cuda:0")
    eng = CudaEngine(tensor_core=True)
    out = [probe(kind, d, dev, eng) for kind in ("normalised", "cancelling", "wide") for d in (64, 200, 800, 2000)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tc_accum_probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
