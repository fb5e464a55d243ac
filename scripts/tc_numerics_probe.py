"""What does tcgen05.mma (kind::f16, fp32 accumulate in TMEM) do to the low bits?

The tensor-core scan's error bound (csrc/tc.h) needs ONE hardware fact: how far the accumulator
of an MMA instruction can be from the exact value of  acc_in + sum_{k<16} a_k b_k  (the products
of two bf16 / fp16 values are exact in fp32).  This probe measures it through the scan's dump mode
(kge_rank_args_t.tc_dump) with operands that are exactly representable, so the lo planes are zero
and every accumulator is the result of ceil(K/16) "real" instructions:

  crafted  -- one product of magnitude 1 plus up to 15 products of 2^-j: shows at which bit
              addends are cut (guard bits), whether the cut is truncation or rounding, and
              whether the incoming accumulator is treated like a product;
  random   -- operands with 8-bit significands and exponents spread over 2^-w..2^w: the largest
              observed  |result - exact| / (2^-24 * M),  M = sum over instructions of
              (|acc_in| + sum |products|)  -- the constant the bound needs.

    python scripts/tc_numerics_probe.py [--fp16]   ->  gpurun_out/tc_numerics_probe[_fp16].json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torchkge_b200 import _lib  # noqa: E402
from torchkge_b200.engine import CudaEngine, ModelSpec  # noqa: E402

DEV = torch.device("cuda:0")


def run_pairs(eng, a, b):
    """a, b: (n, d) float32 CPU, exactly representable in the operand format.  Returns the
    tensor-core accumulator of pair i = (a[i], b[i]) as float64 (n,)."""
    n, d = a.shape
    ent = torch.cat([b, a]).to(DEV).contiguous()
    rel = torch.ones(1, d, device=DEV)
    spec = ModelSpec(_lib.DISTMULT, d, n, 1, ent[:n].contiguous(), None, rel, None)
    hrows = ent[n:].contiguous().view(n, 1, d)
    r_idx = torch.zeros(n, dtype=torch.int64, device=DEV)
    tcp = eng.pack_tc(spec)
    dump = torch.full((n, n), float("nan"), device=DEV)
    raw = torch.zeros(n, dtype=torch.int32, device=DEV)
    eng.rank_side(spec, None, _lib.SIDE_TAIL, hrows, hrows, r_idx, r_idx, None, raw, torch.zeros_like(raw),
                  tc_packed=tcp, tc_dump=dump)
    torch.cuda.synchronize()
    return dump.diagonal().cpu().double()


def crafted(eng):
    out = []
    names, A, B = [], [], []

    def add(name, prods, d=16):
        a = torch.ones(d)
        b = torch.zeros(d)
        b[:len(prods)] = torch.tensor(prods, dtype=torch.float32)
        names.append(name); A.append(a); B.append(b)

    for j in range(18, 34):
        s = 2.0 ** -j
        add("1 + 15 x 2^-%d" % j, [1.0] + [s] * 15)
        add("1 - 15 x 2^-%d" % j, [1.0] + [-s] * 15)
        add("-1 + 15 x 2^-%d" % j, [-1.0] + [s] * 15)
        add("15 x 2^-%d + 1 (big last)" % j, [s] * 15 + [1.0])
        add("1 + 1 x 2^-%d" % j, [1.0, s])
        add("1 - 1 x 2^-%d" % j, [1.0, -s])
        add("1 + 1.5 x 2^-%d" % j, [1.0, 1.5 * s])
        add("1 + 1.75 x 2^-%d" % j, [1.0, 1.75 * s])
    res16 = run_pairs(eng, torch.stack(A), torch.stack(B))
    for nm, a, b, r in zip(names, A, B, res16.tolist()):
        exact = float((a.double() * b.double()).sum())
        big = 1.0 if exact > 0 else -1.0
        out.append({"case": nm, "k": 16, "exact_minus_big_in_ulp": (exact - big) * 2 ** 23,
                    "got_minus_big_in_ulp": (r - big) * 2 ** 23})
    # two instructions: the 1.0 arrives through the accumulator
    names, A, B = [], [], []
    for j in range(18, 34):
        s = 2.0 ** -j
        a = torch.ones(32)
        b = torch.zeros(32); b[0] = 1.0; b[16:31] = s
        names.append("acc=1 then 15 x 2^-%d" % j); A.append(a); B.append(b)
        b2 = torch.zeros(32); b2[0] = 1.0; b2[16:31] = -s
        names.append("acc=1 then -15 x 2^-%d" % j); A.append(a); B.append(b2)
        b3 = torch.zeros(32); b3[0] = 1.0; b3[16] = s
        names.append("acc=1 then 1 x 2^-%d" % j); A.append(a); B.append(b3)
        b4 = torch.zeros(32); b4[0] = 1.0; b4[16] = 1.5 * s
        names.append("acc=1 then 1.5 x 2^-%d" % j); A.append(a); B.append(b4)
    res32 = run_pairs(eng, torch.stack(A), torch.stack(B))
    for nm, a, b, r in zip(names, A, B, res32.tolist()):
        exact = float((a.double() * b.double()).sum())
        out.append({"case": nm, "k": 32, "exact_minus_big_in_ulp": (exact - 1.0) * 2 ** 23,
                    "got_minus_big_in_ulp": (r - 1.0) * 2 ** 23})
    return out


def quantize(x, bits):
    """round to `bits` significant bits (exactly representable in bf16 for 8, fp16 for 11)"""
    m, e = torch.frexp(x)
    return torch.ldexp(torch.round(m * 2 ** bits) / 2 ** bits, e)


def random_cases(eng, bits, max_abs_exp):
    out = []
    g = torch.Generator().manual_seed(1)
    n = 2048
    for k in (16, 32, 64, 208, 800, 2000):
        for width in (0, 2, 6, 12):
            w = min(width, max_abs_exp)
            for sign in ("mixed", "same"):
                a = torch.randn(n, k, generator=g)
                b = torch.randn(n, k, generator=g)
                if sign == "same":
                    a, b = a.abs(), b.abs()
                if w:
                    a = a * torch.exp2(torch.randint(-w, w + 1, (n, k), generator=g).float())
                    b = b * torch.exp2(torch.randint(-w, w + 1, (n, k), generator=g).float())
                a, b = quantize(a, bits), quantize(b, bits)
                got = run_pairs(eng, a, b)
                p = a.double() * b.double()                       # exact products
                groups = p.view(n, k // 16, 16)
                gsum = groups.sum(2)
                acc_in = torch.cumsum(gsum, 1) - gsum              # exact running value before each instruction
                m_run = (acc_in.abs() + groups.abs().sum(2)).sum(1)
                m_tot = p.abs().sum(1)
                exact = p.sum(1)
                err = (got - exact).abs()
                # the last rounding to fp32 of the result itself is unavoidable: report both
                out.append({"k": k, "instructions": k // 16, "exp_width": w, "sign": sign,
                            "max_err_over_u_running": float((err / (2.0 ** -24 * m_run)).max()),
                            "max_err_over_u_sumabs_per_instr": float((err / (2.0 ** -24 * m_tot * (k // 16))).max()),
                            "max_err_over_u_sumabs": float((err / (2.0 ** -24 * m_tot)).max()),
                            "mean_signed_err_over_u_result": float(((got - exact) / (2.0 ** -24 * exact.abs().clamp_min(1e-300))).mean())})
    return out


def main():
    fp16 = "--fp16" in sys.argv
    if fp16:
        os.environ["KGE_TC_FP16"] = "1"
    eng = CudaEngine(tensor_core=True)
    res = {"operand_format": "fp16" if fp16 else "bf16", "crafted": crafted(eng),
           "random": random_cases(eng, 11 if fp16 else 8, 6 if fp16 else 12)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "tc_numerics_probe%s.json" % ("_fp16" if fp16 else ""))
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    for r in res["crafted"]:
        print("%-34s k=%2d exact %+10.4f ulp   got %+10.4f ulp" % (r["case"], r["k"], r["exact_minus_big_in_ulp"],
                                                                  r["got_minus_big_in_ulp"]))
    for r in res["random"]:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
