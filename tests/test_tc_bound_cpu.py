"""CPU checks of the error bounds the bound-and-refine scans rest on (csrc/tc.h, csrc/api.cu).

Two of the three terms of `tc_gamma` are pure arithmetic facts that need no GPU:
  * split:  x = hi + lo + r with two bf16 roundings; dropping lo*lo and the residuals costs at
            most 3 * 2^-16 * sum|a_k b_k|
  * ref:    the reference's own fp32 evaluation (products rounded once, summed in ATen's order) is
            within (depth + 4) * 2^-24 * sum|a_k b_k| of the real dot product, `depth` being
            kge_schedule_depth of the schedule the exact kernels replay;
            for the L2 norm: within (depth + 10) * 2^-24 of sum x_k^2 (relative)
(the third, the tensor core's fp32 accumulation, is measured on the GPU by tests/test_tc_gpu.py).
Random and adversarial (cancelling, wide dynamic range) vectors; float64 is the yardstick.
"""
import numpy as np
import pytest
import torch

from torchkge_b200 import _lib


def _bf16_round(x):
    """round-to-nearest-even to bfloat16, returned as float32 (numpy)"""
    return torch.from_numpy(x.astype(np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def _vectors(rng, n, d, kind):
    a = rng.standard_normal((n, d)).astype(np.float32)
    b = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "cancelling":          # products of alternating sign and equal magnitude
        b = np.abs(b) * np.where(np.arange(d) % 2 == 0, 1, -1).astype(np.float32) * np.sign(a + 1e-30)
    elif kind == "wide":              # 2^-20 .. 2^20 dynamic range
        a *= np.exp2(rng.integers(-20, 21, (n, d))).astype(np.float32)
        b *= np.exp2(rng.integers(-20, 21, (n, d))).astype(np.float32)
    elif kind == "normalised":
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
    return a, b


@pytest.mark.parametrize("kind", ["normalised", "cancelling", "wide"])
@pytest.mark.parametrize("d", [13, 64, 200, 800])
def test_bf16_split_term_of_the_bound(kind, d):
    rng = np.random.default_rng(d)
    a, b = _vectors(rng, 400, d, kind)
    a_hi = _bf16_round(a); a_lo = _bf16_round(a - a_hi)
    b_hi = _bf16_round(b); b_lo = _bf16_round(b - b_hi)
    A, B = a.astype(np.float64), b.astype(np.float64)
    kept = (a_hi.astype(np.float64) * b_hi + a_lo.astype(np.float64) * b_hi
            + a_hi.astype(np.float64) * b_lo).sum(1)            # what the three MMAs add up, exactly
    exact = (A * B).sum(1)
    bound = 3.0 * 2.0 ** -16 * (np.abs(A) * np.abs(B)).sum(1)
    assert (np.abs(kept - exact) <= bound).all()
    assert (bound <= 3.0 * 2.0 ** -16 * np.linalg.norm(A, axis=1) * np.linalg.norm(B, axis=1) * (1 + 1e-12)).all()


@pytest.mark.parametrize("kind", ["normalised", "cancelling", "wide"])
@pytest.mark.parametrize("d", [7, 13, 64, 200, 400, 1000])
def test_reference_sum_is_within_depth_times_u_of_the_real_dot(kind, d):
    """DistMult-style `(q * c).sum(dim=-1)` in ATen vs float64, against (depth + 4) u sum|terms|."""
    lib = _lib.load()
    depth = lib.kge_schedule_depth(_lib.DISTMULT, d)
    assert depth >= 0
    rng = np.random.default_rng(100 + d)
    a, b = _vectors(rng, 300, d, kind)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    ref = (ta.view(300, 1, d) * tb.view(300, 1, d)).sum(dim=2).view(-1).double().numpy()
    exact = (a.astype(np.float64) * b.astype(np.float64)).sum(1)
    bound = (depth + 4) * 2.0 ** -24 * (np.abs(a.astype(np.float64)) * np.abs(b.astype(np.float64))).sum(1)
    assert (np.abs(ref - exact) <= bound).all(), float((np.abs(ref - exact) / bound).max())


@pytest.mark.parametrize("planes", [2, 3])
@pytest.mark.parametrize("kind", ["normalised", "cancelling", "wide"])
@pytest.mark.parametrize("d", [7, 13, 100, 200, 520])
def test_reference_multi_plane_sum_is_within_depth_times_u(planes, kind, d):
    """ComplEx (`q0*c0 + q1*c1`, bilinear.py:514-515) and Analogy (`q0*c0 + qm*cm + q1*c1`,
    bilinear.py:695-698) add their rounded products element-wise before the one sum over the last
    axis: two resp. three rounded products and one resp. two rounded additions per element still fit
    the `+ 4` of (depth + 4) u sum|products| -- the reference-side term of tc_gamma for the
    concatenated K = planes * d contraction the tensor-core scan evaluates.  `cancelling` makes the
    planes cancel each other (small element, large products)."""
    lib = _lib.load()
    model = _lib.COMPLEX if planes == 2 else _lib.ANALOGY
    depth = lib.kge_schedule_depth(model, d)
    assert depth >= 0
    rng = np.random.default_rng(1000 * planes + d)
    n = 300
    a, b = _vectors(rng, n, planes * d, kind)
    if kind == "cancelling":    # plane 1 against plane 0: nearly equal products of opposite sign
        a[:, d:2 * d] = a[:, :d]
        b[:, d:2 * d] = -b[:, :d] * np.float32(1 + 2 ** -12)
    ta, tb = torch.from_numpy(a).view(n, planes, d), torch.from_numpy(b).view(n, planes, d)
    q = [ta[:, p].view(n, 1, d) for p in range(planes)]
    c = [tb[:, p].view(n, 1, d) for p in range(planes)]
    if planes == 2:
        ref = (q[0] * c[0] + q[1] * c[1]).sum(dim=2)
    else:
        ref = (q[0] * c[0] + q[1] * c[1] + q[2] * c[2]).sum(dim=2)
    ref = ref.view(-1).double().numpy()
    A, B = a.astype(np.float64), b.astype(np.float64)
    exact = (A * B).sum(1)
    bound = (depth + 4) * 2.0 ** -24 * (np.abs(A) * np.abs(B)).sum(1)
    assert (np.abs(ref - exact) <= bound).all(), float((np.abs(ref - exact) / bound).max())


@pytest.mark.parametrize("kind", ["normalised", "wide"])
@pytest.mark.parametrize("d", [8, 50, 200, 203, 1000])
def test_reference_l2_norm_is_within_depth_times_u(kind, d):
    """`(q - c).norm(p=2, dim=-1) ** 2` in ATen vs float64, against (depth + 10) u sum x^2."""
    lib = _lib.load()
    depth = lib.kge_schedule_depth(_lib.TRANSE_L2, d)
    rng = np.random.default_rng(200 + d)
    a, b = _vectors(rng, 300, d, kind)
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    ref = ((ta.view(300, 1, d) - tb.view(300, 1, d)).norm(p=2, dim=-1) ** 2).view(-1).double().numpy()
    x = a.astype(np.float64) - b.astype(np.float64)
    exact = (x * x).sum(1)
    assert (np.abs(ref - exact) <= (depth + 10) * 2.0 ** -24 * exact).all()


def test_gamma_formulas_match_the_header():
    """The Python mirror of tc.h used by tests/test_tc_gpu.py, pinned to a few known values."""
    lib = _lib.load()
    d = 200
    depth_sum, depth_norm = lib.kge_schedule_depth(_lib.DISTMULT, d), lib.kge_schedule_depth(_lib.TRANSE_L2, d)
    assert (depth_sum, depth_norm) == (16, 31)
    gamma_dot = 3 * 2.0 ** -16 + 2 * (3 * 13 + 2) * 2.0 ** -22 + (depth_sum + 4) * 2.0 ** -24
    assert gamma_dot == pytest.approx(6.652e-5, rel=1e-3)
    assert (depth_norm + 42) * 2.0 ** -24 == pytest.approx(4.351e-6, rel=1e-3)


@pytest.mark.parametrize("d", [16, 200, 1000])
def test_rotate_two_level_sum_stays_inside_the_relative_bound(d):
    """RotatE bound-and-refine (csrc/api.cu: rel_eps): an fp32 evaluation with FMA and a two-level
    sum (32 terms per stage, then the stage sums) is within g |s~| of the ATen-order score, with
    g = (depth_exact + 32 + ceil(d/32) + 23) * 2^-24.  numpy float32 stands in for the fp32 pipes
    (its correctly rounded sqrt is inside the 2^-21 allowed for MUFU.SQRT)."""
    lib = _lib.load()
    depth_e = lib.kge_schedule_depth(_lib.ROTATE, d)
    g = (32 + (d + 31) // 32 + depth_e + 4 + 11 + 8) * 2.0 ** -24 * 1.001
    rng = np.random.default_rng(d)
    nq, nc = 40, 60
    q = (rng.random((nq, 2, d)).astype(np.float32) * 2 - 1) * 0.05
    c = (rng.random((nc, 2, d)).astype(np.float32) * 2 - 1) * 0.05
    tq, tc_ = torch.from_numpy(q), torch.from_numpy(c)
    dre = tq[:, 0].view(nq, 1, d) - tc_[:, 0].view(1, nc, d)
    dim_ = tq[:, 1].view(nq, 1, d) - tc_[:, 1].view(1, nc, d)
    exact = (-torch.stack([dre, dim_], dim=0).norm(dim=0).sum(dim=2)).numpy().astype(np.float64)
    dr = (q[:, None, 0, :] - c[None, :, 0, :]).astype(np.float32)
    di = (q[:, None, 1, :] - c[None, :, 1, :]).astype(np.float32)
    x = (di.astype(np.float64) * di + (dr * dr).astype(np.float32)).astype(np.float32)   # fma(di, di, dr*dr)
    m = np.sqrt(x).astype(np.float32)
    total = np.zeros((nq, nc), dtype=np.float32)
    for k0 in range(0, d, 32):
        stage = np.zeros((nq, nc), dtype=np.float32)
        for k in range(k0, min(d, k0 + 32)):
            stage = (stage + m[:, :, k]).astype(np.float32)
        total = (total + stage).astype(np.float32)
    approx = -total.astype(np.float64)
    assert (np.abs(approx - exact) <= g * np.abs(approx)).all()


def _tc_constants(k_total, depth, l2):
    gamma = 3.0 * 2.0 ** -16 + 2.0 * (3.0 * ((k_total + 15) // 16) + 2.0) * 2.0 ** -22 + (0.0 if l2 else (depth + 4.0) * 2.0 ** -24)
    gamma2 = (depth + 42.0) * 2.0 ** -24
    return np.float32(gamma), np.float32(gamma2)


@pytest.mark.parametrize("l2", [False, True])
@pytest.mark.parametrize("d", [50, 200])
def test_threshold_decisions_never_contradict_the_exact_comparison(l2, d):
    """The tensor-core scan's decision rule (csrc/tc.cu epilogue), restated in numpy: the
    accumulator f = sum(a_hi b_hi + a_lo b_hi + a_hi b_lo) [L2: - |b|^2/2 from three bf16 pieces],
    thresholds T_hi / T_lo from the per-block largest candidate bound.  Wherever the rule says
    "greater" (f > T_hi) the ATen-order score must be >= s_true, wherever it says "smaller"
    (f < T_lo) it must be < s_true -- on random rows, near-duplicates (near ties) and exact
    duplicates; and the undecided band must stay small."""
    lib = _lib.load()
    model = _lib.TRANSE_L2 if l2 else _lib.DISTMULT
    depth = lib.kge_schedule_depth(model, d)
    k_total = d + 3 if l2 else d
    gamma, gamma2 = _tc_constants(k_total, depth, l2)
    rng = np.random.default_rng(7 * d + l2)
    nq, nc = 48, 512
    a = rng.standard_normal((nq, d)).astype(np.float32)
    b = rng.standard_normal((nc, d)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    b[1::7] = b[0:-1:7] * np.float32(1 + 2 ** -20) + np.float32(1e-7) * rng.standard_normal((len(b[1::7]), d)).astype(np.float32)
    b[2::7] = b[0:-2:7]                                    # exact duplicates
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    if l2:
        S = -((ta.view(nq, 1, d) - tb.view(1, nc, d)).norm(p=2, dim=-1) ** 2)
    else:
        S = (ta.view(nq, 1, d) * tb.view(1, nc, d)).sum(dim=2)
    S = S.numpy()
    true_idx = rng.integers(0, nc, nq)
    st = S[np.arange(nq), true_idx]
    a_hi = _bf16_round(a); a_lo = _bf16_round(a - a_hi)
    b_hi = _bf16_round(b); b_lo = _bf16_round(b - b_hi)
    f = (a_hi.astype(np.float64) @ b_hi.T.astype(np.float64) + a_lo.astype(np.float64) @ b_hi.T.astype(np.float64)
         + a_hi.astype(np.float64) @ b_lo.T.astype(np.float64))
    qb = (np.linalg.norm(a.astype(np.float64), axis=1) * (1 + 1e-6)).astype(np.float32) + np.float32(1e-30)
    cb = (np.linalg.norm(b.astype(np.float64), axis=1) * (1 + 1e-6)).astype(np.float32) + np.float32(1e-30)
    infl = np.float32(1 + 2.0 ** -19)
    if l2:
        cn = (b.astype(np.float64) ** 2).sum(1).astype(np.float32)
        qn = (a.astype(np.float64) ** 2).sum(1).astype(np.float32)
        rest = (-0.5 * cn).astype(np.float32)
        pieces = np.zeros_like(rest, dtype=np.float64)
        for _ in range(3):
            p = _bf16_round(rest)
            pieces += p
            rest = (rest - p).astype(np.float32)
        f = f + pieces[None, :]
    f = f.astype(np.float32)                                # the fp32 accumulator the epilogue reads
    up = lambda x: np.nextafter(x.astype(np.float32), np.float32(np.inf))       # noqa: E731
    down = lambda x: np.nextafter(x.astype(np.float32), np.float32(-np.inf))    # noqa: E731
    cbmax = cb.reshape(-1, 32).max(1).repeat(32)            # largest bound of each 32-candidate block
    if l2:
        k1 = (np.float32(2) * (gamma + gamma2) * qb * infl).astype(np.float32)
        k0 = (gamma2 * qb * qb * infl).astype(np.float32)
        E = ((cbmax[None, :] * (gamma2 * infl * cbmax[None, :] + k1[:, None]) + k0[:, None]) * infl).astype(np.float32)
        base = (qn + st).astype(np.float32)[:, None]
        t_hi = up(up(base + E) * np.float32(0.5))
        t_lo = down(down(base - E) * np.float32(0.5))
    else:
        E = ((gamma * qb * infl)[:, None] * cbmax[None, :]).astype(np.float32)
        t_hi = up(st[:, None] + E)
        t_lo = down(st[:, None] - E)
    gt, lt = f > t_hi, f < t_lo
    assert (S[gt] >= np.broadcast_to(st[:, None], S.shape)[gt]).all()
    assert (S[lt] < np.broadcast_to(st[:, None], S.shape)[lt]).all()
    amb = ~(gt | lt)
    assert amb[np.arange(nq), true_idx].all()               # the true entity itself is always rechecked
    assert amb.mean() < 0.02
