"""Tensor-core bound-and-refine path (csrc/tc.cu): the approximate scores must stay inside
their error bound, and ranks must remain bit-identical to the oracle's."""
import ctypes

import numpy as np
import pytest
import torch

import torchkge_b200 as tk
from oracle import kge_oracle as oracle
from tests import helpers
from torchkge_b200 import _lib
from torchkge_b200.engine import CudaEngine, ModelSpec, rank_link_prediction
from torchkge_b200.data import filter_csr

pytestmark = pytest.mark.gpu
TC_KINDS = ["transe_l2", "distmult", "complex", "rescal"]


def _approx_scores(eng, spec, side, h, t, r):
    hrows, trows = eng.gather_rows(spec, h), eng.gather_rows(spec, t)
    packed, tcp = eng.pack(spec), eng.pack_tc(spec)
    assert tcp is not None
    n = h.shape[0]
    dump = torch.full((n, spec.n_rows), float("nan"), device=h.device)
    raw = torch.zeros(n, dtype=torch.int32, device=h.device)
    sub = torch.zeros_like(raw)
    eng.rank_side(spec, packed, side, hrows, trows, r, t if side == 0 else h, None, raw, sub,
                  tc_packed=tcp, tc_dump=dump)
    torch.cuda.synchronize()
    return dump.cpu()


@pytest.fixture(params=[0, 1], ids=["bf16", "fp16"])
def operand_format(request):
    """Both operand formats of the split (csrc/tc.h): restored to the environment's default afterwards."""
    default = _lib.tc_bound_constants(_lib.DISTMULT, 16)[3]
    _lib.tc_configure(fp16=request.param)
    yield request.param
    _lib.tc_configure(fp16=int(default))


def _prefix_factor(x, k_total=None):
    """P(x) = sqrt(sum_i |x_{<= 16 i}|^2) per row over the ceil(k_total / 16) MMA k-steps
    (csrc/tc.h: tc_gamma_p), x (n, k) float64; k_total > k: trailing k-steps see the whole |x|^2."""
    k = x.shape[1]
    pad = (-k) % 16
    sq = torch.nn.functional.pad(x * x, (0, pad)).view(x.shape[0], -1, 16).sum(2)
    pre = torch.cumsum(sq, 1)
    extra = 0 if k_total is None else (k_total + 15) // 16 - (k + 15) // 16
    return (pre.sum(1) + extra * pre[:, -1]).sqrt()


def _kernel_bound(code, d, k_total, na, nb, max_a, max_b, max_norm2_b, l2, pa, pb):
    """The bound E(q, c) the scan's threshold test uses (csrc/tc.cu epilogue, csrc/tc.h), from the
    library's own constants: na (nq, 1), nb (1, nc) row norms; pa (nq, 1), pb (1, nc) the
    running-magnitude factors of the operands."""
    gamma, gamma2, gamma_p, fp16 = _lib.tc_bound_constants(code, d)
    e_abs = 0.0
    if fp16:
        na = na + max_a * 2.0 ** -11 * (k_total ** 0.5) * 1.01 / 3.0
        nb = nb + max_b * 2.0 ** -11 * (k_total ** 0.5) * 1.01 / 3.0
        if l2 and max_norm2_b > 0:
            import math
            phi = 2.0 ** (14 - math.frexp(0.5 * max_norm2_b)[1])
            e_abs = 2.0 ** -22 / phi
    if l2:
        return 2 * gamma * na * nb + 2 * gamma_p * pa * pb + gamma2 * (na + nb) ** 2 + e_abs
    return gamma * na * nb + gamma_p * pa * pb


@pytest.mark.parametrize("kind", TC_KINDS)
@pytest.mark.parametrize("d", [13, 50, 64, 200])
def test_approximate_scores_within_half_the_bound(kind, d, cuda_device, operand_format):
    if kind == "rescal" and d > 64:
        pytest.skip("rescal d^2 tables: small dims only")
    n_ent, n_rel, b = 1300, 7, 150
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=d)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.0 + torch.rand(p.shape[0], 1))
    model = model.to(cuda_device)
    P = helpers.oracle_params(kind, model)
    g = torch.Generator().manual_seed(3)
    h = torch.randint(0, n_ent, (b,), generator=g)
    t = torch.randint(0, n_ent, (b,), generator=g)
    r = torch.randint(0, n_rel, (b,), generator=g)
    eng = CudaEngine(tensor_core=True)
    spec = ModelSpec.from_model(model)
    for side, name in ((0, "tail"), (1, "head")):
        got = _approx_scores(eng, spec, side, h.to(cuda_device), t.to(cuda_device), r.to(cuda_device))
        want = oracle.scores_all(kind, P, h, t, r, name).double()
        assert not torch.isnan(got).any()
        l2 = kind == "transe_l2"
        k_total = 2 * d if kind == "complex" else (d + 3 if l2 else d)  # csrc/api.cu: tc_k_total
        # operand norms
        if kind == "complex":
            cand = torch.cat([P["re_ent"], P["im_ent"]], 1).double()
            re_h, im_h, re_t, im_t = P["re_ent"][h], P["im_ent"][h], P["re_ent"][t], P["im_ent"][t]
            re_r, im_r = P["re_rel"][r], P["im_rel"][r]
            q = (torch.cat([re_h * re_r - im_h * im_r, re_h * im_r + im_h * re_r], 1) if side == 0 else
                 torch.cat([re_r * re_t + im_r * im_t, re_r * im_t - im_r * re_t], 1)).double()
        elif kind == "distmult":
            cand = P["ent"].double()
            q = ((P["ent"][h] * P["rel"][r]) if side == 0 else (P["rel"][r] * P["ent"][t])).double()
        elif kind == "rescal":
            cand = P["ent"].double()
            m = P["rel_mat"][r].view(-1, d, d)
            q = (torch.matmul(P["ent"][h].view(b, 1, d), m).view(b, d) if side == 0 else
                 torch.matmul(m, P["ent"][t].view(b, d, 1)).view(b, d)).double()
        else:
            cand = P["ent"].double()
            q = ((P["ent"][h] + P["rel"][r]) if side == 0 else (P["ent"][t] - P["rel"][r])).double()
        na, nb = q.norm(dim=1).view(-1, 1), cand.norm(dim=1).view(1, -1)
        pa, pb = _prefix_factor(q, k_total).view(-1, 1), _prefix_factor(cand, k_total).view(1, -1)
        if l2 and side == 1:   # head side: the kernel bounds |t - r| by |t| + |r| (see tc.cu)
            nt = (P["ent"][t].double().norm(dim=1) + P["rel"][r].double().norm(dim=1)).view(-1, 1)
            pa = pa * nt / na.clamp_min(1e-300)
            na = nt
        bound = _kernel_bound(spec.code, d, k_total, na, nb, q.abs().max().item(), cand.abs().max().item(),
                              (cand ** 2).sum(1).max().item(), l2, pa, pb)
        ratio = ((got.double() - want).abs() / bound).max().item()
        assert ratio < 0.5, "%s %s d=%d: error / bound = %.3f" % (kind, name, d, ratio)


def _operands(shape, n_q, n_c, d, g):
    """Adversarial operand families for the error bound (the CPU twin is tests/test_tc_bound_cpu.py)."""
    a = torch.randn(n_q, d, generator=g)
    b = torch.randn(n_c, d, generator=g)
    if shape == "cancelling":        # large products of alternating sign: running sums far below sum |terms|
        sign = torch.where(torch.arange(d) % 2 == 0, 1.0, -1.0)
        a, b = a.abs(), b.abs() * sign
    elif shape == "same_sign":       # every product positive: the accumulator grows monotonically (truncation bias)
        a, b = a.abs(), b.abs()
    elif shape == "wide":            # 2^-10 .. 2^10 per element
        a = a * torch.exp2(torch.randint(-10, 11, (n_q, d), generator=g).float())
        b = b * torch.exp2(torch.randint(-10, 11, (n_c, d), generator=g).float())
    elif shape == "one_big":         # one dominant product per 16-term instruction, 15 small ones below its ulp
        a, b = a.abs() * 2.0 ** -13, b.abs() * 2.0 ** -13
        a[:, ::16] = 1.0 + a[:, ::16]
        b[:, ::16] = 1.0 + b[:, ::16]
    elif shape == "normalised":
        a = torch.nn.functional.normalize(a, dim=1)
        b = torch.nn.functional.normalize(b, dim=1)
    return a.contiguous(), b.contiguous()


@pytest.mark.parametrize("shape", ["normalised", "cancelling", "same_sign", "wide", "one_big"])
@pytest.mark.parametrize("kind,d", [("distmult", 203), ("distmult", 800), ("distmult", 1024),
                                    ("transe_l2", 200), ("transe_l2", 797), ("complex", 400)])
def test_error_bound_holds_on_adversarial_operands(shape, kind, d, cuda_device, operand_format):
    """|tensor-core score - reference fp32 score| <= the kernel's bound, pair by pair, on operand
    families built to stress each term of the bound: K = 203, 800, 1024 (13, 50, 64 MMA k-steps x 3
    instructions).  Queries are table rows with an identity relation, so the operands are exactly
    the generated vectors."""
    g = torch.Generator().manual_seed(1000 + d)
    n_q, n_c = 96, 1100
    dev = cuda_device
    code = {"distmult": _lib.DISTMULT, "transe_l2": _lib.TRANSE_L2, "complex": _lib.COMPLEX}[kind]
    l2 = kind == "transe_l2"
    if kind == "complex":
        a0, b0 = _operands(shape, n_q, n_c, d, g)
        a1, b1 = _operands(shape, n_q, n_c, d, g)
        ent0, ent1 = torch.cat([b0, a0]), torch.cat([b1, a1])
        P = {"re_ent": ent0, "im_ent": ent1, "re_rel": torch.ones(1, d), "im_rel": torch.zeros(1, d)}
        spec = ModelSpec(code, d, n_c, 1, ent0[:n_c].contiguous().to(dev), ent1[:n_c].contiguous().to(dev),
                         P["re_rel"].to(dev), P["im_rel"].to(dev))
        hrows = torch.stack([a0, a1], 1).contiguous().to(dev)        # (n_q, 2, d); q = h o (1 + 0i) = h
        qv, cv = torch.cat([a0, a1], 1).double(), torch.cat([b0, b1], 1).double()
        k_total = 2 * d
    else:
        a, b = _operands(shape, n_q, n_c, d, g)
        rel = torch.ones(1, d) if kind == "distmult" else torch.zeros(1, d)   # h * 1 = h ; h + 0 = h
        ent = torch.cat([b, a])
        P = {"ent": ent, "rel": rel}
        spec = ModelSpec(code, d, n_c, 1, ent[:n_c].contiguous().to(dev), None, rel.to(dev), None)
        hrows = a.view(n_q, 1, d).contiguous().to(dev)
        qv, cv = a.double(), b.double()
        k_total = d + 3 if l2 else d
    h_idx = torch.arange(n_c, n_c + n_q)
    r_idx = torch.zeros(n_q, dtype=torch.int64)
    want = oracle.scores_all(kind, P, h_idx, h_idx, r_idx, "tail")[:, :n_c].double()
    eng = CudaEngine(tensor_core=True)
    tcp = eng.pack_tc(spec)
    dump = torch.full((n_q, n_c), float("nan"), device=dev)
    raw = torch.zeros(n_q, dtype=torch.int32, device=dev)
    eng.rank_side(spec, None, _lib.SIDE_TAIL, hrows, hrows, r_idx.to(dev), r_idx.to(dev), None, raw,
                  torch.zeros_like(raw), tc_packed=tcp, tc_dump=dump)
    torch.cuda.synchronize()
    got = dump.cpu().double()
    assert torch.isfinite(got).all()
    na, nb = qv.norm(dim=1).view(-1, 1), cv.norm(dim=1).view(1, -1)
    bound = _kernel_bound(code, d, k_total, na, nb, qv.abs().max().item(), cv.abs().max().item(),
                          (cv ** 2).sum(1).max().item(), l2, _prefix_factor(qv, k_total).view(-1, 1),
                          _prefix_factor(cv, k_total).view(1, -1))
    ratio = ((got - want).abs() / bound).max().item()
    assert ratio <= 1.0, "%s %s d=%d: error / bound = %.3f" % (kind, shape, d, ratio)
    print("tc bound %s %s d=%d fp16=%d: max error / bound = %.3f" % (kind, shape, d, operand_format, ratio))


@pytest.mark.parametrize("kind", TC_KINDS)
@pytest.mark.parametrize("d", [13, 50, 100])
def test_ranks_equal_oracle_with_tensor_cores(kind, d, cuda_device, operand_format):
    if kind == "rescal" and not helpers.rescal_order_matches_here(d):
        pytest.skip("oneMKL on this CPU sums RESCAL's batched matmul in another order than the authoring machine")
    n_ent, n_rel = 1500, 9
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=7000, n_test=300, seed=100 + d)
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=d).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 64)
    eng = CudaEngine(tensor_core=True)
    spec = ModelSpec.from_model(model)
    dev = cuda_device
    csr_t = tuple(x.to(dev) for x in filter_csr(dt, kg.head_idx, kg.relations, kg.tail_idx))
    csr_h = tuple(x.to(dev) for x in filter_csr(dh, kg.tail_idx, kg.relations, kg.head_idx))
    got = rank_link_prediction(spec, kg.head_idx.to(dev), kg.tail_idx.to(dev), kg.relations.to(dev),
                               csr_t, csr_h, engine=eng)
    assert len(eng.tc_stats) == 2
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)
    found = sum(int(s[0]) for s in eng.tc_stats)
    assert found < 0.02 * 2 * 300 * n_ent      # the near-tie band is a small fraction


def test_exact_ties_and_duplicates_with_tensor_cores(cuda_device):
    n_ent, n_rel, d = 600, 5, 32
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=3000, n_test=200, seed=9)
    model = helpers.make_model("distmult", d, n_ent, n_rel, seed=9)
    with torch.no_grad():
        model.ent_emb.weight[100:300] = model.ent_emb.weight[0:200].clone()   # exact ties
        model.ent_emb.weight[500:] = 0.0
    P = helpers.oracle_params("distmult", model)
    ref = oracle.link_prediction("distmult", P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 50)
    ev = tk.LinkPredictionEvaluator(model.to(cuda_device), kg)
    ev.evaluate(b_size=64, verbose=False)
    for a, b in zip((ev.rank_true_heads, ev.rank_true_tails, ev.filt_rank_true_heads,
                     ev.filt_rank_true_tails), ref):
        assert torch.equal(a, b)


# (bk, resident, ct_group, max_ctas): every geometry of the scan, with grids small enough that
# each CTA walks many work units (ring phases wrap, the resident query image is replaced, both
# TMEM accumulators are reused)
TC_CONFIGS = [(32, 1, 1, 3), (32, 1, 2, 5), (32, 1, 0, 0), (32, 0, 2, 4), (64, 0, 1, 3), (64, 0, 0, 0)]


@pytest.mark.parametrize("cfg", TC_CONFIGS, ids=lambda c: "bk%d_res%d_grp%d_ctas%d" % c)
@pytest.mark.parametrize("kind,d", [("transe_l2", 200), ("distmult", 72), ("complex", 120)])
def test_every_scan_geometry_gives_oracle_ranks(cfg, kind, d, cuda_device):
    n_ent, n_rel = 2700, 6
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=9000, n_test=520, seed=41 + d)
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=d).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 128)
    dev = cuda_device
    csr_t = tuple(x.to(dev) for x in filter_csr(dt, kg.head_idx, kg.relations, kg.tail_idx))
    csr_h = tuple(x.to(dev) for x in filter_csr(dh, kg.tail_idx, kg.relations, kg.head_idx))
    try:
        _lib.tc_configure(*cfg)
        eng = CudaEngine(tensor_core=True)
        spec = ModelSpec.from_model(model)
        got = rank_link_prediction(spec, kg.head_idx.to(dev), kg.tail_idx.to(dev), kg.relations.to(dev),
                                   csr_t, csr_h, engine=eng)
        torch.cuda.synchronize()
    finally:
        _lib.tc_configure(32, 1, 0, 0)
    assert len(eng.tc_stats) == 2
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("d", [16, 100, 1000])
@pytest.mark.parametrize("refine", [True, False])
def test_rotate_bound_and_refine_gives_oracle_ranks(d, refine, cuda_device):
    """RotatE has no tensor-core form: its bound-and-refine (KGE_FLAG_APPROX_SCAN) runs on the fp32
    pipes with approximate square roots; ranks must equal the oracle's with it on and off, and the
    near-tie band must stay small."""
    # the CPU oracle's RotatE is slow (stack + norm over (b, n_ent, d)): keep d * n_ent * n_test small
    n_ent, n_rel, n_test = (1500, 9, 300) if d <= 100 else (400, 5, 70)
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=7000 if d <= 100 else 1500, n_test=n_test, seed=200 + d)
    model = helpers.make_model("rotate", d, n_ent, n_rel, seed=d)
    with torch.no_grad():
        model.re_ent_emb.weight[100:200] = model.re_ent_emb.weight[0:100].clone()   # exact ties
        model.im_ent_emb.weight[100:200] = model.im_ent_emb.weight[0:100].clone()
    model = model.to(cuda_device)
    P = helpers.oracle_params("rotate", model)
    ref = oracle.link_prediction("rotate", P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 64)
    eng = CudaEngine(tensor_core=refine)
    spec = ModelSpec.from_model(model)
    dev = cuda_device
    csr_t = tuple(x.to(dev) for x in filter_csr(dt, kg.head_idx, kg.relations, kg.tail_idx))
    csr_h = tuple(x.to(dev) for x in filter_csr(dh, kg.tail_idx, kg.relations, kg.head_idx))
    got = rank_link_prediction(spec, kg.head_idx.to(dev), kg.tail_idx.to(dev), kg.relations.to(dev),
                               csr_t, csr_h, engine=eng)
    for a, b in zip(got, ref):
        assert torch.equal(a.cpu(), b)
    if refine:
        assert len(eng.tc_stats) == 2
        found = sum(int(s[0]) for s in eng.tc_stats)
        assert 2 * n_test <= found < 0.05 * 2 * n_test * n_ent     # at least the true entities themselves
    else:
        assert len(eng.tc_stats) == 0


def test_cached_operand_image_follows_in_place_weight_updates(cuda_device):
    """The engine keeps the tensor-core operand image between evaluations, guarded by a device-side
    content checksum of the table: in-place updates that leave data_ptr and _version untouched
    (``weight.data.mul_``, an optimizer step on ``.data``) must be picked up, an unchanged table must
    not be rebuilt (same ranks either way)."""
    import torchkge_b200.engine as engine_mod
    n_ent, n_rel, d = 900, 6, 48
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=5000, n_test=260, seed=77)
    model = helpers.make_model("complex", d, n_ent, n_rel, seed=7).to(cuda_device)
    eng = CudaEngine(tensor_core=True)
    assert eng.tc_cache_entries > 0
    old = engine_mod._default_engine
    engine_mod._default_engine = eng
    try:
        for step in range(3):
            if step == 1:
                pass                                            # unchanged table: served from the cache
            if step == 2:
                with torch.no_grad():
                    model.re_ent_emb.weight.data[::3] *= -1.5   # same storage, same _version
                    model.im_ent_emb.weight.data[5] = 0.25
            ev = tk.LinkPredictionEvaluator(model, kg)
            ev.evaluate(b_size=64, verbose=False)
            P = helpers.oracle_params("complex", model)
            ref = oracle.link_prediction("complex", P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 64)
            for a, b in zip((ev.rank_true_heads, ev.rank_true_tails, ev.filt_rank_true_heads,
                             ev.filt_rank_true_tails), ref):
                assert torch.equal(a, b), "step %d" % step
            assert len(eng._tc_cache) == 1
    finally:
        engine_mod._default_engine = old
