"""TorusE (torus_L1 / torus_L2) link prediction on the GPU against the unmodified reference's
golden outputs (tests/golden/torus_*.npz) and the oracle."""
import pytest
import torch

import torchkge_b200 as tk
from oracle import kge_oracle as oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def _model(g, dev):
    diss = "torus_L1" if g["kind"] == "toruse_l1" else "torus_L2"
    model = tk.TorusEModel(g["dim"], g["n_ent"], g["n_rel"], diss)
    model.load_state_dict(g["state"])
    return model.to(dev)


@pytest.mark.parametrize("case", helpers.TORUS_CASES)
def test_ranks_equal_the_unmodified_reference(case, cuda_device):
    g = helpers.load_golden_torus(case)
    model = _model(g, cuda_device)
    kg = tk.KnowledgeGraph(g["heads"], g["tails"], g["rels"], g["n_ent"], g["n_rel"],
                           dict_of_heads=g["dh"], dict_of_tails=g["dt"])
    ev = tk.LinkPredictionEvaluator(model, kg)
    ev.evaluate(b_size=g["b_size"], verbose=False)
    for nm in ("rank_true_heads", "rank_true_tails", "filt_rank_true_heads", "filt_rank_true_tails"):
        assert torch.equal(getattr(ev, nm), torch.from_numpy(g["raw"][nm])), nm
    h, t, r = (g[k][:8].to(cuda_device) for k in ("heads", "tails", "rels"))
    he, te, re_, cands = model.inference_prepare_candidates(h, t, r, entities=True)
    assert helpers.bits_equal(model.inference_scoring_function(he, cands, re_),
                              torch.from_numpy(g["raw"]["scores_tail"])).all()
    assert helpers.bits_equal(model.inference_scoring_function(cands, te, re_),
                              torch.from_numpy(g["raw"]["scores_head"])).all()
    got = model.scoring_function(g["heads"].to(cuda_device), g["tails"].to(cuda_device), g["rels"].to(cuda_device))
    torch.testing.assert_close(got.cpu(), torch.from_numpy(g["raw"]["triple_scores"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind,d", [("toruse_l1", 13), ("toruse_l2", 64), ("toruse_l1", 520)])
def test_ranks_equal_oracle(kind, d, cuda_device):
    n_ent, n_rel = 700, 6
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=4000, n_test=200, seed=31 + d)
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=d)
    with torch.no_grad():
        model.ent_emb.weight.mul_(37.0)
        model.rel_emb.weight.mul_(37.0)
        model.normalize_parameters()
        model.ent_emb.weight[50:90] = model.ent_emb.weight[0:40].clone()   # exact ties
    model = model.to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 64)
    ev = tk.LinkPredictionEvaluator(model, kg)
    ev.evaluate(b_size=64, verbose=False)
    for got, want in zip((ev.rank_true_heads, ev.rank_true_tails, ev.filt_rank_true_heads,
                          ev.filt_rank_true_tails), ref):
        assert torch.equal(got, want)


@pytest.mark.parametrize("diss", ["torus_L1", "torus_L2"])
def test_toruse_training_kernels_match_the_reference_expression(diss, cuda_device):
    """scoring_function (translation.py:706-720: -dissimilarity(frac(h) + frac(r), frac(t))) and its
    gradients from the per-triple CUDA kernels against the same expression in torch on the CPU
    (1e-5 relative on scores, rtol 1e-4 on dense gradients), and the fused margin step on top."""
    from torchkge_b200.training import fused_margin_step
    n_ent, n_rel, d, n = 300, 7, 37, 400
    torch.manual_seed(5)
    model = tk.TorusEModel(d, n_ent, n_rel, diss)
    with torch.no_grad():                      # values beyond one period, both signs
        model.ent_emb.weight.copy_((torch.rand(n_ent, d) - 0.5) * 5)
        model.rel_emb.weight.copy_((torch.rand(n_rel, d) - 0.5) * 5)
    g = torch.Generator().manual_seed(1)
    h = torch.randint(0, n_ent, (n,), generator=g)
    t = torch.randint(0, n_ent, (n,), generator=g)
    r = torch.randint(0, n_rel, (n,), generator=g)
    E = model.ent_emb.weight.detach().clone().requires_grad_(True)
    R = model.rel_emb.weight.detach().clone().requires_grad_(True)
    fn = tk.models.l1_torus_dissimilarity if diss == "torus_L1" else tk.models.l2_torus_dissimilarity
    hh, tt, rr = E[h], E[t], R[r]
    hf, tf, rf = hh - hh.detach().trunc(), tt - tt.detach().trunc(), rr - rr.detach().trunc()   # frac, unit slope
    want = -fn(hf + rf, tf)
    w = torch.rand(n, generator=g)
    (want * w).sum().backward()
    model = model.to(cuda_device)
    got = model.scoring_function(h.to(cuda_device), t.to(cuda_device), r.to(cuda_device))
    torch.testing.assert_close(got.cpu(), want.detach(), rtol=1e-5, atol=1e-5)
    (got * w.to(cuda_device)).sum().backward()
    torch.testing.assert_close(model.ent_emb.weight.grad.cpu(), E.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(model.rel_emb.weight.grad.cpu(), R.grad, rtol=1e-4, atol=1e-5)
    # fused step with supplied negatives == scoring + MarginLoss composed
    nh = torch.randint(0, n_ent, (2 * n,), generator=g)
    nt = t.repeat(2)
    loss = fused_margin_step(model, h.to(cuda_device), t.to(cuda_device), r.to(cuda_device), 0.5,
                             negatives=(nh.to(cuda_device), nt.to(cuda_device)))
    pos, neg = model(h.to(cuda_device), t.to(cuda_device), r.to(cuda_device), nh.to(cuda_device), nt.to(cuda_device))
    ref_loss = tk.MarginLoss(0.5)(pos, neg)
    torch.testing.assert_close(loss, ref_loss, rtol=1e-5, atol=1e-5)
