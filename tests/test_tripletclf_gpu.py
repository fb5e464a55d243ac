"""PositionalNegativeSampler (torchkge/sampling.py:330-503) and TripletClassificationEvaluator
(torchkge/evaluation.py:428-580) on the GPU.  The sampler is checked for its law (support, one
end corrupted, Bernoulli head frequency); the evaluator's thresholds and accuracy are checked
against a CPU restatement with the oracle's scores on the SAME negatives."""
import pytest
import torch

import torchkge_b200 as tk
from oracle import kge_oracle as oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def _graphs(n_ent=400, n_rel=6, seed=4):
    h, t, r = helpers.random_graph(n_ent, n_rel, 5000, seed=seed)
    mk = lambda a, b: tk.KnowledgeGraph(h[a:b], t[a:b], r[a:b], n_ent, n_rel, dict_of_heads={}, dict_of_tails={})  # noqa: E731
    return mk(0, 3000), mk(3000, 4000), mk(4000, 5000)


def test_positional_sampler_law(cuda_device):
    kg, kg_val, kg_test = _graphs()
    s = tk.PositionalNegativeSampler(kg, kg_val=kg_val, kg_test=kg_test, seed=3)
    ph, pt = s.possible_heads, s.possible_tails
    both = torch.cat([kg.relations, kg_val.relations])
    for rel in range(kg.n_rel):
        m = both == rel
        assert set(ph[rel]) == set(torch.cat([kg.head_idx, kg_val.head_idx])[m].tolist())
        assert set(pt[rel]) == set(torch.cat([kg.tail_idx, kg_val.tail_idx])[m].tolist())
        assert s.n_poss_heads[rel].item() == len(ph[rel])
    reps = 20
    h = kg.head_idx.repeat(reps).to(cuda_device)
    t = kg.tail_idx.repeat(reps).to(cuda_device)
    r = kg.relations.repeat(reps).to(cuda_device)
    nh, nt = s.corrupt_batch(h, t, r)
    assert nh.dtype == torch.int64 and nh.shape == h.shape and nh.device == h.device
    hc, tc_ = nh != h, nt != t
    assert not (hc & tc_).any()
    nh_c, nt_c, r_c = nh.cpu(), nt.cpu(), r.cpu()
    for rel in range(kg.n_rel):
        m = r_c == rel
        assert set(nh_c[m & hc.cpu()].tolist()) <= set(ph[rel])
        assert set(nt_c[m & tc_.cpu()].tolist()) <= set(pt[rel])
        n = int(m.sum())
        p = s.bern_probs[rel].item()
        # a draw may coincide with the original entity with probability 1 / n_poss
        slack = 1.0 / max(1, len(ph[rel]))
        f = hc.cpu()[m].float().mean().item()
        assert p - slack - 5 * (0.25 / n) ** 0.5 <= f <= p + 5 * (0.25 / n) ** 0.5
    ch, ct = s.corrupt_kg(512, True, which="test")
    assert ch.shape == (kg_test.n_facts,) and not ch.is_cuda


@pytest.mark.parametrize("kind", ["transe_l2", "distmult", "complex"])
def test_triplet_classification_matches_cpu_restatement(kind, cuda_device):
    kg, kg_val, kg_test = _graphs(seed=5)
    model = helpers.make_model(kind, 32, kg.n_ent, kg.n_rel, seed=1).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ev = tk.TripletClassificationEvaluator(model, kg_val, kg_test)
    gen = torch.Generator().manual_seed(9)
    fixed = {"main": (torch.randint(0, kg.n_ent, (kg_val.n_facts,), generator=gen), kg_val.tail_idx.clone()),
             "test": (kg_test.head_idx.clone(), torch.randint(0, kg.n_ent, (kg_test.n_facts,), generator=gen))}

    class Fixed:
        def corrupt_kg(self, batch_size, use_cuda, which="main"):
            return fixed[which]

    ev.sampler = Fixed()
    ev.evaluate(b_size=300)
    neg = oracle.score_triples(kind, P, fixed["main"][0], fixed["main"][1], kg_val.relations)
    want = torch.zeros(kg.n_rel)
    for i in range(kg.n_rel):
        m = kg_val.relations == i
        want[i] = neg[m].max() if m.sum() > 0 else neg.max()
    torch.testing.assert_close(ev.thresholds.cpu(), want, rtol=2e-5, atol=2e-6)
    acc = ev.accuracy(b_size=300)
    pos_s = oracle.score_triples(kind, P, kg_test.head_idx, kg_test.tail_idx, kg_test.relations)
    neg_s = oracle.score_triples(kind, P, fixed["test"][0], fixed["test"][1], kg_test.relations)
    thr = want[kg_test.relations]
    margin = 1e-4 * (1 + thr.abs())       # facts within rounding distance of a threshold may flip
    lo = ((pos_s > thr + margin).sum() + (neg_s < thr - margin).sum()).item() / (2 * kg_test.n_facts)
    hi = ((pos_s > thr - margin).sum() + (neg_s < thr + margin).sum()).item() / (2 * kg_test.n_facts)
    assert lo <= acc <= hi
    # the default sampler is the reference's
    assert isinstance(tk.TripletClassificationEvaluator(model, kg_val, kg_test).sampler,
                      tk.PositionalNegativeSampler)
