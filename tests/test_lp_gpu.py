"""GPU parity tests of the link-prediction path: CUDA engine (through the public classes and
the C ABI) against the CPU oracle on the same seeded inputs.

Bar (north_star): ranks bit-identical, ties broken identically; scores within 1e-5 relative
-- and in fact required bit-identical here: every reduction is an ATen / oneMKL order we replay
(RESCAL's query preparation `h^T M_r`, `M_r t` goes through MKL's batched GEMM on the CPU side; its
summation order is replayed too, see csrc/reduce.cuh:rescal_query_component and DESIGN.md 2.4).
"""
import numpy as np
import pytest
import torch

import torchkge_b200 as tk
from oracle import kge_oracle as oracle
from tests import helpers

pytestmark = pytest.mark.gpu

EXACT_KINDS = ["transe_l1", "transe_l2", "distmult", "complex", "rotate"]
ALL_KINDS = EXACT_KINDS + ["rescal"]


def _ranks_gpu(model, kg, b_size=64):
    ev = tk.LinkPredictionEvaluator(model, kg)
    ev.evaluate(b_size=b_size, verbose=False)
    return ev


def _assert_ranks_equal(ev, ref, kind):
    names = ["rank_true_heads", "rank_true_tails", "filt_rank_true_heads", "filt_rank_true_tails"]
    for name, want in zip(names, ref):
        got = getattr(ev, name)
        assert got.dtype == torch.long and got.shape == want.shape
        bad = (got != want).nonzero().flatten()
        assert bad.numel() == 0, "%s %s: %d / %d ranks differ, first at %d: got %d want %d" % (
            kind, name, bad.numel(), want.numel(), bad[0], got[bad[0]], want[bad[0]])


def _needs_reference_mkl_order(kind, d):
    """RESCAL vs an oracle computed on THIS machine: only meaningful where this machine's MKL sums
    the query preparation like the authoring machine's (the committed golden fixtures are compared
    unconditionally in test_ranks_equal_the_unmodified_reference)."""
    if kind == "rescal" and not helpers.rescal_order_matches_here(d):
        pytest.skip("oneMKL on this CPU sums RESCAL's batched matmul in another order than the machine "
                    "the golden fixtures come from: the reference's own bits differ here")


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("d", [50, 64, 100, 13])
def test_ranks_match_oracle_exactly(kind, d, cuda_device):
    if kind == "rescal" and d > 64:
        pytest.skip("rescal d^2 tables: small dims only")
    _needs_reference_mkl_order(kind, d)
    n_ent, n_rel = 1000, 11
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=6000, n_test=300, seed=d)
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=d).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, b_size=64)
    ev = _ranks_gpu(model, kg)
    _assert_ranks_equal(ev, ref, kind)


@pytest.mark.parametrize("kind", ALL_KINDS)
def test_dense_scores_match_oracle(kind, cuda_device):
    """inference_prepare_candidates + inference_scoring_function == oracle scores."""
    n_ent, n_rel, d = 777, 7, 40
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=3).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    g = torch.Generator().manual_seed(5)
    b = 37
    h = torch.randint(0, n_ent, (b,), generator=g)
    t = torch.randint(0, n_ent, (b,), generator=g)
    r = torch.randint(0, n_rel, (b,), generator=g)
    he, te, re_, cands = model.inference_prepare_candidates(h.to(cuda_device), t.to(cuda_device),
                                                            r.to(cuda_device), entities=True)
    for side, args in (("tail", (he, cands, re_)), ("head", (cands, te, re_))):
        got = model.inference_scoring_function(*args).cpu()
        want = oracle.scores_all(kind, P, h, t, r, side)
        assert got.shape == want.shape == (b, n_ent)
        if kind == "rescal" and not helpers.rescal_order_matches_here(d):
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)   # this CPU's MKL sums differently
        else:
            same = (got.numpy().view(np.uint32) == want.numpy().view(np.uint32)) | (got == want).numpy()
            assert same.all(), "%s %s: %d scores differ in bits (max abs diff %g)" % (
                kind, side, (~same).sum(), (got - want).abs().max())


@pytest.mark.parametrize("kind", ["transe_l2", "distmult", "complex"])
def test_ties_zero_rows_and_duplicates(kind, cuda_device):
    """Exact ties must count against the true entity, everywhere the same way."""
    n_ent, n_rel, d = 300, 5, 32
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=2000, n_test=200, seed=9)
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=9)
    with torch.no_grad():
        for emb in [m for n, m in model.named_children() if "ent" in n]:
            emb.weight[100:200] = emb.weight[0:100]      # duplicate rows -> exact ties
            emb.weight[250:] = 0.0                       # zero rows -> many equal scores
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, b_size=50)
    ev = _ranks_gpu(model.to(cuda_device), kg)
    _assert_ranks_equal(ev, ref, kind)


def test_all_zero_table_rank_is_n_ent(cuda_device):
    n_ent, n_rel, d = 200, 3, 16
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=500, n_test=100, seed=1)
    model = helpers.make_model("distmult", d, n_ent, n_rel)
    with torch.no_grad():
        model.ent_emb.weight.zero_()
    ev = _ranks_gpu(model.to(cuda_device), kg)
    assert (ev.rank_true_heads == n_ent).all() and (ev.rank_true_tails == n_ent).all()
    P = helpers.oracle_params("distmult", model)
    ref = oracle.link_prediction("distmult", P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 64)
    _assert_ranks_equal(ev, ref, "distmult")


def test_true_entity_missing_from_filter_dict_means_no_filtering(cuda_device):
    """get_true_targets quirk (modeling.py:78-88): KeyError on remove => row unfiltered."""
    n_ent, n_rel, d = 150, 4, 24
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=1500, n_test=120, seed=4)
    # drop the true tail of the first 40 test facts from its own filter set
    for i in range(40):
        key = (kg.head_idx[i].item(), kg.relations[i].item())
        dt[key].discard(kg.tail_idx[i].item())
    model = helpers.make_model("transe_l1", d, n_ent, n_rel, seed=2)
    P = helpers.oracle_params("transe_l1", model)
    ref = oracle.link_prediction("transe_l1", P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 32)
    ev = _ranks_gpu(model.to(cuda_device), kg)
    _assert_ranks_equal(ev, ref, "transe_l1")


@pytest.mark.parametrize("b_size", [1, 7, 1000])
def test_b_size_never_changes_results(b_size, cuda_device):
    n_ent, n_rel, d = 400, 6, 20
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=1200, n_test=130, seed=8)
    model = helpers.make_model("complex", d, n_ent, n_rel, seed=8).to(cuda_device)
    a = _ranks_gpu(model, kg, b_size=b_size)
    b = _ranks_gpu(model, kg, b_size=64)
    for n in ("rank_true_heads", "rank_true_tails", "filt_rank_true_heads", "filt_rank_true_tails"):
        assert torch.equal(getattr(a, n), getattr(b, n))


def test_metrics_and_not_yet_evaluated(cuda_device):
    n_ent, n_rel, d = 120, 3, 16
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=600, n_test=60, seed=6)
    model = helpers.make_model("transe_l2", d, n_ent, n_rel).to(cuda_device)
    ev = tk.LinkPredictionEvaluator(model, kg)
    with pytest.raises(tk.NotYetEvaluatedError):
        ev.mean_rank()
    ev.evaluate(b_size=len(kg), verbose=False)
    mr, hit, mrr = oracle.lp_metrics(ev.rank_true_heads, ev.rank_true_tails,
                                     ev.filt_rank_true_heads, ev.filt_rank_true_tails, k=10)
    assert ev.mean_rank() == pytest.approx(mr)
    assert ev.hit_at_k(10) == pytest.approx(hit)
    assert ev.mrr() == pytest.approx(mrr)
    assert ev.filt_rank_true_heads.le(ev.rank_true_heads).all()


@pytest.mark.parametrize("kind,d", [("distmult", 520), ("complex", 512), ("transe_l2", 1001)])
def test_large_dim_cascade(kind, d, cuda_device):
    """dim >= 512 switches ATen's sum to a two-level cascade; dim % 8 != 0 adds scalar tails."""
    n_ent, n_rel = 300, 3
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=900, n_test=70, seed=12)
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=12).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, b_size=35)
    ev = _ranks_gpu(model, kg)
    _assert_ranks_equal(ev, ref, kind)


@pytest.mark.parametrize("kind", ["transe_l2", "transe_l1", "complex"])
def test_filter_index_graph_gives_reference_ranks(kind, cuda_device):
    """KnowledgeGraph(filter_facts=...) keeps the filter sets as sorted arrays resident on the
    device instead of Python dictionaries; ranks must not change."""
    n_ent, n_rel, d = 700, 8, 40
    h, t, r = helpers.random_graph(n_ent, n_rel, 5000, seed=31)
    dh, dt = oracle.build_filter_dicts(h, t, r)
    # test facts: 200 real ones plus 40 corrupted ones whose true entity is not in its filter set
    g = torch.Generator().manual_seed(1)
    th = torch.cat([h[:200], torch.randint(0, n_ent, (40,), generator=g)])
    tt = torch.cat([t[:200], torch.randint(0, n_ent, (40,), generator=g)])
    tr = torch.cat([r[:200], torch.randint(0, n_rel, (40,), generator=g)])
    kg = tk.KnowledgeGraph(th, tt, tr, n_ent, n_rel, filter_facts=(h, t, r))
    model = helpers.make_model(kind, d, n_ent, n_rel, seed=31).to(cuda_device)
    P = helpers.oracle_params(kind, model)
    ref = oracle.link_prediction(kind, P, th, tt, tr, dh, dt, b_size=60)
    ev = _ranks_gpu(model, kg)
    _assert_ranks_equal(ev, ref, kind)


@pytest.mark.parametrize("case", helpers.GOLDEN_CASES)
@pytest.mark.parametrize("tensor_core", [True, False])
def test_ranks_equal_the_unmodified_reference(case, tensor_core, cuda_device, monkeypatch):
    """The committed outputs of torchkge's own LinkPredictionEvaluator (tests/golden/*.npz): all
    four rank vectors and the metric getters, through both the tensor-core and the scalar scan."""
    import torchkge_b200.engine as engine_mod
    g = helpers.load_golden(case)
    model = helpers.model_from_golden(g).to(cuda_device)
    kg = tk.KnowledgeGraph(g["heads"], g["tails"], g["rels"], g["n_ent"], g["n_rel"],
                           dict_of_heads=g["dh"], dict_of_tails=g["dt"])
    monkeypatch.setattr(engine_mod, "_default_engine", engine_mod.CudaEngine(tensor_core=tensor_core))
    ev = tk.LinkPredictionEvaluator(model, kg)
    ev.evaluate(b_size=g["b_size"], verbose=False)
    raw = g["raw"]
    names = ["rank_true_heads", "rank_true_tails", "filt_rank_true_heads", "filt_rank_true_tails"]
    for nm in names:     # RESCAL included: its MKL query preparation is replayed in the reference's order
        assert torch.equal(getattr(ev, nm), torch.from_numpy(raw[nm])), nm
    got = [*ev.mean_rank(), *ev.hit_at_k(10), *ev.mrr()]
    for a, b in zip(got, raw["metrics"]):
        assert a == pytest.approx(float(b), rel=1e-6)


def test_empty_and_single_fact_graphs(cuda_device):
    model = helpers.make_model("distmult", 16, 40, 3, seed=0).to(cuda_device)
    e = torch.zeros(0, dtype=torch.long)
    kg = tk.KnowledgeGraph(e, e, e, 40, 3, dict_of_heads={}, dict_of_tails={})
    ev = tk.LinkPredictionEvaluator(model, kg)
    ev.evaluate(b_size=8, verbose=False)
    assert ev.rank_true_heads.shape == (0,) and ev.evaluated
    one = torch.tensor([0])          # entity 0 as head and tail, relation 0
    kg1 = tk.KnowledgeGraph(one, one, one, 40, 3)
    ev1 = tk.LinkPredictionEvaluator(model, kg1)
    ev1.evaluate(b_size=8, verbose=False)
    P = helpers.oracle_params("distmult", model)
    dh, dt = oracle.build_filter_dicts(one, one, one)
    ref = oracle.link_prediction("distmult", P, one, one, one, dh, dt, 8)
    _assert_ranks_equal(ev1, ref, "distmult")


@pytest.mark.parametrize("n_test", [120, 150])
def test_non_finite_embeddings_rank_like_the_reference(n_test, cuda_device):
    """NaN / inf rows: comparisons with NaN are false on both sides (operations.py:61).  150 test
    facts leave padding rows in the last 128-query tile beyond the 64-aligned true-score buffer:
    they must not reach the near-tie list."""
    n_ent, n_rel, d = 300, 4, 24
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=1500, n_test=n_test, seed=17)
    model = helpers.make_model("distmult", d, n_ent, n_rel, seed=17)
    with torch.no_grad():
        model.ent_emb.weight[5] = float("nan")
        model.ent_emb.weight[9, 3] = float("inf")
        model.ent_emb.weight[11, 0] = -float("inf")
    model = model.to(cuda_device)
    P = helpers.oracle_params("distmult", model)
    ref = oracle.link_prediction("distmult", P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, 64)
    _assert_ranks_equal(_ranks_gpu(model, kg), ref, "distmult")


def test_model_on_a_non_current_device(cuda_device):
    """A model on cuda:1 while cuda:0 is the current device: every launching entry point of the
    library switches to the device that owns its buffers (csrc/api.cu: DeviceScope), and kernel
    attributes are set per device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    other = torch.device("cuda", 1)
    assert torch.cuda.current_device() == 0
    n_ent, n_rel, d = 600, 5, 40
    kg, dh, dt = helpers.make_kg(n_ent, n_rel, n_facts=3000, n_test=150, seed=4)
    for kind in ("transe_l2", "transe_l1", "complex"):
        model = helpers.make_model(kind, d, n_ent, n_rel, seed=4).to(other)
        P = helpers.oracle_params(kind, model)
        ref = oracle.link_prediction(kind, P, kg.head_idx, kg.tail_idx, kg.relations, dh, dt, b_size=64)
        ev = _ranks_gpu(model, kg)
        _assert_ranks_equal(ev, ref, kind)
        assert torch.cuda.current_device() == 0
