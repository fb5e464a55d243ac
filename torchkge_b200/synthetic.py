"""Seeded synthetic workloads of the shapes BASELINE.json names:
"FB15k-237-shaped" skewed graphs (heads/tails ~ Zipf(0.8), relations ~ Zipf(1.0), ~21 facts
per entity, 20,466 test triples) and embedding tables with the reference constructors'
distribution (Xavier-uniform, entity rows L2-normalised where the reference does so).

Used by bench.py and the large-size GPU tests; nothing here is on the scoring path.
Tables are generated in fixed blocks of rows seeded by the block index, so any rank can
materialise exactly its shard and every world size sees the same table.
"""
import math

import torch

from . import _lib

BLOCK = 65536

WORKLOADS = {
    # name: (model code, TransE dissimilarity or None, dim, n_ent, n_rel, n_facts, n_test)
    "c1": dict(model="TransE", diss="L2", dim=50, n_ent=14541, n_rel=237, n_facts=310116, n_test=20466),
    "c2": dict(model="TransE", diss="L2", dim=200, n_ent=1000000, n_rel=1000, n_facts=21000000, n_test=20466),
    "c3": dict(model="ComplEx", diss=None, dim=400, n_ent=1000000, n_rel=1000, n_facts=21000000, n_test=20466),
    "c4": dict(model="RotatE", diss=None, dim=1000, n_ent=5000000, n_rel=1000, n_facts=21000000, n_test=20466),
    # C4's model at a tenth of its table: small enough for the CPU oracle to score in one piece (parity sample)
    "c4s": dict(model="RotatE", diss=None, dim=1000, n_ent=500000, n_rel=1000, n_facts=10500000, n_test=20466),
    "tiny": dict(model="TransE", diss="L2", dim=32, n_ent=2000, n_rel=20, n_facts=30000, n_test=512),
}

# C5 (BASELINE.json configs[4]): DistMult training step, Bernoulli corruption fused with scoring + margin loss
C5 = dict(model="DistMult", dim=200, n_ent=1000000, n_rel=1000, n_neg=256)

MODEL_CODE = {("TransE", "L1"): _lib.TRANSE_L1, ("TransE", "L2"): _lib.TRANSE_L2,
              ("DistMult", None): _lib.DISTMULT, ("RESCAL", None): _lib.RESCAL,
              ("ComplEx", None): _lib.COMPLEX, ("RotatE", None): _lib.ROTATE}
ORACLE_KIND = {_lib.TRANSE_L1: "transe_l1", _lib.TRANSE_L2: "transe_l2", _lib.DISTMULT: "distmult",
               _lib.RESCAL: "rescal", _lib.COMPLEX: "complex", _lib.ROTATE: "rotate"}


def _xavier_rows(lo, hi, cols, fan_rows, seed, device, normalise):
    """Rows [lo, hi) of a (fan_rows, cols) Xavier-uniform table, block-seeded."""
    a = math.sqrt(6.0 / (fan_rows + cols))
    out = torch.empty((hi - lo, cols), dtype=torch.float32, device=device)
    b0, b1 = lo // BLOCK, (hi + BLOCK - 1) // BLOCK
    for b in range(b0, b1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + b)
        rows = (torch.rand((BLOCK, cols), generator=g, device=device) * 2 - 1) * a
        s, e = max(lo, b * BLOCK), min(hi, (b + 1) * BLOCK)
        part = rows[s - b * BLOCK:e - b * BLOCK]
        if normalise:
            part = torch.nn.functional.normalize(part, p=2, dim=1)
        out[s - lo:e - lo] = part
    return out


def _xavier_small(rows, cols, seed, device, normalise):
    """Whole (rows, cols) Xavier-uniform table in one draw (relation tables)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    a = math.sqrt(6.0 / (rows + cols))
    out = (torch.rand((rows, cols), generator=g, device=device) * 2 - 1) * a
    return torch.nn.functional.normalize(out, p=2, dim=1) if normalise else out


def make_tables(code, dim, n_ent, n_rel, lo, hi, seed, device):
    """dict of fp32 tensors for entity rows [lo, hi) and all relations, keyed like the
    reference state_dict contract (ent0/ent1/rel0/rel1 in ModelSpec order)."""
    norm_ent = code in (_lib.TRANSE_L1, _lib.TRANSE_L2, _lib.DISTMULT, _lib.RESCAL)
    t = {"ent0": _xavier_rows(lo, hi, dim, n_ent, seed + 1, device, norm_ent), "ent1": None,
         "rel1": None}
    if code in (_lib.COMPLEX, _lib.ROTATE):
        t["ent1"] = _xavier_rows(lo, hi, dim, n_ent, seed + 2, device, False)
    if code == _lib.RESCAL:
        t["rel0"] = _xavier_small(n_rel, dim * dim, seed + 3, device, False)
    elif code == _lib.ROTATE:
        g = torch.Generator(device=device)
        g.manual_seed(seed + 3)
        ph = (torch.rand((n_rel, dim), generator=g, device=device) * 2 - 1) * math.pi
        t["rel0"], t["rel1"] = torch.cos(ph), torch.sin(ph)
    else:
        t["rel0"] = _xavier_small(n_rel, dim, seed + 3, device,
                                  code in (_lib.TRANSE_L1, _lib.TRANSE_L2))
        if code == _lib.COMPLEX:
            t["rel1"] = _xavier_small(n_rel, dim, seed + 4, device, False)
    return t


def rows_by_id(code, dim, n_ent, seed, ids, device):
    """{"ent0": (len(ids), dim), "ent1": ... or None}: the rows ``ids`` of the entity tables
    ``make_tables`` generates, without materialising the tables (each needed block of 65,536 rows is
    regenerated from its seed)."""
    norm_ent = code in (_lib.TRANSE_L1, _lib.TRANSE_L2, _lib.DISTMULT, _lib.RESCAL)
    ids = ids.long().cpu()
    out = {"ent0": torch.empty((ids.numel(), dim), dtype=torch.float32, device=device), "ent1": None}
    if code in (_lib.COMPLEX, _lib.ROTATE):
        out["ent1"] = torch.empty_like(out["ent0"])
    for b in torch.unique(ids // BLOCK).tolist():
        lo, hi = b * BLOCK, min(n_ent, (b + 1) * BLOCK)
        sel = torch.nonzero(ids // BLOCK == b).view(-1)
        local = (ids[sel] - lo).to(device)
        out["ent0"][sel.to(device)] = _xavier_rows(lo, hi, dim, n_ent, seed + 1, device, norm_ent)[local]
        if out["ent1"] is not None:
            out["ent1"][sel.to(device)] = _xavier_rows(lo, hi, dim, n_ent, seed + 2, device, False)[local]
    return out


def make_graph(n_ent, n_rel, n_facts, n_test, seed, device):
    """Skewed, deduplicated facts and a test split.  Returns dict of int64 device tensors:
    heads, tails, rels (all facts) and test_h, test_t, test_r (the first n_test of a seeded
    shuffle)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w_e = 1.0 / torch.arange(1, n_ent + 1, dtype=torch.float32, device=device) ** 0.8
    w_r = 1.0 / torch.arange(1, n_rel + 1, dtype=torch.float32, device=device)

    def draw(w, n):
        # inverse-CDF sampling (torch.multinomial caps the number of categories at 2^24)
        cdf = torch.cumsum(w.double(), 0)
        u = torch.rand(n, generator=g, device=device, dtype=torch.float64) * cdf[-1]
        return torch.searchsorted(cdf, u).clamp_(max=w.numel() - 1)

    perm_e = torch.randperm(n_ent, generator=g, device=device)  # popular ids spread over the table
    h = perm_e[draw(w_e, n_facts)]
    t = perm_e[draw(w_e, n_facts)]
    r = draw(w_r, n_facts)
    key = torch.unique((h * n_rel + r) * n_ent + t)
    key = key[torch.randperm(key.numel(), generator=g, device=device)]
    t = key % n_ent
    hr = key // n_ent
    h, r = hr // n_rel, hr % n_rel
    n_test = min(n_test, key.numel())
    return {"heads": h, "tails": t, "rels": r, "test_h": h[:n_test].contiguous(),
            "test_t": t[:n_test].contiguous(), "test_r": r[:n_test].contiguous()}


def _csr_for(keys_all, vals_all, keys_q, true_q):
    """CSR over queries: values of all facts sharing the query's key, minus the true value."""
    order = torch.argsort(keys_all)
    ks, vs = keys_all[order], vals_all[order]
    lo = torch.searchsorted(ks, keys_q)
    hi = torch.searchsorted(ks, keys_q, right=True)
    cnt = hi - lo
    q_of = torch.repeat_interleave(torch.arange(keys_q.numel(), device=keys_q.device), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    pos = torch.arange(q_of.numel(), device=keys_q.device) - start[q_of] + lo[q_of]
    ids = vs[pos]
    has_true = torch.zeros(keys_q.numel(), dtype=torch.bool, device=keys_q.device)
    has_true[q_of[ids == true_q[q_of]]] = True
    keep = (ids != true_q[q_of]) & has_true[q_of]   # get_true_targets: no true in set -> no filter
    ids, q_of = ids[keep], q_of[keep]
    offs = torch.zeros(keys_q.numel() + 1, dtype=torch.int64, device=keys_q.device)
    offs[1:] = torch.cumsum(torch.bincount(q_of, minlength=keys_q.numel()), 0)
    return offs, ids.contiguous(), q_of.to(torch.int32).contiguous()


def make_filters(graph, n_ent, n_rel):
    """Device CSRs of the filter sets of the test triples (true entity removed), computed from
    ALL facts of the graph: (csr_tail, csr_head) with csr = (offs int64 (n+1,), ids int64, rows int32)."""
    h, t, r = graph["heads"], graph["tails"], graph["rels"]
    th, tt, tr = graph["test_h"], graph["test_t"], graph["test_r"]
    # (offs, ids, row of every entry): the third array spares the filter kernel a bisection per entry
    csr_t = _csr_for(h * n_rel + r, t, th * n_rel + tr, tt)
    csr_h = _csr_for(t * n_rel + r, h, tt * n_rel + tr, th)
    return csr_t, csr_h


def filters_as_dicts(graph, n_ent, n_rel, limit=None):
    """(dict_of_heads, dict_of_tails) holding the FULL sets (true entity included) of exactly
    the keys the first `limit` test triples touch -- what the reference-side evaluator needs."""
    from collections import defaultdict
    h, t, r = graph["heads"], graph["tails"], graph["rels"]
    th, tt, tr = graph["test_h"], graph["test_t"], graph["test_r"]
    if limit is not None:
        th, tt, tr = th[:limit], tt[:limit], tr[:limit]
    out = []
    for keys_all, vals_all, keys_q, k1, k2 in (
            (t * n_rel + r, h, tt * n_rel + tr, tt, tr),     # dict_of_heads keyed (t, r)
            (h * n_rel + r, t, th * n_rel + tr, th, tr)):    # dict_of_tails keyed (h, r)
        order = torch.argsort(keys_all)
        ks, vs = keys_all[order], vals_all[order]
        lo = torch.searchsorted(ks, keys_q).tolist()
        hi = torch.searchsorted(ks, keys_q, right=True).tolist()
        vs_c = vs.cpu()
        d = defaultdict(set)
        for a, b, l, u in zip(k1.tolist(), k2.tolist(), lo, hi):
            if (a, b) not in d:
                d[(a, b)] = set(vs_c[l:u].tolist())
        out.append(d)
    return out[0], out[1]
