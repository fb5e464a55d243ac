"""CPU oracle for the link-prediction / scoring hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs may import this module; the product (``torchkge_b200``) never does and has no CPU
fallback.

What this is: a restatement, in plain PyTorch CPU tensor ops, of the algorithm torchkge
v0.17.7 (commit 3adb934) runs for
  * all-entity scoring      models/interfaces.py:240-260, models/bilinear.py:98-121,
                            224-245, 501-528, 682-711 (+ utils/dissimilarities.py:11-25)
  * filtering               utils/modeling.py:53-102
  * ranking                 utils/operations.py:37-61
  * the evaluator loop      evaluation.py:263-308
  * relation prediction     evaluation.py:64-112 with the relation case of
                            inference_scoring_function (interfaces.py:261-272,
                            bilinear.py:115-121, 241-245, 524-528)
  * per-triple scoring      models/translation.py:69-81, models/bilinear.py:60-71,
                            188-199, 460-473, models/interfaces.py:39-82
  * Bernoulli corruption    sampling.py:259-327, utils/operations.py:116-179
  * margin loss             utils/losses.py:12-44
The reference's arithmetic lives in ATen (third-party, torch>=1.2 per its setup.py:11; here
torch 2.11.0), so the oracle issues the same tensor ops in the same order and therefore
produces the same bits as the reference on the same machine.

Pinning: ``tests/golden/*.npz`` hold inputs and outputs of the UNMODIFIED reference
(generated in the authoring container by ``tests/golden/make_golden.py`` with
``PYTHONPATH=/root/reference``); ``tests/test_oracle_golden.py`` requires this module to
reproduce them exactly (ranks, scores, Bernoulli probabilities, loss).  The reference's own
known-answer tests for this path (tests/test_utils.py:97-107, 119-127, 156-163, 77-95) are
replayed there too.  RotatE does not exist in the reference: ``rotate_*`` below is a
restatement of Sun et al. 2019 in torchkge's ComplEx style and is **parity unpinned**.
"""
from collections import defaultdict

import torch

KINDS = ("transe_l1", "transe_l2", "distmult", "rescal", "complex", "rotate", "toruse_l1", "toruse_l2",
         "analogy")


# --------------------------------------------------------------------------- dissimilarities
def l1_diss(a, b):
    """utils/dissimilarities.py:11-16"""
    return (a - b).norm(p=1, dim=-1)


def l2_diss(a, b):
    """utils/dissimilarities.py:19-25 -- note: 2-norm first, THEN squared."""
    return (a - b).norm(p=2, dim=-1) ** 2


def l1_torus_diss(a, b):
    """utils/dissimilarities.py:28-34 (TorusE)"""
    return 2 * torch.min(torch.abs(a - b), 1 - torch.abs(a - b)).sum(dim=-1)


def l2_torus_diss(a, b):
    """utils/dissimilarities.py:37-43 (TorusE)"""
    return 4 * torch.min((a - b) ** 2, 1 - (a - b) ** 2).sum(dim=-1)


# --------------------------------------------------------------------------- all-entity scores
def _rows(P, names, idx):
    return [P[n][idx] for n in names]


def scores_all(kind, P, h_idx, t_idx, r_idx, side):
    """(b, n_ent) scores of every entity as tail (side='tail') or head (side='head').

    Follows inference_prepare_candidates + inference_scoring_function of the model; the
    candidates tensor is the weight matrix broadcast along the batch axis.
    """
    b = h_idx.shape[0]
    tail = side == "tail"
    if kind in ("transe_l1", "transe_l2", "toruse_l1", "toruse_l2"):
        # TorusE (translation.py:655-767) is TransE's inference_scoring_function on tables that
        # hold fractional parts (normalize_parameters), with the torus dissimilarities
        diss = {"transe_l1": l1_diss, "transe_l2": l2_diss, "toruse_l1": l1_torus_diss,
                "toruse_l2": l2_torus_diss}[kind]
        E, R = P["ent"], P["rel"]
        d = E.shape[1]
        cand = E.view(1, -1, d).expand(b, -1, d)
        h, t, r = E[h_idx], E[t_idx], R[r_idx]
        if tail:  # interfaces.py:249-254
            hr = (h + r).view(b, 1, d)
            return -diss(hr, cand)
        # interfaces.py:256-260
        return -diss(cand + r.view(b, 1, d), t.view(b, 1, d))
    if kind == "distmult":
        E, R = P["ent"], P["rel"]
        d = E.shape[1]
        cand = E.view(1, -1, d).expand(b, -1, d)
        h, t, r = E[h_idx], E[t_idx], R[r_idx]
        if tail:  # bilinear.py:231-235
            return ((h * r).view(b, 1, d) * cand).sum(dim=2)
        return (cand * (r * t).view(b, 1, d)).sum(dim=2)  # bilinear.py:236-240
    if kind == "rescal":
        E, M = P["ent"], P["rel_mat"]
        d = E.shape[1]
        cand = E.view(1, -1, d).expand(b, -1, d)
        h, t, m = E[h_idx], E[t_idx], M[r_idx].view(-1, d, d)
        if tail:  # bilinear.py:110-114
            hr = torch.matmul(h.view(b, 1, d), m).view(b, 1, d)
            return (hr * cand).sum(dim=2)
        tr = torch.matmul(m, t.view(b, d, 1)).view(b, 1, d)  # bilinear.py:105-109
        return (cand * tr).sum(dim=2)
    if kind == "complex":
        d = P["re_ent"].shape[1]
        re_c = P["re_ent"].view(1, -1, d).expand(b, -1, d)
        im_c = P["im_ent"].view(1, -1, d).expand(b, -1, d)
        re_h, im_h = _rows(P, ("re_ent", "im_ent"), h_idx)
        re_t, im_t = _rows(P, ("re_ent", "im_ent"), t_idx)
        re_r, im_r = _rows(P, ("re_rel", "im_rel"), r_idx)
        if tail:  # bilinear.py:511-515
            return ((re_h * re_r - im_h * im_r).view(b, 1, d) * re_c
                    + (re_h * im_r + im_h * re_r).view(b, 1, d) * im_c).sum(dim=2)
        # bilinear.py:517-522
        return (re_c * (re_r * re_t + im_r * im_t).view(b, 1, d)
                + im_c * (re_r * im_t - im_r * re_t).view(b, 1, d)).sum(dim=2)
    if kind == "analogy":
        names = ("sc_ent", "re_ent", "im_ent")
        ds, dc = P["sc_ent"].shape[1], P["re_ent"].shape[1]
        sc_c = P["sc_ent"].view(1, -1, ds).expand(b, -1, ds)
        re_c = P["re_ent"].view(1, -1, dc).expand(b, -1, dc)
        im_c = P["im_ent"].view(1, -1, dc).expand(b, -1, dc)
        sc_h, re_h, im_h = _rows(P, names, h_idx)
        sc_t, re_t, im_t = _rows(P, names, t_idx)
        sc_r, re_r, im_r = _rows(P, ("sc_rel", "re_rel", "im_rel"), r_idx)
        if tail:  # bilinear.py:692-698
            return ((sc_h * sc_r).view(b, 1, ds) * sc_c
                    + (re_h * re_r - im_h * im_r).view(b, 1, dc) * re_c
                    + (re_h * im_r + im_h * re_r).view(b, 1, dc) * im_c).sum(dim=2)
        # bilinear.py:699-705
        return (sc_c * (sc_r * sc_t).view(b, 1, ds)
                + re_c * (re_r * re_t + im_r * im_t).view(b, 1, dc)
                + im_c * (re_r * im_t - im_r * re_t).view(b, 1, dc)).sum(dim=2)
    if kind == "rotate":
        return rotate_scores_all(P, h_idx, t_idx, r_idx, side)
    raise ValueError(kind)


def rotate_rel_planes(P):
    """(cos, sin) of the relation phases.  Kept separate so GPU tests can feed the kernel the
    very same fp32 tables (libm cos/sin differ between CPU and GPU in the last ulp)."""
    if "re_rel" in P and "im_rel" in P:
        return P["re_rel"], P["im_rel"]
    return torch.cos(P["rel_phase"]), torch.sin(P["rel_phase"])


def rotate_scores_all(P, h_idx, t_idx, r_idx, side):
    """RotatE (Sun et al. 2019, eq. 5): score = -sum_k |h_k * r_k - t_k| with |r_k| = 1.

    NOT IN THE REFERENCE -- parity unpinned.  Written in the style of ComplExModel
    (bilinear.py:501-556): separate re/im tables, a (b, 1, d) query against the broadcast
    candidate tables.  Tail: q = h o r, distance to candidates.  Head: as in the authors'
    head-batch mode, q = t o conj(r), distance to candidates.
    """
    d = P["re_ent"].shape[1]
    b = h_idx.shape[0]
    re_c = P["re_ent"].view(1, -1, d).expand(b, -1, d)
    im_c = P["im_ent"].view(1, -1, d).expand(b, -1, d)
    re_rel, im_rel = rotate_rel_planes(P)
    re_r, im_r = re_rel[r_idx], im_rel[r_idx]
    if side == "tail":
        re_e, im_e = _rows(P, ("re_ent", "im_ent"), h_idx)
        re_q = re_e * re_r - im_e * im_r
        im_q = re_e * im_r + im_e * re_r
    else:
        re_e, im_e = _rows(P, ("re_ent", "im_ent"), t_idx)
        re_q = re_r * re_e + im_r * im_e
        im_q = re_r * im_e - im_r * re_e
    dre = re_q.view(b, 1, d) - re_c
    dim_ = im_q.view(b, 1, d) - im_c
    return -complex_modulus(dre, dim_).sum(dim=2)


def complex_modulus(re, im, authors_form=False):
    """|re + i im| element-wise, as the authors' code computes it: ``stack([re, im]).norm(dim=0)``,
    i.e. sqrt_rn(fl(fl(re^2) + fl(im^2))) with a CORRECTLY ROUNDED square root (ATen's element-wise
    fp32 ``torch.sqrt`` is not: 0.35 % of values are off by one ulp on CPU).  The stacked norm is a
    reduction over an axis of length 2 and runs at a few tens of M elements/s; the default form
    takes the square root in float64 and rounds once more to float32 -- the same bits (a float64 sqrt
    of a float32 value rounds to the correctly rounded float32 sqrt; tests/test_oracle_golden.py checks the
    two forms against each other) at memory speed, which is what makes oracle samples at the C4
    table size affordable."""
    if authors_form:
        return torch.stack([re, im], dim=0).norm(dim=0)
    return torch.sqrt((re * re + im * im).double()).float()


# --------------------------------------------------------------------------- filter + rank
def filter_targets(dictionary, k1, k2, true_e):
    """utils/modeling.py:53-88 for one row.  Returns the list of entities to mask or None.

    Quirks kept: the lookup goes through the (default)dict as-is; if the true entity is not
    in the set the `remove` raises KeyError and the row is left unfiltered.
    """
    try:
        s = dictionary[k1, k2].copy()
        s.remove(true_e)
    except KeyError:
        return None
    if len(s) == 0:
        return None
    return list(s)


def filtered_scores(scores, dictionary, key1, key2, true_idx):
    """utils/modeling.py:91-102"""
    out = scores.clone()
    for i in range(scores.shape[0]):
        tg = filter_targets(dictionary, key1[i].item(), key2[i].item(), true_idx[i].item())
        if tg is not None:
            out[i][torch.tensor(tg).long()] = -float("inf")
    return out


def rank_of_true(scores, true_idx):
    """utils/operations.py:56-61 (high scores are good; ties count against the true one)."""
    true_scores = scores.gather(1, true_idx.long().view(-1, 1))
    return (scores >= true_scores).sum(dim=1)


def link_prediction(kind, P, heads, tails, rels, dict_of_heads, dict_of_tails, b_size):
    """evaluation.py:263-308: returns (rank_true_heads, rank_true_tails,
    filt_rank_true_heads, filt_rank_true_tails), int64 (n_facts,)."""
    n = heads.shape[0]
    out = [torch.empty(n, dtype=torch.long) for _ in range(4)]
    for lo in range(0, n, b_size):
        hi = min(n, lo + b_size)
        h, t, r = heads[lo:hi], tails[lo:hi], rels[lo:hi]
        s = scores_all(kind, P, h, t, r, "tail")
        fs = filtered_scores(s, dict_of_tails, h, r, t)
        out[1][lo:hi] = rank_of_true(s, t)
        out[3][lo:hi] = rank_of_true(fs, t)
        s = scores_all(kind, P, h, t, r, "head")
        fs = filtered_scores(s, dict_of_heads, t, r, h)
        out[0][lo:hi] = rank_of_true(s, h)
        out[2][lo:hi] = rank_of_true(fs, h)
    return tuple(out)


# --------------------------------------------------------------------------- relation prediction
def relation_scores_all(kind, P, h_idx, t_idx):
    """(b, n_rel) scores of every relation for the pairs (h, t): inference_prepare_candidates(...,
    entities=False) + the relation case of inference_scoring_function -- interfaces.py:261-272
    (TransE), bilinear.py:115-121 (RESCAL), 241-245 (DistMult), 524-528 (ComplEx)."""
    b = h_idx.shape[0]
    if kind in ("transe_l1", "transe_l2"):
        diss = l1_diss if kind == "transe_l1" else l2_diss
        d = P["ent"].shape[1]
        h, t = P["ent"][h_idx], P["ent"][t_idx]
        cands = P["rel"].view(1, -1, d).expand(b, -1, -1)
        return -diss(h.view(b, -1, d) + cands, t.view(b, -1, d))
    if kind == "distmult":
        d = P["ent"].shape[1]
        h, t = P["ent"][h_idx], P["ent"][t_idx]
        cands = P["rel"].view(1, -1, d).expand(b, -1, -1)
        hr = h.view(b, 1, d) * cands
        return (hr * t.view(b, 1, d)).sum(dim=2)
    if kind == "rescal":
        d = P["ent"].shape[1]
        n_rel = P["rel_mat"].shape[0]
        h, t = P["ent"][h_idx], P["ent"][t_idx]
        cands = P["rel_mat"].view(1, n_rel, d, d).expand(b, n_rel, d, d)
        hr = torch.matmul(h.view(b, 1, 1, d), cands).view(b, n_rel, d)
        return (hr * t.view(b, 1, d)).sum(dim=2)
    if kind == "complex":
        d = P["re_ent"].shape[1]
        re_h, im_h = _rows(P, ("re_ent", "im_ent"), h_idx)
        re_t, im_t = _rows(P, ("re_ent", "im_ent"), t_idx)
        re_r = P["re_rel"].view(1, -1, d).expand(b, -1, -1)
        im_r = P["im_rel"].view(1, -1, d).expand(b, -1, -1)
        return ((re_h * re_t + im_h * im_t).view(b, 1, d) * re_r
                + (re_h * im_t - im_h * re_t).view(b, 1, d) * im_r).sum(dim=2)
    if kind == "analogy":   # bilinear.py:706-711
        names = ("sc_ent", "re_ent", "im_ent")
        ds, dc = P["sc_ent"].shape[1], P["re_ent"].shape[1]
        sc_h, re_h, im_h = _rows(P, names, h_idx)
        sc_t, re_t, im_t = _rows(P, names, t_idx)
        sc_r = P["sc_rel"].view(1, -1, ds).expand(b, -1, -1)
        re_r = P["re_rel"].view(1, -1, dc).expand(b, -1, -1)
        im_r = P["im_rel"].view(1, -1, dc).expand(b, -1, -1)
        return (sc_r * (sc_h * sc_t).view(b, 1, ds)
                + re_r * (re_h * re_t + im_h * im_t).view(b, 1, dc)
                + im_r * (re_h * im_t - im_h * re_t).view(b, 1, dc)).sum(dim=2)
    raise ValueError(kind)


def relation_prediction(kind, P, heads, tails, rels, dict_of_rels, b_size, directed=True):
    """RelationPredictionEvaluator.evaluate, evaluation.py:64-112: (rank_true_rels,
    filt_rank_true_rels).  Undirected: the (t, ?, h) scores are concatenated after the (h, ?, t)
    ones and the true relation keeps its index in the first half."""
    n = heads.shape[0]
    out = [torch.empty(n, dtype=torch.long) for _ in range(2)]
    for lo in range(0, n, b_size):
        hi = min(n, lo + b_size)
        h, t, r = heads[lo:hi], tails[lo:hi], rels[lo:hi]
        s = relation_scores_all(kind, P, h, t)
        fs = filtered_scores(s, dict_of_rels, h, t, r)
        if not directed:
            s2 = relation_scores_all(kind, P, t, h)
            fs2 = filtered_scores(s2, dict_of_rels, h, t, r)
            s, fs = torch.cat((s, s2), dim=1), torch.cat((fs, fs2), dim=1)
        out[0][lo:hi] = rank_of_true(s, r)
        out[1][lo:hi] = rank_of_true(fs, r)
    return tuple(out)


def build_rel_dict(heads, tails, rels):
    """data_structures.py:386-397: dict_of_rels keyed (h, t)."""
    dr = defaultdict(set)
    for h, t, r in zip(heads.tolist(), tails.tolist(), rels.tolist()):
        dr[(h, t)].add(r)
    return dr


def lp_metrics(rh, rt, frh, frt, k=10):
    """evaluation.py:310-397: (mean_rank, hit@k, mrr), each a (raw, filtered) pair."""
    mr = ((rh.float().mean() + rt.float().mean()).item() / 2,
          (frh.float().mean() + frt.float().mean()).item() / 2)
    hit = (((rh <= k).float().mean().item() + (rt <= k).float().mean().item()) / 2,
           ((frh <= k).float().mean().item() + (frt <= k).float().mean().item()) / 2)
    mrr = (((rh.float() ** (-1)).mean() + (rt.float() ** (-1)).mean()).item() / 2,
           ((frh.float() ** (-1)).mean() + (frt.float() ** (-1)).mean()).item() / 2)
    return mr, hit, mrr


# --------------------------------------------------------------------------- per-triple scores
def score_triples(kind, P, h_idx, t_idx, r_idx):
    """Model.scoring_function: translation.py:69-81, bilinear.py:60-71, 188-199, 460-473.
    TransE / DistMult / RESCAL L2-normalise the gathered entity rows first."""
    nrm = torch.nn.functional.normalize
    if kind in ("transe_l1", "transe_l2"):
        diss = l1_diss if kind == "transe_l1" else l2_diss
        h = nrm(P["ent"][h_idx], p=2, dim=1)
        t = nrm(P["ent"][t_idx], p=2, dim=1)
        return -diss(h + P["rel"][r_idx], t)
    if kind == "distmult":
        h = nrm(P["ent"][h_idx], p=2, dim=1)
        t = nrm(P["ent"][t_idx], p=2, dim=1)
        return (h * P["rel"][r_idx] * t).sum(dim=1)
    if kind == "rescal":
        d = P["ent"].shape[1]
        h = nrm(P["ent"][h_idx], p=2, dim=1)
        t = nrm(P["ent"][t_idx], p=2, dim=1)
        m = P["rel_mat"][r_idx].view(-1, d, d)
        hr = torch.matmul(h.view(-1, 1, d), m)
        return (hr.view(-1, d) * t).sum(dim=1)
    if kind == "complex":
        re_h, im_h = _rows(P, ("re_ent", "im_ent"), h_idx)
        re_t, im_t = _rows(P, ("re_ent", "im_ent"), t_idx)
        re_r, im_r = _rows(P, ("re_rel", "im_rel"), r_idx)
        return (re_h * (re_r * re_t + im_r * im_t) + im_h * (re_r * im_t - im_r * re_t)).sum(dim=1)
    if kind == "analogy":   # bilinear.py:634-650
        sc_h, re_h, im_h = _rows(P, ("sc_ent", "re_ent", "im_ent"), h_idx)
        sc_t, re_t, im_t = _rows(P, ("sc_ent", "re_ent", "im_ent"), t_idx)
        sc_r, re_r, im_r = _rows(P, ("sc_rel", "re_rel", "im_rel"), r_idx)
        return ((sc_h * sc_r * sc_t).sum(dim=1) +
                (re_h * (re_r * re_t + im_r * im_t) + im_h * (re_r * im_t - im_r * re_t)).sum(dim=1))
    if kind == "rotate":
        re_h, im_h = _rows(P, ("re_ent", "im_ent"), h_idx)
        re_t, im_t = _rows(P, ("re_ent", "im_ent"), t_idx)
        re_rel, im_rel = rotate_rel_planes(P)
        re_r, im_r = re_rel[r_idx], im_rel[r_idx]
        dre = (re_h * re_r - im_h * im_r) - re_t
        dim_ = (re_h * im_r + im_h * re_r) - im_t
        return -complex_modulus(dre, dim_).sum(dim=1)
    raise ValueError(kind)


def forward_pos_neg(kind, P, h, t, r, nh, nt):
    """Model.forward, models/interfaces.py:39-82 (negative_relations=None)."""
    pos = score_triples(kind, P, h, t, r)
    if nh.shape[0] > r.shape[0]:
        n_neg = int(nh.shape[0] / r.shape[0])
        pos = pos.repeat(n_neg)
        neg = score_triples(kind, P, nh, nt, r.repeat(n_neg))
    else:
        neg = score_triples(kind, P, nh, nt, r)
    return pos, neg


def margin_loss(pos, neg, margin):
    """utils/losses.py:19-44: MarginRankingLoss(margin, reduction='sum') with target +1."""
    crit = torch.nn.MarginRankingLoss(margin=margin, reduction="sum")
    return crit(pos, neg, target=torch.ones_like(pos))


# --------------------------------------------------------------------------- Bernoulli sampler
def bernoulli_probs(heads, tails, rels, n_rel):
    """p_r = tph / (tph + hpt)  (utils/operations.py:116-179, sampling.py:263-276).

    tph = mean over (head, rel) groups of the group size, per relation; hpt likewise over
    (rel, tail) groups; relations absent from the graph get 0.5.  The reference does this
    with pandas groupby().count().groupby().mean() in float64; so do we, with torch.unique.
    """
    out = torch.full((n_rel,), 0.5, dtype=torch.float64)
    hr = torch.stack([heads, rels], 1)
    tr = torch.stack([tails, rels], 1)
    uhr, chr_ = torch.unique(hr, dim=0, return_counts=True)
    utr, ctr = torch.unique(tr, dim=0, return_counts=True)
    for rel in torch.unique(rels).tolist():
        tph = chr_[uhr[:, 1] == rel].double().mean()
        hpt = ctr[utr[:, 1] == rel].double().mean()
        out[rel] = tph / (tph + hpt)
    return out.float()


def corrupt_batch(heads, tails, rels, bern_probs, n_ent, n_neg):
    """BernoulliNegativeSampler.corrupt_batch, sampling.py:278-327 (torch global RNG)."""
    b = heads.shape[0]
    nh = heads.repeat(n_neg)
    nt = tails.repeat(n_neg)
    mask = torch.bernoulli(bern_probs[rels].repeat(n_neg)).double()
    n_h = int(mask.sum().item())
    nh[mask == 1] = torch.randint(1, n_ent, (n_h,))
    nt[mask == 0] = torch.randint(1, n_ent, (b * n_neg - n_h,))
    return nh.long(), nt.long()


# --------------------------------------------------------------------------- graph helpers
def build_filter_dicts(heads, tails, rels):
    """data_structures.py:386-397 (dict_of_heads keyed (t, r); dict_of_tails keyed (h, r))."""
    dh, dt = defaultdict(set), defaultdict(set)
    for h, t, r in zip(heads.tolist(), tails.tolist(), rels.tolist()):
        dh[(t, r)].add(h)
        dt[(h, r)].add(t)
    return dh, dt


def init_params(kind, n_ent, n_rel, d, generator=None):
    """Parameter tensors with the reference constructors' distribution: Xavier-uniform
    (utils/modeling.py:21-28), entity rows L2-normalised for TransE / DistMult / RESCAL,
    TransE relations normalised too (translation.py:58-67; bilinear.py:51-58, 180-186);
    ComplEx raw (bilinear.py:455-458).  RotatE: ComplEx-like entities, phases U(-pi, pi)."""
    def xavier(rows, cols):
        a = (6.0 / (rows + cols)) ** 0.5
        return (torch.rand(rows, cols, generator=generator) * 2 - 1) * a
    nrm = torch.nn.functional.normalize
    if kind in ("transe_l1", "transe_l2"):
        return {"ent": nrm(xavier(n_ent, d), p=2, dim=1), "rel": nrm(xavier(n_rel, d), p=2, dim=1)}
    if kind == "distmult":
        return {"ent": nrm(xavier(n_ent, d), p=2, dim=1), "rel": xavier(n_rel, d)}
    if kind == "rescal":
        return {"ent": nrm(xavier(n_ent, d), p=2, dim=1), "rel_mat": xavier(n_rel, d * d)}
    if kind == "complex":
        return {"re_ent": xavier(n_ent, d), "im_ent": xavier(n_ent, d),
                "re_rel": xavier(n_rel, d), "im_rel": xavier(n_rel, d)}
    if kind == "analogy":   # bilinear.py:620-631: d = emb_dim, split in halves (scalar_share 0.5); raw Xavier
        ds = d // 2
        dc = d - ds
        return {"sc_ent": xavier(n_ent, ds), "re_ent": xavier(n_ent, dc), "im_ent": xavier(n_ent, dc),
                "sc_rel": xavier(n_rel, ds), "re_rel": xavier(n_rel, dc), "im_rel": xavier(n_rel, dc)}
    if kind == "rotate":
        ph = (torch.rand(n_rel, d, generator=generator) * 2 - 1) * 3.141592653589793
        return {"re_ent": xavier(n_ent, d), "im_ent": xavier(n_ent, d), "rel_phase": ph}
    raise ValueError(kind)
