"""Minimal knowledge-graph container for the hot path.

The reference's ``KnowledgeGraph`` (torchkge/data_structures.py:17-415) stays usable as-is:
the evaluator and the sampler only read the attributes listed in SURVEY.md section 8b
(``n_ent, n_rel, n_facts, head_idx, tail_idx, relations, dict_of_heads, dict_of_tails``).
This class provides exactly those for scripts / tests / benchmarks that do not have torchkge
installed; building graphs from data frames, splitting, label dictionaries etc. are out of
scope here.
"""
from collections import defaultdict

import torch


class KnowledgeGraph:
    """Index-tensor view of a set of facts.

    Parameters
    ----------
    heads, tails, relations: torch.LongTensor (n_facts,), CPU
    n_ent, n_rel: int
    dict_of_heads: mapping (t, r) -> set of heads, optional
    dict_of_tails: mapping (h, r) -> set of tails, optional
        Filter dictionaries.  When omitted they are built from this graph's own facts
        (the reference does the same in ``evaluate_dicts``, data_structures.py:386-397).
        Pass the full-graph dictionaries when evaluating a split.
    """

    def __init__(self, heads, tails, relations, n_ent, n_rel, dict_of_heads=None,
                 dict_of_tails=None):
        if not (heads.shape == tails.shape == relations.shape and heads.dim() == 1):
            raise ValueError("heads, tails, relations must be 1-D tensors of equal length")
        self.head_idx = heads.long().cpu()
        self.tail_idx = tails.long().cpu()
        self.relations = relations.long().cpu()
        self.n_ent, self.n_rel = int(n_ent), int(n_rel)
        self.n_facts = int(heads.shape[0])
        if dict_of_heads is None or dict_of_tails is None:
            dict_of_heads, dict_of_tails = build_filter_dicts(self.head_idx, self.tail_idx,
                                                              self.relations)
        self.dict_of_heads = dict_of_heads
        self.dict_of_tails = dict_of_tails

    def __len__(self):
        return self.n_facts


def build_filter_dicts(heads, tails, relations):
    """(dict_of_heads keyed (t, r), dict_of_tails keyed (h, r)) as defaultdict(set)."""
    order = torch.arange(heads.shape[0])
    dh, dt = defaultdict(set), defaultdict(set)
    hl, tl, rl = heads.tolist(), tails.tolist(), relations.tolist()
    for i in order.tolist():
        dh[(tl[i], rl[i])].add(hl[i])
        dt[(hl[i], rl[i])].add(tl[i])
    return dh, dt


def filter_csr(dictionary, key1, key2, true_idx):
    """CSR (offs, ids) of the entities ``filter_scores`` would mask for each row.

    Row i lists dictionary[(key1[i], key2[i])] minus true_idx[i]; following
    get_true_targets (torchkge/utils/modeling.py:53-88) the row is EMPTY when the key is
    unknown or when the true entity is not in the set (the reference's ``remove`` raises
    KeyError there and the row is left unfiltered).  The per-row work is two dict/set probes;
    set contents are flattened at C speed and the true entities are masked out vectorised.
    """
    import itertools

    import numpy as np
    n = int(key1.shape[0])
    get = dictionary.get
    picked, lens = [], np.zeros(n, dtype=np.int64)
    for i, (a, b, c) in enumerate(zip(key1.tolist(), key2.tolist(), true_idx.tolist())):
        s = get((a, b))
        if s is not None and c in s:
            picked.append(s)
            lens[i] = len(s)
    total = int(lens.sum())
    flat = np.fromiter(itertools.chain.from_iterable(picked), dtype=np.int64, count=total)
    true_rep = np.repeat(true_idx.numpy().astype(np.int64, copy=False), lens)
    keep = flat != true_rep
    # list-valued dictionaries (the reference's own test fixture) may repeat the true entity;
    # sets cannot: either way every occurrence is dropped, as ``remove`` + masking would not
    # -- only sets are produced by KnowledgeGraph, so this is moot in practice.
    row_of = np.repeat(np.arange(n, dtype=np.int64), lens)[keep]
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(row_of, minlength=n), out=offs[1:])
    return torch.from_numpy(offs), torch.from_numpy(flat[keep])
