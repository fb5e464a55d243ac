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


class FilterIndex:
    """Sorted-array form of the filter dictionaries.

    The reference keeps ``dict_of_tails[(h, r)] -> set`` / ``dict_of_heads[(t, r)] -> set``
    built by a Python loop over all facts (data_structures.py:386-397, ~26 us per fact) and
    walks them row by row at evaluation time (utils/modeling.py:53-102).  The same information
    as two sorted key arrays lets the per-row filter lists of a whole test set be produced by
    two ``searchsorted`` calls and a gather -- vectorised, on the host, while the GPU is busy
    with the dense scans.
    """

    def __init__(self, heads, tails, relations, n_ent, n_rel):
        # built where the facts live: on a GPU the two sorts of tens of millions of keys take tens of
        # milliseconds instead of seconds on the host
        heads, tails, relations = (x.long() for x in (heads, tails, relations))
        self.n_ent, self.n_rel = int(n_ent), int(n_rel)
        if self.n_ent * self.n_rel * self.n_ent >= 2 ** 63:
            raise ValueError("FilterIndex packs (entity, relation, entity) into one int64 key: "
                             "n_ent^2 * n_rel = %d * %d * %d does not fit" % (self.n_ent, self.n_rel, self.n_ent))
        for name, x, hi in (("heads", heads, self.n_ent), ("tails", tails, self.n_ent),
                            ("relations", relations, self.n_rel)):
            if x.numel() and (int(x.min()) < 0 or int(x.max()) >= hi):
                raise ValueError("FilterIndex: %s outside [0, %d)" % (name, hi))
        # one sorted, deduplicated int64 array per side: (key1 * n_rel + rel) * n_ent + value
        built = {
            "tail": torch.unique((heads * self.n_rel + relations) * self.n_ent + tails),
            "head": torch.unique((tails * self.n_rel + relations) * self.n_ent + heads),
        }
        self._on_device = {heads.device: built}

    @property
    def kv(self):
        """the two key arrays on the host (copied there on first use when the index was built on a GPU)"""
        return self.on("cpu")

    def on(self, device):
        """The two key arrays on ``device`` (moved there once and kept, like model weights)."""
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device not in self._on_device:
            src = next(iter(self._on_device.values()))
            self._on_device[device] = {k: v.to(device) for k, v in src.items()}
        return self._on_device[device]

    def csr(self, which, key1, key2, true_idx):
        """CSR (offs, ids) with exactly the semantics of ``filter_csr`` on the dictionaries:
        row i = {v : (key1[i], key2[i], v) is a fact} minus true_idx[i], empty when the true
        entity is not in that set (get_true_targets' KeyError quirk).  Runs on the device of
        ``key1`` (CPU or CUDA); everything is searchsorted / gather, no Python loop."""
        dev = key1.device
        kv = self.on(dev)[which]
        kq = (key1.long() * self.n_rel + key2.long()) * self.n_ent
        true_idx = true_idx.long()
        lo = torch.searchsorted(kv, kq)
        hi = torch.searchsorted(kv, kq + self.n_ent)
        kvq = kq + true_idx
        pos_true = torch.searchsorted(kv, kvq)
        has_true = (pos_true < kv.numel()) & (kv[pos_true.clamp(max=max(kv.numel() - 1, 0))] == kvq)
        zero = torch.zeros_like(lo)
        cnt = torch.where(has_true, hi - lo - 1, zero)
        n = kq.numel()
        offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=offs[1:])
        full = torch.where(has_true, hi - lo, zero)
        row = torch.repeat_interleave(torch.arange(n, device=dev), full)
        start = torch.cumsum(full, 0) - full
        pos = torch.arange(row.numel(), device=dev) - start[row] + lo[row]
        ids = kv[pos] - kq[row]
        keep = ids != true_idx[row]
        # third array: the row every entry belongs to (the filter kernel then needs no bisection)
        return offs, ids[keep].contiguous(), row[keep].to(torch.int32).contiguous()

    def as_dicts(self):
        """(dict_of_heads, dict_of_tails) for code that wants the reference's containers."""
        out = {}
        for which in ("head", "tail"):
            d = defaultdict(set)
            kv = self.kv[which]
            keys = torch.div(kv, self.n_ent, rounding_mode="floor")
            vals = (kv - keys * self.n_ent).tolist()
            for k, v in zip(keys.tolist(), vals):
                d[(k // self.n_rel, k % self.n_rel)].add(v)
            out[which] = d
        return out["head"], out["tail"]


class KnowledgeGraph:
    """Index-tensor view of a set of facts.

    Parameters
    ----------
    heads, tails, relations: torch.LongTensor (n_facts,), CPU
    n_ent, n_rel: int
    dict_of_heads: mapping (t, r) -> set of heads, optional
    dict_of_tails: mapping (h, r) -> set of tails, optional
        Filter dictionaries as in the reference.  Pass the full-graph dictionaries when
        evaluating a split.
    filter_facts: (heads, tails, relations) of ALL known facts, optional
        Alternative to the dictionaries: the filter sets are then kept as a ``FilterIndex``
        (sorted arrays; seconds to build for tens of millions of facts) and the dictionaries
        are materialised only if somebody asks for them.
    When neither is given the filter sets are built from this graph's own facts, as the
    reference does in ``evaluate_dicts`` (data_structures.py:386-397).
    """

    def __init__(self, heads, tails, relations, n_ent, n_rel, dict_of_heads=None,
                 dict_of_tails=None, filter_facts=None):
        if not (heads.shape == tails.shape == relations.shape and heads.dim() == 1):
            raise ValueError("heads, tails, relations must be 1-D tensors of equal length")
        self.head_idx = heads.long().cpu()
        self.tail_idx = tails.long().cpu()
        self.relations = relations.long().cpu()
        self.n_ent, self.n_rel = int(n_ent), int(n_rel)
        self.n_facts = int(heads.shape[0])
        self.filter_index = None
        self._dict_of_heads, self._dict_of_tails = dict_of_heads, dict_of_tails
        if dict_of_heads is None or dict_of_tails is None:
            fh, ft, fr = filter_facts if filter_facts is not None else (self.head_idx, self.tail_idx,
                                                                        self.relations)
            self.filter_index = FilterIndex(fh, ft, fr, self.n_ent, self.n_rel)

    def _dicts(self):
        if self._dict_of_heads is None or self._dict_of_tails is None:
            self._dict_of_heads, self._dict_of_tails = self.filter_index.as_dicts()
        return self._dict_of_heads, self._dict_of_tails

    @property
    def dict_of_heads(self):
        return self._dicts()[0]

    @property
    def dict_of_tails(self):
        return self._dicts()[1]

    @property
    def dict_of_rels(self):
        """(h, t) -> set of relations, as ``evaluate_dicts`` builds it (data_structures.py:386-397);
        built from this graph's own facts on first use unless ``dict_of_rels`` was assigned."""
        if getattr(self, "_dict_of_rels", None) is None:
            d = defaultdict(set)
            for h, t, r in zip(self.head_idx.tolist(), self.tail_idx.tolist(), self.relations.tolist()):
                d[(h, t)].add(r)
            self._dict_of_rels = d
        return self._dict_of_rels

    @dict_of_rels.setter
    def dict_of_rels(self, value):
        self._dict_of_rels = value

    def __len__(self):
        return self.n_facts


class DataLoader:
    """Batches of (heads, tails, relations) index tensors of a graph, in fact order, for the tutorial
    training loop ``for h, t, r in DataLoader(kg, batch_size, use_cuda='all')`` (torchkge/utils/data.py:
    83-151; no shuffling there either).  ``use_cuda``: None (batches stay where the graph's tensors
    are), 'all' (the three index tensors are moved to the current CUDA device once) or 'batch' (every
    batch is moved when it is handed out)."""

    def __init__(self, kg, batch_size, use_cuda=None):
        if use_cuda not in (None, "all", "batch"):
            raise ValueError("use_cuda must be None, 'all' or 'batch'")
        if int(batch_size) < 1:
            raise ValueError("batch_size must be a positive integer")
        self.batch_size, self.use_cuda = int(batch_size), use_cuda
        self._columns = [kg.head_idx, kg.tail_idx, kg.relations]
        if use_cuda == "all":
            self._columns = [x.cuda() for x in self._columns]

    def __len__(self):
        return -(-self._columns[0].shape[0] // self.batch_size)

    def __iter__(self):
        n = self._columns[0].shape[0]
        for lo in range(0, n, self.batch_size):
            batch = tuple(x[lo:lo + self.batch_size] for x in self._columns)
            yield tuple(x.cuda() for x in batch) if self.use_cuda == "batch" else batch


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


def expand_unique_sets(row_uid, has_true, uoffs, flat, true_idx):
    """Device-side CSR from per-row references to unique sets.

    row_uid (n,) int64: which unique set a row uses (any value where has_true is False);
    has_true (n,) bool: the row's true entity is a member of its set (otherwise the row is left
    unfiltered, get_true_targets' KeyError quirk); uoffs (u+1,), flat (m,): the unique sets, CSR;
    true_idx (n,).  All on one device.  Returns (offs (n+1,), ids, row of every entry (int32)) with
    the true entity removed.
    """
    dev = row_uid.device
    n = row_uid.numel()
    uid = row_uid.clamp(min=0)
    lens = (uoffs[1:] - uoffs[:-1])[uid] if uoffs.numel() > 1 else torch.zeros_like(uid)
    zero = torch.zeros_like(lens)
    full = torch.where(has_true, lens, zero)
    cnt = torch.where(has_true, lens - 1, zero)
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt, 0, out=offs[1:])
    row = torch.repeat_interleave(torch.arange(n, device=dev), full)
    start = torch.cumsum(full, 0) - full
    pos = torch.arange(row.numel(), device=dev) - start[row] + uoffs[uid][row]
    ids = flat[pos]
    keep = ids != true_idx[row]
    return offs, ids[keep].contiguous(), row[keep].to(torch.int32).contiguous()


def dict_filter_csr(kg, which, key1, key2, true_idx, device):
    """Device CSR of the filter sets of ``kg``'s facts taken from the REFERENCE's containers
    (``dict_of_tails[(h, r)]`` / ``dict_of_heads[(t, r)]``: defaultdict(set),
    data_structures.py:386-397), with ``filter_csr``'s semantics, fast enough to sit in front of
    the kernels: each DISTINCT key's set is flattened once (test sets repeat popular keys), the
    per-row lists are expanded on the device, and the result is cached on the graph object (keyed on
    the dictionary's identity and size and on the index tensors), so that every evaluation after
    the first one does no host work at all.  ``kg._b200_filter_cache = {}`` drops the cache (needed
    only if the dictionaries are edited in place between evaluations)."""
    import itertools

    import numpy as np
    dictionary = kg.dict_of_tails if which == "tail" else kg.dict_of_heads
    cache = kg.__dict__.setdefault("_b200_filter_cache", {})
    key = (which, str(device), id(dictionary), len(dictionary), key1.data_ptr(), key2.data_ptr(),
           true_idx.data_ptr(), int(key1.shape[0]), int(key1.sum()), int(key2.sum()), int(true_idx.sum()))
    hit = cache.get(key)
    if hit is not None:
        return hit      # ((offs, ids, row of entry), bytes uploaded)
    n = int(key1.shape[0])
    get = dictionary.get
    # per-row work in C-level loops (zip / dict.fromkeys / comprehensions): ~0.1 us per row
    keys = list(zip(key1.tolist(), key2.tolist()))
    uid_of = {k: u for u, k in enumerate(dict.fromkeys(keys))}          # distinct keys, first-seen order
    found = [get(k) for k in uid_of]                                     # their sets (None: unknown key)
    empty = frozenset()
    sets = [x if x is not None else empty for x in found]
    uids = [uid_of[k] for k in keys]
    row_uid = np.fromiter(uids, dtype=np.int64, count=n)
    # get_true_targets: a row whose true entity is not in its set is left unfiltered
    has_true = np.fromiter((c in sets[u] for c, u in zip(true_idx.tolist(), uids)), dtype=bool, count=n)
    lens = np.fromiter(map(len, sets), dtype=np.int64, count=len(sets))
    uoffs = np.zeros(len(sets) + 1, dtype=np.int64)
    np.cumsum(lens, out=uoffs[1:])
    flat = np.fromiter(itertools.chain.from_iterable(sets), dtype=np.int64, count=int(uoffs[-1]))
    to = lambda x: torch.from_numpy(x).to(device, non_blocking=True)   # noqa: E731
    csr = expand_unique_sets(to(row_uid), to(has_true), to(uoffs), to(flat),
                             true_idx.to(device, non_blocking=True))
    nbytes = 8 * (row_uid.size + uoffs.size + flat.size) + has_true.size
    while len(cache) >= 4:          # tail + head of the current test set, and one generation back
        cache.pop(next(iter(cache)))
    cache[key] = (csr, nbytes)
    return cache[key]
