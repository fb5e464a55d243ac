"""``LinkPredictionEvaluator`` with the reference's constructor, attributes and metric
methods (torchkge/evaluation.py:207-425).  ``evaluate`` hands the whole job -- scoring every
entity as head and as tail of every fact, discounting the filter sets, ranking -- to the
CUDA engine; no (batch, n_entities) score matrix exists at any point.
"""
import torch

from . import _lib
from .data import dict_filter_csr, filter_csr
from .engine import (DEFAULT_CHUNK, ModelSpec, QueryShard, default_engine, rank_link_prediction,
                     rank_relation_prediction)
from .exceptions import NotYetEvaluatedError


def _check_index_range(heads, tails, rels, n_ent, n_rel):
    """nn.Embedding raises on an out-of-range index; the kernels would read out of bounds.  The index
    tensors of a graph live on the host, so the check is a few microseconds there."""
    for name, x, hi in (("head", heads, n_ent), ("tail", tails, n_ent), ("relation", rels, n_rel)):
        if x.numel() and not x.is_cuda and (int(x.min()) < 0 or int(x.max()) >= hi):
            raise IndexError("%s index out of range [0, %d)" % (name, hi))


class LinkPredictionEvaluator(object):
    """Evaluate an embedding model by link prediction (Bordes et al. 2013).

    Parameters
    ----------
    model: a model of ``torchkge_b200.models`` (or the reference's class of the same name),
        living on a CUDA device.
    knowledge_graph: object exposing ``n_facts, head_idx, tail_idx, relations,
        dict_of_heads, dict_of_tails`` (``torchkge.data_structures.KnowledgeGraph`` or
        ``torchkge_b200.data.KnowledgeGraph``).
    shard: ``torchkge_b200.engine.EntityShard`` or ``QueryShard``, optional (extension).
        EntityShard: every rank of the group scans only its range of entity rows and the
        per-fact counters are summed with one all-reduce.  QueryShard: every rank ranks its
        contiguous slice of the facts against the whole (replicated) table and the rank vectors
        are all-gathered.  Either way all ranks end up with the full rank vectors.

    Attributes (as in the reference, evaluation.py:236-261)
    ----------
    rank_true_heads, rank_true_tails, filt_rank_true_heads, filt_rank_true_tails:
        torch.LongTensor (n_facts,), on CPU after ``evaluate``.
    evaluated: bool
    """

    def __init__(self, model, knowledge_graph, shard=None):
        self.model = model
        self.kg = knowledge_graph
        n = knowledge_graph.n_facts
        self.rank_true_heads = torch.empty(size=(n,)).long()
        self.rank_true_tails = torch.empty(size=(n,)).long()
        self.filt_rank_true_heads = torch.empty(size=(n,)).long()
        self.filt_rank_true_tails = torch.empty(size=(n,)).long()
        self.evaluated = False
        self.shard = shard
        self.last_stats = {}

    def evaluate(self, b_size, verbose=True):
        """Rank all facts of the graph.

        ``b_size`` is accepted for signature compatibility; as in the reference it never
        changes the result.  Here it does not bound memory either (nothing of size
        b_size x n_entities is allocated); facts are processed in chunks of
        ``engine.DEFAULT_CHUNK``.  ``verbose`` is accepted and ignored (no per-batch loop to
        report on).
        """
        if b_size is None or int(b_size) < 1:
            raise ValueError("b_size must be a positive integer")
        spec = ModelSpec.from_model(self.model)
        qshard = self.shard if isinstance(self.shard, QueryShard) else None
        eshard = None if qshard is not None else self.shard
        if eshard is not None and eshard.local_storage:
            spec.ent_lo, spec.n_ent = eshard.lo, eshard.n_ent
        if not spec.ent0.is_cuda:
            raise _lib.KgeLibraryError(
                "LinkPredictionEvaluator.evaluate needs the model on a CUDA device "
                "(model.cuda()); this package has no CPU execution path")
        dev = spec.ent0.device
        kg = self.kg
        heads, tails, rels = kg.head_idx, kg.tail_idx, kg.relations
        if qshard is not None:      # this rank's contiguous slice of the facts
            heads, tails, rels = (x[qshard.lo:qshard.hi] for x in (heads, tails, rels))
        n_here = int(heads.shape[0])
        _check_index_range(heads, tails, rels, spec.n_ent, spec.n_rel)
        h_d = heads.to(dev, non_blocking=True)
        t_d = tails.to(dev, non_blocking=True)
        r_d = rels.to(dev, non_blocking=True)
        stats = {"h2d_bytes": 8 * 3 * n_here, "d2h_bytes": 8 * (4 * kg.n_facts + 1)}

        # Filter sets -> CSR (same per-row semantics as get_true_targets).  Built lazily: the
        # engine asks for them after the dense scans are enqueued, so host work overlaps the GPU.
        index = getattr(kg, "filter_index", None)
        if index is not None:
            # sorted-array filters (torchkge_b200.data.KnowledgeGraph): the index is resident on
            # the device (uploaded at first use, like weights); the per-row lists of this test set
            # come from searchsorted + gather on the device
            # ... once per (graph, test slice, device): the lists depend only on the graph, so they are
            # kept on the graph object like the index itself (keyed on the host index tensors)
            def from_index(which, k1, k2, true, k1_d, k2_d, true_d):
                def build():
                    cache = kg.__dict__.setdefault("_b200_filter_cache", {})
                    # identity of the host tensors + a cheap content fingerprint (in-place edits of a
                    # test set are unusual, but must not be served stale lists)
                    key = (which, str(dev), id(index), k1.data_ptr(), k2.data_ptr(), true.data_ptr(), n_here,
                           int(k1.sum()), int(k2.sum()), int(true.sum()))
                    hit = cache.get(key)
                    if hit is None:
                        while len(cache) >= 4:
                            cache.pop(next(iter(cache)))
                        hit = cache[key] = index.csr(which, k1_d, k2_d, true_d)
                    return hit
                return build
            csr_tail = from_index("tail", heads, rels, tails, h_d, r_d, t_d)
            csr_head = from_index("head", tails, rels, heads, t_d, r_d, h_d)
        else:
            # the reference's dictionaries (torchkge.data_structures.KnowledgeGraph): distinct
            # keys flattened once on the host, expanded on the device, cached on the graph
            def from_dicts(which, k1, k2, true):
                def build():
                    csr, nbytes = dict_filter_csr(kg, which, k1, k2, true, dev)
                    stats["h2d_bytes"] += nbytes
                    return csr
                return build
            csr_tail = from_dicts("tail", heads, rels, tails)
            csr_head = from_dicts("head", tails, rels, heads)
        engine = default_engine()
        lazy = rank_link_prediction(spec, h_d, t_d, r_d, csr_tail, csr_head, shard=eshard,
                                    engine=engine, chunk=DEFAULT_CHUNK, sync=False)

        def to_host(ranks, flag):
            flag = torch.zeros(1, dtype=torch.int64, device=dev) if flag is None else flag.long().view(1)
            if qshard is not None:
                ranks = qshard.all_gather(ranks)
                flag = qshard.all_reduce_sum(flag)   # every rank takes the same decision below
            return torch.cat([x.view(-1) for x in ranks] + [flag]).cpu()   # ONE device -> host copy

        host = to_host(lazy.ranks, lazy.overflow)
        if int(host[-1]) > 0:        # near-tie list overflow somewhere: exact recomputation
            host = to_host(lazy.get(flag_host=int(host[-1])), None)
        n = kg.n_facts
        self.last_stats = stats
        self.rank_true_heads, self.rank_true_tails = host[0:n], host[n:2 * n]
        self.filt_rank_true_heads, self.filt_rank_true_tails = host[2 * n:3 * n], host[3 * n:4 * n]
        self.evaluated = True

    def _check(self):
        if not self.evaluated:
            raise NotYetEvaluatedError('Evaluator not evaluated call '
                                       'LinkPredictionEvaluator.evaluate')

    def mean_rank(self):
        """(mean rank, filtered mean rank), heads and tails averaged (evaluation.py:310-330)."""
        self._check()
        raw = (self.rank_true_heads.float().mean() + self.rank_true_tails.float().mean()).item()
        filt = (self.filt_rank_true_heads.float().mean()
                + self.filt_rank_true_tails.float().mean()).item()
        return raw / 2, filt / 2

    def hit_at_k_heads(self, k=10):
        self._check()
        return ((self.rank_true_heads <= k).float().mean().item(),
                (self.filt_rank_true_heads <= k).float().mean().item())

    def hit_at_k_tails(self, k=10):
        self._check()
        return ((self.rank_true_tails <= k).float().mean().item(),
                (self.filt_rank_true_tails <= k).float().mean().item())

    def hit_at_k(self, k=10):
        """(Hits@k, filtered Hits@k), heads and tails averaged (evaluation.py:354-374)."""
        self._check()
        hh, fhh = self.hit_at_k_heads(k=k)
        th, fth = self.hit_at_k_tails(k=k)
        return (hh + th) / 2, (fhh + fth) / 2

    def mrr(self):
        """(MRR, filtered MRR), heads and tails averaged (evaluation.py:376-397)."""
        self._check()
        head = (self.rank_true_heads.float() ** (-1)).mean()
        tail = (self.rank_true_tails.float() ** (-1)).mean()
        fhead = (self.filt_rank_true_heads.float() ** (-1)).mean()
        ftail = (self.filt_rank_true_tails.float() ** (-1)).mean()
        return (head + tail).item() / 2, (fhead + ftail).item() / 2

    def print_results(self, k=None, n_digits=3):
        """Same report as the reference (evaluation.py:399-425)."""
        if k is None:
            k = 10
        if type(k) == int:
            k = [k]
        for i in k:
            print('Hit@{} : {} \t\t Filt. Hit@{} : {}'.format(
                i, round(self.hit_at_k(k=i)[0], n_digits),
                i, round(self.hit_at_k(k=i)[1], n_digits)))
        print('Mean Rank : {} \t Filt. Mean Rank : {}'.format(
            int(self.mean_rank()[0]), int(self.mean_rank()[1])))
        print('MRR : {} \t\t Filt. MRR : {}'.format(
            round(self.mrr()[0], n_digits), round(self.mrr()[1], n_digits)))


class RelationPredictionEvaluator(object):
    """Evaluate an embedding model by relation prediction (torchkge/evaluation.py:16-204): every
    fact's true relation is ranked against all relations, raw and filtered by
    ``knowledge_graph.dict_of_rels``.

    Parameters
    ----------
    model: TransE (L1/L2), DistMult or ComplEx model on a CUDA device.
    knowledge_graph: object exposing ``n_facts, head_idx, tail_idx, relations, dict_of_rels``.
    directed: bool (default True).  False: both (h, ?, t) and (t, ?, h) are scored and ranked
        together against the directed true score (evaluation.py:99-107).

    Attributes: ``rank_true_rels``, ``filt_rank_true_rels`` (LongTensor (n_facts,), CPU after
    ``evaluate``), ``evaluated``, ``directed``.
    """

    def __init__(self, model, knowledge_graph, directed=True):
        self.model = model
        self.kg = knowledge_graph
        self.directed = directed
        self.rank_true_rels = torch.empty(size=(knowledge_graph.n_facts,)).long()
        self.filt_rank_true_rels = torch.empty(size=(knowledge_graph.n_facts,)).long()
        self.evaluated = False

    def evaluate(self, b_size, verbose=True):
        """``b_size`` / ``verbose`` are accepted for signature compatibility (see
        ``LinkPredictionEvaluator.evaluate``)."""
        if b_size is None or int(b_size) < 1:
            raise ValueError("b_size must be a positive integer")
        spec = ModelSpec.from_model(self.model)
        if not spec.ent0.is_cuda:
            raise _lib.KgeLibraryError(
                "RelationPredictionEvaluator.evaluate needs the model on a CUDA device "
                "(model.cuda()); this package has no CPU execution path")
        dev = spec.ent0.device
        kg = self.kg
        _check_index_range(kg.head_idx, kg.tail_idx, kg.relations, spec.n_ent, spec.n_rel)
        h_d, t_d, r_d = (x.to(dev, non_blocking=True) for x in (kg.head_idx, kg.tail_idx, kg.relations))
        csr = filter_csr(kg.dict_of_rels, kg.head_idx, kg.tail_idx, kg.relations)
        csr = tuple(x.to(dev, non_blocking=True) for x in csr)
        rr, frr = rank_relation_prediction(spec, h_d, t_d, r_d, csr, directed=self.directed,
                                           engine=default_engine(), chunk=DEFAULT_CHUNK)
        self.rank_true_rels = rr.cpu()
        self.filt_rank_true_rels = frr.cpu()
        self.evaluated = True

    def _check(self):
        if not self.evaluated:
            raise NotYetEvaluatedError('Evaluator not evaluated call '
                                       'LinkPredictionEvaluator.evaluate')

    def mean_rank(self):
        """(mean rank, filtered mean rank) of the true relation (evaluation.py:114-131)."""
        self._check()
        return self.rank_true_rels.float().mean().item(), self.filt_rank_true_rels.float().mean().item()

    def hit_at_k(self, k=10):
        """(Hit@k, filtered Hit@k) (evaluation.py:133-154)."""
        self._check()
        return ((self.rank_true_rels <= k).float().mean().item(),
                (self.filt_rank_true_rels <= k).float().mean().item())

    def mrr(self):
        """(MRR, filtered MRR) (evaluation.py:156-174)."""
        self._check()
        return ((self.rank_true_rels.float() ** (-1)).mean().item(),
                (self.filt_rank_true_rels.float() ** (-1)).mean().item())

    def print_results(self, k=None, n_digits=3):
        """Same report as the reference (evaluation.py:176-204)."""
        if k is None:
            k = 10
        if k is not None and type(k) == int:
            print('Hit@{} : {} \t\t Filt. Hit@{} : {}'.format(
                k, round(self.hit_at_k(k=k)[0], n_digits), k, round(self.hit_at_k(k=k)[1], n_digits)))
        if k is not None and type(k) == list:
            for i in k:
                print('Hit@{} : {} \t\t Filt. Hit@{} : {}'.format(
                    i, round(self.hit_at_k(k=i)[0], n_digits), i, round(self.hit_at_k(k=i)[1], n_digits)))
        print('Mean Rank : {} \t Filt. Mean Rank : {}'.format(
            int(self.mean_rank()[0]), int(self.mean_rank()[1])))
        print('MRR : {} \t\t Filt. MRR : {}'.format(
            round(self.mrr()[0], n_digits), round(self.mrr()[1], n_digits)))


class TripletClassificationEvaluator(object):
    """Evaluate an embedding model by triplet classification (Socher et al. 2013),
    torchkge/evaluation.py:428-580: one threshold per relation (the best-scoring negative of the
    validation facts of that relation), then accuracy on the test facts and their negatives.

    Scores come from ``model.scoring_function`` (the CUDA per-triple scorer), negatives from
    ``PositionalNegativeSampler(kg_val, kg_test=kg_test)`` as in the reference; ``sampler`` may be
    replaced after construction.
    """

    def __init__(self, model, kg_val, kg_test):
        from .sampling import PositionalNegativeSampler
        self.model = model
        self.kg_val = kg_val
        self.kg_test = kg_test
        self.is_cuda = next(self.model.parameters()).is_cuda
        self.evaluated = False
        self.thresholds = None
        self.sampler = PositionalNegativeSampler(self.kg_val, kg_test=self.kg_test)

    def get_scores(self, heads, tails, relations, batch_size):
        """Scores of the given triplets, computed batch by batch (evaluation.py:478-511)."""
        if not self.is_cuda:
            raise _lib.KgeLibraryError("TripletClassificationEvaluator needs the model on a CUDA device")
        dev = next(self.model.parameters()).device
        scores = []
        with torch.no_grad():
            for lo in range(0, heads.shape[0], batch_size):
                sl = slice(lo, lo + batch_size)
                scores.append(self.model.scoring_function(heads[sl].to(dev), tails[sl].to(dev),
                                                          relations[sl].to(dev)))
        return torch.cat(scores, dim=0)

    def evaluate(self, b_size):
        """Thresholds from the validation graph (evaluation.py:513-541): for relation i the largest
        score among the negatives of its validation facts; relations absent from the validation set
        get the largest negative score overall."""
        r_idx = self.kg_val.relations
        neg_heads, neg_tails = self.sampler.corrupt_kg(b_size, self.is_cuda, which='main')
        neg_scores = self.get_scores(neg_heads, neg_tails, r_idx, b_size)
        n_rel = self.kg_val.n_rel
        r_dev = r_idx.to(neg_scores.device)
        per_rel = torch.full((n_rel,), -float("inf"), device=neg_scores.device)
        per_rel = per_rel.scatter_reduce(0, r_dev, neg_scores, reduce="amax", include_self=True)
        present = torch.bincount(r_dev, minlength=n_rel) > 0
        self.thresholds = torch.where(present, per_rel, neg_scores.max()).detach().cpu()
        self.evaluated = True

    def accuracy(self, b_size):
        """Share of test facts scored above, and of their negatives scored below, the threshold
        of their relation (evaluation.py:543-580)."""
        if not self.evaluated:
            self.evaluate(b_size)
        r_idx = self.kg_test.relations
        neg_heads, neg_tails = self.sampler.corrupt_kg(b_size, self.is_cuda, which='test')
        scores = self.get_scores(self.kg_test.head_idx, self.kg_test.tail_idx, r_idx, b_size)
        neg_scores = self.get_scores(neg_heads, neg_tails, r_idx, b_size)
        if self.is_cuda:
            self.thresholds = self.thresholds.to(scores.device)
        thr = self.thresholds[r_idx.to(self.thresholds.device)]
        return ((scores > thr).sum().item() + (neg_scores < thr).sum().item()) / (2 * self.kg_test.n_facts)
