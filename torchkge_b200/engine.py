"""Host-side driver of the CUDA engine: turns torch tensors into C-ABI calls.

Nothing here computes scores: torch is used for device memory, streams and (when the entity
table is range-partitioned) the two NCCL all-reduces.  The only engine is the CUDA one; tests
may substitute an object with the same four methods (``pack``, ``gather_rows``,
``rank_side``, ``score_all``) to exercise the sharding logic on CPU.
"""
import ctypes
import os

import torch

from . import _lib

#: default number of test triples ranked per kernel launch (bounds workspace memory only;
#: results never depend on it)
DEFAULT_CHUNK = 65536


class ModelSpec:
    """What the kernels read from a model: which score function, and raw fp32 tables.

    Mirrors the state_dict contract of the reference models:
    ``ent_emb.weight`` / ``rel_emb.weight`` (TransE models/translation.py:63-64, DistMult
    models/bilinear.py:183-184), ``ent_emb.weight`` / ``rel_mat.weight`` (RESCAL
    models/bilinear.py:55-56), ``re_/im_ent_emb.weight`` + ``re_/im_rel_emb.weight``
    (ComplEx models/bilinear.py:455-458); Analogy (models/bilinear.py:623-631) has three planes per
    table -- ``sc_/re_/im_ent_emb.weight``, ``sc_/re_/im_rel_emb.weight`` -- handed over as views of one
    stacked (3, n, dim) copy (``stacked``): the C ABI takes planes 0 and 1 and finds plane 2 at the
    same spacing (include/kge_b200.h, "three-plane tables").
    """

    def __init__(self, code, dim, n_ent, n_rel, ent0, ent1, rel0, rel1, ent_lo=0, ent2=None, rel2=None):
        self.code = code
        self.dim = int(dim)
        self.n_ent = int(n_ent)      # global number of entities
        self.n_rel = int(n_rel)
        self.ent0, self.ent1, self.rel0, self.rel1 = ent0, ent1, rel0, rel1
        self.ent2, self.rel2 = ent2, rel2
        self.ent_lo = int(ent_lo)    # global id of row 0 of ent0/ent1
        self.n_rows = int(ent0.shape[0])
        if code == _lib.ANALOGY:
            for name, planes in (("entity", (ent0, ent1, ent2)), ("relation", (rel0, rel1, rel2))):
                if planes[0] is None and name == "relation":
                    continue         # relation prediction: the candidates are the relation rows
                if any(x is None for x in planes) or not _equally_spaced(*planes):
                    raise ValueError("Analogy %s planes must be equally spaced views of one stacked "
                                     "tensor (ModelSpec.stacked)" % name)

    @property
    def cand_planes(self):
        return 3 if self.code == _lib.ANALOGY else (2 if self.code in (_lib.COMPLEX, _lib.ROTATE) else 1)

    @staticmethod
    def stacked(planes):
        """[plane tensors (n, d)] -> the views (p0, p1, p2) of one contiguous (3, n, d) copy."""
        s = torch.stack([ModelSpec._f32(x) for x in planes])
        return s[0], s[1], s[2]

    def narrowed(self, lo, hi):
        """Same model restricted to entity rows [lo, hi) (views, no copy)."""
        a, b = lo - self.ent_lo, hi - self.ent_lo
        if a < 0 or b > self.n_rows:
            raise ValueError("shard [%d,%d) outside held rows" % (lo, hi))
        e1 = None if self.ent1 is None else self.ent1[a:b]
        e2 = None if self.ent2 is None else self.ent2[a:b]
        return ModelSpec(self.code, self.dim, self.n_ent, self.n_rel, self.ent0[a:b], e1,
                         self.rel0, self.rel1, ent_lo=lo, ent2=e2, rel2=self.rel2)

    @staticmethod
    def _f32(t):
        t = t.detach()
        if t.dtype != torch.float32:
            raise TypeError("embedding tables must be float32, got %s" % t.dtype)
        return t.contiguous()

    @classmethod
    def from_model(cls, model):
        """Accepts torchkge_b200 models and (duck-typed) the reference's own model classes."""
        name = type(model).__name__
        f = cls._f32
        if name == "TransEModel":
            dname = getattr(model.dissimilarity, "__name__", str(model.dissimilarity))
            if dname == "l1_dissimilarity":
                code = _lib.TRANSE_L1
            elif dname == "l2_dissimilarity":
                code = _lib.TRANSE_L2
            else:
                raise NotImplementedError("TransE dissimilarity %s is not on the CUDA path" % dname)
            return cls(code, model.emb_dim, model.n_ent, model.n_rel,
                       f(model.ent_emb.weight), None, f(model.rel_emb.weight), None)
        if name == "TorusEModel":
            dname = getattr(model.dissimilarity, "__name__", str(model.dissimilarity))
            code = {"l1_torus_dissimilarity": _lib.TORUSE_L1, "l2_torus_dissimilarity": _lib.TORUSE_L2,
                    "l1_dissimilarity": _lib.TRANSE_L1}.get(dname)
            if code is None:
                raise NotImplementedError("TorusE dissimilarity %s is not on the CUDA path" % dname)
            if not getattr(model, "normalized", False):
                model.normalize_parameters()   # as inference_prepare_candidates does (translation.py:745-746)
            return cls(code, model.emb_dim, model.n_ent, model.n_rel,
                       f(model.ent_emb.weight), None, f(model.rel_emb.weight), None)
        if name == "DistMultModel":
            return cls(_lib.DISTMULT, model.emb_dim, model.n_ent, model.n_rel,
                       f(model.ent_emb.weight), None, f(model.rel_emb.weight), None)
        if name == "RESCALModel":
            return cls(_lib.RESCAL, model.emb_dim, model.n_ent, model.n_rel,
                       f(model.ent_emb.weight), None, f(model.rel_mat.weight), None)
        if name == "ComplExModel":
            return cls(_lib.COMPLEX, model.emb_dim, model.n_ent, model.n_rel,
                       f(model.re_ent_emb.weight), f(model.im_ent_emb.weight),
                       f(model.re_rel_emb.weight), f(model.im_rel_emb.weight))
        if name == "RotatEModel":
            re_r, im_r = model.relation_planes()
            return cls(_lib.ROTATE, model.emb_dim, model.n_ent, model.n_rel,
                       f(model.re_ent_emb.weight), f(model.im_ent_emb.weight), f(re_r), f(im_r))
        if name == "AnalogyModel":
            if model.scalar_dim != model.complex_dim:
                # the reference's own inference_scoring_function adds (b, n, scalar_dim) and
                # (b, n, complex_dim) tensors (bilinear.py:695-698): it needs equal widths too
                raise NotImplementedError("Analogy link prediction needs scalar_dim == complex_dim "
                                          "(got %d and %d)" % (model.scalar_dim, model.complex_dim))
            e = cls.stacked([model.sc_ent_emb.weight, model.re_ent_emb.weight, model.im_ent_emb.weight])
            r = cls.stacked([model.sc_rel_emb.weight, model.re_rel_emb.weight, model.im_rel_emb.weight])
            return cls(_lib.ANALOGY, model.scalar_dim, model.n_ent, model.n_rel, e[0], e[1], r[0], r[1],
                       ent2=e[2], rel2=r[2])
        raise NotImplementedError(
            "%s has no CUDA link-prediction path (supported: TransE L1/L2, TorusE torus_L1/torus_L2, "
            "DistMult, RESCAL, ComplEx, Analogy, RotatE)" % name)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _equally_spaced(p0, p1, p2):
    """planes of a three-plane table as the C ABI expects them: plane 2 at p1 + (p1 - p0)"""
    return (p0.shape == p1.shape == p2.shape and p0.is_contiguous() and p1.is_contiguous() and p2.is_contiguous()
            and p2.data_ptr() - p1.data_ptr() == p1.data_ptr() - p0.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.KgeLibraryError(
                "the link-prediction engine runs on CUDA tensors only (got a %s tensor); "
                "move the model to a GPU -- there is no CPU fallback" % t.device)


def _device_guard(device):
    """The library launches on the CURRENT device (cudaGetDevice) with the stream it is handed:
    every public entry point runs its calls under the device of the tensors it was given."""
    if getattr(device, "type", None) == "cuda":
        return torch.cuda.device(device)
    import contextlib
    return contextlib.nullcontext()


class CudaEngine:
    """Direct calls into libkge_b200.so on the current CUDA stream."""

    def __init__(self, tensor_core=None):
        self.lib = _lib.load()
        self.launches = 0  # kernels launched through this engine (bench.py reports it)
        if tensor_core is None:
            tensor_core = os.environ.get("KGE_TENSOR_CORE", "1") != "0"
        #: use the tcgen05 bound-and-refine scan for models that have one (ranks unchanged)
        self.tensor_core = bool(tensor_core)
        self.tc_stats = []  # (device tensor [found, capacity]) per tensor-core call, for checks
        #: tensor-core operand images kept between evaluations, each with a device-side content
        #: checksum of the table it was built from (kge_tc_pack_table_cached): an unchanged table
        #: costs one read instead of a rebuild, a changed one is always rebuilt.  KGE_TC_CACHE=0 or
        #: ``tc_cache_entries = 0`` disables it; ``clear_cache()`` frees the images.
        self.tc_cache_entries = 0 if os.environ.get("KGE_TC_CACHE", "1") == "0" else 2
        self._tc_cache = {}
        #: KGE_TRACE=1: rank_link_prediction appends (label, host seconds, CUDA event) marks here
        self.trace = [] if os.environ.get("KGE_TRACE") else None

    def mark(self, label):
        """Debug aid (KGE_TRACE=1): a host timestamp and a CUDA event on the current stream."""
        if self.trace is None:
            return
        import time
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.trace.append((label, time.perf_counter(), ev))

    def trace_report(self):
        """[(label, host ms since first mark, device ms since first mark)] of the recorded marks."""
        if not self.trace:
            return []
        torch.cuda.synchronize()
        l0, t0, e0 = self.trace[0]
        out = [(lab, 1e3 * (t - t0), e0.elapsed_time(ev)) for lab, t, ev in self.trace]
        self.trace = []
        return out

    # ---- table packing: once per evaluate() ----
    def pack(self, spec):
        _need_cuda(spec.ent0, spec.ent1)
        n_floats = self.lib.kge_packed_table_floats(spec.code, spec.n_rows, spec.dim)
        packed = torch.empty(max(n_floats, 1), dtype=torch.float32, device=spec.ent0.device)
        if spec.n_rows > 0:
            _lib.check(self.lib.kge_pack_table(spec.code, _ptr(spec.ent0), _ptr(spec.ent1),
                                               spec.n_rows, spec.dim, _ptr(packed),
                                               _stream(packed.device)), "kge_pack_table")
            self.launches += 1
        return packed

    def clear_cache(self):
        self._tc_cache.clear()

    def pack_tc(self, spec):
        """Tensor-core operand image of the shard, or None when the model has no such path."""
        if not self.tensor_core:
            return None
        nbytes = self.lib.kge_tc_packed_bytes(spec.code, spec.n_rows, spec.dim)
        if nbytes == 0:
            return None
        dev = spec.ent0.device
        # a three-plane spec reads a stacked COPY of the weights made for this call (ModelSpec.stacked):
        # its address changes every time, so there is nothing to cache an image under
        if self.tc_cache_entries <= 0 or getattr(spec, "ent2", None) is not None:
            out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(self.lib.kge_tc_pack_table(spec.code, _ptr(spec.ent0), _ptr(spec.ent1), spec.n_rows,
                                                  spec.dim, _ptr(out), _stream(dev)), "kge_tc_pack_table")
            self.launches += 4
            return out
        key = (spec.code, spec.ent0.data_ptr(), None if spec.ent1 is None else spec.ent1.data_ptr(),
               spec.n_rows, spec.dim, self.lib.kge_tc_layout_id(), str(dev))
        hit = self._tc_cache.pop(key, None)
        if hit is None:
            while len(self._tc_cache) >= self.tc_cache_entries:
                self._tc_cache.pop(next(iter(self._tc_cache)))
            # the table tensors are kept too: their storage cannot be freed and re-used at the same
            # address by another table while the entry lives
            hit = (torch.empty(nbytes, dtype=torch.uint8, device=dev),
                   torch.zeros(4, dtype=torch.int64, device=dev), spec.ent0, spec.ent1)
        self._tc_cache[key] = hit        # most recently used last
        out, guard = hit[0], hit[1]
        _lib.check(self.lib.kge_tc_pack_table_cached(spec.code, _ptr(spec.ent0), _ptr(spec.ent1), spec.n_rows,
                                                     spec.dim, _ptr(out), _ptr(guard), _stream(dev)),
                   "kge_tc_pack_table_cached")
        self.launches += 8 if spec.ent1 is None else 9
        return out

    def gather_rows(self, spec, idx):
        _need_cuda(spec.ent0, idx)
        n = idx.shape[0]
        out = torch.empty((n, spec.cand_planes, spec.dim), dtype=torch.float32,
                          device=spec.ent0.device)
        _lib.check(self.lib.kge_gather_rows(spec.code, _ptr(spec.ent0), _ptr(spec.ent1),
                                            spec.ent_lo, spec.n_rows, spec.dim, _ptr(idx), n,
                                            _ptr(out), _stream(out.device)), "kge_gather_rows")
        self.launches += 1
        return out

    def rank_side(self, spec, packed, side, hrows, trows, r_idx, true_idx, filt, raw_count,
                  filt_sub, true_score=None, tc_packed=None, tc_dump=None, true_rows=None,
                  true_score_in=None, approx=False, stats=None):
        """Adds this shard's counts for one side into raw_count / filt_sub (int32, device).
        ``stats``: optional int64[2] device tensor receiving (near-ties found, capacity) of a
        bound-and-refine call (one is allocated when absent)."""
        n = r_idx.shape[0] if r_idx is not None else hrows.shape[0]
        dev = raw_count.device
        flags = _lib.FLAG_TENSOR_CORE if tc_packed is not None else 0
        if tc_packed is None and approx and spec.code == _lib.ROTATE:
            flags = _lib.FLAG_APPROX_SCAN
        ws_bytes = self.lib.kge_rank_workspace_bytes(spec.code, side, spec.dim, n, spec.n_rows, flags)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a = _lib.RankArgs()
        a.model, a.side, a.dim, a.flags = spec.code, side, spec.dim, flags
        if flags:
            if stats is None:
                stats = torch.zeros(2, dtype=torch.int64, device=dev)
            a.tc_packed, a.tc_stats, a.tc_dump = _ptr(tc_packed), _ptr(stats), _ptr(tc_dump)
            self.tc_stats.append(stats)
            del self.tc_stats[:-256]
        else:
            stats = None
        a.n, a.n_ent, a.ent_lo, a.n_rows = n, spec.n_ent, spec.ent_lo, spec.n_rows
        a.packed, a.ent0, a.ent1 = _ptr(packed), _ptr(spec.ent0), _ptr(spec.ent1)
        a.rel0, a.rel1 = _ptr(spec.rel0), _ptr(spec.rel1)
        a.hrows, a.trows, a.r_idx, a.true_idx = _ptr(hrows), _ptr(trows), _ptr(r_idx), _ptr(true_idx)
        if filt is not None:
            offs, ids = filt[0], filt[1]
            a.filt_offs, a.filt_ids, a.n_filt = _ptr(offs), _ptr(ids), ids.shape[0]
            a.filt_qid = _ptr(filt[2]) if len(filt) > 2 else None
        a.raw_count, a.filt_sub, a.true_score = _ptr(raw_count), _ptr(filt_sub), _ptr(true_score)
        a.true_rows, a.true_score_in = _ptr(true_rows), _ptr(true_score_in)
        a.workspace, a.workspace_bytes, a.stream = _ptr(ws), ws_bytes, _stream(dev)
        _lib.check(self.lib.kge_rank_side(ctypes.byref(a)), "kge_rank_side")
        # prep, pack_queries, pad fill, true scores, scan (+ filter)
        self.launches += 5 + (1 if filt is not None and filt[1].shape[0] > 0 else 0)
        # handle for filter_side; keeps buffers alive
        return (a, ws, packed, hrows, trows, tc_packed, stats, true_rows, true_score_in)

    def filter_side(self, handle, filt, filt_sub):
        """Sparse filter pass for a side whose dense scan was enqueued earlier by rank_side
        (with filt=None); ``handle`` is what that call returned."""
        a = handle[0]
        offs, ids = filt[0], filt[1]
        if ids.shape[0] == 0:
            return
        a.filt_offs, a.filt_ids, a.n_filt = _ptr(offs), _ptr(ids), ids.shape[0]
        # optional third array: the CSR row of every entry (int32), spares the kernel a bisection per entry
        a.filt_qid = _ptr(filt[2]) if len(filt) > 2 and filt[2] is not None else None
        self._keep = filt
        a.filt_sub = _ptr(filt_sub)
        a.stream = _stream(filt_sub.device)
        _lib.check(self.lib.kge_filter_side(ctypes.byref(a)), "kge_filter_side")
        self.launches += 1

    def score_all(self, spec, packed, side, hrows, trows, r_idx):
        n = hrows.shape[0]
        dev = hrows.device
        scores = torch.empty((n, spec.n_rows), dtype=torch.float32, device=dev)
        ws_bytes = self.lib.kge_rank_workspace_bytes(spec.code, side, spec.dim, n, 0, 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a = _lib.ScoreAllArgs()
        a.model, a.side, a.dim = spec.code, side, spec.dim
        a.n, a.n_rows = n, spec.n_rows
        a.packed, a.rel0, a.rel1 = _ptr(packed), _ptr(spec.rel0), _ptr(spec.rel1)
        a.hrows, a.trows, a.r_idx = _ptr(hrows), _ptr(trows), _ptr(r_idx)
        a.scores, a.workspace, a.workspace_bytes = _ptr(scores), _ptr(ws), ws_bytes
        a.stream = _stream(dev)
        _lib.check(self.lib.kge_score_all(ctypes.byref(a)), "kge_score_all")
        self.launches += 4
        return scores

    def topk_side(self, spec, packed, side, hrows, trows, r_idx, k, mask=None):
        """(pred int64 (n, k), scores float32 (n, k)) of the k best candidates per query, exact
        scores, best first; ``mask`` = device CSR (offs, ids ascending per row) of candidates to set to
        -inf.  No (n, n_rows) matrix is allocated (include/kge_b200.h: kge_topk_side)."""
        n = hrows.shape[0]
        dev = hrows.device
        pred = torch.empty((n, k), dtype=torch.int64, device=dev)
        scores = torch.empty((n, k), dtype=torch.float32, device=dev)
        ws_bytes = self.lib.kge_topk_workspace_bytes(spec.code, side, spec.dim, n, spec.n_rows, k)
        if ws_bytes == 0:
            raise _lib.KgeLibraryError("kge_topk_workspace_bytes: unsupported arguments (k must be in [1, 1024])")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a = _lib.TopkArgs()
        a.model, a.side, a.dim, a.k = spec.code, side, spec.dim, k
        a.n, a.n_rows = n, spec.n_rows
        a.packed, a.rel0, a.rel1 = _ptr(packed), _ptr(spec.rel0), _ptr(spec.rel1)
        a.hrows, a.trows, a.r_idx = _ptr(hrows), _ptr(trows), _ptr(r_idx)
        if mask is not None and mask[1].numel() > 0:
            a.mask_offs, a.mask_ids = _ptr(mask[0]), _ptr(mask[1])
        a.pred, a.scores = _ptr(pred), _ptr(scores)
        a.workspace, a.workspace_bytes, a.stream = _ptr(ws), ws_bytes, _stream(dev)
        _lib.check(self.lib.kge_topk_side(ctypes.byref(a)), "kge_topk_side")
        self.launches += 7   # prep, pack, fill, finish + at least one collect scan / merge pair
        return pred, scores

    def rescal_rel_scores(self, spec, hrows, trows):
        """(n, n_rel) scores ((h^T M_c) * t).sum() of every relation matrix c: RESCAL's relation case
        (bilinear.py:115-121), dense (the candidates are per-fact vectors, nothing to scan)."""
        n, dev = hrows.shape[0], hrows.device
        scores = torch.empty((n, spec.n_rel), dtype=torch.float32, device=dev)
        _lib.check(self.lib.kge_rescal_rel_scores(_ptr(hrows), _ptr(trows), _ptr(spec.rel0), spec.dim, n,
                                                  spec.n_rel, _ptr(scores), _stream(dev)), "kge_rescal_rel_scores")
        self.launches += 1
        return scores

    def rank_dense(self, scores, true_idx, filt, raw_count, filt_sub, true_score=None, true_score_in=None):
        """get_rank + filter_scores on a dense (n, n_cand) matrix, counters added into."""
        n, n_c = scores.shape
        if filt is not None and filt[1].numel() == 0:
            filt = None      # nothing to discount (an empty tensor has no device pointer)
        offs, ids = filt if filt is not None else (None, None)
        _lib.check(self.lib.kge_rank_dense(_ptr(scores), n, n_c, _ptr(true_idx), _ptr(true_score_in), _ptr(offs),
                                           _ptr(ids), _ptr(raw_count), _ptr(filt_sub), _ptr(true_score),
                                           _stream(scores.device)), "kge_rank_dense")
        self.launches += 1

    def topk_dense(self, scores, k, mask=None):
        """(pred, scores) of the k best columns per row of a dense matrix (kge_topk_dense)."""
        n, n_c = scores.shape
        dev = scores.device
        pred = torch.empty((n, k), dtype=torch.int64, device=dev)
        vals = torch.empty((n, k), dtype=torch.float32, device=dev)
        ws_bytes = self.lib.kge_topk_dense_workspace_bytes(n, n_c, k)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        if mask is not None and mask[1].numel() == 0:
            mask = None
        offs, ids = mask if mask is not None else (None, None)
        _lib.check(self.lib.kge_topk_dense(_ptr(scores), n, n_c, k, _ptr(offs), _ptr(ids), _ptr(pred), _ptr(vals),
                                           _ptr(ws), ws_bytes, _stream(dev)), "kge_topk_dense")
        self.launches += 3
        return pred, vals

    def finalize(self, raw_count, filt_sub):
        n = raw_count.shape[0]
        ranks = torch.empty(n, dtype=torch.int64, device=raw_count.device)
        filt = torch.empty(n, dtype=torch.int64, device=raw_count.device)
        _lib.check(self.lib.kge_finalize_ranks(_ptr(raw_count), _ptr(filt_sub), n, _ptr(ranks),
                                               _ptr(filt), _stream(raw_count.device)),
                   "kge_finalize_ranks")
        self.launches += 1
        return ranks, filt


_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = CudaEngine()
    return _default_engine


class EntityShard:
    """Range partition of the entity table over the ranks of a process group
    (SURVEY.md section 8e): rank g holds rows [g*ceil(nE/G), (g+1)*ceil(nE/G))."""

    def __init__(self, n_ent, rank=0, world=1, group=None, local_storage=False):
        per = (n_ent + world - 1) // world
        self.n_ent, self.rank, self.world, self.group = n_ent, rank, world, group
        #: True when the model on this rank HOLDS only rows [lo, hi) (its row 0 is entity lo);
        #: False when every rank holds the full table and merely scans its own range
        self.local_storage = local_storage
        self.lo = min(n_ent, rank * per)
        self.hi = min(n_ent, (rank + 1) * per)

    @classmethod
    def from_group(cls, n_ent, group=None, local_storage=False):
        import torch.distributed as dist
        return cls(n_ent, dist.get_rank(group), dist.get_world_size(group), group, local_storage)

    def all_reduce_sum(self, t):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class QueryShard:
    """Contiguous split of n test triples over the ranks of a process group, for tables that fit
    one GPU (second form): the entity table is replicated, every rank ranks
    ITS triples against all entities -- independent units, no collective on the data path -- and
    the rank vectors are all-gathered at the end."""

    def __init__(self, n, rank=0, world=1, group=None):
        self.n, self.rank, self.world, self.group = int(n), rank, world, group
        self.per = (self.n + world - 1) // world
        self.lo = min(self.n, rank * self.per)
        self.hi = min(self.n, (rank + 1) * self.per)

    @classmethod
    def from_group(cls, n, group=None):
        import torch.distributed as dist
        return cls(n, dist.get_rank(group), dist.get_world_size(group), group)

    def slice(self, *tensors):
        """This rank's rows of each (n,) tensor."""
        return tuple(x[self.lo:self.hi].contiguous() for x in tensors)

    def csr(self, filt):
        """This rank's rows of a CSR over the n triples (offsets rebased)."""
        return None if filt is None else _csr_slice(filt, self.lo, self.hi, self.n)

    def all_reduce_sum(self, t):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, parts):
        """Per-rank result vectors (one entry per local triple, same dtype) -> full-length vectors on
        every rank, with ONE collective for all of them."""
        if self.world == 1:
            return list(parts)
        import torch.distributed as dist
        k = len(parts)
        mine = torch.zeros((k, self.per), dtype=parts[0].dtype, device=parts[0].device)
        for i, x in enumerate(parts):
            mine[i, :x.numel()] = x
        everyone = torch.empty((self.world, k, self.per), dtype=mine.dtype, device=mine.device)
        dist.all_gather([everyone[i] for i in range(self.world)], mine, group=self.group)
        # rank-major slices of length `per` concatenate to the original order of the triples
        full = everyone.permute(1, 0, 2).reshape(k, self.world * self.per)[:, :self.n]
        return [full[i] for i in range(k)]


def _csr_slice(filt, lo, hi, n):
    """CSR rows [lo, hi) with offsets (and row ids, when present) rebased to 0."""
    offs, ids = filt[0], filt[1]
    if lo == 0 and hi == n:
        return filt
    base = int(offs[lo].item())
    end = int(offs[hi].item())
    out = ((offs[lo:hi + 1] - base).contiguous(), ids[base:end].contiguous())
    if len(filt) > 2 and filt[2] is not None:
        out = out + ((filt[2][base:end] - lo).contiguous(),)
    return out


class LazyRanks:
    """Result of ``rank_link_prediction(..., sync=False)``: the four rank vectors are enqueued on
    the device but nothing has been synchronised yet.  ``overflow`` is a device scalar > 0 iff
    some bound-and-refine call ran out of room in its near-tie list (adversarial tables only);
    ``get()`` looks at it (one host sync) and, in that case, recomputes everything on the exact
    scalar scan.  Callers that copy the ranks to the host anyway pass the flag along in that copy
    (``get(flag_host=...)``)."""

    def __init__(self, ranks, overflow, redo):
        self.ranks, self.overflow, self._redo = ranks, overflow, redo

    def get(self, flag_host=None):
        if self.overflow is not None:
            flag = int(self.overflow.item()) if flag_host is None else int(flag_host)
            if flag > 0:
                import warnings
                warnings.warn("torchkge_b200: the near-tie list of %d bound-and-refine call(s) overflowed "
                              "(adversarial or degenerate embeddings?); recomputing the ranks on the exact "
                              "scalar scan" % flag, RuntimeWarning)
                self.ranks, self.overflow = self._redo(), None
        return self.ranks


def rank_link_prediction(spec, h_idx, t_idx, r_idx, filt_tail, filt_head, shard=None,
                         engine=None, chunk=DEFAULT_CHUNK, packed=None, exact=False, sync=True):
    """Rank every triple's true tail and head against all entities, raw and filtered.

    spec       ModelSpec holding either the full entity table or exactly this rank's shard
    h/t/r_idx  int64 device tensors (n,)
    filt_*     (offs int64 (n+1,), ids int64 (m,)) device CSR of entities to discount, None,
               or a zero-argument callable returning such a CSR -- callables are invoked only
               after every dense scan has been enqueued, so that building the CSR on the host
               overlaps with the scans on the device
    shard      EntityShard when the table is range-partitioned over a process group
    exact      True: scalar ATen-order scan only (no tensor-core / approximate bound-and-refine)
    sync       False: return a LazyRanks (no host synchronisation at all inside this call)
    Returns (rank_heads, rank_tails, filt_rank_heads, filt_rank_tails), int64 device tensors.
    """
    engine = engine or default_engine()
    n = h_idx.shape[0]
    if shard is not None and shard.world > 1:
        if (spec.ent_lo, spec.n_rows) != (shard.lo, shard.hi - shard.lo):
            spec = spec.narrowed(shard.lo, shard.hi)
            packed = None
    mark = getattr(engine, "mark", lambda label: None)
    mark("step begin")
    dev = spec.ent0.device
    with _device_guard(dev):
        tc_packed = engine.pack_tc(spec) if (hasattr(engine, "pack_tc") and not exact) else None
        mark("pack_tc")
        if packed is None and tc_packed is None:
            packed = engine.pack(spec)   # scalar-scan layout: only when there is no tensor-core path
        mark("pack")
        # RotatE has no tensor-core form; its bound-and-refine runs on the fp32 pipes (KGE_FLAG_APPROX_SCAN)
        refine = (not exact and tc_packed is None and spec.code == _lib.ROTATE
                  and getattr(engine, "tensor_core", False))
        if tc_packed is not None or refine:
            # the near-tie list of one call holds n * n_rows / 128 pairs at most 2^28 (2 GB): keep the
            # facts per call small enough for tables of many millions of rows, so that running out of
            # room stays the exception it is meant to be
            chunk = min(chunk, max(1024, ((1 << 35) // max(spec.n_rows, 1)) // 128 * 128))
        # counters raw_t, sub_t, raw_h, sub_h and, behind them, one slot for the overflow flag so
        # that a sharded run needs a single all-reduce
        buf = torch.zeros(4 * n + 1, dtype=torch.int32, device=dev)
        counters = buf[:4 * n].view(4, n)
        n_calls = 2 * ((n + chunk - 1) // chunk)
        stats_all = (torch.zeros((max(n_calls, 1), 2), dtype=torch.int64, device=dev)
                     if (tc_packed is not None or refine) else None)
        pending = []
        call = 0
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            h, t, r = h_idx[lo:hi], t_idx[lo:hi], r_idx[lo:hi]
            hrows = engine.gather_rows(spec, h)
            trows = engine.gather_rows(spec, t)
            if shard is not None and shard.world > 1:
                # every row is owned by exactly one rank; the others contribute zeros
                shard.all_reduce_sum(hrows)
                shard.all_reduce_sum(trows)
            for side, true_idx, which, raw, sub in ((_lib.SIDE_TAIL, t, 0, counters[0], counters[1]),
                                                    (_lib.SIDE_HEAD, h, 1, counters[2], counters[3])):
                if tc_packed is not None:
                    handle = engine.rank_side(spec, packed, side, hrows, trows, r, true_idx, None,
                                              raw[lo:hi], sub[lo:hi], tc_packed=tc_packed,
                                              stats=stats_all[call])
                elif refine:
                    handle = engine.rank_side(spec, packed, side, hrows, trows, r, true_idx, None,
                                              raw[lo:hi], sub[lo:hi], approx=True, stats=stats_all[call])
                else:
                    handle = engine.rank_side(spec, packed, side, hrows, trows, r, true_idx, None,
                                              raw[lo:hi], sub[lo:hi])
                call += 1
                pending.append((handle, which, lo, hi, sub))
                mark("rank_side %d enqueued" % side)
        filts = [filt_tail, filt_head]
        for k in (0, 1):
            if callable(filts[k]):
                filts[k] = filts[k]()
        # chunk boundaries of the CSRs: one device -> host read for all of them (none for a single chunk)
        cuts = [None, None]
        if n > chunk:
            starts = torch.tensor(list(range(0, n, chunk)) + [n], device=dev)
            for k in (0, 1):
                if filts[k] is not None:
                    cuts[k] = dict(zip(starts.tolist(), filts[k][0][starts].tolist()))
        for handle, which, lo, hi, sub in pending:
            if filts[which] is not None:
                offs, ids = filts[which][0], filts[which][1]
                qid = filts[which][2] if len(filts[which]) > 2 else None
                if cuts[which] is None:
                    part = filts[which]
                else:
                    a, b = cuts[which][lo], cuts[which][hi]
                    part = ((offs[lo:hi + 1] - a).contiguous(), ids[a:b].contiguous())
                    if qid is not None:
                        part = part + ((qid[a:b] - lo).contiguous(),)
                engine.filter_side(handle, part, sub[lo:hi])
        mark("filters enqueued")
        del pending
        overflow = None
        if stats_all is not None:
            # the near-tie list of a bound-and-refine call is bounded; a call that ran out of room
            # (never seen on real or synthetic embeddings, possible on adversarial ones) reports
            # found = capacity + 1.  The flag travels with the counters; nobody waits for it here.
            buf[4 * n:] = (stats_all[:, 0] > stats_all[:, 1]).sum().to(torch.int32)
            overflow = buf[4 * n]
        if shard is not None and shard.world > 1:
            shard.all_reduce_sum(buf)  # the single collective on the rank counters (+ flag)
        mark("collective")
        rank_t, filt_t = engine.finalize(counters[0], counters[1])
        rank_h, filt_h = engine.finalize(counters[2], counters[3])
        mark("finalize")
    result = (rank_h, rank_t, filt_h, filt_t)

    def redo():
        return rank_link_prediction(spec, h_idx, t_idx, r_idx, filts[0], filts[1], shard=shard,
                                    engine=engine, chunk=chunk, exact=True)

    lazy = LazyRanks(result, overflow, redo)
    return lazy.get() if sync else lazy


def relation_spec(spec):
    """The model seen from relation prediction: the candidate table is the RELATION table
    (``inference_prepare_candidates(..., entities=False)``: translation.py:118-121,
    bilinear.py:263-265, 551-554)."""
    cand2 = None
    if spec.code in (_lib.TRANSE_L1, _lib.TRANSE_L2, _lib.DISTMULT):
        cand0, cand1 = spec.rel0, None
    elif spec.code == _lib.COMPLEX:
        cand0, cand1 = spec.rel0, spec.rel1
    elif spec.code == _lib.ANALOGY:
        cand0, cand1, cand2 = spec.rel0, spec.rel1, spec.rel2
    else:
        raise NotImplementedError(
            "%s has no scan-based relation-prediction path (TransE L1/L2, DistMult, ComplEx, Analogy have; "
            "RESCAL goes through the dense kge_rescal_rel_scores)" % _lib.MODEL_NAMES.get(spec.code, spec.code))
    return ModelSpec(spec.code, spec.dim, spec.n_rel, spec.n_rel, cand0, cand1, None, None, ent2=cand2)


def rank_relation_prediction(spec, h_idx, t_idx, r_idx, filt, directed=True, engine=None,
                             chunk=DEFAULT_CHUNK):
    """Rank every fact's true relation against all relations (RelationPredictionEvaluator,
    torchkge/evaluation.py:64-112).

    filt       device CSR (offs, ids) of the relations to discount per fact (dict_of_rels[(h, t)]
               minus the true one), or None
    directed   False: the scores of (t, ?, h) are ranked together with those of (h, ?, t), against
               the directed true score (evaluation.py:99-107)
    Returns (rank_true_rels, filt_rank_true_rels), int64 device tensors.
    """
    engine = engine or default_engine()
    n = h_idx.shape[0]
    dev = spec.ent0.device
    counters = torch.zeros((2, n), dtype=torch.int32, device=dev)
    if spec.code == _lib.RESCAL:
        # the candidates are the relation MATRICES: per-fact vectors h^T M_c, a dense (n, n_rel)
        # score matrix (bilinear.py:115-121) ranked by kge_rank_dense
        for lo in range(0, n, min(chunk, 4096)):
            hi = min(n, lo + min(chunk, 4096))
            h, t, r = h_idx[lo:hi], t_idx[lo:hi], r_idx[lo:hi].contiguous()
            hrows = engine.gather_rows(spec, h).view(hi - lo, spec.dim)
            trows = engine.gather_rows(spec, t).view(hi - lo, spec.dim)
            f = None if filt is None else _csr_slice(filt, lo, hi, n)
            s_true = torch.empty(hi - lo, dtype=torch.float32, device=dev)
            scores = engine.rescal_rel_scores(spec, hrows, trows)
            engine.rank_dense(scores, r, f, counters[0][lo:hi], counters[1][lo:hi], true_score=s_true)
            if not directed:
                scores2 = engine.rescal_rel_scores(spec, trows, hrows)
                engine.rank_dense(scores2, r, f, counters[0][lo:hi], counters[1][lo:hi], true_score_in=s_true)
        return engine.finalize(counters[0], counters[1])
    rspec = relation_spec(spec)
    packed = engine.pack(rspec)
    keep = []
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        h, t, r = h_idx[lo:hi], t_idx[lo:hi], r_idx[lo:hi]
        hrows, trows = engine.gather_rows(spec, h), engine.gather_rows(spec, t)
        rrows = engine.gather_rows(rspec, r)
        f = None if filt is None else _csr_slice(filt, lo, hi, n)
        s_true = torch.empty(hi - lo, dtype=torch.float32, device=dev)
        keep.append(engine.rank_side(rspec, packed, _lib.SIDE_REL, hrows, trows, None, r, f,
                                     counters[0][lo:hi], counters[1][lo:hi], true_score=s_true,
                                     true_rows=rrows))
        if not directed:
            keep.append(engine.rank_side(rspec, packed, _lib.SIDE_REL, trows, hrows, None, r, f,
                                         counters[0][lo:hi], counters[1][lo:hi], true_rows=rrows,
                                         true_score_in=s_true))
        keep.append((s_true, f))
    ranks, filt_ranks = engine.finalize(counters[0], counters[1])
    del keep
    return ranks, filt_ranks
