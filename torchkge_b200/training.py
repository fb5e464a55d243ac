"""Training-side entry points: differentiable per-triple scoring (``Model.scoring_function``),
``MarginLoss`` and the fused sample + score + hinge step, as autograd Functions over the CUDA
kernels of csrc/train.cu.  Gradients are dense tables (what ``nn.Embedding`` yields in the
reference), accumulated with atomics.
"""
import ctypes

import torch

from . import _lib
from .engine import ModelSpec, _ptr, _stream


def _param_tensors(model, code):
    """Parameters in ModelSpec order (ent0, ent1, rel0, rel1); RotatE's relation planes are
    differentiable functions (cos, sin) of its phase parameter."""
    if code in (_lib.TRANSE_L1, _lib.TRANSE_L2, _lib.DISTMULT, _lib.TORUSE_L1, _lib.TORUSE_L2):
        return model.ent_emb.weight, None, model.rel_emb.weight, None
    if code == _lib.RESCAL:
        return model.ent_emb.weight, None, model.rel_mat.weight, None
    if code == _lib.COMPLEX:
        return (model.re_ent_emb.weight, model.im_ent_emb.weight, model.re_rel_emb.weight,
                model.im_rel_emb.weight)
    if code == _lib.ROTATE:
        ph = model.rel_emb.weight
        return model.re_ent_emb.weight, model.im_ent_emb.weight, torch.cos(ph), torch.sin(ph)
    if code == _lib.ANALOGY:
        # three planes per table: one stacked (3, n, dim) tensor in the ent0 / rel0 slot (autograd
        # splits its gradient back onto the three embeddings); see _plane_ptrs
        return (torch.stack([model.sc_ent_emb.weight, model.re_ent_emb.weight, model.im_ent_emb.weight]), None,
                torch.stack([model.sc_rel_emb.weight, model.re_rel_emb.weight, model.im_rel_emb.weight]), None)
    raise NotImplementedError(code)


def _kernel_dim(model, code):
    """Width of one plane: emb_dim, or Analogy's scalar_dim (= complex_dim on this path)."""
    return model.scalar_dim if code == _lib.ANALOGY else model.emb_dim


def _plane_ptrs(x0, x1):
    """(plane 0, plane 1) pointers of a table: a stacked (3, n, dim) tensor stands for three equally
    spaced planes, of which the C ABI takes the first two (include/kge_b200.h, "three-plane tables")."""
    if x0 is not None and x0.dim() == 3:
        return _ptr(x0[0]), _ptr(x0[1])
    return _ptr(x0), _ptr(x1)


def _tables(code, dim, tensors):
    tb = _lib.Tables()
    tb.model, tb.dim = code, dim
    tb.ent0, tb.ent1 = _plane_ptrs(tensors[0], tensors[1])
    tb.rel0, tb.rel1 = _plane_ptrs(tensors[2], tensors[3])
    return tb


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.KgeLibraryError("training kernels need CUDA tensors (got %s); there is no "
                                       "CPU fallback" % t.device)


def _idx(t, dev):
    return t.to(device=dev, dtype=torch.int64).contiguous()


def _zero_grads(tensors):
    gs = [None if x is None else torch.zeros_like(x, dtype=torch.float32) for x in tensors]
    g = _lib.Grads()
    g.ent0, g.ent1 = _plane_ptrs(gs[0], gs[1])
    g.rel0, g.rel1 = _plane_ptrs(gs[2], gs[3])
    return gs, g


class _ScoreTriples(torch.autograd.Function):
    @staticmethod
    def forward(ctx, code, dim, h, t, r, ent0, ent1, rel0, rel1):
        tensors = [None if x is None else x.detach().contiguous() for x in (ent0, ent1, rel0, rel1)]
        _check_cuda(tensors[0], h, t, r)
        dev = tensors[0].device
        h, t, r = _idx(h, dev), _idx(t, dev), _idx(r, dev)
        n = h.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=dev)
        tb = _tables(code, dim, tensors)
        _lib.check(_lib.load().kge_score_triples_fwd(ctypes.byref(tb), _ptr(h), _ptr(t), _ptr(r), n,
                                                     _ptr(out), _stream(dev)), "kge_score_triples_fwd")
        ctx.code, ctx.dim = code, dim
        ctx.save_for_backward(h, t, r, *[x for x in tensors if x is not None])
        ctx.present = [x is not None for x in tensors]
        return out

    @staticmethod
    def backward(ctx, gout):
        saved = list(ctx.saved_tensors)
        h, t, r = saved[:3]
        it = iter(saved[3:])
        tensors = [next(it) if p else None for p in ctx.present]
        dev = h.device
        gout = gout.contiguous().float()
        tb = _tables(ctx.code, ctx.dim, tensors)
        gs, g = _zero_grads(tensors)
        _lib.check(_lib.load().kge_score_triples_bwd(ctypes.byref(tb), ctypes.byref(g), _ptr(h), _ptr(t),
                                                     _ptr(r), h.shape[0], _ptr(gout), _stream(dev)),
                   "kge_score_triples_bwd")
        return (None, None, None, None, None, *gs)


def _training_code(model):
    """Kernel selector of the training-side kernels.  TorusE: the torus dissimilarities have their own
    per-triple kernels (no row normalisation, fractional parts taken on the fly, translation.py:706-720);
    its plain-'L1' variant is not on the CUDA path (the TransE-L1 kernels L2-normalise the entity rows)."""
    if type(model).__name__ == "TorusEModel":
        dname = getattr(model.dissimilarity, "__name__", str(model.dissimilarity))
        code = {"l1_torus_dissimilarity": _lib.TORUSE_L1, "l2_torus_dissimilarity": _lib.TORUSE_L2}.get(dname)
        if code is None:
            raise NotImplementedError("TorusE with dissimilarity %s has no training kernel" % dname)
        return code
    if type(model).__name__ == "AnalogyModel":
        if model.scalar_dim != model.complex_dim:
            raise NotImplementedError("the Analogy training kernels need scalar_dim == complex_dim")
        return _lib.ANALOGY
    return ModelSpec.from_model(model).code


def score_triples(model, h_idx, t_idx, r_idx):
    """``model.scoring_function(h_idx, t_idx, r_idx)`` -> (n,) float scores, differentiable with
    respect to the model's embedding tables."""
    code = _training_code(model)
    ent0, ent1, rel0, rel1 = _param_tensors(model, code)
    return _ScoreTriples.apply(code, _kernel_dim(model, code), h_idx, t_idx, r_idx, ent0, ent1, rel0, rel1)


class _MarginLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, neg, margin):
        _check_cuda(pos, neg)
        pos, neg = pos.detach().contiguous().float(), neg.detach().contiguous().float()
        if pos.shape != neg.shape:
            raise ValueError("positive and negative score tensors must have the same shape")
        loss = torch.zeros((), dtype=torch.float32, device=pos.device)
        _lib.check(_lib.load().kge_margin_loss_fwd(_ptr(pos), _ptr(neg), pos.numel(), float(margin),
                                                   _ptr(loss), _stream(pos.device)), "kge_margin_loss_fwd")
        ctx.margin = float(margin)
        ctx.save_for_backward(pos, neg)
        return loss

    @staticmethod
    def backward(ctx, gl):
        pos, neg = ctx.saved_tensors
        gl = gl.contiguous().float()
        gp, gn = torch.empty_like(pos), torch.empty_like(neg)
        _lib.check(_lib.load().kge_margin_loss_bwd(_ptr(pos), _ptr(neg), pos.numel(), ctx.margin,
                                                   _ptr(gl), _ptr(gp), _ptr(gn), _stream(pos.device)),
                   "kge_margin_loss_bwd")
        return gp, gn, None


def margin_loss(pos, neg, margin):
    return _MarginLoss.apply(pos, neg, margin)


class _PairLoss(torch.autograd.Function):
    """LogisticLoss / BinaryCrossEntropyLoss (kge_pair_loss_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, pos, neg, kind):
        _check_cuda(pos, neg)
        pos, neg = pos.detach().contiguous().float(), neg.detach().contiguous().float()
        if pos.shape != neg.shape:
            raise ValueError("positive and negative score tensors must have the same shape")
        loss = torch.zeros((), dtype=torch.float32, device=pos.device)
        _lib.check(_lib.load().kge_pair_loss_fwd(kind, _ptr(pos), _ptr(neg), pos.numel(), _ptr(loss),
                                                 _stream(pos.device)), "kge_pair_loss_fwd")
        ctx.kind = kind
        ctx.save_for_backward(pos, neg)
        return loss

    @staticmethod
    def backward(ctx, gl):
        pos, neg = ctx.saved_tensors
        gl = gl.contiguous().float()
        gp, gn = torch.empty_like(pos), torch.empty_like(neg)
        _lib.check(_lib.load().kge_pair_loss_bwd(ctx.kind, _ptr(pos), _ptr(neg), pos.numel(), _ptr(gl),
                                                 _ptr(gp), _ptr(gn), _stream(pos.device)),
                   "kge_pair_loss_bwd")
        return gp, gn, None


def pair_loss(pos, neg, kind):
    return _PairLoss.apply(pos, neg, kind)


class _MarginStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, code, dim, n_ent, margin, n_neg, h, t, r, nh, nt, probs, seed, offset,
                ent0, ent1, rel0, rel1):
        tensors = [None if x is None else x.detach().contiguous() for x in (ent0, ent1, rel0, rel1)]
        _check_cuda(tensors[0], h, t, r)
        dev = tensors[0].device
        h, t, r = _idx(h, dev), _idx(t, dev), _idx(r, dev)
        if nh is not None:
            nh, nt = _idx(nh, dev), _idx(nt, dev)
        if probs is not None:
            probs = probs.to(device=dev, dtype=torch.float32).contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        a = _MarginStep._args(code, dim, n_ent, margin, n_neg, h, t, r, nh, nt, probs, seed, offset,
                              tensors, loss, dev)
        _lib.check(_lib.load().kge_margin_step_fwd(ctypes.byref(a)), "kge_margin_step_fwd")
        ctx.meta = (code, dim, n_ent, margin, n_neg, seed, offset)
        ctx.present = [x is not None for x in tensors]
        ctx.has_neg, ctx.has_probs = nh is not None, probs is not None
        extra = ([nh, nt] if nh is not None else []) + ([probs] if probs is not None else [])
        ctx.save_for_backward(h, t, r, *extra, *[x for x in tensors if x is not None])
        return loss

    @staticmethod
    def _args(code, dim, n_ent, margin, n_neg, h, t, r, nh, nt, probs, seed, offset, tensors, loss, dev):
        a = _lib.MarginStepArgs()
        a.tb = _tables(code, dim, tensors)
        a.n_neg, a.margin, a.b, a.n_ent = n_neg, float(margin), h.shape[0], n_ent
        a.h, a.t, a.r, a.nh, a.nt, a.bern_probs = (_ptr(x) for x in (h, t, r, nh, nt, probs))
        a.seed, a.offset = int(seed), int(offset)
        a.loss, a.stream = _ptr(loss), _stream(dev)
        return a

    @staticmethod
    def backward(ctx, gl):
        code, dim, n_ent, margin, n_neg, seed, offset = ctx.meta
        saved = list(ctx.saved_tensors)
        h, t, r = saved[:3]
        k = 3
        nh = nt = probs = None
        if ctx.has_neg:
            nh, nt = saved[k], saved[k + 1]
            k += 2
        if ctx.has_probs:
            probs = saved[k]
            k += 1
        it = iter(saved[k:])
        tensors = [next(it) if p else None for p in ctx.present]
        dev = h.device
        gl = gl.contiguous().float()
        dummy = torch.zeros((), dtype=torch.float32, device=dev)
        a = _MarginStep._args(code, dim, n_ent, margin, n_neg, h, t, r, nh, nt, probs, seed, offset,
                              tensors, dummy, dev)
        gs, g = _zero_grads(tensors)
        _lib.check(_lib.load().kge_margin_step_bwd(ctypes.byref(a), ctypes.byref(g), _ptr(gl)),
                   "kge_margin_step_bwd")
        return (None,) * 13 + tuple(gs)


def fused_margin_step(model, heads, tails, relations, margin, n_neg=1, negatives=None,
                      bern_probs=None, seed=0, offset=0):
    """Loss of one training step, fused: Bernoulli corruption (or the given ``negatives =
    (neg_heads, neg_tails)``), ``model(h, t, r, nh, nt)`` and ``MarginLoss(margin)`` in a
    single kernel, differentiable with respect to the embedding tables.

    Equivalent to the tutorial loop body (docs/tutorials/transe.rst:47-57)
        nh, nt = sampler.corrupt_batch(h, t, r); pos, neg = model(h, t, r, nh, nt)
        loss = criterion(pos, neg)
    without materialising nh, nt, pos, neg.
    """
    spec_code = _training_code(model)
    ent0, ent1, rel0, rel1 = _param_tensors(model, spec_code)
    nh = nt = None
    if negatives is not None:
        nh, nt = negatives
        n_neg = int(nh.shape[0] // heads.shape[0])
    elif bern_probs is None:
        raise ValueError("either negatives or bern_probs must be given")
    return _MarginStep.apply(spec_code, _kernel_dim(model, spec_code), model.n_ent, margin, n_neg, heads, tails,
                             relations, nh, nt, bern_probs, seed, offset, ent0, ent1, rel0, rel1)
