"""Perf probe: time the rank scan of one side on random tables (device-resident), per model; QP_MODELS selects (l2,l1,dm,cx,rot,rot1k), QP_TC=0 forces the exact scalar scan."""
import sys, time
import torch
sys.path.insert(0, '.')
from torchkge_b200 import _lib
from torchkge_b200.engine import ModelSpec, CudaEngine

def run(code, d, n_ent, n_q, n_rel=1000, reps=3):
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    planes = 2 if code in (_lib.COMPLEX, _lib.ROTATE) else 1
    ent0 = torch.randn(n_ent, d, device=dev, generator=g) * 0.07
    ent1 = torch.randn(n_ent, d, device=dev, generator=g) * 0.07 if planes == 2 else None
    rel0 = torch.randn(n_rel, d, device=dev, generator=g) * 0.07
    rel1 = torch.randn(n_rel, d, device=dev, generator=g) * 0.07 if planes == 2 else None
    spec = ModelSpec(code, d, n_ent, n_rel, ent0, ent1, rel0, rel1)
    import os
    eng = CudaEngine(tensor_core=os.environ.get('QP_TC', '1') == '1')
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); packed = eng.pack(spec); tcp = eng.pack_tc(spec); t1.record(); torch.cuda.synchronize()
    pack_ms = t0.elapsed_time(t1)
    h = torch.randint(0, n_ent, (n_q,), device=dev, generator=g)
    t = torch.randint(0, n_ent, (n_q,), device=dev, generator=g)
    r = torch.randint(0, n_rel, (n_q,), device=dev, generator=g)
    hrows = eng.gather_rows(spec, h); trows = eng.gather_rows(spec, t)
    out = {}
    pairs = n_q * n_ent
    for side in (0, 1):
        raw = torch.zeros(n_q, dtype=torch.int32, device=dev); sub = torch.zeros_like(raw)
        best = 1e9
        for _ in range(reps):
            raw.zero_()
            t0.record()
            ws = eng.rank_side(spec, packed, side, hrows, trows, r, t if side == 0 else h, None, raw, sub,
                               tc_packed=tcp, approx=eng.tensor_core)
            t1.record(); torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        out[side] = best
        assert int(raw.min()) >= 1, 'true entity must count itself'
        if eng.tc_stats: out['amb%d' % side] = int(eng.tc_stats[-1][0]) / pairs
    pairs = n_q * n_ent
    print('%-10s d=%4d nE=%8d nq=%6d pack %.2f ms | tail %.2f ms (%.2f Tpair-dim/s) head %.2f ms (%.2f) | mean raw rank %.0f | near-tie frac %s' % (
        _lib.MODEL_NAMES[code], d, n_ent, n_q, pack_ms, out[0], pairs * d / out[0] / 1e9, out[1], pairs * d / out[1] / 1e9, raw.float().mean().item(), (out.get('amb0'), out.get('amb1'))), flush=True)

if __name__ == '__main__':
    n_ent = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    import os
    sel = os.environ.get("QP_MODELS", "l2,l1,dm,cx,rot").split(",")
    table = {"l2": (_lib.TRANSE_L2, 200), "l1": (_lib.TRANSE_L1, 200), "dm": (_lib.DISTMULT, 200),
             "cx": (_lib.COMPLEX, 400), "rot": (_lib.ROTATE, 200), "rot1k": (_lib.ROTATE, 1000)}
    for k in sel:
        run(table[k][0], table[k][1], n_ent, n_q)
