#!/usr/bin/env python
"""bench.py -- filtered link-prediction throughput (triples/s) on synthetic KGs.

    python bench.py --gpus N --steps K --warmup W [--workload c2] [--impl reference]

A "step" is ONE full filtered link-prediction evaluation of the workload's test set
(raw + filtered ranks of the true head and of the true tail of every test triple against all
entities) -- the metric BASELINE.json names.  Default workload: c2 = TransE-L2 d=200,
|E|=1M, |R|=1k, 20,466 test triples (the single-GPU configuration the metric is quoted on).

Our arm (default) prints one JSON line with
  value      whole-job triples/s, inputs already on the device (table, indices, filter CSR)
  e2e        the same through the public API (LinkPredictionEvaluator.evaluate) from HOST
             index tensors / filter dictionaries to rank vectors on the host
  roofline   the dense scan kernel: algorithmic bytes per launch / CUDA-event duration
  cpu_baseline  the CPU oracle (a PyTorch-CPU restatement of torchkge's path) timed on a
             bounded sample of the same test set on this box's host cores, and a parity
             check of the GPU ranks on that sample
With N > 1 (torchrun) the entity table is range-partitioned over the ranks (each rank holds
and scans |E|/N rows of every query) and the rank counters are summed by one NCCL
all-reduce: the total work is fixed, so "scaling" is "strong".

--impl reference times the reference's own CPU implementation of the path; the reference is
pure Python over ATen and cannot be installed on the GPU box, so this arm runs the oracle
port (oracle/kge_oracle.py, pinned against the unmodified reference by tests/golden) on a
bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-sample", type=int, default=24, help="test triples in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--split", default="auto", choices=["auto", "queries", "entities"],
                    help="multi-GPU decomposition: shard the test triples (replicated table) or the entity table")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the warm-up and timed steps run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1]); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- helpers
def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def row_bytes(code, dim):
    from torchkge_b200 import _lib
    return (8 if code in (_lib.COMPLEX, _lib.ROTATE) else 4) * dim


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 1400.0, "fallback (B200_PROFILING.md sustained)"


class HostKG:
    """What LinkPredictionEvaluator reads from a knowledge graph."""

    def __init__(self, h, t, r, n_ent, n_rel, dh, dt):
        self.head_idx, self.tail_idx, self.relations = h, t, r
        self.n_ent, self.n_rel, self.n_facts = n_ent, n_rel, h.shape[0]
        self.dict_of_heads, self.dict_of_tails = dh, dt

    def __len__(self):
        return self.n_facts


class TableModel(torch.nn.Module):
    """A model object of the reference's class name wrapping pre-generated tables, so that
    the public evaluator can be driven with synthetic weights of any size."""

    def __init__(self, cls_name, diss, dim, n_ent, n_rel, tabs):
        super().__init__()
        self.__class__ = type(cls_name, (TableModel,), {})
        self.emb_dim, self.n_ent, self.n_rel = dim, n_ent, n_rel
        P = torch.nn.Parameter
        mk = lambda w: torch.nn.Embedding.from_pretrained(w, freeze=True)  # noqa: E731
        if cls_name == "TransEModel":
            from torchkge_b200.models import l1_dissimilarity, l2_dissimilarity
            self.dissimilarity = l1_dissimilarity if diss == "L1" else l2_dissimilarity
            self.ent_emb, self.rel_emb = mk(tabs["ent0"]), mk(tabs["rel0"])
        elif cls_name == "DistMultModel":
            self.ent_emb, self.rel_emb = mk(tabs["ent0"]), mk(tabs["rel0"])
        elif cls_name == "RESCALModel":
            self.ent_emb, self.rel_mat = mk(tabs["ent0"]), mk(tabs["rel0"])
        elif cls_name == "ComplExModel":
            self.re_ent_emb, self.im_ent_emb = mk(tabs["ent0"]), mk(tabs["ent1"])
            self.re_rel_emb, self.im_rel_emb = mk(tabs["rel0"]), mk(tabs["rel1"])
        elif cls_name == "RotatEModel":
            self.re_ent_emb, self.im_ent_emb = mk(tabs["ent0"]), mk(tabs["ent1"])
            self._planes = (tabs["rel0"], tabs["rel1"])
        del P

    def relation_planes(self):
        return self._planes


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """CPU oracle port on a bounded sample per step; rank 0 only."""
    if rank != 0:
        return
    from oracle import kge_oracle as oracle
    from torchkge_b200 import synthetic as S
    wl = S.WORKLOADS[args.workload]
    code = S.MODEL_CODE[(wl["model"], wl["diss"])]
    kind = S.ORACLE_KIND[code]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    tabs = S.make_tables(code, wl["dim"], wl["n_ent"], wl["n_rel"], 0, wl["n_ent"], args.seed, dev)
    graph = S.make_graph(wl["n_ent"], wl["n_rel"], wl["n_facts"], wl["n_test"], args.seed, dev)
    per_step = 8 if wl["n_ent"] >= 100000 else 1024
    total = min(per_step * (args.steps + args.warmup), graph["test_h"].numel())
    per_step = max(1, total // (args.steps + args.warmup))
    dh, dt = S.filters_as_dicts(graph, wl["n_ent"], wl["n_rel"], limit=total)
    P = oracle_params_from_tables(kind, {k: (v.cpu() if v is not None else None) for k, v in tabs.items()})
    th, tt, tr = (graph[k][:total].cpu() for k in ("test_h", "test_t", "test_r"))
    b_size = 4 if wl["n_ent"] >= 100000 else 256

    def step(i):
        lo, hi = i * per_step, (i + 1) * per_step
        oracle.link_prediction(kind, P, th[lo:hi], tt[lo:hi], tr[lo:hi], dh, dt, b_size)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    el = time.perf_counter() - t0
    value = args.steps * per_step / el
    sample = "%d test triples per step (b_size %d) of workload %s" % (per_step, b_size, args.workload)
    out = {
        "impl": "reference", "metric": "filtered link-prediction triples/sec (full hits@k/MRR eval)",
        "value": value, "unit": "triples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * el / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl, world, split_mode(args, wl, world)),
        "cpu_baseline": {"value": value, "unit": "triples/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def oracle_params_from_tables(kind, tabs):
    if kind in ("transe_l1", "transe_l2", "distmult"):
        return {"ent": tabs["ent0"], "rel": tabs["rel0"]}
    if kind == "rescal":
        return {"ent": tabs["ent0"], "rel_mat": tabs["rel0"]}
    return {"re_ent": tabs["ent0"], "im_ent": tabs["ent1"], "re_rel": tabs["rel0"], "im_rel": tabs["rel1"]}


def split_mode(args, wl, world):
    """single | queries (replicated table, test triples sharded) | entities (table range-partitioned)"""
    if world == 1:
        return "single"
    if args.split != "auto":
        return args.split
    table_bytes = wl["n_ent"] * wl["dim"] * (8 if wl["model"] in ("ComplEx", "RotatE") else 4)
    return "entities" if table_bytes > 24e9 else "queries"


def workload_config(name, wl, world, mode="single"):
    return {"workload": "%s: %s%s dim=%d |E|=%d |R|=%d, %d test triples, full filtered LP (head+tail)" % (
                name, wl["model"], ("-" + wl["diss"]) if wl["diss"] else "", wl["dim"], wl["n_ent"],
                wl["n_rel"], wl["n_test"]),
            "n_facts_requested": wl["n_facts"],
            "parallelism": {"single": "single GPU",
                            "queries": "test triples sharded x%d, table replicated, no data-path collective "
                                       "(rank vectors all-gathered)" % world,
                            "entities": "entity-range shards x%d, 1 all-reduce of rank counters" % world}[mode],
            "l2_policy": l2_policy(wl)}


def l2_policy(wl):
    mb = wl["n_ent"] * wl["dim"] * (8 if wl["model"] in ("ComplEx", "RotatE") else 4) / 1e6
    if mb > 126:
        return "inputs larger than L2 (table %.0f MB >> 126 MB)" % mb
    return ("table %.1f MB is L2-resident: plumbing / parity configuration, no L2 flush between steps, "
            "HBM fractions are meaningless here" % mb)


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, local, world):
    import torch.distributed as dist
    from torchkge_b200 import _lib, synthetic as S
    from torchkge_b200.engine import (CudaEngine, EntityShard, ModelSpec, QueryShard,
                                      rank_link_prediction)
    import torchkge_b200.engine as engine_mod
    from torchkge_b200.evaluation import LinkPredictionEvaluator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    wl = S.WORKLOADS[args.workload]
    code = S.MODEL_CODE[(wl["model"], wl["diss"])]
    n_ent, n_rel, dim = wl["n_ent"], wl["n_rel"], wl["dim"]
    # How the job is split over ranks: a table that fits one GPU is
    # replicated and the TEST TRIPLES are sharded (independent units, no collective on the data
    # path, ranks all-gathered at the end); a table that does not (c4) is range-partitioned over
    # the ranks and the rank counters are summed by one all-reduce.
    mode = split_mode(args, wl, world)
    shard = EntityShard(n_ent, rank, world, None, local_storage=True) if mode == "entities" else None
    lo, hi = (shard.lo, shard.hi) if shard else (0, n_ent)

    tabs = S.make_tables(code, dim, n_ent, n_rel, lo, hi, args.seed, dev)
    graph = S.make_graph(n_ent, n_rel, wl["n_facts"], wl["n_test"], args.seed, dev)
    t0 = time.perf_counter()
    csr_t, csr_h = S.make_filters(graph, n_ent, n_rel)
    torch.cuda.synchronize()
    csr_build_s = time.perf_counter() - t0
    n_test = graph["test_h"].numel()
    # this rank's contiguous slice of the test set and of the filter CSRs (all of it unless the
    # test triples are what is sharded)
    qshard = QueryShard(n_test, rank, world) if mode == "queries" else QueryShard(n_test, 0, 1)
    q_lo, q_hi = qshard.lo, qshard.hi
    my_h, my_t, my_r = qshard.slice(graph["test_h"], graph["test_t"], graph["test_r"])
    my_csr_t, my_csr_h = qshard.csr(csr_t), qshard.csr(csr_h)
    spec = ModelSpec(code, dim, n_ent, n_rel, tabs["ent0"], tabs["ent1"], tabs["rel0"], tabs["rel1"],
                     ent_lo=lo)
    eng = CudaEngine()
    engine_mod._default_engine = eng  # the evaluator uses the same instance (launch counting)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return rank_link_prediction(spec, my_h, my_t, my_r, my_csr_t, my_csr_h, shard=shard, engine=eng)

    def gather_ranks(parts):
        """all ranks' slices -> full-length vectors (query-sharded mode only)"""
        return qshard.all_gather(parts) if mode == "queries" else parts

    # ---- device-resident timing -------------------------------------------------------
    # clocks / throttle reasons are sampled from the first warm-up step to the end of the timed
    # region (same load throughout; a timed region of a few hundred ms alone may fall between two
    # nvidia-smi samples)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        ranks_dev = device_step()
    barrier()
    _lib.scan_timing_enable(True)
    for kind in (0, 1, 2):
        _lib.scan_timing_read(kind)
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        ranks_dev = device_step()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    if eng.trace is not None and rank == 0:   # KGE_TRACE=1: where the steps spent their time
        for lab, host_ms, dev_ms_ in eng.trace_report():
            print("trace %-28s host %9.3f ms  device %9.3f ms" % (lab, host_ms, dev_ms_), file=sys.stderr)
    launches = eng.launches - launches0
    scan_n, scan_ms = _lib.scan_timing_read(0)
    tc_n, tc_ms = _lib.scan_timing_read(1)
    rc_n, rc_ms = _lib.scan_timing_read(2)
    _lib.scan_timing_enable(False)
    clock_rec = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = t_ms.item()
    value = args.steps * n_test / (dev_ms / 1000.0)

    # ---- end-to-end through the public API, host buffers in, host ranks out ----------------
    cls = wl["model"] + "Model"
    model = TableModel(cls, wl["diss"], dim, n_ent, n_rel, tabs)
    # host-side knowledge graph: test facts + the filter sets of the FULL graph as sorted arrays
    from torchkge_b200.data import KnowledgeGraph
    t0 = time.perf_counter()
    kg = KnowledgeGraph(my_h.cpu(), my_t.cpu(), my_r.cpu(), n_ent, n_rel,
                        filter_facts=(graph["heads"].cpu(), graph["tails"].cpu(), graph["rels"].cpu()))
    kg.head_idx, kg.tail_idx, kg.relations = (x.pin_memory() for x in (kg.head_idx, kg.tail_idx, kg.relations))
    filter_index_build_s = time.perf_counter() - t0
    evaluator = LinkPredictionEvaluator(model, kg, shard=shard)  # sharded: model holds its rows only
    for _ in range(min(args.warmup, 2)):
        evaluator.evaluate(b_size=256, verbose=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        evaluator.evaluate(b_size=256, verbose=False)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = t_e2e.item()
    e2e_value = args.steps * n_test / e2e_s
    same = all(torch.equal(a.cpu(), b) for a, b in zip(
        ranks_dev, (evaluator.rank_true_heads, evaluator.rank_true_tails,
                    evaluator.filt_rank_true_heads, evaluator.filt_rank_true_tails)))
    ranks_dev = gather_ranks(ranks_dev)

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peak, peak_src = measured_peaks()
    rows_here = hi - lo
    rb = row_bytes(code, dim)
    sm_mhz = (clock_rec or {}).get("sm_mhz") or 1965.0
    near_ties = None
    if tc_n > 0:
        # tensor-core bound-and-refine scan: bf16x3 split GEMM (3 bf16 MMAs per fp32 product)
        k_total = dim * (2 if code == _lib.COMPLEX else 1)
        ms_per_launch = tc_ms / tc_n
        n_my = q_hi - q_lo
        alg_flops = 2.0 * n_my * rows_here * k_total            # the fp32 contraction itself (this rank)
        tensor_peak = measured_tensor_peak()
        near_ties = sum(int(s_[0]) for s_ in eng.tc_stats[-2 * args.steps:]) / max(1, args.steps)
        roofline = {
            "kernel": "tc_scan_kernel (tcgen05 bf16x3 split GEMM + threshold epilogue)", "bound": "tensor",
            "achieved": alg_flops / (ms_per_launch / 1000.0) / 1e12, "peak": tensor_peak[0],
            "unit": "TFLOP/s", "frac": alg_flops / (ms_per_launch / 1000.0) / 1e12 / tensor_peak[0],
            "peak_source": tensor_peak[1],
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at c2 (20,466 x 1M), from the
            # committed capture profiles/r01_tc_scan_v2_ncu_summary.md; null for other shapes
            "traffic": (1.029e9 + 0.177e9) if (args.workload == "c2" and world == 1) else None,
            "traffic_unit": "bytes per launch (ncu --set full, profiles/r01_tc_scan_v2_ncu_summary.md)",
            "launches_timed": tc_n, "ms_per_launch": ms_per_launch,
            "scan_share_of_step": tc_ms / dev_ms if dev_ms > 0 else None,
            "recheck_ms_per_launch": rc_ms / max(1, rc_n),
            "recheck_share_of_step": rc_ms / dev_ms if dev_ms > 0 else None,
            "near_tie_pairs_per_step": near_ties,
            "near_tie_fraction": near_ties / (2.0 * max(1, n_my) * rows_here),
            "executed_bf16_tflops": 3 * alg_flops / (ms_per_launch / 1000.0) / 1e12,
            "executed_frac_of_peak": 3 * alg_flops / (ms_per_launch / 1000.0) / 1e12 / tensor_peak[0],
            "note": "algorithmic flops = 2 x queries x rows x K (fp32 contraction); the kernel executes "
                    "3 bf16 MMAs per product (hi*hi + lo*hi + hi*lo), i.e. 3x the algorithmic flops, and "
                    "runs power-capped (sw_power_cap) like the sustained cuBLAS measurement the peak "
                    "comes from; near-ties are re-scored exactly so ranks stay bit-identical",
        }
    else:
        # scalar fp32 scan.  algorithmic bytes per launch = queries x candidate rows x row_bytes
        # (2*nE*row_bytes per triple = nE*row_bytes per (triple, side) launch unit)
        alg_bytes_per_launch = float(q_hi - q_lo) * rows_here * rb
        scan_ms_per_launch = scan_ms / max(1, scan_n)
        achieved = alg_bytes_per_launch / (scan_ms_per_launch / 1000.0) / 1e9
        refine_rot = code == _lib.ROTATE and eng.tensor_core   # approximate-sqrt bound-and-refine scan
        ops_per_elem = {_lib.TRANSE_L1: 2.5, _lib.TRANSE_L2: 3.5, _lib.DISTMULT: 2.0, _lib.RESCAL: 2.0,
                        _lib.COMPLEX: 4.0, _lib.ROTATE: 5.0 if refine_rot else 16.0}[code]  # fp32 ops per (q,c,k)
        lane_ops = float(q_hi - q_lo) * rows_here * dim * ops_per_elem
        fp32_peak = 148 * 128 * sm_mhz * 1e6
        roofline = {
            "kernel": "scan_kernel (dense rank scan, fp32 pipes)", "bound": "hbm", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": None,
            "launches_timed": scan_n, "ms_per_launch": scan_ms_per_launch,
            "scan_share_of_step": scan_ms / dev_ms if dev_ms > 0 else None,
            "note": "algorithmic bytes = queries x rows x row_bytes per launch; every streamed candidate "
                    "tile is shared by 64 queries per CTA, so DRAM traffic is ~1/64 of this and the "
                    "kernel is fp32-issue bound (see fp32_issue)",
            "fp32_issue": {"achieved_tlaneops": lane_ops / (scan_ms_per_launch / 1000.0) / 1e12,
                           "peak_tlaneops": fp32_peak / 1e12,
                           "frac": lane_ops / (scan_ms_per_launch / 1000.0) / fp32_peak,
                           "ops_per_element": ops_per_elem},
        }
        if refine_rot:
            # one MUFU.SQRT per element at 16 per clock and SM is the binding unit of this form
            elems = float(q_hi - q_lo) * rows_here * dim
            mufu_peak = 148 * 16 * sm_mhz * 1e6
            roofline["mufu"] = {"achieved_telem": elems / (scan_ms_per_launch / 1000.0) / 1e12,
                                "peak_telem": mufu_peak / 1e12,
                                "frac": elems / (scan_ms_per_launch / 1000.0) / mufu_peak}
            roofline["kernel"] = "scan_kernel<EL_ROT, APPROX> (approximate-sqrt bound-and-refine, fp32 pipes + MUFU)"

    # ---- CPU baseline (oracle port) on a bounded sample + parity on that sample -----------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import kge_oracle as oracle
        kind = S.ORACLE_KIND[code]
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        big = n_ent >= 100000
        ns = min(n_test, args.cpu_sample if big else 2048)
        if mode == "entities":
            full = S.make_tables(code, dim, n_ent, n_rel, 0, n_ent, args.seed, dev)
        else:
            full = tabs
        P = oracle_params_from_tables(kind, {k: (v.cpu() if v is not None else None) for k, v in full.items()})
        th, tt, tr = (graph[k][:ns].cpu() for k in ("test_h", "test_t", "test_r"))
        dh, dt = S.filters_as_dicts(graph, n_ent, n_rel, limit=ns)  # reference-style dicts, sample keys
        b_size = 4 if big else 256
        t0 = time.perf_counter()
        ref = oracle.link_prediction(kind, P, th, tt, tr, dh, dt, b_size)
        cpu_s = time.perf_counter() - t0
        got = [x[:ns].cpu() for x in ranks_dev]
        equal = [bool(torch.equal(a, b)) for a, b in zip(got, ref)]
        cpu = {"value": ns / cpu_s, "unit": "triples/s", "cores": cores, "kind": "port",
               "sample": "first %d test triples, b_size %d, torch %s CPU, %d threads" % (
                   ns, b_size, torch.__version__, cores),
               "parity_on_sample": {"ranks_equal": all(equal), "n": ns,
                                    "vectors": dict(zip(["rank_heads", "rank_tails", "filt_rank_heads",
                                                         "filt_rank_tails"], equal))}}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        "metric": "filtered link-prediction triples/sec (full hits@k/MRR eval)",
        "value": value, "unit": "triples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl, world, mode),
        "e2e": {"value": e2e_value, "unit": "triples/s",
                "h2d_bytes_per_step": evaluator.last_stats.get("h2d_bytes"),
                "d2h_bytes_per_step": evaluator.last_stats.get("d2h_bytes"),
                "ms_per_step": 1000 * e2e_s / args.steps,
                "api": "LinkPredictionEvaluator(model, kg).evaluate(b_size=256), kg = torchkge_b200.KnowledgeGraph "
                       "(host index tensors + sorted-array filter index of all facts)",
                "equals_device_ranks": bool(same)},
        "gpu_launches": launches,
        "clocks": clock_rec,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "filter_csr_build_s": csr_build_s,
        "filter_index_build_s": filter_index_build_s,
        "mean_filter_set": float(csr_t[1].numel() + csr_h[1].numel()) / (2 * n_test),
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank, local, world = dist_info()
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (
            args.gpus, args.gpus, world))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local, world)


if __name__ == "__main__":
    main()
