"""Markdown tables from an Nsight Compute report (what profiles/*_ncu_summary.md are made of).

    python scripts/ncu_summary.py gpurun_out/tc_scan_c2.ncu-rep            # key metrics per kernel
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep --hot 25            # + hottest SASS lines (needs --import-source)
    python scripts/ncu_summary.py --launches gpurun_out/launches.csv       # per-kernel time shares of a launch list

Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (ncu must be on PATH;
no GPU needed).
"""
import argparse
import collections
import csv
import io
import subprocess
import sys

KEY = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_table(rep, extra):
    rows = ncu_csv(rep, "raw")
    hdr, units, kernels = rows[0], rows[1], rows[2:]
    want = KEY + list(extra)
    name_col = hdr.index("Kernel Name")
    print("| metric | " + " | ".join(k[name_col].split("(")[0][-40:] for k in kernels) + " | unit |")
    print("|---|" + "---|" * (len(kernels) + 1))
    for m in want:
        if m in hdr:
            i = hdr.index(m)
            print("| %s | %s | %s |" % (m, " | ".join(k[i] for k in kernels), units[i]))


def hot_lines(rep, n):
    rows = ncu_csv(rep, "source")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix["# Samples"]]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    print("\n| SASS | executed | samples | share |\n|---|---|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
        s = int(r[ix["# Samples"]])
        print("| `%s` | %s | %d | %.1f %% |" % (r[ix["Source"]].strip()[:70], r[ix["Instructions Executed"]], s, 100.0 * s / tot))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        v = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else v)
        a = agg.setdefault(r[ix["Kernel Name"]].split("(")[0][-64:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.2f | %.3f |" % (k, a[0], a[1], a[1] / tot))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report", nargs="?")
    ap.add_argument("--hot", type=int, default=0, help="also list the N hottest SASS lines")
    ap.add_argument("--metric", action="append", default=[], help="extra raw metric names")
    ap.add_argument("--launches", help="launch-list csv (ncu --metrics gpu__time_duration.sum --csv)")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches)
    if a.report:
        raw_table(a.report, a.metric)
        if a.hot:
            hot_lines(a.report, a.hot)
    if not a.launches and not a.report:
        ap.print_help()
        sys.exit(1)


if __name__ == "__main__":
    main()
