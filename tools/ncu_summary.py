"""Per-kernel summary of an Nsight Compute capture — runs where there is no GPU (`ncu -i` only reads the report).

    python tools/ncu_summary.py gpurun_out/ncu/conv.ncu-rep [more.ncu-rep ...] [--out profiles/ncu_summary.md] [--top 20]
    ncu -i x.ncu-rep --page raw --csv | python tools/ncu_summary.py -              # or a raw-page CSV on stdin

For every kernel (template arguments kept, parameter list dropped): launches, summed duration and its share of the
capture, and — duration-weighted over its launches — SM throughput, DRAM throughput, active warps, tensor-pipe
utilisation, plus registers per thread and DRAM traffic.  The raw-page metric names are the ones
`/opt/skills/guides/B200_PROFILING.md` greps for; `aggregate()` is also what tests/test_gpu_ncu_report.py uses for the
profile it takes inside the GPU test run.  Numbers taken under the profiler are shares and utilisations, never step times."""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

DURATION = "gpu__time_duration.sum"
SM = "sm__throughput.avg.pct_of_peak_sustained_elapsed"
DRAM = "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
WARPS = "sm__warps_active.avg.pct_of_peak_sustained_active"
REGS = "launch__registers_per_thread"
RD, WR = "dram__bytes_read.sum", "dram__bytes_write.sum"
TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
BASE_METRICS = [DURATION, SM, DRAM, WARPS, REGS, RD, WR]


def short_name(name: str) -> str:
    name = name.split("(")[0]
    for pre in ("void hz::", "hz::", "void "):
        if name.startswith(pre):
            name = name[len(pre):]
    return name[:60]


def _scale(unit: str, metric: str) -> float:
    """raw-page values come in the unit of the second header row; normalise durations to ns and traffic to bytes"""
    u = (unit or "").strip().lower()
    if metric == DURATION:
        return {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(u, 1.0)
    if metric in (RD, WR):
        return {"byte": 1.0, "bytes": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return 1.0


def aggregate(text: str):
    """raw-page CSV text (anything before the header line is skipped) -> (n_launches, OrderedDict kernel -> sums, columns)
    or None if the text holds no capture."""
    start = text.find('"ID"')
    if start < 0:
        return None
    rows = list(csv.reader(io.StringIO(text[start:])))
    if len(rows) < 3:
        return None
    header, units = rows[0], rows[1]
    rows = [r for r in rows[2:] if len(r) == len(header)]
    col = {h: i for i, h in enumerate(header)}
    if "Kernel Name" not in col or DURATION not in col:
        return None
    mul = {m: _scale(units[col[m]] if col[m] < len(units) else "", m) for m in (DURATION, RD, WR) if m in col}

    def num(r, m):
        try:
            return float(r[col[m]].replace(",", "")) * mul.get(m, 1.0)
        except (KeyError, ValueError):
            return float("nan")

    def nz(v):
        return v if v == v else 0.0
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short_name(r[col["Kernel Name"]]),
                           {"launches": 0, "ns": 0.0, "sm": 0.0, "dram": 0.0, "occ": 0.0, "tensor": 0.0, "regs": 0, "bytes": 0.0})
        t = nz(num(r, DURATION))
        a["launches"] += 1
        a["ns"] += t
        a["sm"] += nz(num(r, SM)) * t                      # duration-weighted utilisations
        a["dram"] += nz(num(r, DRAM)) * t
        a["occ"] += nz(num(r, WARPS)) * t
        a["tensor"] += nz(num(r, TENSOR)) * t
        a["regs"] = max(a["regs"], int(nz(num(r, REGS))))
        a["bytes"] += nz(num(r, RD)) + nz(num(r, WR))
    return len(rows), agg, col


def summarize(agg, col, top: int):
    """rows for the `top` kernels by summed duration"""
    total = sum(a["ns"] for a in agg.values()) or 1.0
    out = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])[:top]:
        t = a["ns"] or 1.0
        out.append({"kernel": k, "launches": a["launches"], "us": round(a["ns"] / 1e3, 1), "share": round(a["ns"] / total, 3),
                    "sm_pct": round(a["sm"] / t, 1), "dram_pct": round(a["dram"] / t, 1),
                    "warps_active_pct": round(a["occ"] / t, 1),
                    "tensor_pipe_pct": (round(a["tensor"] / t, 1) if TENSOR in col else None),
                    "regs": a["regs"], "dram_MB": round(a["bytes"] / 1e6, 2)})
    return total, out


def raw_csv_of(report: str) -> str:
    exe = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    r = subprocess.run([exe, "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        raise SystemExit(f"ncu -i {report} failed: {(r.stderr or r.stdout)[-400:]}")
    return r.stdout


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("reports", nargs="+", help=".ncu-rep files, or - for a raw-page CSV on stdin")
    ap.add_argument("--out", default=None, help="write a markdown table here as well (e.g. profiles/ncu_summary.md)")
    ap.add_argument("--top", type=int, default=20)
    args = ap.parse_args()
    md = []
    for rep in args.reports:
        text = sys.stdin.read() if rep == "-" else raw_csv_of(rep)
        res = aggregate(text)
        if res is None:
            print(f"{rep}: no raw-page capture found", file=sys.stderr)
            continue
        n, agg, col = res
        total, rows = summarize(agg, col, args.top)
        print(json.dumps({"report": rep, "launches": n, "distinct_kernels": len(agg), "sum_of_durations_us": round(total / 1e3, 1)}))
        md += [f"### {os.path.basename(rep)} — {n} launches, {len(agg)} kernels, {total / 1e3:.1f} µs summed (serialised under the profiler)", "",
               "| kernel | launches | µs | share | SM % | DRAM % | warps active % | tensor pipe % | regs | DRAM MB |", "|---|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            print(json.dumps(r))
            md.append(f"| `{r['kernel']}` | {r['launches']} | {r['us']} | {r['share']:.3f} | {r['sm_pct']} | {r['dram_pct']} | "
                      f"{r['warps_active_pct']} | {r['tensor_pipe_pct'] if r['tensor_pipe_pct'] is not None else '–'} | {r['regs']} | {r['dram_MB']} |")
        md.append("")
    if args.out and md:
        with open(args.out, "w") as f:
            f.write("\n".join(md))
        print("wrote", args.out)


if __name__ == "__main__":
    main()
