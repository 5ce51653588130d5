"""Benchmark sweep + comparison reports (reference: main.py:17-61 ``run_benchmarks`` and
main.py:64-390 ``generate_comparison_graphs``).

Same programme: for every sample size run data-parallel, layer-parallel, tensor-parallel in turn,
then emit the eight comparison artefacts (accuracy, loss, training time, compute-vs-comm, CPU,
memory, idle, overall radar).  Differences: the real ``world_size`` is used to pick the
last pipeline rank (the reference hard-codes 5), a failed strategy degrades gracefully instead of
crashing, and because matplotlib/seaborn are optional the figures are always drawn as
``*_comparison.svg`` by the built-in renderer (``svgplot.py``) next to the numbers behind them
(``*_comparison.csv`` + ``benchmark_summary.json``); PNGs are drawn as well when matplotlib imports.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional

import numpy as np
import pandas as pd

from .config import TrainConfig
from .trainers import run_data_parallel, run_model_parallel, run_tensor_parallel

STRATEGIES = ["data_parallel", "model_parallel", "tensor_parallel"]
LABEL = {"data_parallel": "Data Parallel", "model_parallel": "Model Parallel",
         "tensor_parallel": "Tensor Parallel"}
FIGURES = ["accuracy", "loss", "training_time", "compute_vs_comm", "cpu_utilization", "memory_usage",
           "idle_time", "overall_performance"]


def run_benchmarks(sample_sizes: List[int], world_size: int, epochs: int, cfg: Optional[TrainConfig] = None,
                   strategies: Optional[List[str]] = None) -> Dict[str, Dict[int, Optional[pd.DataFrame]]]:
    """Sequential sweep; returns {'data_parallel'|'model_parallel'|'tensor_parallel': {N: DataFrame|None}}."""
    results: Dict[str, Dict[int, Optional[pd.DataFrame]]] = {s: {} for s in STRATEGIES}
    runners = {"data_parallel": (run_data_parallel, "data_parallel_logs"),
               "model_parallel": (run_model_parallel, "model_parallel_logs"),
               "tensor_parallel": (run_tensor_parallel, "tensor_parallel_logs")}
    base = cfg or TrainConfig()
    for n in sample_sizes:
        print(f"\n{'=' * 50}\nRunning benchmarks with sample size {n}\n{'=' * 50}")
        for s in (strategies or STRATEGIES):
            fn, logs = runners[s]
            ws = world_size
            if s == "model_parallel":
                ws = min(world_size, 5)       # five atomic blocks (layer_…:54-55)
            print(f"\nRunning {LABEL[s]} training with {n} samples...")
            try:
                df = fn(ws, epochs, n, logs_dir=base.logs_dir or logs, cfg=base)
            except Exception as e:  # noqa: BLE001
                print(f"{LABEL[s]} training failed: {e}")
                df = None
            results[s][n] = df
            if df is not None:
                t = df["total_training_time"].iloc[0]
                print(f"{LABEL[s]} training completed in {t:.2f} seconds")
                if df["epoch"].max() < epochs:
                    print(f"WARNING: {LABEL[s]} only completed {df['epoch'].max()} of {epochs} epochs")
            else:
                print(f"WARNING: {LABEL[s]} training with {n} samples did not complete successfully")
    return results


def _rows_for_curves(strategy: str, df: pd.DataFrame) -> pd.DataFrame:
    """Rows that carry loss/accuracy: all ranks for DP/TP (mean per epoch, main.py:73-85), the last
    pipeline rank for layer-parallel (true last rank, not a hard-coded 4)."""
    if strategy == "model_parallel":
        return df[df["worker"] == df["worker"].max()]
    return df


def summarize(results) -> Dict:
    """All numbers behind the reference's eight figures."""
    out: Dict = {"curves": [], "bars": []}
    for s, per_n in results.items():
        for n, df in per_n.items():
            if df is None or len(df) == 0:
                continue
            cur = _rows_for_curves(s, df).groupby("epoch")[["loss", "accuracy"]].mean().reset_index()
            for _, r in cur.iterrows():
                out["curves"].append({"strategy": s, "sample_size": n, "epoch": int(r["epoch"]),
                                      "loss": float(r["loss"]), "accuracy": float(r["accuracy"])})
            bar = {"strategy": s, "sample_size": n,
                   "epoch_time": float(df["epoch_time"].mean()),
                   "compute_time": float(df["compute_time"].mean()),
                   "comm_time": float(df["comm_time"].mean()),
                   "idle_time": float(df["idle_time"].mean()),
                   "avg_cpu": float(df["avg_cpu"].mean()),
                   "avg_memory": float(df["avg_memory"].mean()),
                   "total_training_time": float(df["total_training_time"].iloc[0]),
                   "final_accuracy": float(cur["accuracy"].iloc[-1]) if len(cur) else 0.0}
            if "images_per_sec" in df:
                bar["images_per_sec"] = float(df.groupby("epoch")["images_per_sec"].max().mean())
            out["bars"].append(bar)
    return out


def radar_scores(bars: List[Dict], sample_size: int) -> Dict[str, Dict[str, float]]:
    """Six normalised metrics at the largest sample size; 'lower is better' ones as 1 − x/max
    (main.py:335-351)."""
    rows = [b for b in bars if b["sample_size"] == sample_size]
    if not rows:
        return {}

    def norm(key, lower_better):
        mx = max(r[key] for r in rows) or 1.0
        return {r["strategy"]: (1 - r[key] / mx) if lower_better else (r[key] / mx) for r in rows}

    acc = norm("final_accuracy", False)
    tt, comm = norm("epoch_time", True), norm("comm_time", True)
    idle, cpu, mem = norm("idle_time", True), norm("avg_cpu", True), norm("avg_memory", True)
    return {r["strategy"]: {"Accuracy": acc[r["strategy"]], "Training Speed": tt[r["strategy"]],
                            "Communication Efficiency": comm[r["strategy"]],
                            "Idle Time": idle[r["strategy"]], "CPU Efficiency": cpu[r["strategy"]],
                            "Memory Efficiency": mem[r["strategy"]]} for r in rows}


def generate_comparison_graphs(results, output_dir: str = "benchmark_results") -> Dict[str, str]:
    """Writes the eight comparison artefacts. Returns {figure: path}."""
    os.makedirs(output_dir, exist_ok=True)
    summ = summarize(results)
    curves, bars = pd.DataFrame(summ["curves"]), pd.DataFrame(summ["bars"])
    written: Dict[str, str] = {}
    sizes = sorted({b["sample_size"] for b in summ["bars"]})
    radar = radar_scores(summ["bars"], sizes[-1]) if sizes else {}
    tables = {
        "accuracy": curves[["strategy", "sample_size", "epoch", "accuracy"]] if len(curves) else curves,
        "loss": curves[["strategy", "sample_size", "epoch", "loss"]] if len(curves) else curves,
        "training_time": bars[["strategy", "sample_size", "epoch_time"]] if len(bars) else bars,
        "compute_vs_comm": bars[["strategy", "sample_size", "compute_time", "comm_time"]] if len(bars) else bars,
        "cpu_utilization": bars[["strategy", "sample_size", "avg_cpu"]] if len(bars) else bars,
        "memory_usage": bars[["strategy", "sample_size", "avg_memory"]] if len(bars) else bars,
        "idle_time": bars[["strategy", "sample_size", "idle_time"]] if len(bars) else bars,
        "overall_performance": pd.DataFrame(radar).T.reset_index().rename(columns={"index": "strategy"}),
    }
    for name, tab in tables.items():
        p = os.path.join(output_dir, f"{name}_comparison.csv")
        tab.to_csv(p, index=False)
        written[name + "_csv"] = p
    with open(os.path.join(output_dir, "benchmark_summary.json"), "w") as fh:
        json.dump({"bars": summ["bars"], "radar": radar}, fh, indent=1)
    _draw_svg(summ, radar, output_dir, written)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        print("matplotlib not available: wrote CSV/JSON summaries and SVG figures")
        return written
    _draw_all(plt, curves, bars, radar, output_dir, written)
    return written


def _draw_svg(summ: Dict, radar: Dict, out: str, written: Dict[str, str]) -> None:
    """The reference's eight figures (main.py:64-390) as SVG, with no plotting dependency."""
    from . import svgplot

    def save(name, svg):
        p = os.path.join(out, f"{name}_comparison.svg")
        with open(p, "w") as fh:
            fh.write(svg)
        written[name + "_svg"] = p

    curves, bars = summ["curves"], summ["bars"]
    for metric, ylabel in (("accuracy", "Accuracy (%)"), ("loss", "Loss")):
        series: Dict[str, list] = {}
        for c in curves:
            series.setdefault(f"{LABEL[c['strategy']]} ({c['sample_size']} samples)", []).append((c["epoch"], c[metric]))
        save(metric, svgplot.line_chart(series, f"{ylabel} Comparison", "Epoch", ylabel))
    sizes = sorted({b["sample_size"] for b in bars})
    strategies = [s for s in STRATEGIES if any(b["strategy"] == s for b in bars)]

    def val(s, n, col):
        for b in bars:
            if b["strategy"] == s and b["sample_size"] == n:
                return float(b[col])
        return 0.0

    for name, col, ylabel in (("training_time", "epoch_time", "Average Epoch Time (s)"),
                              ("cpu_utilization", "avg_cpu", "CPU Utilization (%)"),
                              ("memory_usage", "avg_memory", "Memory Usage (MB)"),
                              ("idle_time", "idle_time", "Idle Time (s)")):
        series = {LABEL[s]: [val(s, n, col) for n in sizes] for s in strategies}
        save(name, svgplot.grouped_bars([str(n) for n in sizes], series, ylabel, "Sample size", ylabel))
    short = {"data_parallel": "DP", "model_parallel": "MP", "tensor_parallel": "TP"}
    groups = [f"{short[s]} {n}" for n in sizes for s in strategies]
    series = {"Compute": [val(s, n, "compute_time") for n in sizes for s in strategies],
              "Communication": [val(s, n, "comm_time") for n in sizes for s in strategies]}
    save("compute_vs_comm", svgplot.grouped_bars(groups, series, "Compute vs Communication Time",
                                                 "Strategy (DP data / MP model / TP tensor parallel) and sample size",
                                                 "Time (s, cumulative)", stacked=True))
    if radar:
        save("overall_performance", svgplot.radar({LABEL[s]: sc for s, sc in radar.items()},
                                                  f"Overall Performance ({sizes[-1]} samples)" if sizes else "Overall"))


def _draw_all(plt, curves, bars, radar, out, written):  # pragma: no cover - needs matplotlib
    def save(name):
        p = os.path.join(out, f"{name}_comparison.png")
        plt.tight_layout(); plt.savefig(p); plt.close()
        written[name] = p

    for metric, ylabel in (("accuracy", "Accuracy (%)"), ("loss", "Loss")):
        plt.figure(figsize=(12, 8))
        for (s, n), g in curves.groupby(["strategy", "sample_size"]):
            plt.plot(g["epoch"], g[metric], marker="o", label=f"{LABEL[s]} ({n} samples)")
        plt.xlabel("Epoch"); plt.ylabel(ylabel); plt.title(f"{ylabel} Comparison"); plt.legend(); plt.grid(True)
        save(metric)
    for name, col, ylabel in (("training_time", "epoch_time", "Average Epoch Time (s)"),
                              ("cpu_utilization", "avg_cpu", "CPU Utilization (%)"),
                              ("memory_usage", "avg_memory", "Memory Usage (MB)"),
                              ("idle_time", "idle_time", "Idle Time (s)")):
        plt.figure(figsize=(12, 8))
        piv = bars.pivot(index="sample_size", columns="strategy", values=col)
        piv.rename(columns=LABEL).plot(kind="bar", ax=plt.gca())
        plt.ylabel(ylabel); plt.title(ylabel)
        save(name)
    sizes = sorted(bars["sample_size"].unique())
    fig, axes = plt.subplots(1, max(len(sizes), 1), figsize=(6 * max(len(sizes), 1), 6), squeeze=False)
    for ax, n in zip(axes[0], sizes):
        sub = bars[bars["sample_size"] == n]
        ax.bar([LABEL[s] for s in sub["strategy"]], sub["compute_time"], label="Compute")
        ax.bar([LABEL[s] for s in sub["strategy"]], sub["comm_time"], bottom=sub["compute_time"], label="Comm")
        ax.set_title(f"{n} samples"); ax.legend()
    save("compute_vs_comm")
    if radar:
        cats = list(next(iter(radar.values())).keys())
        ang = np.linspace(0, 2 * np.pi, len(cats), endpoint=False).tolist()
        ang += ang[:1]
        plt.figure(figsize=(10, 10)); ax = plt.subplot(111, polar=True)
        for s, sc in radar.items():
            v = [sc[c] for c in cats]; v += v[:1]
            ax.plot(ang, v, label=LABEL[s]); ax.fill(ang, v, alpha=0.1)
        ax.set_xticks(ang[:-1]); ax.set_xticklabels(cats); plt.legend()
        save("overall_performance")


def load_results(sample_sizes: Optional[List[int]] = None, logs_root: str = ".",
                 logs_dirs: Optional[Dict[str, str]] = None) -> Dict[str, Dict[int, Optional[pd.DataFrame]]]:
    """Re-read ``combined_results_{N}.csv`` files written by earlier runs (the offline analysis step the
    reference README promises as ``analyze_results.py``, README.md:21, but never ships)."""
    import glob
    import re
    dirs = {"data_parallel": "data_parallel_logs", "model_parallel": "model_parallel_logs",
            "tensor_parallel": "tensor_parallel_logs"}
    dirs.update(logs_dirs or {})
    results: Dict[str, Dict[int, Optional[pd.DataFrame]]] = {s: {} for s in STRATEGIES}
    for s in STRATEGIES:
        d = os.path.join(logs_root, dirs[s])
        found = {}
        for f in glob.glob(os.path.join(d, "combined_results_*.csv")):
            m = re.search(r"combined_results_(\d+)\.csv$", f)
            if m:
                found[int(m.group(1))] = f
        for n in (sample_sizes or sorted(found)):
            results[s][n] = pd.read_csv(found[n]) if n in found else None
    return results


def analyze_main(argv=None) -> int:
    import argparse
    p = argparse.ArgumentParser(description="Rebuild the comparison reports from existing training logs")
    p.add_argument('--sample_sizes', type=int, nargs='*', default=None, help='default: every size found')
    p.add_argument('--logs_root', default='.', help='directory holding data_/model_/tensor_parallel_logs')
    p.add_argument('--output_dir', default='benchmark_results')
    args = p.parse_args(argv)
    results = load_results(args.sample_sizes, args.logs_root)
    if not any(df is not None for per in results.values() for df in per.values()):
        print(f"no combined_results_*.csv under {args.logs_root}")
        return 1
    written = generate_comparison_graphs(results, args.output_dir)
    print(f"Analysis completed: {len(written)} artefacts in {args.output_dir}")
    return 0


def main(argv=None) -> int:
    import argparse
    from .config import add_train_flags, config_from_args
    p = argparse.ArgumentParser(description="Compare different parallelism strategies")
    p.add_argument('--sample_sizes', type=int, nargs='+', default=[1000, 10000, 50000],
                   help='Sample sizes to benchmark')
    p.add_argument('--output_dir', type=str, default='benchmark_results',
                   help='Directory to save comparison graphs')
    p.add_argument('--strategies', nargs='+', default=None, choices=STRATEGIES)
    add_train_flags(p, "data")
    args = p.parse_args(argv)
    cfg = config_from_args(args, "data")
    results = run_benchmarks(args.sample_sizes, args.world_size, args.epochs, cfg, args.strategies)
    generate_comparison_graphs(results, args.output_dir)
    print(f"Benchmarking completed. Results saved to {args.output_dir}")
    return 0
