"""Dependency-free SVG charts for the benchmark report.

The reference's product is eight matplotlib/seaborn figures (main.py:64-390); neither library exists in this
environment, so the same figures are always emitted as ``*_comparison.svg`` by this ~150-line renderer (line charts,
grouped bars, stacked-bar panels, radar) and additionally as PNG when matplotlib imports."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

PALETTE = ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728", "#9467bd", "#8c564b", "#e377c2", "#7f7f7f", "#bcbd22", "#17becf"]
W, H = 900, 560
ML, MR, MT, MB = 80, 220, 50, 70          # margins (legend lives in the right margin)


def _esc(s) -> str:
    return str(s).replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;")


def _fmt(v: float) -> str:
    if v == 0:
        return "0"
    a = abs(v)
    if a >= 1000:
        return f"{v:.0f}"
    if a >= 10:
        return f"{v:.1f}"
    if a >= 1:
        return f"{v:.2f}"
    return f"{v:.3g}"


def _nice_max(v: float) -> float:
    if v <= 0:
        return 1.0
    e = 10 ** math.floor(math.log10(v))
    for m in (1, 2, 2.5, 5, 10):
        if v <= m * e:
            return m * e
    return 10 * e


def _frame(title: str, xlabel: str, ylabel: str, ymax: float, body: List[str], legend: Sequence[Tuple[str, str]],
           ymin: float = 0.0) -> str:
    pw, ph = W - ML - MR, H - MT - MB
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" viewBox="0 0 {W} {H}" '
           f'font-family="Helvetica,Arial,sans-serif" font-size="12">',
           f'<rect width="{W}" height="{H}" fill="white"/>',
           f'<text x="{ML + pw / 2}" y="28" text-anchor="middle" font-size="16" font-weight="bold">{_esc(title)}</text>']
    for i in range(6):                                           # horizontal grid + y ticks
        yv = ymin + (ymax - ymin) * i / 5
        y = MT + ph - ph * i / 5
        out.append(f'<line x1="{ML}" y1="{y:.1f}" x2="{ML + pw}" y2="{y:.1f}" stroke="#dddddd"/>')
        out.append(f'<text x="{ML - 8}" y="{y + 4:.1f}" text-anchor="end">{_fmt(yv)}</text>')
    out.append(f'<rect x="{ML}" y="{MT}" width="{pw}" height="{ph}" fill="none" stroke="#333333"/>')
    out.append(f'<text x="{ML + pw / 2}" y="{H - 18}" text-anchor="middle" font-size="13">{_esc(xlabel)}</text>')
    out.append(f'<text transform="translate(20,{MT + ph / 2}) rotate(-90)" text-anchor="middle" font-size="13">'
               f'{_esc(ylabel)}</text>')
    out += body
    for i, (lab, col) in enumerate(legend):
        y = MT + 10 + 20 * i
        out.append(f'<rect x="{W - MR + 16}" y="{y - 9}" width="14" height="10" fill="{col}"/>')
        out.append(f'<text x="{W - MR + 36}" y="{y}">{_esc(lab)}</text>')
    out.append("</svg>")
    return "\n".join(out)


def line_chart(series: Dict[str, List[Tuple[float, float]]], title: str, xlabel: str, ylabel: str) -> str:
    """series: {label: [(x, y), ...]}"""
    pts = [p for s in series.values() for p in s]
    if not pts:
        return _frame(title, xlabel, ylabel, 1.0, [], [])
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    x0, x1 = min(xs), max(xs)
    if x1 == x0:
        x0, x1 = x0 - 0.5, x1 + 0.5
    ymax = _nice_max(max(ys))
    pw, ph = W - ML - MR, H - MT - MB
    sx = lambda x: ML + (x - x0) / (x1 - x0) * pw        # noqa: E731
    sy = lambda y: MT + ph - max(y, 0.0) / ymax * ph     # noqa: E731
    body, legend = [], []
    for xv in sorted(set(xs)):
        body.append(f'<text x="{sx(xv):.1f}" y="{MT + ph + 18}" text-anchor="middle">{_fmt(xv)}</text>')
    for i, (lab, s) in enumerate(series.items()):
        col = PALETTE[i % len(PALETTE)]
        s = sorted(s)
        body.append(f'<polyline fill="none" stroke="{col}" stroke-width="2" points="' +
                    " ".join(f"{sx(x):.1f},{sy(y):.1f}" for x, y in s) + '"/>')
        body += [f'<circle cx="{sx(x):.1f}" cy="{sy(y):.1f}" r="3.5" fill="{col}"/>' for x, y in s]
        legend.append((lab, col))
    return _frame(title, xlabel, ylabel, ymax, body, legend)


def grouped_bars(groups: Sequence[str], series: Dict[str, List[float]], title: str, xlabel: str, ylabel: str,
                 stacked: bool = False) -> str:
    """groups: x-axis categories; series: {label: [value per group]} — side by side, or stacked."""
    n_g, labels = len(groups), list(series)
    tot = [sum(series[k][g] for k in labels) for g in range(n_g)] if stacked else \
        [max([series[k][g] for k in labels] or [0]) for g in range(n_g)]
    ymax = _nice_max(max(tot or [1.0]))
    pw, ph = W - ML - MR, H - MT - MB
    gw = pw / max(n_g, 1)
    body, legend = [], [(k, PALETTE[i % len(PALETTE)]) for i, k in enumerate(labels)]
    for g, name in enumerate(groups):
        gx = ML + g * gw
        body.append(f'<text x="{gx + gw / 2:.1f}" y="{MT + ph + 18}" text-anchor="middle">{_esc(name)}</text>')
        base = 0.0
        bw = gw * 0.8 / (1 if stacked else max(len(labels), 1))
        for i, k in enumerate(labels):
            v = max(series[k][g], 0.0)
            h = v / ymax * ph
            x = gx + gw * 0.1 + (0 if stacked else i * bw)
            y = MT + ph - h - (base / ymax * ph if stacked else 0)
            body.append(f'<rect x="{x:.1f}" y="{y:.1f}" width="{bw:.1f}" height="{h:.1f}" fill="{PALETTE[i % len(PALETTE)]}">'
                        f'<title>{_esc(k)} / {_esc(name)}: {_fmt(v)}</title></rect>')
            if stacked:
                base += v
    return _frame(title, xlabel, ylabel, ymax, body, legend)


def radar(scores: Dict[str, Dict[str, float]], title: str) -> str:
    """scores: {label: {axis: value in [0, 1]}}"""
    cats = list(next(iter(scores.values())).keys()) if scores else []
    cx, cy, R = (W - MR) / 2 + 20, H / 2 + 15, 200
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" viewBox="0 0 {W} {H}" '
           f'font-family="Helvetica,Arial,sans-serif" font-size="12">', f'<rect width="{W}" height="{H}" fill="white"/>',
           f'<text x="{cx}" y="28" text-anchor="middle" font-size="16" font-weight="bold">{_esc(title)}</text>']
    n = max(len(cats), 1)
    ang = [-math.pi / 2 + 2 * math.pi * i / n for i in range(n)]
    for frac in (0.25, 0.5, 0.75, 1.0):
        out.append('<polygon fill="none" stroke="#dddddd" points="' +
                   " ".join(f"{cx + R * frac * math.cos(a):.1f},{cy + R * frac * math.sin(a):.1f}" for a in ang) + '"/>')
    for a, c in zip(ang, cats):
        out.append(f'<line x1="{cx}" y1="{cy}" x2="{cx + R * math.cos(a):.1f}" y2="{cy + R * math.sin(a):.1f}" stroke="#bbbbbb"/>')
        out.append(f'<text x="{cx + (R + 22) * math.cos(a):.1f}" y="{cy + (R + 22) * math.sin(a) + 4:.1f}" '
                   f'text-anchor="middle">{_esc(c)}</text>')
    for i, (lab, sc) in enumerate(scores.items()):
        col = PALETTE[i % len(PALETTE)]
        pts = " ".join(f"{cx + R * max(min(sc[c], 1.0), 0.0) * math.cos(a):.1f},"
                       f"{cy + R * max(min(sc[c], 1.0), 0.0) * math.sin(a):.1f}" for a, c in zip(ang, cats))
        out.append(f'<polygon points="{pts}" fill="{col}" fill-opacity="0.12" stroke="{col}" stroke-width="2"/>')
        y = MT + 10 + 20 * i
        out.append(f'<rect x="{W - MR + 16}" y="{y - 9}" width="14" height="10" fill="{col}"/>')
        out.append(f'<text x="{W - MR + 36}" y="{y}">{_esc(lab)}</text>')
    out.append("</svg>")
    return "\n".join(out)
