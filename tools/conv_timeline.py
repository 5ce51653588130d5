"""In-kernel timeline of the tcgen05 implicit-GEMM convolution: per-CTA clock64 stamps at the phase boundaries
(entry → prologue → griddepcontrol.wait → first operand stage landed → last MMA issued → accumulator complete →
tile staged (cluster split-K reduced) → rows + BN sums written → done), for every distinct ResNet-18 conv at
batch 64, forward and dgrad, with L2 flushed (weights cold, as in the benchmark) and hot.
Diagnosis only (the stamps cost a few predicated stores).  Usage: python tools/conv_timeline.py out.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops  # noqa: E402
from horizonml_b200.ops import native_backend as nb  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_timeline.json"
dev = torch.device("cuda", 0)
ops.set_backend("native")
C = nb.C
GHZ = 1.965          # SM clock under load on this part (bench.py samples 1965 MHz)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PH = ["prologue", "pdl_wait", "first_stage", "mma_issue", "mma_drain", "reduce_stage", "store_stats", "teardown"]
CONVS = [(64, 64, 8, 8, 64, 3, 1, 1), (64, 64, 8, 8, 128, 3, 2, 1), (64, 128, 4, 4, 128, 3, 1, 1),
         (64, 256, 2, 2, 256, 3, 1, 1), (64, 512, 1, 1, 512, 3, 1, 1), (64, 256, 2, 2, 512, 1, 2, 0)]


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def run(kind, cfg, cold):
    N, Cin, H, W, Cout, R, s, p = cfg
    g = torch.Generator().manual_seed(1)
    x = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(dev).bfloat16())
    w = cl((torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(dev).bfloat16())
    w._hz_stable = True
    Ho = (H + 2 * p - R) // s + 1
    dy = cl((torch.randn(N, Cout, Ho, Ho, generator=g) * 0.5).to(dev).bfloat16())
    fn = (lambda: nb.conv_fwd(x, w, s, p, True)) if kind == "fwd" else (lambda: nb.conv_dgrad(dy, w, x.shape, s, p))
    for _ in range(3):
        fn()
    dbg = torch.zeros(512 * 16, dtype=torch.int64, device=dev)
    res = []
    for rep in range(5):
        dbg.zero_()
        if cold:
            flush.fill_(rep)
        torch.cuda.synchronize()
        C.conv_set_debug(dbg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        C.conv_set_debug(None)
        st = dbg.view(-1, 16).cpu()
        st = st[st[:, 0] != 0][:, :9].double()
        if st.numel() == 0:
            continue
        # stamps a CTA did not reach (e.g. k_iters == 0 splits never wait for operands) inherit the previous one
        for j in range(1, 9):
            st[:, j] = torch.where(st[:, j] == 0, st[:, j - 1], st[:, j])
        d = (st[:, 1:] - st[:, :-1]) / GHZ / 1e3                      # us per phase per CTA
        res.append({"event_us": e0.elapsed_time(e1) * 1e3, "ctas": int(st.shape[0]),
                    "cta_span_us_median": float(((st[:, 8] - st[:, 0]) / GHZ / 1e3).median()),
                    "phases_us_median": {k: round(float(d[:, i].median()), 3) for i, k in enumerate(PH)}})
    res.sort(key=lambda r: r["event_us"])
    return res[len(res) // 2] if res else None


rows = []
for cfg in CONVS:
    for kind in ("fwd", "dgrad"):
        for cold in (True, False):
            r = run(kind, cfg, cold)
            row = {"conv": list(cfg), "kind": kind, "l2": "flushed" if cold else "hot", **(r or {})}
            rows.append(row)
            print(json.dumps(row), flush=True)
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump({"sm_ghz": GHZ, "phases": PH, "rows": rows}, open(out, "w"), indent=1)
