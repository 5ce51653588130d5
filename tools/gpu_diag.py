"""GPU diagnostics: every native kernel vs the PyTorch oracle, detailed error report as JSON.
Run on the GPU box:  python tools/gpu_diag.py gpurun_out/diag.json"""
import json
import sys
import traceback

import torch

sys.path.insert(0, ".")
from horizonml_b200.ops import native_backend as nb  # noqa: E402
from horizonml_b200.ops import torch_backend as tb  # noqa: E402

dev = torch.device("cuda", 0)
out = {"cases": []}


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def err(a, b):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    den = b.abs().max().item() + 1e-6
    bad = (d > 0.05 * den)
    return {"max_abs": d.max().item(), "ref_max": den, "rel": d.max().item() / den,
            "frac_bad": bad.float().mean().item(), "nan": bool(torch.isnan(a).any().item())}


def bad_pattern(a, b):
    """where are the bad entries: per physical row (n,h,w) / channel summary"""
    a, b = a.float(), b.float()
    d = (a - b).abs() > 0.05 * (b.abs().max().item() + 1e-6)
    if d.dim() == 4:
        per_c = d.sum(dim=(0, 2, 3)).tolist()
        per_n = d.sum(dim=(1, 2, 3)).tolist()
        per_hw = d.sum(dim=(0, 1)).flatten().tolist()
        return {"bad_per_channel_first32": per_c[:32], "bad_per_image_first16": per_n[:16], "bad_per_hw": per_hw[:64]}
    return {}


def record(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        r["name"] = name
        r["ok"] = bool(r.get("rel", 1) < 0.03 and not r.get("nan", False))
    except Exception as e:  # noqa: BLE001
        r = {"name": name, "ok": False, "exc": repr(e), "tb": traceback.format_exc()[-1500:]}
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa: BLE001
            r["sync_exc"] = repr(e2)
    out["cases"].append(r)
    print(("PASS " if r["ok"] else "FAIL ") + name + " " + json.dumps({k: v for k, v in r.items() if k in ("rel", "frac_bad", "exc", "nan")}), flush=True)
    return r["ok"]


CONVS = [  # (N, Cin, H, W, Cout, R, stride, pad)
    (64, 64, 8, 8, 64, 3, 1, 1), (64, 64, 8, 8, 128, 3, 2, 1), (64, 64, 8, 8, 128, 1, 2, 0),
    (64, 128, 4, 4, 128, 3, 1, 1), (64, 128, 4, 4, 256, 3, 2, 1), (64, 128, 4, 4, 256, 1, 2, 0),
    (64, 256, 2, 2, 256, 3, 1, 1), (64, 256, 2, 2, 512, 3, 2, 1), (64, 256, 2, 2, 512, 1, 2, 0),
    (64, 512, 1, 1, 512, 3, 1, 1), (16, 64, 8, 8, 64, 3, 1, 1), (8, 64, 16, 16, 64, 3, 1, 1),
]


def conv_case(kind, cfg):
    N, Cin, H, W, Cout, R, s, p = cfg
    g = torch.Generator(device="cpu").manual_seed(1)
    x = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(dev).bfloat16())
    w = cl((torch.randn(Cout, Cin, R, R, generator=g) * (1.0 / (Cin * R * R) ** 0.5)).to(dev).bfloat16())
    Ho, Wo = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    dy = cl((torch.randn(N, Cout, Ho, Wo, generator=g) * 0.5).to(dev).bfloat16())
    before = dict(nb.FALLBACKS)
    if kind == "fwd":
        y, st = nb.conv_fwd(x, w, s, p, True)
        yr, sr = tb.conv_fwd(x.float(), w.float(), s, p, True)
        e = err(y, yr)
        e["stats"] = err(st, sr)
        if e["rel"] > 0.03:
            e.update(bad_pattern(y, yr))
    elif kind == "dgrad":
        dx = nb.conv_dgrad(dy, w, x.shape, s, p)
        dxr = tb.conv_dgrad(dy.float(), w.float(), x.shape, s, p)
        e = err(dx, dxr)
        if e["rel"] > 0.03:
            e.update(bad_pattern(dx, dxr))
    else:
        gbuf = torch.zeros(Cout * R * R * Cin, device=dev)
        gv = gbuf.view(Cout, R, R, Cin).permute(0, 3, 1, 2)
        nb.conv_wgrad(dy, x, w.shape, s, p, gv, False)
        ref = torch.zeros(Cout, Cin, R, R, device=dev)
        tb.conv_wgrad(dy.float(), x.float(), w.shape, s, p, ref, False)
        e = err(gv, ref)
        if e["rel"] > 0.03:
            d = ((gv - ref).abs() > 0.05 * ref.abs().max()).float()
            e["bad_per_tap"] = d.sum(dim=(0, 1)).flatten().tolist()
            e["bad_per_cout_first16"] = d.sum(dim=(1, 2, 3)).tolist()[:16]
    e["fallback"] = dict(nb.FALLBACKS) != before
    if e["fallback"]:
        e["rel"] = 1.0
    return e


GROUP = sys.argv[2] if len(sys.argv) > 2 else "driver"
if GROUP == "driver":
    import os, subprocess
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/diag.json"
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    merged = {"cases": [], "groups": {}}
    for grp in ("elem", "comm", "conv_fwd", "conv_dgrad", "conv_wgrad", "stem"):
        gp = path + "." + grp
        try:
            r = subprocess.run([sys.executable, __file__, gp, grp], timeout=600)
            merged["groups"][grp] = r.returncode
        except subprocess.TimeoutExpired:
            merged["groups"][grp] = "timeout"
        if os.path.exists(gp):
            d = json.load(open(gp))
            merged["cases"] += d["cases"]
            merged.setdefault("launches", {}).update(d.get("launches", {}))
            os.remove(gp)
    merged["n_fail"] = sum(1 for c in merged["cases"] if not c["ok"])
    json.dump(merged, open(path, "w"), indent=1)
    print(f"diag: {len(merged['cases'])} cases, {merged['n_fail']} failed, groups={merged['groups']}")
    sys.exit(0)

if GROUP.startswith("conv_"):
    kind = GROUP.split("_")[1]
    for cfg in CONVS:
        record(f"conv_{kind}_{cfg}", lambda k=kind, c=cfg: conv_case(k, c))


def stem_case(kind):
    g = torch.Generator(device="cpu").manual_seed(2)
    x = cl(torch.randn(64, 3, 32, 32, generator=g).to(dev).bfloat16())
    w = cl((torch.randn(64, 3, 7, 7, generator=g) * 0.08).to(dev).bfloat16())
    dy = cl(torch.randn(64, 64, 16, 16, generator=g).to(dev).bfloat16())
    if kind == "fwd":
        y, st = nb.conv_fwd(x, w, 2, 3, True)
        yr, sr = tb.conv_fwd(x.float(), w.float(), 2, 3, True)
        e = err(y, yr); e["stats"] = err(st, sr)
        return e
    gbuf = torch.zeros(64 * 147, device=dev)
    gv = gbuf.view(64, 7, 7, 3).permute(0, 3, 1, 2)
    nb.conv_wgrad(dy, x, w.shape, 2, 3, gv, False)
    ref = torch.zeros(64, 3, 7, 7, device=dev)
    tb.conv_wgrad(dy.float(), x.float(), w.shape, 2, 3, ref, False)
    return err(gv, ref)


if GROUP == "stem":
    record("stem_fwd", lambda: stem_case("fwd"))
    record("stem_wgrad", lambda: stem_case("wgrad"))


def bn_case(C, hw, res, relu):
    g = torch.Generator(device="cpu").manual_seed(3)
    y = cl(torch.randn(64, C, hw, hw, generator=g).to(dev).bfloat16())
    r = cl(torch.randn(64, C, hw, hw, generator=g).to(dev).bfloat16()) if res else None
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev); beta = torch.randn(C, generator=g).to(dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    rm2, rv2 = rm.clone(), rv.clone()
    o, m, i = nb.bn_act_fwd(y, None, gamma, beta, rm, rv, 0.1, 1e-5, r, relu, True)
    o2, m2, i2 = tb.bn_act_fwd(y, None, gamma, beta, rm2, rv2, 0.1, 1e-5, r, relu, True)
    e = err(o, o2); e["mean"] = err(m, m2); e["invstd"] = err(i, i2); e["rmean"] = err(rm, rm2); e["rvar"] = err(rv, rv2)
    dout = cl(torch.randn(64, C, hw, hw, generator=g).to(dev).bfloat16())
    dy, dg, db, dr = nb.bn_act_bwd(dout, o2, y, m2, i2, gamma, relu, res)
    dy2, dg2, db2, dr2 = tb.bn_act_bwd(dout, o2, y, m2, i2, gamma, relu, res)
    e["bwd_dy"] = err(dy, dy2); e["bwd_dg"] = err(dg, dg2); e["bwd_db"] = err(db, db2)
    if res:
        e["bwd_dres"] = err(dr, dr2)
    e["rel"] = max(e["rel"], e["bwd_dy"]["rel"], e["bwd_dg"]["rel"], e["bwd_db"]["rel"], e["mean"]["rel"], e["rvar"]["rel"])
    return e


for (C, hw, res, relu) in ([(64, 16, False, True), (64, 8, True, True), (128, 4, False, False), (512, 1, True, True), (256, 2, False, True)] if GROUP == "elem" else []):
    record(f"bn_{C}_{hw}_{res}_{relu}", lambda a=(C, hw, res, relu): bn_case(*a))


def pool_case():
    g = torch.Generator(device="cpu").manual_seed(4)
    x = cl(torch.randn(64, 64, 16, 16, generator=g).clamp_min(0).to(dev).bfloat16())
    (y, aux), (y2, aux2) = nb.maxpool_fwd(x, True), tb.maxpool_fwd(x, True)
    dy = cl(torch.randn(64, 64, 8, 8, generator=g).to(dev).bfloat16())
    dx = nb.maxpool_bwd(dy, aux); dx2 = tb.maxpool_bwd(dy, aux2)
    e = err(y, y2); e["bwd"] = err(dx, dx2); e["rel"] = max(e["rel"], e["bwd"]["rel"])
    return e


if GROUP == "elem":
    record("maxpool", pool_case)


def head_case(hw):
    g = torch.Generator(device="cpu").manual_seed(5)
    f = cl(torch.randn(64, 512, hw, hw, generator=g).to(dev).bfloat16())
    W = (torch.randn(10, 512, generator=g) * 0.05).to(dev); b = torch.randn(10, generator=g).to(dev) * 0.1
    lab = torch.randint(0, 10, (64,), generator=g).to(dev)
    dW, db = torch.zeros(10, 512, device=dev), torch.zeros(10, device=dev)
    dW2, db2 = torch.zeros(10, 512, device=dev), torch.zeros(10, device=dev)
    l, c, df, lg = nb.head_fwd_bwd(f, W, b, lab, 1.0, 10, dW, db, False, True)
    l2, c2, df2, lg2 = tb.head_fwd_bwd(f, W, b, lab, 1.0, 10, dW2, db2, False, True)
    e = err(l.view(1), l2.view(1)); e["correct"] = [c.item(), c2.item()]
    e["dfeat"] = err(df, df2); e["dW"] = err(dW, dW2); e["db"] = err(db, db2)
    e["rel"] = max(e["rel"], e["dfeat"]["rel"], e["dW"]["rel"], e["db"]["rel"], abs(c.item() - c2.item()))
    return e


if GROUP == "elem":
    record("head_1x1", lambda: head_case(1))
    record("head_2x2", lambda: head_case(2))


def adam_case():
    g = torch.Generator(device="cpu").manual_seed(6)
    n = 4096 * 33
    p = torch.randn(n, generator=g).to(dev); gr = torch.randn(n, generator=g).to(dev) * 0.01
    p2 = p.clone()
    m, v, m2, v2 = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    sh = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    st, st2 = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    opt = torch.optim.Adam([torch.nn.Parameter(p2)], lr=1e-3)
    for _ in range(3):
        nb.adam_step(p, gr, m, v, sh, st, 1e-3, 0.9, 0.999, 1e-8, 1.0)
        opt.param_groups[0]["params"][0].grad = gr.clone()
        opt.step()
    ref = opt.param_groups[0]["params"][0].detach()
    e = err(p - ref + 1.0, torch.ones_like(p)); e["shadow"] = err(sh, ref.bfloat16()); e["step"] = st.item()
    d = nb.grad_diff_sq(gr, torch.zeros_like(gr)); e["diff"] = [d.item(), (gr * gr).sum().item()]
    return e


if GROUP == "elem":
    record("adam", adam_case)


def allreduce_local(world, algo, wire_bf16, n):
    C = nb.C
    comms = [C.PeerComm(r, world, 0, n * 4, 32) for r in range(world)]
    C.PeerComm.link_local(comms)
    g = torch.Generator(device="cpu").manual_seed(7)
    grads = [torch.randn(n, generator=g).to(dev) for _ in range(world)]
    ref = sum(gg.bfloat16().float() / world if wire_bf16 else gg / world for gg in grads) if False else None
    ref = torch.zeros(n, device=dev)
    for gg in grads:
        t = gg * (1.0 / world)
        ref += t.bfloat16().float() if wire_bf16 else t
    streams = [torch.cuda.Stream() for _ in range(world)]
    for rep in range(3):
        work = [gg.clone() for gg in grads]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                comms[r].allreduce(work[r], algo, wire_bf16, 1.0 / world)
        torch.cuda.synchronize()
    e = {"rel": 0.0, "nan": False}
    for r in range(world):
        ee = err(work[r], ref if not (algo == "twoshot" and wire_bf16) else ref.bfloat16().float())
        e["rel"] = max(e["rel"], ee["rel"]); e["nan"] = e["nan"] or ee["nan"]
    e["identical_across_ranks"] = all(torch.equal(work[0], work[r]) for r in range(world))
    e["errflag"] = [c.error() for c in comms]
    if any(e["errflag"]):
        e["rel"] = 1.0
    return e


for world in ((2, 4) if GROUP == "comm" else ()):
    for algo in ("oneshot", "twoshot"):
        for wire in (True, False):
            record(f"allreduce_local_w{world}_{algo}_{'bf16' if wire else 'fp32'}",
                   lambda a=(world, algo, wire): allreduce_local(a[0], a[1], a[2], 1 << 20))
if GROUP == "comm":
    record("allreduce_local_w2_oneshot_small", lambda: allreduce_local(2, "oneshot", True, 512))

out["launches"] = dict(nb.LAUNCHES)
out["fallbacks"] = dict(nb.FALLBACKS)
out["n_fail"] = sum(1 for c in out["cases"] if not c["ok"])
path = sys.argv[1]
import os
os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
print(f"diag: {len(out['cases'])} cases, {out['n_fail']} failed -> {path}")
