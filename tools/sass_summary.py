"""Regenerate profiles/sass/ from the objects of the current build: per kernel, the count of the SASS mnemonics that
prove which hardware paths it uses (tcgen05 / TMEM / TMA / mbarrier / cluster / multimem / peer flags), plus the
tensor-op excerpt of the convolution kernel.  CPU-only (cuobjdump).  Usage: python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "csrc", "build")
OUT = os.path.join(ROOT, "profiles", "sass")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS",
        "UCGABAR", "CGAERRBAR", "MAPA", "LDGMC", "REDG", "ATOMG", "ACQBULK", "HMMA", "STRONG.SYS", "CCTL.IVALL", "ERRBAR"]
HEAD = """# SASS evidence — `cuobjdump -sass csrc/build/*.cu.o` (sm_100a), per kernel (regenerate: `python tools/sass_summary.py`)

| SASS | PTX it comes from |
|---|---|
| `UTCHMMA` | `tcgen05.mma.cta_group::1.kind::f16` |
| `LDTM` | `tcgen05.ld.32x32b.x32` |
| `UTMALDG.2D/.4D` | `cp.async.bulk.tensor.{2d,4d}` (TMA) |
| `UTCBAR` | `tcgen05.commit…mbarrier::arrive::one` |
| `UTCATOMSWS` | `tcgen05.alloc / dealloc / relinquish_alloc_permit` |
| `SYNCS.*` | `mbarrier.*` |
| `UCGABAR_ARV / UCGABAR_WAIT`, `MAPA`, `LD.E.128` on a mapped address | `barrier.cluster.arrive/wait`, `mapa.shared::cluster`, `ld.shared::cluster` (cluster split-K over DSMEM) |
| `ACQBULK` | `griddepcontrol.wait` (programmatic dependent launch) |
| `LDGMC.E.*ADD*` | `multimem.ld_reduce` (NVLS in-switch reduction); the matching `multimem.st` is `STG.E.128.STRONG.SYS` on the multicast address |
| `LDG/STG.E.STRONG.SYS` | `ld.acquire.sys` / `st.release.sys` peer flags |
| `REDG.E.ADD.F32` | `red.global.add(.v4).f32` (BN-statistics / wgrad atomics) |
| `STG.E.128.STRONG.SYS` / `LDG.E.128.STRONG.SYS` in `igemm_tp_kernel`, `tp_head_kernel`, `tp_allreduce_bf16_kernel`, `allreduce_kernel<…,3,…>` | `st.volatile.global.v4` / `multimem.st` / `ld.volatile.global.v4`: the flag-in-data ("LL") words {data, epoch, data, epoch} pushed into every peer and polled locally |
| `REDG.E.ADD.STRONG.SYS` | `red.release.sys.global.add.u32` / `multimem.red.release.sys.add.u32`: tile arrival counters of the bandwidth protocol |
| `UBLKCP.S.G` | `cp.async.bulk.shared::cluster.global`: peers' partial tiles pulled straight into the idle operand ring |

No `HMMA` (legacy `mma.sync`) appears in any kernel.
"""


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = [HEAD]
    excerpt = []
    for obj in sorted(f for f in os.listdir(BUILD) if f.endswith(".cu.o")):
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        for part in re.split(r"\n\s*Function : ", txt)[1:]:
            name, _, body = part.partition("\n")
            name = name.strip()
            ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", body)
            cnt = collections.Counter()
            for i in ins:
                for k in KEYS:
                    if k in i:
                        cnt[i] += 1
                        break
            if not cnt:
                continue
            lines.append(f"### {obj[:-2]} — `{name}`\n{len(ins)} SASS instructions: " +
                         ", ".join(f"`{k}`×{v}" for k, v in sorted(cnt.items())) + "\n")
            if "igemm_kernelILi64ELb0" in name:
                for ln in body.splitlines():
                    if re.search(r"UTCHMMA|LDTM|UTMALDG|UTCBAR|UTCATOMSWS|UCGABAR|MAPA|ACQBULK|SYNCS", ln):
                        excerpt.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
    open(os.path.join(OUT, "SASS_SUMMARY.md"), "w").write("\n".join(lines))
    open(os.path.join(OUT, "conv_gemm_tensor_ops.sass"), "w").write(
        "// tensor-core / TMEM / TMA / mbarrier / cluster instructions of hz::igemm_kernel<64,false> (forward conv)\n" +
        "\n".join(excerpt) + "\n")
    print("wrote", OUT, len(lines) - 1, "kernels")


if __name__ == "__main__":
    main()
