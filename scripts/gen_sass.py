#!/usr/bin/env python
"""Regenerate docs/sass/: one `cuobjdump -sass` listing per kernel of accl_b200/_C*.so (xz-compressed above 256 KB:
k_call and k_engine inline every collective body for every dtype and run to tens of MB of text) plus
docs/sass/HOT_LOOPS.md — the inner loops that matter, cut out of those listings around the instructions that prove
the Blackwell / NVLink code paths (LDGMC = multimem.ld_reduce, multimem stores / reds, system-scope flag accesses,
UTCHMMA[.2CTA] = tcgen05.mma, UTMALDG / UTMAREDG = TMA load / reduce-add, LDTM = tcgen05.ld).

  python scripts/gen_sass.py            # needs the built extension; no GPU
"""
import glob
import lzma
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "docs", "sass")

# (title, kernel file stem, regex of the anchor instruction, lines before, lines after, comment)
EXCERPTS = [
    ("NVLS two-shot all-reduce: multimem.ld_reduce -> multimem.st (fp32, 8 x 16 B in flight per thread)", "k_call",
     r"LDGMC\.E\.ADD\.F32x4", 4, 40,
     "`LDGMC.E.ADD.F32x4.RN.STRONG.SYS` is `multimem.ld_reduce.relaxed.sys.global.add.v4.f32`: the NVSwitch returns the sum over "
     "all ranks' copies of the addressed 16 bytes; the stores that follow go to the multicast address (`multimem.st`), i.e. "
     "the reduced shard lands in every rank's result buffer.  No SM arithmetic on this path."),
    ("NVLS bf16 reduction with fp32 accumulation in the switch", "k_call", r"LDGMC\.E\.HPADD\.BF16x8", 2, 12,
     "`multimem.ld_reduce ... .add.acc::f32.v4.bf16x2`."),
    ("LL exchange: 8 payload bytes + 8 flag bytes per 16-byte store, consumer spins on the data", "k_call",
     r"ST\.E\.128\.STRONG\.SYS", 6, 24,
     "`st.volatile.global.v2.u64` to the peer's staging region (each 64-bit half = payload word | sequence number << 32) and the "
     "matching `LD.E.128.STRONG.SYS` poll loop on the local staging region: no fence, no separate flag."),
    ("system-scope release / acquire flags of the rendezvous meetings", "k_call", r"ST\.E\.STRONG\.SYS", 8, 16,
     "`st.release.sys` (MEMBAR.ALL.SYS + ST.E.STRONG.SYS) on the peer's sync pad and the `LD.E.STRONG.SYS` acquire spin."),
    ("persistent engine: command fetch / parked-call loop of the control CTA", "k_engine", r"NANOSLEEP", 20, 20,
     "the control CTA polls the command ring with `ld.acquire.sys`, steps every call in flight and backs off with NANOSLEEP "
     "only after 100 us without work."),
    ("tcgen05 GEMM, CTA pair: MMA issue loop", "k_plugin_gemm_rs_true_false", r"UTCHMMA\.2CTA", 6, 30,
     "`tcgen05.mma.cta_group::2.kind::f16` (M = 256 across the two CTAs of the cluster) issued by one elected thread from "
     "shared-memory descriptors; `UTCBAR.2CTA.MULTICAST` = `tcgen05.commit ... multicast::cluster` releasing the smem stage in both CTAs."),
    ("tcgen05 GEMM: TMA loads", "k_plugin_gemm_rs_true_false", r"UTMALDG\.2D\.2CTA", 4, 10,
     "`cp.async.bulk.tensor.2d.cta_group::2 ... mbarrier::complete_tx::bytes` signalling the leader's barrier."),
    ("tcgen05 GEMM epilogue: TMEM -> registers -> swizzled smem -> TMA reduce-add into the owner rank's shard", "k_plugin_gemm_rs_true_false",
     r"UTMAREDG\.2D\.ADD", 30, 6,
     "`LDTM.x32` = `tcgen05.ld.32x32b.x32`; `UTMAREDG.2D.ADD` = `cp.reduce.async.bulk.tensor.2d ... .add`: the tile is added into "
     "peer memory over NVLink by the TMA unit (device API `reduce_scatter_emit_tile`)."),
    ("tcgen05 GEMM, single CTA: MMA issue loop", "k_plugin_gemm_rs_false_false", r"UTCHMMA ", 6, 24, "`tcgen05.mma.cta_group::1.kind::f16`."),
]


def main():
    so = glob.glob(os.path.join(ROOT, "accl_b200", "_C*.so"))
    if not so:
        sys.exit("build the extension first (python -m accl_b200.utils.build)")
    txt = subprocess.run(["cuobjdump", "-sass", so[0]], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"(?=\t\tFunction : )", txt)
    os.makedirs(OUT, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, "*.sass*")):
        os.remove(f)
    listing = {}
    index = []
    for p in parts[1:]:
        name = p.split("\n")[0].split(": ")[1]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        stem = re.sub(r"\(.*", "", dem).replace("accl::cuda::", "").replace("void ", "")
        stem = stem.replace("<", "_").replace(">", "").replace(", ", "_").replace("(bool)1", "true").replace("(bool)0", "false")
        listing[stem] = p
        ninstr = len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@!]", p))
        if len(p) > (256 << 10):
            path = os.path.join(OUT, stem + ".sass.xz")
            with lzma.open(path, "wt", preset=6) as fh:
                fh.write(p)
        else:
            path = os.path.join(OUT, stem + ".sass")
            with open(path, "w") as fh:
                fh.write(p)
        ops = {}
        for m in re.finditer(r"\b(LDGMC[.\w]*|UTCHMMA[.\w]*|UTMALDG[.\w]*|UTMAREDG[.\w]*|UTCBAR[.\w]*|LDTM[.\w]*|UTCATOMSWS[.\w]*|"
                             r"REDG?\.E\.[.\w]*STRONG\.SYS|ST\.E\.128\.STRONG\.SYS|LD\.E\.128\.STRONG\.SYS|MEMBAR\.ALL\.SYS)", p):
            ops[m.group(1)] = ops.get(m.group(1), 0) + 1
        index.append((stem, dem, os.path.basename(path), ninstr, ops))
    with open(os.path.join(OUT, "README.md"), "w") as fh:
        fh.write("# SASS listings (cuobjdump -sass of accl_b200/_C*.so, sm_100a)\n\n"
                 "Regenerate with `python scripts/gen_sass.py`; read the compressed ones with `xz -dc <file> | less`.\n"
                 "`HOT_LOOPS.md` shows the inner loops.\n\n| kernel | listing | instructions | opcodes of interest |\n|---|---|---|---|\n")
        for stem, dem, path, n, ops in sorted(index):
            o = ", ".join(f"`{k}` x{v}" for k, v in sorted(ops.items()))
            fh.write(f"| `{dem[:110]}` | [{path}]({path}) | {n} | {o} |\n")
    with open(os.path.join(OUT, "HOT_LOOPS.md"), "w") as fh:
        fh.write("# Hot loops, cut out of the listings next to this file\n\n")
        for title, stem, rx, before, after, note in EXCERPTS:
            p = listing.get(stem)
            if p is None:
                continue
            lines = [ln for ln in p.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln)]
            idx = next((i for i, ln in enumerate(lines) if re.search(rx, ln)), None)
            fh.write(f"## {title}\n\n{note}\n\n")
            if idx is None:
                fh.write("(instruction not found in this build)\n\n")
                continue
            fh.write(f"`{stem}`:\n\n```\n")
            for ln in lines[max(0, idx - before): idx + after]:
                fh.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip() + "\n")
            fh.write("```\n\n")
    print("wrote", len(index), "listings to", OUT)


if __name__ == "__main__":
    main()
