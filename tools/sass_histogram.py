#!/usr/bin/env python
"""Opcode histogram of the tcgen05 / TMA kernels of libnvae_b200.so -> profiles/r02_sass_conv_tc.md (runs without a GPU).
usage: python tools/sass_histogram.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "STTM", "LDTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "UTCCP", "HMMA")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_conv_tc.md")
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nvae_tf_b200", "libnvae_b200.so")], capture_output=True,
                         text=True, check=True).stdout
    out = ["# SASS opcode histogram of the tcgen05 / TMA kernels in libnvae_b200.so (sm_100a), round 2\n",
           "Produced by `python tools/sass_histogram.py` (`cuobjdump -sass`, CUDA 12.9) from the committed sources.  Mnemonics that",
           "prove the Blackwell path: `UTCHMMA` / `UTCHMMA.2CTA` (tcgen05.mma, kind::tf32 and kind::f16), `UTMALDG.{2D,4D,5D}` (TMA",
           "tensor loads), `STTM` / `LDTM` (tcgen05.st / tcgen05.ld: TMEM stores of the split A operand, accumulator reads), `UTCBAR`",
           "(tcgen05.commit), `SYNCS.*` (mbarrier), `UTCATOMSWS` (TMEM allocation).\n"]
    hmma = 0
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ops = collections.Counter(m.group(1) for m in re.finditer(
            r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", f, re.M))
        hmma += sum(n for o, n in ops.items() if o.startswith("HMMA"))
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(conv_tc_kernel<[^>]*>)", dem)
        if not m:
            continue
        out.append(f"## `{m.group(1)}`  ({sum(ops.values())} instructions)\n")
        out.append("| mnemonic | count |\n|---|---:|")
        out += [f"| `{op}` | {n} |" for op, n in sorted(ops.items(), key=lambda kv: (-kv[1], kv[0])) if op.split(".")[0] in KEEP]
        out.append("")
    out.append(f"`HMMA` (mma.sync) instructions in the whole library: {hmma}.")
    open(out_path, "w").write("\n".join(out) + "\n")
    print(f"wrote {out_path}")


if __name__ == "__main__":
    main()
