"""Attribute ncu's per-SASS-instruction 'Instructions Executed' counts to CUDA source lines.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libcgs.so; nvdisasm --print-line-info conv_gemm.sm_100a.cubin > dis.txt
    python tools/ncu_inst_by_line.py src.csv dis.txt <mangled-kernel-substring> [index among matching launches]

The n-th instruction of the kernel in the ncu listing is the n-th instruction of the same kernel in nvdisasm.
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(dis, sub):
    out, on, cur = [], False, ("?", 0)
    for ln in open(dis):
        if ln.startswith("\t.section"):
            on = ln.startswith("\t.section\t.text.") and sub in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    src, dis, sub = sys.argv[1], sys.argv[2], sys.argv[3]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(src)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    named = [b for b in blocks if sub.split("ILi")[0] in b["name"].replace("<(int)", "ILi")] or blocks
    b = named[which % len(named)]
    ix = {h: i for i, h in enumerate(b["hdr"])}
    lines = sass_lines(dis, sub)
    print("kernel:", b["name"][:80], "| ncu instrs", len(b["data"]), "| nvdisasm instrs", len(lines))
    agg, smp = defaultdict(int), defaultdict(int)
    for r, key in zip(b["data"], lines):
        agg[key] += int(r[ix["Instructions Executed"]])
        smp[key] += int(r[ix["# Samples"]])
    tot = sum(agg.values())
    print("total warp instructions", tot)
    text = {}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:50]:
        if k[0] not in text:
            try:
                text[k[0]] = open("/root/repo/collaborative-gan-sampling_b200/csrc/" + k[0]).read().split("\n")
            except OSError:
                text[k[0]] = []
        t = text[k[0]][k[1] - 1].strip()[:90] if 0 < k[1] <= len(text[k[0]]) else ""
        print("%9d %5.1f%% smp %4d  %s:%d  %s" % (v, 100.0 * v / tot, smp[k], k[0], k[1], t))


if __name__ == "__main__":
    main()
