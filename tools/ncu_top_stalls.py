"""Summarise `ncu -i X.ncu-rep --page source --csv` output: hottest SASS instructions with their stall reasons."""
import csv
import sys


def main(path, topn=30):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for b in blocks:
        hdr, data = b["hdr"], b["data"]
        ix = {h: i for i, h in enumerate(hdr)}
        s_i = ix["# Samples"]
        tot = sum(int(r[s_i]) for r in data)
        print("==", b["name"][:90], "total samples", tot)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
        print("   overall:", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
        for r in sorted(data, key=lambda r: -int(r[s_i]))[:topn]:
            st = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print("%7s  %-72s %s" % (r[s_i], r[ix["Source"]].strip()[:72], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
