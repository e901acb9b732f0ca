"""Build profiles/<round>_traffic.json from ncu metric passes (CSV log files) of tools/profile_layers.py.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
        -k regex:"conv_gemm_tc|edge_" --csv --log-file X_mnist.csv python tools/profile_layers.py --workload mnist --reps 0
    python tools/make_traffic_json.py out.json mnist=X_mnist.csv dcgan64=X_dcgan64.csv

Per workload and kernel family (conv_gemm_tc = tcgen05 gathered GEMM, edge = image-edge streaming kernels): launches of
one forward + one data-gradient chain at batch 1024, DRAM bytes read + written, ncu time, time-weighted tensor-pipe
activity.  bench.py reads `traffic_bytes` / `tensor_pipe_active_pct_time_weighted` of the conv_gemm_tc family for the
`roofline` object and the edge family for `roofline.edge`.
"""
import csv
import json
import sys


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr, recs = None, {}
    for r in rows:
        if "Kernel Name" in r and "Metric Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            rec = recs.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d.get("Metric Unit", "")
            if d["Metric Name"].startswith("dram__bytes"):
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            if d["Metric Name"].startswith("gpu__time"):
                v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1e-3)
            rec[d["Metric Name"]] = v
    return [recs[k] for k in sorted(recs)]


def family(name):
    return "conv_gemm_tc" if "conv_gemm_tc" in name else ("edge" if "edge_" in name else None)


def summarise(recs):
    out = {}
    for fam in ("conv_gemm_tc", "edge"):
        sel = [r for r in recs if family(r["name"]) == fam]
        if not sel:
            continue
        rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in sel)
        wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in sel)
        t = sum(r.get("gpu__time_duration.sum", 0.0) for r in sel)
        tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
        out[fam] = {
            "launches": len(sel), "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
            "ncu_time_us": round(t, 3),
            "tensor_pipe_active_pct_time_weighted": round(sum(r.get(tp, 0.0) * r.get("gpu__time_duration.sum", 0.0) for r in sel) / max(t, 1e-9), 2),
            "per_launch": [{"kernel": ("edge_wide" if "edge_wide" in r["name"] else "edge_narrow" if "edge_narrow" in r["name"] else
                                       r["name"][r["name"].find("conv_gemm_tc"):][:24]),
                            "us": round(r.get("gpu__time_duration.sum", 0.0), 1),
                            "dram_MB": round((r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)) / 1e6, 1),
                            "tensor_pct": round(r.get(tp, 0.0), 1)} for r in sel]}
    return out


def main():
    out = {}
    for arg in sys.argv[2:]:
        key, path = arg.split("=", 1)
        out[key] = summarise(parse(path))
    out["_how"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active... "
                   "--clock-control none -k regex:'conv_gemm_tc|edge_' python tools/profile_layers.py --workload {mnist,dcgan64_l1} "
                   "--reps 0 (batch 1024): every GEMM / edge launch of one forward + one data-gradient chain, layer order fwd,bwd; "
                   "tools/make_traffic_json.py")
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    for k, v in out.items():
        if k != "_how":
            print(k, {f: (d["launches"], round(d["traffic_bytes"] / 1e6, 1), d["ncu_time_us"], d["tensor_pipe_active_pct_time_weighted"]) for f, d in v.items()})


if __name__ == "__main__":
    main()
