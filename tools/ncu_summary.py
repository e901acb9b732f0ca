"""Summarise an ncu report (raw page) into a small markdown table: python tools/ncu_summary.py X.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("smsp__cycles_active.avg", "cycles")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|" + "---|" * len(WANT))
    for n, d in enumerate(data):
        name = d[ix["Kernel Name"]]
        name = name[name.find("conv_gemm"):][:28] if "conv_gemm" in name else name[:28]
        cells = []
        for key, _ in WANT:
            if key in ix:
                v, u = d[ix[key]], units[ix[key]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, u) if u not in ("", "%") else v)
            else:
                cells.append("-")
        print("| %d | %s | %s |" % (n, name, " | ".join(cells)))


if __name__ == "__main__":
    main(sys.argv[1])
