"""Summarise an .ncu-rep (read offline with `ncu -i`): one row per captured launch with the metrics DESIGN.md and
profiles/README.md quote. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--csv out.csv]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clk"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "inst"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    out = [["kernel", "grid"] + [f"{short} [{units[hdr.index(m)]}]" if m in hdr else short for m, short in WANT]]
    for r in body:
        name = r[hdr.index("Kernel Name")]
        name = name.replace("void ", "").replace("vlmfw::<unnamed>::", "").replace("unnamed>::", "")[:70]
        line = [name, r[hdr.index("Grid Size")]]
        for m, _ in WANT:
            line.append(r[hdr.index(m)] if m in hdr else "")
        out.append(line)
    if len(sys.argv) > 3 and sys.argv[2] == "--csv":
        with open(sys.argv[3], "w", newline="") as f:
            csv.writer(f).writerows(out)
    for line in out:
        print(" | ".join(line))


if __name__ == "__main__":
    main()
