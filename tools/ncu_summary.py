"""Per-launch summary table of an .ncu-rep (ncu --set full): duration, DRAM bytes, pipe utilisation."""
import csv, io, subprocess, sys

KEYS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf")]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), n) for k, n in KEYS if k in hdr]
    print(" | ".join(n for _, n in idx))
    for r in rows[2:]:
        vals = []
        for i, n in idx:
            v = r[i]
            if n == "kernel":
                v = v.split("(")[0].replace("nrc::", "")[:28]
            elif n == "us":
                u = units[i]
                f = float(v)
                v = f"{f/1000:.1f}" if u in ("ns", "nsecond") else f"{f:.1f}"
            elif n in ("dram_rd", "dram_wr", "l2_bytes"):
                f = float(v); u = units[i]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                v = f"{f*mult/1e6:.2f}MB"
            else:
                try:
                    v = f"{float(v):.1f}"
                except ValueError:
                    pass
            vals.append(v)
        print(" | ".join(vals))

if __name__ == "__main__":
    main(sys.argv[1])
