"""Condenses an .ncu-rep (read with `ncu -i ... --page raw --csv`, no GPU needed) into the per-launch
figures the roofline discussion uses.  usage: python tools/ncu_summary.py REP [label ...]"""
import csv, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def main():
    rep, labels = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for k, r in enumerate(rows[2:]):
        name = r[kn].split("(")[0]
        print("launch %d: %s%s" % (k, name, ("   [%s]" % labels[k]) if k < len(labels) else ""))
        to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = {}
        for metric, label in WANT:
            if metric in hdr:
                i = hdr.index(metric)
                vals[metric] = (r[i], units[i])
                print("    %-26s %s %s" % (label, r[i], units[i]))
        try:
            dur = float(vals["gpu__time_duration.sum"][0]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[vals["gpu__time_duration.sum"][1]]
            dr = float(vals["dram__bytes_read.sum"][0]) * to_bytes[vals["dram__bytes_read.sum"][1]]
            dw = float(vals["dram__bytes_write.sum"][0]) * to_bytes[vals["dram__bytes_write.sum"][1]]
            xb = float(vals["l1tex__m_xbar2l1tex_read_bytes.sum"][0]) * to_bytes[vals["l1tex__m_xbar2l1tex_read_bytes.sum"][1]]
            print("    %-26s %.0f bytes/launch = %.2f TB/s" % ("dram traffic", dr + dw, (dr + dw) / dur / 1e12))
            print("    %-26s %.2f TB/s" % ("L2 -> SM rate", xb / dur / 1e12))
        except Exception:
            pass


if __name__ == "__main__":
    main()
