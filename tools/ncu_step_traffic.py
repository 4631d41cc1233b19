"""DRAM traffic of ONE training step per plan op, from an ncu CSV (no GPU needed to parse):

    NPVC_NVTX=1 ncu --nvtx --print-nvtx-rename kernel --clock-control none --csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --log-file gpurun_out/r2_step_traffic.csv python tools/ncu_target.py 16384 adam
    python tools/ncu_step_traffic.py gpurun_out/r2_step_traffic.csv profiles/ncu_traffic.json

tools/ncu_target.py runs exactly one pack + fwd + bwd (+ Adam) pass with no warm-up; with NPVC_NVTX=1 the library wraps
every plan op in an NVTX range and ncu renames the kernels after it, so a row of the CSV is a plan op.  The result
(profiles/ncu_traffic.json) feeds bench.py: roofline.traffic (the dominant op's bytes per launch) and
roofline.step_dram_bytes (the whole step)."""
import csv
import json
import sys


def main(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Name" in r)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ii = hdr.index("ID")
    launches = {}
    for r in rows:
        if r is hdr or not r[ii].isdigit():
            continue
        d = launches.setdefault(int(r[ii]), {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        unit = r[ui].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1.0)
        d[r[mi]] = v * scale
    per_op, ms_op, order = {}, {}, []
    for i in sorted(launches):
        d = launches[i]
        name = d["name"].split("(")[0].strip()
        if "/" in name:                                   # "<plan op>/<kernel function>" (NVTX rename): keep the plan op
            name = name.split("/")[0]
        else:
            name = name.replace("void ", "").replace("npvc::", "").split("<")[0]
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        if name not in per_op:
            order.append(name)
        per_op[name] = per_op.get(name, 0.0) + b
        ms_op[name] = ms_op.get(name, 0.0) + d.get("gpu__time_duration.sum", 0.0)
    out = {
        "source": src, "how": __doc__.split("\n\n")[1].strip(),
        "launches": len(launches),
        "step_dram_bytes": sum(per_op.values()),
        "step_kernel_ms_under_ncu": sum(ms_op.values()),
        "bytes_per_launch": {k: per_op[k] for k in order},
        "ms_under_ncu": {k: round(ms_op[k], 5) for k in order},
    }
    json.dump(out, open(dst, "w"), indent=1)
    print("%d launches, %.3f GB of DRAM traffic per step, %.3f ms of kernel time (serialised, cold caches)"
          % (len(launches), out["step_dram_bytes"] / 1e9, out["step_kernel_ms_under_ncu"]))
    for k in sorted(per_op, key=lambda k: -per_op[k])[:12]:
        print("  %-18s %8.1f MB  %.4f ms" % (k, per_op[k] / 1e6, ms_op[k]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
