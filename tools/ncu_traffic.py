#!/usr/bin/env python
"""Summarise `ncu --set full` captures (here, no GPU needed): per launch the duration, DRAM bytes, pipe activity.
   python tools/ncu_traffic.py gpurun_out/r2p           -> profiles/r2_ncu_<name>.txt and profiles/ncu_traffic.json
bench.py reads profiles/ncu_traffic.json for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch
of the dominant kernel)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
           "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
           "launch__registers_per_thread", "launch__cluster_size"]


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h == "Kernel Name":
                d["kernel"] = r[i].split("(")[0].replace("void ", "")
            elif h in METRICS:
                try:
                    d[h] = (float(r[i].replace(",", "")), units[i])
                except ValueError:
                    pass
        res.append(d)
    return res


def to_bytes(v):
    val, unit = v
    return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    src = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
    traffic = {}
    for name in sorted(f[:-8] for f in os.listdir(src) if f.endswith(".ncu-rep")):
        ls = launches(os.path.join(src, name + ".ncu-rep"))
        path = os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.txt")
        with open(path, "w") as f:
            f.write(f"# ncu --set full --clock-control none, python tools/ncu_target.py {name} "
                    f"(from {src}/{name}.ncu-rep, `ncu -i ... --page raw --csv`)\n")
            for i, d in enumerate(ls):
                f.write(f"\nlaunch {i}: {d.get('kernel')}\n")
                for m in METRICS:
                    if m in d:
                        f.write(f"  {m} [{d[m][1]}] = {d[m][0]:.6g}\n")
        print("wrote", path, len(ls), "launches")
        for d in ls:
            if name == "swiglu" and "256, 3, 1>" in d.get("kernel", ""):
                tot = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
                traffic["gemm_gateup_4096x16384x2048"] = {
                    "dram_bytes": tot, "read": to_bytes(d["dram__bytes_read.sum"]),
                    "write": to_bytes(d["dram__bytes_write.sum"]), "duration_us": d["gpu__time_duration.sum"][0],
                    "tensor_pipe_active_pct": d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0],
                    "algorithmic_bytes": 4096 * 2048 * 2 + 16384 * 2048 * 2 + 4096 * 16384 * 2 + 4096 * 8192 * 2,
                    "source": f"profiles/{tag}_ncu_swiglu.txt (fused w1|w3 GEMM + SwiGLU, gemm_tc_kernel<.., 256, EPI_SWIGLU_FWD, "
                              "pair>): dram__bytes_read.sum + dram__bytes_write.sum of one launch"}
    if traffic:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
