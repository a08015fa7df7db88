"""Compact per-launch summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_hmma_inst_pct", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("ipc", "sm__inst_executed.avg.per_cycle_elapsed"),
    ("smem_B", "launch__shared_mem_per_block_dynamic"),
]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("# " + rep)
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.split("(")[0].split("::")[-1][:48]
    out = [f"{name:48s}"]
    for label, key in KEYS:
        hits = [h for h in hdr if h.endswith(key)]
        if not hits:
            continue
        i = idx[hits[0]]
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            continue
        u = units[i]
        if label.endswith("_MB") or label == "time_us":
            v *= UNIT.get(u, 1.0)
        out.append(f"{label}={v:.2f}" if isinstance(v, float) and v != int(v) else f"{label}={int(v)}")
    print(" ".join(out))
