"""Per-kernel-family DRAM bytes of one headline step from an ncu launch list that captured
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum:
    python scripts/dram_traffic.py profiles/r1_cXX_launches.csv > profiles/r1_dram_traffic.json
(bench.py reports these as roofline.traffic; they are per step = 6 iterations.)"""
import csv, json, sys
from collections import defaultdict

FAMILIES = [("corr_sample", ("corr_sample_tc_kernel", "corr_patch_tc_kernel", "corr_patch_t_kernel", "corr_sample_simt_kernel")),
            ("encoder", ("conv3x3_tc_kernel", "conv_stem_kernel", "norm_act_kernel", "instnorm_", "gather_s2_kernel",
                         "upsample_concat", "l2norm_rows", "avgpool2")),
            ("gemm", ("gemm_split3", "gemm_qkv_time_attn")),
            ("layernorm", ("layernorm_split_kernel",)),
            ("attention", ("attention_tc_kernel", "attention_combine_kernel", "attention_simt", "attn_p2v"))]
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
SECS = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
out = defaultdict(lambda: {"ids": set(), "dram_read": 0.0, "dram_write": 0.0, "ncu_time_s": 0.0})
for r in rows:
    fam = next((f for f, keys in FAMILIES if any(k in r["Kernel Name"] for k in keys)), None)
    if fam is None:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    o = out[fam]
    o["ids"].add(r["ID"])
    if r["Metric Name"] == "dram__bytes_read.sum":
        o["dram_read"] += v * BYTES[r["Metric Unit"]]
    elif r["Metric Name"] == "dram__bytes_write.sum":
        o["dram_write"] += v * BYTES[r["Metric Unit"]]
    elif r["Metric Name"] == "gpu__time_duration.sum":
        o["ncu_time_s"] += v * SECS[r["Metric Unit"]]
res = {}
for fam, o in out.items():
    res[fam] = {"launches_per_step": len(o["ids"]), "dram_bytes_per_step": o["dram_read"] + o["dram_write"],
                "dram_read": o["dram_read"], "dram_write": o["dram_write"], "ncu_time_s": o["ncu_time_s"]}
res["_how"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one "
               "headline step (scripts/profile_step.py), N=6400 T=16; summed per kernel family by scripts/dram_traffic.py "
               "from " + sys.argv[1])
print(json.dumps(res, indent=1))
