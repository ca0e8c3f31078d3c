"""Aggregates an ncu per-launch metric list of the dense-conv kernels of one training step
(`ncu -k regex:conv_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,...`)
into profiles/roofline_traffic.json, which bench.py reports as `roofline.traffic` (DRAM bytes per launch of the
implicit-GEMM fprop + dgrad class).

  python tools/ncu_traffic.py gpurun_out/<tag>_conv_launches.csv.gz profiles/roofline_traffic.json
"""
import collections
import csv
import gzip
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
op = gzip.open if src.endswith(".gz") else open
rows = [ln for ln in op(src, "rt") if ln.startswith('"')]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1,
         "ms": 1e3, "msecond": 1e3, "%": 1}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for x in csv.DictReader(rows):
    name = x["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(x["Metric Value"].replace(",", "")) * SCALE.get(x["Metric Unit"], 1)
    agg[name][x["Metric Name"]] += v
    if x["Metric Name"] == "gpu__time_duration.sum":
        cnt[name] += 1
out = {"kernels": {}}
for k, d in sorted(agg.items()):
    n = cnt[k]
    out["kernels"][k] = {"launches": n, "time_us_per_launch": d["gpu__time_duration.sum"] / n,
                         "dram_read_bytes_per_launch": d["dram__bytes_read.sum"] / n,
                         "dram_write_bytes_per_launch": d["dram__bytes_write.sum"] / n,
                         "l2_bytes_per_launch": d["lts__t_bytes.sum"] / n,
                         "tensor_pipe_active_pct": d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] / n}
gemm = [k for k in out["kernels"] if k.startswith(("conv_gemm", "conv3_kernel"))]
n = sum(out["kernels"][k]["launches"] for k in gemm)
tot = sum((out["kernels"][k]["dram_read_bytes_per_launch"] + out["kernels"][k]["dram_write_bytes_per_launch"]) *
          out["kernels"][k]["launches"] for k in gemm)
out["conv_gemm"] = {"launches": n, "dram_bytes_per_launch": tot / max(n, 1), "kernels": gemm,
                    "source": "ncu per-launch metrics of one eager training step (batch 32 @384x384), " + src}
out["source"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum per launch over one eager step "
                 "(tools/ncu_traffic.py, %s), %s" % (src, sys.argv[3] if len(sys.argv) > 3 else "build not named"))
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["conv_gemm"]))
