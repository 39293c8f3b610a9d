"""Dev tool: aggregate an ncu CSV of one eager UNet call (scripts/one_unet_call.py; metrics gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum) per kernel -> profiles/<tag>_unet_call_dram_traffic.json.  bench.py reads
that file for `roofline.traffic` (DRAM bytes per launch of the dominant kernel, from this one capture)."""
import collections, csv, json, sys
from pathlib import Path
src, dst = Path(sys.argv[1]), Path(sys.argv[2])
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0,
              "usecond": 1e3, "msecond": 1e6, "second": 1e9}
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("emote::", "").replace("void ", "")
    fam = k.split("<")[0]
    v = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r["Metric Unit"], 1.0)
    for key in (k, "family:" + fam):
        a = agg.setdefault(key, {"launches": 0, "time_ns": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            a["launches"] += 1
            a["time_ns"] += v
        elif m == "dram__bytes_read.sum":
            a["dram_read_bytes"] += v
        elif m == "dram__bytes_write.sum":
            a["dram_write_bytes"] += v
gemm = {"launches": 0, "time_ns": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0}
for k, a in agg.items():
    if k.startswith("family:gemm"):
        for f in gemm:
            gemm[f] += a[f]
out = {"command": "ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                  "dram__bytes_write.sum python scripts/one_unet_call.py   (one eager UNet3D call, config #2)",
       "all_gemm_kernels": dict(gemm, dram_bytes_per_launch=(gemm["dram_read_bytes"] + gemm["dram_write_bytes"]) / max(1, gemm["launches"])),
       "total_dram_bytes_one_unet_call": sum(a["dram_read_bytes"] + a["dram_write_bytes"] for k, a in agg.items() if not k.startswith("family:")),
       "total_time_ns_serialised": sum(a["time_ns"] for k, a in agg.items() if not k.startswith("family:")),
       "per_kernel": {k: a for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_ns"]) if not k.startswith("family:")}}
dst.write_text(json.dumps(out, indent=1))
print(json.dumps({k: v for k, v in out.items() if k != "per_kernel"}, indent=1))
