"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the three hash launches of one training step, out of an `ncu --set full` capture of
`bench.py` (read here, no GPU) -> profiles/r2_ncu_traffic.json, which bench.py reports as `roofline.traffic` (so the figure is a measurement of a
committed capture, not a literal in the bench).
    python scripts/ncu_traffic.py gpurun_out/x/prof_hash.ncu-rep profiles/r2_ncu_traffic.json"""
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rep, dst = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else None      # use only the which-th captured step (e.g. skip the graph capture's warm-up on placeholder rays)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__grid_size")}
fwd, bwd = [], []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    by = sum(float(r[col[k]].replace(",", "")) * UNIT[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    rec = {"bytes": by, "grid": int(r[col["launch__grid_size"]].replace(",", "")), "us": float(r[col["gpu__time_duration.sum"]].replace(",", ""))}
    if "hash_fwd_kernel" in name:
        fwd.append(rec)
    elif "hash_bwd_kernel" in name:
        bwd.append(rec)
# a step has two forward launches: the coarse one is the smaller grid
steps = []
for i in range(0, len(fwd) - 1, 2):
    a, b = sorted(fwd[i:i + 2], key=lambda x: x["grid"])
    steps.append((a, b))
if which is not None:
    steps, bwd = steps[which:which + 1], bwd[which:which + 1]
res = {"source": rep, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, cold-cache replays)"}
if steps:
    res["hash_encode_fwd_coarse"] = sum(s[0]["bytes"] for s in steps) / len(steps)
    res["hash_encode_fwd_fine"] = sum(s[1]["bytes"] for s in steps) / len(steps)
if bwd:
    res["hash_encode_bwd"] = sum(b["bytes"] for b in bwd) / len(bwd)
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res))
