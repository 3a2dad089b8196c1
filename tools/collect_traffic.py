"""Per-workload DRAM traffic of the hot kernels (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch) from the
csv launch lists tools/gpu_profile_r02.sh writes -> profiles/top_kernel_traffic.json, keyed by workload then by bench.py's
kernel name.  The upsampler / 300M captures run at a reduced batch; traffic of these kernels is linear in the number of
sequences, so it is scaled to the workload's batch (factor recorded)."""
import csv
import json
import os
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"base40M-imagevec-1024pt-b64": 1.0, "upsample-4096pt-b128": 1.0, "base300M-upsample-4096pt-b64": 4.0}


def kernel_key(name):
    if "attn_" in name:
        return "flash_attention"
    if "cast_rowstats" in name:
        return "cast_rowstats"
    m = re.search(r"gemm_bf16_tc2_kernel<(?:\(int\))?(\d+), (?:\(bool\))?(\d), (?:\(bool\))?(\d)>", name)
    if m:  # <epilogue, bf16 output, deep-K>: 4 = c_qkv, 5 = c_fc + GELU, 3 = residual + statistics (deep-K: mlp.c_proj)
        epi, _, deepk = int(m.group(1)), int(m.group(2)), int(m.group(3))
        return {4: "gemm_qkv_lnfold", 5: "gemm_fc1_lnfold_gelu"}.get(epi, "gemm_fc2_resid_stats" if deepk else "gemm_attn_proj_resid_stats")
    return None


out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, inside a real denoiser forward), keyed by "
                   "workload then bench.py kernel name; written by tools/collect_traffic.py from tools/gpu_profile_r02.sh"}
for wl, scale in SCALE.items():
    path = os.path.join(src, f"traffic_{wl}.csv")
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None:
        continue
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows:
        if r is hdr or not r[idi].isdigit():
            continue
        d = launches.setdefault(int(r[idi]), {"name": r[ki]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    per = {}
    for i in sorted(launches):
        d = launches[i]
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        key = kernel_key(d["name"])
        if key and key not in per:
            per[key] = b * scale
    out[wl] = dict(per, _scaled_by=scale)
    print(wl, {k: (round(v / 1e6, 1) if isinstance(v, float) else v) for k, v in per.items()}, "MB per launch")
json.dump(out, open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json"), "w"), indent=1)
