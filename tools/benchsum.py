"""Human-readable summary of a bench.py JSON line: tools/benchsum.py gpurun_out/bench.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"{d['config']['workload']}: {d['value']:.2f} {d['unit']} (e2e {d['e2e']['value']:.2f}), {d['ms_per_step']:.1f} ms/step, "
      f"{d['n_gpus']} GPU, clocks {d.get('clocks')}")
r = d.get("roofline", {})
print("roofline:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "peak_source"})
for k, v in d.get("kernels", {}).items():
    print(f"  {k:28s} {v['ms']*1e3:8.1f} us  {v['achieved']:8.1f} {v['unit']:8s} frac {v['frac']:.3f}  "
          f"share {v.get('share_of_step', 0):.3f}  in-step {v.get('in_step_ms', 0)*1e3:7.1f} us")
for o in d.get("other_workloads", []):
    if "error" in o:
        print("  OTHER", o["workload"], "ERROR", o["error"])
    elif "value" in o:
        dk = o.get("dominant_kernel", {})
        print(f"  OTHER {o['workload']:32s} {o['value']:8.3f} {o['unit']} (e2e {o['e2e']['value']:.3f}) batch {o['per_gpu_batch']}/gpu "
              f"{o['ms_per_step']/1e3:6.1f} s/step whole-step frac {o['whole_step']['frac']:.3f}; dominant {dk.get('kernel')} "
              f"frac {dk.get('frac')}; cpu {o.get('cpu_baseline', {}).get('value')}; wall {o.get('bench_wall_s')} s")
    else:
        print("  OTHER", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in o.items() if k != "note"})
print("cpu_baseline:", d.get("cpu_baseline"))
