"""Print the key figures of a bench.py JSON line.  usage: python tools/show_bench.py gpurun_out/bench.json"""
import json
import sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.1f}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
for k, v in d.get("stages", {}).items():
    print(f"  {k:20s} {v['ms']:8.2f} ms  hbm_frac {v.get('hbm_frac', 0):.3f}  fp32_frac {v.get('fp32_frac', 0):.3f}  {v.get('write_only_frac', '')}")
r = d.get("roofline", {})
print("  roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
for k in ("e2e", "e2e_planes_to_host", "single_clip_30s", "extras", "final_gather_logmel", "cpu_baseline", "strong_scaling", "host_copy"):
    if k in d:
        print(" ", k, d[k])
