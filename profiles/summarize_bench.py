"""Prints a compact summary of bench.py JSON lines read from stdin (helper for profiles/)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print(f"reference: {d['value']:.4g} {d['unit']} cores={d['cpu_baseline']['cores']} sample={d['cpu_baseline']['sample']}")
        continue
    print(f"n_gpus={d['n_gpus']} value={d['value'] / 1e9:.3f} G point-updates/s  ms/step={d['ms_per_step']:.3f}  "
          f"e2e={d['e2e']['value'] / 1e9:.3f} G  launches={d['gpu_launches']}  clocks={d['clocks']}")
    print("  kernel ms/step:", {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    r = d["roofline"]
    print(f"  roofline: {r['kernel']} {r['achieved']:.0f}/{r['peak']:.0f} GB/s = {r['frac']:.3f}; whole iteration "
          f"{r['whole_iteration']['achieved']:.0f} GB/s = {r['whole_iteration']['frac']:.3f}")
    print("  setup:", d["config"].get("setup_s"))
    if "cpu_baseline" in d:
        print("  cpu_baseline:", d["cpu_baseline"])
