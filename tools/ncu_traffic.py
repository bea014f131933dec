"""profiles/ncu_traffic.json from one `ncu --set full` capture: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per
launch of the kernels behind bench.py's launch groups, stamped with the commit. bench.py reads the file for
`roofline.traffic` (it cannot run ncu on itself).

  # on the GPU box (plain run first, then the capture; the persistent cluster kernel is profiled as a plain launch)
  DPI_B200_SMALL_COOP=0 ncu --set full --clock-control none -k regex:'flow_small_kernel|clip_adam_kernel' -c 6 \
      -o gpurun_out/c1_traffic python bench.py --no-extra --no-graph --steps 1 --warmup 3 --cpu-steps 1
  # here
  ncu -i gpurun_out/c1_traffic.ncu-rep --page raw --csv > gpurun_out/c1_traffic.csv
  python tools/ncu_traffic.py c1 gpurun_out/c1_traffic.csv flow_reverse_fwd=flow_small_kernel:0 \
      flow_reverse_bwd=flow_small_kernel:1 clip_adam=clip_adam_kernel:0
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "KB": 1e3, "MB": 1e6, "GB": 1e9, "B": 1}.get(u, 1)


def main():
    workload, path = sys.argv[1], sys.argv[2]
    maps = dict(a.split("=") for a in sys.argv[3:])
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    head, units, data = rows[0], rows[1], rows[2:]
    kcol = head.index("Kernel Name")
    rcol, wcol = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
    tcol = head.index("gpu__time_duration.sum") if "gpu__time_duration.sum" in head else None
    seen, per = {}, {}
    for r in data:
        name = r[kcol]
        for key in set(v.split(":")[0] for v in maps.values()):
            if key in name:
                i = seen.get(key, 0)
                seen[key] = i + 1
                per[f"{key}:{i}"] = dict(
                    bytes=float(r[rcol].replace(",", "")) * unit_scale(units[rcol]) + float(r[wcol].replace(",", "")) * unit_scale(units[wcol]),
                    ns=(float(r[tcol].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(units[tcol], 1)) if tcol is not None else None)
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    out["commit"] = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT).decode().strip()
    out[workload] = {g: int(per[k]["bytes"]) for g, k in maps.items() if k in per}
    out.setdefault("detail", {})[workload] = {g: per[k] for g, k in maps.items() if k in per}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out[workload]))


if __name__ == "__main__":
    main()
