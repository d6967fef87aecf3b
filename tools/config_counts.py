"""Parse tools/config_counts.sh's ncu log: per config, the launches of its LAST call (pure kernels), summed."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ix = {h: j for j, h in enumerate(hdr)}
launches = {}
for r in rows[start + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    if r[ix["Metric Name"]] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ix["Metric Unit"]], 1.0)
    d[r[ix["Metric Name"]]] = v
# trace launches (not probes: PROBE is the 4th template argument) in order
def is_trace(name):
    args = [a.strip() for a in name.split("trace_lean_kernel<")[1].split(">")[0].split(",")]
    return args[3] == "0"


traces = [d for _, d in sorted(launches.items()) if is_trace(d["name"])]
# run_configs --configs 2 3 4 5 --repeat 2: every config = 2 x timed() = 2 x (warm call + timed call); launches per call:
# config 2: 1, config 3: 1, config 4: 2 (2^27 rays + the rest), config 5: 1
per_call = {"config2": 1, "config3": 1, "config4": 2, "config5_achromat": 1}
rays = {"config2": 3 * 4096 ** 2, "config3": 32 * 2048 ** 2, "config4": 16001 * 16000, "config5_achromat": 11181 * 11180}
out, pos = {}, 0
for name, n in per_call.items():
    calls = [traces[pos + k * n: pos + (k + 1) * n] for k in range(4)]
    pos += 4 * n
    last = calls[-1]
    fp64 = sum(d["smsp__inst_executed_pipe_fp64.sum"] for d in last)
    total = sum(d["smsp__inst_executed.sum"] for d in last)
    out[name] = {"kernel": last[0]["name"], "rays": rays[name], "fp64_pipe_instr_per_ray": fp64 / (rays[name] / 32.0),
                 "instr_per_ray": total / (rays[name] / 32.0), "ncu_ms": sum(d["gpu__time_duration.sum"] for d in last)}
print(json.dumps({"source": "tools/config_counts.sh: ncu smsp__inst_executed_pipe_fp64.sum / smsp__inst_executed.sum of the lean trace "
                            "kernels (pure instantiations) of each config's last call, per warp of 32 rays", **out}, indent=1))
