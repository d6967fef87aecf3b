"""profiles/r02_traffic.json from the ncu captures of tools/r02_profiles_final.sh: per-ray instruction counts (warp
instructions per warp of 32 rays), DRAM bytes per ray and FP64 pipe utilisation of the bench-step kernel and of the
final-slab kernel.  python tools/make_traffic_json.py gpurun_out/r2f_prof_grid.ncu-rep gpurun_out/r2f_prof_fast.ncu-rep RAYS"""
import csv
import json
import subprocess
import sys


def page(path, which):
    out = subprocess.run(["ncu", "-i", path, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarize(path, rays):
    raw = page(path, "raw")
    d = dict(zip(raw[0], raw[2]))
    num = lambda k: float(d[k].replace(",", ""))
    src = page(path, "source")
    hdr = src[1]
    i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
    total = fp64 = 0
    for r in src[2:]:
        if len(r) <= i_exec:
            continue
        n = int(r[i_exec] or 0)
        t = r[i_src].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        total += n
        fp64 += n if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") else 0
    warps = rays / 32.0
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units = dict(zip(raw[0], raw[1]))
    dram = sum(num(k) * scale[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    return {"kernel_name": src[0][1], "gpu_time_ms": num("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "msecond": 1, "usecond": 1e-3}.get(units["gpu__time_duration.sum"], 1),
            "dram_bytes_per_ray": dram / rays, "instr_per_ray": round(total / warps),
            "fp64_pipe_instr_per_ray": round(fp64 / warps), "other_instr_per_ray": round((total - fp64) / warps),
            "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active")}


if __name__ == "__main__":
    grid, fast, rays = sys.argv[1], sys.argv[2], int(float(sys.argv[3]))
    out = summarize(grid, rays)
    out = {"source": "ncu --set full, tools/r02_profiles_final.sh: the pure lean kernel (trace_lean_kernel<1,0,0,0,2>, picked by "
                     "the verdict cache) on the bench workload, %d rays x 10 surfaces" % rays,
           "rays": rays, **out, "kernel": "final slab + statistics + 2048^2 grid (the bench step)",
           "fast_kernel": {**summarize(fast, rays), "kernel": "final slab only"}}
    print(json.dumps(out, indent=1))
