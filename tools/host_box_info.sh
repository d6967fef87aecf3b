#!/bin/bash
# What the host side of this GPU box looks like (NUMA layout, GPU affinity, PCIe copy rates): context for e2e numbers.
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"
nvidia-smi topo -m 2>/dev/null | head -20
python - <<'PY'
import os
print("affinity of this process:", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:4], "...")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (x >> b) & 1]
    print("NVML ideal cpu affinity of GPU 0:", len(cpus), "cpus", cpus[:4], "...", cpus[-4:])
    try:
        print("NUMA node of GPU 0:", pynvml.nvmlDeviceGetNumaNodeId(h))
    except Exception as e:
        print("numa id n/a", e)
except Exception as e:
    print("pynvml n/a:", e)
PY
python tools/pcie_probe.py
