"""Dynamic opcode mix of a kernel from `ncu --page source --csv` (needs --import-source / SASS in the report)."""
import collections
import csv
import re
import subprocess
import sys


def main(path, warps=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    i_src, i_exec, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops = collections.Counter()
    samples = collections.Counter()
    total = 0
    for r in rows[2:]:
        if len(r) <= i_exec:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
        if not m:
            continue
        n = int(r[i_exec] or 0)
        ops[m.group(1)] += n
        samples[m.group(1)] += int(r[i_samp] or 0)
        total += n
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"{path}: {total} warp instructions, FP64-pipe {fp64} ({100 * fp64 / total:.1f}%)")
    if warps:
        print(f"  per warp: total {total / warps:.0f}, FP64 {fp64 / warps:.0f}, other {(total - fp64) / warps:.0f}")
    tot_s = sum(samples.values())
    for k, v in ops.most_common(28):
        per = f"{v / warps:8.1f}/warp" if warps else ""
        print(f"  {k:12s} {v:14d} {100 * v / total:5.1f}%  {per}  stall-samples {100 * samples[k] / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
