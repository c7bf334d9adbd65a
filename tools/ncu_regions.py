"""Instruction / stall share of an ncu report per code region (a region = a named line range of a source file).

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep k_wave
Regions are read from the `// @region name` markers? No: they are the function boundaries of kernels.cuh found by a
regex (every `RT1W_DEV` / `template` definition starts a region) plus whole files for the others.
"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_starts(path):
    starts = []
    for n, line in enumerate(open(path), 1):
        m = re.match(r"(?:template <[^>]*>\s*)?(?:RT1W_DEV|__global__|static|RT1W_HD)\s.*?(\w+)\(", line)
        if m and not line.startswith(" "):
            starts.append((n, m.group(1)))
    return starts


def main():
    rep, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    path = func = hdr = None
    seen, agg = set(), {}
    starts = {}
    first_func = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or want not in (func or ""):
            continue
        if first_func is None:
            first_func = func
        if func != first_func or (path, int(r[0])) in seen:
            continue
        seen.add((path, int(r[0])))
        d = dict(zip(hdr[2:], r[2:]))
        try:
            inst, thr, st = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["Warp Stall Sampling (All Samples)"] or 0)
        except (KeyError, ValueError):
            continue
        if path not in starts:
            full = None
            for base in ("raytracing-1w_b200/csrc", "raytracing-1w_b200/host"):
                if os.path.exists(os.path.join(ROOT, base, path)):
                    full = os.path.join(ROOT, base, path)
            starts[path] = function_starts(full) if full else []
        name = path
        for n, fn in starts[path]:
            if n <= int(r[0]):
                name = f"{path}:{fn}"
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += inst
        a[1] += thr
        a[2] += st
    tot = sum(a[0] for a in agg.values()) or 1
    stt = sum(a[2] for a in agg.values()) or 1
    print(f"== {first_func}: {tot} warp instructions")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * a[0] / tot:5.1f}% inst {100 * a[2] / stt:5.1f}% stall  thr {a[1] / max(a[0], 1):4.1f}  {name}")


if __name__ == "__main__":
    main()
