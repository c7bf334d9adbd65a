"""Stall samples of an ncu report (--set full --import-source on) per device function and stall reason.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep k_wave [reason-substring ...]
Prints the source page's column names once (stderr), then for every stall-reason column whose name contains one of the
substrings (default: no_inst, long_sb, wait, short_sb, branch) the functions that hold most of its samples.
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_regions import function_starts  # noqa: E402


def main():
    rep, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    reasons = sys.argv[3:] or ["no_inst", "long_sb", "wait", "short_sb", "branch"]
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    path = func = hdr = first_func = None
    seen, agg, starts = set(), {}, {}
    cols = []
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
            if not cols:
                cols = [c for c in hdr if "stall" in c.lower()]
                print("columns:", hdr, file=sys.stderr)
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or want not in (func or ""):
            continue
        if first_func is None:
            first_func = func
        if func != first_func or (path, int(r[0])) in seen:
            continue
        seen.add((path, int(r[0])))
        d = dict(zip(hdr, r))
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
        a = agg.setdefault(name, {})
        for c in cols + ["Instructions Executed"]:
            try:
                a[c] = a.get(c, 0) + int(d.get(c) or 0)
            except ValueError:
                pass
    print(f"== {first_func}")
    for c in cols:
        if not any(s in c for s in reasons):
            continue
        tot = sum(a.get(c, 0) for a in agg.values())
        if tot == 0:
            continue
        print(f"-- {c}: {tot} samples")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1].get(c, 0))[:14]:
            if a.get(c, 0) == 0:
                break
            print(f"   {100 * a[c] / tot:5.1f}%  ({a[c] / max(a.get('Instructions Executed', 0), 1) * 1e3:7.3f} samples per 1000 warp instructions)  {name}")


if __name__ == "__main__":
    main()
