"""Per-source-line summary of an ncu report captured with --import-source on (-lineinfo build).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top-N]
Aggregates `ncu --page source --print-source cuda,sass --csv` rows that carry a CUDA line number.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    path, func, hdr, done = None, None, None, set()
    agg = {}
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
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        if want not in (func or ""):
            continue
        key = (func, path, int(r[0]))
        if key in agg:
            continue  # the same kernel captured more than once: keep the first
        d = dict(zip(hdr[2:], r[2:]))
        try:
            agg[key] = (int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["Warp Stall Sampling (All Samples)"] or 0), r[1].strip())
        except (KeyError, ValueError):
            pass
    # only the first function instance
    funcs = []
    for k in agg:
        if k[0] not in funcs:
            funcs.append(k[0])
    for f in funcs[:1]:
        items = [(v, k) for k, v in agg.items() if k[0] == f]
        tot = sum(v[0] for v, _ in items) or 1
        st = sum(v[2] for v, _ in items) or 1
        print(f"== {f}: {tot} warp instructions, {sum(v[1] for v, _ in items) / tot:.1f} threads/instruction, {st} stall samples")
        items.sort(key=lambda x: -x[0][0])
        for v, k in items[:top]:
            print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[2] / st:5.1f}% stall  thr {v[1] / max(v[0], 1):4.1f}  {k[1]}:{k[2]:<4d} | {v[3][:100]}")


if __name__ == "__main__":
    main()
