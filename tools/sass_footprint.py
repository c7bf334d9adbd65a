"""Static instruction footprint of one kernel per source function (inlined code is charged to the function whose lines it
carries), from `nvdisasm -g` of the built library: what the 32 KB instruction cache has to hold.

    python tools/sass_footprint.py 'k_waveILb0ELb1ELb1ELb0ELi0E'         # substring of the mangled kernel name
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_regions import function_starts  # noqa: E402


def main():
    want = sys.argv[1]
    lib = os.environ.get("RT1W_LIB") or os.path.join(ROOT, "raytracing-1w_b200", "_build", "librt1w.so")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "render", lib], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    starts, agg, total = {}, {}, 0
    inside, name = False, "?"
    for line in text.splitlines():
        if line.startswith(".text."):
            inside = want in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            path, n = m.group(1), int(m.group(2))
            base = os.path.basename(path)
            if base not in starts:
                starts[base] = function_starts(path) if os.path.exists(path) and "/csrc/" in path else []
            name = base
            for s, fn in starts[base]:
                if s <= n:
                    name = f"{base}:{fn}"
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", line):
            agg[name] = agg.get(name, 0) + 1
            total += 1
    print(f"{want}: {total} instructions, {total * 16 / 1024:.1f} KB")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"{v:6d}  {v * 16 / 1024:6.1f} KB  {k}")


if __name__ == "__main__":
    main()
