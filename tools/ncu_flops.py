"""profiles/r02_flops.json from the captures of tools/profiles_capture.sh: per BASELINE config, the fp32 / fp64 flops per ray that
ncu COUNTED over all wave launches of one whole render (fadd + fmul + 2 ffma, dadd + dmul + 2 dfma; predicated-on
threads), the thread instructions per ray, and the DRAM bytes of the steady-state launch of the `--set full` capture.
bench.py multiplies flops per ray by the rays it counts live (SURVEY.md section 8d's secondary roofline figure).

    python tools/ncu_flops.py r02        # reads gpurun_out/r02_<config>_flops.{csv,json} and r02_<config>_wave.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = ["C1", "C2", "C2w", "C3", "C5"]


def metric_sums(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    name, value, kern = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    sums, launches, kernels = {}, set(), set()
    for r in rows[1:]:
        sums[r[name]] = sums.get(r[name], 0.0) + float(r[value].replace(",", ""))
        launches.add(r[0])
        kernels.add(r[kern].split("(")[0])
    return sums, len(launches), sorted(kernels)


def full_capture(path):
    if not os.path.exists(path):
        return {}
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    if len(rows) < 3:
        return {}
    hdr, units, r = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def get(k):
        return float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] else None

    out = {}
    if "dram__bytes_read.sum" in hdr:
        out["dram_bytes_per_launch"] = sum(get(k) * scale.get(units[hdr.index(k)], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    for key, label in (("gpu__time_duration.sum", "launch_duration"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_instruction"),
                       ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_slot_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
                       ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
                       ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                       ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct")):
        v = get(key)
        if v is not None:
            out[label] = v
            if key == "gpu__time_duration.sum":
                out["launch_duration_unit"] = units[hdr.index(key)]
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
    out = {}
    # a partial re-capture (profiles_capture.sh TAG "C2") keeps the other configs' records: those of an earlier partial digest
    # into the same directory, else the committed ones
    for earlier in (os.path.join(out_dir, f"{tag}_flops.json"), os.path.join(ROOT, "profiles", f"{tag}_flops.json")):
        if os.path.exists(earlier):
            out = json.load(open(earlier))
            break
    for c in CONFIGS:
        base = os.path.join(ROOT, "gpurun_out", f"{tag}_{c}")
        if not os.path.exists(base + "_flops.csv"):
            continue
        sums, launches, kernels = metric_sums(base + "_flops.csv")
        perf = [json.loads(l) for l in open(base + "_flops.json") if l.startswith("{")][-1]
        rays = perf["mrays_s"] * perf["render_ms"] * 1e3

        def m(op):
            return sums.get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum", 0.0)

        f32 = m("fadd") + m("fmul") + 2.0 * m("ffma")
        f64 = m("dadd") + m("dmul") + 2.0 * m("dfma")
        rec = {"capture": f"profiles/{tag}_{c}_*: ncu over the {launches} wave launches of one render of {perf['scene']} "
                          f"{perf['image'][0]}x{perf['image'][1]} x {perf['spp']} spp ({int(rays)} rays)", "kernels": kernels,
               "fp32_flops_per_ray": f32 / rays, "fp64_flops_per_ray": f64 / rays,
               "thread_instructions_per_ray": sums.get("smsp__thread_inst_executed.sum", 0.0) / rays, "rays_per_path": perf["rays_per_path"]}
        rec.update(full_capture(base + "_wave.ncu-rep"))
        out[c] = rec
    with open(os.path.join(out_dir, f"{tag}_flops.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
