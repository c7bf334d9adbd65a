#!/usr/bin/env python
"""Headline benchmark: Mpaths/s of the Cornell box (main.rs scene 5: 600x600, 100 spp, depth 50) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W              # the CUDA wavefront renderer (this repo)
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host CPU cores

A "step" is one render of the whole image for one sample range (W*H*spp paths).  With N > 1 (torchrun, one
rank per GPU) rank r renders samples [r*spp, (r+1)*spp) of every pixel (weak scaling: per-GPU work fixed) and
the partial radiance sums are combined with one NCCL reduce.  Prints ONE JSON line (rank 0).

value   device-resident: scene committed, result left in HBM, timed with CUDA events on the render stream
e2e     through the C ABI the reference-side host calls (rt1w_render): camera/params in, radiance sums copied
        back to a pinned HOST buffer inside the timed region
roofline  the wave kernel (scatter + closest hit + regroup, one launch per wave) against the measured HBM peak, on
        SURVEY.md section 8(d)'s algorithmic bytes (148 B per ray segment + 60 B per path); kernel time measured
        live with CUDA events around every launch (RT1W_FLAG_PROFILE pass)
cpu_baseline  the C++ f64 restatement of the reference (oracle/, kind "port": the Rust crate cannot be built in
        this image) on the host cores, on a bounded sample of the same workload
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SCENE = "cornel_box"
WIDTH, HEIGHT, SPP, DEPTH = 600, 600, 100, 50
# algorithmic bytes of one ray segment in a wavefront with fp32 SoA queues
BYTES_PER_RAY = 148.0       # whole pipeline
BYTES_PER_PATH = 60.0       # generate write + accumulate
METRIC = "Mpaths/s (Cornell box 600x600, 100 spp per GPU, depth 50)"


def ncu_traffic():
    """DRAM bytes of one steady-state k_wave launch (8.39 M work items) from the committed `ncu --set full` capture."""
    try:
        vals = {}
        for line in open(os.path.join(ROOT, "profiles", "r01_wave_metrics.txt")):
            k, v, *unit = line.split()
            if k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                vals[k] = float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit[0]]
        return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def run_reference(args):
    """The reference arm: the reference's own CPU algorithm (oracle port) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rt = importlib.import_module("raytracing-1w_b200")
    api = rt.api
    import oracle_binding  # the one place besides tests/smoke where bench.py runs the oracle: as the measured reference arm

    hs = api.HostScene(SCENE, seed=1)
    osc = oracle_binding.OracleScene(hs.desc)
    cam = hs.camera()
    spp = args.ref_spp
    total_s, paths, rays, threads = 0.0, 0, 0, 1
    for step in range(args.warmup + args.steps):
        p = hs.params(width=WIDTH, spp=(step + 1) * spp, sample_begin=step * spp)
        _, _, st = osc.render(cam, p, threads=0)
        if step >= args.warmup:
            total_s += st.seconds
            paths += st.paths
            rays += st.rays
            threads = st.threads
    value = paths / total_s / 1e6
    sample = f"{WIDTH}x{HEIGHT} x {spp} spp per step ({WIDTH * HEIGHT * spp} paths), depth {DEPTH}, all rows"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cornel_box 600x600 depth 50 (main.rs:395-512,867-894), bounded sample", "sample": sample},
        "mrays_per_s": rays / total_s / 1e6, "rays_per_path": rays / max(paths, 1),
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-spp", type=int, default=8, help="samples per pixel of one reference-arm step")
    ap.add_argument("--cpu-spp", type=int, default=32, help="samples per pixel of the cpu_baseline sample")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    rt = importlib.import_module("raytracing-1w_b200")
    api = rt.api
    hs = api.HostScene(SCENE, seed=1)
    ctx = api.Context(local)
    scene = api.Scene(ctx, hs.desc)
    cam = hs.camera()
    n_pixels = WIDTH * HEIGHT
    s0, s1 = rank * SPP, (rank + 1) * SPP  # this rank's sample range

    def params(flags=0):
        return hs.params(width=WIDTH, height=HEIGHT, spp=s1, sample_begin=s0, seed=0, flags=flags, pool_paths=args.pool)

    stream = torch.cuda.current_stream()
    accum = torch.empty(n_pixels * 3, dtype=torch.float32, device="cuda")
    host_out = torch.empty(n_pixels * 3, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB of L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        st = scene.render_device(cam, params(), accum.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        return st

    def step_e2e():
        if world == 1:
            return scene.render_into(cam, params(), host_out.numpy())
        st = scene.render_device(cam, params(), accum.data_ptr(), stream.cuda_stream)
        dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            host_out.copy_(accum, non_blocking=False)
        return st

    # ---- device-resident arm
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    total_ms, launches, rays, paths = 0.0, 0, 0, 0
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = step_device()
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        launches += st.launches
        rays += st.rays
        paths += st.paths
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, float(rays), float(paths), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0])
        rays, paths, launches = float(t[1]), float(t[2]), float(t[3])
    value = paths / (total_ms * 1e-3) / 1e6
    mrays = rays / (total_ms * 1e-3) / 1e6

    # ---- end-to-end arm (host buffers through the C ABI)
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = (n_pixels * SPP * world * args.steps) / float(te[0]) / 1e6

    # ---- per-kernel times (profiling pass, rank 0 only; events around every launch)
    roofline, kernels = None, {}
    if rank == 0:
        stp = scene.render_device(cam, params(api.FLAG_PROFILE), accum.data_ptr(), stream.cuda_stream)
        ksum = sum(stp.kernel_ms)
        for k, name in enumerate(api.KERNEL_NAMES):
            if stp.kernel_launches[k]:
                kernels[name] = {"ms": round(stp.kernel_ms[k], 3), "launches": int(stp.kernel_launches[k]),
                                 "share": round(stp.kernel_ms[k] / ksum, 4)}
        peak, which = measured_peaks()
        wave_ms, wave_n = stp.kernel_ms[0], max(stp.kernel_launches[0], 1)
        alg_bytes = BYTES_PER_RAY * stp.rays + BYTES_PER_PATH * stp.paths  # all launches of one render
        achieved = alg_bytes / (wave_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_wave", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(), "peak_source": which,
                    "traffic_note": "ncu dram read+write of ONE steady-state launch (2^23 work items: 1.34e9 algorithmic bytes); "
                                    "the per-launch average on the left includes the small launches of the tail",
                    "algorithmic_bytes_per_launch": alg_bytes / wave_n, "avg_launch_ms": wave_ms / wave_n,
                    "note": "148 B per ray segment + 60 B per path (SURVEY.md 8d) over all k_wave launches of one render; "
                            "the kernel moves 160 B per segment through HBM and is bound by instruction issue / latency, see DESIGN.md"}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_binding  # checker used as the reported CPU baseline (never on the product path)

        osc = oracle_binding.OracleScene(hs.desc)
        pc = hs.params(width=WIDTH, height=HEIGHT, spp=args.cpu_spp)
        _, _, ost = osc.render(cam, pc, threads=0)
        cpu = {"value": ost.paths / ost.seconds / 1e6, "unit": "Mpaths/s", "cores": int(ost.threads), "kind": "port",
               "sample": f"{WIDTH}x{HEIGHT} x {args.cpu_spp} spp ({ost.paths} paths) of the same scene, {ost.seconds:.1f} s",
               "mrays_per_s": ost.rays / ost.seconds / 1e6}

    if rank == 0:
        info = scene.info()
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 intersection / f32 shading", "data": "synthetic",
            "config": {"workload": "cornel_box 600x600, 100 spp per GPU, depth 50 (main.rs:395-512,867-894; BASELINE.json configs[0])",
                       "sharding": f"sample ranges, rank r renders [{SPP}r, {SPP}(r+1)); NCCL reduce of {n_pixels * 12} B" if world > 1 else "single GPU",
                       "l2": "flushed between timed iterations (256 MiB memset)", "pool_paths": args.pool or (1 << 23),
                       "prims": info.n_prims, "bvh_nodes": info.n_bvh_nodes},
            "mrays_per_s": mrays, "rays_per_path": rays / max(paths, 1.0),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": 312, "d2h_bytes_per_step": n_pixels * 12},
            "gpu_launches": int(launches), "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    scene.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
