"""BASELINE.json config 4: the Cornell box at 3840x2160, 4096 spp, sharded by sample range across the ranks, one NCCL
reduce of the 99.5 MB radiance sums to rank 0.  Not a bench line (bench.py measures config 1); prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4.py [--spp 4096]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=4096, help="total samples per pixel over all ranks")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = importlib.import_module("raytracing-1w_b200")
    api = rt.api
    shard = importlib.import_module("raytracing-1w_b200.shard")
    hs = api.HostScene("cornel_box", seed=1)
    ctx = api.Context(local)
    scene = api.Scene(ctx, hs.desc)
    cam = hs.camera(aspect=args.width / args.height)  # Camera::new gets the 16:9 aspect
    s0, s1 = shard.sample_range(rank, world, args.spp)
    accum = torch.empty(args.width * args.height * 3, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        st = scene.render_device(cam, hs.params(width=args.width, height=args.height, spp=s1, sample_begin=s0, pool_paths=1 << 23), accum.data_ptr(),
                                 stream.cuda_stream)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        return st

    # warm-up with the same wave capacity, so that the queues are allocated outside the timed region
    scene.render_device(cam, hs.params(width=args.width, height=args.height, spp=s0 + 1, sample_begin=s0, pool_paths=1 << 23), accum.data_ptr(), stream.cuda_stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    st = step()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), float(st.rays), float(st.paths)], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(tmax[0])
        img = accum.cpu().numpy().reshape(args.height, args.width, 3)
        print(json.dumps({"config": f"cornel_box {args.width}x{args.height}, {args.spp} spp over {world} GPU(s), depth 50", "ms": round(ms, 1),
                          "mpaths_s": round(float(t[2]) / ms / 1e3, 1), "mrays_s": round(float(t[1]) / ms / 1e3, 1),
                          "reduce_bytes": args.width * args.height * 12, "mean_radiance": float(img[img == img].mean() / args.spp)}), flush=True)
    scene.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
