"""Multi-GPU sharding of the sample loop (one process per GPU).

The reference's pixel loop has no cross-pixel or cross-sample state (main.rs:957-993), so the path shards by
SAMPLE RANGE: rank r renders samples [begin_r, end_r) of every pixel with the Philox counter keyed by the global
sample index, and the partial radiance sums are added with one NCCL reduce over NVLink (the only exchange step).
The result is independent of the rank count up to fp32 summation order.
"""
import torch
import torch.distributed as dist


def sample_range(rank, world, spp_total):
    """Contiguous split of [0, spp_total) into `world` ranges whose sizes differ by at most one (strong scaling)."""
    base, extra = divmod(spp_total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def weak_sample_range(rank, spp_per_rank):
    """Weak scaling: every rank renders spp_per_rank samples; the image ends up with world * spp_per_rank."""
    return rank * spp_per_rank, (rank + 1) * spp_per_rank


def reduce_radiance(partial, dst=0):
    """Sum of the per-rank radiance-sum buffers on rank `dst` (NCCL on GPU tensors, gloo on CPU tensors)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM)
    return partial


def resolve(total_sum, spp_total):
    """Color::into_sampled on the reduced buffer (color.rs:14-21): NaN sum -> 0, then the mean."""
    return torch.where(torch.isnan(total_sum), torch.zeros_like(total_sum), total_sum) / float(spp_total)
