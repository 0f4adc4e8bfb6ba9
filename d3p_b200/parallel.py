"""Sharding one minibatch over the GPUs of a box (new; the reference is single-device).

One process per GPU (``torch.distributed``, NCCL over NVLink).  All ranks hold the same
``DPSVIState`` and the same batchifier state, so
  * the sampler runs redundantly on every rank (ALU-only, no traffic) and yields the bit-identical
    index list everywhere — no collective for indices;
  * rank r computes per-example gradients only for batch positions ``[r*ceil(B/G), (r+1)*ceil(B/G))``
    (per-example Threefry keys are addressed by position, so they match the single-GPU run);
  * the ``P + 2`` clipped sums (gradient, loss, valid count) are all-reduced once per step;
  * every rank draws the same ChaCha noise and applies the same optimizer step.
The sum order differs from the single-GPU run, so parameters agree to fp32 reassociation error.
"""
import torch
import torch.distributed as dist

from . import _native as _n


def shard_dpsvi(svi, rank=None, world_size=None, group=None):
    """Make ``svi.update`` process only this rank's slice of every batch and all-reduce the sums."""
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    buf = {}

    def reduce_fn(ws, n_partials, P):
        key = (ws.device, P)
        if key not in buf:
            buf[key] = torch.empty(P + 2, dtype=torch.float32, device=ws.device)
        out = buf[key]
        _n.check(_n.lib().d3p_reduce_partials_f32(_n.ptr(ws), n_partials, P, _n.ptr(out), _n.stream_ptr()),
                 "reduce_partials")
        if world_size > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out, 1

    svi.shard = (rank, world_size, reduce_fn)
    return svi


def position_range(B, rank, world_size):
    per = (B + world_size - 1) // world_size
    return min(B, rank * per), min(B, (rank + 1) * per)
