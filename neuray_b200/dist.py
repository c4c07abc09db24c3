"""Multi-GPU plumbing for the rendering path: one process per GPU, torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests).

The reference has no distributed code (train/trainer.py:65-67 raises NotImplementedError), so this is new
functionality with one correctness contract each:
  * inference: rays are independent -> shard contiguous ray ranges over ranks, render locally, all-gather the
    rendered tiles; the result must equal the single-rank render bit for bit
  * training: data parallel, one batch per rank -> all-reduce(sum)/world of the gradients in ONE flat bucket
    (3.08 M fp32 = 12.3 MB: latency-bound, so one launch rather than per-tensor reductions)
"""
import torch
import torch.distributed as dist


def ray_range(n_rays, rank, world):
    """Contiguous, balanced [start, stop) of rank's rays (first n_rays % world ranks get one more)."""
    base, extra = divmod(n_rays, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def render_sharded(render_fn, que_imgs_info, ref_imgs_info, is_train=False, group=None, keys=None):
    """Renders rank's share of que_imgs_info['coords'] with `render_fn(que, ref, is_train) -> dict` and all-gathers
    every output along the ray axis.  Every rank returns the full-image dict."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    coords = que_imgs_info["coords"]
    n = coords.shape[1]
    s, e = ray_range(n, rank, world)
    q = dict(que_imgs_info)
    q["coords"] = coords[:, s:e].contiguous()
    out = render_fn(q, ref_imgs_info, is_train)
    if world == 1:
        return out
    max_len = -(-n // world)
    full = {}
    for k in sorted(out if keys is None else keys):
        v = out[k]
        as_bool = v.dtype == torch.bool
        if as_bool:
            v = v.to(torch.uint8)
        pad = torch.zeros((v.shape[0], max_len) + tuple(v.shape[2:]), dtype=v.dtype, device=v.device)
        pad[:, : v.shape[1]] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        pieces = []
        for r, part in enumerate(parts):
            rs, re = ray_range(n, r, world)
            pieces.append(part[:, : re - rs])
        v = torch.cat(pieces, 1)
        full[k] = v.bool() if as_bool else v
    return full


def allreduce_gradients(parameters, group=None):
    """Averages .grad of `parameters` over the group with a single flat all-reduce (parameters without a grad
    contribute zeros, so every rank issues the same collective)."""
    params = [p for p in parameters if p.requires_grad]
    if not params or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
