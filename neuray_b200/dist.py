"""Multi-GPU plumbing for the rendering path: one process per GPU, torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests).

The reference has no distributed code (train/trainer.py:65-67 raises NotImplementedError), so this is new
functionality with one correctness contract each:
  * inference: rays are independent -> shard contiguous ray ranges over ranks, render locally, all-gather the
    rendered tiles; the result must equal the single-rank render bit for bit.  Every output key travels in ONE
    collective: the per-ray outputs are laid side by side as fp32 columns of one [rays, C] tile (bool masks as 0/1,
    which is exact), so a frame costs one all_gather launch, not one per key.
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


def _events(device):
    if device.type != "cuda":
        return None
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def render_sharded(render_fn, que_imgs_info, ref_imgs_info, is_train=False, group=None, keys=None, timing=None):
    """Renders rank's share of que_imgs_info['coords'] with `render_fn(que, ref, is_train) -> dict` and all-gathers
    every output (or `keys`) along the ray axis.  Every rank returns the full-image dict.

    render_fn must return per-ray tensors [1, rays, ...] for every key -- also for an EMPTY shard (zero-length tensors of
    the right trailing shape; renderer.render does), so that every rank issues the same collective even when there are
    fewer rays than ranks.  `timing`: optional list; (start, stop) CUDA events around the collective are appended."""
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
    names = sorted(out if keys is None else keys)
    if not names:
        raise RuntimeError("render_sharded: the render function returned no outputs (an empty shard must still return "
                           "zero-length tensors for every key)")
    # one fp32 tile [max_len, C]: key k occupies columns [off_k, off_k + prod(trailing shape))
    layout, cols = [], []
    for k in names:
        v = out[k]
        if v.shape[0] != 1 or v.shape[1] != e - s:
            raise RuntimeError(f"render_sharded: output {k} has shape {tuple(v.shape)}, expected [1, {e - s}, ...]")
        width = 1
        for d in v.shape[2:]:
            width *= d
        layout.append((k, v.dtype, tuple(v.shape[2:]), width))
        cols.append(v.reshape(e - s, width).to(torch.float32))
    max_len = -(-n // world)
    total = sum(wd for *_, wd in layout)
    dev = coords.device
    tile = torch.zeros(max_len, total, dtype=torch.float32, device=dev)
    if e > s:
        tile[: e - s] = torch.cat(cols, 1)
    gathered = torch.empty(world, max_len, total, dtype=torch.float32, device=dev)
    ev = _events(dev) if timing is not None else None
    if ev:
        ev[0].record()
    dist.all_gather_into_tensor(gathered.view(world * max_len, total), tile, group=group)
    if ev:
        ev[1].record()
        timing.append(ev)
    pieces = []
    for r in range(world):
        rs, re = ray_range(n, r, world)
        pieces.append(gathered[r, : re - rs])
    full_tile = torch.cat(pieces, 0)                                   # [n, C]
    full, off = {}, 0
    for k, dtype, trail, width in layout:
        v = full_tile[:, off:off + width].reshape((1, n) + trail)
        full[k] = (v != 0) if dtype == torch.bool else v.to(dtype).contiguous()
        off += width
    return full


def allreduce_gradients(parameters, group=None, timing=None):
    """Averages .grad of `parameters` over the group with a single flat all-reduce (parameters without a grad
    contribute zeros, so every rank issues the same collective).  `timing`: optional list; (start, stop) CUDA events
    around the collective are appended."""
    params = [p for p in parameters if p.requires_grad]
    if not params or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    ev = _events(flat.device) if timing is not None else None
    if ev:
        ev[0].record()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if ev:
        ev[1].record()
        timing.append(ev)
    flat /= world
    grads = torch.split(flat, [p.numel() for p in params])
    for p, g in zip(params, grads):
        if p.grad is None:
            p.grad = g.view_as(p).clone()
        else:
            p.grad.copy_(g.view_as(p))
