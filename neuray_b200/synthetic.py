"""Seeded synthetic inputs in the reference's `imgs_info` dict format.

No dataset or checkpoint is available offline, so tests, goldens and bench.py all render synthetic scenes:
look-at cameras on an arc around the origin, pin-hole K from a field of view, random images and feature
maps of the shape the reference's encoders would produce (32 channels at 1/4 resolution), and full-image
query coords laid out like reference utils/imgs_info.py:122-131 (x fastest).

Everything is built with numpy RandomState on the CPU (portable across torch versions) and returned as CPU
fp32 torch tensors.
"""
import math

import numpy as np
import torch

from . import modules


def look_at_pose(centre, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """World->camera [R|t] (OpenCV convention: x right, y down, z forward), x_cam = R x + t."""
    c = np.asarray(centre, np.float64)
    z = np.asarray(target, np.float64) - c
    z /= np.linalg.norm(z)
    x = np.cross(z, np.asarray(up, np.float64))
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)
    return np.concatenate([R, -(R @ c)[:, None]], 1).astype(np.float32)


def _smooth_noise(rs, shape, smooth):
    """N(0,1) noise; `smooth` > 1 generates on a coarser grid and upsamples (bilinear) so that the maps
    vary slowly from texel to texel, like encoder outputs do."""
    n, c, h, w = shape
    if smooth <= 1:
        return rs.standard_normal(shape).astype(np.float32)
    ch, cw = max(2, h // smooth), max(2, w // smooth)
    coarse = torch.from_numpy(rs.standard_normal((n, c, ch, cw)).astype(np.float32))
    return torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True).numpy()


def make_scene(h, w, rfn, que_h=None, que_w=None, fov_x=0.6911112, radius=4.0, depth_range=(2.0, 6.0), seed=0,
               arc_deg=60.0, pad=16, smooth=1, with_que_imgs=True, focal=None):
    """Returns (que_imgs_info, ref_imgs_info) as dicts of CPU tensors.

    Reference images are `h x w` padded up to a multiple of `pad` (reference utils/imgs_info.py:60-75);
    feature maps are [rfn,32,H/4,W/4] of the padded size.  The query image is `que_h x que_w` (defaults h,w).
    """
    rs = np.random.RandomState(seed)
    que_h, que_w = que_h or h, que_w or w
    f = focal if focal is not None else 0.5 * w / math.tan(0.5 * fov_x)
    ph, pw = (h + pad - 1) // pad * pad, (w + pad - 1) // pad * pad

    def K_for(width, height, scale=1.0):
        return np.array([[f * scale, 0, 0.5 * width], [0, f * scale, 0.5 * height], [0, 0, 1]], np.float32)

    def cam_centre(az_deg, el_deg, r):
        az, el = math.radians(az_deg), math.radians(el_deg)
        return (r * math.cos(el) * math.cos(az), r * math.cos(el) * math.sin(az), r * math.sin(el))

    az = np.linspace(-arc_deg / 2, arc_deg / 2, rfn) if rfn > 1 else np.zeros(1)
    el = 25.0 + 10.0 * rs.uniform(-1, 1, rfn)
    rr = radius * (1.0 + 0.05 * rs.uniform(-1, 1, rfn))
    ref_poses = np.stack([look_at_pose(cam_centre(a, e, r)) for a, e, r in zip(az, el, rr)], 0)
    ref_Ks = np.stack([K_for(w, h) for _ in range(rfn)], 0)
    near, far = depth_range
    ref_dr = np.stack([np.array([near * (1 + 0.03 * rs.uniform(-1, 1)), far * (1 + 0.03 * rs.uniform(-1, 1))],
                                np.float32) for _ in range(rfn)], 0)
    ref = {
        "imgs": rs.uniform(0, 1, (rfn, 3, ph, pw)).astype(np.float32),
        "poses": ref_poses, "Ks": ref_Ks, "depth_range": ref_dr,
        "ray_feats": _smooth_noise(rs, (rfn, 32, ph // 4, pw // 4), smooth),
        "img_feats": _smooth_noise(rs, (rfn, 32, ph // 4, pw // 4), smooth),
    }
    que_pose = look_at_pose(cam_centre(0.13 * arc_deg, 27.0, radius * 0.98))[None]
    xs, ys = np.meshgrid(np.arange(que_w), np.arange(que_h))
    coords = np.stack([xs, ys], -1).reshape(1, -1, 2).astype(np.float32)
    que = {
        "poses": que_pose, "Ks": K_for(que_w, que_h, que_w / w)[None], "coords": coords,
        "depth_range": np.array([[near, far]], np.float32),
    }
    if with_que_imgs:
        que["imgs"] = rs.uniform(0, 1, (1, 3, que_h, que_w)).astype(np.float32)
    to_t = lambda d: {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()}
    return to_t(que), to_t(ref)


def make_weights(cfg, seed=0, sigma_gain=8.0, sigma_bias=0.3):
    """Random-init weights under the reference's state-dict names (coarse + optional fine nets).

    Plain random init yields near-zero density everywhere; the last density layer is
    rescaled so that rays actually terminate and the fine resampling has structure to follow.
    """
    torch.manual_seed(seed)
    nets = {
        "dist_decoder": modules.MixtureLogisticsDistDecoder(cfg.get("dist_decoder_cfg", {})),
        "agg_net": modules.DefaultAggregationNet(cfg.get("agg_net_cfg", {})),
    }
    if cfg.get("use_hierarchical_sampling", False):
        nets["fine_dist_decoder"] = modules.MixtureLogisticsDistDecoder(cfg.get("fine_dist_decoder_cfg", {}))
        nets["fine_agg_net"] = modules.DefaultAggregationNet(cfg.get("fine_agg_net_cfg", {}))
    W = {}
    rs = np.random.RandomState(seed + 1)
    for name, net in nets.items():
        for k, v in net.state_dict().items():
            v = v.detach().clone().float()
            if k.endswith(".bias") and v.dim() == 1 and v.numel() > 1:
                v = v + torch.from_numpy(rs.uniform(-0.1, 0.1, v.shape).astype(np.float32))  # exercise biases
            W[f"{name}.{k}"] = v
    for agg in ("agg_net", "fine_agg_net"):
        key = f"{agg}.agg_impl.out_geometry_fc.2"
        if key + ".weight" in W:
            W[key + ".weight"] = W[key + ".weight"] * sigma_gain
            W[key + ".bias"] = W[key + ".bias"] + sigma_bias
    return W


def slice_rays(que, start, stop):
    q = dict(que)
    q["coords"] = que["coords"][:, start:stop].contiguous()
    return q


def to_device(d, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()}
