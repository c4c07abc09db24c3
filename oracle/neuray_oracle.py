"""TEST INFRASTRUCTURE ONLY -- CPU oracle for NeuRay's per-ray rendering hot path.

A functional fp32 restatement (plain torch tensor arithmetic on the CPU, explicit bilinear taps, no
nn.Module, no F.grid_sample) of the algorithm in the reference (liuyuan-pal/NeuRay @ 17a52d3):

    network/render_ops.py:4-229       ray geometry, sampling, reprojection, gather, compositing, resample
    network/ops.py:14-34              interpolate_feats (grid_sample wrapper)
    network/dist_decoder.py:6-140     mixture-of-logistics visibility decoder
    network/aggregate_net.py:8-68     aggregation front-end
    network/ibrnet.py:7-102,239-369   IBRNetWithNeuRay + ray self-attention
    network/renderer.py:67-83,127-254 render_by_depth / render_impl / render chunk loop
and, next to the path (second half of this file):
    network/init_net.py:29-101        get_diff_feats, DepthInitNet (extract_depth_for_init, ResEncoder, depth_skip, conv_out)
    network/ops.py:43-312             ResidualBlock, BasicBlock, ResUNetLight / ResEncoder
    network/vis_encoder.py:6-21       DefaultVisEncoder
    network/renderer.py:280-316, network/loss.py:17-132   predict_mean_for_depth_loss and the three losses
    network/init_net.py:103-254, network/mvsnet/          CostVolumeInitNet with its MVSNet (feature net, homography variance
                                      volume, 3-D regulariser, softmax + depth regression)

Pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container by
oracle/gen_golden*.py and committed under tests/golden/ (checked by tests/test_oracle_golden.py,
test_diff_feats.py, test_encoders_cpu.py, test_losses.py, test_mvsnet.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product path (neuray_b200/) never does.

Weights are passed as a flat dict keyed by the reference's state-dict names, e.g.
'dist_decoder.mean_decoder.0.weight', 'agg_net.agg_impl.base_fc.0.weight'.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# small helpers


def _lin(W, name, x):
    """y = x @ weight^T + bias for the Linear called `name` (bias optional)."""
    y = x @ W[name + ".weight"].t()
    b = W.get(name + ".bias")
    return y if b is None else y + b


def _elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def _softplus(x):
    # torch.nn.Softplus(beta=1, threshold=20)
    return torch.where(x > 20, x, torch.log1p(torch.exp(x)))


def _sigmoid(x):
    return 1.0 / (1.0 + torch.exp(-x))


# --------------------------------------------------------------------------------------------------
# render_ops.py


def coords2rays(coords, poses, Ks):
    """render_ops.py:4-25.  coords [n,rn,2], poses [n,3,4] (world->cam), Ks [n,3,3]."""
    n, rn, _ = coords.shape
    Rt = poses[:, :, :3].transpose(1, 2)              # cam->world rotation  [n,3,3]
    centre = -(Rt @ poses[:, :, 3:])                   # camera centre        [n,3,1]
    hom = torch.cat([coords, torch.ones(n, rn, 1, dtype=torch.float32)], 2)  # [n,rn,3]
    cam = torch.inverse(Ks)[:, None] @ hom[..., None]  # [n,rn,3,1]
    world = Rt[:, None] @ cam + centre[:, None]
    centres = centre[:, None, :, 0].expand(n, rn, 3)
    # the reference subtracts the centre again instead of skipping the add (render_ops.py:22-23)
    directions = world[..., 0] - centres
    return centres, directions


def depth2points(que_imgs_info, que_depth):
    """render_ops.py:27-39."""
    o, d = coords2rays(que_imgs_info["coords"], que_imgs_info["poses"], que_imgs_info["Ks"])
    pts = o[:, :, None] + d[:, :, None] * que_depth[..., None]
    dn = que_depth.shape[-1]
    view_dir = -d / torch.linalg.norm(d, dim=2, keepdim=True)
    return pts, view_dir[:, :, None].expand(-1, -1, dn, -1).contiguous()


def depth2dists(depth):
    """render_ops.py:41-44."""
    tail = torch.full_like(depth[..., :1], 1e6)
    return torch.cat([depth[..., 1:] - depth[..., :-1], tail], -1)


def depth2inv_dists(depth, depth_range):
    """render_ops.py:46-52."""
    a = (-1 / depth_range[:, 0])[:, None, None]
    b = (-1 / depth_range[:, 1])[:, None, None]
    return depth2dists((-1 / depth - a) / (b - a))


def _unnormalise(g, size, align_corners):
    if align_corners:
        return (g + 1) / 2 * (size - 1)
    return ((g + 1) * size - 1) / 2


def bilinear_sample(feats, pts, h=None, w=None, padding_mode="zeros", align_corners=False):
    """ops.py:14-34 (interpolate_feats) with F.grid_sample(mode='bilinear') written out tap by tap.

    feats [b,c,fh,fw]; pts [b,n,2] in pixel units of an (h,w) image -> [b,n,c].
    """
    b, c, fh, fw = feats.shape
    if h is None and w is None:
        h, w = fh, fw
    gx = pts[..., 0] / (w - 1) * 2 - 1
    gy = pts[..., 1] / (h - 1) * 2 - 1
    ix = _unnormalise(gx, fw, align_corners)
    iy = _unnormalise(gy, fh, align_corners)
    if padding_mode == "border":
        ix = ix.clamp(0, fw - 1)
        iy = iy.clamp(0, fh - 1)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    tx = ix - x0
    ty = iy - y0
    flat = feats.reshape(b, c, fh * fw)
    out = torch.zeros(b, c, pts.shape[1], dtype=torch.float32)
    for dy, dx, wgt in ((0, 0, (1 - tx) * (1 - ty)), (0, 1, tx * (1 - ty)),
                        (1, 0, (1 - tx) * ty), (1, 1, tx * ty)):
        xi = x0 + dx
        yi = y0 + dy
        inside = (xi >= 0) & (xi <= fw - 1) & (yi >= 0) & (yi <= fh - 1)
        idx = (yi.clamp(0, fh - 1) * fw + xi.clamp(0, fw - 1)).long()   # [b,n]
        tap = torch.gather(flat, 2, idx[:, None, :].expand(b, c, -1))   # [b,c,n]
        out = out + tap * (wgt * inside.float())[:, None, :]
    return out.permute(0, 2, 1)


def interpolate_feature_map(ray_feats, coords, mask, h, w, border_type="border"):
    """render_ops.py:54-70."""
    fh, fw = ray_feats.shape[-2:]
    same = (fh == h) and (fw == w)
    return bilinear_sample(ray_feats, coords, h, w, border_type, same) * mask.float()[..., None]


def alpha_values2hit_prob(alpha):
    """render_ops.py:72-80."""
    keep = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1)
    return alpha * torch.cumprod(keep, -1)[..., :-1]


def project_points_coords(pts, Rt, K):
    """render_ops.py:82-104.  pts [pn,3] -> pix [rfn,pn,2], valid [rfn,pn], depth [rfn,pn,1]."""
    rfn = Rt.shape[0]
    P = K @ Rt                                                         # [rfn,3,4]
    bottom = torch.zeros(rfn, 1, 4)
    bottom[:, :, 3] = 1
    H = torch.cat([P, bottom], 1)
    hom = torch.cat([pts, torch.ones(pts.shape[0], 1)], 1)
    cam = (H[:, None] @ hom[None, :, :, None])[:, :, :3, 0]           # [rfn,pn,3]
    z = cam[:, :, 2:].clone()
    degenerate = z.abs() < 1e-4
    z[degenerate] = 1e-3
    return cam[:, :, :2] / z, ~degenerate[..., 0], z


def project_points_directions(poses, points):
    """render_ops.py:106-115."""
    centre = -(poses[:, :, :3].transpose(1, 2) @ poses[:, :, 3:])     # [rfn,3,1]
    d = points[None] - centre.transpose(1, 2)
    return -d / torch.linalg.norm(d, dim=2, keepdim=True).clamp_min(1e-5)


def project_points_ref_views(ref_imgs_info, que_points):
    """render_ops.py:117-130."""
    pix, ok, z = project_points_coords(que_points, ref_imgs_info["poses"], ref_imgs_info["Ks"])
    h, w = ref_imgs_info["imgs"].shape[-2:]
    outside = (pix[..., 0] < -0.5) | (pix[..., 0] >= w - 0.5) | (pix[..., 1] < -0.5) | (pix[..., 1] >= h - 0.5)
    return project_points_directions(ref_imgs_info["poses"], que_points), pix, z, ok & ~outside


def project_points_dict(ref_imgs_info, que_pts):
    """render_ops.py:132-144."""
    qn, rn, dn, _ = que_pts.shape
    pdir, pix, z, ok = project_points_ref_views(ref_imgs_info, que_pts.reshape(-1, 3))
    rfn, _, h, w = ref_imgs_info["imgs"].shape
    d = {
        "dir": pdir, "pts": pix, "depth": z, "mask": ok.float(),
        "ray_feats": interpolate_feature_map(ref_imgs_info["ray_feats"], pix, ok, h, w),
        "rgb": interpolate_feature_map(ref_imgs_info["imgs"], pix, ok, h, w),
    }
    return {k: v.reshape(rfn, qn, rn, dn, -1) for k, v in d.items()}


def sample_depth(depth_range, coords, sample_num, random_sample=False, jitter=None):
    """render_ops.py:146-170.  `jitter` (qn,rn,dn-2 uniform[0,1)) replaces torch.rand when random_sample."""
    qn, rn, _ = coords.shape
    near, far = depth_range[:, 0], depth_range[:, 1]
    dn = sample_num
    assert dn > 2
    span = 1 / far - 1 / near
    step = span / (dn - 1)
    k = torch.arange(1, dn - 1, dtype=torch.float32)[None, None, :]
    if random_sample:
        k = k + (jitter - 0.5) * 0.999
    else:
        k = k + torch.zeros(qn, rn, dn - 2)
    ticks = torch.cat([torch.zeros(qn, rn, 1), step[:, None, None] * k, span[:, None, None].expand(qn, rn, 1)], -1)
    depth = 1 / (1 / near[:, None, None] + ticks)
    nxt = torch.cat([depth[..., 1:], torch.full((qn, rn, 1), 1e6)], -1)
    return depth, nxt - depth


def fine_sample_u(fdn):
    """render_ops.py:199-202: deterministic bin-centre quantiles used at eval time."""
    step = 1 / fdn
    return 0.5 * step + torch.arange(fdn) * step


def sample_fine_depth(depth, hit_prob, depth_range, sample_num, random_sample, u=None, inv_mode=True):
    """render_ops.py:172-229.  `u` [qn,rn,fdn] replaces torch.rand when random_sample.  inv_mode=False (render_ops.py:182,224):
    the resampling runs directly in depth instead of normalised inverse depth."""
    if inv_mode:
        a = -1 / depth_range[0, 0]
        b = -1 / depth_range[0, 1]
        t = (-1 / depth - a) / (b - a)
    else:
        t = depth
    edges = torch.cat([t[..., :1], (t[..., 1:] + t[..., :-1]) / 2, t[..., -1:]], -1)   # dn+1
    p = hit_prob + 1e-5
    pdf = p / p.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)        # dn+1
    if not random_sample:
        u = fine_sample_u(sample_num).expand(*cdf.shape[:-1], sample_num)
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp_min(0)
    hi = hi.clamp_max(cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    e0, e1 = torch.gather(edges, -1, lo), torch.gather(edges, -1, hi)
    den = c1 - c0
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    tf = e0 + (u - c0) / den * (e1 - e0)
    return -1 / (tf * (b - a) + a) if inv_mode else tf


# --------------------------------------------------------------------------------------------------
# dist_decoder.py


def dist_decoder_forward(W, pre, feats, use_vis, bias_val=0.05):
    """dist_decoder.py:99-107.  `pre` = 'dist_decoder' | 'fine_dist_decoder'."""
    def trunk(name):
        h = _elu(_lin(W, f"{pre}.{name}.0", feats))
        h = _elu(_lin(W, f"{pre}.{name}.2", h))
        return _lin(W, f"{pre}.{name}.4", h)
    mean = _softplus(trunk("mean_decoder"))
    var = _softplus(trunk("var_decoder")) + bias_val                            # AddBias (ops.py:78-84)
    aw = _sigmoid(trunk("aw_decoder"))
    vis = _sigmoid(trunk("vis_decoder")) if use_vis else None
    return mean, var, vis, aw


def get_near_far_points(depth, interval, depth_range, is_ref):
    """dist_decoder.py:6-51 (fixed_interval=False)."""
    shape = (-1,) + (1,) * (depth.dim() - 1)
    a = (-1 / depth_range[:, 0]).reshape(shape)
    b = (-1 / depth_range[:, 1]).reshape(shape)
    t = (-1 / depth.clamp(min=1e-5) - a) / (b - a)
    half = interval / 2
    if is_ref:
        before = torch.cat([half[..., :1], half[..., :-1]], -1)
        return t - before, t + half
    first = t[..., :1] - half[..., :1]
    last = t[..., -1:] + half[..., -1:]
    edges = torch.cat([first, (t[..., :-1] + t[..., 1:]) / 2, last], -1)
    return edges[..., :-1], edges[..., 1:]


def compute_prob(depth, interval, mean, var, vis, aw, is_ref, depth_range, use_vis):
    """dist_decoder.py:109-140."""
    lo, hi = get_near_far_points(depth, interval, depth_range, is_ref)
    mix = torch.cat([aw, 1 - aw], -1)
    c0 = 0.5 + 0.5 * torch.tanh((lo[..., None] - mean) * var)
    c1 = 0.5 + 0.5 * torch.tanh((hi[..., None] - mean) * var)
    if use_vis:
        c0, c1 = c0 * vis, c1 * vis
    visibility = ((1 - c0) * mix).sum(-1)
    hit = ((c1 - c0) * mix).sum(-1)
    alpha = torch.log(hit / (visibility - hit + 1e-5) + 1e-5)
    return alpha, visibility, hit


# --------------------------------------------------------------------------------------------------
# ibrnet.py / aggregate_net.py


def posenc_table(n_samples, d_hid=16):
    """ibrnet.py:305-313 (float64 numpy table cast to fp32)."""
    pos = np.arange(n_samples, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).float()[None]


def weighted_mean_var(x, w):
    """ibrnet.py:112-116."""
    m = (x * w).sum(2, keepdim=True)
    return m, (w * (x - m) ** 2).sum(2, keepdim=True)


def ray_attention(W, pre, x, row_mask):
    """ibrnet.py:52-102 + 7-27: 4 heads, d_k = d_v = 4, temperature 2, post-LN (eps 1e-6)."""
    B, L, _ = x.shape
    nh, dk = 4, 4
    q = (x @ W[f"{pre}.w_qs.weight"].t()).view(B, L, nh, dk).transpose(1, 2)
    k = (x @ W[f"{pre}.w_ks.weight"].t()).view(B, L, nh, dk).transpose(1, 2)
    v = (x @ W[f"{pre}.w_vs.weight"].t()).view(B, L, nh, dk).transpose(1, 2)
    logits = (q / (dk ** 0.5)) @ k.transpose(2, 3)                     # [B,nh,Lq,Lk]
    # mask [B,L,1] -> [B,1,L,1]: broadcasts over KEYS, i.e. it blanks whole query rows (ibrnet.py:20)
    logits = logits.masked_fill(row_mask[:, None] == 0, -1e9)
    a = torch.softmax(logits, -1)
    o = (a @ v).transpose(1, 2).reshape(B, L, nh * dk)
    o = o @ W[f"{pre}.fc.weight"].t() + x
    return F.layer_norm(o, (16,), W[f"{pre}.layer_norm.weight"], W[f"{pre}.layer_norm.bias"], 1e-6)


def ibrnet_forward(W, pre, rgb_feat, neuray_feat, ray_diff, mask, pos_enc):
    """ibrnet.py:315-369.  rgb_feat [R,dn,rfn,35], neuray_feat [...,32], ray_diff [...,4], mask [...,1]."""
    nv = rgb_feat.shape[2]
    dfeat = _elu(_lin(W, f"{pre}.ray_dir_fc.2", _elu(_lin(W, f"{pre}.ray_dir_fc.0", ray_diff))))
    rgb_in = rgb_feat[..., :3]
    rgb_feat = rgb_feat + dfeat
    weight = mask / (mask.sum(2, keepdim=True) + 1e-8)

    gate = _lin(W, f"{pre}.neuray_fc.2", _elu(_lin(W, f"{pre}.neuray_fc.0", neuray_feat)))
    weight0 = _sigmoid(gate) * weight
    m0, v0 = weighted_mean_var(rgb_feat, weight0)
    m1, v1 = weighted_mean_var(rgb_feat, weight)
    glob = torch.cat([m0, v0, m1, v1], -1).expand(-1, -1, nv, -1)

    x = torch.cat([glob, rgb_feat, neuray_feat], -1)                    # 140 + 35 + 32 = 207
    x = _elu(_lin(W, f"{pre}.base_fc.2", _elu(_lin(W, f"{pre}.base_fc.0", x))))

    xv = _elu(_lin(W, f"{pre}.vis_fc.2", _elu(_lin(W, f"{pre}.vis_fc.0", x * weight))))
    x_res, vis = xv[..., :-1], xv[..., -1:]
    vis = _sigmoid(vis) * mask
    x = x + x_res
    vis = _sigmoid(_lin(W, f"{pre}.vis_fc2.2", _elu(_lin(W, f"{pre}.vis_fc2.0", x * vis)))) * mask
    weight = vis / (vis.sum(2, keepdim=True) + 1e-8)

    m, v = weighted_mean_var(x, weight)
    g = torch.cat([m.squeeze(2), v.squeeze(2), weight.mean(2)], -1)     # 65
    g = _elu(_lin(W, f"{pre}.geometry_fc.2", _elu(_lin(W, f"{pre}.geometry_fc.0", g))))
    n_valid = mask.sum(2)                                                # [R,dn,1]
    g = g + pos_enc
    g = ray_attention(W, f"{pre}.ray_attention", g, (n_valid > 1).float())
    sigma = torch.relu(_lin(W, f"{pre}.out_geometry_fc.2", _elu(_lin(W, f"{pre}.out_geometry_fc.0", g))))
    sigma = sigma.masked_fill(n_valid < 1, 0.0)

    c = torch.cat([x, vis, ray_diff], -1)                                # 37
    c = _elu(_lin(W, f"{pre}.rgb_fc.0", c))
    c = _elu(_lin(W, f"{pre}.rgb_fc.2", c))
    c = _lin(W, f"{pre}.rgb_fc.4", c)
    c = c.masked_fill(mask == 0, -1e9)
    blend = torch.softmax(c, 2)
    return torch.cat([(rgb_in * blend).sum(2), sigma], -1)


def agg_net_forward(W, pre, prj, que_dir, pos_enc):
    """aggregate_net.py:34-68.  Returns density [qn,rn,dn], colours [qn,rn,dn,3]."""
    rfn, qn, rn, dn, _ = prj["mask"].shape
    hit = (prj["hit_prob"] - 0.5) * 2
    vis = (prj["vis"] - 0.5) * 2
    emb = torch.cat([prj["ray_feats"], hit, vis], -1)
    emb = _lin(W, f"{pre}.prob_embed.2", torch.relu(_lin(W, f"{pre}.prob_embed.0", emb)))

    def to_rays(t):  # [rfn,qn,rn,dn,c] -> [qn*rn,dn,rfn,c]
        return t.reshape(rfn, qn * rn, dn, -1).permute(1, 2, 0, 3)
    diff = prj["dir"] - que_dir[None]
    dot = (prj["dir"] * que_dir[None]).sum(-1, keepdim=True)
    out = ibrnet_forward(W, f"{pre}.agg_impl", to_rays(torch.cat([prj["rgb"], prj["img_feats"]], -1)),
                         to_rays(emb), to_rays(torch.cat([diff, dot], -1)), to_rays(prj["mask"]), pos_enc)
    return out[..., 3].reshape(qn, rn, dn), out[..., :3].reshape(qn, rn, dn, 3)


# --------------------------------------------------------------------------------------------------
# renderer.py


DEFAULT_CFG = {
    "use_hierarchical_sampling": False, "fine_depth_sample_num": 64, "fine_depth_use_all": False,
    "ray_batch_num": 2048, "depth_sample_num": 64, "use_ray_mask": True, "ray_mask_view_num": 2,
    "ray_mask_point_num": 8, "render_depth": False,
    "dist_decoder_use_vis": True, "fine_dist_decoder_use_vis": True,
    "agg_sample_num": 64, "fine_agg_sample_num": 64, "dist_decoder_bias_val": 0.05, "fine_dist_decoder_bias_val": 0.05,
}


def render_by_depth(W, cfg, que_depth, que, ref, is_train, is_fine, keep=None):
    """renderer.py:168-203 (+ predict_proj_ray_prob :67-83, get_img_feats :127-135, network_rendering :157-166).

    `keep`: optional dict that receives intermediate tensors (for stage-level parity tests).
    """
    que_dists = depth2inv_dists(que_depth, que["depth_range"])
    que_pts, que_dir = depth2points(que, que_depth)
    prj = project_points_dict(ref, que_pts)
    rfn, qn, rn, dn, _ = prj["mask"].shape

    dec = "fine_dist_decoder" if is_fine else "dist_decoder"
    mean, var, vis, aw = dist_decoder_forward(W, dec, prj["ray_feats"], cfg[dec + "_use_vis"], cfg[dec + "_bias_val"])
    # compute_prob is always the COARSE decoder's method (renderer.py:75) -> its use_vis flag decides
    _, visibility, hit = compute_prob(prj["depth"].squeeze(-1), que_dists[None], mean, var, vis, aw, True,
                                      ref["depth_range"], cfg["dist_decoder_use_vis"])
    prj["vis"] = visibility.reshape(rfn, qn, rn, dn, 1) * prj["mask"]
    prj["hit_prob"] = hit.reshape(rfn, qn, rn, dn, 1) * prj["mask"]

    h, w = ref["imgs"].shape[-2:]
    prj["img_feats"] = interpolate_feature_map(ref["img_feats"], prj["pts"].reshape(rfn, -1, 2),
                                               prj["mask"].reshape(rfn, -1), h, w).reshape(rfn, qn, rn, dn, -1)

    agg = "fine_agg_net" if is_fine else "agg_net"
    pos_enc = posenc_table(cfg["fine_agg_sample_num" if is_fine else "agg_sample_num"])
    density, colors = agg_net_forward(W, agg, prj, que_dir, pos_enc)
    alpha = 1.0 - torch.exp(-torch.relu(density))
    hit_prob = alpha_values2hit_prob(alpha)
    out = {"pixel_colors_nr": (hit_prob[..., None] * colors).sum(2), "hit_prob_nr": hit_prob}
    if "imgs" in que:
        out["pixel_colors_gt"] = bilinear_sample(que["imgs"], que["coords"], align_corners=True)
    if cfg["use_ray_mask"]:
        seen = prj["mask"].int().sum(0) > cfg["ray_mask_view_num"]          # qn,rn,dn,1
        out["ray_mask"] = (seen.sum(2) > cfg["ray_mask_point_num"])[..., 0]
    if cfg["render_depth"]:
        out["render_depth"] = (hit_prob * que_depth).sum(-1)
    if keep is not None:
        keep.update({"que_dists": que_dists, "que_pts": que_pts, "que_dir": que_dir, "density": density,
                     "colors": colors, **{"prj_" + k: v for k, v in prj.items()}})
    return out


def predict_self_hit_prob(W, cfg, que, que_depth, que_dists, is_fine):
    """renderer.py:137-155: the query view's own ray_feats decoded along its rays (fine-tuning configs).
    que must hold 'ray_feats' [qn,32,fh,fw] and 'imgs'."""
    dec = "fine_dist_decoder" if is_fine else "dist_decoder"
    h, w = que["imgs"].shape[-2:]
    mask = torch.ones(que["coords"].shape[:2])
    feats = interpolate_feature_map(que["ray_feats"], que["coords"], mask, h, w)          # qn,rn,32
    mean, var, vis, aw = dist_decoder_forward(W, dec, feats, cfg[dec + "_use_vis"], cfg[dec + "_bias_val"])
    un = lambda t: None if t is None else t.unsqueeze(2)
    # the decoder's OWN compute_prob here (renderer.py:146), so its own use_vis flag
    _, _, hit = compute_prob(que_depth, que_dists, un(mean), un(var), un(vis), un(aw), False, que["depth_range"], cfg[dec + "_use_vis"])
    return hit


def render_impl(W, cfg, que, ref, is_train, fine_u=None, fine_depth_override=None):
    """renderer.py:205-226.  `fine_u`: training-time uniforms for sample_fine_depth; `fine_depth_override`
    injects externally computed (sorted) fine-pass depths (searchsorted is discontinuous)."""
    que_depth, _ = sample_depth(que["depth_range"], que["coords"], cfg["depth_sample_num"], False)
    out = render_by_depth(W, cfg, que_depth, que, ref, is_train, False)
    if cfg["use_hierarchical_sampling"]:
        if fine_depth_override is not None:
            fd = fine_depth_override
        else:
            fd = sample_fine_depth(que_depth, out["hit_prob_nr"].detach(), que["depth_range"],
                                   cfg["fine_depth_sample_num"], is_train, fine_u)
            if cfg["fine_depth_use_all"]:
                fd = torch.cat([que_depth, fd], -1)
            fd = torch.sort(fd, -1)[0]
        out["que_depth_fine"] = fd                                              # oracle-only extra
        for k, v in render_by_depth(W, cfg, fd, que, ref, is_train, True).items():
            out[k + "_fine"] = v
    out["que_depth"] = que_depth                                                # oracle-only extra
    return out


def render(W, cfg, que, ref, is_train, ray_batch_num=None):
    """renderer.py:237-254: python chunk loop (image_encoder / vis_encoder are upstream, not part of the path)."""
    que = dict(que)
    coords = que["coords"]
    step = ray_batch_num or cfg["ray_batch_num"]
    chunks = {}
    for s in range(0, coords.shape[1], step):
        que["coords"] = coords[:, s:s + step]
        for k, v in render_impl(W, cfg, que, ref, is_train).items():
            if is_train or not k.startswith("hit_prob"):
                chunks.setdefault(k, []).append(v)
    return {k: torch.cat(v, 1) for k, v in chunks.items()}


# --------------------------------------------------------------------------------------------------
# DepthInitNet.get_diff_feats (network/init_net.py:14-61)

def masked_mean_var(feats, mask, dim):
    """ops.py:36-41."""
    mask = mask.float()
    mask_sum = torch.clamp_min(torch.sum(mask, dim, keepdim=True), min=1e-4)
    mean = torch.sum(feats * mask, dim, keepdim=True) / mask_sum
    var = torch.sum((feats - mean) ** 2 * mask, dim, keepdim=True) / mask_sum
    return mean, var


def depth2pts3d(depth, ref_Ks, ref_poses):
    """init_net.py:13-27: every pixel of every view lifted to the world with its depth.  depth [rfn,1,h,w] -> [rfn,h*w,3]."""
    rfn, dn, h, w = depth.shape
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    hom = torch.stack([xs, ys, torch.ones_like(xs)], -1).float()                   # h,w,3 = (x, y, 1)
    pts = depth.permute(0, 2, 3, 1).unsqueeze(-1) * hom[None, :, :, None, :]        # rfn,h,w,dn,3
    pts = pts.reshape(rfn, h * w * dn, 3).permute(0, 2, 1)
    pts = torch.inverse(ref_Ks) @ pts
    R = ref_poses[:, :, :3].permute(0, 2, 1)
    t = -R @ ref_poses[:, :, 3:]
    return (R @ pts + t).permute(0, 2, 1)


def get_diff_feats(ref, depth_in):
    """init_net.py:29-61.  ref: imgs [rfn,3,h,w], poses, Ks, depth_range; depth_in [rfn,1,h,w] in [0,1] -> [rfn,8,h,w]."""
    imgs = ref["imgs"]
    near = ref["depth_range"][:, 0][:, None, None]
    far = ref["depth_range"][:, 1][:, None, None]
    near_inv, far_inv = -1 / near[..., None], -1 / far[..., None]
    depth = -1 / (depth_in * (far_inv - near_inv) + near_inv)
    rfn, _, h, w = imgs.shape
    pts3d = depth2pts3d(depth, ref["Ks"], ref["poses"])
    _, pts2d, dpt_prj, valid = project_points_ref_views(ref, pts3d.reshape(-1, 3))
    dpt_int = bilinear_sample(depth, pts2d, padding_mode="border", align_corners=True)
    rgb_int = bilinear_sample(imgs, pts2d, padding_mode="border", align_corners=True)
    rgb_diff = torch.abs(rgb_int - imgs.permute(0, 2, 3, 1).reshape(1, rfn * h * w, 3))
    dpt_diff = torch.abs(-1 / torch.clamp(dpt_int, min=1e-5) + 1 / torch.clamp(dpt_prj, min=1e-5))
    dpt_diff = torch.clamp(dpt_diff / ((-1 / far) - (-1 / near)), max=1.5)         # the range of the view projected into
    valid = valid.float().unsqueeze(-1)
    dm, dv = masked_mean_var(dpt_diff, valid, 0)
    rm, rv = masked_mean_var(rgb_diff, valid, 0)
    shape = lambda t, c: t.reshape(rfn, h, w, c).permute(0, 3, 1, 2)
    return torch.cat([shape(rm, 3), shape(rv, 3), shape(dm, 1), shape(dv, 1)], 1)


# --------------------------------------------------------------------------------------------------
# encoders upstream of the ray path: functional restatement of
#   ResUNetLight(3, [1,2,6,4], 32, inplanes=16)   network/ops.py:150-230 (built at renderer.py:59)
#   DefaultVisEncoder                              network/vis_encoder.py:6-21
# pinned by tests/golden/encoders.npz (oracle/gen_golden_encoders.py runs the unmodified modules).


def _conv2d(x, w, b=None, stride=1, pad=None):
    """nn.Conv2d(k, stride, padding=(k-1)//2 unless given, padding_mode='reflect') (ops.py:129-134, conv3x3 / conv1x1)."""
    p = (w.shape[-1] - 1) // 2 if pad is None else pad
    if p:
        x = F.pad(x, (p, p, p, p), mode="reflect")
    return F.conv2d(x, w, b, stride=stride)


def _inorm(W, name, x, eps=1e-5):
    """nn.InstanceNorm2d(C, track_running_stats=False, affine=True): per-(image, channel) biased statistics."""
    m = x.mean((2, 3), keepdim=True)
    v = ((x - m) ** 2).mean((2, 3), keepdim=True)
    return (x - m) / torch.sqrt(v + eps) * W[name + ".weight"][None, :, None, None] + W[name + ".bias"][None, :, None, None]


def _basic_block(W, pre, x, stride):
    """BasicBlock.forward (ops.py:107-124)."""
    out = torch.relu(_inorm(W, pre + ".bn1", _conv2d(x, W[pre + ".conv1.weight"], None, stride)))
    out = _inorm(W, pre + ".bn2", _conv2d(out, W[pre + ".conv2.weight"]))
    if pre + ".downsample.0.weight" in W:
        x = _inorm(W, pre + ".downsample.1", _conv2d(x, W[pre + ".downsample.0.weight"], None, stride))
    return torch.relu(out + x)


def _conv_bn_elu(W, pre, x):
    """`conv` module (ops.py:126-138): ELU(IN(conv3x3 reflect + bias))."""
    return F.elu(_inorm(W, pre + ".bn", _conv2d(x, W[pre + ".conv.weight"], W[pre + ".conv.bias"])))


def _skipconnect(x1, x2):
    """ops.py:199-208: zero-pad the skip to the upsampled size, upsampled first."""
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    return torch.cat([x2, x1], 1)


def res_unet_light(W, pre, imgs, blocks=(1, 2, 6), first_pad=None):
    """ResUNetLight.forward (ops.py:210-228); W keyed by state-dict names under `pre` ('' or 'image_encoder.').
    ResEncoder.forward (ops.py:296-312) is the same graph with blocks (2, 2, 2) and an 8x8 stride-2 first conv of padding 2."""
    W = {k[len(pre):]: v for k, v in W.items() if k.startswith(pre)}
    x = torch.relu(_inorm(W, "bn1", _conv2d(imgs, W["conv1.weight"], None, 2, first_pad)))
    feats = []
    for li, nb in enumerate(blocks):
        for bi in range(nb):
            x = _basic_block(W, f"layer{li + 1}.{bi}", x, 2 if bi == 0 else 1)
        feats.append(x)
    x1, x2, x3 = feats
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    x = _conv_bn_elu(W, "upconv3.conv", up(x3))
    x = _conv_bn_elu(W, "iconv3", _skipconnect(x2, x))
    x = _conv_bn_elu(W, "upconv2.conv", up(x))
    x = _conv_bn_elu(W, "iconv2", _skipconnect(x1, x))
    return F.conv2d(x, W["out_conv.weight"], W["out_conv.bias"])


def vis_encoder(W, pre, ray_feats, img_feats):
    """DefaultVisEncoder.forward (vis_encoder.py:19-21) with ResidualBlock (ops.py:43-76, use_norm, no shortcut conv)."""
    W = {k[len(pre):]: v for k, v in W.items() if k.startswith(pre)}
    x = _conv2d(torch.cat([img_feats, ray_feats], 1), W["out_conv.0.weight"])
    for i in (1, 2):
        p = f"out_conv.{i}.conv"
        t = _conv2d(torch.relu(_inorm(W, p + ".0", x)), W[p + ".2.weight"])
        t = _conv2d(torch.relu(_inorm(W, p + ".3", t)), W[p + ".5.weight"])
        x = t + x
    return F.conv2d(x, W["out_conv.3.weight"])


def encoder_test_weights(template, seed):
    """Deterministic non-trivial parameters for a state dict `template` ({name: shape}): conv weights ~ N(0, 2/fan_in),
    norm weights ~ U(0.5, 1.5), biases ~ N(0, 0.1).  numpy RandomState keeps the stream stable across versions, so the
    goldens store only inputs and outputs, not the 2 M parameters."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in template.items():
        shape = tuple(shape)
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            a = rs.standard_normal(shape) * math.sqrt(2.0 / fan_in)
        elif name.endswith(".weight"):
            a = rs.uniform(0.5, 1.5, shape)
        else:
            a = rs.standard_normal(shape) * 0.1
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


# --------------------------------------------------------------------------------------------------
# training extras: network/renderer.py:280-316 and network/loss.py:17-132,
# pinned by tests/golden/losses.npz (oracle/gen_golden_losses.py runs the unmodified classes).


def predict_mean(W, pre, ray_feats, coords, h, w):
    """renderer.py:291-296: interpolate_feature_map(ray_feats, coords, ones, h, w) + dist_decoder.predict_mean -> [rfn,pn,2]."""
    mask = torch.ones(coords.shape[:2], dtype=torch.float32)
    f = interpolate_feature_map(ray_feats, coords.float(), mask, h, w)
    x = _elu(_lin(W, pre + ".mean_decoder.0", f))
    x = _elu(_lin(W, pre + ".mean_decoder.2", x))
    return _softplus(_lin(W, pre + ".mean_decoder.4", x))


def render_loss(pr, gt, ray_mask=None):
    """RenderLoss.compute_loss (loss.py:58-66)."""
    loss = torch.sum((pr - gt) ** 2, -1)
    if ray_mask is None:
        return torch.mean(loss, 1)
    m = ray_mask.float()
    return torch.sum(loss * m, 1) / (torch.sum(m, 1) + 1e-3)


def depth_loss(depth_pr, coords, true_depth, depth_range, loss_type="l2", beta=0.05, aug_depth=None, thresh=0.02):
    """DepthLoss.__call__ (loss.py:92-127); aug_depth = the 'gso' branch."""
    rfn, _, h, w = true_depth.shape
    near, far = -1 / depth_range[:, 0:1], -1 / depth_range[:, 1:2]

    def process(d):
        d = -1 / torch.clamp(d, min=1e-5)
        return torch.clamp((d - near) / (far - near), min=0, max=1.0)

    at = lambda m: bilinear_sample(m, coords.float(), h, w, padding_mode="border", align_corners=True)[..., 0]
    gt = process(at(true_depth))
    if loss_type == "l2":
        loss = (gt - depth_pr) ** 2
    else:
        d = (gt - depth_pr).abs()
        loss = torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta)
    if aug_depth is None:
        return torch.mean(loss, 1)
    m = ((process(at(aug_depth)) - gt).abs() < thresh).float()
    return torch.sum(loss * m, 1) / (torch.sum(m, 1) + 1e-4)


def consistency_loss(prob0, prob1):
    """ConsistencyLoss.__call__ (loss.py:30-37)."""
    ce = -prob0 * torch.log(prob1 + 1e-5) - (1 - prob0) * torch.log(1 - prob1 + 1e-5)
    return torch.mean(torch.mean(ce, -1), 1)


def extract_depth_for_init(depth_range, depth):
    """init_net.py:63-74: metric depth [rfn,1,h,w] -> normalised inverse depth in [0, 1]."""
    near, far = -1 / depth_range[:, 0][:, None, None, None], -1 / depth_range[:, 1][:, None, None, None]
    d = -1 / torch.clamp(depth, min=1e-5)
    return torch.clamp((d - near) / (far - near), min=0, max=1.0)


def depth_init_net(W, pre, ref, diff=None):
    """DepthInitNet.forward (init_net.py:93-101); ref = {imgs, depth, depth_range, poses, Ks}; W under `pre` ('' or 'init_net.').
    diff: get_diff_feats' output when the caller already has it (tests that isolate the convolution stack)."""
    W = {k[len(pre):]: v for k, v in W.items() if k.startswith(pre)}
    depth = extract_depth_for_init(ref["depth_range"], ref["depth"])
    if diff is None:
        diff = get_diff_feats(ref, depth)
    feats = res_unet_light(W, "res_net.", torch.cat([ref["imgs"], depth, diff], 1), blocks=(2, 2, 2), first_pad=2)
    d = torch.relu(F.conv2d(depth, W["depth_skip.0.weight"], W["depth_skip.0.bias"], stride=2))
    d = F.conv2d(d, W["depth_skip.2.weight"], W["depth_skip.2.bias"], stride=2)
    return F.conv2d(torch.cat([d, feats], 1), W["conv_out.weight"], W["conv_out.bias"])


# --------------------------------------------------------------------------------------------------
# CostVolumeInitNet's frozen MVSNet: network/mvsnet/mvsnet.py:7-66,124-141,
# network/mvsnet/modules.py:25-70, network/init_net.py:103-168; pinned by tests/golden/mvsnet.npz.
# ABN / InPlaceABN in eval mode = batch_norm with the running statistics + leaky_relu(0.01) (oracle/ref_import.py).


def _abn(W, pre, x):
    x = F.batch_norm(x, W[pre + ".running_mean"], W[pre + ".running_var"], W[pre + ".weight"], W[pre + ".bias"], False, 0.0, 1e-5)
    return F.leaky_relu(x, 0.01)


def mvs_feature_net(W, pre, x):
    """FeatureNet.forward (mvsnet.py:25-29): zero-padded 3x3 / 5x5 convolutions + ABN, a final biased 3x3."""
    for name, stride, pad in (("conv0", 1, 1), ("conv1", 1, 1), ("conv2", 2, 2), ("conv3", 1, 1), ("conv4", 1, 1), ("conv5", 2, 2), ("conv6", 1, 1)):
        x = _abn(W, f"{pre}{name}.bn", F.conv2d(x, W[f"{pre}{name}.conv.weight"], None, stride, pad))
    return F.conv2d(x, W[pre + "feature.weight"], W[pre + "feature.bias"], 1, 1)


def mvs_homo_warp(src_feat, src_proj, ref_proj_inv, depth_values):
    """modules.py:25-63 for one reference view: src_feat [C,h,w], 4x4 matrices, depth_values [D] -> [C,D,h,w]."""
    C, h, w = src_feat.shape
    T = src_proj @ ref_proj_inv
    R, t = T[:3, :3], T[:3, 3:]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    grid = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w)], 0)              # [3,hw]
    pts = (R @ (grid[:, None, :] * depth_values.view(1, -1, 1)).reshape(3, -1)) + t            # [3,D*hw]
    z = pts[2:].clone()
    z[z < 1e-4] = 1e-4
    xy = (pts[:2] / z).t()[None]                                                               # [1,D*hw,2] pixel coordinates
    out = bilinear_sample(src_feat[None], xy, h, w, padding_mode="zeros", align_corners=True)  # [1,D*hw,C]
    return out[0].t().reshape(C, -1, h, w)


def mvs_cost_reg(W, pre, vol):
    """CostRegNet.forward (mvsnet.py:52-66) on vol [1,32,D,h,w] -> [1,1,D,h,w]."""
    c3 = lambda name, x, s: _abn(W, f"{pre}{name}.bn", F.conv3d(x, W[f"{pre}{name}.conv.weight"], None, s, 1))
    up = lambda name, x: _abn(W, f"{pre}{name}.1", F.conv_transpose3d(x, W[f"{pre}{name}.0.weight"], None, 2, 1, 1))
    conv0 = c3("conv0", vol, 1)
    conv2 = c3("conv2", c3("conv1", conv0, 2), 1)
    conv4 = c3("conv4", c3("conv3", conv2, 2), 1)
    x = c3("conv6", c3("conv5", conv4, 2), 1)
    x = conv4 + up("conv7", x)
    x = conv2 + up("conv9", x)
    x = conv0 + up("conv11", x)
    return F.conv3d(x, W[pre + "prob.weight"], W[pre + "prob.bias"], 1, 1)


def mvs_project_matrix(ratio, Ks, poses):
    """init_net.py:103-111."""
    S = torch.diag(torch.tensor([ratio, ratio, 1.0]))
    P = S[None] @ Ks @ poses
    return torch.cat([P, torch.tensor([0.0, 0.0, 0.0, 1.0]).view(1, 1, 4).repeat(P.shape[0], 1, 1)], 1)


def mvs_depth_vals(depth_range, dn):
    """init_net.py:162-168."""
    near, far = depth_range[:, 0], depth_range[:, 1]
    interval = (1 / far - 1 / near) / (dn - 1)
    vals = 1 / (1 / near[:, None] + torch.arange(0, dn - 1)[None, :] * interval[:, None])
    return torch.cat([vals, far[:, None]], 1)


def mvs_cost_volume(W, pre, ref, src, dn, is_train):
    """construct_cost_volume_with_src (init_net.py:113-160): returns (softmaxed cost volume [rfn,dn,h/4,w/4], depth [rfn,h/4,w/4])."""
    imgs_r, imgs_s = ref["imgs"], src["imgs"]
    rfn, _, h, w = imgs_r.shape
    ratio = 1.0
    if not is_train and max(h, w) >= 800:
        size = (576, 768) if (h, w) == (768, 1024) else (640, 640) if (h, w) == (800, 800) else None
        if size is not None:
            imgs_r, imgs_s = F.interpolate(imgs_r, size, mode="bilinear"), F.interpolate(imgs_s, size, mode="bilinear")
            ratio = size[0] / h
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    rf = mvs_feature_net(W, pre + "feature.", (imgs_r - mean) / std)
    sf = mvs_feature_net(W, pre + "feature.", (imgs_s - mean) / std)
    rp = mvs_project_matrix(0.25 * ratio, ref["Ks"], ref["poses"])
    sp = mvs_project_matrix(0.25 * ratio, src["Ks"], src["poses"])
    vals = mvs_depth_vals(ref["depth_range"], dn)
    nn_ids = ref["nn_ids"]
    out = []
    for i in range(rfn):
        s = rf[i][:, None].repeat(1, dn, 1, 1)
        q = s ** 2
        inv = torch.inverse(rp[i])
        for k in nn_ids[i].tolist():
            wv = mvs_homo_warp(sf[k], sp[k], inv, vals[i])
            s = s + wv
            q = q + wv ** 2
        n = nn_ids.shape[1] + 1
        var = q / n - (s / n) ** 2
        out.append(mvs_cost_reg(W, pre + "cost_regularization.", var[None])[0, 0])
    cost = torch.stack(out, 0)
    cost[torch.isnan(cost)] = 0
    if ratio != 1.0:
        cost = F.interpolate(cost, (h // 4, w // 4), mode="bilinear")
    cost = F.softmax(cost, 1)
    return cost, torch.sum(cost * vals.view(rfn, dn, 1, 1), 1)


def mvs_test_weights(template, seed):
    """Seeded MVSNet parameters: conv weights ~ N(0, 2/fan_in), norm weights U(0.5,1.5), running_var U(0.5,1.5), rest N(0,0.1)."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in template.items():
        shape = tuple(shape)
        if len(shape) >= 4:
            a = rs.standard_normal(shape) * math.sqrt(2.0 / (int(np.prod(shape[1:])) if "conv" in name and ".0.weight" not in name else int(np.prod(shape[2:])) * shape[0]))
        elif name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
            a = rs.uniform(0.5, 1.5, shape)
        else:
            a = rs.standard_normal(shape) * 0.1
        if "prob.weight" in name:
            a = a * 0.05          # keep the logits of the depth softmax moderate: a saturated (one-hot) volume would test an argmax
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def _conv_stack(W, pre, x, nrb=1):
    """conv3x3 -> ResidualBlock x nrb -> conv1x1 (init_net.py:230-249; ops.py:43-76), all reflect / bias-free."""
    x = _conv2d(x, W[f"{pre}.0.weight"])
    for i in range(1, nrb + 1):
        p = f"{pre}.{i}.conv"
        t = _conv2d(torch.relu(_inorm(W, p + ".0", x)), W[p + ".2.weight"])
        t = _conv2d(torch.relu(_inorm(W, p + ".3", t)), W[p + ".5.weight"])
        x = t + x
    return F.conv2d(x, W[f"{pre}.{nrb + 1}.weight"])


def cost_volume_init_net(W, pre, ref, src, is_train, sn=64):
    """CostVolumeInitNet.forward (init_net.py:247-254); W under `pre` ('' or 'init_net.'), ref carries nn_ids."""
    W = {k[len(pre):]: v for k, v in W.items() if k.startswith(pre)}
    cost, depth = mvs_cost_volume(W, "mvsnet.", ref, src, sn, is_train)
    ref_feats = res_unet_light(W, "res_net.", ref["imgs"], blocks=(2, 3, 6))
    volume_feats = _conv_stack(W, "volume_conv2d", cost)
    d = extract_depth_for_init(ref["depth_range"], depth[:, None])
    depth_feats = _conv_stack(W, "depth_conv", d)
    return _conv_stack(W, "out_conv", torch.cat([ref_feats, torch.cat([volume_feats, depth_feats], 1)], 1))
