"""TEST INFRASTRUCTURE: PyTorch restatement of one render pass (same arithmetic as reference network/renderer.py:168-203 and
the modules it calls) and of predict_self_hit_prob, as differentiable torch ops that run on any device.  Used by the tests
as the GPU-side A/B reference of the native backward (autograd over these functions).  Nothing under neuray_b200/ imports
this file.
"""
import os

import torch
import torch.nn.functional as F


def _lin(P, name, x):
    w = P[name + ".weight"]
    b = P.get(name + ".bias")
    return F.linear(x, w, b)


def _bilinear(feats, pts, h, w, align_corners):
    """interpolate_feats with padding 'border' (reference ops.py:14-34): feats [b,c,fh,fw], pts [b,n,2] -> [b,n,c]."""
    gx = pts[..., 0] / (w - 1) * 2 - 1
    gy = pts[..., 1] / (h - 1) * 2 - 1
    grid = torch.stack([gx, gy], -1).unsqueeze(1)
    out = F.grid_sample(feats, grid, mode="bilinear", padding_mode="border", align_corners=align_corners)
    return out.squeeze(2).permute(0, 2, 1)


def render_pass_torch(P, dec, agg, cfgv, que_depth, coords, que_pose, que_K, que_range, ref, pos_enc):
    """One render_by_depth pass -> (pixel_colors [1,rn,3], hit_prob [1,rn,dn], render_depth [1,rn]).

    P: {name: tensor} parameters (reference state-dict names); dec/agg: module prefixes; cfgv: dict(use_vis_prob,
    use_vis_head, var_bias); ref: dict(poses, Ks, depth_range, imgs, ray_feats, img_feats)."""
    qn, rn, dn = que_depth.shape
    # depth2inv_dists / depth2points
    a = (-1 / que_range[:, 0])[:, None, None]
    b = (-1 / que_range[:, 1])[:, None, None]
    t = (-1 / que_depth - a) / (b - a)
    que_dists = torch.cat([t[..., 1:] - t[..., :-1], torch.full_like(t[..., :1], 1e6)], -1)
    rot = que_pose[:, :, :3].transpose(1, 2)
    centre = -(rot @ que_pose[:, :, 3:])
    hom = torch.cat([coords, torch.ones_like(coords[..., :1])], -1)
    cam = torch.inverse(que_K)[:, None] @ hom[..., None]
    dirs = (rot[:, None] @ cam + centre[:, None])[..., 0] - centre[:, None, :, 0]
    pts = centre[:, None, None, :, 0] + dirs[:, :, None] * que_depth[..., None]
    que_dir = (-dirs / dirs.norm(dim=2, keepdim=True))[:, :, None].expand(-1, -1, dn, -1)
    # project_points_ref_views
    rfn, _, h, w = ref["imgs"].shape
    flat = pts.reshape(-1, 3)
    KRt = ref["Ks"] @ ref["poses"]
    camp = flat @ KRt[:, :, :3].transpose(1, 2) + KRt[:, None, :, 3]            # [rfn,pn,3]
    z = camp[..., 2:]
    degenerate = z.abs() < 1e-4
    z = torch.where(degenerate, torch.full_like(z, 1e-3), z)
    pix = camp[..., :2] / z
    outside = (pix[..., 0] < -0.5) | (pix[..., 0] >= w - 0.5) | (pix[..., 1] < -0.5) | (pix[..., 1] >= h - 0.5)
    mask = (~degenerate[..., 0] & ~outside).float()                              # [rfn,pn]
    rcentre = -(ref["poses"][:, :, :3].transpose(1, 2) @ ref["poses"][:, :, 3:]).transpose(1, 2)
    d = flat[None] - rcentre
    prj_dir = -d / d.norm(dim=2, keepdim=True).clamp_min(1e-5)
    fh, fw = ref["ray_feats"].shape[-2:]
    al = (fh == h and fw == w)
    m = mask[..., None]
    ray_feats = _bilinear(ref["ray_feats"], pix, h, w, al) * m
    img_feats = _bilinear(ref["img_feats"], pix, h, w, al) * m
    rgb = _bilinear(ref["imgs"], pix, h, w, True) * m
    # dist decoder + compute_prob
    def head(name, nout):
        x = F.elu(_lin(P, f"{dec}.{name}.0", ray_feats))
        x = F.elu(_lin(P, f"{dec}.{name}.2", x))
        return _lin(P, f"{dec}.{name}.4", x)
    mean = F.softplus(head("mean_decoder", 2))
    var = F.softplus(head("var_decoder", 2)) + cfgv["var_bias"]
    aw = torch.sigmoid(head("aw_decoder", 1))
    ra = (-1 / ref["depth_range"][:, 0])[:, None]
    rb = (-1 / ref["depth_range"][:, 1])[:, None]
    tz = ((-1 / z[..., 0].clamp(min=1e-5)) - ra) / (rb - ra)                     # [rfn,pn]
    half = (que_dists / 2).reshape(1, qn * rn, dn)
    before = torch.cat([half[..., :1], half[..., :-1]], -1).reshape(1, -1)
    lo = (tz - before)[..., None]
    hi = (tz + half.reshape(1, -1))[..., None]
    c0 = 0.5 + 0.5 * torch.tanh((lo - mean) * var)
    c1 = 0.5 + 0.5 * torch.tanh((hi - mean) * var)
    if cfgv["use_vis_prob"]:
        visd = torch.sigmoid(head("vis_decoder", 1))
        c0, c1 = c0 * visd, c1 * visd
    mix = torch.cat([aw, 1 - aw], -1)
    vis = ((1 - c0) * mix).sum(-1, keepdim=True) * m
    hit = ((c1 - c0) * mix).sum(-1, keepdim=True) * m
    # aggregation net
    emb = torch.cat([ray_feats, (hit - 0.5) * 2, (vis - 0.5) * 2], -1)
    emb = _lin(P, f"{agg}.prob_embed.2", F.relu(_lin(P, f"{agg}.prob_embed.0", emb)))
    qd = que_dir.reshape(1, -1, 3)
    ray_diff = torch.cat([prj_dir - qd, (prj_dir * qd).sum(-1, keepdim=True)], -1)
    R = qn * rn

    def rays(x):   # [rfn,pn,c] -> [R,dn,rfn,c]
        return x.reshape(rfn, R, dn, -1).permute(1, 2, 0, 3)
    rgb_feat, nf, rdiff, msk = rays(torch.cat([rgb, img_feats], -1)), rays(emb), rays(ray_diff), rays(m)
    ib = f"{agg}.agg_impl"
    rgb_in = rgb_feat[..., :3]
    rgb_feat = rgb_feat + F.elu(_lin(P, f"{ib}.ray_dir_fc.2", F.elu(_lin(P, f"{ib}.ray_dir_fc.0", rdiff))))
    weight = msk / (msk.sum(2, keepdim=True) + 1e-8)
    weight0 = torch.sigmoid(_lin(P, f"{ib}.neuray_fc.2", F.elu(_lin(P, f"{ib}.neuray_fc.0", nf)))) * weight

    def mv(x, wgt):
        mu = (x * wgt).sum(2, keepdim=True)
        return mu, (wgt * (x - mu) ** 2).sum(2, keepdim=True)
    m0, v0 = mv(rgb_feat, weight0)
    m1, v1 = mv(rgb_feat, weight)
    x = torch.cat([torch.cat([m0, v0, m1, v1], -1).expand(-1, -1, rfn, -1), rgb_feat, nf], -1)
    x = F.elu(_lin(P, f"{ib}.base_fc.2", F.elu(_lin(P, f"{ib}.base_fc.0", x))))
    xv = F.elu(_lin(P, f"{ib}.vis_fc.2", F.elu(_lin(P, f"{ib}.vis_fc.0", x * weight))))
    x = x + xv[..., :-1]
    visw = torch.sigmoid(xv[..., -1:]) * msk
    visw = torch.sigmoid(_lin(P, f"{ib}.vis_fc2.2", F.elu(_lin(P, f"{ib}.vis_fc2.0", x * visw)))) * msk
    weight = visw / (visw.sum(2, keepdim=True) + 1e-8)
    mu, vr = mv(x, weight)
    g = torch.cat([mu.squeeze(2), vr.squeeze(2), weight.mean(2)], -1)
    g = F.elu(_lin(P, f"{ib}.geometry_fc.2", F.elu(_lin(P, f"{ib}.geometry_fc.0", g))))
    nvalid = msk.sum(2)
    g = g + pos_enc[None]
    at = f"{ib}.ray_attention"
    q = F.linear(g, P[f"{at}.w_qs.weight"]).view(R, dn, 4, 4).transpose(1, 2)
    k = F.linear(g, P[f"{at}.w_ks.weight"]).view(R, dn, 4, 4).transpose(1, 2)
    vv = F.linear(g, P[f"{at}.w_vs.weight"]).view(R, dn, 4, 4).transpose(1, 2)
    logits = (q / 2.0) @ k.transpose(2, 3)
    logits = logits.masked_fill(((nvalid > 1).float()[:, None]) == 0, -1e9)
    o = (torch.softmax(logits, -1) @ vv).transpose(1, 2).reshape(R, dn, 16)
    o = F.layer_norm(F.linear(o, P[f"{at}.fc.weight"]) + g, (16,), P[f"{at}.layer_norm.weight"], P[f"{at}.layer_norm.bias"], 1e-6)
    sigma = F.relu(_lin(P, f"{ib}.out_geometry_fc.2", F.elu(_lin(P, f"{ib}.out_geometry_fc.0", o))))
    sigma = sigma.masked_fill(nvalid < 1, 0.0)
    c = torch.cat([x, visw, rdiff], -1)
    c = _lin(P, f"{ib}.rgb_fc.4", F.elu(_lin(P, f"{ib}.rgb_fc.2", F.elu(_lin(P, f"{ib}.rgb_fc.0", c)))))
    blend = torch.softmax(c.masked_fill(msk == 0, -1e9), 2)
    colors = (rgb_in * blend).sum(2).reshape(qn, rn, dn, 3)
    alpha = 1.0 - torch.exp(-F.relu(sigma[..., 0].reshape(qn, rn, dn)))
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    hit_prob = alpha * T
    return (hit_prob[..., None] * colors).sum(2), hit_prob, (hit_prob * que_depth).sum(-1)


def self_hit_prob_torch(P, dec, use_vis_prob, var_bias, que_ray_feats, coords, h, w, que_depth, que_range):
    """predict_self_hit_prob (reference renderer.py:137-155 + dist_decoder.py compute_prob with is_ref=False): the query
    view's own visibility features decoded along its rays (fine-tuning configs).  PyTorch restatement: the A/B reference of
    nr_self_hit_prob.
    que_ray_feats [qn,32,fh,fw], coords [qn,rn,2], que_depth [qn,rn,dn] -> hit_prob_self [qn,rn,dn]."""
    fh, fw = que_ray_feats.shape[-2:]
    feats = _bilinear(que_ray_feats, coords, h, w, fh == h and fw == w)                      # [qn,rn,32]

    def head(name):
        x = F.elu(_lin(P, f"{dec}.{name}.0", feats))
        x = F.elu(_lin(P, f"{dec}.{name}.2", x))
        return _lin(P, f"{dec}.{name}.4", x)
    mean = F.softplus(head("mean_decoder"))[:, :, None]                                       # [qn,rn,1,2]
    var = (F.softplus(head("var_decoder")) + var_bias)[:, :, None]
    aw = torch.sigmoid(head("aw_decoder"))[:, :, None]
    a = (-1 / que_range[:, 0])[:, None, None]
    b = (-1 / que_range[:, 1])[:, None, None]
    t = (-1 / que_depth.clamp(min=1e-5) - a) / (b - a)
    dists = torch.cat([t[..., 1:] - t[..., :-1], torch.full_like(t[..., :1], 1e6)], -1)      # depth2inv_dists
    half = dists / 2
    edges = torch.cat([t[..., :1] - half[..., :1], (t[..., :-1] + t[..., 1:]) / 2, t[..., -1:] + half[..., -1:]], -1)
    lo, hi = edges[..., :-1, None], edges[..., 1:, None]
    c0 = 0.5 + 0.5 * torch.tanh((lo - mean) * var)
    c1 = 0.5 + 0.5 * torch.tanh((hi - mean) * var)
    if use_vis_prob:
        visd = torch.sigmoid(head("vis_decoder"))[:, :, None]
        c0, c1 = c0 * visd, c1 * visd
    mix = torch.cat([aw, 1 - aw], -1)
    return ((c1 - c0) * mix).sum(-1)
