"""Native backward of one render pass: `nr_render_pass_bwd` fills the tapes, the weight gradients are GEMMs over them.

Reference: the gradients `loss.backward()` sends through `NeuralRayBaseRenderer.render_by_depth` (renderer.py:168-203)
into `dist_decoder` / `agg_net` parameters and the encoder outputs `ray_feats` / `img_feats`.
"""
import ctypes as C

import torch

from . import _lib

HEADS = ("mean_decoder", "var_decoder", "aw_decoder", "vis_decoder")
_WS = {}


def _ws(key, numel, dev):
    buf = _WS.get((key, str(dev)))
    if buf is None or buf.numel() < numel:
        _WS.pop((key, str(dev)), None)
        buf = torch.empty(numel, dtype=torch.float32, device=dev)
        _WS[(key, str(dev))] = buf
    return buf[:numel]


def tape_shapes(rfn, n_points):
    s = _lib.bwd_slot
    rows = rfn * n_points
    return {"tape_row": (s("R_SLOTS"), rows), "grad_row": (s("G_SLOTS"), rows), "tape_point": (s("P_SLOTS"), n_points),
            "grad_point": (s("GP_SLOTS"), n_points)}


def assemble_param_grads(names, dec, agg, n_heads, tape_row, grad_row, tape_point, grad_point):
    """dW = dz @ x^T for every Linear of the pass (bias = row sums of dz), from the tapes of nr_render_pass_bwd.
    names: parameter names of the pass (state-dict names).  Returns {name: grad or None}."""
    s = _lib.bwd_slot
    TR, GR, TP, GP = tape_row, grad_row, tape_point, grad_point

    def lin(out, g_tape, g_slot, n_out, x_tape, x_slot, n_in, w, b):
        dz = g_tape[g_slot:g_slot + n_out]
        out[w] = dz @ x_tape[x_slot:x_slot + n_in].t()
        if b is not None:
            out[b] = dz.sum(1)

    out = {}
    for hd, head in enumerate(HEADS):
        if f"{dec}.{head}.0.weight" not in names:
            continue
        if hd >= n_heads:        # a head the pass does not evaluate (fine decoder's vis head under a coarse use_vis=False)
            continue
        n_out = 2 if hd < 2 else 1
        lin(out, GR, s("G_DD0") + 32 * hd, 32, TR, s("R_RF"), 32, f"{dec}.{head}.0.weight", f"{dec}.{head}.0.bias")
        lin(out, GR, s("G_DD1") + 32 * hd, 32, TR, s("R_H1") + 32 * hd, 32, f"{dec}.{head}.2.weight", f"{dec}.{head}.2.bias")
        lin(out, GR, s("G_DD2") + 2 * hd, n_out, TR, s("R_H2") + 32 * hd, 32, f"{dec}.{head}.4.weight", f"{dec}.{head}.4.bias")
    ib = f"{agg}.agg_impl"
    lin(out, GR, s("G_PE0"), 32, TR, s("R_RF"), 34, f"{agg}.prob_embed.0.weight", f"{agg}.prob_embed.0.bias")
    lin(out, GR, s("G_PE1"), 32, TR, s("R_P1"), 32, f"{agg}.prob_embed.2.weight", f"{agg}.prob_embed.2.bias")
    lin(out, GR, s("G_RD0"), 16, TR, s("R_DD"), 4, f"{ib}.ray_dir_fc.0.weight", f"{ib}.ray_dir_fc.0.bias")
    lin(out, GR, s("G_RD1"), 35, TR, s("R_R16"), 16, f"{ib}.ray_dir_fc.2.weight", f"{ib}.ray_dir_fc.2.bias")
    lin(out, GR, s("G_NF0"), 8, TR, s("R_NF"), 32, f"{ib}.neuray_fc.0.weight", f"{ib}.neuray_fc.0.bias")
    lin(out, GR, s("G_NF1"), 1, TR, s("R_Q8"), 8, f"{ib}.neuray_fc.2.weight", f"{ib}.neuray_fc.2.bias")
    dz0 = GR[s("G_B0"):s("G_B0") + 64]
    out[f"{ib}.base_fc.0.weight"] = torch.cat([GP[s("GP_B0SUM"):s("GP_B0SUM") + 64] @ TP[s("P_GLOB"):s("P_GLOB") + 140].t(),
                                               dz0 @ TR[s("R_RGBF"):s("R_RGBF") + 67].t()], 1)
    out[f"{ib}.base_fc.0.bias"] = dz0.sum(1)
    lin(out, GR, s("G_B1"), 32, TR, s("R_B1"), 64, f"{ib}.base_fc.2.weight", f"{ib}.base_fc.2.bias")
    lin(out, GR, s("G_V0"), 32, TR, s("R_U"), 32, f"{ib}.vis_fc.0.weight", f"{ib}.vis_fc.0.bias")
    lin(out, GR, s("G_V1"), 33, TR, s("R_VH"), 32, f"{ib}.vis_fc.2.weight", f"{ib}.vis_fc.2.bias")
    lin(out, GR, s("G_V20"), 32, TR, s("R_U2"), 32, f"{ib}.vis_fc2.0.weight", f"{ib}.vis_fc2.0.bias")
    lin(out, GR, s("G_V21"), 1, TR, s("R_WH"), 32, f"{ib}.vis_fc2.2.weight", f"{ib}.vis_fc2.2.bias")
    lin(out, GR, s("G_C0"), 16, TR, s("R_X2"), 37, f"{ib}.rgb_fc.0.weight", f"{ib}.rgb_fc.0.bias")
    lin(out, GR, s("G_C1"), 8, TR, s("R_CH1"), 16, f"{ib}.rgb_fc.2.weight", f"{ib}.rgb_fc.2.bias")
    lin(out, GR, s("G_C2"), 1, TR, s("R_CH2"), 8, f"{ib}.rgb_fc.4.weight", f"{ib}.rgb_fc.4.bias")
    lin(out, GP, s("GP_GEO0"), 64, TP, s("P_GIN"), 65, f"{ib}.geometry_fc.0.weight", f"{ib}.geometry_fc.0.bias")
    lin(out, GP, s("GP_GEO1"), 16, TP, s("P_GH"), 64, f"{ib}.geometry_fc.2.weight", f"{ib}.geometry_fc.2.bias")
    at = f"{ib}.ray_attention"
    lin(out, GP, s("GP_DQ"), 16, TP, s("P_AX"), 16, f"{at}.w_qs.weight", None)
    lin(out, GP, s("GP_DK"), 16, TP, s("P_AX"), 16, f"{at}.w_ks.weight", None)
    lin(out, GP, s("GP_DV"), 16, TP, s("P_AX"), 16, f"{at}.w_vs.weight", None)
    lin(out, GP, s("GP_DFC"), 16, TP, s("P_O"), 16, f"{at}.fc.weight", None)
    dy = GP[s("GP_DLNY"):s("GP_DLNY") + 16]
    out[f"{at}.layer_norm.weight"] = (dy * TP[s("P_XH"):s("P_XH") + 16]).sum(1)
    out[f"{at}.layer_norm.bias"] = dy.sum(1)
    lin(out, GP, s("GP_OG0"), 16, TP, s("P_Y"), 16, f"{ib}.out_geometry_fc.0.weight", f"{ib}.out_geometry_fc.0.bias")
    lin(out, GP, s("GP_OG1"), 1, TP, s("P_T16"), 16, f"{ib}.out_geometry_fc.2.weight", f"{ib}.out_geometry_fc.2.bias")
    return {n: out.get(n) for n in names}


def feat_grads_to_nchw(d_feat):
    """[rfn,fh,fw,64] channel-last gradient -> (d_ray_feats, d_img_feats), both [rfn,32,fh,fw]."""
    g = d_feat.permute(0, 3, 1, 2)
    return g[:, :32].contiguous(), g[:, 32:].contiguous()


def render_pass_backward(p, names, dec, agg, g_pix, g_hit, g_depth, want_feat_grads, feat_shape, stream):
    """p: the NrPassParams of the forward launch (CUDA).  Returns ({name: grad}, d_ray_feats, d_img_feats)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    n_points = p.rn * p.dn
    shapes = tape_shapes(p.rfn, n_points)
    bufs = {k: _ws(k, sh[0] * sh[1], dev).view(sh) for k, sh in shapes.items()}
    d_feat = torch.zeros(feat_shape, dtype=torch.float32, device=dev) if want_feat_grads else None
    b = _lib.NrBwdParams()
    keep = [t.contiguous().float() if t is not None else None for t in (g_pix, g_hit, g_depth)]
    b.d_pixel_colors, b.d_hit_prob, b.d_render_depth = (_lib.ptr(t) for t in keep)
    b.tape_row, b.grad_row = _lib.ptr(bufs["tape_row"]), _lib.ptr(bufs["grad_row"])
    b.tape_point, b.grad_point = _lib.ptr(bufs["tape_point"]), _lib.ptr(bufs["grad_point"])
    b.d_feat = _lib.ptr(d_feat)
    _lib.check(_lib.lib().nr_render_pass_bwd(C.byref(p), C.byref(b), stream), "nr_render_pass_bwd")
    _lib.count_launches(5)
    grads = assemble_param_grads(names, dec, agg, 4 if p.use_vis else 3, bufs["tape_row"], bufs["grad_row"], bufs["tape_point"],
                                 bufs["grad_point"])
    if d_feat is None:
        return grads, None, None
    drf, dimf = feat_grads_to_nchw(d_feat)
    return grads, drf, dimf
