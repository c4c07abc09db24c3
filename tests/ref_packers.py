"""TEST INFRASTRUCTURE: the round-1 PyTorch packers, kept as the independent restatement `nr_pack_weights` (csrc/nr_pack.cu) is
held to bit for bit (tests/test_pack_gpu.py), plus the torch expressions of the reference's camera terms that
`nr_camera_blocks` replaces.  Nothing under neuray_b200/ imports this file.

Packing of the reference's parameters into the flat buffers the kernels stage to shared memory.

Source: any mapping {state-dict name -> tensor} with the reference's names, e.g.
`dict(renderer.named_parameters())` of a reference NeuralRayBaseRenderer, or of neuray_b200.renderer.NeuralRayRenderPath.
The layout (offsets in floats) is owned by the library: `_lib.weight_layout()` (include/neuray_b200.h, NrWeightLayout).
"PyTorch is plumbing": only cat / transpose / copy here, no arithmetic.
"""
import numpy as np
import torch

from neuray_b200 import _lib


def posenc_table(n_samples, d_hid=16):
    """Sinusoid table of reference network/ibrnet.py:305-313 (built in float64 numpy, cast to fp32) -> [n,16]."""
    pos = np.arange(n_samples, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table).float()


def _put(buf, off, t):
    t = t.detach().reshape(-1).to(buf.dtype)
    buf[off:off + t.numel()] = t


def pack_pass_weights(params, dec, agg, device=None):
    """Returns (w_point [total_point], w_ray [total_ray]) for one pass.

    params: mapping name -> tensor; dec: 'dist_decoder' | 'fine_dist_decoder'; agg: 'agg_net' | 'fine_agg_net'.
    A decoder without a vis head (use_vis False) leaves that head block zero.
    """
    L = _lib.weight_layout()
    g = lambda name: params[name]
    device = device if device is not None else g(f"{dec}.mean_decoder.0.weight").device
    wp = torch.zeros(L.total_point, dtype=torch.float32, device=device)
    wr = torch.zeros(L.total_ray, dtype=torch.float32, device=device)

    for hi, head in enumerate(("mean_decoder", "var_decoder", "aw_decoder", "vis_decoder")):
        if f"{dec}.{head}.0.weight" not in params:
            continue
        base = L.dd_head + hi * L.dd_head_stride
        _put(wp, base + L.dd_l0_w, g(f"{dec}.{head}.0.weight").t())
        _put(wp, base + L.dd_l0_b, g(f"{dec}.{head}.0.bias"))
        _put(wp, base + L.dd_l1_w, g(f"{dec}.{head}.2.weight").t())
        _put(wp, base + L.dd_l1_b, g(f"{dec}.{head}.2.bias"))
        _put(wp, base + L.dd_l2_w, g(f"{dec}.{head}.4.weight"))          # [outs,32] row-major
        _put(wp, base + L.dd_l2_b, g(f"{dec}.{head}.4.bias"))

    b = L.grp_b
    _put(wp, b + L.pe0_w, g(f"{agg}.prob_embed.0.weight").t())
    _put(wp, b + L.pe0_b, g(f"{agg}.prob_embed.0.bias"))
    _put(wp, b + L.pe1_w, g(f"{agg}.prob_embed.2.weight").t())
    _put(wp, b + L.pe1_b, g(f"{agg}.prob_embed.2.bias"))
    ib = f"{agg}.agg_impl"
    _put(wp, b + L.rd0_w, g(f"{ib}.ray_dir_fc.0.weight").t())
    _put(wp, b + L.rd0_b, g(f"{ib}.ray_dir_fc.0.bias"))
    rd1 = torch.zeros(16, 36, dtype=torch.float32, device=device)
    rd1[:, :35] = g(f"{ib}.ray_dir_fc.2.weight").detach().t()
    _put(wp, b + L.rd1_w, rd1)
    _put(wp, b + L.rd1_b, g(f"{ib}.ray_dir_fc.2.bias"))
    _put(wp, b + L.nf0_w, g(f"{ib}.neuray_fc.0.weight").t())
    _put(wp, b + L.nf0_b, g(f"{ib}.neuray_fc.0.bias"))
    _put(wp, b + L.nf1_w, g(f"{ib}.neuray_fc.2.weight"))
    _put(wp, b + L.nf1_b, g(f"{ib}.neuray_fc.2.bias"))

    w0 = g(f"{ib}.base_fc.0.weight")                                      # [64, 207] = [glob 140 | rgb_feat 35 | neuray 32]
    _put(wp, L.hoist_w, w0[:, :140].t())
    _put(wp, L.hoist_b, g(f"{ib}.base_fc.0.bias"))
    _put(wp, L.base0_w, w0[:, 140:].t())
    _put(wp, L.base1_w, g(f"{ib}.base_fc.2.weight").t())
    _put(wp, L.base1_b, g(f"{ib}.base_fc.2.bias"))

    d = L.grp_d1
    _put(wp, d + L.vis0_w, g(f"{ib}.vis_fc.0.weight").t())
    _put(wp, d + L.vis0_b, g(f"{ib}.vis_fc.0.bias"))
    v1w, v1b = g(f"{ib}.vis_fc.2.weight"), g(f"{ib}.vis_fc.2.bias")       # [33,32], [33]
    _put(wp, d + L.vis1_w, v1w[:32].t())
    _put(wp, d + L.vis1_b, v1b[:32])
    _put(wp, d + L.vis1l_w, v1w[32])
    _put(wp, d + L.vis1l_b, v1b[32:])
    _put(wp, d + L.v20_w, g(f"{ib}.vis_fc2.0.weight").t())
    _put(wp, d + L.v20_b, g(f"{ib}.vis_fc2.0.bias"))
    _put(wp, d + L.v21_w, g(f"{ib}.vis_fc2.2.weight"))
    _put(wp, d + L.v21_b, g(f"{ib}.vis_fc2.2.bias"))
    _put(wp, d + L.rgb0_w, g(f"{ib}.rgb_fc.0.weight").t())
    _put(wp, d + L.rgb0_b, g(f"{ib}.rgb_fc.0.bias"))
    _put(wp, d + L.rgb1_w, g(f"{ib}.rgb_fc.2.weight").t())
    _put(wp, d + L.rgb1_b, g(f"{ib}.rgb_fc.2.bias"))
    _put(wp, d + L.rgb2_w, g(f"{ib}.rgb_fc.4.weight"))
    _put(wp, d + L.rgb2_b, g(f"{ib}.rgb_fc.4.bias"))

    e = L.grp_d2
    _put(wp, e + L.geo0_w, g(f"{ib}.geometry_fc.0.weight").t())
    _put(wp, e + L.geo0_b, g(f"{ib}.geometry_fc.0.bias"))
    _put(wp, e + L.geo1_w, g(f"{ib}.geometry_fc.2.weight").t())
    _put(wp, e + L.geo1_b, g(f"{ib}.geometry_fc.2.bias"))

    at = f"{ib}.ray_attention"
    _put(wr, L.wq, g(f"{at}.w_qs.weight").t())
    _put(wr, L.wk, g(f"{at}.w_ks.weight").t())
    _put(wr, L.wv, g(f"{at}.w_vs.weight").t())
    _put(wr, L.wfc, g(f"{at}.fc.weight").t())
    _put(wr, L.ln_w, g(f"{at}.layer_norm.weight"))
    _put(wr, L.ln_b, g(f"{at}.layer_norm.bias"))
    _put(wr, L.og0_w, g(f"{ib}.out_geometry_fc.0.weight").t())
    _put(wr, L.og0_b, g(f"{ib}.out_geometry_fc.0.bias"))
    _put(wr, L.og1_w, g(f"{ib}.out_geometry_fc.2.weight"))
    _put(wr, L.og1_b, g(f"{ib}.out_geometry_fc.2.bias"))
    return wp, wr


_SWZ_CACHE = {}


def _sw128_perm(n_rows, device):
    """perm[n*32+k] = float index of element (n,k) inside a K-major SWIZZLE_128B tile of 32-wide fp32 rows
    (csrc/nr_tc.cuh sw128_index): 8-row atoms of 1024 B, 16-byte chunk index XORed with the row index mod 8."""
    key = (n_rows, str(device))
    if key not in _SWZ_CACHE:
        n = torch.arange(n_rows)[:, None]
        k = torch.arange(32)[None, :]
        idx = (n >> 3) * 256 + (n & 7) * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3))
        _SWZ_CACHE[key] = idx.reshape(-1).to(device)
    return _SWZ_CACHE[key]


def _tc_tiles(wmat, k_cols):
    """wmat [N, K] fp32 (already in the kernel's K order, un-padded) -> (hi, lo) flat tensors of ceil(k_cols/32)
    swizzled slabs each; hi = w with the low 13 mantissa bits cleared (exact tf32), lo = w - hi (exact in fp32)."""
    n, k = wmat.shape
    slabs = (k_cols + 31) // 32
    full = torch.zeros(n, slabs * 32, dtype=torch.float32, device=wmat.device)
    full[:, :k] = wmat.detach().float()
    hi = (full.view(torch.int32) & -8192).view(torch.float32)      # 0xffffe000
    lo = full - hi
    perm = _sw128_perm(n, wmat.device)
    out = []
    for part in (hi, lo):
        tiles = torch.empty(slabs, n * 32, dtype=torch.float32, device=wmat.device)
        for s_ in range(slabs):
            tiles[s_, perm] = part[:, 32 * s_:32 * s_ + 32].reshape(-1)
        out.append(tiles.reshape(-1))
    return out


def pack_tc_weights(params, dec, agg, device=None, _comp=None):
    """Tensor-core weight buffer of one pass (layout: csrc/nr_common.cuh namespace tcl / NrTcLayout)."""
    T = _lib.tc_layout()
    g = lambda name: params[name]
    device = device if device is not None else g(f"{dec}.mean_decoder.0.weight").device
    buf = torch.zeros(T.total, dtype=torch.float32, device=device)

    def put(off, wmat, k_cols, lo_off):
        hi, lo = _tc_tiles(wmat.to(device), k_cols)
        buf[off:off + hi.numel()] = hi
        buf[off + lo_off:off + lo_off + lo.numel()] = lo

    for hi_, head in enumerate(("mean_decoder", "var_decoder", "aw_decoder", "vis_decoder")):
        if f"{dec}.{head}.0.weight" not in params:
            continue
        base = T.head0 + hi_ * T.stage
        put(base, g(f"{dec}.{head}.0.weight"), 32, 1024)
        put(base + 2048, g(f"{dec}.{head}.2.weight"), 32, 1024)
    ib = f"{agg}.agg_impl"
    put(T.pe0, g(f"{agg}.prob_embed.0.weight"), 40, 2048)                 # [32, 34] -> K 40
    # prob_embed.2 carries 8 extra output rows: neuray_fc.0 applied to its (linear) output, so the point kernel gets the
    # neuray_fc hidden layer out of the same MMA (rows 32..39 = W_nf0 @ W_pe2; its bias is folded in the kernel's setup)
    wpe1 = g(f"{agg}.prob_embed.2.weight").detach().to(device)
    wnf0 = g(f"{ib}.neuray_fc.0.weight").detach().to(device)
    w48 = torch.zeros(48, 32, dtype=torch.float32, device=device)
    w48[:32] = wpe1.float()
    w48[32:40] = (wnf0.double() @ wpe1.double()).float() if _comp is None else _comp.to(device)
    put(T.pe1, w48, 32, 1536)
    w0 = g(f"{ib}.base_fc.0.weight").detach().float().to(device)        # [64, 207]
    b0 = torch.zeros(64, 72, dtype=torch.float32, device=device)          # K order: rgb_feat 35 | 5 zeros | neuray_feat 32
    b0[:, :35] = w0[:, 140:175]
    b0[:, 35] = g(f"{ib}.base_fc.0.bias").detach().float().to(device)     # bias column: the pm3 kernel feeds a constant 1 at K = 35
    b0[:, 40:72] = w0[:, 175:207]
    hi, lo = _tc_tiles(b0, 72)
    for s_ in range(3):
        buf[T.b0 + s_ * T.stage: T.b0 + s_ * T.stage + 2048] = hi[s_ * 2048:(s_ + 1) * 2048]
        buf[T.b0 + s_ * T.stage + 2048: T.b0 + (s_ + 1) * T.stage] = lo[s_ * 2048:(s_ + 1) * 2048]
    put(T.b1, g(f"{ib}.base_fc.2.weight"), 64, 2048)
    put(T.v01, g(f"{ib}.vis_fc.0.weight"), 32, 1024)
    put(T.v01 + 2048, g(f"{ib}.vis_fc.2.weight")[:32], 32, 1024)
    put(T.v2r, g(f"{ib}.vis_fc2.0.weight"), 32, 1024)
    put(T.v2r + 2048, g(f"{ib}.rgb_fc.0.weight"), 40, 1024)               # [16, 37] -> K 40, two 512-float slabs
    wrd = torch.zeros(48, 16, dtype=torch.float32, device=device)         # ray_dir_fc.2 [35, 16] -> 48 rows (MMA N % 16 == 0)
    wrd[:35] = g(f"{ib}.ray_dir_fc.2.weight").detach().float().to(device)
    put(T.rd1, wrd, 16, 1536)
    # view-pooled inputs of base_fc.0 (ibrnet.py:338-342): K = 32*round + 8*stat + i <-> reference column stat*35 + 8*round + i
    bh = torch.zeros(64, 160, dtype=torch.float32, device=device)
    for r_ in range(5):
        for s_ in range(4):
            n_ = min(8, 35 - 8 * r_)
            bh[:, 32 * r_ + 8 * s_: 32 * r_ + 8 * s_ + n_] = w0[:, s_ * 35 + 8 * r_: s_ * 35 + 8 * r_ + n_]
    hi, lo = _tc_tiles(bh, 160)
    buf[T.hst: T.hst + 10240] = hi
    buf[T.hst + 10240: T.hst + 20480] = lo
    # geometry_fc.0 on the second view pooling (ibrnet.py:352-354): K = 32*round + 16*stat + i <-> reference column stat*32 + 16*round + i;
    # third stage: K 64 = mean weight, K 65 = bias (constant-1 input)
    wg = g(f"{ib}.geometry_fc.0.weight").detach().float().to(device)     # [64, 65]: mean 32 | var 32 | mean weight
    bg = torch.zeros(64, 96, dtype=torch.float32, device=device)
    for r_ in range(2):
        for s_ in range(2):
            bg[:, 32 * r_ + 16 * s_: 32 * r_ + 16 * s_ + 16] = wg[:, s_ * 32 + 16 * r_: s_ * 32 + 16 * r_ + 16]
    bg[:, 64] = wg[:, 64]
    bg[:, 65] = g(f"{ib}.geometry_fc.0.bias").detach().float().to(device)
    hi, lo = _tc_tiles(bg, 96)
    for s_ in range(3):
        buf[T.g0 + s_ * T.stage: T.g0 + s_ * T.stage + 2048] = hi[s_ * 2048:(s_ + 1) * 2048]
        buf[T.g0 + s_ * T.stage + 2048: T.g0 + (s_ + 1) * T.stage] = lo[s_ * 2048:(s_ + 1) * 2048]
    return buf


def camera_block(pose, K, depth_range):
    """que_cam [24] = R^T (9) | centre (3) | K^-1 (9) | near, far, 0, built with the reference's own torch
    expressions (render_ops.py:14-20) so the rounding matches."""
    rot = pose[:, :3].t()
    centre = -(rot @ pose[:, 3:])
    kinv = torch.inverse(K)
    pad = torch.zeros(1, dtype=torch.float32, device=pose.device)
    return torch.cat([rot.reshape(-1), centre.reshape(-1), kinv.reshape(-1), depth_range.reshape(-1)[:2], pad]).contiguous()


def view_param_block(poses, Ks, depth_range=None):
    """view_params [rfn,20] = K@Rt (12) | centre (3) | -1/near, -1/far | pad (3)  (render_ops.py:95, 112; dist_decoder.py:17-20)."""
    rfn = poses.shape[0]
    KRt = Ks @ poses
    centre = -(poses[:, :, :3].transpose(1, 2) @ poses[:, :, 3:])
    if depth_range is None:
        inv = torch.zeros(rfn, 2, dtype=torch.float32, device=poses.device)
    else:
        inv = -1 / depth_range
    pad = torch.zeros(rfn, 3, dtype=torch.float32, device=poses.device)
    return torch.cat([KRt.reshape(rfn, 12), centre.reshape(rfn, 3), inv, pad], 1).contiguous()


def point_index_map_cpu(params, dec, agg):
    """CPU twin of neuray_b200.weights.point_index_map (which needs the GPU packer): position in w_point -> 1 + flat
    parameter index, from packing parameters whose elements are their own index with the PyTorch packer above."""
    names = list(params)
    coded, off = {}, 1
    for k in names:
        n = params[k].numel()
        coded[k] = torch.arange(off, off + n, dtype=torch.float32).reshape(params[k].shape)
        off += n
    wp, _ = pack_pass_weights(coded, dec, agg, torch.device("cpu"))
    return wp.long(), names, [params[k].numel() for k in names], [tuple(params[k].shape) for k in names]
