"""CPU: the C-ABI library loads and exports exactly what include/neuray_b200.h declares; host-side packing logic."""
import os
import re

import pytest
import torch

import ref_packers as weights
from neuray_b200 import _lib, synthetic
from neuray_b200 import weights as nr_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "neuray_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(nr_[a-z0-9_]+)\s*\(", src))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    handle = _lib.lib()
    for n in names:
        assert getattr(handle, n) is not None
    assert handle.nr_abi_version() == _lib.ABI_VERSION


def test_weight_layout_is_consistent():
    L = _lib.weight_layout()
    assert L.total_point > 0 and L.total_ray == 1348
    assert L.dd_head_stride == 2180 and L.grp_b == 4 * 2180
    for name in ("grp_b", "hoist_w", "base0_w", "base1_w", "grp_d1", "grp_d2"):
        assert getattr(L, name) % 4 == 0, name          # 16-byte alignment for 128-bit staging copies
    assert L.grp_d2 + L.grp_d2_size == L.total_point


def test_pack_covers_every_parameter_exactly_once():
    """Give every parameter element a unique value; after packing, each value must appear exactly once."""
    cfg = {"use_hierarchical_sampling": True}
    W = synthetic.make_weights(cfg)
    for pre_d, pre_a in (("dist_decoder", "agg_net"), ("fine_dist_decoder", "fine_agg_net")):
        params = {k: v for k, v in W.items() if k.startswith(pre_d + ".") or k.startswith(pre_a + ".")}
        counter = 1.0
        uniq = {}
        for k in sorted(params):
            n = params[k].numel()
            uniq[k] = torch.arange(counter, counter + n, dtype=torch.float32).reshape(params[k].shape)
            counter += n
        wp, wr = weights.pack_pass_weights(uniq, pre_d, pre_a, torch.device("cpu"))
        packed = torch.cat([wp, wr])
        vals = packed[packed != 0]
        assert vals.numel() == int(counter - 1), (vals.numel(), counter - 1)
        assert torch.equal(torch.sort(vals)[0], torch.arange(1.0, counter))


def test_pack_spot_checks():
    cfg = {"dist_decoder_cfg": {"use_vis": False}}
    W = synthetic.make_weights(cfg)
    L = _lib.weight_layout()
    wp, wr = weights.pack_pass_weights(W, "dist_decoder", "agg_net", torch.device("cpu"))
    w = W["agg_net.agg_impl.base_fc.0.weight"]
    assert wp[L.hoist_w + 5 * 64 + 7] == w[7, 5]                  # WT[k][j] = W[j][k]
    assert wp[L.base0_w + 3 * 64 + 9] == w[9, 140 + 3]
    assert wp[L.grp_d1 + L.vis1l_w + 4] == W["agg_net.agg_impl.vis_fc.2.weight"][32, 4]
    assert wr[L.wq + 2 * 16 + 5] == W["agg_net.agg_impl.ray_attention.w_qs.weight"][5, 2]
    vis_block = wp[L.dd_head + 3 * L.dd_head_stride: L.dd_head + 4 * L.dd_head_stride]
    assert float(vis_block.abs().sum()) == 0.0                     # no vis head when use_vis is False


def test_ops_refuse_cpu_tensors():
    from neuray_b200 import render_ops
    with pytest.raises(_lib.NeurayB200Error):
        render_ops.depth2dists(torch.rand(1, 4, 8))


def test_posenc_matches_oracle():
    import neuray_oracle as orc
    assert torch.equal(nr_weights.posenc_table(48), orc.posenc_table(48)[0])


def test_tc_weight_pack_roundtrip():
    """Un-swizzle the packed tensor-core tiles and check hi + lo == W exactly, hi is tf32-exact, K padding is zero."""
    cfg = {"use_hierarchical_sampling": False}
    W = synthetic.make_weights(cfg)
    T = _lib.tc_layout()
    buf = weights.pack_tc_weights(W, "dist_decoder", "agg_net", torch.device("cpu"))
    assert buf.numel() == T.total == 84992

    def unswz(flat, n):                # [slabs*n*32] -> [n, slabs*32]
        slabs = flat.numel() // (n * 32)
        perm = weights._sw128_perm(n, torch.device("cpu"))
        return torch.cat([flat[s * n * 32:(s + 1) * n * 32][perm].reshape(n, 32) for s in range(slabs)], 1)

    w = W["agg_net.agg_impl.base_fc.2.weight"]                      # [32, 64], stage B1: hi 0..2048, lo 2048..4096
    hi, lo = unswz(buf[T.b1:T.b1 + 2048], 32), unswz(buf[T.b1 + 2048:T.b1 + 4096], 32)
    assert torch.equal(hi + lo, w)
    assert torch.equal((hi.view(torch.int32) & 8191), torch.zeros_like(hi, dtype=torch.int32))
    w0 = W["agg_net.agg_impl.base_fc.0.weight"]
    rec = torch.cat([unswz(buf[T.b0 + s * T.stage:T.b0 + s * T.stage + 2048], 64) +
                     unswz(buf[T.b0 + s * T.stage + 2048:T.b0 + (s + 1) * T.stage], 64) for s in range(3)], 1)   # [64, 96]
    assert torch.equal(rec[:, :35], w0[:, 140:175]) and torch.equal(rec[:, 40:72], w0[:, 175:207])
    assert torch.equal(rec[:, 35], W["agg_net.agg_impl.base_fc.0.bias"])             # bias column (constant-1 input)
    assert float(rec[:, 36:40].abs().sum()) == 0.0 and float(rec[:, 72:].abs().sum()) == 0.0
    hst = unswz(buf[T.hst:T.hst + 10240], 64) + unswz(buf[T.hst + 10240:T.hst + 20480], 64)          # [64, 160]
    for r_, s_, i_ in ((0, 0, 0), (2, 1, 3), (4, 3, 2), (3, 2, 7)):
        assert torch.equal(hst[:, 32 * r_ + 8 * s_ + i_], w0[:, s_ * 35 + 8 * r_ + i_])
    for s_ in range(4):                                                                                  # features 35..39 do not exist
        assert float(hst[:, 32 * 4 + 8 * s_ + 3: 32 * 4 + 8 * s_ + 8].abs().sum()) == 0.0
    wg = W["agg_net.agg_impl.geometry_fc.0.weight"]
    geo = torch.cat([unswz(buf[T.g0 + s * T.stage:T.g0 + s * T.stage + 2048], 64) +
                     unswz(buf[T.g0 + s * T.stage + 2048:T.g0 + (s + 1) * T.stage], 64) for s in range(3)], 1)      # [64, 96]
    for r_, s_, i_ in ((0, 0, 0), (1, 1, 15), (1, 0, 7), (0, 1, 7)):
        assert torch.equal(geo[:, 32 * r_ + 16 * s_ + i_], wg[:, s_ * 32 + 16 * r_ + i_])
    assert torch.equal(geo[:, 64], wg[:, 64]) and torch.equal(geo[:, 65], W["agg_net.agg_impl.geometry_fc.0.bias"])
    assert float(geo[:, 66:].abs().sum()) == 0.0
    # prob_embed.2 + neuray_fc.0 behind it: 48-row tile, hi 0..1536, lo 1536..3072
    rec = unswz(buf[T.pe1:T.pe1 + 1536], 48) + unswz(buf[T.pe1 + 1536:T.pe1 + 3072], 48)
    wpe, wnf = W["agg_net.prob_embed.2.weight"], W["agg_net.agg_impl.neuray_fc.0.weight"]
    assert torch.equal(rec[:32], wpe) and float(rec[40:].abs().sum()) == 0.0
    assert torch.allclose(rec[32:40], wnf @ wpe, rtol=1e-5, atol=1e-6)
    rec = unswz(buf[T.rd1:T.rd1 + 1536], 48) + unswz(buf[T.rd1 + 1536:T.rd1 + 3072], 48)   # ray_dir_fc.2, resident tile
    wrd = W["agg_net.agg_impl.ray_dir_fc.2.weight"]
    assert torch.equal(rec[:35, :16], wrd) and float(rec[35:].abs().sum()) == 0.0 and float(rec[:, 16:].abs().sum()) == 0.0
    wr = W["agg_net.agg_impl.rgb_fc.0.weight"]                      # [16, 37] at V2R+2048 (hi, two 512 slabs), lo at +1024
    rec = unswz(buf[T.v2r + 2048:T.v2r + 3072], 16) + unswz(buf[T.v2r + 3072:T.v2r + 4096], 16)
    assert torch.equal(rec[:, :37], wr) and float(rec[:, 37:].abs().sum()) == 0.0


def test_pass_weight_struct_fields_cover_the_state_dict(monkeypatch):
    """Host glue of nr_pack_weights: every parameter of a pass lands in exactly one field of NrPassWeights (pointer
    identity checked on the CPU by stubbing the device-pointer accessor), a missing vis head leaves its block NULL."""
    monkeypatch.setattr(_lib, "ptr", lambda t: None if t is None else t.data_ptr())
    for use_vis in (True, False):
        cfg = {"dist_decoder_cfg": {"use_vis": use_vis}}
        W = synthetic.make_weights(cfg)
        w, keep = nr_weights.pass_weight_struct(W, "dist_decoder", "agg_net")
        want = {v.data_ptr() for v in W.values()}
        got = []

        def walk(x):
            if isinstance(x, _lib.NrLinear):
                got.extend([x.w, x.b])
            elif hasattr(x, "__len__"):
                for y in x:
                    walk(y)
        for name, _ in _lib.NrPassWeights._fields_:
            v = getattr(w, name)
            walk(v) if not isinstance(v, (int, type(None))) else got.append(v)
        nonnull = [g for g in got if g]
        assert len(nonnull) == len(set(nonnull)) == len(W), (len(nonnull), len(W))
        assert set(nonnull) == want
        vis = w.dist_decoder[3]
        assert all((l.w is None) == (not use_vis) for l in vis)
