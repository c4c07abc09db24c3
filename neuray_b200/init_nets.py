"""Host side of DepthInitNet, the init net of the neuray_gen_depth model (reference network/init_net.py:63-101), forward /
inference:

    depth      = extract_depth_for_init(ref_imgs_info)                       nr_extract_depth
    diff_feats = get_diff_feats(ref_imgs_info, depth)                        nr_diff_feats (neuray_b200.init_ops)
    feats      = res_net(cat([imgs, depth, diff_feats]))                     ResEncoder (ops.py:232-312) on the tensor cores
    ray_feats  = conv_out(cat([depth_skip(depth), feats]))                   nr_depth_init_fwd

`DepthInitNet` is a parameter container under the reference's state-dict names whose forward is the native path (it returns
the reference's NCHW [rfn,32,h/4,w/4] tensor); `forward_into` writes channel-last straight into the frame pack, where the
vis encoder expects the init net's ray_feats.  `patch.install()` puts `forward_or_reference` over the reference class'
forward: with a gradient wanted through the net (training) the reference's own torch code runs.  CUDA tensors only.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib, encoders, init_ops

BLOCKS = (2, 2, 2)
LAUNCHES = 3 + 2 + 3 * (5 + 4) + 2 + 8 + 2 + 1 + 1 + 1       # input assembly, first layer, 3 stages, decoder, out_conv, depth_skip, conv_out


def param_names():
    """state_dict() order of DepthInitNet = the order nr_depth_init_pack expects (72 tensors)."""
    return (["res_net." + n for n in encoders.unet_param_names(BLOCKS)] +
            ["depth_skip.0.weight", "depth_skip.0.bias", "depth_skip.2.weight", "depth_skip.2.bias", "conv_out.weight", "conv_out.bias"])


def _shapes():
    shapes = {"res_net." + k: v for k, v in encoders.unet_param_shapes((32, 12, 8, 8), BLOCKS).items()}
    shapes.update({"depth_skip.0.weight": (8, 1, 2, 2), "depth_skip.0.bias": (8,), "depth_skip.2.weight": (16, 8, 2, 2), "depth_skip.2.bias": (16,),
                   "conv_out.weight": (32, 48, 1, 1), "conv_out.bias": (32,)})
    return {n: shapes[n] for n in param_names()}


def dims(h, w):
    fh, fw = C.c_int(), C.c_int()
    _lib.check(_lib.lib().nr_depth_init_dims(h, w, C.byref(fh), C.byref(fw)), "nr_depth_init_dims")
    return fh.value, fw.value


def usable(module, ref_imgs_info):
    if not ref_imgs_info["imgs"].is_cuda:
        return False
    have = dict(module.named_parameters())
    if not all(n in have for n in param_names()):
        return False
    return not (torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()))


def extract_depth_for_init(ref_imgs_info):
    """init_net.py:76-79 -> [rfn,1,h,w] normalised inverse depth."""
    depth, rng = ref_imgs_info["depth"], ref_imgs_info["depth_range"]
    rfn, _, h, w = depth.shape
    d, r = depth.detach().contiguous().float(), rng.detach().contiguous().float()
    out = torch.empty_like(d)
    with _lib.on_device(depth):
        _lib.check(_lib.lib().nr_extract_depth(_lib.ptr(d), _lib.ptr(r), rfn, h, w, _lib.ptr(out), _lib.stream_of(depth)), "nr_extract_depth")
    _lib.count_launches(1)
    return out


def forward_into(module, ref_imgs_info, out, out_off):
    """DepthInitNet.forward (init_net.py:93-101) into channels [out_off, out_off + 32) of the channel-last buffer out."""
    imgs = ref_imgs_info["imgs"]
    if not imgs.is_cuda:
        raise _lib.NeurayB200Error("DepthInitNet needs CUDA tensors (no CPU fallback)")
    rfn, _, h, w = imgs.shape
    dev = imgs.device
    depth = extract_depth_for_init(ref_imgs_info)
    diff = init_ops.get_diff_feats(ref_imgs_info, depth)
    packed = encoders._packed(module, param_names(), "depth_init", dev)
    nbytes = _lib.lib().nr_depth_init_workspace(rfn, h, w)
    ws = encoders._workspace("depth_init", nbytes, dev)
    x = imgs.detach().contiguous().float()
    with _lib.on_device(imgs):
        _lib.check(_lib.lib().nr_depth_init_fwd(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(depth), _lib.ptr(diff.contiguous()), rfn, h, w, _lib.ptr(out),
                                                out.shape[-1], out_off, int(encoders.PRECISION == "tf32"), ws.data_ptr(), nbytes,
                                                _lib.stream_of(imgs)), "nr_depth_init_fwd")
    _lib.count_launches(LAUNCHES)
    return out


def forward_nchw(module, ref_imgs_info):
    imgs = ref_imgs_info["imgs"]
    rfn, _, h, w = imgs.shape
    fh, fw = dims(h, w)
    buf = torch.empty(rfn, fh, fw, 32, dtype=torch.float32, device=imgs.device)
    forward_into(module, ref_imgs_info, buf, 0)
    return encoders.from_channel_last(buf, 0, 32)


class DepthInitNet(nn.Module):
    """reference init_net.py:76-101 with the reference's parameter names; forward(ref_imgs_info, src_imgs_info, is_train)."""
    default_cfg = {}

    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = {**self.default_cfg, **(cfg or {})}
        encoders._plant(self, _shapes())

    def forward(self, ref_imgs_info, src_imgs_info=None, is_train=False):
        if not usable(self, ref_imgs_info):
            raise _lib.NeurayB200Error("DepthInitNet: the native init net is forward-only and CUDA-only; run it under torch.no_grad() "
                                       "(training keeps the torch module upstream of the boundary)")
        return forward_nchw(self, ref_imgs_info)


def forward_or_reference(reference_forward):
    """The forward patch.install() puts over the reference's DepthInitNet: native when no gradient is wanted through the net."""
    def forward(self, ref_imgs_info, src_imgs_info, is_train):
        if usable(self, ref_imgs_info):
            return forward_nchw(self, ref_imgs_info)
        return reference_forward(self, ref_imgs_info, src_imgs_info, is_train)
    return forward
