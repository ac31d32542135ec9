"""TEST INFRASTRUCTURE ONLY -- fp32 CPU/torch restatement of the reference network (the parity oracle for
the floating-point side of the hot path).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs may import this module; the product package never does.

Restates, from the reference's behaviour (no code copied):
  DeepRecurrNet.forward / reset_states           models/model.py:294-344
  FeatsExtract                                   models/model.py:20-45
  TimePropagation (local + global correlation)   models/model.py:48-153
  STFusion (DCN alignment, attention, decoder)   models/model.py:156-291
  ConvLayer / UpsampleConvLayer / ResidualBlock  models/submodules.py:159-200, 254-299, 347-409
  RecurrentConvLayer / ConvGRU / MLP             models/submodules.py:302-344, 474-514, 67-77
  CropSize                                       models/model_util.py:41-48, 133-164
  DCN_sep.forward + the modulated deformable conv kernel semantics
                                                 models/DCNv2/dcn_v2.py:214-227,
                                                 models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:25-54,125-195,
                                                 models/DCNv2/src/cuda/dcn_v2_cuda.cu:67-92

It is a pure function of (state_dict with the reference's key names, input, recurrent states).
Parity status: PINNED by tests/test_oracle_model.py against tests/golden/model_golden.npz, which
tests/golden/make_golden_model.py produced by running the reference's own models/model.py (imported from
/root/reference with the two import stubs of SURVEY.md 8c; `_ext.dcn_v2_forward` served by
torchvision.ops.deform_conv2d, which SURVEY 8c verified bit-identical to the patched reference CPU kernel).
"""
from math import ceil, floor

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# deterministic parameters shared by the golden generator, the oracle tests and the GPU tests
# ------------------------------------------------------------------------------------------------
def param_shapes(inch=2, basech=8, num_frame=3):
    """Ordered {reference state_dict key: shape} of DeepRecurrNet(inch, basech, num_frame) (68 tensors)."""
    b, C = basech, 8 * basech
    s = {}

    def conv(name, co, ci, k=3):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    conv("head.conv2d", b, inch)
    conv("feat_extract.convblock.0.conv2d", 2 * b, b)
    conv("feat_extract.convblock.1.conv2d", 4 * b, 2 * b)
    conv("feat_extract.convblock.2.conv2d", 8 * b, 4 * b)
    conv("time_propagate.pred_map.0.conv2d", C, 2 * C)
    conv("time_propagate.pred_map.1.conv2d", 1, C)
    conv("time_propagate.local_fusion.0.conv1", 3 * C, 3 * C)
    conv("time_propagate.local_fusion.0.conv2", 3 * C, 3 * C)
    conv("time_propagate.local_fusion.1.conv2d", C, 3 * C)
    conv("time_propagate.lstm.conv.conv2d", C, C)
    conv("time_propagate.lstm.recurrent_block.reset_gate", C, 2 * C)
    conv("time_propagate.lstm.recurrent_block.update_gate", C, 2 * C)
    conv("time_propagate.lstm.recurrent_block.out_gate", C, 2 * C)
    conv("time_propagate.global_fusion.conv2d", C, 2 * C, 1)
    conv("spacetime_fuse.offset.0.conv2d", C, 2 * C)
    conv("spacetime_fuse.offset.1.conv2d", C, C)
    s["spacetime_fuse.dcn.weight"] = (C, C, 3, 3)
    s["spacetime_fuse.dcn.bias"] = (C,)
    conv("spacetime_fuse.dcn.conv_offset_mask", 8 * 3 * 9, C)
    conv("spacetime_fuse.convblock.0.conv2d", C, 2 * C)
    conv("spacetime_fuse.convblock.1.conv2d", C, C)
    conv("spacetime_fuse.kernel.conv2d", 2, C, 1)
    s["spacetime_fuse.fc.0.layers.0.weight"] = (C // 2, C)
    s["spacetime_fuse.fc.0.layers.0.bias"] = (C // 2,)
    s["spacetime_fuse.fc.0.layers.1.weight"] = (2 * C, C // 2)
    s["spacetime_fuse.fc.0.layers.1.bias"] = (2 * C,)
    conv("spacetime_fuse.dcn_fusion.0.conv2d", C, 2 * C)
    conv("spacetime_fuse.dcn_fusion.1.conv2d", C, C)
    conv("spacetime_fuse.dense_fusion.0.conv2d", C, num_frame * C)
    conv("spacetime_fuse.dense_fusion.1.conv2d", C, C)
    conv("spacetime_fuse.attens.0.conv2d", 1, C)
    conv("spacetime_fuse.attens.1.conv2d", 1, C // 2)
    conv("spacetime_fuse.attens.2.conv2d", 1, C // 4)
    conv("spacetime_fuse.recons.0.conv2d", C // 2, C)
    conv("spacetime_fuse.recons.1.conv2d", C // 4, C // 2)
    conv("spacetime_fuse.recons.2.conv2d", C // 8, C // 4)
    conv("tail.conv2d", inch, b)
    return s


def seeded_state_dict(seed=0, inch=2, basech=8, num_frame=3, offset_std=0.02):
    """A reproducible state_dict (fan-in scaled normal weights, small biases) that any implementation with the
    reference's key names can load.  conv_offset_mask gets NON-zero weights (the reference's zero init never
    exercises bilinear sampling)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(inch, basech, num_frame).items():
        if k.endswith(".weight"):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            std = offset_std if "conv_offset_mask" in k else (1.0 / fan_in) ** 0.5
            sd[k] = torch.randn(shp, generator=g) * std
        else:
            sd[k] = torch.randn(shp, generator=g) * (0.3 if "conv_offset_mask" in k else 0.05)
    return sd


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def _act(x, name):
    if name is None:
        return x
    return getattr(torch, name)(x)


def _conv(sd, name, x, stride=1, act=None, pad=None):
    w = sd[name + ".weight"]
    if pad is None:
        pad = w.shape[-1] // 2
    return _act(F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=pad), act)


def dcn_v2_forward(inp, weight, bias, offset, mask, dg):
    """Modulated deformable 3x3 conv, stride 1, pad 1, dilation 1 (dcn_v2_im2col_cuda.cu:125-195).

    offset [B, dg*18, H, W]: for group g, tap k = i*3+j: channel g*18+2k is the row (h) offset, +1 the col offset.
    mask   [B, dg*9, H, W].  Sample position (y-1+i+off_h, x-1+j+off_w); contributes only if > -1 and < H/W;
    bilinear corners outside the image read as zero (:25-54).  columns[c*9+k] = val*mask; out = W[Co,Ci*9]·columns + b."""
    B, Ci, H, W = inp.shape
    Co = weight.shape[0]
    cpg = Ci // dg
    ys = torch.arange(H, dtype=inp.dtype).view(1, 1, H, 1)
    xs = torch.arange(W, dtype=inp.dtype).view(1, 1, 1, W)
    cols = inp.new_zeros(B, Ci, 9, H, W)
    flat = inp.reshape(B, Ci, H * W)
    for k in range(9):
        i, j = k // 3, k % 3
        off_h = offset[:, [g * 18 + 2 * k for g in range(dg)]]          # [B, dg, H, W]
        off_w = offset[:, [g * 18 + 2 * k + 1 for g in range(dg)]]
        m = mask[:, [g * 9 + k for g in range(dg)]]
        h_im = ys - 1 + i + off_h
        w_im = xs - 1 + j + off_w
        valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low, w_low = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        h_high, w_high = h_low + 1, w_low + 1

        def corner(hi, wi, ok):
            ok = ok & valid
            idx = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1))             # [B, dg, H, W]
            idx = idx.repeat_interleave(cpg, dim=1).reshape(B, Ci, H * W)
            v = torch.gather(flat, 2, idx).reshape(B, Ci, H, W)
            return v * ok.repeat_interleave(cpg, dim=1)

        v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))
        v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
        v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
        v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
        r = lambda t: t.repeat_interleave(cpg, dim=1)
        val = r(hh * hw) * v1 + r(hh * lw) * v2 + r(lh * hw) * v3 + r(lh * lw) * v4
        cols[:, :, k] = val * r(m)
    out = torch.einsum("ok,bkp->bop", weight.reshape(Co, Ci * 9), cols.reshape(B, Ci * 9, H * W))
    return out.reshape(B, Co, H, W) + bias.view(1, Co, 1, 1)


def _dcn_sep(sd, inp, fea, dcn_fn):
    """DCN_sep.forward (dcn_v2.py:214-227), deformable_groups = 8."""
    out = _conv(sd, "spacetime_fuse.dcn.conv_offset_mask", fea)
    o1, o2, mask = torch.chunk(out, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    mask = torch.sigmoid(mask)
    return dcn_fn(inp, sd["spacetime_fuse.dcn.weight"], sd["spacetime_fuse.dcn.bias"], offset, mask, 8)


def _local_time_corre(sd, f0, f1, f2):
    p = "time_propagate."

    def pred_map(x):
        return _conv(sd, p + "pred_map.1.conv2d", _conv(sd, p + "pred_map.0.conv2d", x, act="relu"), act="sigmoid")

    m0 = pred_map(torch.cat([f0, f1], 1))
    m1 = pred_map(torch.cat([f1, f2], 1))
    x = torch.cat([f0 * m0, f1, f2 * m1], 1)
    # ResidualBlock(192,192): conv1 relu conv2 (+x) relu   (submodules.py:391-409)
    r = F.relu(F.conv2d(x, sd[p + "local_fusion.0.conv1.weight"], sd[p + "local_fusion.0.conv1.bias"], padding=1))
    r = F.conv2d(r, sd[p + "local_fusion.0.conv2.weight"], sd[p + "local_fusion.0.conv2.bias"], padding=1)
    r = F.relu(r + x)
    return _conv(sd, p + "local_fusion.1.conv2d", r) + f1


def _gru_step(sd, x, h):
    """RecurrentConvLayer -> ConvGRU (submodules.py:340-344, 496-514)."""
    p = "time_propagate.lstm."
    x = _conv(sd, p + "conv.conv2d", x, act="relu")
    if h is None:
        h = torch.zeros_like(x)
    xh = torch.cat([x, h], 1)
    z = torch.sigmoid(F.conv2d(xh, sd[p + "recurrent_block.update_gate.weight"], sd[p + "recurrent_block.update_gate.bias"], padding=1))
    r = torch.sigmoid(F.conv2d(xh, sd[p + "recurrent_block.reset_gate.weight"], sd[p + "recurrent_block.reset_gate.bias"], padding=1))
    o = torch.tanh(F.conv2d(torch.cat([x, h * r], 1), sd[p + "recurrent_block.out_gate.weight"],
                            sd[p + "recurrent_block.out_gate.bias"], padding=1))
    return h * (1 - z) + o * z


def forward(sd, inp, states=None, dcn_fn=dcn_v2_forward, gtc_frozen=False):
    """DeepRecurrNet.forward.  inp [B,N,2,H,W] fp32 -> (out [B,2,H,W], new_states [h_fwd, h_rev])."""
    B, N, Cin, H, W = inp.shape
    x = inp
    Hc, Wc = 8 * ceil(H / 8), 8 * ceil(W / 8)
    need_crop = (H % 8 != 0) or (W % 8 != 0)
    if need_crop:                                                        # CropSize.pad (model_util.py:148-152)
        pt, pb = ceil(0.5 * (Hc - H)), floor(0.5 * (Hc - H))
        pl, pr = ceil(0.5 * (Wc - W)), floor(0.5 * (Wc - W))
        x = F.pad(x, (pl, pr, pt, pb))
    x = x.reshape(B * N, Cin, Hc, Wc)
    x = _conv(sd, "head.conv2d", x, act="relu")
    feats = []
    for i in range(3):
        x = _conv(sd, f"feat_extract.convblock.{i}.conv2d", x, stride=2, act="relu")
        feats.append(x)
    feats = feats[::-1]                                                  # [64@h, 32@2h, 16@4h]
    C, h, w = feats[0].shape[1:]
    f = feats[0].view(B, N, C, h, w)

    # ---- TimePropagation (model.py:126-153)
    ltc = []
    for i in range(N):
        idx = [0, 0, 1] if i == 0 else ([N - 2, N - 1, N - 1] if i == N - 1 else [i - 1, i, i + 1])
        ltc.append(_local_time_corre(sd, f[:, idx[0]], f[:, idx[1]], f[:, idx[2]]))
    ltc = torch.stack(ltc, 1)
    st_f, st_r = (states if states is not None else (None, None))
    xs_f, xs_r = [], []
    for i in range(N):
        if gtc_frozen:
            st_f, st_r = None, None
        st_f = _gru_step(sd, ltc[:, i], st_f)
        st_r = _gru_step(sd, ltc[:, N - 1 - i], st_r)
        xs_f.append(st_f)
        xs_r.append(st_r)
    new_states = [None, None] if gtc_frozen else [st_f, st_r]
    xf = torch.stack(xs_f, 1)
    xr = torch.stack(xs_r[::-1], 1)
    g = torch.cat([xf, xr], 2).view(B * N, 2 * C, h, w)
    g = _conv(sd, "time_propagate.global_fusion.conv2d", g, act="relu", pad=0).view(B, N, C, h, w)
    tp = g + f

    # ---- STFusion.dense_fuse (model.py:208-251)
    mid = (N - 1) // 2
    p = "spacetime_fuse."
    fused = []
    for i in list(range(mid)) + list(range(mid + 1, N)):
        f0, f1 = tp[:, i], tp[:, mid]
        off = _conv(sd, p + "offset.1.conv2d", _conv(sd, p + "offset.0.conv2d", torch.cat([f0, f1], 1), act="relu"))
        al = F.relu(_dcn_sep(sd, f0, off, dcn_fn))
        ft = _conv(sd, p + "convblock.1.conv2d", _conv(sd, p + "convblock.0.conv2d", torch.cat([al, f1], 1), act="relu"))
        sk = _conv(sd, p + "kernel.conv2d", ft, act="sigmoid", pad=0)           # [B,2,h,w]
        mx = ft.view(B, C, h * w).max(dim=2)[0]                                  # global max pool
        ck = F.relu(F.linear(mx, sd[p + "fc.0.layers.0.weight"], sd[p + "fc.0.layers.0.bias"]))
        ck = torch.sigmoid(F.linear(ck, sd[p + "fc.0.layers.1.weight"], sd[p + "fc.0.layers.1.bias"]))  # [B,2C]
        y0 = al * sk[:, 0:1] * ck[:, :C, None, None]
        y1 = f1 * sk[:, 1:2] * ck[:, C:, None, None]
        fused.append(_conv(sd, p + "dcn_fusion.1.conv2d",
                           _conv(sd, p + "dcn_fusion.0.conv2d", torch.cat([y0, y1], 1), act="relu")))
    fused.append(tp[:, mid])
    x = _conv(sd, p + "dense_fusion.1.conv2d", _conv(sd, p + "dense_fusion.0.conv2d", torch.cat(fused, 1), act="relu"))

    # ---- scale_aggre + recons (model.py:253-291)
    for idx, ft in enumerate(feats):
        at = _conv(sd, p + f"attens.{idx}.conv2d", ft, act="sigmoid")            # [BN,1,.,.]
        agg = (ft * at).view(B, N, ft.shape[1], ft.shape[2], ft.shape[3]).mean(1)
        x = x + agg
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = _conv(sd, p + f"recons.{idx}.conv2d", x, act="relu")
    x = _conv(sd, "tail.conv2d", x, act="relu")
    if need_crop:                                                        # CropSize.crop (model_util.py:154-164)
        cx, cy = floor(Wc / 2), floor(Hc / 2)
        x = x[..., cy - floor(H / 2): cy + ceil(H / 2), cx - floor(W / 2): cx + ceil(W / 2)].contiguous()
    return x, new_states


class OracleNet:
    """Stateful convenience wrapper with the reference's call pattern (forward / reset_states)."""

    def __init__(self, state_dict, dcn_fn=dcn_v2_forward):
        self.sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.dcn_fn = dcn_fn
        self.states = None

    def reset_states(self):
        self.states = None

    @torch.no_grad()
    def __call__(self, inp):
        out, self.states = forward(self.sd, inp, self.states, self.dcn_fn)
        return out
