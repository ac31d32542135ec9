"""Training step of the hot path (train_ours_cnt_seq.py:206-235, 767-782).

The reference sums MSELoss(pred, gt) over the L-2 windows of a sequence (ConvGRU state carried, so gradients flow back
through time across windows), calls backward once, lets DDP all-reduce the 1 813 120 gradients and steps
Adam(lr, weight_decay, amsgrad).  Here

  * every ConvLayer (models/submodules.py:159-200) is `conv2d` below: an autograd.Function whose forward AND backward
    are the sm_100a operators esr_conv2d_forward / esr_conv2d_backward (tcgen05 implicit GEMM for the 64-multiple
    layers incl. dx and dw, CUDA-core kernels for the narrow full-resolution layers; include/esr_b200.h);
  * DCN_sep (models/DCNv2/dcn_v2.py:17-68) is `dcn_v2`: esr_dcn_v2_forward / esr_dcn_v2_backward;
  * the loss and the optimizer are esr_mse_loss / esr_adam_step (one launch over the flat parameter buffer);
  * autograd itself, the concatenations / gating products / bilinear x2 / global max / 64-32-128 MLP between those
    operators are torch on the GPU -- plumbing, < 2 % of the step's FLOPs.

`forward_window` is the differentiable counterpart of esr_b200.DeepRecurrNet.forward (which runs the fused inference
plan and cannot be differentiated); DeepRecurrNet.forward dispatches here when gradients are enabled.
"""
import math

import torch
import torch.nn.functional as F

from . import _lib, dcn_v2_ext

_ACT = {None: 0, "none": 0, "relu": 1, "sigmoid": 2, "tanh": 3}


def _ws(nbytes, device):
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)


class _Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, act):
        if not x.is_cuda:
            raise _lib.ESRError("esr_b200.train.conv2d needs CUDA tensors (there is no CPU path)")
        x, w, b = x.contiguous().float(), w.contiguous().float(), b.contiguous().float()
        B, Cin, H, W = x.shape
        Cout, _, k, _ = w.shape
        pad = k // 2
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        L = _lib.lib()
        y = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = L.esr_conv2d_workspace_bytes(B, Cin, H, W, Cout, k, stride)
            ws = _ws(nbytes, x.device)
            _lib.check(L.esr_conv2d_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), B, Cin, H, W, Cout, k, stride, act,
                                            _lib.ptr(y), _lib.ptr(ws), nbytes, _lib.stream_ptr()), "esr_conv2d_forward")
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (stride, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        stride, act = ctx.cfg
        B, Cin, H, W = x.shape
        Cout, _, k, _ = w.shape
        dy = dy.contiguous().float()
        L = _lib.lib()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.empty_like(w), torch.empty((Cout,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = L.esr_conv2d_workspace_bytes(B, Cin, H, W, Cout, k, stride)
            ws = _ws(nbytes, x.device)
            _lib.check(L.esr_conv2d_backward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), _lib.ptr(dy), B, Cin, H, W, Cout, k, stride,
                                             act, _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), nbytes,
                                             _lib.stream_ptr()), "esr_conv2d_backward")
        return dx, dw, db, None, None


def conv2d(x, w, b, stride=1, act=None):
    """act(conv2d(x, w, b, stride, padding=k//2)), differentiable; k = 3 or 1 (ConvLayer, models/submodules.py:159-200)."""
    return _Conv2dFn.apply(x, w, b, int(stride), _ACT[act])


class _DCNFn(torch.autograd.Function):
    """_DCNv2 (models/DCNv2/dcn_v2.py:17-68): same save_for_backward set, same five gradients."""

    @staticmethod
    def forward(ctx, inp, offset, mask, weight, bias, dg):
        ctx.dg = dg
        ctx.save_for_backward(inp, offset, mask, weight, bias)
        return dcn_v2_ext.dcn_v2_forward(inp, weight, bias, offset, mask, 3, 3, 1, 1, 1, 1, 1, 1, dg)

    @staticmethod
    def backward(ctx, grad_output):
        inp, offset, mask, weight, bias = ctx.saved_tensors
        gi, go, gm, gw, gb = dcn_v2_ext.dcn_v2_backward(inp, weight, bias, offset, mask, grad_output.contiguous(), 3, 3, 1, 1,
                                                        1, 1, 1, 1, ctx.dg)
        return gi, go, gm, gw, gb, None


def dcn_v2(inp, offset, mask, weight, bias, dg=8):
    return _DCNFn.apply(inp.contiguous(), offset.contiguous(), mask.contiguous(), weight, bias, dg)


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous().float(), target.contiguous().float()
        loss = torch.empty((1,), dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().esr_mse_loss(_lib.ptr(pred), _lib.ptr(target), pred.numel(), _lib.ptr(loss), _lib.ptr(grad),
                                               1.0, _lib.stream_ptr()), "esr_mse_loss")
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def mse_loss(pred, target):
    """nn.MSELoss() (train_ours_cnt_seq.py:774): mean over all elements; value and gradient from one kernel."""
    return _MSEFn.apply(pred, target)


# ------------------------------------------------------------------------------------------------------------------
# the differentiable window forward (structure of models/model.py:314-344 and the sub-blocks it calls)
# ------------------------------------------------------------------------------------------------------------------
def _cl(m, x, stride=1, act=None):
    return conv2d(x, m.conv2d.weight, m.conv2d.bias, stride, act)


def _local_time_corre(tp, f0, f1, f2):
    """TimePropagation.local_time_corre (models/model.py:77-89)."""
    def pred_map(a, b):
        return _cl(tp.pred_map[1], _cl(tp.pred_map[0], torch.cat([a, b], 1), act="relu"), act="sigmoid")

    x = torch.cat([f0 * pred_map(f0, f1), f1, f2 * pred_map(f1, f2)], 1)
    rb = tp.local_fusion[0]                                   # ResidualBlock (models/submodules.py:391-409)
    r = conv2d(x, rb.conv1.weight, rb.conv1.bias, 1, "relu")
    r = torch.relu(conv2d(r, rb.conv2.weight, rb.conv2.bias, 1, None) + x)
    return _cl(tp.local_fusion[1], r) + f1


def _gru_step(tp, x, h):
    """RecurrentConvLayer + ConvGRU (models/submodules.py:340-344, 496-514)."""
    g = tp.lstm.recurrent_block
    x = _cl(tp.lstm.conv, x, act="relu")
    if h is None:
        h = torch.zeros_like(x)
    xh = torch.cat([x, h], 1)
    z = conv2d(xh, g.update_gate.weight, g.update_gate.bias, 1, "sigmoid")
    r = conv2d(xh, g.reset_gate.weight, g.reset_gate.bias, 1, "sigmoid")
    o = conv2d(torch.cat([x, h * r], 1), g.out_gate.weight, g.out_gate.bias, 1, "tanh")
    return h * (1 - z) + o * z


def forward_window(model, inp, states=None):
    """inp BxNx2xHxW -> (Bx2xHxW, [h_fwd, h_rev]); differentiable w.r.t. the parameters, the input and `states`."""
    cfg = model._cfg
    N = cfg["num_frame"]
    B, n_in, Cin, H, W = inp.shape
    assert n_in == N
    Hc, Wc = 8 * math.ceil(H / 8), 8 * math.ceil(W / 8)
    x = inp.float()
    if (Hc, Wc) != (H, W):                                     # CropSize.pad (models/model_util.py:148-152)
        pt, pb = math.ceil(0.5 * (Hc - H)), math.floor(0.5 * (Hc - H))
        pl, pr = math.ceil(0.5 * (Wc - W)), math.floor(0.5 * (Wc - W))
        x = F.pad(x, (pl, pr, pt, pb))
    x = _cl(model.head, x.reshape(B * N, Cin, Hc, Wc), act="relu")
    pyramid = []
    for blk in model.feat_extract.convblock:                    # FeatsExtract (models/model.py:20-45)
        x = _cl(blk, x, stride=2, act="relu")
        pyramid.append(x)
    pyramid.reverse()
    C, h, w = pyramid[0].shape[1:]
    f = pyramid[0].view(B, N, C, h, w)

    tp = model.time_propagate                                   # TimePropagation.forward (models/model.py:126-153)
    ltc = []
    for i in range(N):
        lo, hi = max(i - 1, 0), min(i + 1, N - 1)
        ltc.append(_local_time_corre(tp, f[:, lo], f[:, i], f[:, hi]))
    h_f, h_r = states if states is not None else (None, None)
    fwd, rev = [], []
    for i in range(N):
        if cfg["gtc_frozen"]:
            h_f, h_r = None, None
        h_f = _gru_step(tp, ltc[i], h_f)
        h_r = _gru_step(tp, ltc[N - 1 - i], h_r)
        fwd.append(h_f)
        rev.append(h_r)
    new_states = [None, None] if cfg["gtc_frozen"] else [h_f, h_r]
    both = torch.cat([torch.stack(fwd, 1), torch.stack(rev[::-1], 1)], 2).view(B * N, 2 * C, h, w)
    prop = _cl(tp.global_fusion, both, act="relu").view(B, N, C, h, w) + f

    sf = model.spacetime_fuse                                   # STFusion.forward (models/model.py:208-291)
    mid = (N - 1) // 2
    center = prop[:, mid]
    fused = []
    for i in range(N):
        if i == mid:
            continue
        nb = prop[:, i]
        off_feat = _cl(sf.offset[1], _cl(sf.offset[0], torch.cat([nb, center], 1), act="relu"))
        om = conv2d(off_feat, sf.dcn.conv_offset_mask.weight, sf.dcn.conv_offset_mask.bias, 1, None)
        o1, o2, msk = torch.chunk(om, 3, dim=1)                 # DCN_sep.forward (models/DCNv2/dcn_v2.py:214-227)
        aligned = torch.relu(dcn_v2(nb, torch.cat((o1, o2), 1), torch.sigmoid(msk), sf.dcn.weight, sf.dcn.bias, 8))
        ft = _cl(sf.convblock[1], _cl(sf.convblock[0], torch.cat([aligned, center], 1), act="relu"))
        sk = _cl(sf.kernel, ft, act="sigmoid")
        mlp = sf.fc[0].layers
        ck = torch.relu(F.linear(ft.flatten(2).max(dim=2)[0], mlp[0].weight, mlp[0].bias))
        ck = torch.sigmoid(F.linear(ck, mlp[1].weight, mlp[1].bias))
        y = torch.cat([aligned * sk[:, 0:1] * ck[:, :C, None, None], center * sk[:, 1:2] * ck[:, C:, None, None]], 1)
        fused.append(_cl(sf.dcn_fusion[1], _cl(sf.dcn_fusion[0], y, act="relu")))
    fused.append(center)
    x = _cl(sf.dense_fusion[1], _cl(sf.dense_fusion[0], torch.cat(fused, 1), act="relu"))
    for lvl, ft in enumerate(pyramid):                          # scale_aggre + recons (models/model.py:253-291)
        att = _cl(sf.attens[lvl], ft, act="sigmoid")
        x = x + (ft * att).view(B, N, *ft.shape[1:]).mean(1)
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = _cl(sf.recons[lvl], x, act="relu")
    x = _cl(model.tail, x, act="relu")
    if (Hc, Wc) != (H, W):                                     # CropSize.crop (models/model_util.py:154-164)
        cy, cx = Hc // 2, Wc // 2
        x = x[..., cy - H // 2: cy + math.ceil(H / 2), cx - W // 2: cx + math.ceil(W / 2)].contiguous()
    return x, new_states


# ------------------------------------------------------------------------------------------------------------------
# optimizer and the step
# ------------------------------------------------------------------------------------------------------------------
class Adam:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay, amsgrad) semantics (the reference's optimizer,
    train_ours_cnt_seq.py:781 + config optimizer args) with one kernel launch per step: parameters and gradients are
    re-homed as views of two flat fp32 buffers (so DDP buckets and the update see contiguous memory)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not self.params[0].is_cuda:
            raise _lib.ESRError("esr_b200.train.Adam needs CUDA parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty((n,), dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros((n,), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += k
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.max_exp_avg_sq = torch.zeros_like(self.flat) if amsgrad else None
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.param_groups = [dict(params=self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)]

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        off = 0
        for p in self.params:                                    # keep the views if something replaced .grad
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += k

    def step(self):
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        with torch.cuda.device(self.flat.device):
            _lib.check(_lib.lib().esr_adam_step(_lib.ptr(self.flat), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                                _lib.ptr(self.exp_avg_sq), _lib.ptr(self.max_exp_avg_sq), self.flat.numel(),
                                                self.step_count, lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                _lib.stream_ptr()), "esr_adam_step")
        torch._C._increment_version(self.params)                  # the kernel wrote the parameters behind autograd's back:
        #                                                           bump their versions so cached inference blobs repack


def train_step(model, optimizer, frames, gt, num_frame=3, all_reduce=None):
    """One reference training iteration (train_ours_cnt_seq.py:209-235) on a batch of sequences.

    frames: BxLx2xHxW input count tensors (inp_scaled_cnt of each frame); gt: BxLx2xHxW target count tensors.
    Windows slide by one frame (dataloader/h5dataloader.py:229-231); the loss is the sum over windows of
    MSE(pred, gt[:, window middle]); one backward; optional `all_reduce(flat_grad)` (DDP's role); one Adam step.
    Returns the summed loss (a 0-dim tensor)."""
    L = frames.shape[1]
    mid = (num_frame - 1) // 2
    optimizer.zero_grad()
    net = model.module if hasattr(model, "module") else model
    net.reset_states()
    loss = 0
    for w in range(L - num_frame + 1):
        pred = model(frames[:, w:w + num_frame])
        loss = loss + mse_loss(pred, gt[:, w + mid])
    loss.backward()
    if all_reduce is not None:
        all_reduce(optimizer.flat_grad)
    optimizer.step()
    return loss.detach()
