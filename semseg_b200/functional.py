"""Autograd functions of the hot path, operating on NHWC bf16 activations.

Each Function is a thin scheduler of C-ABI kernel launches (semseg_b200/ops.py); the fp32 master weights
stay ordinary nn.Parameters (OIHW) so torch.optim.SGD, DistributedDataParallel and state_dict see exactly
what the reference's modules expose (tool/train.py:134-140,157).

conv + BatchNorm(train) + ReLU (+ residual) is three launches forward:
    conv_fprop (raw bf16 + per-tile partial statistics)  ->  bn_merge/finalize  ->  bn_apply
because training-mode BN needs the batch (and, under SyncBN, cross-rank) statistics of the complete conv
output before anything can be normalised (model/resnet.py:77-83). In eval mode the BN folds into the conv
epilogue and the whole thing is a single kernel.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from . import p2p
from . import precision
from .dist_utils import gather_rank_stats
from .ops import EPI_RAW, EPI_AFFINE, EPI_F32


# ------------------------------------------------------------------------------------------------ helpers
def packed(conv, need_dgrad=True, split=False):
    """bf16 operand slabs of conv.weight (hi + lo slabs when split), re-packed only when the parameter changed
    (optimizer step / load). The cache key is (parameter version, storage pointer): in-place edits through `.data`
    do not bump the version — call `invalidate_packs(model)` after such an edit."""
    w = conv.weight
    key = (w._version, w.data_ptr(), need_dgrad, split)
    cache = conv.__dict__.get("_sb_pack")
    if cache is not None and cache[0] == key:
        return cache[1]
    if cache is not None and cache[0][:2] == key[:2] and cache[0][3] == split and cache[0][2]:
        return cache[1]                       # a pack with dgrad slabs also serves a forward-only request
    pw = ops.pack_weights(w, need_dgrad=need_dgrad, split=split)
    conv.__dict__["_sb_pack"] = (key, pw)
    return pw


def invalidate_packs(model):
    """Forget every cached operand slab of `model` (needed after weights were modified through `.data`, which does not
    bump the version counter the cache is keyed on)."""
    for m in model.modules():
        m.__dict__.pop("_sb_pack", None)
        m.__dict__.pop("_sb_pack_patches", None)
    model.__dict__.pop("_sb_pack_plan", None)


def packed_patches(conv, split=False):
    """Operand slab of the stem conv in its patch form (ops.im2col3x3s2): wf [1][Cout][32] with column
    (r*3+s)*Cin + c = weight[:, c, r, s], zero padded ([2][1][Cout][32] = hi, lo slabs when split); re-made only when the
    parameter changed."""
    w = conv.weight
    key = (w._version, w.data_ptr(), "patches", split)
    cache = conv.__dict__.get("_sb_pack_patches")
    if cache is not None and cache[0] == key:
        return cache[1]
    cout, cin = w.shape[0], w.shape[1]
    flat = torch.zeros((1, cout, 32), dtype=torch.float32, device=w.device)
    flat[0, :, :9 * cin] = w.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    hi = flat.to(torch.bfloat16)
    wf = torch.stack([hi, (flat - hi.float()).to(torch.bfloat16)], 0) if split else hi
    conv.__dict__["_sb_pack_patches"] = (key, wf)
    return wf


def _is_patch_conv(conv, x):
    """The 3-channel stride-2 stem conv (model/resnet.py:106-108) on an input that needs no gradient."""
    return (conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.dilation == (1, 1) and conv.padding == (1, 1)
            and conv.in_channels <= 3 and conv.groups == 1 and x.shape[-1] >= 4)


def prepack(model, force=False):
    """Refresh the operand slabs of every native conv of `model` in one launch when its weights changed (call at the top
    of a training forward; force=True while the step is being captured into a CUDA graph, so that the re-pack is part
    of every replay). Convs keep working without it: `packed` falls back to the per-layer kernels."""
    convs = [m for m in model.modules() if isinstance(m, nn.Conv2d) and m.weight.is_cuda and
             m.weight.dtype == torch.float32 and m.weight.is_contiguous() and m.kernel_size[0] == m.kernel_size[1] and
             m.kernel_size[0] * m.kernel_size[1] <= ops.MAX_TAPS]
    if not convs:
        return
    split = precision.split_enabled()
    keys = [(c.weight._version, c.weight.data_ptr(), True, split) for c in convs]
    if not force and all(c.__dict__.get("_sb_pack", (None,))[0] == k for c, k in zip(convs, keys)):
        return                                            # nothing changed since the last pack
    plan = model.__dict__.get("_sb_pack_plan")
    weights = [c.weight.detach() for c in convs]
    if plan is None or not plan.valid_for(weights, split):
        plan = ops.WeightPackPlan(weights, split)
        model.__dict__["_sb_pack_plan"] = plan
    plan.refresh()
    for c, k, pw in zip(convs, keys, plan.packs):
        c.__dict__["_sb_pack"] = (k, pw)


def _sync_group(bn):
    """Process group when `bn` is a SyncBatchNorm that must synchronise, else None."""
    if isinstance(bn, nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
        pg = bn.process_group if bn.process_group is not None else dist.group.WORLD
        if dist.get_world_size(pg) > 1:
            return pg
    return None


def _bn_momentum(bn):
    # nn.BatchNorm semantics: momentum None = cumulative moving average
    if bn.momentum is None:
        return 1.0 / float(bn.num_batches_tracked.item() + 1)
    return bn.momentum


def _finalize_stats(stats, bn, pg, count_batches=True):
    """stats [3][C] local (mean, M2, count) -> (mean_invstd, scale_shift, world). Updates running stats."""
    world = 1
    if pg is not None:
        world = dist.get_world_size(pg)
        stats = gather_rank_stats(stats, pg)
    track = bn.track_running_stats and bn.running_mean is not None
    mom = _bn_momentum(bn) if track else 0.0
    mi, ss = ops.bn_finalize(stats, bn.weight, bn.bias, bn.eps, mom, bn.running_mean if track else None,
                             bn.running_var if track else None)
    if count_batches and track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return mi, ss, world


def _bn_backward(ctx_pg, world, dy, y, raw, mi, gamma, relu, want_dres, ss=None):
    """Shared BN(+ReLU) backward: returns (d_raw, dres, dgamma, dbeta). The ReLU mask comes from the saved output `y`,
    or — when `y` is None and `ss` (scale/shift) is given, i.e. no residual — is recomputed from `raw` (saves one
    tensor read in each of the two passes)."""
    n, h, w, c = raw.shape[-4:]
    if not dy.is_contiguous():
        dy = dy.contiguous()
    px = p2p.get_exchange(ctx_pg) if ctx_pg is not None else None
    if px is not None and 2 * c <= p2p.SLOT_FLOATS:
        # cross-rank sum over NVLink peer memory inside the reduction kernel (no NCCL call)
        local, sums = ops.bn_bwd_reduce_p2p(dy, y if relu else None, raw, mi, relu, ss if y is None else None, px)
        dbeta, dgamma = local[0], local[1]
    else:
        sums = ops.bn_bwd_reduce(dy, y if relu else None, raw, mi, relu, scale_shift=ss if y is None else None)
        dbeta, dgamma = sums[0], sums[1]
        if ctx_pg is not None:
            dbeta, dgamma = dbeta.clone(), dgamma.clone()  # local sums feed dgamma/dbeta (DDP averages them)
            dist.all_reduce(sums, group=ctx_pg)
    # under SyncBN the per-channel sample count is the one the forward exchange measured (mi row 2): exact also when the
    # ranks hold different numbers of pixels, like torch.nn.SyncBatchNorm's gathered counts
    count = float(n * h * w) if world == 1 else 0.0
    d_raw, dres, _ = ops.bn_bwd_apply(dy, y if relu else None, raw, mi, gamma, sums, count, relu, want_dres=want_dres,
                                      scale_shift=ss if y is None else None)
    return d_raw, dres, dgamma, dbeta


# ------------------------------------------------------------------------------------------------ conv+bn+act
class _CbaState:
    """What one conv+BN(+residual)(+ReLU) stage keeps for its backward pass."""
    __slots__ = ("xin", "raw", "y", "mi", "gamma", "ss", "pw", "k", "dil", "stride", "relu", "pg", "world",
                 "in_shape", "has_res")


def cba_forward(x, conv, bn, relu, residual, out=None, input_needs_grad=True):
    """conv (1x1 / 3x3; stride 1 with any dilation, or stride 2) + training BatchNorm + optional residual + ReLU.
    Returns (y, state). Three launches: conv_fprop (raw + per-CTA statistics) -> finalise (+ SyncBN exchange) -> apply.

    Stride-2 convs (stem conv1, layer2.0 conv2 / downsample — model/resnet.py:108,130-137) run on the same
    stride-1 tensor-core kernel through a 2x2 phase decomposition of the input (ops.space_to_phases): tap (r, s)
    reads phase ((r+1)&1, (s+1)&1) shifted by -1 or 0; dgrad is one small conv per phase, wgrad reads the phases."""
    split = ops.is_split(x)
    pw = packed(conv, split=split)
    k, dil, stride = conv.kernel_size[0], conv.dilation[0], conv.stride[0]
    n, h, w, cx = x.shape[-4:]
    wf = pw.wf
    if stride == 1:
        xin, img_add, out_nhw = x, None, None
        taps = ops.conv_taps(k, dil)
    elif not input_needs_grad and _is_patch_conv(conv, x):
        # stem conv: one 1x1 conv over 27-value input patches instead of 9 taps of a 3(->64)-channel K block
        xin, img_add, out_nhw = ops.im2col3x3s2(x, conv.in_channels), None, None
        taps, wf, stride = ops.conv_taps(1, 1), packed_patches(conv, split), 0    # stride 0 marks the patch form
    else:
        xin = ops.space_to_phases(x)                     # [4N, Hh, Wh, C]
        t2 = ops.conv_taps_s2(k, n)
        taps, img_add = [t[:3] for t in t2], [t[3] for t in t2]
        out_nhw = (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1)
    raw, sp = ops.conv_fprop(xin, wf, pw.cout, taps, stats=True, img_add=img_add, out_nhw=out_nhw)
    pg = _sync_group(bn)
    track = bn.track_running_stats and bn.running_mean is not None
    mom = _bn_momentum(bn) if track else 0.0
    rm, rv = (bn.running_mean, bn.running_var) if track else (None, None)
    px = p2p.get_exchange(pg) if pg is not None else None
    if pg is None:
        # single rank: merge the per-CTA partials and finalise in one launch
        mi, ss = ops.bn_finalize_partials(sp, bn.weight, bn.bias, bn.eps, mom, rm, rv)
        world = 1
    elif px is not None and 3 * pw.cout <= p2p.SLOT_FLOATS:
        # SyncBN: statistics exchanged over NVLink peer memory inside the finalise kernel (no NCCL call)
        mi, ss = ops.bn_finalize_p2p(sp, bn.weight, bn.bias, bn.eps, mom, rm, rv, px)
        world = px.world
    else:
        mi, ss, world = _finalize_stats(ops.bn_merge_partials(sp), bn, pg, count_batches=False)
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    y = ops.bn_apply(raw, ss, residual=residual, relu=relu, out=out)
    st = _CbaState()
    # ReLU mask for backward: with a residual it needs the saved output, otherwise it is recomputed from raw
    need_y = relu and residual is not None
    st.xin, st.raw, st.y, st.mi, st.gamma = xin, raw, (y if need_y else None), mi, bn.weight
    st.ss = ss if (relu and not need_y) else None
    st.pw, st.k, st.dil, st.stride, st.relu, st.pg, st.world = pw, k, dil, stride, relu, pg, world
    st.in_shape, st.has_res = (n, h, w, cx), residual is not None
    return y, st


def cba_backward(st, dy, need_dx=True, need_dw=True, need_dres=False, dx_add=None):
    """Backward of cba_forward: returns (dx, dw, dgamma, dbeta, dres). `dx_add` (same shape as dx) is summed into
    dx inside the dgrad epilogue (AFFINE mode with a residual operand) — this is how gradient fan-in is fused."""
    pw = st.pw
    n, h, w, cx = st.in_shape
    d_raw, dres, dgamma, dbeta = _bn_backward(st.pg, st.world, dy, st.y, st.raw, st.mi, st.gamma, st.relu,
                                              st.has_res and need_dres, st.ss)
    dx = dw = None
    if st.stride == 0:      # patch form of the stem conv (no input gradient by construction)
        if need_dx:
            raise RuntimeError("semseg_b200: the patch form of the stem conv was chosen but dx is requested")
        if need_dw:
            cin = pw.cin
            dwp = ops.conv_wgrad(st.xin, d_raw, 32, pw.cout, ops.conv_taps(1, 1))          # [Cout, 32, 1, 1]
            dw = dwp[:, :9 * cin, 0, 0].reshape(pw.cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous()
        return dx, dw, dgamma, dbeta, dres
    if st.stride == 1:
        if need_dx:
            if dx_add is not None:
                dx, _ = ops.conv_fprop(d_raw, pw.wd, pw.cin, ops.conv_taps(st.k, st.dil, transpose=True),
                                       epi=EPI_AFFINE, residual=dx_add)
            else:
                dx, _ = ops.conv_fprop(d_raw, pw.wd, pw.cin, ops.conv_taps(st.k, st.dil, transpose=True))
        if need_dw:
            dw = ops.conv_wgrad(st.xin, d_raw, cx, pw.cout, ops.conv_taps(st.k, st.dil))
    else:
        t2 = ops.conv_taps_s2(st.k, n)
        if need_dx:
            if pw.cin % 64 != 0:
                raise NotImplementedError("semseg_b200: input gradient of a stride-2 conv needs Cin % 64 == 0")
            hh, wh = st.xin.shape[-3], st.xin.shape[-2]
            dxp = ops.empty_act((4 * n, hh, wh, pw.cin), ops.is_split(d_raw), dy.device)
            for q in range(4):
                sub = [(-t[0], -t[1], t[2]) for t in t2 if t[4] == (q >> 1, q & 1)]
                part = ops.act_batch_slice(dxp, q * n, (q + 1) * n)
                if sub:      # dx of phase q: conv of d_raw with the taps that read this phase (mirrored shifts)
                    ops.conv_fprop(d_raw, pw.wd, pw.cin, sub, out=part, out_nhw=(n, hh, wh))
                else:
                    part.zero_()
            dx = ops.phases_to_space(dxp, n, h, w)
            if dx_add is not None:
                dx = ops.add_act(dx, dx_add)
        if need_dw:
            dw = ops.conv_wgrad(st.xin, d_raw, cx, pw.cout, [t[:2] for t in t2], img_add=[t[3] for t in t2])
    if dw is not None and cx != pw.cin and st.stride != 0:
        dw = dw[:, :pw.cin].contiguous()          # input channels were zero-padded to a multiple of 8 (stem)
    return dx, dw, dgamma, dbeta, dres


class _ConvBnAct(torch.autograd.Function):
    """Autograd wrapper of one cba_forward / cba_backward stage."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, conv, bn, relu, out):
        y, st = cba_forward(x, conv, bn, relu, residual, out, input_needs_grad=ctx.needs_input_grad[0])
        ctx.st = st
        if out is not None:
            ctx.mark_dirty(out)
        return y

    @staticmethod
    def backward(ctx, dy):
        ni = ctx.needs_input_grad
        dx, dw, dgamma, dbeta, dres = cba_backward(ctx.st, dy, need_dx=ni[0], need_dw=ni[1], need_dres=ni[4])
        ctx.st = None
        return dx, dw, dgamma, dbeta, dres, None, None, None, None


class _BottleneckFn(torch.autograd.Function):
    """A whole Bottleneck (model/resnet.py:74-94) as one autograd node: conv1-bn1-relu, conv2-bn2-relu, conv3-bn3,
    (+ downsample conv-bn), residual add, relu. Besides saving three autograd nodes per block, the backward pass
    fuses the gradient fan-in (dx = dgrad(conv1) + d(residual branch)) into the dgrad epilogue instead of a separate
    elementwise add."""

    @staticmethod
    def forward(ctx, x, blk, *params):
        y1, s1 = cba_forward(x, blk.conv1, blk.bn1, True, None)
        y2, s2 = cba_forward(y1, blk.conv2, blk.bn2, True, None)
        if blk.downsample is not None:
            res, sd = cba_forward(x, blk.downsample[0], blk.downsample[1], False, None)
        else:
            res, sd = x, None
        y3, s3 = cba_forward(y2, blk.conv3, blk.bn3, True, res)
        ctx.states = (s1, s2, s3, sd)
        return y3

    @staticmethod
    def backward(ctx, dy):
        s1, s2, s3, sd = ctx.states
        ctx.states = None
        need_dx = ctx.needs_input_grad[0]
        d2, dw3, dg3, db3, dres = cba_backward(s3, dy, need_dres=True)
        d1, dw2, dg2, db2, _ = cba_backward(s2, d2)
        grads_ds = ()
        if sd is not None:
            dxd, dwd, dgd, dbd, _ = cba_backward(sd, dres, need_dx=need_dx)
            dres_to_x = dxd
            grads_ds = (dwd, dgd, dbd)
        else:
            dres_to_x = dres
        dx, dw1, dg1, db1, _ = cba_backward(s1, d1, need_dx=need_dx, dx_add=dres_to_x if need_dx else None)
        return (dx, None, dw1, dg1, db1, dw2, dg2, db2, dw3, dg3, db3) + grads_ds


def bottleneck(x, blk):
    """Fused Bottleneck when every stage is covered by the native kernels in training mode, else stage by stage."""
    convs = [blk.conv1, blk.conv2, blk.conv3] + ([blk.downsample[0]] if blk.downsample is not None else [])
    bns = [blk.bn1, blk.bn2, blk.bn3] + ([blk.downsample[1]] if blk.downsample is not None else [])
    cins = [x.shape[-1], blk.conv1.out_channels, blk.conv2.out_channels, x.shape[-1]]
    fused = (torch.is_grad_enabled() and all(b.training or b.running_mean is None for b in bns) and
             all(_is_native_conv(c, ci) for c, ci in zip(convs, cins)) and
             (blk.downsample is None or len(blk.downsample) == 2))
    if fused:
        params = [blk.conv1.weight, blk.bn1.weight, blk.bn1.bias, blk.conv2.weight, blk.bn2.weight, blk.bn2.bias,
                  blk.conv3.weight, blk.bn3.weight, blk.bn3.bias]
        if blk.downsample is not None:
            params += [blk.downsample[0].weight, blk.downsample[1].weight, blk.downsample[1].bias]
        return _BottleneckFn.apply(x, blk, *params)
    y = conv_bn_act(x, blk.conv1, blk.bn1, relu=True)
    y = conv_bn_act(y, blk.conv2, blk.bn2, relu=True)
    residual = conv_bn_act(x, blk.downsample[0], blk.downsample[1], relu=False) if blk.downsample is not None else x
    return conv_bn_act(y, blk.conv3, blk.bn3, relu=True, residual=residual)


def _is_native_conv(conv, cin):
    """Convs the tensor-core kernel covers: 1x1 / 3x3 'same' convs, stride 1 (any dilation) or stride 2 (dilation 1)."""
    ok = (conv.kernel_size in ((1, 1), (3, 3)) and conv.groups == 1 and conv.bias is None and
          conv.padding == (conv.dilation[0] * (conv.kernel_size[0] // 2),) * 2 and
          conv.dilation[0] == conv.dilation[1] and cin % 8 == 0 and conv.out_channels % 64 == 0)
    if conv.stride == (1, 1):
        return ok
    return ok and conv.stride == (2, 2) and conv.dilation == (1, 1)


def _require_native(conv, cin):
    """There is exactly one backend: a convolution the sm_100a kernel does not cover is an error, never a library
    (cuDNN) fallback. Every convolution of PSPNet / PSANet (model/resnet.py, model/pspnet.py, model/psanet.py) is covered."""
    if not _is_native_conv(conv, cin):
        raise NotImplementedError(
            "semseg_b200: convolution %r on %d input channels is outside the tensor-core kernel's coverage (1x1 / 3x3 "
            "'same' convs without bias, stride 1 with any dilation or stride 2 undilated, Cin %% 8 == 0, Cout %% 64 == 0); "
            "there is no library fallback" % (conv, cin))


def conv_bn_act(x, conv, bn, relu=True, residual=None, out=None):
    """NHWC activation -> NHWC activation: conv -> BatchNorm -> (+residual) -> (ReLU), training or eval semantics of `bn`."""
    use_batch_stats = bn.training or (bn.running_mean is None)
    _require_native(conv, x.shape[-1])
    if use_batch_stats:
        return _ConvBnAct.apply(x, conv.weight, bn.weight, bn.bias, residual, conv, bn, relu, out)
    # Eval-mode BatchNorm: conv + folded BN + residual + ReLU are ONE kernel. It has no backward: the reference's
    # validate() (tool/train.py:353-359) calls model.eval()(input) without torch.no_grad() and never back-propagates,
    # so the result is returned detached (a later .backward() through it raises torch's usual "does not require grad").
    ss = ops.bn_fold_eval(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    split = ops.is_split(x)
    with torch.no_grad():
        pw = packed(conv, need_dgrad=False, split=split)
        if conv.stride == (1, 1):
            y, _ = ops.conv_fprop(x, pw.wf, pw.cout, ops.conv_taps(conv.kernel_size[0], conv.dilation[0]),
                                  epi=EPI_AFFINE, relu=relu, scale=ss[0], shift=ss[1], residual=residual, out=out)
        else:
            n, h, w, _ = x.shape[-4:]
            t2 = ops.conv_taps_s2(conv.kernel_size[0], n)
            y, _ = ops.conv_fprop(ops.space_to_phases(x), pw.wf, pw.cout, [t[:3] for t in t2], epi=EPI_AFFINE,
                                  relu=relu, scale=ss[0], shift=ss[1], residual=residual, out=out,
                                  img_add=[t[3] for t in t2], out_nhw=(n, (h - 1) // 2 + 1, (w - 1) // 2 + 1))
    return y


# ------------------------------------------------------------------------------------------------ classifier
class _ConvBiasF32(torch.autograd.Function):
    """1x1 conv with bias producing fp32 NHWC logits (model/pspnet.py:69,77)."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv):
        pw = packed(conv, split=ops.is_split(x))
        y, _ = ops.conv_fprop(x, pw.wf, pw.cout, ops.conv_taps(1, 1), epi=EPI_F32, shift=bias)
        ctx.save_for_backward(x)
        ctx.pw = pw
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        pw = ctx.pw
        n, h, w, c = dy.shape
        cp = pw.wd.shape[-1]  # Cout rounded up to 8 (zero padded operand)
        dyb = ops.f32_to_act(dy.contiguous(), ops.is_split(x))      # [N,h,w,cp], padding columns zero
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx, _ = ops.conv_fprop(dyb, pw.wd, pw.cin, ops.conv_taps(1, 1))
        if ctx.needs_input_grad[1]:
            dwp = ops.conv_wgrad(x, dyb, pw.cin, cp, ops.conv_taps(1, 1))
            dw = dwp[:c].contiguous()
        if ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 1, 2))
        return dx, dw, db, None


def conv_bias_f32(x, conv):
    assert conv.kernel_size == (1, 1) and conv.stride == (1, 1) and x.shape[-1] % 8 == 0
    if torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad):
        return _ConvBiasF32.apply(x, conv.weight, conv.bias, conv)
    pw = packed(conv, need_dgrad=False, split=ops.is_split(x))
    y, _ = ops.conv_fprop(x, pw.wf, pw.cout, ops.conv_taps(1, 1), epi=EPI_F32, shift=conv.bias)
    return y


# ------------------------------------------------------------------------------------------------ fused tail
class _UpsampleCE(torch.autograd.Function):
    """bilinear x8 upsample (align_corners) + CrossEntropyLoss(ignore_index, mean) + argmax in one kernel each way
    (model/pspnet.py:94-103) — the [N, classes, H, W] logits tensor is never materialised."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        info, amax, lse = ops.upsample_ce_fwd(logits, target, ignore_index)
        ctx.save_for_backward(logits, target, lse, info)
        ctx.ignore_index = ignore_index
        ctx.mark_non_differentiable(amax)
        return info[0], amax

    @staticmethod
    def backward(ctx, grad_loss, _grad_amax):
        logits, target, lse, info = ctx.saved_tensors
        return ops.upsample_ce_bwd(logits, target, ctx.ignore_index, lse, info, grad_loss), None, None


def fused_tail_supported(criterion, logits, target, zoom_factor, x_size=None):
    """The fused kernel implements exactly nn.CrossEntropyLoss(ignore_index=k) with default options at zoom 8.
    `logits` fp32 NHWC, or None with the NCHW input size `x_size` (decision before the network has run)."""
    if not (type(criterion) is nn.CrossEntropyLoss and criterion.weight is None and criterion.reduction == 'mean'
            and getattr(criterion, 'label_smoothing', 0.0) == 0.0 and zoom_factor == 8 and target is not None
            and target.dtype == torch.int64 and target.dim() == 3):
        return False
    if logits is None:
        return target.shape[1] == x_size[2] and target.shape[2] == x_size[3]
    return (logits.shape[-1] <= 256 and target.shape[1] == 8 * (logits.shape[1] - 1) + 1
            and target.shape[2] == 8 * (logits.shape[2] - 1) + 1)


def upsample_ce(logits, target, ignore_index):
    """-> (mean CE loss scalar, argmax int64 [N,H,W])."""
    return _UpsampleCE.apply(logits, target.contiguous(), ignore_index)


# ------------------------------------------------------------------------------------------------ pyramid pooling
class _PPMLink:
    """Carries the identity-branch gradient of x from the concat node to the pooling node of one PPM invocation.

    x feeds both AdaptiveAvgPool (all bins) and the concat (model/pspnet.py:20-26); autograd would add the two gradients
    with a strided ATen kernel (one of them is a channel slice of the 4096-wide concat gradient). The concat node always
    runs first in backward (the pooled branch reaches x only through it), so it parks its slice here and returns no
    gradient for x; the pooling node's kernel adds the slice while it writes dx."""
    __slots__ = ("dx_identity",)

    def __init__(self):
        self.dx_identity = None


class _PPMPool(torch.autograd.Function):
    """AdaptiveAvgPool2d of every bin in one launch (model/pspnet.py:14)."""

    @staticmethod
    def forward(ctx, x, bins, link):
        ctx.bins, ctx.shape, ctx.link = bins, tuple(x.shape[-4:]), link
        return tuple(ops.ppm_pool(x, bins))

    @staticmethod
    def backward(ctx, *dpooled):
        n, h, w, c = ctx.shape
        add = None
        if ctx.link is not None:
            add, ctx.link.dx_identity = ctx.link.dx_identity, None
        return ops.ppm_pool_bwd(list(dpooled), ctx.bins, n, h, w, c, add=add), None, None


class _PPMUpsampleConcat(torch.autograd.Function):
    """cat([x, bilinear(f_1), ..., bilinear(f_nb)], channel) written in place (model/pspnet.py:25-26)."""

    @staticmethod
    def forward(ctx, x, bins, link, *feats):
        ctx.bins, ctx.c, ctx.cr, ctx.link = bins, x.shape[-1], feats[0].shape[-1], link
        return ops.ppm_upsample_concat(x, list(feats), bins)

    @staticmethod
    def backward(ctx, dout):
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dfeats = ops.ppm_upsample_bwd(dout, ctx.c, ctx.bins, ctx.cr)
        dx = dout[..., :ctx.c]
        if ctx.link is not None and ctx.needs_input_grad[0] and all(ctx.needs_input_grad[3:]):
            ctx.link.dx_identity = dx      # summed into dx by the pooling node's kernel (see _PPMLink)
            dx = None
        return (dx, None, None) + tuple(dfeats)


def ppm_pool(x, bins, link=None):
    return _PPMPool.apply(x, tuple(bins), link)


def ppm_upsample_concat(x, feats, bins, link=None):
    return _PPMUpsampleConcat.apply(x, tuple(bins), link, *feats)


def ppm_link():
    return _PPMLink()


# ------------------------------------------------------------------------------------------------ fused PSA attention
class _PSAAttend(torch.autograd.Function):
    """psa_mask -> softmax over the source positions -> aggregation bmm -> 1/normalization_factor (model/psanet.py:81-91)
    as one kernel each way; nothing [HW x HW] is written to HBM (the backward recomputes the probabilities from the
    logits and the saved per-target (max, 1/sum))."""

    @staticmethod
    def forward(ctx, attn, feat, psa_type, mask_h, mask_w, scale):
        out, stats = ops.psa_attend(attn, feat, psa_type, mask_h, mask_w, scale)
        ctx.save_for_backward(attn, feat, out, stats)
        ctx.cfg = (psa_type, mask_h, mask_w, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        attn, feat, out, stats = ctx.saved_tensors
        psa_type, mask_h, mask_w, scale = ctx.cfg
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dattn = dfeat = None
        if ctx.needs_input_grad[1]:
            dfeat, _ = ops.psa_attend(attn, dout, psa_type, mask_h, mask_w, scale, stats=stats, mode=1)
        if ctx.needs_input_grad[0]:
            dattn = ops.psa_attend_bwd_attn(attn, stats, feat, out, dout, psa_type, mask_h, mask_w, scale)
        return dattn, dfeat, None, None, None, None


def psa_attend(attn, feat, psa_type, mask_h, mask_w, scale):
    """attn fp32 NHWC [N,h,w,mask_h*mask_w], feat NHWC activation [N,h,w,512] -> aggregated features [N,h,w,512]."""
    return _PSAAttend.apply(attn.contiguous(), feat, psa_type, mask_h, mask_w, scale)


def psa_attend_supported(feat, mask_h, mask_w):
    return feat.shape[-1] == 512 and feat.shape[-2] <= 128 and mask_h % 2 == 1 and mask_w % 2 == 1


# ------------------------------------------------------------------------------------------------ bilinear resize
class _ResizeBilinear(torch.autograd.Function):
    """F.interpolate(mode='bilinear', align_corners=True) on an NHWC activation (model/psanet.py:61,97), deterministic
    gather backward."""

    @staticmethod
    def forward(ctx, x, size):
        ctx.in_size = tuple(x.shape[-3:-1])
        return ops.resize_bilinear(x, size)

    @staticmethod
    def backward(ctx, dy):
        return ops.resize_bilinear_bwd(dy if dy.is_contiguous() else dy.contiguous(), ctx.in_size), None


def resize_bilinear(x, size):
    return _ResizeBilinear.apply(x, tuple(size))


# ------------------------------------------------------------------------------------------------ misc NHWC ops
def to_nhwc_bf16(x_nchw):
    """fp32 NCHW module input -> NHWC activation (channels padded to a multiple of 8 with zeros) in the storage form of
    the current precision mode (precision.py): plain bf16, or (hi, lo) bf16 planes for bf16x3."""
    return ops.nchw_to_nhwc_bf16(x_nchw.contiguous().float(), split=precision.split_enabled())


class _ActToF32(torch.autograd.Function):
    """activation -> fp32 NHWC (differentiable)."""

    @staticmethod
    def forward(ctx, x):
        ctx.split = ops.is_split(x)
        return ops.act_to_f32(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.f32_to_act(dy.contiguous(), ctx.split)


class _F32ToAct(torch.autograd.Function):
    """fp32 NHWC -> activation in the requested storage form (differentiable)."""

    @staticmethod
    def forward(ctx, x, split):
        return ops.f32_to_act(x.contiguous(), split)

    @staticmethod
    def backward(ctx, dy):
        return ops.act_to_f32(dy.contiguous()), None


def act_to_f32(x):
    return _ActToF32.apply(x)


def f32_to_act(x, split):
    return _F32ToAct.apply(x, split)


def to_nchw_f32(y):
    """NHWC activation -> fp32 NCHW (what the reference's modules return)."""
    if not (torch.is_grad_enabled() and y.requires_grad):
        return ops.nhwc_bf16_to_nchw(y)
    return act_to_f32(y).permute(0, 3, 1, 2)


class _Fork(torch.autograd.Function):
    """x -> k aliases of x whose gradients are summed by the split-aware add kernel (autograd's own accumulation would
    add the hi and lo planes of two split gradients separately, losing the error compensation)."""

    @staticmethod
    def forward(ctx, x, k):
        return tuple(x.view_as(x) for _ in range(k))

    @staticmethod
    def backward(ctx, *grads):
        acc = None
        for g in grads:
            if g is None:
                continue
            g = g if g.is_contiguous() else g.contiguous()
            acc = g if acc is None else ops.add_act(acc, g)
        return acc, None


def fork(x, k=2):
    """k handles on x for k consumers (gradient fan-in through one native add per extra branch)."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        return (x,) * k
    return _Fork.apply(x, k)


class _ScaleNC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.save_for_backward(scale)
        return ops.scale_nc(x, scale)

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        return ops.scale_nc(dy if dy.is_contiguous() else dy.contiguous(), scale), None


def dropout2d_nhwc(x, p, training):
    """nn.Dropout2d on NHWC: one Bernoulli per (n, c), kept channels scaled by 1/(1-p) (model/pspnet.py:68,76)."""
    if not training or p == 0.0:
        return x
    n, c = x.shape[-4], x.shape[-1]
    scale = torch.empty((n, c), device=x.device, dtype=torch.float32).bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))
    return _ScaleNC.apply(x, scale)


class _MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, code = ops.maxpool3x3s2_fwd(x, want_argcode=True)
        ctx.save_for_backward(code)
        ctx.in_shape = tuple(x.shape[-4:])
        return y

    @staticmethod
    def backward(ctx, dy):
        (code,) = ctx.saved_tensors
        return ops.maxpool3x3s2_bwd(code, dy, ctx.in_shape)


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def maxpool_nhwc(x, pool):
    if (_pair(pool.kernel_size) == (3, 3) and _pair(pool.stride) == (2, 2) and _pair(pool.padding) == (1, 1)
            and _pair(pool.dilation) == (1, 1) and not pool.ceil_mode and x.shape[-1] % 8 == 0):
        return _MaxPool3x3s2.apply(x)
    raise NotImplementedError("semseg_b200: only MaxPool2d(kernel_size=3, stride=2, padding=1) (model/resnet.py:115) is "
                              "implemented; there is no library fallback (got %r)" % (pool,))
