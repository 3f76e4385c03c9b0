"""Tensor-level wrappers over the C-ABI (semseg_b200/_lib.py): every function takes CUDA tensors,
launches on torch's current stream and returns tensors. PyTorch is used for device memory and streams
only; the arithmetic happens in libsemseg_b200.so. No function here has a CPU or eager fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvDesc, WgradDesc, EPI_RAW, EPI_AFFINE, EPI_F32, MAX_TAPS


# bf16x3: K blocks (64-channel block x tap) one tensor-core accumulation chain may span (8 x 4 x 3 = 96 MMA steps; the
# truncating fp32 accumulation of tcgen05 loses ~2^-24 per step towards zero, tools/probe_accum.py)
X3_MAX_KBLOCKS = 8

_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """torch's current CUDA stream of the current device as a raw cudaStream_t. Called once per kernel launch (~700
    times per training step), so it goes through the two C accessors instead of building a torch.cuda.Stream object
    (which was a quarter of the host-side step time, tools/profile_cpu.py)."""
    if _raw_stream is not None and _cur_device is not None:
        return ctypes.c_void_p(_raw_stream(_cur_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SemsegError("semseg_b200 ops require CUDA tensors (no CPU fallback); got device %s" % t.device)


# ------------------------------------------------------------------------------------------------ activation storage
# An activation is either a plain bf16 NHWC tensor [N,H,W,C] ("bf16", the speed configuration) or a SPLIT tensor
# [2,N,H,W,C]: plane 0 = hi = bf16(v), plane 1 = lo = bf16(v - hi) (16 mantissa bits, csrc/act.cuh). The conv kernels
# consume split operands as three K segments (bf16x3); every other kernel reads hi + lo and re-splits its result. The
# storage form is chosen once per model call (precision.py) and every op follows the form of its input.
def is_split(t):
    return t is not None and t.dim() == 5


def _nhwc_meta(t):
    """(N, H, W, C, pitch) of a bf16 NHWC activation (plain 4-D or split 5-D) whose channel dim may be a slice of a
    wider buffer."""
    assert t.dtype == torch.bfloat16 and t.dim() in (4, 5), (t.shape, t.dtype)
    if t.dim() == 5:
        assert t.shape[0] == 2, "split activation must be [2,N,H,W,C] (hi, lo planes)"
    n, h, w, c = t.shape[-4:]
    sn, sh, sw, sc = t.stride()[-4:]
    assert sc == 1 and sh == sw * w and sn == sh * h, "NHWC tensor must be pixel-contiguous (stride %s)" % (t.stride(),)
    return n, h, w, c, sw


def _lo(t):
    """Pointer to the lo plane of a split activation, NULL for a plain one."""
    if t is not None and t.dim() == 5:
        return ctypes.c_void_p(t.data_ptr() + 2 * t.stride(0))
    return ctypes.c_void_p(0)


def _lo_int(t):
    return t.data_ptr() + 2 * t.stride(0) if (t is not None and t.dim() == 5) else 0


def empty_act(shape, split, device):
    """Uninitialised activation of NHWC shape `shape` in the requested storage form."""
    return torch.empty(((2,) + tuple(shape)) if split else tuple(shape), dtype=torch.bfloat16, device=device)


def _same_form(*ts):
    forms = {t.dim() == 5 for t in ts if t is not None}
    assert len(forms) <= 1, "activations of one call must all be plain or all be split"


def act_batch_slice(t, a, b):
    """Images a..b of an activation (either storage form)."""
    return t[:, a:b] if t.dim() == 5 else t[a:b]


def round_up(a, b):
    return (a + b - 1) // b * b


# ------------------------------------------------------------------------------------------------ psa mask
def psamask_fwd(x, psa_type, mask_h, mask_w):
    _require_cuda(x)
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, h * w, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.semseg_psamask_fwd(psa_type, _ptr(x), _ptr(out), n, h, w, mask_h, mask_w, _stream()),
               "semseg_psamask_fwd")
    return out


def psamask_bwd(grad_out, psa_type, mask_h, mask_w):
    _require_cuda(grad_out)
    lib = _lib.load()
    assert grad_out.dtype == torch.float32 and grad_out.is_contiguous()
    n, hw, h, w = grad_out.shape
    din = torch.empty((n, mask_h * mask_w, h, w), dtype=torch.float32, device=grad_out.device)
    _lib.check(lib.semseg_psamask_bwd(psa_type, _ptr(grad_out), _ptr(din), n, h, w, mask_h, mask_w, _stream()),
               "semseg_psamask_bwd")
    return din


# ------------------------------------------------------------------------------------------------ fused PSA attention
def psa_attend(attn, feat, psa_type, mask_h, mask_w, scale, stats=None, mode=0):
    """mode 0: (out, stats) = fused mask-gather -> softmax -> aggregation (model/psanet.py:81-91) of the fp32 NHWC logits
    `attn` [N,h,w,>=mask_h*mask_w] and the NHWC activation `feat` [N,h,w,512]; mode 1: the feature gradient (pass dout as
    `feat` and the forward's `stats`)."""
    _require_cuda(attn, feat)
    lib = _lib.load()
    assert attn.dtype == torch.float32 and attn.dim() == 4 and attn.is_contiguous()
    n, h, w, c, fp = _nhwc_meta(feat)
    assert tuple(attn.shape[:3]) == (n, h, w) and attn.shape[3] >= mask_h * mask_w
    out = empty_act((n, h, w, c), is_split(feat), feat.device)
    if stats is None:
        assert mode == 0
        stats = torch.empty((n, h * w, 2), dtype=torch.float32, device=feat.device)
    _lib.check(lib.semseg_psa_attend(mode, psa_type, _ptr(attn), attn.shape[3], _ptr(feat), _lo(feat), fp, _ptr(stats),
                                     _ptr(out), _lo(out), c, n, h, w, mask_h, mask_w, c, float(scale), _stream()),
               "semseg_psa_attend")
    return out, stats


def psa_attend_bwd_attn(attn, stats, feat, out, dout, psa_type, mask_h, mask_w, scale):
    """Gradient of psa_attend w.r.t. the attention logits (same shape as attn, zero outside the mask windows)."""
    lib = _lib.load()
    n, h, w, c, fp = _nhwc_meta(feat)
    op, dp = _nhwc_meta(out)[4], _nhwc_meta(dout)[4]
    _same_form(feat, out, dout)
    dattn = torch.empty_like(attn)
    _lib.check(lib.semseg_psa_attend_bwd_attn(psa_type, _ptr(attn), attn.shape[3], _ptr(stats), _ptr(feat), _lo(feat), fp,
                                              _ptr(out), _lo(out), op, _ptr(dout), _lo(dout), dp, _ptr(dattn), n, h, w,
                                              mask_h, mask_w, c, float(scale), _stream()),
               "semseg_psa_attend_bwd_attn")
    return dattn


# ------------------------------------------------------------------------------------------------ weights
class PackedWeight:
    """bf16 operand slabs of one conv weight: wf [taps][Cout][Cin_p] (fprop), wd [taps][Cin][Cout_p] (dgrad); with
    split=True each is [2][taps][rows][cols] = the hi slab followed by the lo slab (bf16x3 operand mode)."""
    __slots__ = ("wf", "wd", "cout", "cin", "taps", "ksize", "split")

    def __init__(self, wf, wd, cout, cin, taps, ksize, split=False):
        self.wf, self.wd, self.cout, self.cin, self.taps, self.ksize, self.split = wf, wd, cout, cin, taps, ksize, split


def _slab(taps, rows, cols, split, device):
    return torch.empty(((2,) if split else ()) + (taps, rows, cols), dtype=torch.bfloat16, device=device)


def pack_weights(w, need_dgrad=True, split=False):
    """w: fp32 OIHW parameter -> PackedWeight (one fused launch pair)."""
    _require_cuda(w)
    lib = _lib.load()
    w = w.detach()
    assert w.dtype == torch.float32 and w.dim() == 4
    if not w.is_contiguous():
        w = w.contiguous()
    cout, cin, kh, kw = w.shape
    assert kh == kw
    taps = kh * kw
    cin_p, cout_p = round_up(cin, 8), round_up(cout, 8)
    wf = _slab(taps, cout, cin_p, split, w.device)
    wd = _slab(taps, cin, cout_p, split, w.device) if need_dgrad else None
    _lib.check(lib.semseg_pack_weights(_ptr(w), cout, cin, taps, _ptr(wf), cout, cin_p, _ptr(wd),
                                       cin if need_dgrad else 0, cout_p if need_dgrad else 0, int(bool(split)),
                                       _stream()),
               "semseg_pack_weights")
    return PackedWeight(wf, wd, cout, cin, taps, kh, bool(split))


class WeightPackPlan:
    """Persistent bf16 operand slabs for a list of conv weights, refreshed in ONE launch (semseg_pack_weights_multi).

    The fp32 OIHW parameters stay the masters (optimizer / DDP / checkpoints); after an optimizer step every conv of the
    model needs new slabs, which costs 2 launches per conv on the per-layer path. The plan owns one (wf, wd) pair per
    conv and a device-side item table; `refresh()` re-packs all of them into the same buffers."""

    def __init__(self, weights, split=False):
        import ctypes
        _require_cuda(*weights)
        self.split = bool(split)
        self.weights = [w for w in weights]
        self.ptrs = [w.data_ptr() for w in weights]
        self.packs = []
        items = (_lib.PackItem * len(weights))()
        tile0, max_taps = 0, 1
        for k, w in enumerate(weights):
            assert w.dtype == torch.float32 and w.dim() == 4 and w.is_contiguous() and w.shape[2] == w.shape[3]
            cout, cin, kh, _ = w.shape
            taps = kh * kh
            assert taps <= MAX_TAPS
            cin_p, cout_p = round_up(cin, 8), round_up(cout, 8)
            wf = _slab(taps, cout, cin_p, self.split, w.device)
            wd = _slab(taps, cin, cout_p, self.split, w.device)
            self.packs.append(PackedWeight(wf, wd, cout, cin, taps, kh, self.split))
            it = items[k]
            it.w, it.wf, it.wd = w.data_ptr(), wf.data_ptr(), wd.data_ptr()
            it.Cout, it.Cin, it.taps, it.cols_f, it.cols_d = cout, cin, taps, cin_p, cout_p
            it.tile0, it.tiles_ci, it.split = tile0, (cin_p + 31) // 32, int(self.split)
            tile0 += it.tiles_ci * ((cout_p + 31) // 32)
            max_taps = max(max_taps, taps)
        self.n_items, self.n_tiles, self.max_taps = len(weights), tile0, max_taps
        raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8)
        self.items_dev = raw.to(weights[0].device)

    def valid_for(self, weights, split=False):
        return (bool(split) == self.split and len(weights) == len(self.ptrs) and
                all(w.data_ptr() == p for w, p in zip(weights, self.ptrs)))

    def refresh(self):
        lib = _lib.load()
        _lib.check(lib.semseg_pack_weights_multi(_ptr(self.items_dev), self.n_items, self.n_tiles, self.max_taps,
                                                 _stream()), "semseg_pack_weights_multi")


def conv_taps(ksize, dilation, transpose=False):
    """[(dh, dw, wtap)] of a stride-1 'same' conv; transpose=True gives the dgrad taps."""
    taps = []
    half = ksize // 2
    for r in range(ksize):
        for s in range(ksize):
            dh, dw = (r - half) * dilation, (s - half) * dilation
            if transpose:
                dh, dw = -dh, -dw
            taps.append((dh, dw, r * ksize + s))
    return taps


def _fill_taps(desc, taps, with_wtap=True, img_add=None):
    assert 1 <= len(taps) <= MAX_TAPS
    desc.taps = len(taps)
    for i, t in enumerate(taps):
        desc.dh[i], desc.dw[i] = t[0], t[1]
        if with_wtap:
            desc.wtap[i] = t[2]
        desc.img_add[i] = img_add[i] if img_add is not None else 0
    desc.img_mul = 1


def conv_taps_s2(ksize, n):
    """Taps of a stride-2 'same' conv (k in {1,3}) on the 2x2 phase tensor [4N,Hh,Wh,C] (space_to_phases):
    [(dh, dw, wtap, img_add, (ph, pw))]: input row 2*ho + r - k//2 lives in phase (r+1)&1 at row ho + dh."""
    taps = []
    for r in range(ksize):
        for s in range(ksize):
            if ksize == 3:
                ph, dh = (0, 0) if r == 1 else (1, -1 if r == 0 else 0)
                pw, dw = (0, 0) if s == 1 else (1, -1 if s == 0 else 0)
            else:
                ph = pw = dh = dw = 0
            taps.append((dh, dw, r * ksize + s, (ph * 2 + pw) * n, (ph, pw)))
    return taps


def _planes(*ts):
    """[(plane views...)] of activations that share a storage form: one tuple for plain tensors, two for split."""
    _same_form(*ts)
    if ts[0].dim() == 5:
        return [tuple(t[0] for t in ts), tuple(t[1] for t in ts)]
    return [tuple(ts)]


def im2col3x3s2(x, cin):
    """NHWC bf16 x (first `cin` <= 3 channels real) -> patches [N, Ho, Wo, 32] of the 3x3 / stride 2 / pad 1 stem conv."""
    lib = _lib.load()
    n, h, w, _, p = _nhwc_meta(x)
    out = empty_act((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 32), is_split(x), x.device)
    for xi, oi in _planes(x, out):      # pure data movement: the hi and lo planes are gathered independently
        _lib.check(lib.semseg_im2col3x3s2(_ptr(xi), p, n, h, w, int(cin), _ptr(oi), _stream()), "semseg_im2col3x3s2")
    return out


def space_to_phases(x):
    """x [N,H,W,C] bf16 -> [4N, (H+1)//2, (W+1)//2, C] (phase-major)."""
    _require_cuda(x)
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    xp = empty_act((4 * n, (h + 1) // 2, (w + 1) // 2, c), is_split(x), x.device)
    for xi, oi in _planes(x, xp):
        _lib.check(lib.semseg_space_to_phases(_ptr(xi), p, n, h, w, c, _ptr(oi), _stream()), "semseg_space_to_phases")
    return xp


def phases_to_space(xp, n, h, w):
    lib = _lib.load()
    c = xp.shape[-1]
    assert xp.is_contiguous() and xp.shape[-4] == 4 * n
    x = empty_act((n, h, w, c), is_split(xp), xp.device)
    for xi, oi in _planes(xp, x):
        _lib.check(lib.semseg_phases_to_space(_ptr(xi), n, h, w, c, _ptr(oi), _stream()), "semseg_phases_to_space")
    return x


# ------------------------------------------------------------------------------------------------ conv
def conv_stats_rows(n, h, w, cout):
    return int(_lib.load().semseg_conv_stats_rows(n, h, w, cout))


def conv_fprop(x, w3d, cout, taps, *, out=None, epi=EPI_RAW, relu=False, scale=None, shift=None, residual=None,
               stats=False, out_f32=None, img_add=None, out_nhw=None):
    """Implicit-GEMM conv of NHWC bf16 `x` with packed weights `w3d` [n_wtaps][rows][cols].

    Returns (y, stats_partial); y is bf16 NHWC [N,H,W,cout] (or the fp32 tensor in F32 mode); stats_partial is the
    per-CTA [rows][3][cout] (sum, sum of squares, count) buffer when stats=True.
    """
    _require_cuda(x, w3d)
    lib = _lib.load()
    nin, hin, win, cin, xp = _nhwc_meta(x)
    split = is_split(x)
    n, h, w = out_nhw if out_nhw is not None else (nin, hin, win)   # output pixel grid (differs for phase tensors)
    d = ConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cout = n, h, w, cin, cout
    d.x, d.Nin, d.Hin, d.Win, d.x_pitch = x.data_ptr(), nin, hin, win, xp
    d.x_lo = _lo_int(x)
    assert w3d.dtype == torch.bfloat16 and w3d.is_contiguous() and w3d.dim() == (4 if split else 3), \
        "packed weights must be [taps][rows][cols] (plain) or [2][taps][rows][cols] (split, hi then lo slab)"
    d.w, d.n_wtaps, d.w_rows, d.w_cols = w3d.data_ptr(), w3d.shape[-3], w3d.shape[-2], w3d.shape[-1]
    d.w_split = int(split)
    _fill_taps(d, taps, img_add=img_add)
    d.epi_mode, d.relu = epi, int(bool(relu))
    sp = None
    if epi == EPI_F32:
        if out_f32 is None:
            out_f32 = torch.empty((n, h, w, cout), dtype=torch.float32, device=x.device)
        assert out_f32.dtype == torch.float32 and out_f32.stride(-1) == 1
        d.out_f32, d.out_pitch = out_f32.data_ptr(), out_f32.stride(2)
        y = out_f32
    else:
        if out is None:
            out = empty_act((n, h, w, cout), split, x.device)
        on, oh, ow, oc, op = _nhwc_meta(out)
        assert (on, oh, ow, oc) == (n, h, w, cout) and is_split(out) == split
        d.y, d.y_pitch, d.y_lo = out.data_ptr(), op, _lo_int(out)
        y = out
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.numel() >= cout
        d.scale = scale.data_ptr()
    if shift is not None:
        assert shift.dtype == torch.float32 and shift.numel() >= cout
        d.shift = shift.data_ptr()
    if residual is not None:
        rn, rh, rw, rc, rp = _nhwc_meta(residual)
        assert (rn, rh, rw, rc) == (n, h, w, cout) and is_split(residual) == split
        d.residual, d.res_pitch, d.residual_lo = residual.data_ptr(), rp, _lo_int(residual)
    # bf16x3: convs with long K are K-sliced (fp32 partials summed round-to-nearest by conv_splitk_finish), because one
    # tensor-core accumulation chain loses ~2^-24 per MMA step towards zero (tools/probe_accum.py)
    k_slices = int(lib.semseg_conv_k_slices(cin, len(taps), X3_MAX_KBLOCKS)) if (split and epi != EPI_F32) else 1
    if k_slices > 1:
        part = torch.empty((k_slices, n, h, w, cout), dtype=torch.float32, device=x.device)
        d.epi_mode, d.relu = EPI_F32, 0
        d.out_f32, d.out_pitch = part.data_ptr(), cout
        d.k_slices, d.slice_stride = k_slices, part.stride(0)
        sc, sh, rs, rsl, rpitch, ylo = d.scale, d.shift, d.residual, d.residual_lo, d.res_pitch, d.y_lo
        d.scale = d.shift = d.residual = d.residual_lo = d.y_lo = None
        _lib.check(lib.semseg_conv_fprop(ctypes.byref(d), _stream()), "semseg_conv_fprop (K-sliced)")
        m = n * h * w
        if stats:
            assert epi == EPI_RAW
            sp = torch.empty((int(lib.semseg_conv_splitk_rows(m)), 3, cout), dtype=torch.float32, device=x.device)
        _lib.check(lib.semseg_conv_splitk_finish(_ptr(part), k_slices, part.stride(0), cout, m, cout, epi,
                                                 int(bool(relu)), sc, sh, rs, rsl, rpitch, d.y, ylo, d.y_pitch,
                                                 _ptr(sp), _stream()), "semseg_conv_splitk_finish")
        return y, sp
    if stats:
        assert epi == EPI_RAW
        sp = torch.empty((conv_stats_rows(n, h, w, cout), 3, cout), dtype=torch.float32, device=x.device)
        d.stats_partial = sp.data_ptr()
    _lib.check(lib.semseg_conv_fprop(ctypes.byref(d), _stream()), "semseg_conv_fprop")
    return y, sp


def conv_wgrad(x, dy, cin, cout, taps, grad_out=None, accumulate=False, img_add=None):
    """dW (fp32 OIHW [cout][cin][k][k]) from NHWC bf16 x and dy. The pixel grid is dy's; `x` may be a phase tensor
    (stride-2 convs) addressed through `img_add`."""
    _require_cuda(x, dy)
    lib = _lib.load()
    nin, hin, win, xc, xp = _nhwc_meta(x)
    n, h, w, dc, dp = _nhwc_meta(dy)
    assert xc >= cin and dc >= cout
    assert img_add is not None or (nin, hin, win) == (n, h, w)
    _same_form(x, dy)
    d = WgradDesc()
    d.N, d.H, d.W, d.Cin, d.Cout = n, h, w, cin, cout
    d.x, d.Nin, d.Hin, d.Win, d.x_pitch = x.data_ptr(), nin, hin, win, xp
    d.dy, d.dy_pitch = dy.data_ptr(), dp
    d.x_lo, d.dy_lo = _lo_int(x), _lo_int(dy)
    _fill_taps(d, taps, with_wtap=False, img_add=img_add)
    d.n_splits = 0
    splits = lib.semseg_conv_wgrad_splits(ctypes.byref(d))
    if splits <= 0:
        raise _lib.SemsegError("semseg_conv_wgrad_splits failed (%d)" % splits)
    ntaps = len(taps)
    part = torch.empty((splits, ntaps, cout, cin), dtype=torch.float32, device=x.device)
    d.dw_partial = part.data_ptr()
    d.n_splits = splits
    _lib.check(lib.semseg_conv_wgrad(ctypes.byref(d), _stream()), "semseg_conv_wgrad")
    k = int(round(ntaps ** 0.5))
    if grad_out is None:
        grad_out = torch.empty((cout, cin, k, k), dtype=torch.float32, device=x.device)
        accumulate = False
    assert grad_out.is_contiguous() and grad_out.dtype == torch.float32
    _lib.check(lib.semseg_wgrad_reduce(_ptr(part), splits, ntaps, cout, cin, _ptr(grad_out), int(accumulate),
                                       _stream()), "semseg_wgrad_reduce")
    return grad_out


# ------------------------------------------------------------------------------------------------ layout
def nchw_to_nhwc_bf16(x, pad_to=8, split=False):
    _require_cuda(x)
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c, h, w = x.shape
    cp = round_up(c, pad_to)
    shape = ((2,) if split else ()) + (n, h, w, cp)
    out = (torch.zeros if cp != c else torch.empty)(shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.semseg_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), _lo(out), n, c, h, w, cp, _stream()),
               "semseg_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_bf16_to_nchw(x):
    _require_cuda(x)
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.semseg_nhwc_bf16_to_nchw_f32(_ptr(x), _lo(x), _ptr(out), n, c, h, w, p, _stream()),
               "semseg_nhwc_bf16_to_nchw_f32")
    return out


def nhwc_f32_to_nchw(x):
    _require_cuda(x)
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.stride(-1) == 1
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.semseg_nhwc_f32_to_nchw_f32(_ptr(x), _ptr(out), n, c, h, w, x.stride(2), _stream()),
               "semseg_nhwc_f32_to_nchw_f32")
    return out


# ------------------------------------------------------------------------------------------------ batch norm
def bn_workspace(m, c, device):
    nf = int(_lib.load().semseg_bn_workspace_floats(m, c))
    return torch.empty((nf,), dtype=torch.float32, device=device), nf


def bn_merge_partials(stats_partial):
    lib = _lib.load()
    t, _, c = stats_partial.shape
    out = torch.empty((3, c), dtype=torch.float32, device=stats_partial.device)
    _lib.check(lib.semseg_bn_merge_partials(_ptr(stats_partial), t, c, _ptr(out), _stream()),
               "semseg_bn_merge_partials")
    return out


def bn_stats(x):
    """(mean, M2, count) [3][C] of an NHWC bf16 tensor."""
    _require_cuda(x)
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    m = n * h * w
    ws, nf = bn_workspace(m, c, x.device)
    out = torch.empty((3, c), dtype=torch.float32, device=x.device)
    _lib.check(lib.semseg_bn_stats(_ptr(x), m, c, p, _ptr(ws), nf, _ptr(out), _stream()), "semseg_bn_stats")
    return out


def bn_finalize(rank_stats, gamma, beta, eps, momentum, running_mean, running_var):
    """rank_stats [R][3][C] -> (mean_invstd [2][C], scale_shift [2][C]); updates running stats in place."""
    lib = _lib.load()
    if rank_stats.dim() == 2:
        rank_stats = rank_stats.unsqueeze(0)
    r, _, c = rank_stats.shape
    assert rank_stats.is_contiguous()
    mi = torch.empty((3, c), dtype=torch.float32, device=rank_stats.device)      # mean, invstd, total count
    ss = torch.empty((2, c), dtype=torch.float32, device=rank_stats.device)
    _lib.check(lib.semseg_bn_finalize(_ptr(rank_stats), r, c, _ptr(gamma), _ptr(beta), float(eps), float(momentum),
                                      _ptr(running_mean), _ptr(running_var), _ptr(mi), _ptr(ss), _stream()),
               "semseg_bn_finalize")
    return mi, ss


def bn_finalize_partials(stats_partial, gamma, beta, eps, momentum, running_mean, running_var):
    """Per-tile conv partials -> (mean_invstd, scale_shift) in one launch (single-rank BatchNorm)."""
    lib = _lib.load()
    t, _, c = stats_partial.shape
    buf = torch.empty((5, c), dtype=torch.float32, device=stats_partial.device)
    mi, ss = buf[:3], buf[3:]                               # (mean, invstd, total count), (scale, shift)
    _lib.check(lib.semseg_bn_finalize_partials(_ptr(stats_partial), t, c, _ptr(gamma), _ptr(beta),
                                               float(eps), float(momentum), _ptr(running_mean), _ptr(running_var),
                                               _ptr(mi), _ptr(ss), _stream()), "semseg_bn_finalize_partials")
    return mi, ss


def bn_fold_eval(gamma, beta, running_mean, running_var, eps):
    lib = _lib.load()
    c = running_mean.numel()
    ss = torch.empty((2, c), dtype=torch.float32, device=running_mean.device)
    _lib.check(lib.semseg_bn_fold_eval(_ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), float(eps), c,
                                       _ptr(ss), _stream()), "semseg_bn_fold_eval")
    return ss


def bn_apply(x, scale_shift, residual=None, relu=True, out=None):
    _require_cuda(x)
    lib = _lib.load()
    n, h, w, c, xp = _nhwc_meta(x)
    if out is None:
        out = empty_act((n, h, w, c), is_split(x), x.device)
    _, _, _, oc, op = _nhwc_meta(out)
    assert oc == c
    rp = 0
    if residual is not None:
        _, _, _, rc, rp = _nhwc_meta(residual)
        assert rc == c
    _same_form(x, out, residual)
    _lib.check(lib.semseg_bn_apply(_ptr(x), _lo(x), xp, _ptr(scale_shift), _ptr(residual), _lo(residual), rp,
                                   _ptr(out), _lo(out), op, n * h * w, c, int(bool(relu)), _stream()),
               "semseg_bn_apply")
    return out


def bn_bwd_reduce(dy, y, x, mean_invstd, relu, scale_shift=None):
    lib = _lib.load()
    n, h, w, c, dp = _nhwc_meta(dy)
    _, _, _, _, xp = _nhwc_meta(x)
    yp = _nhwc_meta(y)[4] if y is not None else 0
    m = n * h * w
    ws, nf = bn_workspace(m, c, dy.device)
    sums = torch.empty((2, c), dtype=torch.float32, device=dy.device)
    _same_form(dy, y, x)
    _lib.check(lib.semseg_bn_bwd_reduce(_ptr(dy), _lo(dy), dp, _ptr(y), _lo(y), yp, _ptr(x), _lo(x), xp,
                                        _ptr(mean_invstd), _ptr(scale_shift), m, c, int(bool(relu)), _ptr(ws), nf,
                                        _ptr(sums), _stream()),
               "semseg_bn_bwd_reduce")
    return sums


def bn_bwd_apply(dy, y, x, mean_invstd, gamma, sums, count, relu, want_dres=False, scale_shift=None):
    """Returns (dx bf16, dres bf16 or None, dgamma_dbeta [2][C])."""
    lib = _lib.load()
    n, h, w, c, dp = _nhwc_meta(dy)
    _, _, _, _, xp = _nhwc_meta(x)
    yp = _nhwc_meta(y)[4] if y is not None else 0
    _same_form(dy, y, x)
    split = is_split(dy)
    dx = empty_act((n, h, w, c), split, dy.device)
    dres = empty_act((n, h, w, c), split, dy.device) if want_dres else None
    dgb = torch.empty((2, c), dtype=torch.float32, device=dy.device)
    _lib.check(lib.semseg_bn_bwd_apply(_ptr(dy), _lo(dy), dp, _ptr(y), _lo(y), yp, _ptr(x), _lo(x), xp,
                                       _ptr(mean_invstd), _ptr(gamma), _ptr(scale_shift), _ptr(sums), float(count),
                                       n * h * w, c, int(bool(relu)), _ptr(dx), _lo(dx), c, _ptr(dres), _lo(dres), c,
                                       _ptr(dgb), _stream()), "semseg_bn_bwd_apply")
    return dx, dres, dgb


def add_act(a, b, out=None):
    """a + b of two activations (either storage form; a split-aware add, unlike adding the planes)."""
    lib = _lib.load()
    n, h, w, c, ap = _nhwc_meta(a)
    bp = _nhwc_meta(b)[4]
    if out is None:
        out = empty_act((n, h, w, c), is_split(a), a.device)
    op = _nhwc_meta(out)[4]
    _same_form(a, b, out)
    _lib.check(lib.semseg_add_act(_ptr(a), _lo(a), ap, _ptr(b), _lo(b), bp, _ptr(out), _lo(out), op, n * h * w, c,
                                  _stream()), "semseg_add_act")
    return out


def scale_nc(x, scale):
    """x[n, :, :, c] * scale[n, c] (scale fp32 [N, C]): Dropout2d's per-(image, channel) factor."""
    lib = _lib.load()
    n, h, w, c, xp = _nhwc_meta(x)
    assert scale.dtype == torch.float32 and scale.is_contiguous() and scale.numel() == n * c
    out = empty_act((n, h, w, c), is_split(x), x.device)
    _lib.check(lib.semseg_scale_nc(_ptr(x), _lo(x), xp, _ptr(scale), _ptr(out), _lo(out), c, n, h * w, c, _stream()),
               "semseg_scale_nc")
    return out


def f32_to_act(x, split, pad_to=8):
    """fp32 NHWC [N,H,W,C] (channel-contiguous, any pixel pitch) -> activation [N,H,W,Cp], Cp = C rounded up to
    `pad_to`, padding zero filled."""
    _require_cuda(x)
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.stride(-1) == 1
    n, h, w, c = x.shape
    sn, sh, sw, _ = x.stride()
    assert sh == sw * w and sn == sh * h
    cp = round_up(c, pad_to)
    out = empty_act((n, h, w, cp), split, x.device)
    _lib.check(lib.semseg_f32_to_act(_ptr(x), sw, _ptr(out), _lo(out), cp, n * h * w, c, cp, _stream()),
               "semseg_f32_to_act")
    return out


def act_to_f32(x):
    """activation -> fp32 NHWC [N,H,W,C]."""
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    _lib.check(lib.semseg_act_to_f32(_ptr(x), _lo(x), p, _ptr(out), c, n * h * w, c, _stream()), "semseg_act_to_f32")
    return out


# ------------------------------------------------------------------------------------------------ fused tail
def upsample_ce_fwd(logits, target, ignore_index, want_argmax=True):
    """logits fp32 NHWC [N,h,w,C], target int64 [N,Ho,Wo] -> (loss_info [2] = (mean CE, count), argmax, lse)."""
    _require_cuda(logits, target)
    lib = _lib.load()
    assert logits.dtype == torch.float32 and logits.dim() == 4 and logits.stride(-1) == 1
    assert target.dtype == torch.int64 and target.is_contiguous()
    n, h, w, c = logits.shape
    _, ho, wo = target.shape
    nws = int(lib.semseg_upsample_ce_workspace_floats(n, ho, wo))
    ws = torch.empty((nws,), dtype=torch.float32, device=logits.device)
    info = torch.empty((2,), dtype=torch.float32, device=logits.device)
    amax = torch.empty((n, ho, wo), dtype=torch.int64, device=logits.device) if want_argmax else None
    lse = torch.empty((n, ho, wo), dtype=torch.float32, device=logits.device)
    _lib.check(lib.semseg_upsample_ce_fwd(_ptr(logits), logits.stride(2), n, h, w, c, _ptr(target), ho, wo,
                                          int(ignore_index), _ptr(ws), _ptr(info), _ptr(amax), _ptr(lse), _stream()),
               "semseg_upsample_ce_fwd")
    return info, amax, lse


def upsample_ce_bwd(logits, target, ignore_index, lse, info, grad_out):
    lib = _lib.load()
    n, h, w, c = logits.shape
    _, ho, wo = target.shape
    dl = torch.empty((n, h, w, c), dtype=torch.float32, device=logits.device)
    ws = torch.empty((int(lib.semseg_upsample_ce_bwd_workspace_floats(n, ho, w, c)),), dtype=torch.float32,
                     device=logits.device)
    g = grad_out.reshape(1).float().contiguous()
    _lib.check(lib.semseg_upsample_ce_bwd(_ptr(logits), logits.stride(2), n, h, w, c, _ptr(target), ho, wo,
                                          int(ignore_index), _ptr(lse), _ptr(info), _ptr(g), _ptr(ws), _ptr(dl),
                                          _stream()),
               "semseg_upsample_ce_bwd")
    return dl


# ------------------------------------------------------------------------------------------------ pyramid pooling
def _bin_args(bins, tensors):
    """(bins[], hi pointers[], lo pointers[] or NULL, nb)"""
    nb = len(bins)
    barr = (ctypes.c_int * nb)(*bins)
    parr = (ctypes.c_void_p * nb)(*[t.data_ptr() for t in tensors])
    larr = (ctypes.c_void_p * nb)(*[_lo_int(t) for t in tensors]) if is_split(tensors[0]) else None
    return barr, parr, larr, nb


def ppm_pool(x, bins):
    """x NHWC bf16 -> [pooled_k [N,b,b,C] bf16 for b in bins] (AdaptiveAvgPool2d of every bin, one launch)."""
    _require_cuda(x)
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    outs = [empty_act((n, b, b, c), is_split(x), x.device) for b in bins]
    barr, parr, larr, nb = _bin_args(bins, outs)
    _lib.check(lib.semseg_ppm_pool(_ptr(x), _lo(x), p, n, h, w, c, barr, parr, larr, nb, _stream()), "semseg_ppm_pool")
    return outs


def ppm_pool_bwd(dpooled, bins, n, h, w, c, add=None):
    """dx of ppm_pool; `add` (NHWC bf16, possibly a channel slice of a wider tensor) is summed in."""
    lib = _lib.load()
    dpooled = [d.contiguous() for d in dpooled]
    dx = empty_act((n, h, w, c), is_split(dpooled[0]), dpooled[0].device)
    barr, parr, larr, nb = _bin_args(bins, dpooled)
    ap = _nhwc_meta(add)[4] if add is not None else 0
    _same_form(dx, add)
    _lib.check(lib.semseg_ppm_pool_bwd(parr, larr, barr, nb, n, h, w, c, _ptr(dx), _lo(dx), c, _ptr(add), _lo(add), ap,
                                       _stream()), "semseg_ppm_pool_bwd")
    return dx


def ppm_upsample_concat(x, feats, bins):
    """-> out [N,H,W,C + nb*Cr] bf16 = cat([x, bilinear(feats_k)...], channel)."""
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    feats = [f.contiguous() for f in feats]
    cr = feats[0].shape[-1]
    out = empty_act((n, h, w, c + len(bins) * cr), is_split(x), x.device)
    barr, parr, larr, nb = _bin_args(bins, feats)
    _same_form(x, feats[0])
    _lib.check(lib.semseg_ppm_upsample_concat(_ptr(x), _lo(x), p, parr, larr, barr, nb, n, h, w, c, cr, _ptr(out),
                                              _lo(out), out.shape[-1], _stream()), "semseg_ppm_upsample_concat")
    return out


def ppm_upsample_bwd(dout, c_off, bins, cr):
    lib = _lib.load()
    n, h, w, ct, p = _nhwc_meta(dout)
    dfeats = [empty_act((n, b, b, cr), is_split(dout), dout.device) for b in bins]
    barr, parr, larr, nb = _bin_args(bins, dfeats)
    _lib.check(lib.semseg_ppm_upsample_bwd(_ptr(dout), _lo(dout), p, c_off, parr, larr, barr, nb, n, h, w, cr,
                                           _stream()), "semseg_ppm_upsample_bwd")
    return dfeats


# ------------------------------------------------------------------------------------------------ bilinear resize
def resize_bilinear(x, size):
    """NHWC activation [N,Hi,Wi,C] -> [N,Ho,Wo,C], bilinear with align_corners=True."""
    _require_cuda(x)
    lib = _lib.load()
    n, hi, wi, c, p = _nhwc_meta(x)
    ho, wo = int(size[0]), int(size[1])
    y = empty_act((n, ho, wo, c), is_split(x), x.device)
    _lib.check(lib.semseg_resize_bilinear_fwd(_ptr(x), _lo(x), p, n, hi, wi, c, ho, wo, _ptr(y), _lo(y), c, _stream()),
               "semseg_resize_bilinear_fwd")
    return y


def resize_bilinear_bwd(dy, in_size):
    """Adjoint of resize_bilinear: dy [N,Ho,Wo,C] -> dx [N,Hi,Wi,C]."""
    lib = _lib.load()
    n, ho, wo, c, p = _nhwc_meta(dy)
    hi, wi = int(in_size[0]), int(in_size[1])
    dx = empty_act((n, hi, wi, c), is_split(dy), dy.device)
    _lib.check(lib.semseg_resize_bilinear_bwd(_ptr(dy), _lo(dy), p, n, hi, wi, c, ho, wo, _ptr(dx), _lo(dx), c, _stream()),
               "semseg_resize_bilinear_bwd")
    return dx


# ------------------------------------------------------------------------------------------------ max pool
def maxpool3x3s2_fwd(x, want_argcode=True):
    """-> (y, argcode uint8 or None)."""
    _require_cuda(x)
    lib = _lib.load()
    n, h, w, c, p = _nhwc_meta(x)
    assert p == c, "maxpool expects a dense NHWC tensor"
    shape = (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c)
    y = empty_act(shape, is_split(x), x.device)
    code = torch.empty(shape, dtype=torch.uint8, device=x.device) if want_argcode else None
    _lib.check(lib.semseg_maxpool3x3s2_fwd(_ptr(x), _lo(x), _ptr(y), _lo(y), _ptr(code), n, h, w, c, _stream()),
               "semseg_maxpool3x3s2_fwd")
    return y, code


def maxpool3x3s2_bwd(argcode, dy, in_shape):
    lib = _lib.load()
    n, h, w, c = in_shape
    dy = dy.contiguous()
    dx = empty_act((n, h, w, c), is_split(dy), dy.device)
    _lib.check(lib.semseg_maxpool3x3s2_bwd(_ptr(argcode), _ptr(dy), _lo(dy), _ptr(dx), _lo(dx), n, h, w, c,
                                           _stream()), "semseg_maxpool3x3s2_bwd")
    return dx


# ------------------------------------------------------------------------------------------------ SyncBN over NVLink
def bn_finalize_p2p(stats_partial, gamma, beta, eps, momentum, running_mean, running_var, px):
    """Like bn_finalize_partials, with the cross-rank exchange done inside the kernel over peer memory (px)."""
    lib = _lib.load()
    t, _, c = stats_partial.shape
    buf = torch.empty((5, c), dtype=torch.float32, device=stats_partial.device)
    mi, ss = buf[:3], buf[3:]
    slot, seq = px.next()
    _lib.check(lib.semseg_bn_finalize_p2p(_ptr(stats_partial), t, c, _ptr(gamma), _ptr(beta), float(eps),
                                          float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(mi), _ptr(ss),
                                          px.data_ptrs, px.flag_ptrs, _ptr(px.counter), px.world, px.rank, slot,
                                          int(__import__("semseg_b200.p2p", fromlist=["x"]).SLOT_FLOATS), seq,
                                          _ptr(px.step), _stream()), "semseg_bn_finalize_p2p")
    return mi, ss


def bn_bwd_reduce_p2p(dy, y, x, mean_invstd, relu, scale_shift, px):
    """-> (sums_local [2][C], sums_total [2][C]) with the all-reduce done over peer memory."""
    lib = _lib.load()
    n, h, w, c, dp = _nhwc_meta(dy)
    _, _, _, _, xp = _nhwc_meta(x)
    yp = _nhwc_meta(y)[4] if y is not None else 0
    m = n * h * w
    ws, nf = bn_workspace(m, c, dy.device)
    out = torch.empty((2, 2, c), dtype=torch.float32, device=dy.device)
    slot, seq = px.next()
    _same_form(dy, y, x)
    _lib.check(lib.semseg_bn_bwd_reduce_p2p(_ptr(dy), _lo(dy), dp, _ptr(y), _lo(y), yp, _ptr(x), _lo(x), xp,
                                            _ptr(mean_invstd),
                                            _ptr(scale_shift), m, c, int(bool(relu)), _ptr(ws), nf, _ptr(out[0]),
                                            _ptr(out[1]), px.data_ptrs, px.flag_ptrs, _ptr(px.counter), px.world,
                                            px.rank, slot,
                                            int(__import__("semseg_b200.p2p", fromlist=["x"]).SLOT_FLOATS), seq,
                                            _ptr(px.step), _stream()), "semseg_bn_bwd_reduce_p2p")
    return out[0], out[1]
