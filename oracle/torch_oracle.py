"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (never imported by semseg_b200/).

fp32 functional restatement of the reference's PSPNet / PSANet forward (training and eval) on top of
torch.nn.functional. The reference delegates all arithmetic on this path to PyTorch (ATen / oneDNN / cuDNN,
not vendored in /root/reference), so the restatement calls the same ATen operators in the same order; what it
restates is the *network*: which operator runs on which tensor with which hyper-parameters, keyed by the
reference's state_dict names. Runs on CPU (the cpu_baseline / `--impl reference` legs of bench.py) or on a GPU
in strict fp32 (TF32 off) as the large-size oracle of the parity tests.

Parity pin: tests/golden/make_golden.py imports the real reference modules in the build container, runs them on
seeded inputs and commits logits / losses / gradient norms; tests/test_oracle_cpu.py replays this file against
those fixtures (rel. L2 <= 1e-5 — the reference's own 1-vs-8-thread reorder noise is 8e-7 .. 2e-5).

Reference lines restated:
  stem / maxpool      model/resnet.py:106-115,148-152 (wired as layer0 at model/pspnet.py:46)
  Bottleneck          model/resnet.py:74-94
  layer strides/dil.  model/resnet.py:116-119 + dilation patch model/pspnet.py:49-58 (same model/psanet.py:123-132)
  PPM                 model/pspnet.py:21-26
  PSA                 model/psanet.py:53-98  (psa_mask = lib/psa/functions/psamask.py, restated in psamask_oracle.c)
  heads / tail        model/pspnet.py:64-78,93-105
"""
import torch
import torch.nn.functional as F

BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}
EPS, MOMENTUM = 1e-5, 0.1


def psa_mask_torch(x, psa_type, mask_h, mask_w):
    """Differentiable torch restatement of psa_mask (index gather), used so the oracle has gradients.
    Checked against psamask_oracle.c / the reference extension in tests/test_oracle_cpu.py."""
    n, c, h, w = x.shape
    hh, hw = (mask_h - 1) // 2, (mask_w - 1) // 2
    dev = x.device
    i = torch.arange(h, device=dev).view(h, 1, 1, 1)
    j = torch.arange(w, device=dev).view(1, w, 1, 1)
    hs = torch.arange(h, device=dev).view(1, 1, h, 1)
    ws = torch.arange(w, device=dev).view(1, 1, 1, w)
    a = i - hs + hh                      # mask row that maps (h) -> target row i
    b = j - ws + hw
    valid = ((a >= 0) & (a < mask_h) & (b >= 0) & (b < mask_w))
    ch = (a.clamp(0, mask_h - 1) * mask_w + b.clamp(0, mask_w - 1)).expand(h, w, h, w)
    ch = ch.reshape(1, h * w, h, w).expand(n, -1, -1, -1)
    col = torch.gather(x, 1, ch) * valid.expand(h, w, h, w).reshape(1, h * w, h, w).to(x.dtype)
    if psa_type == 0:
        return col
    return col.view(n, h * w, h * w).transpose(1, 2).reshape(n, h * w, h, w)


class Oracle:
    """Functional model over a reference-format state_dict (tensors are used in place; BN buffers are updated in
    training mode exactly like nn.BatchNorm2d)."""

    def __init__(self, sd, arch='psp', layers=50, bins=(1, 2, 3, 6), classes=2, zoom_factor=8, psa_type=2,
                 compact=False, shrink_factor=2, mask_h=59, mask_w=59, normalization_factor=1.0, psa_softmax=True,
                 ignore_index=255, dropout=0.0):
        self.sd = {k[7:] if k.startswith('module.') else k: v for k, v in sd.items()}
        self.arch, self.layers, self.bins, self.classes, self.zoom = arch, layers, bins, classes, zoom_factor
        self.psa_type, self.compact, self.shrink = psa_type, compact, shrink_factor
        self.mask_h, self.mask_w = mask_h, mask_w
        self.norm = mask_h * mask_w if normalization_factor is None else normalization_factor
        self.psa_softmax, self.ignore_index, self.dropout = psa_softmax, ignore_index, dropout
        self.training = True

    # -------------------------------------------------------------------------------------------- primitives
    def conv(self, x, name, stride=1, padding=0, dilation=1):
        return F.conv2d(x, self.sd[name + '.weight'], self.sd.get(name + '.bias'), stride, padding, dilation)

    def bn(self, x, name):
        sd = self.sd
        if self.training:
            nbt = sd.get(name + '.num_batches_tracked')
            if nbt is not None:
                nbt += 1
        return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'],
                            sd[name + '.bias'], self.training, MOMENTUM, EPS)

    def cbr(self, x, conv, bn, relu=True, **kw):
        y = self.bn(self.conv(x, conv, **kw), bn)
        return F.relu(y) if relu else y

    # -------------------------------------------------------------------------------------------- backbone
    def bottleneck(self, x, p, stride, dilation, has_down):
        y = self.cbr(x, p + '.conv1', p + '.bn1')
        y = self.cbr(y, p + '.conv2', p + '.bn2', stride=stride, padding=dilation, dilation=dilation)
        y = self.bn(self.conv(y, p + '.conv3'), p + '.bn3')
        res = self.bn(self.conv(x, p + '.downsample.0', stride=stride), p + '.downsample.1') if has_down else x
        return F.relu(y + res)

    def backbone(self, x):
        x = self.cbr(x, 'layer0.0', 'layer0.1', stride=2, padding=1)
        x = self.cbr(x, 'layer0.3', 'layer0.4', padding=1)
        x = self.cbr(x, 'layer0.6', 'layer0.7', padding=1)
        x = F.max_pool2d(x, 3, 2, 1)
        feats = {}
        # (stride of the first block's 3x3 and downsample, dilation of every 3x3)
        cfg = {1: (1, 1), 2: (2, 1), 3: (1, 2), 4: (1, 4)}
        for li, nblocks in enumerate(BLOCKS[self.layers], start=1):
            stride, dil = cfg[li]
            for b in range(nblocks):
                x = self.bottleneck(x, 'layer%d.%d' % (li, b), stride if b == 0 else 1, dil, b == 0)
            feats[li] = x
        return feats[3], feats[4]

    # -------------------------------------------------------------------------------------------- PPM / PSA
    def ppm(self, x):
        out = [x]
        for i, b in enumerate(self.bins):
            p = 'ppm.features.%d' % i
            y = self.cbr(F.adaptive_avg_pool2d(x, b), p + '.1', p + '.2')
            out.append(F.interpolate(y, x.shape[2:], mode='bilinear', align_corners=True))
        return torch.cat(out, 1)

    def _psa_branch(self, x, red, att, mask_type):
        t = self.cbr(x, red + '.0', red + '.1')
        n, c, h, w = t.shape
        if self.shrink != 1:
            h, w = (h - 1) // self.shrink + 1, (w - 1) // self.shrink + 1
            t = F.interpolate(t, size=(h, w), mode='bilinear', align_corners=True)
        y = self.conv(self.cbr(t, att + '.0', att + '.1'), att + '.3')
        if self.compact:
            if mask_type == 1:
                y = y.view(n, h * w, h * w).transpose(1, 2).reshape(n, h * w, h, w)
        else:
            y = psa_mask_torch(y, mask_type, self.mask_h, self.mask_w)
        if self.psa_softmax:
            y = F.softmax(y, dim=1)
        t = torch.bmm(t.view(n, c, h * w), y.view(n, h * w, h * w)).view(n, c, h, w) * (1.0 / self.norm)
        return t, (h, w)

    def psa(self, x):
        if self.psa_type in (0, 1):
            t, (h, w) = self._psa_branch(x, 'psa.reduce', 'psa.attention', self.psa_type)
        else:
            tc, (h, w) = self._psa_branch(x, 'psa.reduce', 'psa.attention', 0)
            td, _ = self._psa_branch(x, 'psa.reduce_p', 'psa.attention_p', 1)
            t = torch.cat([tc, td], 1)
        t = self.cbr(t, 'psa.proj.0', 'psa.proj.1')
        if self.shrink != 1:
            h, w = (h - 1) * self.shrink + 1, (w - 1) * self.shrink + 1
            t = F.interpolate(t, size=(h, w), mode='bilinear', align_corners=True)
        return torch.cat((x, t), 1)

    # -------------------------------------------------------------------------------------------- heads / tail
    def head(self, x, name):
        y = self.cbr(x, name + '.0', name + '.1', padding=1)
        if self.training and self.dropout > 0:
            y = F.dropout2d(y, self.dropout, True)
        return self.conv(y, name + '.4')

    def logits_lowres(self, x):
        """(main, aux) logits before the final upsample; aux only in training."""
        f3, f4 = self.backbone(x)
        f = self.ppm(f4) if self.arch == 'psp' else self.psa(f4)
        main = self.head(f, 'cls')
        aux = self.head(f3, 'aux') if self.training else None
        return main, aux

    def forward(self, x, y=None):
        hh = int((x.shape[2] - 1) / 8 * self.zoom + 1)
        ww = int((x.shape[3] - 1) / 8 * self.zoom + 1)
        main, aux = self.logits_lowres(x)
        if self.zoom != 1:
            main = F.interpolate(main, size=(hh, ww), mode='bilinear', align_corners=True)
        if not self.training:
            return main
        if self.zoom != 1:
            aux = F.interpolate(aux, size=(hh, ww), mode='bilinear', align_corners=True)
        main_loss = F.cross_entropy(main, y, ignore_index=self.ignore_index)
        aux_loss = F.cross_entropy(aux, y, ignore_index=self.ignore_index)
        return main.max(1)[1], main_loss, aux_loss

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)


def merge_moments(blocks):
    """Chan merge of [(mean, M2, count)] blocks (numpy or torch arrays) — the formula bn_finalize implements,
    restated for the host-side / gloo tests."""
    mean, m2, n = blocks[0]
    for (mb, m2b, nb) in blocks[1:]:
        tot = n + nb
        d = mb - mean
        mean = mean + d * (nb / tot)
        m2 = m2 + m2b + d * d * (n * nb / tot)
        n = tot
    return mean, m2, n
