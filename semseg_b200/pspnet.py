"""PSPNet, drop-in for the reference's model/pspnet.py (same constructor / forward signatures, same child
module names and state_dict keys, same return values — model/pspnet.py:29-105), executed on NHWC bf16
activations by the sm_100a kernels behind semseg_b200/functional.py.
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import functional as SF
from . import graphs
from . import ops
from . import p2p
from . import resnet as models


class PPM(nn.Module):
    """Pyramid pooling module (model/pspnet.py:8-26): per bin AdaptiveAvgPool2d -> 1x1 conv -> BN -> ReLU ->
    bilinear upsample (align_corners=True) -> concat with the input."""

    def __init__(self, in_dim, reduction_dim, bins):
        super(PPM, self).__init__()
        self.features = []
        for bin in bins:
            self.features.append(nn.Sequential(
                nn.AdaptiveAvgPool2d(bin),
                nn.Conv2d(in_dim, reduction_dim, kernel_size=1, bias=False),
                nn.BatchNorm2d(reduction_dim),
                nn.ReLU(inplace=True)
            ))
        self.features = nn.ModuleList(self.features)

    def forward_nhwc(self, x):
        bins = []
        for f in self.features:
            b = f[0].output_size
            bins.append(b if isinstance(b, int) else b[0])
        link = SF.ppm_link()                                     # the two gradients of x are summed in the pool-bwd kernel
        pooled = SF.ppm_pool(x, bins, link)                      # one launch: every bin's AdaptiveAvgPool2d
        feats = [SF.conv_bn_act(p, f[1], f[2], relu=True) for p, f in zip(pooled, self.features)]
        return SF.ppm_upsample_concat(x, feats, bins, link)      # one launch: upsample x4 + concat, written in place

    def forward(self, x):
        return SF.to_nchw_f32(self.forward_nhwc(SF.to_nhwc_bf16(x)))


def head_forward_nhwc(head, t):
    """cls / aux head: 3x3 conv + BN + ReLU + Dropout2d + 1x1 conv with bias -> fp32 NHWC logits."""
    t = SF.conv_bn_act(t, head[0], head[1], relu=True)
    t = SF.dropout2d_nhwc(t, head[3].p, head[3].training)
    return SF.conv_bias_f32(t, head[4])


def upsample_logits(logits_nhwc, size, zoom_factor):
    """fp32 NHWC logits -> NCHW (view) bilinearly upsampled to `size` (model/pspnet.py:94-95)."""
    x = logits_nhwc.permute(0, 3, 1, 2)
    if zoom_factor != 1:
        x = F.interpolate(x, size=size, mode='bilinear', align_corners=True)
    return x


class PSPNet(nn.Module):
    def __init__(self, layers=50, bins=(1, 2, 3, 6), dropout=0.1, classes=2, zoom_factor=8, use_ppm=True,
                 criterion=nn.CrossEntropyLoss(ignore_index=255), pretrained=True):
        super(PSPNet, self).__init__()
        assert layers in [50, 101, 152]
        assert 2048 % len(bins) == 0
        assert classes > 1
        assert zoom_factor in [1, 2, 4, 8]
        self.zoom_factor = zoom_factor
        self.use_ppm = use_ppm
        self.criterion = criterion

        if layers == 50:
            resnet = models.resnet50(pretrained=pretrained)
        elif layers == 101:
            resnet = models.resnet101(pretrained=pretrained)
        else:
            resnet = models.resnet152(pretrained=pretrained)
        self.layer0 = resnet.stem()
        self.layer1, self.layer2, self.layer3, self.layer4 = resnet.layer1, resnet.layer2, resnet.layer3, resnet.layer4

        # output stride 8: dilate layer3 / layer4 instead of striding (model/pspnet.py:49-58)
        for n, m in self.layer3.named_modules():
            if 'conv2' in n:
                m.dilation, m.padding, m.stride = (2, 2), (2, 2), (1, 1)
            elif 'downsample.0' in n:
                m.stride = (1, 1)
        for n, m in self.layer4.named_modules():
            if 'conv2' in n:
                m.dilation, m.padding, m.stride = (4, 4), (4, 4), (1, 1)
            elif 'downsample.0' in n:
                m.stride = (1, 1)

        fea_dim = 2048
        if use_ppm:
            self.ppm = PPM(fea_dim, int(fea_dim / len(bins)), bins)
            fea_dim *= 2
        self.cls = nn.Sequential(
            nn.Conv2d(fea_dim, 512, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(512),
            nn.ReLU(inplace=True),
            nn.Dropout2d(p=dropout),
            nn.Conv2d(512, classes, kernel_size=1)
        )
        if self.training:
            self.aux = nn.Sequential(
                nn.Conv2d(1024, 256, kernel_size=3, padding=1, bias=False),
                nn.BatchNorm2d(256),
                nn.ReLU(inplace=True),
                nn.Dropout2d(p=dropout),
                nn.Conv2d(256, classes, kernel_size=1)
            )

    _sb_head_modules = ("layer0", "layer1", "layer2")     # modules whose parameters lie before graphs.note_boundary

    def forward(self, x, y=None):
        x_size = x.size()
        assert (x_size[2] - 1) % 8 == 0 and (x_size[3] - 1) % 8 == 0
        if (self.training and torch.is_grad_enabled() and y is not None and
                SF.fused_tail_supported(self.criterion, None, y, self.zoom_factor, x_size)):
            # whole training step (forward and, later, backward) as two replayed CUDA graphs behind one autograd node
            out = graphs.train_step(self, self._forward_impl, x, y)
            if out is not None:
                return out
        return self._forward_impl(x, y)

    def _forward_impl(self, x, y=None):
        x_size = x.size()
        h = int((x_size[2] - 1) / 8 * self.zoom_factor + 1)
        w = int((x_size[3] - 1) / 8 * self.zoom_factor + 1)

        if self.training and torch.is_grad_enabled():
            SF.prepack(self, force=graphs.capturing())   # all conv operand slabs refreshed in one launch
            p2p.begin_step(force=graphs.capturing())     # new SyncBN exchange epoch (device-resident step counter)
        t = SF.to_nhwc_bf16(x)
        t = self.layer0.forward_nhwc(t)
        t = self.layer1.forward_nhwc(t)
        t = graphs.note_boundary(self.layer2.forward_nhwc(t))     # where a captured backward is cut in two
        t_tmp = self.layer3.forward_nhwc(t)
        t_aux = None
        if self.training:       # layer3's output feeds layer4 and the aux head: explicit fan-out (native gradient add)
            t_tmp, t_aux = SF.fork(t_tmp, 2)
        t = self.layer4.forward_nhwc(t_tmp)
        if self.use_ppm:
            t = self.ppm.forward_nhwc(t)
        logits = head_forward_nhwc(self.cls, t)

        if self.training:
            aux_logits = head_forward_nhwc(self.aux, t_aux)
            if SF.fused_tail_supported(self.criterion, logits, y, self.zoom_factor):
                # upsample + cross-entropy + argmax fused: [N, classes, H, W] never exists (model/pspnet.py:94-103)
                main_loss, pred = SF.upsample_ce(logits, y, self.criterion.ignore_index)
                aux_loss, _ = SF.upsample_ce(aux_logits, y, self.criterion.ignore_index)
                return pred, main_loss, aux_loss
            x = upsample_logits(logits, (h, w), self.zoom_factor)
            aux = upsample_logits(aux_logits, (h, w), self.zoom_factor)
            main_loss = self.criterion(x, y)
            aux_loss = self.criterion(aux, y)
            return x.max(1)[1], main_loss, aux_loss
        else:
            x = ops.nhwc_f32_to_nchw(logits) if not logits.requires_grad else logits.permute(0, 3, 1, 2).contiguous()
            if self.zoom_factor != 1:
                x = F.interpolate(x, size=(h, w), mode='bilinear', align_corners=True)
            return x
