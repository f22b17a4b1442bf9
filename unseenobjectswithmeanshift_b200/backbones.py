"""Backbones that FEED the hot path (out of scope as kernels - plain cuDNN work, they stay
PyTorch). They exist so that the benchmarks and the META_ARCH wrappers can run at the reference's real API boundary
- images in, instances out (``PretrainedMeanShiftMaskFormer.forward(batched_inputs)``,
pretrained_meanshiftformer_model.py:244-378) - instead of stopping at backbone features.

``ResNet50Features``   config #2 (configs/tabletop_pretrained_ResNet50.yaml: ``build_resnet_backbone``, DEPTH 50,
                       STRIDE_IN_1X1 False, torchvision-format weights, FrozenBN): that IS torchvision's ResNet-50 in
                       eval mode, so torchvision's module is used (third-party code on both sides, like detectron2's
                       builder in the reference). Returns {"res2".."res5"} with strides 4/8/16/32, 256..2048 channels.
``SegnetEmbedding``    configs #1 / #3 (lib/networks/SEG.py:88-120 ``SEGNET`` with INPUT RGBD, FUSION_TYPE add:
                       two ResNet34-8s streams, RGB and XYZ, summed; lib/networks/resnet_dilated.py:287-327: dilated
                       ResNet-34 at output stride 8, 1x1 classifier to 64 channels, bilinear upsampling
                       (``upsample_bilinear`` = align_corners True) to the input size). Restated with torchvision's
                       BasicBlock ResNet-34 and ``replace_stride_with_dilation``-style surgery done by hand (BasicBlock
                       refuses dilation in torchvision). Random init; no checkpoint exists in the container.
cuDNN math: PyTorch's own default for convolutions on Ampere and later is TF32 (``torch.backends.cudnn.allow_tf32``),
which is therefore what the reference's unmodified code does with its backbone on this GPU - and the default here
(channels_last). ``set_tf32(False)`` before construction selects strict fp32 (NCHW: measured on B200 at batch 8, 640x480,
ResNet-50: TF32 channels_last 4.65 ms, fp32 NCHW 16.5 ms, fp32 channels_last 26.5 ms - cuDNN's fp32 CUDA-core kernels are
not a tuned path on this part). The parity claims of this package concern the head and are stated on identical backbone
features (the head itself never uses TF32: precision.py, split-precision tensor-core kernels).
"""
import contextlib
import os

import torch
import torch.nn.functional as F
from torch import nn

_TF32 = True


def set_tf32(flag):
    """cuDNN math of backbones constructed AFTER this call: True (default) = PyTorch's own default on Ampere and later
    (TF32 tensor cores, channels_last), i.e. what the reference's stock code does on a GPU; False = strict fp32 (NCHW),
    the reference's CPU arithmetic."""
    global _TF32
    _TF32 = bool(flag)


@contextlib.contextmanager
def _conv_math(tf32):
    """cuDNN settings of the backbones' convolutions: TF32 on / off, and (MSM_CUDNN_BENCHMARK=1) cuDNN's own autotuner
    instead of its heuristics - the choices are made in the eager warm-up calls and reused under graph capture."""
    old, old_b = torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    if os.environ.get("MSM_CUDNN_BENCHMARK", "0") == "1":
        torch.backends.cudnn.benchmark = True
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32 = old
        torch.backends.cudnn.benchmark = old_b


def _fold_bn(conv, bn):
    """eval-mode BatchNorm folded into the preceding convolution (same arithmetic up to fp32 rounding)."""
    w = conv.weight * (bn.weight / torch.sqrt(bn.running_var + bn.eps)).reshape(-1, 1, 1, 1)
    b = bn.bias - bn.running_mean * bn.weight / torch.sqrt(bn.running_var + bn.eps)
    if conv.bias is not None:
        b = b + conv.bias * bn.weight / torch.sqrt(bn.running_var + bn.eps)
    out = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                    conv.groups, bias=True)
    out.weight.data.copy_(w)
    out.bias.data.copy_(b)
    return out


class ResNet50Features(nn.Module):
    """torchvision ResNet-50 trunk -> {"res2","res3","res4","res5"} (detectron2 naming). ``fold_bn``: inference-time
    folding of the frozen BatchNorms into the convolutions (halves the elementwise launches; off = the plain module)."""

    size_divisibility = 32

    def __init__(self, seed=0, fold_bn=True, fused_relu=True):
        super().__init__()
        import torchvision
        g = torch.random.get_rng_state()
        torch.manual_seed(5000 + seed)
        net = torchvision.models.resnet50(weights=None)
        torch.random.set_rng_state(g)
        net.eval()
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.res2, self.res3, self.res4, self.res5 = net.layer1, net.layer2, net.layer3, net.layer4
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()
        self.tf32 = _TF32
        self.fmt = torch.channels_last if _TF32 else torch.contiguous_format
        self.fused_relu = False
        if fold_bn:
            self.fold_(fused_relu)
        self.to(memory_format=self.fmt)   # weights converted ONCE (else cuDNN re-lays them out per call)

    def fold_(self, fused_relu=True):
        """Inference-time folding of the frozen BatchNorms into the convolutions, in place; ``fused_relu``: forward then
        uses cuDNN's conv + bias + (residual) + ReLU kernels."""
        with torch.no_grad():
            self.stem[0], self.stem[1] = _fold_bn(self.stem[0], self.stem[1]), nn.Identity()
            for layer in (self.res2, self.res3, self.res4, self.res5):
                for blk in layer:
                    blk.conv1, blk.bn1 = _fold_bn(blk.conv1, blk.bn1), nn.Identity()
                    blk.conv2, blk.bn2 = _fold_bn(blk.conv2, blk.bn2), nn.Identity()
                    blk.conv3, blk.bn3 = _fold_bn(blk.conv3, blk.bn3), nn.Identity()
                    if blk.downsample is not None:
                        blk.downsample = nn.Sequential(_fold_bn(blk.downsample[0], blk.downsample[1]))
                        # both biases are added before the block's ReLU: one of them is enough (PyTorch adds a
                        # Conv2d bias with a separate elementwise kernel - 101 us on the 157 MB res2 map)
                        blk.conv3.bias += blk.downsample[0].bias
                        blk.downsample[0].bias = None
        for p in self.parameters():
            p.requires_grad_(False)
        self.fused_relu = bool(fused_relu)   # cuDNN conv + bias + (residual) + ReLU kernels at inference
        self.to(memory_format=self.fmt)
        return self

    def train(self, mode=True):  # FrozenBN semantics: never leaves eval mode
        return super().train(False)

    @staticmethod
    def _conv_relu(conv, x):
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)

    def _bottleneck_fused(self, blk, x):
        """torchvision Bottleneck.forward with cuDNN's fused conv + bias + ReLU and conv + bias + residual add + ReLU
        (the BatchNorms are folded): three launches per block instead of eight."""
        y = self._conv_relu(blk.conv1, x)
        y = self._conv_relu(blk.conv2, y)
        idt = x if blk.downsample is None else blk.downsample(x)
        c3 = blk.conv3
        return torch.cudnn_convolution_add_relu(y, c3.weight, idt, 1.0, c3.bias, c3.stride, c3.padding, c3.dilation,
                                                c3.groups)

    def forward(self, x):
        x = x.contiguous(memory_format=self.fmt)
        fused = self.fused_relu and x.is_cuda and not torch.is_grad_enabled()
        with _conv_math(self.tf32):
            if fused:
                x = self._conv_relu(self.stem[0], x)
                if self.fmt == torch.channels_last:
                    from . import ops
                    x = ops.maxpool3x3s2_channels_last(x)   # ATen's NHWC max-pool: 187 us at [8,64,240,320]
                else:
                    x = self.stem[3](x)
            else:
                x = self.stem(x)
            out = {}
            for name in ("res2", "res3", "res4", "res5"):
                if fused:
                    for blk in getattr(self, name):
                        x = self._bottleneck_fused(blk, x)
                else:
                    x = getattr(self, name)(x)
                # channels_last maps go to the head as they are: its 1x1 convolutions read them as token-major rows
                # (ops.conv1x1_nhwc) - no NCHW copy (0.32 ms per step at batch 8)
                out[name] = x if self.fmt == torch.channels_last and x.is_cuda else x.contiguous()
        return out


class _BasicBlock(nn.Module):
    """ResNet BasicBlock with dilation (lib/networks/resnet.py's block; torchvision's refuses dilation > 1)."""

    def __init__(self, cin, cout, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, dilation, dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, dilation, dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = downsample

    def fold(self):
        """eval-mode BatchNorms folded into the convolutions (inference): enables cuDNN's fused conv + bias + ReLU."""
        with torch.no_grad():
            self.conv1, self.bn1 = _fold_bn(self.conv1, self.bn1), nn.Identity()
            self.conv2, self.bn2 = _fold_bn(self.conv2, self.bn2), nn.Identity()
            if self.downsample is not None:
                self.downsample = nn.Sequential(_fold_bn(self.downsample[0], self.downsample[1]))
                self.conv2.bias += self.downsample[0].bias   # one bias before the ReLU instead of two (see ResNet50Features)
                self.downsample[0].bias = None
        self.folded = True

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        if getattr(self, "folded", False) and x.is_cuda and not torch.is_grad_enabled():
            c1, c2 = self.conv1, self.conv2
            y = torch.cudnn_convolution_relu(x, c1.weight, c1.bias, c1.stride, c1.padding, c1.dilation, c1.groups)
            return torch.cudnn_convolution_add_relu(y, c2.weight, idt, 1.0, c2.bias, c2.stride, c2.padding, c2.dilation,
                                                    c2.groups)
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + idt)


class _Resnet34_8s(nn.Module):
    """lib/networks/resnet_dilated.py:287-327: ResNet-34, output stride 8 (layer3 / layer4 dilated 2 / 4 instead of
    strided), fully-convolutional 1x1 classifier to ``num_classes`` channels, bilinear upsampling to the input size."""

    def __init__(self, num_classes=64):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cfg = [(64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 1, 2), (512, 3, 1, 4)]  # (planes, blocks, stride, dilation)
        layers, cin = [], 64
        for planes, blocks, stride, dil in cfg:
            ds = None
            if stride != 1 or cin != planes:
                ds = nn.Sequential(nn.Conv2d(cin, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
            blk = [_BasicBlock(cin, planes, stride, dil, ds)]
            blk += [_BasicBlock(planes, planes, 1, dil) for _ in range(blocks - 1)]
            layers.append(nn.Sequential(*blk))
            cin = planes
        self.layer1, self.layer2, self.layer3, self.layer4 = layers
        self.fc = nn.Conv2d(512, num_classes, 1)

    def forward(self, x, upsample=True):
        size = x.shape[-2:]
        x = F.relu(self.bn1(self.conv1(x)))
        if x.is_cuda and not torch.is_grad_enabled() and not x.is_contiguous() and x.permute(0, 2, 3, 1).is_contiguous():
            from . import ops
            x = ops.maxpool3x3s2_channels_last(x)   # same pool as nn.MaxPool2d(3, 2, 1); ATen's NHWC kernel is 4x slower
        else:
            x = self.maxpool(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.fc(x)
        if not upsample:   # the caller sums the two streams at 1/8 resolution and up-samples once (linear: same result)
            return x
        return F.interpolate(x, size=size, mode="bilinear", align_corners=True)  # upsample_bilinear (:325)


class SegnetEmbedding(nn.Module):
    """SEGNET.forward(img, label, depth) for INPUT RGBD / FUSION_TYPE add (lib/networks/SEG.py:88-120): two ResNet34-8s
    streams, features summed; -> [B,64,H,W] (the META_ARCH wrapper L2-normalises it, :298)."""

    def __init__(self, seed=0, use_depth=True):
        super().__init__()
        g = torch.random.get_rng_state()
        torch.manual_seed(6000 + seed)
        self.fcn = _Resnet34_8s(64)
        self.fcn_depth = _Resnet34_8s(64) if use_depth else None
        torch.random.set_rng_state(g)
        self.eval()
        for net in (self.fcn, self.fcn_depth):
            if net is not None:
                with torch.no_grad():
                    net.conv1, net.bn1 = _fold_bn(net.conv1, net.bn1), nn.Identity()
                for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
                    for blk in layer:
                        blk.fold()
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()
        self.tf32 = _TF32
        self.fmt = torch.channels_last if _TF32 else torch.contiguous_format
        self.to(memory_format=self.fmt)

    def forward(self, img, label=None, depth=None):
        img = img.contiguous(memory_format=self.fmt)
        with _conv_math(self.tf32):
            if img.is_cuda and not torch.is_grad_enabled():
                # up(a) + up(b) == up(a + b): one up-sampling of the summed 1/8-resolution maps, written NCHW-contiguous
                # by the library's resample kernel (ATen's NHWC kernel: 165 us per stream at 480x640, plus the add and
                # the NHWC -> NCHW copy of the 79 MB map)
                from . import ops
                if self.fcn_depth is not None and depth is not None:
                    # the two streams are independent and, at batch 1, too small to fill the GPU one after the other:
                    # the depth stream runs on a side stream (a parallel branch of the step's CUDA graph)
                    main = torch.cuda.current_stream()
                    if getattr(self, "_side", None) is None:
                        self._side = torch.cuda.Stream()
                    self._side.wait_stream(main)
                    with torch.cuda.stream(self._side):
                        d = self.fcn_depth(depth.contiguous(memory_format=self.fmt), upsample=False)
                    s = self.fcn(img, upsample=False)
                    main.wait_stream(self._side)
                    d.record_stream(main)
                    s = s + d
                else:
                    s = self.fcn(img, upsample=False)
                return ops.resample_bilinear(s.float().contiguous(), img.shape[-2:], align_corners=True)
            f = self.fcn(img)
            if self.fcn_depth is not None and depth is not None:
                f = f + self.fcn_depth(depth.contiguous(memory_format=self.fmt))
        return f.contiguous()
