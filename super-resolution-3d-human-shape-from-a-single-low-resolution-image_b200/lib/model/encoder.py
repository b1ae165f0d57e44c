"""The image encoder of SuRS in plain PyTorch (kept in PyTorch by design: it is the producer of the
two feature maps the CUDA path consumes).

Module / parameter names mirror the reference so that a reference checkpoint loads with
``strict=True``:
  ``super_resolution``  <- lib/model/SuRSSR_v3.py:26-181   (bicubic x2 -> U-Net with PixelShuffle)
  ``image_filter_lr``   <- lib/model/HGFilters.py:121-208  (down_type 'low_res': ConvBlock + 3 stacked hourglasses)
  ``image_filter_hr``   <- lib/model/HGFilters.py:121-208  (down_type 'high_res': only the 1x1 ``conv5`` runs)
Unused-but-present modules of the reference (conv1/bn1/conv3/conv4 of both filters, the hourglass
stack inside ``image_filter_hr``, ``bn4`` of equal-width ConvBlocks, ``sub_mean`` / ``add_mean``) are kept so
the state dict has the same 553 - 20 tensors.  Written from the architecture description, not copied.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv3x3(cin, cout, stride=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=bias)


def _norm(kind, ch):
    return nn.BatchNorm2d(ch) if kind == "batch" else nn.GroupNorm(32, ch)


class ConvBlock(nn.Module):
    """Pre-activation residual block: three 3x3 convs of widths out/2, out/4, out/4, concatenated."""

    def __init__(self, in_planes, out_planes, norm="batch"):
        super().__init__()
        widths = [in_planes, out_planes // 2, out_planes // 4, out_planes // 4]
        for n in range(3):
            setattr(self, "conv%d" % (n + 1), _conv3x3(widths[n], widths[n + 1]))
        for n in range(3):
            setattr(self, "bn%d" % (n + 1), _norm(norm, widths[n]))
        self.bn4 = _norm(norm, in_planes)
        self.downsample = None
        if in_planes != out_planes:
            self.downsample = nn.Sequential(self.bn4, nn.ReLU(True), nn.Conv2d(in_planes, out_planes, 1, bias=False))

    def forward(self, x):
        parts, y = [], x
        for n in (1, 2, 3):
            y = getattr(self, "conv%d" % n)(F.relu(getattr(self, "bn%d" % n)(y), True))
            parts.append(y)
        out = torch.cat(parts, 1)
        return out + (x if self.downsample is None else self.downsample(x))


class HourGlass(nn.Module):
    def __init__(self, depth, n_features, norm="batch"):
        super().__init__()
        self.depth = depth
        for level in range(depth, 0, -1):
            self.add_module("b1_%d" % level, ConvBlock(n_features, n_features, norm))
            self.add_module("b2_%d" % level, ConvBlock(n_features, n_features, norm))
        self.add_module("b2_plus_1", ConvBlock(n_features, n_features, norm))
        for level in range(1, depth + 1):
            self.add_module("b3_%d" % level, ConvBlock(n_features, n_features, norm))

    def _level(self, level, x):
        up = self._modules["b1_%d" % level](x)
        low = self._modules["b2_%d" % level](F.avg_pool2d(x, 2, stride=2))
        low = self._level(level - 1, low) if level > 1 else self._modules["b2_plus_1"](low)
        low = self._modules["b3_%d" % level](low)
        return up + F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)

    def forward(self, x):
        return self._level(self.depth, x)


class HGFilter(nn.Module):
    def __init__(self, stack, depth, in_ch, last_ch, norm="batch", down_type="conv64", use_sigmoid=True):
        super().__init__()
        self.n_stack, self.down_type, self.use_sigmoid = stack, down_type, use_sigmoid
        self.conv1 = nn.Conv2d(in_ch, 64, kernel_size=7, stride=2, padding=3)
        self.bn1 = _norm(norm, 64)
        if down_type == "conv64":
            self.conv2 = ConvBlock(64, 64, norm)
            self.down_conv2 = nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1)
        elif down_type == "low_res":
            self.conv2 = ConvBlock(256, 256, norm)
        elif down_type == "high_res":
            self.conv2 = ConvBlock(64, 128, norm)
        self.conv3 = ConvBlock(128, 128, norm)
        self.conv4 = ConvBlock(128, 256, norm)
        self.conv5 = nn.Conv2d(64, 64, kernel_size=1)
        for s in range(stack):
            self.add_module("m%d" % s, HourGlass(depth, 256, norm))
            self.add_module("top_m_%d" % s, ConvBlock(256, 256, norm))
            self.add_module("conv_last%d" % s, nn.Conv2d(256, 256, kernel_size=1))
            self.add_module("bn_end%d" % s, _norm(norm, 256))
            self.add_module("l%d" % s, nn.Conv2d(256, last_ch, kernel_size=1))
            if s < stack - 1:
                self.add_module("bl%d" % s, nn.Conv2d(256, 256, kernel_size=1))
                self.add_module("al%d" % s, nn.Conv2d(last_ch, 256, kernel_size=1))

    def forward(self, x):
        if self.down_type == "high_res":                     # reference HGFilters.py:179-181
            return [self.conv5(x)]
        m = self._modules
        previous = self.conv2(x)
        outputs = []
        for s in range(self.n_stack):
            ll = m["top_m_%d" % s](m["m%d" % s](previous))
            ll = F.relu(m["bn_end%d" % s](m["conv_last%d" % s](ll)), True)
            out = m["l%d" % s](ll)
            outputs.append(torch.tanh(out) if self.use_sigmoid else out)
            if s < self.n_stack - 1:
                previous = previous + m["bl%d" % s](ll) + m["al%d" % s](out)
        return outputs


class _ResBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(ch, ch, 3, padding=1), nn.ReLU(True), nn.Conv2d(ch, ch, 3, padding=1))

    def forward(self, x):
        return self.body(x) + x


class _MeanShift(nn.Conv2d):
    def __init__(self, rgb_range, mean, sign):
        super().__init__(3, 3, kernel_size=1)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1)
        self.bias.data = sign * rgb_range * torch.tensor(mean)


def _act_conv(cin, cout, stride=1):
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1), nn.LeakyReLU(0.2, True)]


class SuRSSR_v3(nn.Module):
    """reference lib/model/SuRSSR_v3.py: returns (img_SR, feature_lr [B,256,S/2,S/2], feature_hr [B,64,2S,2S])."""

    def __init__(self, opt):
        super().__init__()
        nb = list(opt.n_block)
        mean = (0.4488, 0.4371, 0.4040)
        self.residual = opt.residual
        self.sub_mean = _MeanShift(opt.rgb_range, mean, -1)
        self.add_mean = _MeanShift(opt.rgb_range, mean, 1)
        self.head = nn.Sequential(*_act_conv(3, 32))
        self.down1 = nn.Sequential(*_act_conv(32, 32, 2))
        self.body1 = nn.Sequential(*[_ResBlock(32) for _ in range(nb[0])])
        self.tail1 = nn.Sequential(*(_act_conv(32, 32) + _act_conv(32, 64)))
        self.down2 = nn.Sequential(*_act_conv(64, 64, 2))
        self.body2 = nn.Sequential(*[_ResBlock(64) for _ in range(nb[1])])
        self.tail2 = nn.Sequential(*(_act_conv(64, 64) + _act_conv(64, 128)))
        self.down3 = nn.Sequential(*_act_conv(128, 128, 2))
        self.body3 = nn.Sequential(*[_ResBlock(128) for _ in range(nb[2])])
        self.tail3 = nn.Sequential(*(_act_conv(128, 128) + _act_conv(128, 256)))
        self.bottleneck = nn.Sequential(*_act_conv(256, 256))
        self.bott2 = nn.Sequential(*_act_conv(512, 512))
        self.pixel_shuffle = nn.Sequential(nn.PixelShuffle(2), nn.LeakyReLU(0.2, True))
        self.ups2 = nn.Sequential(*_act_conv(256, 256))
        self.ups3 = nn.Sequential(*_act_conv(128, 128))
        self.ups4 = nn.Sequential(*_act_conv(64, 64))
        self.last = nn.Sequential(*(_act_conv(64, 32) + [nn.Conv2d(32, 3, kernel_size=3, padding=1)]))
        self.upsample = nn.Upsample(scale_factor=opt.scale, mode="bicubic", align_corners=False)

    def forward(self, x):
        h = self.head(self.upsample(x))
        d1 = self.down1(h)
        d1_f = self.tail1(self.body1(d1) if self.residual else d1)
        d2 = self.down2(d1_f)
        d2_f = self.tail2(self.body2(d2) if self.residual else d2)
        d3 = self.down3(d2_f)
        d3_f = self.tail3(self.body3(d3) if self.residual else d3)
        up1 = self.pixel_shuffle(self.bott2(torch.cat((d3_f, self.bottleneck(d3_f)), 1)))
        feature_lr = torch.cat((d2_f, up1), 1)
        up2 = self.pixel_shuffle(self.ups2(feature_lr))
        up3 = self.pixel_shuffle(self.ups3(torch.cat((d1_f, up2), 1)))
        feature_hr = self.ups4(torch.cat((h, up3), 1))
        return self.last(feature_hr), feature_lr, feature_hr


def init_weights(net, gain=0.02):
    """reference lib/net_util.py:99-132 ('normal'): conv / linear weights N(0, gain), zero bias."""
    for m in net.modules():
        name = m.__class__.__name__
        if hasattr(m, "weight") and (name.find("Conv") != -1 or name.find("Linear") != -1 or name == "_MeanShift"):
            nn.init.normal_(m.weight.data, 0.0, gain)
            if getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif name.find("BatchNorm2d") != -1:
            nn.init.normal_(m.weight.data, 1.0, gain)
            nn.init.constant_(m.bias.data, 0.0)
