"""CPU restatement (plain torch functional ops, fp32) of the reference's PMFNet forward.  TEST INFRASTRUCTURE.

Works from a ``state_dict`` with the reference's key names, so it needs neither the
reference tree nor torchvision.  Each function cites the reference lines it follows
(paths relative to /root/reference).  Pinned against the real reference by tests/test_oracle_pinning.py
and tests/golden/pmf_*.npz (see oracle/__init__.py).

``train=True`` uses batch statistics in every BatchNorm (biased variance for normalisation) exactly as
``nn.BatchNorm2d.train()`` does; running-stat side effects are returned in ``new_stats`` instead of being
applied in place.  Dropout2d sites take explicit masks (``dropout`` dict: site -> (B,C,1,1) tensor holding
0 or 1/(1-p)); a missing site means "no dropout" (eval semantics).

Two optional EMULATION switches of ``Ctx`` (off by default; what is pinned to the reference is the plain fp32 path):
``tf32=True`` rounds conv operands to tf32, ``half_pre_bn=True`` rounds the input of a training-mode BatchNorm to fp16 where
the device's "f16" mode stores it in fp16.  They exist so that the kernel tests can tell wiring errors (O(1)) from the
expected rounding noise: the device result is compared with an oracle that rounds at the same points.
"""
import math

import torch
import torch.nn.functional as F

LEAKY = 0.01  # nn.LeakyReLU() default, salsanext.py:13
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

RESNET_LAYERS = {"resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
RESNET_EXPANSION = {"resnet34": 1, "resnet50": 4, "resnet101": 4, "resnet152": 4}


def round_tf32(x):
    """Round-to-nearest (ties away from zero) to the 10-bit tf32 mantissa, i.e. PTX cvt.rna.tf32.f32, as a
    differentiable straight-through op.  Used only by the optional tensor-core EMULATION mode below."""
    with torch.no_grad():
        bits = x.detach().contiguous().view(torch.int32)
        r = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return x + (r - x).detach()


def round_half(x):
    """Round to fp16 (saturating) as a differentiable straight-through op: EMULATION of the "f16" mode's fp16 storage of
    the pre-BatchNorm activation (pmfb_conv_desc.out_half)."""
    with torch.no_grad():
        r = x.detach().clamp(-65504.0, 65504.0).half().float()
    return x + (r - x).detach()


class Ctx:
    """tf32=False (default): the reference's fp32 arithmetic (this is what is pinned to the reference).
    tf32=True: conv operands (input and weight) are rounded to tf32 before every convolution, accumulation stays
    fp32 — an emulation of the tcgen05 kind::tf32 path that lets tests separate indexing bugs (O(1) errors) from
    the expected operand-rounding noise.
    half_pre_bn=True (train mode only): the input of a training-mode BatchNorm that directly follows a convolution the
    engine's "f16" mode stores in fp16 (stride 1, taps inside the halo kernel's window, c_out a multiple of 8 and <= 256)
    is rounded to fp16 first — the batch statistics, the normalisation and both backward passes then see the rounded
    values, as on the device."""

    def __init__(self, sd, train=False, dropout=None, tf32=False, half_pre_bn=False):
        self.sd = sd
        self.train = train
        self.dropout = dropout or {}
        self.new_stats = {}
        self.tf32 = tf32
        self.half_pre_bn = half_pre_bn
        self._half_geom = False  # the engine would store the last convolution's (activated) output as fp16

    def conv(self, x, name, stride=1, padding=0, dilation=1):
        w = self.sd[name + ".weight"]
        if self.tf32:
            x, w = round_tf32(x), round_tf32(w)
        y = F.conv2d(x, w, self.sd.get(name + ".bias"), stride=stride, padding=padding, dilation=dilation)
        if self.half_pre_bn and self.train:
            k = w.shape[2]
            reach = max(padding, dilation * (k - 1) - padding) if k > 1 else 0
            self._half_geom = (stride == 1 and w.shape[0] % 8 == 0 and w.shape[0] <= 256 and (reach <= 2 or k == 7))
        return y

    def bn(self, x, name):
        w, b = self.sd[name + ".weight"], self.sd[name + ".bias"]
        rm, rv = self.sd[name + ".running_mean"], self.sd[name + ".running_var"]
        if not self.train:
            return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)
        if self.half_pre_bn and getattr(self, "_half_geom", False):
            x = round_half(x)
        self._half_geom = False
        rm2, rv2 = rm.detach().clone(), rv.detach().clone()
        y = F.batch_norm(x, rm2, rv2, w, b, True, BN_MOMENTUM, BN_EPS)
        self.new_stats[name + ".running_mean"] = rm2
        self.new_stats[name + ".running_var"] = rv2
        return y

    def drop(self, x, site):
        m = self.dropout.get(site)
        return x if m is None else x * m


def lrelu(x):
    return F.leaky_relu(x, LEAKY)


# ------------------------------------------------------------------ salsanext.py
def res_context_block(c, x, p):
    """salsanext.py:23-36"""
    shortcut = lrelu(c.conv(x, p + ".conv1"))
    a1 = c.bn(lrelu(c.conv(shortcut, p + ".conv2", padding=1)), p + ".bn1")
    a2 = c.bn(lrelu(c.conv(a1, p + ".conv3", padding=2, dilation=2)), p + ".bn2")
    return shortcut + a2


def res_block(c, x, p, pooling=True, drop_out=True):
    """salsanext.py:69-104"""
    shortcut = lrelu(c.conv(x, p + ".conv1"))
    a1 = c.bn(lrelu(c.conv(x, p + ".conv2", padding=1)), p + ".bn1")
    a2 = c.bn(lrelu(c.conv(a1, p + ".conv3", padding=2, dilation=2)), p + ".bn2")
    a3 = c.bn(lrelu(c.conv(a2, p + ".conv4", padding=1, dilation=2)), p + ".bn3")
    cat = torch.cat((a1, a2, a3), dim=1)
    a = c.bn(lrelu(c.conv(cat, p + ".conv5")), p + ".bn4")
    a = shortcut + a
    b = c.drop(a, p + ".dropout") if drop_out else a
    if pooling:
        b = F.avg_pool2d(b, kernel_size=3, stride=2, padding=1)
        return b, a
    return b


def up_block(c, x, skip, p, drop_out=True):
    """salsanext.py:136-164"""
    up_a = F.pixel_shuffle(x, 2)
    if drop_out:
        up_a = c.drop(up_a, p + ".dropout1")
    up_b = torch.cat((up_a, skip), dim=1)
    if drop_out:
        up_b = c.drop(up_b, p + ".dropout2")
    e1 = c.bn(lrelu(c.conv(up_b, p + ".conv1", padding=1)), p + ".bn1")
    e2 = c.bn(lrelu(c.conv(e1, p + ".conv2", padding=2, dilation=2)), p + ".bn2")
    e3 = c.bn(lrelu(c.conv(e2, p + ".conv3", padding=1, dilation=2)), p + ".bn3")
    cat = torch.cat((e1, e2, e3), dim=1)
    e = c.bn(lrelu(c.conv(cat, p + ".conv4")), p + ".bn4")
    if drop_out:
        e = c.drop(e, p + ".dropout3")
    return e


# ------------------------------------------------------------------ pmf_net.py
def fusion_block(c, pcd, img, p):
    """ResidualBasedFusionBlock.forward, pmf_net.py:31-36 (modules :13-29)."""
    cat = torch.cat((pcd, img), dim=1)
    fuse = c.bn(lrelu(c.conv(cat, p + ".fuse_conv.0", padding=1)), p + ".fuse_conv.2")
    att = F.relu(c.bn(c.conv(fuse, p + ".attention.0", padding=1), p + ".attention.1"))
    att = torch.sigmoid(c.bn(c.conv(att, p + ".attention.3", padding=1), p + ".attention.4"))
    return fuse * att + pcd


def aspp(c, x, p):
    """ASPP.forward, pmf_net.py:119-138"""
    size = x.shape[2:]
    img = c.conv(F.adaptive_avg_pool2d(x, 1), p + ".conv")
    img = F.interpolate(img, size=size, mode="bilinear")
    b1 = c.conv(x, p + ".atrous_block1")
    b6 = c.conv(x, p + ".atrous_block6", padding=6, dilation=6)
    b12 = c.conv(x, p + ".atrous_block12", padding=12, dilation=12)
    b18 = c.conv(x, p + ".atrous_block18", padding=18, dilation=18)
    return c.conv(torch.cat([img, b1, b6, b12, b18], dim=1), p + ".conv_1x1_output")


def basic_block(c, x, p, stride):
    """torchvision.models.resnet.BasicBlock.forward (torchvision 0.14/0.26: identical)."""
    out = F.relu(c.bn(c.conv(x, p + ".conv1", stride=stride, padding=1), p + ".bn1"))
    out = c.bn(c.conv(out, p + ".conv2", padding=1), p + ".bn2")
    if (p + ".downsample.0.weight") in c.sd:
        x = c.bn(c.conv(x, p + ".downsample.0", stride=stride), p + ".downsample.1")
    return F.relu(out + x)


def bottleneck(c, x, p, stride):
    """torchvision.models.resnet.Bottleneck.forward (v1.5: stride on the 3x3)."""
    out = F.relu(c.bn(c.conv(x, p + ".conv1"), p + ".bn1"))
    out = F.relu(c.bn(c.conv(out, p + ".conv2", stride=stride, padding=1), p + ".bn2"))
    out = c.bn(c.conv(out, p + ".conv3"), p + ".bn3")
    if (p + ".downsample.0.weight") in c.sd:
        x = c.bn(c.conv(x, p + ".downsample.0", stride=stride), p + ".downsample.1")
    return F.relu(out + x)


def resnet_encoder(c, x, p, backbone):
    """ResNet.forward, pmf_net.py:83-100 (stride-1 7x7 stem :69-70)."""
    h, w = x.shape[2], x.shape[3]
    assert h % 16 == 0 and w % 16 == 0, "invalid input size: {}".format(x.shape)
    block = basic_block if RESNET_EXPANSION[backbone] == 1 else bottleneck
    x = F.relu(c.bn(c.conv(x, p + ".conv1", padding=3), p + ".bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, nblocks in enumerate(RESNET_LAYERS[backbone], start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            x = block(c, x, "%s.layer%d.%d" % (p, li, bi), stride)
        if li >= 3:
            x = c.drop(x, "%s.dropout.layer%d" % (p, li))
        feats.append(x)
    return feats


def salsanext_fusion(c, x, img_feats, p):
    """SalsaNextFusion.forward, pmf_net.py:153-180"""
    d = res_context_block(c, x, p + ".downCntx")
    d = res_context_block(c, d, p + ".downCntx2")
    d = res_context_block(c, d, p + ".downCntx3")
    d0c, d0b = res_block(c, d, p + ".resBlock1", pooling=True, drop_out=False)
    d0c = fusion_block(c, d0c, img_feats[0], p + ".fusionblock_1")
    d1c, d1b = res_block(c, d0c, p + ".resBlock2")
    d1c = fusion_block(c, d1c, img_feats[1], p + ".fusionblock_2")
    d2c, d2b = res_block(c, d1c, p + ".resBlock3")
    d2c = fusion_block(c, d2c, img_feats[2], p + ".fusionblock_3")
    d3c, d3b = res_block(c, d2c, p + ".resBlock4")
    d3c = fusion_block(c, d3c, img_feats[3], p + ".fusionblock_4")
    d5c = aspp(c, res_block(c, d3c, p + ".resBlock5", pooling=False), p + ".aspp")
    u4 = up_block(c, d5c, d3b, p + ".upBlock1")
    u3 = up_block(c, u4, d2b, p + ".upBlock2")
    u2 = up_block(c, u3, d1b, p + ".upBlock3")
    u1 = up_block(c, u2, d0b, p + ".upBlock4", drop_out=False)
    logits = c.conv(u1, p + ".logits")
    return F.softmax(logits, dim=1)


def rgb_decoder(c, feats, p):
    """RGBDecoder.forward, pmf_net.py:214-222"""

    def stage(x, name, padding):
        x = c.bn(lrelu(c.conv(x, name + ".0", padding=padding)), name + ".2")
        return F.interpolate(x, scale_factor=2, mode="bilinear")

    u4 = stage(feats[3], p + ".up_4a", 1)
    u3 = stage(torch.cat((u4, feats[2]), dim=1), p + ".up_3a", 1)
    u2 = stage(torch.cat((u3, feats[1]), dim=1), p + ".up_2a", 1)
    u1 = stage(torch.cat((u2, feats[0]), dim=1), p + ".up_1a", 0)
    return F.softmax(c.conv(u1, p + ".conv", padding=1), dim=1)


def pmf_forward(sd, pcd_feature, img_feature, backbone="resnet34", train=False, dropout=None, return_ctx=False,
                tf32=False):
    """PMFNet.forward, pmf_net.py:242-249 -> (lidar_pred, camera_pred) softmax maps."""
    c = Ctx(sd, train=train, dropout=dropout, tf32=tf32)
    feats = resnet_encoder(c, img_feature, "camera_stream_encoder", backbone)
    lidar = salsanext_fusion(c, pcd_feature, feats, "lidar_stream")
    camera = rgb_decoder(c, feats, "camera_stream_decoder")
    if return_ctx:
        return lidar, camera, c
    return lidar, camera


# ------------------------------------------------------------------ parameter inventory + synthetic weights
def pmf_param_shapes(nclasses=20, base_channels=32, backbone="resnet34", pcd_channels=5, img_channels=3):
    """Ordered {state_dict key: shape} of PMFNet (pmf_net.py:224-240)."""
    shapes = {}

    def conv(name, cin, cout, kh, kw, bias=True):
        shapes[name + ".weight"] = (cout, cin, kh, kw)
        if bias:
            shapes[name + ".bias"] = (cout,)

    def bn(name, ch):
        shapes[name + ".weight"] = (ch,)
        shapes[name + ".bias"] = (ch,)
        shapes[name + ".running_mean"] = (ch,)
        shapes[name + ".running_var"] = (ch,)
        shapes[name + ".num_batches_tracked"] = ()

    exp = RESNET_EXPANSION[backbone]
    e = "camera_stream_encoder"
    conv(e + ".conv1", img_channels, 64, 7, 7, bias=False)
    bn(e + ".bn1", 64)
    inpl = 64
    for li, nblocks in enumerate(RESNET_LAYERS[backbone], start=1):
        planes = 64 * 2 ** (li - 1)
        for bi in range(nblocks):
            p = "%s.layer%d.%d" % (e, li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            if exp == 1:
                conv(p + ".conv1", inpl, planes, 3, 3, False); bn(p + ".bn1", planes)
                conv(p + ".conv2", planes, planes, 3, 3, False); bn(p + ".bn2", planes)
            else:
                conv(p + ".conv1", inpl, planes, 1, 1, False); bn(p + ".bn1", planes)
                conv(p + ".conv2", planes, planes, 3, 3, False); bn(p + ".bn2", planes)
                conv(p + ".conv3", planes, planes * 4, 1, 1, False); bn(p + ".bn3", planes * 4)
            if bi == 0 and (stride != 1 or inpl != planes * exp):
                conv(p + ".downsample.0", inpl, planes * exp, 1, 1, False); bn(p + ".downsample.1", planes * exp)
            inpl = planes * exp
    fch = [64 * exp, 128 * exp, 256 * exp, 512 * exp]
    d, dc = "camera_stream_decoder", 16 * exp
    conv(d + ".up_4a.0", fch[3], dc, 3, 3); bn(d + ".up_4a.2", dc)
    conv(d + ".up_3a.0", fch[2] + dc, dc, 3, 3); bn(d + ".up_3a.2", dc)
    conv(d + ".up_2a.0", fch[1] + dc, dc, 3, 3); bn(d + ".up_2a.2", dc)
    conv(d + ".up_1a.0", fch[0] + dc, dc, 1, 1); bn(d + ".up_1a.2", dc)
    conv(d + ".conv", dc, nclasses, 3, 3)
    l, b = "lidar_stream", base_channels
    for name, cin in ((".downCntx", pcd_channels), (".downCntx2", b), (".downCntx3", b)):
        conv(l + name + ".conv1", cin, b, 1, 1)
        conv(l + name + ".conv2", b, b, 3, 3); bn(l + name + ".bn1", b)
        conv(l + name + ".conv3", b, b, 3, 3); bn(l + name + ".bn2", b)
    for i, (cin, cout) in enumerate(((b, 2 * b), (2 * b, 4 * b), (4 * b, 8 * b), (8 * b, 8 * b), (8 * b, 8 * b)), start=1):
        p = "%s.resBlock%d" % (l, i)
        conv(p + ".conv1", cin, cout, 1, 1)
        conv(p + ".conv2", cin, cout, 3, 3); bn(p + ".bn1", cout)
        conv(p + ".conv3", cout, cout, 3, 3); bn(p + ".bn2", cout)
        conv(p + ".conv4", cout, cout, 2, 2); bn(p + ".bn3", cout)
        conv(p + ".conv5", 3 * cout, cout, 1, 1); bn(p + ".bn4", cout)
    for i, (cin, cout) in enumerate(((8 * b, 4 * b), (4 * b, 4 * b), (4 * b, 2 * b), (2 * b, b)), start=1):
        p = "%s.upBlock%d" % (l, i)
        conv(p + ".conv1", cin // 4 + 2 * cout, cout, 3, 3); bn(p + ".bn1", cout)
        conv(p + ".conv2", cout, cout, 3, 3); bn(p + ".bn2", cout)
        conv(p + ".conv3", cout, cout, 2, 2); bn(p + ".bn3", cout)
        conv(p + ".conv4", 3 * cout, cout, 1, 1); bn(p + ".bn4", cout)
    conv(l + ".logits", b, nclasses, 1, 1)
    for i, (pc, ic) in enumerate(((2 * b, fch[0]), (4 * b, fch[1]), (8 * b, fch[2]), (8 * b, fch[3])), start=1):
        p = "%s.fusionblock_%d" % (l, i)
        conv(p + ".fuse_conv.0", pc + ic, pc, 3, 3); bn(p + ".fuse_conv.2", pc)
        conv(p + ".attention.0", pc, pc, 3, 3); bn(p + ".attention.1", pc)
        conv(p + ".attention.3", pc, pc, 3, 3); bn(p + ".attention.4", pc)
    a = l + ".aspp"
    conv(a + ".conv", 8 * b, 8 * b, 1, 1)
    conv(a + ".atrous_block1", 8 * b, 8 * b, 1, 1)
    conv(a + ".atrous_block6", 8 * b, 8 * b, 3, 3)
    conv(a + ".atrous_block12", 8 * b, 8 * b, 3, 3)
    conv(a + ".atrous_block18", 8 * b, 8 * b, 3, 3)
    conv(a + ".conv_1x1_output", 40 * b, 8 * b, 1, 1)
    return shapes


def synth_state_dict(shapes, seed=1):
    """Deterministic, construction-order-independent weights: every tensor is drawn from a generator
    seeded by (seed, crc32(key)).  Conv weights are He-scaled so activations stay O(1) through 50 layers;
    BN affine/running stats are perturbed away from their defaults so eval-mode BN is non-trivial."""
    import zlib

    sd = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(1.5 / fan_in)
        elif k.endswith(".weight"):  # BN gamma
            sd[k] = 0.75 + 0.5 * torch.rand(shp, generator=g)
        else:  # conv / BN bias
            sd[k] = 0.05 * torch.randn(shp, generator=g)
    return sd
