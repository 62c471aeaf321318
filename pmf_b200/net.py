"""The PMF dual-stream graph expressed over pmf_b200.engine ops (one fused C-ABI call chain per reference module).

Each function cites the reference forward it replaces (paths relative to the reference tree).  Concats never
materialise: producers write straight into channel slices of the consumer's input buffer.
"""
import torch

from .engine import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, BNParam, ConvParam

RESNET_LAYERS = {"resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
RESNET_EXPANSION = {"resnet34": 1, "resnet50": 4, "resnet101": 4, "resnet152": 4}


class ModuleParams:
    """conv(name)/bn(name) lookups on the LIVE module tree (so replaceBN / DDP / load_state_dict are honoured)."""

    def __init__(self, root, prefix=""):
        self.mods = dict(root.named_modules())
        self.prefix = prefix
        self._conv = {}

    def conv(self, name):
        name = name.lstrip(".")
        cp = self._conv.get(name)
        if cp is None:
            m = self.mods[name]
            dil, pad, stride = m.dilation[0], m.padding[0], m.stride[0]
            assert m.dilation[0] == m.dilation[1] and m.padding[0] == m.padding[1] and m.stride[0] == m.stride[1]
            stem = (m.kernel_size == (7, 7) and m.in_channels <= 4)
            cp = ConvParam(self.prefix + name, m.weight, m.bias, dil, pad, stride, stem=stem)
            self._conv[name] = cp
        return cp

    def bn(self, name):
        name = name.lstrip(".")
        return BNParam(self.prefix + name, self.mods[name])


# ---------------------------------------------------------------------------------------------- salsanext.py
def res_context_block(E, x, p):
    """ResContextBlock.forward, salsanext.py:23-36."""
    s = E.conv_act(x, p + ".conv1", ACT_LEAKY)
    a1 = E.conv_act_bn(s, p + ".conv2", p + ".bn1")
    return E.conv_act_bn(a1, p + ".conv3", p + ".bn2", shortcut=s)


def res_block(E, x, p, pooling=True, drop_out=True, pool_out=None):
    """ResBlock.forward, salsanext.py:69-104.  Returns (pooled, skip) or the un-pooled output."""
    n, h, w, _ = x.shape
    c = E.P.conv(p + ".conv1").c_out
    s = E.conv_act(x, p + ".conv1", ACT_LEAKY)
    cat = E.new(n, h, w, 3 * c)
    a1 = E.conv_act_bn(x, p + ".conv2", p + ".bn1", out=cat.slice(0, c))
    a2 = E.conv_act_bn(a1, p + ".conv3", p + ".bn2", out=cat.slice(c, 2 * c))
    E.conv_act_bn(a2, p + ".conv4", p + ".bn3", out=cat.slice(2 * c, 3 * c))
    mask = E.mask_for(p + ".dropout", n, c) if drop_out else None
    if not pooling:
        return E.conv_act_bn(cat, p + ".conv5", p + ".bn4", shortcut=s, mask=mask)
    res_a = E.conv_act_bn(cat, p + ".conv5", p + ".bn4", shortcut=s)
    res_b = E.pool(res_a, "avg", out=pool_out, mask=mask)  # pool(dropout(x)) == mask * pool(x)
    return res_b, res_a


def up_block(E, x, skip, p, drop_out=True):
    """UpBlock.forward, salsanext.py:136-164."""
    n, h, w, c4 = x.shape
    ca = c4 // 4
    c = E.P.conv(p + ".conv1").c_out
    cat = E.new(n, 2 * h, 2 * w, ca + skip.c)
    m1 = E.mask_for(p + ".dropout1", n, ca) if drop_out else None
    m2 = E.mask_for(p + ".dropout2", n, ca + skip.c) if drop_out else None
    ma = None
    if m1 is not None or m2 is not None:  # both Dropout2d scales of the up-sampled half fold into one
        ma = (m1 if m1 is not None else 1.0) * (m2[:, :ca] if m2 is not None else 1.0)
        ma = ma.contiguous()
    E.pixel_shuffle(x, cat.slice(0, ca), mask=ma)
    E.copy(skip, cat.slice(ca, ca + skip.c), mask=None if m2 is None else m2[:, ca:].contiguous())
    cat2 = E.new(n, 2 * h, 2 * w, 3 * c)
    e1 = E.conv_act_bn(cat, p + ".conv1", p + ".bn1", out=cat2.slice(0, c))
    e2 = E.conv_act_bn(e1, p + ".conv2", p + ".bn2", out=cat2.slice(c, 2 * c))
    E.conv_act_bn(e2, p + ".conv3", p + ".bn3", out=cat2.slice(2 * c, 3 * c))
    m3 = E.mask_for(p + ".dropout3", n, c) if drop_out else None
    return E.conv_act_bn(cat2, p + ".conv4", p + ".bn4", mask=m3)


# ---------------------------------------------------------------------------------------------- pmf_net.py
def fusion_block(E, cat, pcd_c, p, rnd=True):
    """ResidualBasedFusionBlock.forward, pmf_net.py:31-36; ``cat`` already holds [pcd | img] channels.
    rnd=False keeps the result un-rounded (standalone module: the value goes back to the caller, not into a conv)."""
    pcd = cat.slice(0, pcd_c)
    f = E.conv_act_bn(cat, p + ".fuse_conv.0", p + ".fuse_conv.2")
    a = E.conv_bn(f, p + ".attention.0", p + ".attention.1", post=ACT_RELU)
    return E.conv_bn(a, p + ".attention.3", p + ".attention.4", post=ACT_SIGMOID, gate=(f, pcd), rnd=rnd)


def aspp(E, x, p):
    """ASPP.forward, pmf_net.py:119-138."""
    n, h, w, c = x.shape
    d = E.P.conv(p + ".conv").c_out
    cat = E.new(n, h, w, 5 * d)
    g = E.conv_act(E.global_avg(x), p + ".conv", ACT_NONE)
    E.broadcast(g, cat.slice(0, d))
    E.conv_act(x, p + ".atrous_block1", ACT_NONE, out=cat.slice(d, 2 * d))
    E.conv_act(x, p + ".atrous_block6", ACT_NONE, out=cat.slice(2 * d, 3 * d))
    E.conv_act(x, p + ".atrous_block12", ACT_NONE, out=cat.slice(3 * d, 4 * d))
    E.conv_act(x, p + ".atrous_block18", ACT_NONE, out=cat.slice(4 * d, 5 * d))
    return E.conv_act(cat, p + ".conv_1x1_output", ACT_NONE)


def basic_block(E, x, p, has_ds, mask=None):
    """torchvision BasicBlock.forward."""
    idn = x
    if has_ds:  # run first so that its (parity-sparse) stride-2 dgrad accumulates after conv1's dense one
        idn = E.conv_bn(x, p + ".downsample.0", p + ".downsample.1")
    o = E.conv_bn(x, p + ".conv1", p + ".bn1", post=ACT_RELU)
    return E.conv_bn(o, p + ".conv2", p + ".bn2", post=ACT_RELU, identity=idn, mask=mask)


def bottleneck(E, x, p, has_ds, mask=None):
    """torchvision Bottleneck.forward (v1.5: the stride sits on the 3x3)."""
    idn = x
    if has_ds:
        idn = E.conv_bn(x, p + ".downsample.0", p + ".downsample.1")
    o = E.conv_bn(x, p + ".conv1", p + ".bn1", post=ACT_RELU)
    o = E.conv_bn(o, p + ".conv2", p + ".bn2", post=ACT_RELU)
    return E.conv_bn(o, p + ".conv3", p + ".bn3", post=ACT_RELU, identity=idn, mask=mask)


def resnet_encoder(E, img7, p, backbone):
    """ResNet.forward, pmf_net.py:83-100 (stride-1 7x7 stem at :69-70, run as a 7-tap conv over the
    horizontally unrolled input)."""
    block = basic_block if RESNET_EXPANSION[backbone] == 1 else bottleneck
    x = E.conv_bn(img7, p + ".conv1", p + ".bn1", post=ACT_RELU)
    x = E.pool(x, "max")
    feats = []
    for li, nblocks in enumerate(RESNET_LAYERS[backbone], start=1):
        for bi in range(nblocks):
            bp = "%s.layer%d.%d" % (p, li, bi)
            has_ds = (bp + ".downsample.0") in E.P.mods
            mask = None
            if li >= 3 and bi == nblocks - 1:  # self.dropout(layer3/4 output), pmf_net.py:97-98
                c_out = E.P.bn(bp + (".bn2" if block is basic_block else ".bn3")).c
                mask = E.mask_for("%s.dropout.layer%d" % (p, li), x.shape[0], c_out)
            x = block(E, x, bp, has_ds, mask=mask)
        E.mark_ready(x)  # consumers on another stream (Engine.branch) wait for this point only, not for the whole encoder
        feats.append(x)
    return feats


def salsanext_fusion(E, pcd, feats, p, nclasses):
    """SalsaNextFusion.forward, pmf_net.py:153-180 -> NHWC logits Act."""
    n, h, w, _ = pcd.shape
    d = res_context_block(E, pcd, p + ".downCntx")
    d = res_context_block(E, d, p + ".downCntx2")
    d = res_context_block(E, d, p + ".downCntx3")
    skips = []
    x = d
    for i in range(4):
        c = E.P.conv("%s.resBlock%d.conv1" % (p, i + 1)).c_out
        img = feats[i]
        hh, ww = x.shape[1] // 2, x.shape[2] // 2
        cat = E.new(n, hh, ww, c + img.c)
        _, skip = res_block(E, x, "%s.resBlock%d" % (p, i + 1), pooling=True, drop_out=(i > 0), pool_out=cat.slice(0, c))
        E.wait_ready(img)
        E.copy(img, cat.slice(c, c + img.c))
        x = fusion_block(E, cat, c, "%s.fusionblock_%d" % (p, i + 1))
        skips.append(skip)
    x = res_block(E, x, p + ".resBlock5", pooling=False)
    x = aspp(E, x, p + ".aspp")
    x = up_block(E, x, skips[3], p + ".upBlock1")
    x = up_block(E, x, skips[2], p + ".upBlock2")
    x = up_block(E, x, skips[1], p + ".upBlock3")
    x = up_block(E, x, skips[0], p + ".upBlock4", drop_out=False)
    return E.conv_act(x, p + ".logits", ACT_NONE, rnd=False)


def rgb_decoder(E, feats, p):
    """RGBDecoder.forward, pmf_net.py:214-222 -> NHWC logits Act."""
    n = feats[0].shape[0]
    dc = E.P.conv(p + ".up_4a.0").c_out
    x = feats[3]
    for name, skip in ((".up_4a", feats[2]), (".up_3a", feats[1]), (".up_2a", feats[0]), (".up_1a", None)):
        y = E.conv_act_bn(x, p + name + ".0", p + name + ".2")
        _, hh, ww, _ = y.shape
        if skip is not None:
            cat = E.new(n, 2 * hh, 2 * ww, dc + skip.c)
            E.upsample2x(y, cat.slice(0, dc))
            E.copy(skip, cat.slice(dc, dc + skip.c))
            x = cat
        else:
            x = E.upsample2x(y, E.new(n, 2 * hh, 2 * ww, dc))
    return E.conv_act(x, p + ".conv", ACT_NONE, rnd=False)


# ---------------------------------------------------------------------------------------------- epmf_net.py (inference)
def sparse_context_block(E, x, p, out=None):
    """ResContextBlock.forward of epmf_net.py:66-82 (SparseVariantConv at :30-50), eval mode.
    x*mask is the identity for conv1 (the mask is derived from x) and conv2 (its input is LeakyReLU(y*mask))."""
    m0 = E.pixel_mask(x)
    s_pre, m1 = E.sparse_conv(x, m0, p + ".conv1")
    n, h, w, c = s_pre.shape
    s = E.pixel_scale(s_pre, pre=m1, act=ACT_LEAKY)
    a_pre, m2 = E.sparse_conv(s, m1, p + ".conv2")
    a1 = E.pixel_scale(a_pre, pre=m2, act=ACT_LEAKY, bn=p + ".bn1", post=m2)          # bn1(...) * mask: conv3's input
    b_pre, m3 = E.sparse_conv(a1, m2, p + ".conv3")
    return E.pixel_scale(b_pre, out=out, pre=m3, act=ACT_LEAKY, bn=p + ".bn2", r=s, post=m3)  # (shortcut + bn2(...)) * mask


def conv_lrelu_bn_shuffle(E, x, p, out):
    """nn.Sequential(Conv2d 3x3, LeakyReLU, BatchNorm2d, PixelShuffle(2)), epmf_net.py:97-102 / 138-143."""
    y = E.conv_act_bn(x, p + ".0", p + ".2")
    return E.pixel_shuffle(y, out)


def epmf_salsanext_fusion(E, pcd, feats, p, nclasses):
    """SalsaNextFusion.forward, epmf_net.py:104-131 -> (NHWC logits Act, down5c Act).  The LiDAR stream runs at half
    resolution from downCntx3 on and each fusion block PRECEDES its ResBlock."""
    n, h, w, _ = pcd.shape
    b = E.P.conv(p + ".downCntx.conv1.conv").c_out
    d = sparse_context_block(E, pcd, p + ".downCntx")
    d = sparse_context_block(E, d, p + ".downCntx2")
    img = feats[0]
    cat = E.new(n, h // 2, w // 2, b + img.c, needs_grad=False)
    sparse_context_block(E, d, p + ".downCntx3", out=cat.slice(0, b))
    E.wait_ready(img)
    E.copy(img, cat.slice(b, b + img.c))
    x = fusion_block(E, cat, b, p + ".fusionblock_1")
    skips = []
    for i in range(4):
        rb = "%s.resBlock%d" % (p, i + 1)
        c = E.P.conv(rb + ".conv1").c_out
        if i < 3:
            img = feats[i + 1]
            cat = E.new(n, x.shape[1] // 2, x.shape[2] // 2, c + img.c, needs_grad=False)
            _, skip = res_block(E, x, rb, pooling=True, drop_out=(i > 0), pool_out=cat.slice(0, c))
            E.wait_ready(img)
            E.copy(img, cat.slice(c, c + img.c))
            x = fusion_block(E, cat, c, "%s.fusionblock_%d" % (p, i + 2))
        else:
            x, skip = res_block(E, x, rb, pooling=True, drop_out=True)
        skips.append(skip)
    x = res_block(E, x, p + ".resBlock5", pooling=False)
    d5c = aspp(E, x, p + ".aspp")
    x = up_block(E, d5c, skips[3], p + ".upBlock1")
    x = up_block(E, x, skips[2], p + ".upBlock2")
    x = up_block(E, x, skips[1], p + ".upBlock3")
    x = up_block(E, x, skips[0], p + ".upBlock4", drop_out=False)
    up = conv_lrelu_bn_shuffle(E, x, p + ".extraUpSample", E.new(n, 2 * x.shape[1], 2 * x.shape[2], b, needs_grad=False))
    return E.conv_act(up, p + ".logits", ACT_NONE, rnd=False), d5c


def epmf_rgb_decoder(E, feats, lidar_feature, p):
    """RGBDecoder.forward, epmf_net.py:175-183 -> NHWC logits Act."""
    n = feats[0].shape[0]
    dc = E.P.conv(p + ".up_4a.0").c_out
    cu = E.P.conv(p + ".extraUpSample.0").c_out // 4
    ca = E.P.conv(p + ".aspp.conv_1x1_output").c_out
    _, hh, ww, _ = feats[3].shape
    fuse = E.new(n, hh, ww, cu + ca, needs_grad=False)
    conv_lrelu_bn_shuffle(E, lidar_feature, p + ".extraUpSample", fuse.slice(0, cu))
    E.copy(aspp(E, feats[3], p + ".aspp"), fuse.slice(cu, cu + ca))
    x = fuse
    for name, skip in ((".up_4a", feats[2]), (".up_3a", feats[1]), (".up_2a", feats[0]), (".up_1a", None)):
        y = E.conv_act_bn(x, p + name + ".0", p + name + ".2")
        _, hh, ww, _ = y.shape
        if skip is not None:
            cat = E.new(n, 2 * hh, 2 * ww, dc + skip.c, needs_grad=False)
            E.upsample2x(y, cat.slice(0, dc))
            E.copy(skip, cat.slice(dc, dc + skip.c))
            x = cat
        else:
            x = E.upsample2x(y, E.new(n, 2 * hh, 2 * ww, dc, needs_grad=False))
    return E.conv_act(x, p + ".conv", ACT_NONE, rnd=False)


def epmf_forward_packed(E, pcd, img7, backbone, nclasses):
    """EPMFNet.forward, epmf_net.py:209-216, from the packed NHWC inputs."""
    with E.branch():  # the camera encoder runs next to the LiDAR stream (the decoder needs the LiDAR stream's ASPP output)
        feats = resnet_encoder(E, img7, "camera_stream_encoder", backbone)
    lidar_logits, lidar_feature = epmf_salsanext_fusion(E, pcd, feats, "lidar_stream", nclasses)
    E.join_branch()
    camera_logits = epmf_rgb_decoder(E, feats, lidar_feature, "camera_stream_decoder")
    lidar = E.softmax_nchw(lidar_logits, nclasses)
    camera = E.softmax_nchw(camera_logits, nclasses)
    E.finish_forward()
    return lidar, camera, lidar_logits, camera_logits


def epmf_forward(E, pcd_feature, img_feature, backbone, nclasses):
    h, w = img_feature.shape[2], img_feature.shape[3]
    if h % 32 != 0 or w % 32 != 0:  # the half-resolution LiDAR stream pools four more times
        assert False, "invalid input size: {}".format(img_feature.shape)
    img7 = E.input_nchw(img_feature, 32, n_shift=7)
    pcd = E.input_nchw(pcd_feature, (pcd_feature.shape[1] + 3) // 4 * 4)
    return epmf_forward_packed(E, pcd, img7, backbone, nclasses)


def check_input_size(img_feature):
    h, w = img_feature.shape[2], img_feature.shape[3]
    if h % 16 != 0 or w % 16 != 0:
        assert False, "invalid input size: {}".format(img_feature.shape)  # pmf_net.py:87-88


def pmf_forward(E, pcd_feature, img_feature, backbone, nclasses):
    """PMFNet.forward, pmf_net.py:242-249.  Inputs are NCHW torch tensors (any strides); returns
    (lidar_probs, camera_probs, lidar_logits_act, camera_logits_act)."""
    check_input_size(img_feature)
    img7 = E.input_nchw(img_feature, 32, n_shift=7)
    pcd = E.input_nchw(pcd_feature, (pcd_feature.shape[1] + 3) // 4 * 4)
    return pmf_forward_packed(E, pcd, img7, backbone, nclasses)


def pmf_forward_packed(E, pcd, img7, backbone, nclasses):
    """The graph proper, from the packed NHWC inputs (pcd: channel-padded; img7: horizontally unrolled RGB)."""
    with E.branch():  # the camera stream runs next to the LiDAR stream; the four feature maps are handed over by events
        feats = resnet_encoder(E, img7, "camera_stream_encoder", backbone)
        camera_logits = rgb_decoder(E, feats, "camera_stream_decoder")
    lidar_logits = salsanext_fusion(E, pcd, feats, "lidar_stream", nclasses)
    E.join_branch()
    lidar = E.softmax_nchw(lidar_logits, nclasses)
    camera = E.softmax_nchw(camera_logits, nclasses)
    E.finish_forward()
    return lidar, camera, lidar_logits, camera_logits
