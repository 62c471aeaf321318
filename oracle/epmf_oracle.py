"""CPU restatement (plain torch functional ops, fp32) of the reference's EPMFNet forward.  TEST INFRASTRUCTURE.

Follows pc_processor/models/epmf_net.py (paths relative to /root/reference); shares the block restatements of
oracle/pmf_oracle.py.  Pinned against the live reference (bit-identical forward) by tests/test_oracle_pinning.py and
against tests/golden/epmf_*.npz.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import torch
import torch.nn.functional as F

from . import pmf_oracle as po


def sparse_conv(c, x, mask, name, padding=0, stride=1, dilation=1, ksize=3):
    """SparseVariantConv.forward, epmf_net.py:30-50: x*mask -> conv(+conv.bias) -> +bias -> *dilated mask.
    (The normalisation map `mask_conv` computed at :34-41 is never used by the reference: dead code, not restated.)"""
    x = x * mask
    mask = F.max_pool2d(F.pad(mask, (padding, padding, padding, padding)), ksize, stride=stride, padding=0, dilation=dilation)
    w = c.sd[name + ".conv.weight"]
    if c.tf32:
        x, w = po.round_tf32(x), po.round_tf32(w)
    x = F.conv2d(x, w, c.sd[name + ".conv.bias"], stride=stride, padding=padding, dilation=dilation)
    x = x + c.sd[name + ".bias"].view(1, -1, 1, 1)
    return x * mask, mask


def sparse_context_block(c, x, p, stride=1):
    """ResContextBlock.forward, epmf_net.py:66-82 (conv1 is 3x3, optionally strided: :55)."""
    mask = x.abs().sum(1).ne(0).float().unsqueeze(1)
    shortcut, mask = sparse_conv(c, x, mask, p + ".conv1", padding=1, stride=stride)
    shortcut = po.lrelu(shortcut)
    a, mask = sparse_conv(c, shortcut, mask, p + ".conv2", padding=1)
    a1 = c.bn(po.lrelu(a), p + ".bn1")
    a, mask = sparse_conv(c, a1, mask, p + ".conv3", padding=2, dilation=2)
    a2 = c.bn(po.lrelu(a), p + ".bn2")
    return (shortcut + a2) * mask


def conv_lrelu_bn_shuffle(c, x, p):
    """nn.Sequential(Conv2d 3x3 p1, LeakyReLU, BatchNorm2d, PixelShuffle(2)), epmf_net.py:97-102 / 138-143."""
    return F.pixel_shuffle(c.bn(po.lrelu(c.conv(x, p + ".0", padding=1)), p + ".2"), 2)


def salsanext_fusion(c, x, img_feats, p):
    """SalsaNextFusion.forward, epmf_net.py:104-131 -> (softmax map, down5c)."""
    d = sparse_context_block(c, x, p + ".downCntx")
    d = sparse_context_block(c, d, p + ".downCntx2")
    d = sparse_context_block(c, d, p + ".downCntx3", stride=2)
    d = po.fusion_block(c, d, img_feats[0], p + ".fusionblock_1")
    d0c, d0b = po.res_block(c, d, p + ".resBlock1", pooling=True, drop_out=False)
    d0c = po.fusion_block(c, d0c, img_feats[1], p + ".fusionblock_2")
    d1c, d1b = po.res_block(c, d0c, p + ".resBlock2")
    d1c = po.fusion_block(c, d1c, img_feats[2], p + ".fusionblock_3")
    d2c, d2b = po.res_block(c, d1c, p + ".resBlock3")
    d2c = po.fusion_block(c, d2c, img_feats[3], p + ".fusionblock_4")
    d3c, d3b = po.res_block(c, d2c, p + ".resBlock4")
    d5c = po.aspp(c, po.res_block(c, d3c, p + ".resBlock5", pooling=False), p + ".aspp")
    u4 = po.up_block(c, d5c, d3b, p + ".upBlock1")
    u3 = po.up_block(c, u4, d2b, p + ".upBlock2")
    u2 = po.up_block(c, u3, d1b, p + ".upBlock3")
    u1 = po.up_block(c, u2, d0b, p + ".upBlock4", drop_out=False)
    u1 = conv_lrelu_bn_shuffle(c, u1, p + ".extraUpSample")
    return F.softmax(c.conv(u1, p + ".logits"), dim=1), d5c


def rgb_decoder(c, feats, lidar_feature, p):
    """RGBDecoder.forward, epmf_net.py:175-183."""

    def stage(x, name, padding):
        x = c.bn(po.lrelu(c.conv(x, name + ".0", padding=padding)), name + ".2")
        return F.interpolate(x, scale_factor=2, mode="bilinear")

    fuse = torch.cat((conv_lrelu_bn_shuffle(c, lidar_feature, p + ".extraUpSample"), po.aspp(c, feats[3], p + ".aspp")), dim=1)
    u4 = stage(fuse, p + ".up_4a", 1)
    u3 = stage(torch.cat((u4, feats[2]), dim=1), p + ".up_3a", 1)
    u2 = stage(torch.cat((u3, feats[1]), dim=1), p + ".up_2a", 1)
    u1 = stage(torch.cat((u2, feats[0]), dim=1), p + ".up_1a", 0)
    return F.softmax(c.conv(u1, p + ".conv", padding=1), dim=1)


def epmf_forward(sd, pcd_feature, img_feature, backbone="resnet34", tf32=False):
    """EPMFNet.forward, epmf_net.py:209-216 (eval mode) -> (lidar_pred, camera_pred) softmax maps."""
    c = po.Ctx(sd, train=False, tf32=tf32)
    feats = po.resnet_encoder(c, img_feature, "camera_stream_encoder", backbone)
    lidar, lidar_feature = salsanext_fusion(c, pcd_feature, feats, "lidar_stream")
    camera = rgb_decoder(c, feats, lidar_feature, "camera_stream_decoder")
    return lidar, camera


def epmf_param_shapes(nclasses=20, base_channels=32, backbone="resnet34", pcd_channels=5, img_channels=3):
    """Ordered {state_dict key: shape} of EPMFNet: PMFNet's inventory with the sparse context
    blocks (convN.conv.{weight,bias} + convN.bias, conv1 3x3; epmf_net.py:13-22,55-62), the re-dimensioned fusion blocks
    (:91-94), lidar_stream.extraUpSample (:97-102) and the decoder's aspp / extraUpSample / wider up_4a (:138-150)."""
    base = po.pmf_param_shapes(nclasses, base_channels, backbone, pcd_channels, img_channels)
    exp = po.RESNET_EXPANSION[backbone]
    fch = [64 * exp, 128 * exp, 256 * exp, 512 * exp]
    b, dc = base_channels, 16 * exp
    shapes = {}

    def conv(name, cin, cout, kh, kw):
        shapes[name + ".weight"] = (cout, cin, kh, kw)
        shapes[name + ".bias"] = (cout,)

    def bn(name, ch):
        for k, shp in ((".weight", (ch,)), (".bias", (ch,)), (".running_mean", (ch,)), (".running_var", (ch,)),
                       (".num_batches_tracked", ())):
            shapes[name + k] = shp

    # reference order (module registration order; a re-assigned attribute keeps its first position):
    #   camera_stream_encoder | camera_stream_decoder {aspp, extraUpSample, up_4a..up_1a, conv} |
    #   lidar_stream {downCntx*, resBlock*, upBlock*, logits, fusionblock_*, aspp, extraUpSample}
    d, l = "camera_stream_decoder", "lidar_stream"
    for k, v in base.items():
        if k.startswith("camera_stream_encoder."):
            shapes[k] = v
    a = d + ".aspp"
    conv(a + ".conv", fch[3], fch[3], 1, 1)
    conv(a + ".atrous_block1", fch[3], fch[3], 1, 1)
    conv(a + ".atrous_block6", fch[3], fch[3], 3, 3)
    conv(a + ".atrous_block12", fch[3], fch[3], 3, 3)
    conv(a + ".atrous_block18", fch[3], fch[3], 3, 3)
    conv(a + ".conv_1x1_output", 5 * fch[3], fch[3], 1, 1)
    conv(d + ".extraUpSample.0", 8 * b, 8 * b, 3, 3); bn(d + ".extraUpSample.2", 8 * b)
    conv(d + ".up_4a.0", fch[3] + 2 * b, dc, 3, 3)
    for k, v in base.items():
        if k.startswith(d + ".") and not k.startswith(d + ".up_4a.0"):
            shapes[k] = v
    for name, cin in ((".downCntx", pcd_channels), (".downCntx2", b), (".downCntx3", b)):
        for cn, ci in ((".conv1", cin), (".conv2", b)):
            shapes[l + name + cn + ".bias"] = (b,)
            conv(l + name + cn + ".conv", ci, b, 3, 3)
        bn(l + name + ".bn1", b)
        shapes[l + name + ".conv3.bias"] = (b,)
        conv(l + name + ".conv3.conv", b, b, 3, 3)
        bn(l + name + ".bn2", b)
    for k, v in base.items():
        if k.startswith(l + ".resBlock") or k.startswith(l + ".upBlock") or k.startswith(l + ".logits"):
            shapes[k] = v
    for i, (pc, ic) in enumerate(((b, fch[0]), (2 * b, fch[1]), (4 * b, fch[2]), (8 * b, fch[3])), start=1):
        p = "%s.fusionblock_%d" % (l, i)
        conv(p + ".fuse_conv.0", pc + ic, pc, 3, 3); bn(p + ".fuse_conv.2", pc)
        conv(p + ".attention.0", pc, pc, 3, 3); bn(p + ".attention.1", pc)
        conv(p + ".attention.3", pc, pc, 3, 3); bn(p + ".attention.4", pc)
    for k, v in base.items():
        if k.startswith(l + ".aspp."):
            shapes[k] = v
    conv(l + ".extraUpSample.0", b, 4 * b, 3, 3); bn(l + ".extraUpSample.2", 4 * b)
    return shapes
