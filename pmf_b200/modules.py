"""nn.Module mirrors of the reference's hot-path modules, with the reference's constructor/forward signatures and
``state_dict`` keys, whose forward/backward run on libpmf_b200.so.

  PMFNet(pcd_channels=5, img_channels=3, nclasses=20, base_channels=32, imagenet_pretrained=True,
         image_backbone="resnet34").forward(pcd_feature, img_feature) -> (lidar_pred, camera_pred)   pmf_net.py:224-249
  ResidualBasedFusionBlock(pcd_channels, img_channels).forward(pcd_feature, img_feature) -> Tensor    pmf_net.py:10-36

The child modules (nn.Conv2d / nn.BatchNorm2d / nn.Dropout2d ...) are PARAMETER CONTAINERS created in the
reference's construction order (so a given torch seed yields the reference's default initialisation, and
optimisers, DDP, ``replaceBN`` and ``load_state_dict`` see the tree they expect).  Their own ``forward`` is never
called: the whole network runs as one autograd.Function over pmf_b200.engine.  There is no CPU path — calling a
module with CPU tensors raises.
"""
import contextlib
import gc
import os

import torch
import torch.nn as nn

from . import net as G
from . import _lib as _L
from .engine import Act, Engine, WeightCache


def _on_device(t):
    """Context manager making t's CUDA device current (launches take the device from the tensors, not from whatever the
    calling thread had selected).  A no-op for the CPU tensors of the numpy C-ABI model the CPU tests run on."""
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("pmf_b200 modules run on a B200 (sm_100a) only; got a %s tensor. There is no CPU fallback "
                               "(use the reference implementation for CPU runs)." % t.device)


# ------------------------------------------------------------------------------------------ containers (salsanext.py)
class ResContextBlock(nn.Module):
    """Parameters of salsanext.py:9-21."""

    def __init__(self, in_filters, out_filters):
        super().__init__()
        self.conv1 = nn.Conv2d(in_filters, out_filters, kernel_size=(1, 1), stride=1)
        self.act1 = nn.LeakyReLU()
        self.conv2 = nn.Conv2d(out_filters, out_filters, (3, 3), padding=1)
        self.act2 = nn.LeakyReLU()
        self.bn1 = nn.BatchNorm2d(out_filters)
        self.conv3 = nn.Conv2d(out_filters, out_filters, (3, 3), dilation=2, padding=2)
        self.act3 = nn.LeakyReLU()
        self.bn2 = nn.BatchNorm2d(out_filters)


class ResBlock(nn.Module):
    """Parameters of salsanext.py:38-67."""

    def __init__(self, in_filters, out_filters, dropout_rate, kernel_size=(3, 3), stride=1, pooling=True, drop_out=True):
        super().__init__()
        self.pooling = pooling
        self.drop_out = drop_out
        self.conv1 = nn.Conv2d(in_filters, out_filters, kernel_size=(1, 1), stride=stride)
        self.act1 = nn.LeakyReLU()
        self.conv2 = nn.Conv2d(in_filters, out_filters, kernel_size=(3, 3), padding=1)
        self.act2 = nn.LeakyReLU()
        self.bn1 = nn.BatchNorm2d(out_filters)
        self.conv3 = nn.Conv2d(out_filters, out_filters, kernel_size=(3, 3), dilation=2, padding=2)
        self.act3 = nn.LeakyReLU()
        self.bn2 = nn.BatchNorm2d(out_filters)
        self.conv4 = nn.Conv2d(out_filters, out_filters, kernel_size=(2, 2), dilation=2, padding=1)
        self.act4 = nn.LeakyReLU()
        self.bn3 = nn.BatchNorm2d(out_filters)
        self.conv5 = nn.Conv2d(out_filters * 3, out_filters, kernel_size=(1, 1))
        self.act5 = nn.LeakyReLU()
        self.bn4 = nn.BatchNorm2d(out_filters)
        self.dropout = nn.Dropout2d(p=dropout_rate)
        if pooling:
            self.pool = nn.AvgPool2d(kernel_size=kernel_size, stride=2, padding=1)


class UpBlock(nn.Module):
    """Parameters of salsanext.py:107-134."""

    def __init__(self, in_filters, out_filters, dropout_rate, drop_out=True):
        super().__init__()
        self.drop_out = drop_out
        self.in_filters = in_filters
        self.out_filters = out_filters
        self.dropout1 = nn.Dropout2d(p=dropout_rate)
        self.dropout2 = nn.Dropout2d(p=dropout_rate)
        self.conv1 = nn.Conv2d(in_filters // 4 + 2 * out_filters, out_filters, (3, 3), padding=1)
        self.act1 = nn.LeakyReLU()
        self.bn1 = nn.BatchNorm2d(out_filters)
        self.conv2 = nn.Conv2d(out_filters, out_filters, (3, 3), dilation=2, padding=2)
        self.act2 = nn.LeakyReLU()
        self.bn2 = nn.BatchNorm2d(out_filters)
        self.conv3 = nn.Conv2d(out_filters, out_filters, (2, 2), dilation=2, padding=1)
        self.act3 = nn.LeakyReLU()
        self.bn3 = nn.BatchNorm2d(out_filters)
        self.conv4 = nn.Conv2d(out_filters * 3, out_filters, kernel_size=(1, 1))
        self.act4 = nn.LeakyReLU()
        self.bn4 = nn.BatchNorm2d(out_filters)
        self.dropout3 = nn.Dropout2d(p=dropout_rate)


# ------------------------------------------------------------------------------------------ pmf_net.py
def _fusion_layers(mod, pcd_channels, img_channels):
    mod.fuse_conv = nn.Sequential(
        nn.Conv2d(pcd_channels + img_channels, pcd_channels, kernel_size=3, padding=1, stride=1),
        nn.LeakyReLU(),
        nn.BatchNorm2d(pcd_channels))
    mod.attention = nn.Sequential(
        nn.Conv2d(pcd_channels, pcd_channels, kernel_size=3, padding=1, stride=1),
        nn.BatchNorm2d(pcd_channels),
        nn.ReLU(inplace=True),
        nn.Conv2d(pcd_channels, pcd_channels, kernel_size=3, padding=1, stride=1),
        nn.BatchNorm2d(pcd_channels),
        nn.Sigmoid())


class _FusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, record, pcd, img, *params):
        E = Engine(G.ModuleParams(mod), pcd.device, mod.training, record, mod._cache, dropout=False)
        n, cp, h, w = pcd.shape
        ci = img.shape[1]
        cat = E.new(n, h, w, cp + ci, needs_grad=True)
        E.input_nchw(pcd, cp, out=cat.slice(0, cp))
        E.input_nchw(img, ci, out=cat.slice(cp, cp + ci))
        y = G.fusion_block(E, cat, cp, "", rnd=False)
        E.finish_forward()
        ctx.E, ctx.cat, ctx.y, ctx.cp, ctx.names = E, cat, y, cp, [n_ for n_, _ in mod.named_parameters()]
        return E.to_nchw(y.t, cp)

    @staticmethod
    def backward(ctx, dout):
        E, cat, y, cp = ctx.E, ctx.cat, ctx.y, ctx.cp
        E.st = torch.cuda.current_stream(dout.device).cuda_stream
        gy, _ = y.grad_target()
        E.input_nchw(dout, cp, out=Act(gy, needs_grad=False), rnd=False)
        grads = E.run_backward()
        g = cat.grad_read()
        dp = E.to_nchw(g[..., :cp], cp)
        di = E.to_nchw(g[..., cp:], g.shape[3] - cp)
        return (None, None, dp, di) + tuple(grads.get(nm) for nm in ctx.names)


class ResidualBasedFusionBlock(nn.Module):
    """pmf_net.py:10-36: out = f*sigmoid-gate(f) + pcd with f = BN(LeakyReLU(conv3x3(cat(pcd, img))))."""

    def __init__(self, pcd_channels, img_channels):
        super().__init__()
        _fusion_layers(self, pcd_channels, img_channels)
        self._cache = WeightCache()

    def forward(self, pcd_feature, img_feature):
        _require_cuda(pcd_feature, img_feature)
        if pcd_feature.shape[1] % 4 or img_feature.shape[1] % 4:
            raise RuntimeError("pmf_b200 ResidualBasedFusionBlock needs channel counts that are multiples of 4")
        params = [p for _, p in self.named_parameters()]
        record = torch.is_grad_enabled() and (pcd_feature.requires_grad or img_feature.requires_grad
                                              or any(p.requires_grad for p in params))
        return _FusionFn.apply(self, record, pcd_feature, img_feature, *params)


class ResNet(nn.Module):
    """Parameters of pmf_net.py:41-81 (torchvision blocks, stride-1 7x7 stem)."""

    def __init__(self, in_channels=3, backbone="resnet50", dropout_rate=0.2, pretrained=True):
        super().__init__()
        from torchvision.models.resnet import resnet34, resnet50, resnet101, resnet152
        ctors = {"resnet34": (resnet34, 1), "resnet50": (resnet50, 4), "resnet101": (resnet101, 4), "resnet152": (resnet152, 4)}
        if backbone not in ctors:
            raise NotImplementedError("invalid backbone: {}".format(backbone))
        ctor, self.expansion = ctors[backbone]
        net = ctor(pretrained)
        self.feature_channels = [64 * self.expansion, 128 * self.expansion, 256 * self.expansion, 512 * self.expansion]
        self.backbone_name = backbone
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=1, padding=3, bias=False)
        if in_channels == 3:
            self.conv1.weight.data = net.conv1.weight.data
        self.bn1 = net.bn1
        self.relu = net.relu
        self.maxpool = net.maxpool
        self.layer1 = net.layer1
        self.layer2 = net.layer2
        self.layer3 = net.layer3
        self.layer4 = net.layer4
        self.dropout = nn.Dropout2d(p=dropout_rate)


class ASPP(nn.Module):
    """Parameters of pmf_net.py:103-117."""

    def __init__(self, in_channel=512, depth=256):
        super().__init__()
        self.mean = nn.AdaptiveAvgPool2d((1, 1))
        self.conv = nn.Conv2d(in_channel, depth, 1, 1)
        self.atrous_block1 = nn.Conv2d(in_channel, depth, 1, 1)
        self.atrous_block6 = nn.Conv2d(in_channel, depth, 3, 1, padding=6, dilation=6)
        self.atrous_block12 = nn.Conv2d(in_channel, depth, 3, 1, padding=12, dilation=12)
        self.atrous_block18 = nn.Conv2d(in_channel, depth, 3, 1, padding=18, dilation=18)
        self.conv_1x1_output = nn.Conv2d(depth * 5, depth, 1, 1)


class SalsaNextFusion(nn.Module):
    """Parameters of salsanext.py:166-187 + pmf_net.py:141-151, created in the reference's order."""

    def __init__(self, in_channels=8, nclasses=20, base_channels=32, img_feature_channels=()):
        super().__init__()
        b = base_channels
        self.base_channels = b
        self.dropout_ratio = 0.2
        self.downCntx = ResContextBlock(in_channels, b)
        self.downCntx2 = ResContextBlock(b, b)
        self.downCntx3 = ResContextBlock(b, b)
        self.resBlock1 = ResBlock(b, 2 * b, self.dropout_ratio, pooling=True, drop_out=False)
        self.resBlock2 = ResBlock(2 * b, 4 * b, self.dropout_ratio, pooling=True)
        self.resBlock3 = ResBlock(4 * b, 8 * b, self.dropout_ratio, pooling=True)
        self.resBlock4 = ResBlock(8 * b, 8 * b, self.dropout_ratio, pooling=True)
        self.resBlock5 = ResBlock(8 * b, 8 * b, self.dropout_ratio, pooling=False)
        self.upBlock1 = UpBlock(8 * b, 4 * b, self.dropout_ratio)
        self.upBlock2 = UpBlock(4 * b, 4 * b, self.dropout_ratio)
        self.upBlock3 = UpBlock(4 * b, 2 * b, self.dropout_ratio)
        self.upBlock4 = UpBlock(2 * b, b, self.dropout_ratio, drop_out=False)
        self.logits = nn.Conv2d(b, nclasses, kernel_size=(1, 1))
        self.softmax = True
        self.fusionblock_1 = _FusionContainer(2 * b, img_feature_channels[0])
        self.fusionblock_2 = _FusionContainer(4 * b, img_feature_channels[1])
        self.fusionblock_3 = _FusionContainer(8 * b, img_feature_channels[2])
        self.fusionblock_4 = _FusionContainer(8 * b, img_feature_channels[3])
        self.aspp = ASPP(8 * b, 8 * b)


class _FusionContainer(nn.Module):
    def __init__(self, pcd_channels, img_channels):
        super().__init__()
        _fusion_layers(self, pcd_channels, img_channels)


class RGBDecoder(nn.Module):
    """Parameters of pmf_net.py:183-212."""

    def __init__(self, in_channels=(), nclasses=4, base_channels=64):
        super().__init__()

        def stage(cin, k, pad):
            return nn.Sequential(nn.Conv2d(cin, base_channels, k, padding=pad), nn.LeakyReLU(), nn.BatchNorm2d(base_channels),
                                 nn.Upsample(scale_factor=2, mode="bilinear"))

        self.up_4a = stage(in_channels[3], 3, 1)
        self.up_3a = stage(in_channels[2] + base_channels, 3, 1)
        self.up_2a = stage(in_channels[1] + base_channels, 3, 1)
        self.up_1a = stage(in_channels[0] + base_channels, 1, 0)
        self.conv = nn.Conv2d(base_channels, nclasses, kernel_size=3, padding=1)


class _PMFFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, record, pcd, img, *params):
        E = Engine(G.ModuleParams(mod), pcd.device, mod.training, record, mod._cache, dropout=mod._dropout_masks())
        lidar, camera, ll, cl = G.pmf_forward(E, pcd, img, mod.image_backbone, mod.nclasses)
        ctx.mod = mod
        if record:
            ctx.E, ctx.ll, ctx.cl = E, ll, cl
            ctx.names = [n for n, _ in mod.named_parameters()]
            ctx.save_for_backward(lidar, camera)
        else:
            ctx.E = None
        return lidar, camera

    @staticmethod
    def backward(ctx, d_lidar, d_camera):
        E = ctx.E
        if E is None:
            raise RuntimeError("pmf_b200: backward through a forward that was run without gradient recording")
        lidar, camera = ctx.saved_tensors
        with _on_device(lidar):  # autograd's worker thread: take the device from the tensors
            E.st = torch.cuda.current_stream(lidar.device).cuda_stream
            E.softmax_backward(ctx.ll, lidar, d_lidar)
            E.softmax_backward(ctx.cl, camera, d_camera)
            grads = E.run_backward()
            sync = getattr(ctx.mod, "_grad_sync", None)
            if sync is not None and sync.enabled and sync.world() > 1:  # eager pass (first call of a specialisation)
                sync.reduce_list([g for g in grads.values() if g is not None])
        ctx.E = None
        return (None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


@contextlib.contextmanager
def _capture(graph, pool):
    """torch.cuda.graph(...) with Python's cyclic garbage collector out of the way.  torch no longer collects before a
    capture (torch.compiler.config.force_cudagraph_gc), and a capture here creates thousands of Python objects: if the
    collector runs in the middle of it and frees a dead cycle that owns CUDA memory — a module whose graphs were just
    evicted or replaced, an earlier model of the same process — the cudaFree in the capturing thread invalidates the
    capture (cudaErrorStreamCaptureInvalidated at the next launch).  Collect first, keep the collector off until the graph
    is closed."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, pool=pool, capture_error_mode="thread_local"):  # NCCL watchdog threads may poll events
            yield
    finally:
        if was_enabled:
            gc.enable()


GRAPH_CACHE = max(1, int(os.environ.get("PMFB_GRAPH_CACHE", "8")))


def _evict_lru(graphs):
    """Captured specialisations (train / eval x input shape x parameter storage) are kept least-recently-used first: a
    trainer alternating a train and a validation shape keeps both graphs; only the oldest entry goes when the cache
    (PMFB_GRAPH_CACHE, default 8) is full."""
    while len(graphs) >= GRAPH_CACHE:
        graphs.pop(next(iter(graphs)))


class _GraphedPMF:
    """One (shape, mode, parameter-storage) specialisation of PMFNet captured into CUDA graphs.

    The executor issues ~4000 C-ABI launches per training step; replaying them from two CUDA graphs (forward,
    backward) removes the Python/ctypes time from the step.  Only three things stay outside the graphs because they
    touch caller-owned memory whose address changes every call: packing the NCHW inputs, reading d(probabilities)
    in the softmax backward, and cloning the outputs / parameter gradients handed back to autograd.
    Weight packing, BatchNorm running-statistic updates and Dropout2d mask draws are graph nodes, so every replay
    sees the live parameters, updates the live buffers and draws fresh masks (torch's graph-safe Philox)."""

    def __init__(self, mod, pcd, img, record):
        self.mod = mod
        self.record = record
        self.dev = pcd.device
        n, c_pcd, h, w = pcd.shape
        self.shape = (n, c_pcd, h, w)
        frozen = bool(getattr(mod, "_frozen", False)) and not record and not mod.training
        # frozen inference: the eager pass that preceded this capture packed the weights and folded the BatchNorm affines
        # into mod._cache; the captured graph re-uses them and contains no packing / finalisation kernels
        self.cache = mod._cache if frozen else WeightCache(always=True)
        self.cache.precise = _L.get_precision() == "3xtf32"
        self.cache.h16 = _L.get_precision() == "f16" and mod.training
        self.E = None
        self.version = 0
        self.bwd_captured = False
        self.pool = torch.cuda.graph_pool_handle()
        self.g_fwd = torch.cuda.CUDAGraph()
        self.g_bwd = torch.cuda.CUDAGraph() if record else None
        self.names = [nm for nm, _ in mod.named_parameters()]
        # static packed inputs (filled eagerly before every replay)
        self.E_in = Engine(G.ModuleParams(mod), self.dev, mod.training, False, WeightCache(), dropout=False)
        self.img7 = self.E_in.new(n, h, w, 32, needs_grad=False)
        self.pcd = self.E_in.new(n, h, w, (c_pcd + 3) // 4 * 4, needs_grad=False)
        self._stage_inputs(pcd, img)
        if not frozen:
            self.cache.build_table(G.ModuleParams(mod), record, self.dev)  # every weight packed by one launch per pass
        torch.cuda.synchronize(self.dev)
        l0 = _L.launches
        with _capture(self.g_fwd, self.pool):
            E = Engine(G.ModuleParams(mod), self.dev, mod.training, record, self.cache, dropout=mod._dropout_masks())
            self.lidar, self.camera, self.ll, self.cl = G.pmf_forward_packed(E, self.pcd, self.img7, mod.image_backbone,
                                                                             mod.nclasses)
        self.n_fwd_calls = _L.launches - l0  # C-ABI launches one replay of the forward graph stands for
        _L.launches = l0                     # capturing enqueues nothing
        self.E = E
        self.ll_grad = self.cl_grad = None
        # ---- backward segments: K graphs / autograd nodes (DDP's bucketed all-reduce of segment k overlaps segment k+1)
        self.segs = []
        if record:
            k = max(1, int(os.environ.get("PMFB_BWD_SEGMENTS", "4")))
            plan = E.plan_segments(k)
            produced = set(nm for _fs, names in plan for nm in names)
            params = dict(mod.named_parameters())
            # all parameter gradients live in ONE flat buffer laid out segment by segment: one clone per segment hands
            # them to autograd
            offs, total = {}, 0
            for si, (_fs, names) in enumerate(plan):
                start = total
                seg_names = [nm for nm in names if nm in params]
                for nm in seg_names:
                    p = params[nm]
                    offs[nm] = (total, p.numel(), tuple(p.shape))
                    total += (p.numel() + 3) // 4 * 4
                self.segs.append(dict(names=seg_names, lo=start, hi=total, graph=torch.cuda.CUDAGraph(), captured=False, calls=0))
            self.unused = [nm for nm in params if nm not in produced]
            self.flat = torch.zeros(max(total, 4), device=self.dev, dtype=torch.float32)
            self.offs = offs
            E.flat_views = {nm: self.flat[o:o + n_].view(shp) for nm, (o, n_, shp) in offs.items()}
            self.prepared = False

    def _stage_inputs(self, pcd, img):
        self.E_in.st = torch.cuda.current_stream(self.dev).cuda_stream
        self.E_in.input_nchw(img, 32, n_shift=7, out=self.img7)
        self.E_in.input_nchw(pcd, self.pcd.c, out=self.pcd)

    def forward(self, pcd, img):
        self._stage_inputs(pcd, img)
        self.g_fwd.replay()
        _L.launches += self.n_fwd_calls
        self.version += 1
        return self.lidar.clone(), self.camera.clone()

    def backward_segment(self, k, d_lidar=None, d_camera=None):
        """Replays (capturing it the first time) backward segment k; returns {parameter name: gradient}."""
        E = self.E
        seg = self.segs[k]
        if k == 0:
            st = torch.cuda.current_stream(self.dev).cuda_stream
            if self.ll_grad is None:
                self.ll_grad, _ = self.ll.grad_target()
                self.cl_grad, _ = self.cl.grad_target()
            E.softmax_backward_into(self.ll_grad, self.lidar, d_lidar, stream=st)
            E.softmax_backward_into(self.cl_grad, self.camera, d_camera, stream=st)
            if not self.prepared:
                E.prepare_backward()
                self.prepared = True
        if not seg["captured"]:
            torch.cuda.synchronize(self.dev)
            l0 = _L.launches
            with _capture(seg["graph"], self.pool):
                E.run_backward_segment(k)
            seg["calls"] = _L.launches - l0
            _L.launches = l0
            seg["captured"] = True
        seg["graph"].replay()
        _L.launches += seg["calls"]
        out = self.flat[seg["lo"]:seg["hi"]].clone()
        lo = seg["lo"]
        grads = {nm: out[self.offs[nm][0] - lo:self.offs[nm][0] - lo + self.offs[nm][1]].view(self.offs[nm][2]) for nm in seg["names"]}
        sync = getattr(self.mod, "_grad_sync", None)
        if sync is not None and sync.enabled and sync.world() > 1:
            # pmf_b200.dist.FrameParallel: the segment's slice is averaged over the ranks asynchronously (it overlaps the
            # next segments) and handed to the parameters directly, so autograd never reads it before the reduction is done
            params = dict(self.mod.named_parameters())
            plist = [params[nm] for nm in seg["names"] if params[nm].requires_grad]
            sync.reduce(out, plist)
            for nm in seg["names"]:
                p = params[nm]
                if not p.requires_grad:
                    continue
                if p.grad is None:
                    p.grad = grads[nm]
                else:
                    p.grad.add_(grads[nm])
            return {nm: None for nm in seg["names"]}
        return grads


class _PMFGraphHeadFn(torch.autograd.Function):
    """Last node of the forward chain = FIRST backward segment: owns the module outputs."""

    @staticmethod
    def forward(ctx, runner, pcd, img, token, *params):
        lidar, camera = runner.forward(pcd, img)
        ctx.runner, ctx.version, ctx.has_token = runner, runner.version, token is not None
        return lidar, camera

    @staticmethod
    def backward(ctx, d_lidar, d_camera):
        r = ctx.runner
        if not r.record:
            raise RuntimeError("pmf_b200: backward through a forward that was run without gradient recording")
        if ctx.version != r.version:
            raise RuntimeError("pmf_b200: this forward's activations were overwritten by a later forward of the same "
                               "CUDA-graph specialisation; set PMFB_CUDA_GRAPH=0 for call patterns other than "
                               "forward -> backward")
        with _on_device(d_lidar):
            g = r.backward_segment(0, d_lidar, d_camera)
            tok = torch.zeros(1, device=r.dev) if ctx.has_token else None
        return (None, None, None, tok) + tuple(g[nm] for nm in r.segs[0]["names"])


class _PMFGraphSegFn(torch.autograd.Function):
    """Backward segment k >= 1.  In forward it only threads a token through, so that autograd runs the segments in
    order; every node owns the parameters whose gradients its segment produces."""

    @staticmethod
    def forward(ctx, runner, k, token, *params):
        ctx.runner, ctx.k, ctx.has_token = runner, k, token is not None
        return torch.zeros(1, device=runner.dev)

    @staticmethod
    def backward(ctx, d_tok):
        r = ctx.runner
        with _on_device(d_tok):
            g = r.backward_segment(ctx.k)
            tok = torch.zeros(1, device=r.dev) if ctx.has_token else None
        return (None, None, tok) + tuple(g[nm] for nm in r.segs[ctx.k]["names"])


def _graph_apply(runner, pcd, img, named):
    """Chains the segment nodes: the LAST backward segment is applied first.  Parameters no segment produces a gradient
    for (none in PMFNet) ride on the head node and receive None."""
    if not runner.record or len(runner.segs) <= 1:
        names = runner.segs[0]["names"] if runner.segs else []
        return _PMFGraphHeadFn.apply(runner, pcd, img, None, *[named[nm] for nm in names])
    token = None
    for k in range(len(runner.segs) - 1, 0, -1):
        token = _PMFGraphSegFn.apply(runner, k, token, *[named[nm] for nm in runner.segs[k]["names"]])
    return _PMFGraphHeadFn.apply(runner, pcd, img, token, *[named[nm] for nm in runner.segs[0]["names"]])


class PMFNet(nn.Module):
    """pmf_net.py:224-249.  forward returns the two softmax probability maps (B, nclasses, H, W)."""

    def __init__(self, pcd_channels=5, img_channels=3, nclasses=20, base_channels=32, imagenet_pretrained=True,
                 image_backbone="resnet34"):
        super().__init__()
        self.camera_stream_encoder = ResNet(in_channels=img_channels, pretrained=imagenet_pretrained, backbone=image_backbone)
        self.camera_stream_decoder = RGBDecoder(self.camera_stream_encoder.feature_channels, nclasses=nclasses,
                                                base_channels=self.camera_stream_encoder.expansion * 16)
        self.lidar_stream = SalsaNextFusion(in_channels=pcd_channels, nclasses=nclasses, base_channels=base_channels,
                                            img_feature_channels=self.camera_stream_encoder.feature_channels)
        self.nclasses = nclasses
        self.image_backbone = image_backbone
        self._cache = WeightCache()
        self._dropout_override = None  # tests: False (off) or {site: (N,C) mask}
        self._graphs = {}
        self._seen = set()

    def freeze(self, frozen=True):
        """Inference with FIXED parameters: from the next eval forward on, the conv weights are packed
        once and the eval-mode BatchNorm layers folded once into per-channel epilogue vectors — conv -> BN -> ReLU folds in
        front of the activation, the SalsaNext / fusion order conv -> LeakyReLU -> BN behind it — and every later forward
        (CUDA-graph replay) runs the convolutions with their fused epilogues only.  Later changes to the parameters are
        NOT seen until ``freeze(False)`` / ``freeze()`` is called again.  See pmf_b200.export for the on-disk form."""
        self._frozen = bool(frozen)
        self._cache = WeightCache()
        self._cache.frozen = self._frozen
        self._graphs.clear()
        self._seen.clear()
        return self

    def _dropout_masks(self):
        if self._dropout_override is not None:
            return self._dropout_override
        if not self.training:
            return False
        # honour per-module .eval() on the Dropout2d children (sites whose module is in eval mode are skipped)
        return _DropoutSites(self)

    def _check_modes(self):
        """BatchNorm mode is taken from the root module: a tree whose BatchNorm children disagree with it (a frozen-BN
        backbone) would silently normalise with the wrong statistics, so it is refused."""
        for n, m in self.named_modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.training != self.training:
                raise NotImplementedError("pmf_b200.PMFNet: BatchNorm module %r is in %s mode but the model is in %s mode; "
                                          "per-module BatchNorm modes are not supported" %
                                          (n, "train" if m.training else "eval", "train" if self.training else "eval"))

    def _graph_key(self, pcd, img, record, params):
        drop = tuple((n, m.training, m.p) for n, m in self.named_modules() if isinstance(m, nn.Dropout2d))
        ptrs = tuple(p.data_ptr() for p in params) + tuple(b.data_ptr() for b in self.buffers())
        return (tuple(pcd.shape), tuple(img.shape), str(pcd.device), self.training, record, drop, ptrs,
                tuple(p.requires_grad for p in params), _L.get_precision(), bool(getattr(self, "_frozen", False)))

    def forward(self, pcd_feature, img_feature):
        _require_cuda(pcd_feature, img_feature)
        G.check_input_size(img_feature)
        params = [p for _, p in self.named_parameters()]
        if not params or getattr(self, "_is_replica", False):
            # nn.DataParallel replicas (trainer.py:41-47, the reference's legacy non-distributed branch) carry no
            # Parameters and would share this module's caches across threads
            raise RuntimeError("pmf_b200.PMFNet does not support nn.DataParallel; launch one process per GPU and wrap the "
                               "model in DistributedDataParallel (tasks/pmf/trainer.py:34-39)")
        self._check_modes()
        with _on_device(pcd_feature):
            return self._forward(pcd_feature, img_feature, params)

    def _forward(self, pcd_feature, img_feature, params):
        record = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        use_graph = (os.environ.get("PMFB_CUDA_GRAPH", "1") != "0" and self._dropout_override is None
                     and not torch.cuda.is_current_stream_capturing())
        if use_graph:
            key = self._graph_key(pcd_feature, img_feature, record, params)
            runner = self._graphs.get(key)
            if runner is None and key in self._seen:  # second call with this specialisation: capture it
                _evict_lru(self._graphs)
                torch.cuda.empty_cache()
                runner = self._graphs[key] = _GraphedPMF(self, pcd_feature, img_feature, record)
            self._seen.add(key)
            if runner is not None:
                self._graphs[key] = self._graphs.pop(key)  # most recently used last
                return _graph_apply(runner, pcd_feature, img_feature, dict(self.named_parameters()))
        return _PMFFn.apply(self, record, pcd_feature, img_feature, *params)


# ------------------------------------------------------------------------------------------ epmf_net.py (inference)
class SparseVariantConv(nn.Module):
    """Parameters of epmf_net.py:10-28 (``conv`` + separate ``bias``; kaiming-normal fan_out init at :24-28)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, groups=1, dilation=1, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, stride=stride, groups=groups,
                              dilation=dilation)
        self.pool = nn.MaxPool2d(kernel_size, stride=stride, padding=0, dilation=dilation)
        self.bias = nn.Parameter(torch.zeros(out_channels).float()) if bias else None
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_out", nonlinearity="leaky_relu")


class SparseResContextBlock(nn.Module):
    """Parameters of epmf_net.py:52-64."""

    def __init__(self, in_filters, out_filters, stride=1):
        super().__init__()
        self.conv1 = SparseVariantConv(in_filters, out_filters, 3, padding=1, stride=stride)
        self.act1 = nn.LeakyReLU()
        self.conv2 = SparseVariantConv(out_filters, out_filters, (3, 3), padding=(1, 1))
        self.act2 = nn.LeakyReLU()
        self.bn1 = nn.BatchNorm2d(out_filters)
        self.conv3 = SparseVariantConv(out_filters, out_filters, (3, 3), padding=(2, 2), dilation=2)
        self.act3 = nn.LeakyReLU()
        self.bn2 = nn.BatchNorm2d(out_filters)


def _conv_lrelu_bn_shuffle(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.LeakyReLU(), nn.BatchNorm2d(cout), nn.PixelShuffle(2))


class EPMFSalsaNextFusion(nn.Module):
    """Parameters of epmf_net.py:84-102, created in the reference's order: SalsaNext.__init__ (salsanext.py:166-187,
    including its three dense context blocks, which the subclass then REPLACES), then the sparse context blocks, the
    fusion blocks, ASPP and extraUpSample."""

    def __init__(self, in_channels=8, nclasses=20, base_channels=32, img_feature_channels=()):
        super().__init__()
        b = base_channels
        self.base_channels = b
        self.dropout_ratio = 0.2
        self.downCntx = ResContextBlock(in_channels, b)
        self.downCntx2 = ResContextBlock(b, b)
        self.downCntx3 = ResContextBlock(b, b)
        self.resBlock1 = ResBlock(b, 2 * b, self.dropout_ratio, pooling=True, drop_out=False)
        self.resBlock2 = ResBlock(2 * b, 4 * b, self.dropout_ratio, pooling=True)
        self.resBlock3 = ResBlock(4 * b, 8 * b, self.dropout_ratio, pooling=True)
        self.resBlock4 = ResBlock(8 * b, 8 * b, self.dropout_ratio, pooling=True)
        self.resBlock5 = ResBlock(8 * b, 8 * b, self.dropout_ratio, pooling=False)
        self.upBlock1 = UpBlock(8 * b, 4 * b, self.dropout_ratio)
        self.upBlock2 = UpBlock(4 * b, 4 * b, self.dropout_ratio)
        self.upBlock3 = UpBlock(4 * b, 2 * b, self.dropout_ratio)
        self.upBlock4 = UpBlock(2 * b, b, self.dropout_ratio, drop_out=False)
        self.logits = nn.Conv2d(b, nclasses, kernel_size=(1, 1))
        self.softmax = True
        self.downCntx = SparseResContextBlock(in_channels, b)
        self.downCntx2 = SparseResContextBlock(b, b)
        self.downCntx3 = SparseResContextBlock(b, b, stride=2)
        self.fusionblock_1 = _FusionContainer(b, img_feature_channels[0])
        self.fusionblock_2 = _FusionContainer(2 * b, img_feature_channels[1])
        self.fusionblock_3 = _FusionContainer(4 * b, img_feature_channels[2])
        self.fusionblock_4 = _FusionContainer(8 * b, img_feature_channels[3])
        self.aspp = ASPP(8 * b, 8 * b)
        self.extraUpSample = _conv_lrelu_bn_shuffle(b, 4 * b)


class EPMFRGBDecoder(nn.Module):
    """Parameters of epmf_net.py:134-173."""

    def __init__(self, in_channels=(), nclasses=4, base_channels=64, lidar_base_channels=32):
        super().__init__()

        def stage(cin, k, pad):
            return nn.Sequential(nn.Conv2d(cin, base_channels, k, padding=pad), nn.LeakyReLU(), nn.BatchNorm2d(base_channels),
                                 nn.Upsample(scale_factor=2, mode="bilinear"))

        self.aspp = ASPP(in_channels[3], in_channels[3])
        self.extraUpSample = _conv_lrelu_bn_shuffle(lidar_base_channels * 8, lidar_base_channels * 8)
        self.up_4a = stage(in_channels[3] + lidar_base_channels * 2, 3, 1)
        self.up_3a = stage(in_channels[2] + base_channels, 3, 1)
        self.up_2a = stage(in_channels[1] + base_channels, 3, 1)
        self.up_1a = stage(in_channels[0] + base_channels, 1, 0)
        self.conv = nn.Conv2d(base_channels, nclasses, kernel_size=3, padding=1)


class _GraphedEPMF:
    """One (shape, parameter-storage) specialisation of the EPMF eval forward captured into a CUDA graph."""

    def __init__(self, mod, pcd, img):
        self.dev = pcd.device
        n, c_pcd, h, w = pcd.shape
        self.cache = WeightCache(always=True)
        self.cache.precise = _L.get_precision() == "3xtf32"
        self.pool = torch.cuda.graph_pool_handle()
        self.g = torch.cuda.CUDAGraph()
        self.E_in = Engine(G.ModuleParams(mod), self.dev, False, False, WeightCache(), dropout=False)
        self.img7 = self.E_in.new(n, h, w, 32, needs_grad=False)
        self.pcd = self.E_in.new(n, h, w, (c_pcd + 3) // 4 * 4, needs_grad=False)
        self._stage_inputs(pcd, img)
        self.cache.build_table(G.ModuleParams(mod), False, self.dev)
        torch.cuda.synchronize(self.dev)
        l0 = _L.launches
        with _capture(self.g, self.pool):
            E = Engine(G.ModuleParams(mod), self.dev, False, False, self.cache, dropout=False)
            self.lidar, self.camera, _, _ = G.epmf_forward_packed(E, self.pcd, self.img7, mod.image_backbone, mod.nclasses)
        self.n_calls = _L.launches - l0
        _L.launches = l0
        self.E = E

    def _stage_inputs(self, pcd, img):
        self.E_in.st = torch.cuda.current_stream(self.dev).cuda_stream
        self.E_in.input_nchw(img, 32, n_shift=7, out=self.img7)
        self.E_in.input_nchw(pcd, self.pcd.c, out=self.pcd)

    def forward(self, pcd, img):
        self._stage_inputs(pcd, img)
        self.g.replay()
        _L.launches += self.n_calls
        return self.lidar.clone(), self.camera.clone()


class EPMFNet(nn.Module):
    """epmf_net.py:185-216, INFERENCE ONLY in this round: forward under ``torch.no_grad()`` / ``.eval()`` returns the two
    softmax probability maps (B, nclasses, H, W); H and W must be multiples of 32.  A forward that would have to
    record gradients raises NotImplementedError (the sparse convolution has no backward yet)."""

    def __init__(self, pcd_channels=5, img_channels=3, nclasses=20, base_channels=32, imagenet_pretrained=True,
                 image_backbone="resnet34"):
        super().__init__()
        if "resnet" not in image_backbone:
            raise NotImplementedError(image_backbone)
        self.camera_stream_encoder = ResNet(in_channels=img_channels, pretrained=imagenet_pretrained, backbone=image_backbone)
        self.camera_stream_decoder = EPMFRGBDecoder(self.camera_stream_encoder.feature_channels, nclasses=nclasses,
                                                    base_channels=self.camera_stream_encoder.expansion * 16,
                                                    lidar_base_channels=base_channels)
        self.lidar_stream = EPMFSalsaNextFusion(in_channels=pcd_channels, nclasses=nclasses, base_channels=base_channels,
                                                img_feature_channels=self.camera_stream_encoder.feature_channels)
        self.nclasses = nclasses
        self.image_backbone = image_backbone
        self._cache = WeightCache()
        self._graphs = {}
        self._seen = set()

    def forward(self, pcd_feature, img_feature):
        _require_cuda(pcd_feature, img_feature)
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("pmf_b200.EPMFNet is inference-only in this round: call .eval() and run the forward "
                                      "under torch.no_grad()")
        h, w = img_feature.shape[2], img_feature.shape[3]
        if h % 32 != 0 or w % 32 != 0:
            assert False, "invalid input size: {}".format(img_feature.shape)
        use_graph = os.environ.get("PMFB_CUDA_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing()
        if use_graph:
            key = (tuple(pcd_feature.shape), tuple(img_feature.shape), str(pcd_feature.device), _L.get_precision(),
                   tuple(p.data_ptr() for p in self.parameters()) + tuple(b.data_ptr() for b in self.buffers()))
            runner = self._graphs.get(key)
            if runner is None and key in self._seen:
                _evict_lru(self._graphs)
                torch.cuda.empty_cache()
                runner = self._graphs[key] = _GraphedEPMF(self, pcd_feature, img_feature)
            self._seen.add(key)
            if runner is not None:
                self._graphs[key] = self._graphs.pop(key)  # most recently used last
                return runner.forward(pcd_feature, img_feature)
        E = Engine(G.ModuleParams(self), pcd_feature.device, False, False, self._cache, dropout=False)
        lidar, camera, _, _ = G.epmf_forward(E, pcd_feature, img_feature, self.image_backbone, self.nclasses)
        return lidar, camera


class _DropoutSites(dict):
    """Lazily answers Engine.mask_for(site): draws a fresh Dropout2d scale when the site's module is training."""

    def __init__(self, model):
        super().__init__()
        self.mods = dict(model.named_modules())

    def draw(self, site, n, c, device):
        name = site
        if ".dropout.layer" in site:  # camera_stream_encoder.dropout.layer3 -> camera_stream_encoder.dropout
            name = site[:site.index(".dropout.layer")] + ".dropout"
        mod = self.mods.get(name)
        if mod is None or not mod.training or mod.p <= 0.0:
            return None
        keep = 1.0 - float(mod.p)
        return torch.empty((n, c), device=device, dtype=torch.float32).bernoulli_(keep).div_(keep)
