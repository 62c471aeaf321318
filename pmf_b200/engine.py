"""Host-side executor for the PMF hot path: NHWC fp32 activations in HBM, every device op a C-ABI call into
libpmf_b200.so (no torch compute kernels on the path; torch only owns the memory and the stream).

The executor keeps its own backward tape (one closure per fused op) instead of using torch.autograd per op:
gradients are written straight into channel slices of shared concat buffers (``torch.cat`` never materialises),
accumulation is folded into the producing kernel's epilogue, and tf32 rounding of conv operands happens where a
tensor is produced.  ``pmf_b200.modules`` wraps one whole forward/backward in a single autograd.Function.

Reference semantics (paths relative to the reference tree):
  conv_act      nn.Conv2d -> LeakyReLU / identity          salsanext.py:24-25,70-71; pmf_net.py:106-117; salsanext.py:186
  conv_act_bn   nn.Conv2d -> LeakyReLU -> BatchNorm2d       salsanext.py:27-33,73-90,145-160; pmf_net.py:13-18,187-209
  conv_bn       nn.Conv2d -> BatchNorm2d -> ReLU/Sigmoid    torchvision BasicBlock/Bottleneck; pmf_net.py:20-29,35
"""
import contextlib
import ctypes as C
import math
import os

import torch

from . import _lib as L
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, BnFuse, ConvDesc, Epilogue, TmaSrc, View, WeightJob, WgradDesc

BN_EPS_DEFAULT = 1e-5
# forward / dgrad: 16-bit operands from one full 64-channel slab on.  (32-channel layers CAN run on 64-byte operand rows,
# SWIZZLE_64B, with PMFB_H16_MIN_K=32 -- parity-green but measured slower than their tf32 path: 0.30 vs 0.24 ms for the
# full-resolution 32->32 3x3 layers, which are not MMA-bound.)
H16_MIN_K = int(os.environ.get("PMFB_H16_MIN_K", "64"))
H16_WGRAD_MIN = 64                                            # bf16 wgrad: one full 64-channel block per operand
H16_WGRAD = os.environ.get("PMFB_H16_WGRAD", "1") != "0"
# "f16" mode: the pre-BatchNorm activation of a training-mode conv -> [LeakyReLU] -> BN layer is STORED as fp16 by the conv
# epilogue (pmfb_conv_desc.out_half); only the BatchNorm passes read it (apply, backward reduce, backward apply)
PRE_BN_HALF = os.environ.get("PMFB_PRE_BN_HALF", "1") != "0"
# ... and the BatchNorm finalisation of such a layer runs inside its BN-apply launch (pmfb_pointwise16_bn)
BN_FUSE_FINALIZE = os.environ.get("PMFB_BN_FUSE_FINALIZE", "1") != "0"
# forward: the camera stream (encoder + decoder) is enqueued on the auxiliary stream, concurrently with the LiDAR stream
FWD_BRANCH = os.environ.get("PMFB_FWD_BRANCH", "1") != "0"
# backward: the camera stream's closures on a second auxiliary stream next to the LiDAR stream's (Engine._backward_order).
# OFF by default: correct (tests/test_gpu_parity.py::test_backward_branch_order_matches_plain_order) but measured 0.8 ms
# SLOWER at B=8, 480x640 — the backward is already saturated by the weight-gradient overlap, a third chain only adds
# contention (kernel durations sum to 77 ms instead of 70 ms over the same 51 ms step).
BWD_BRANCH = os.environ.get("PMFB_BWD_BRANCH", "0") != "0"


def _rup(x, m):
    return (x + m - 1) // m * m


def _view(t):
    """ctypes View of an NHWC torch view (N,H,W,C) with channel stride 1, or None -> null view."""
    v = View()
    if t is None:
        v.ptr, v.sn, v.sy, v.sx = None, 0, 0, 0
        return v
    assert t.dim() == 4 and (t.shape[3] == 1 or t.stride(3) == 1), (t.shape, t.stride())
    v.ptr, v.sn, v.sy, v.sx = t.data_ptr(), t.stride(0), t.stride(1), t.stride(2)
    v._keep = t
    return v


def _chan_view(m):
    """(N,C) per-image channel scale (Dropout2d mask) as a broadcasting NHWC view."""
    v = View()
    v.ptr, v.sn, v.sy, v.sx = m.data_ptr(), m.stride(0), 0, 0
    v._keep = m  # the View travels into backward closures: keep the mask's storage alive with it
    return v


def _p(t):
    return None if t is None else t.data_ptr()


class Act:
    """An NHWC activation: ``t`` is a (N,H,W,C) torch view; slices share the root's gradient buffer."""
    __slots__ = ("t", "root", "c0", "needs_grad", "_round_grad", "grad", "gcov", "h", "hcov", "hb", "hbcov", "ready")

    def __init__(self, t, root=None, c0=0, needs_grad=True):
        self.t = t
        self.root = root if root is not None else self
        self.c0 = c0
        self.needs_grad = needs_grad
        self._round_grad = False  # root only; see round_grad
        self.grad = None         # root only: dense NHWC gradient buffer
        self.gcov = []           # root only: channel intervals of grad already written
        self.h = None            # root only ("f16" mode): fp16 shadow of t, the operand of the kind::f16 convolutions
        self.hcov = []           # root only: channel intervals of the shadow that hold current values
        self.hb = None           # root only ("f16" mode, recording): bf16 shadow of t, the x operand of the bf16 wgrad
        self.hbcov = []
        self.ready = None        # root only: event recorded when a forward branch (Engine.branch) finished producing t

    # ---- 16-bit shadow ("f16" precision mode)
    def shadow(self, bf16=False):
        """fp16 (or bf16) view matching self.t, or None when this buffer has no such shadow."""
        buf = self.root.hb if bf16 else self.root.h
        return None if buf is None else buf[..., self.c0:self.c0 + self.c]

    def shadow_valid(self, bf16=False):
        a, b = self.c0, self.c0 + self.c
        return any(x <= a and b <= y for (x, y) in (self.root.hbcov if bf16 else self.root.hcov))

    def shadow_mark(self, bf16=False):
        a, b = self.c0, self.c0 + self.c
        r = self.root
        iv = sorted((r.hbcov if bf16 else r.hcov) + [(a, b)])
        merged = [iv[0]]
        for (x, y) in iv[1:]:
            if x <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(merged[-1][1], y))
            else:
                merged.append((x, y))
        if bf16:
            r.hbcov = merged
        else:
            r.hcov = merged

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def c(self):
        return self.t.shape[3]

    @property
    def round_grad(self):
        """True when some producer's backward feeds this buffer's gradient straight into a tensor-core conv:
        every writer of the gradient then stores tf32-rounded values."""
        return self.root._round_grad

    @round_grad.setter
    def round_grad(self, v):
        self.root._round_grad = bool(v)

    def slice(self, a, b):
        s = Act(self.t[..., a:b], root=self.root, c0=self.c0 + a, needs_grad=self.root.needs_grad)
        return s

    # ---- gradient bookkeeping -------------------------------------------------------------------------
    def _cov_state(self, a, b):
        covered = 0
        for (x, y) in self.root.gcov:
            lo, hi = max(a, x), min(b, y)
            if hi > lo:
                covered += hi - lo
        return "full" if covered == b - a else ("none" if covered == 0 else "partial")

    def _cov_add(self, a, b):
        iv = sorted(self.root.gcov + [(a, b)])
        merged = [iv[0]]
        for (x, y) in iv[1:]:
            if x <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(merged[-1][1], y))
            else:
                merged.append((x, y))
        self.root.gcov = merged

    def grad_target(self):
        """(grad view, accumulate?) for a writer of this activation's gradient."""
        r = self.root
        if r.grad is None:
            r.grad = torch.empty(r.t.shape, device=r.t.device, dtype=torch.float32)
        a, b = self.c0, self.c0 + self.c
        st = self._cov_state(a, b)
        g = r.grad[..., a:b]
        if st == "partial":  # rare: zero the unwritten channels, then accumulate everywhere
            pos = a
            for (x, y) in r.gcov + [(b, b)]:
                lo, hi = max(pos, a), min(x, b)
                if hi > lo:
                    L.call("pmfb_pointwise", None, r.grad[..., lo:hi].data_ptr(), r.grad.stride(0), r.grad.stride(1),
                           r.grad.stride(2), r.t.shape[0], r.t.shape[1], r.t.shape[2], hi - lo, C.byref(Epilogue()),
                           torch.cuda.current_stream().cuda_stream)
                pos = max(pos, y)
            self._cov_add(a, b)
            return g, True
        if st == "none":
            self._cov_add(a, b)
            return g, False
        return g, True

    def grad_read(self):
        r = self.root
        a, b = self.c0, self.c0 + self.c
        if r.grad is None or self._cov_state(a, b) != "full":
            raise RuntimeError("gradient of an activation was requested before every consumer produced it")
        return r.grad[..., a:b]


class ConvParam:
    """Geometry + live parameter tensors of one nn.Conv2d."""
    __slots__ = ("name", "weight", "bias", "c_out", "c_in", "kh", "kw", "dil", "pad", "stride", "stem", "c_out_p",
                 "c_in_p")

    def __init__(self, name, weight, bias, dil, pad, stride, stem=False):
        self.name = name
        self.weight, self.bias = weight, bias
        self.c_out, self.c_in, self.kh, self.kw = weight.shape
        self.dil, self.pad, self.stride, self.stem = dil, pad, stride, stem
        self.c_out_p = _rup(self.c_out, 4)
        self.c_in_p = 32 if stem else _rup(self.c_in, 4)

    @property
    def taps(self):
        return self.kh if self.stem else self.kh * self.kw

    def fwd_taps(self):
        """[(dc, dw, dp, dh, weight_tap)] of the forward implicit GEMM."""
        out = []
        if self.stem:
            for i in range(self.kh):
                out.append((0, 0, 0, i - self.pad, i))
            return out
        for i in range(self.kh):
            for j in range(self.kw):
                dh, dw = i * self.dil - self.pad, j * self.dil - self.pad
                if self.stride == 1:
                    out.append((0, dw, 0, dh, i * self.kw + j))
                else:  # input viewed as (2C, W/2, 2, H/2, N): every tap is a plain box
                    ph, pw = dh % 2, dw % 2
                    out.append((pw * self.c_in_p, (dw - pw) // 2, ph, (dh - ph) // 2, i * self.kw + j))
        return out


class BNParam:
    __slots__ = ("name", "weight", "bias", "running_mean", "running_var", "nbt", "momentum", "eps", "c")

    def __init__(self, name, mod):
        self.name = name
        self.weight, self.bias = mod.weight, mod.bias
        self.running_mean, self.running_var = mod.running_mean, mod.running_var
        self.nbt = getattr(mod, "num_batches_tracked", None)
        if mod.momentum is None:
            raise NotImplementedError("pmf_b200: BatchNorm2d(momentum=None) (cumulative moving average) is not implemented; "
                                      "the reference uses the default momentum 0.1 everywhere")
        self.momentum = float(mod.momentum)
        self.eps = float(mod.eps)
        self.c = mod.weight.shape[0]


def _pick_tile(h, w, total):
    best = None
    tw = total
    while tw >= 1:
        th = total // tw
        cost = _rup(w, tw) * _rup(h, th)
        if best is None or cost < best[0]:
            best = (cost, tw, th)
        tw //= 2
    return best[1], best[2]


def _job_table(jobs, device):
    """ctypes WeightJob list -> device-resident table (uint8 tensor).  Built OUTSIDE graph capture (synchronous copy)."""
    arr = (WeightJob * len(jobs))(*jobs)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


class WeightCache:
    """Packed tf32 copies of the conv weights.  Eager mode: re-packed when the parameter's version or storage changes.
    Graph mode (``always=True``): packed once per pass (``begin_pass``) unconditionally, so that the pack kernels are
    part of the captured graph and every replay re-reads the live parameters."""

    def __init__(self, always=False):
        self.entries = {}
        self.always = always
        self.epoch = 0
        self.table = None
        self.precise = False  # set by the Engine of the pass (pmf_b200._lib precision mode)
        self.h16 = False      # "f16" mode: fp16 forward / bf16 dgrad copies of the packed weights
        self.arenas = None
        # frozen (pmf_b200.export / PMFNet.freeze): inference with fixed parameters — weights are packed once and the
        # eval-mode BatchNorm affines (gamma * invstd, beta - mean * gamma * invstd) folded once, kept here, and no pass
        # re-derives them (a captured eval graph then holds the convolutions and their fused epilogues only)
        self.frozen = False
        self.folded = {}

    @staticmethod
    def _convert16(src, dst, dt, stream):
        """flat fp32 packed weights -> 16-bit copy (pmfb_convert16 on a (1,1,size/16,16) view; sizes are multiples of 16)."""
        n16 = src.numel() // 16
        v = View()
        v.ptr, v.sn, v.sy, v.sx = src.data_ptr(), 0, 0, 16
        L.call("pmfb_convert16", C.byref(v), 1, 1, n16, 16, dst.data_ptr(), 0, 0, 16, dt, stream)

    def _split3(self, cp, e, stream):
        """Precise mode: [hi|lo|hi] split packings of the (un-rounded) packed weights (pmfb_split_tf32 mode 1)."""
        for key, rows, cols in (("fwd", cp.c_out_p, cp.c_in_p), ("dgrad", cp.c_in_p, cp.c_out_p)):
            src = e.get(key)
            if src is None:
                e[key + "3"] = None
                continue
            c3 = 3 * _rup(cols, 32)
            dst = e.get(key + "3")
            if dst is None or dst.numel() != cp.taps * rows * c3 or dst.device != src.device:
                dst = torch.empty(cp.taps * rows * c3, device=src.device, dtype=torch.float32)
            v = View()
            v.ptr, v.sn, v.sy, v.sx = src.data_ptr(), 0, rows * cols, cols
            L.call("pmfb_split_tf32", C.byref(v), 1, cp.taps, rows, cols, dst.data_ptr(), 0, rows * c3, c3, 1, stream)
            e[key + "3"] = dst

    def begin_pass(self):
        self.epoch += 1

    # ---- graph mode: every weight of the network packed by ONE launch per pass (pmfb_weight_jobs)
    def build_table(self, params, need_dgrad, device):
        """Allocate the packed buffers of every nn.Conv2d under ``params`` and the device job table that packs them."""
        convs = [params.conv(n) for n, m in params.mods.items() if isinstance(m, torch.nn.Conv2d)]
        if not convs or any(not cp.weight.is_contiguous() for cp in convs):
            self.table = None
            return
        jobs, start = [], 0
        total = sum(cp.taps * cp.c_out_p * cp.c_in_p for cp in convs)
        A = {"fwd": torch.empty(total, device=device, dtype=torch.float32),
             "dgrad": torch.empty(total, device=device, dtype=torch.float32) if need_dgrad else None,
             "fwd16": torch.empty(total, device=device, dtype=torch.float16) if self.h16 else None,
             "dgrad16": torch.empty(total, device=device, dtype=torch.bfloat16) if (self.h16 and need_dgrad) else None}
        self.arenas = A
        for cp in convs:
            size = cp.taps * cp.c_out_p * cp.c_in_p
            e = {k: (None if A[k] is None else A[k][start:start + size]) for k in ("fwd", "dgrad", "fwd16", "dgrad16")}
            e.update(bias=None, tag=None)
            self.entries[cp.name] = e
            j = WeightJob()
            j.src, j.dst, j.dst2 = cp.weight.data_ptr(), e["fwd"].data_ptr(), _p(e["dgrad"])
            j.c_out, j.c_in, j.kh, j.kw, j.stem = cp.c_out, cp.c_in, cp.kh, cp.kw, 1 if cp.stem else 0
            j.c_out_p, j.c_in_p, j.accumulate, j.start = cp.c_out_p, cp.c_in_p, 0, start
            j.no_round = 1 if self.precise else 0
            jobs.append(j)
            start += size
        self.table = (_job_table(jobs, device), len(jobs), start, convs, need_dgrad)

    def pack_all(self, stream):
        tab, n, total, convs, need_dgrad = self.table
        L.call("pmfb_weight_jobs", 0, tab.data_ptr(), n, total, stream)
        if self.h16 and self.arenas is not None:  # ONE conversion launch per arena
            self._convert16(self.arenas["fwd"], self.arenas["fwd16"], L.DT_F16, stream)
            if self.arenas["dgrad16"] is not None:
                self._convert16(self.arenas["dgrad"], self.arenas["dgrad16"], L.DT_BF16, stream)
        tag = ("epoch", self.epoch)
        for cp in convs:
            e = self.entries[cp.name]
            if cp.bias is not None:
                if cp.c_out_p != cp.c_out:
                    b = torch.zeros(cp.c_out_p, device=tab.device, dtype=torch.float32)
                    b[:cp.c_out] = cp.bias.detach()
                    e["bias"] = b
                else:
                    e["bias"] = cp.bias.detach()
            if self.precise:
                self._split3(cp, e, stream)
            e["tag"] = tag

    def get(self, cp, need_dgrad, stream):
        w = cp.weight
        key = cp.name
        e = self.entries.get(key)
        if self.frozen and e is not None and e.get("fwd") is not None and not need_dgrad:
            return e
        if self.always:
            tag = ("epoch", self.epoch)
            if e is not None and e["tag"] == tag and (e["dgrad"] is not None or not need_dgrad):
                return e
        else:
            # re-packed once per pass (Engine): in-place updates through ``.data`` (EMA swaps, manual clipping) do not
            # bump ``_version``, so a version tag alone would keep convolving with stale packed weights
            tag = (self.epoch, w.data_ptr(), None if cp.bias is None else cp.bias.data_ptr(), need_dgrad, self.precise, self.h16)
            if e is not None and e["tag"] == tag:
                return e
        dev = w.device
        if e is None or e["fwd"].device != dev or (need_dgrad and e["dgrad"] is None):
            e = {"fwd": torch.empty(cp.taps * cp.c_out_p * cp.c_in_p, device=dev, dtype=torch.float32),
                 "dgrad": (torch.empty(cp.taps * cp.c_out_p * cp.c_in_p, device=dev, dtype=torch.float32)
                           if need_dgrad else None),
                 "bias": None}
        wd = w.detach()
        if not wd.is_contiguous():
            wd = wd.contiguous()
        L.call("pmfb_pack_weight", wd.data_ptr(), cp.c_out, cp.c_in, cp.kh, cp.kw, 1 if cp.stem else 0, cp.c_out_p,
               cp.c_in_p, e["fwd"].data_ptr(), _p(e["dgrad"]) if need_dgrad else None, 0 if self.precise else 1, stream)
        if self.precise:
            self._split3(cp, e, stream)
        if self.h16:
            e["fwd16"] = torch.empty(e["fwd"].numel(), device=dev, dtype=torch.float16)
            self._convert16(e["fwd"], e["fwd16"], L.DT_F16, stream)
            if need_dgrad:
                e["dgrad16"] = torch.empty(e["dgrad"].numel(), device=dev, dtype=torch.bfloat16)
                self._convert16(e["dgrad"], e["dgrad16"], L.DT_BF16, stream)
        if cp.bias is not None:
            if cp.c_out_p != cp.c_out:
                b = torch.zeros(cp.c_out_p, device=dev, dtype=torch.float32)
                b[:cp.c_out] = cp.bias.detach()
                e["bias"] = b
            else:
                e["bias"] = cp.bias.detach()
        e["tag"] = tag
        self.entries[key] = e
        return e


_side_streams = {}


def _side_stream(device, which=0):
    """Auxiliary CUDA streams per device: 0 for the weight-gradient kernels (Engine._conv_bwd) and the forward camera branch,
    1 for the backward camera branch."""
    key = "%s#%d" % (device, which)
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=device)
    return st


class _Scratch:
    """Bump allocator over one zero-initialised buffer (fp64 reduction cells) or an uninitialised fp32 one.  ``stream`` is a
    handle or a callable returning the engine's CURRENT stream handle: a chunk that has to be added in the middle of a pass
    is zeroed on the stream whose kernels are about to use it (the pass may be on an auxiliary stream, Engine.branch)."""

    def __init__(self, dtype, n, device, stream, zero):
        self.dtype, self.n, self.device, self.stream, self.zero = dtype, n, device, stream, zero
        self.chunks = []
        self._new_chunk()

    def _new_chunk(self):
        self.buf = torch.empty(self.n, device=self.device, dtype=self.dtype)
        if self.zero:
            st = self.stream() if callable(self.stream) else self.stream
            L.call("pmfb_memset_zero", self.buf.data_ptr(), self.n * self.buf.element_size(), st)
        self.chunks.append(self.buf)
        self.pos = 0

    def take(self, k):
        k4 = _rup(k, 4)
        if self.pos + k4 > self.n:
            self.n = max(self.n, k4)
            self._new_chunk()
        out = self.buf[self.pos:self.pos + k]
        self.pos += k4
        return out


class Engine:
    """One forward (and optionally backward) pass.  ``params``: object with conv(name)->ConvParam, bn(name)->BNParam."""

    def __init__(self, params, device, train, record, cache, dropout=None):
        L.require_device()
        self.P = params
        self.device = device
        self.train = bool(train)
        self.record = bool(record)
        if self.record and not self.train:
            raise NotImplementedError("pmf_b200: backward is implemented for train-mode (batch-statistics) BatchNorm only; "
                                      "call .train() or run the eval forward under torch.no_grad()")
        self.cache = cache
        # precision mode of the tensor-core convolutions (pmf_b200._lib): in "3xtf32" nothing is rounded where it is
        # produced (self.R = 0) and every conv operand is split into tf32 hi/lo parts right before the launch
        self.precise = L.get_precision() == "3xtf32"
        # "f16": training passes feed the stride-1 convolutions from 16-bit shadows (fp16 forward operands, bf16 output
        # gradients in dgrad); eval passes fuse everything into the conv epilogues and stay on the tf32 path
        self.h16 = L.get_precision() == "f16" and self.train and str(device).startswith("cuda")
        self.R = 0 if self.precise else 1
        self.cache.h16 = self.h16
        self.cache.precise = self.precise
        self.cache.begin_pass()
        self.st = torch.cuda.current_stream(device).cuda_stream
        self.n_sm = L.query("pmfb_sm_count")
        if self.cache.always and self.cache.table is not None:
            self.cache.pack_all(self.st)
        self._wg_list = []       # recorded convs in forward order (backward: one arena, one memset, one unpack launch)
        self._wg_arena = None
        self._wg_off = {}
        self._unpack_tables = None
        self.segments = None
        self.tape = []
        self.dropout = dropout  # see mask_for
        self.d64 = _Scratch(torch.float64, 1 << 18, device, lambda: self.st, zero=True)
        self.f32 = _Scratch(torch.float32, 1 << 17, device, lambda: self.st, zero=False)
        self._dpre16 = None
        self.param_grads = {}
        self.flat_views = None  # graph mode: {param name: view into one flat gradient buffer}
        self.nbt_list = []
        # backward: the wgrad kernels (tensor pipe + ~2.6 TB/s) run on a side stream next to the BatchNorm-backward /
        # dgrad chain of the following layers (HBM-bound elementwise kernels co-reside with a wgrad CTA on an SM); the
        # operands they read are kept alive until the streams join at the end of run_backward
        self.side = None
        self._side_keep = []
        self.use_side = os.environ.get("PMFB_WGRAD_STREAM", "1") != "0" and str(device).startswith("cuda")
        self.use_branch = FWD_BRANCH and str(device).startswith("cuda")
        self._in_branch = False
        self._branch_done = None
        self._branch_span = None   # [first, one-past-last] tape index recorded inside branch()
        self._marks = []           # (root Act, tape length when a branch marked it ready)
        self._waits = []           # (root Act, tape length when the main stream waited for it)
        self.side2 = None

    # ------------------------------------------------------------------------------------------ forward branches
    @contextlib.contextmanager
    def branch(self):
        """Forward only: the ops issued inside run on the auxiliary stream, concurrently with what the caller issues on the
        main stream afterwards (PMF: the camera stream next to the LiDAR stream — two chains that each alternate a tensor-core
        convolution with HBM-bound elementwise passes, so one chain's elementwise kernels run next to the other's
        convolutions).  Tensors produced inside are handed over with mark_ready / wait_ready; join_branch() before anything
        on the main stream may depend on all of it.  Allocations inside come from the auxiliary stream's allocator pool."""
        if not self.use_branch:
            yield
            return
        if self.side is None:
            self.side = _side_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))  # inputs, packed weights, zeroed scratch are ready
        self.side.wait_event(ev)
        main_st = self.st
        self._branch_span = [len(self.tape), None]
        with torch.cuda.stream(self.side):
            self.st = self.side.cuda_stream
            self._in_branch = True
            try:
                yield
            finally:
                self._in_branch = False
                self.st = main_st
                self._branch_span[1] = len(self.tape)
                self._branch_done = torch.cuda.Event()
                self._branch_done.record(self.side)

    def mark_ready(self, act):
        """Inside branch(): ``act`` is complete at this point of the auxiliary stream."""
        if self._in_branch:
            ev = torch.cuda.Event()
            ev.record(self.side)
            act.root.ready = ev
            self._marks.append((act.root, len(self.tape)))

    def wait_ready(self, act):
        """The current stream waits for a tensor a branch produced (no-op for everything else)."""
        ev = act.root.ready
        if ev is not None and not self._in_branch:
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._waits.append((act.root, len(self.tape)))  # the closure recorded next consumes it on the main stream

    def join_branch(self):
        if self._branch_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._branch_done)
            self._branch_done = None

    # ------------------------------------------------------------------------------------------ helpers
    def _pgrad(self, name, like):
        """Storage for one parameter gradient (a view of the flat buffer in graph mode)."""
        if self.flat_views is not None:
            return self.flat_views[name]
        return torch.empty_like(like, memory_format=torch.contiguous_format)

    def new(self, n, h, w, c, needs_grad=True):
        a = Act(torch.empty((n, h, w, c), device=self.device, dtype=torch.float32), needs_grad=needs_grad)
        if self.h16 and c % 8 == 0 and c >= H16_MIN_K:
            a.h = torch.empty((n, h, w, c), device=self.device, dtype=torch.float16)
            if self.record and H16_WGRAD and c >= H16_WGRAD_MIN:
                a.hb = torch.empty((n, h, w, c), device=self.device, dtype=torch.bfloat16)
        return a

    def _shadow_ptrs(self, out):
        """(fp16 ptr, bf16 ptr) of the shadows of Act ``out`` for a producer that writes them in its own pass (marks them
        valid); (None, None) outside "f16" mode or when the buffer has no shadows."""
        if not self.h16:
            return None, None
        sh, shb = out.shadow(), out.shadow(True)
        for t in (sh, shb):
            if t is not None:
                assert tuple(t.stride()) == tuple(out.t.stride()), (t.stride(), out.t.stride())
        if sh is not None:
            out.shadow_mark()
        if shb is not None:
            out.shadow_mark(True)
        return _p(sh), _p(shb)

    def _ensure_shadow(self, x, bf16=False):
        """fp16 (bf16) shadow view of Act x with current values (converted here if its producer did not write it), or
        None when the buffer has no such shadow."""
        sh = x.shadow(bf16)
        if sh is None:
            return None
        if not x.shadow_valid(bf16):
            n, h, w, c = x.shape
            L.call("pmfb_convert16", C.byref(_view(x.t)), n, h, w, c, sh.data_ptr(), sh.stride(0), sh.stride(1), sh.stride(2),
                   L.DT_BF16 if bf16 else L.DT_F16, self.st)
            x.shadow_mark(bf16)
        return sh

    def _epi(self, alpha1=None, beta1=None, alpha2=None, beta2=None, r1=None, mul=None, r2=None, act=ACT_NONE, rnd=0):
        e = Epilogue()
        e.alpha1, e.beta1, e.alpha2, e.beta2 = _p(alpha1), _p(beta1), _p(alpha2), _p(beta2)
        for name, v in (("r1", r1), ("mul", mul), ("r2", r2)):
            if v is None:
                continue
            setattr(e, name, v if isinstance(v, View) else _view(v))
        e.act, e.round_out = act, (rnd if self.R else 0)
        return e

    def pointwise(self, src, dst, shadow=None, bn_fuse=None, **kw):
        """dst = epilogue(src) elementwise; src None means zeros; src may be a View (broadcast).  ``shadow``: the Act that
        owns ``dst``; in "f16" mode its fp16 shadow is written by the same pass.  ``bn_fuse``: a BnFuse whose finalisation
        (alpha1 / beta1 from the fused channel sums) runs inside this launch (src is then the fp16 pre-BN activation)."""
        n, h, w, c = dst.shape
        e = self._epi(**kw)
        sv = src if isinstance(src, View) else _view(src)
        in_half = 1 if (torch.is_tensor(src) and src.dtype == torch.float16) else 0
        sh = shadow.shadow() if (self.h16 and shadow is not None) else None
        shb = None
        if sh is not None:
            assert tuple(sh.stride()) == tuple(dst.stride()), (sh.stride(), dst.stride())
            shb = shadow.shadow(True)
            shadow.shadow_mark()
            if shb is not None:
                shadow.shadow_mark(True)
        if bn_fuse is not None:
            assert in_half
            L.call("pmfb_pointwise16_bn", C.byref(sv), dst.data_ptr(), dst.stride(0), dst.stride(1), dst.stride(2), n, h, w, c,
                   C.byref(e), _p(sh), L.DT_F16, _p(shb), 1, C.byref(bn_fuse), self.st)
            return
        if sh is not None or in_half:
            L.call("pmfb_pointwise16", C.byref(sv), dst.data_ptr(), dst.stride(0), dst.stride(1), dst.stride(2), n, h, w, c,
                   C.byref(e), _p(sh), L.DT_F16, _p(shb), in_half, self.st)
            return
        L.call("pmfb_pointwise", C.byref(sv), dst.data_ptr(), dst.stride(0), dst.stride(1), dst.stride(2), n, h, w, c,
               C.byref(e), self.st)

    def mask_for(self, site, n, c):
        """Dropout2d scale per (image, channel): 0 or 1/(1-p); None when the site is inactive.
        ``self.dropout``: False = off; dict = explicit {site: (N,C) mask}; otherwise an object with
        draw(site, n, c, device) (pmf_b200.modules._DropoutSites: torch RNG, honours per-module .eval())."""
        if self.dropout is False or self.dropout is None or not self.train:
            return None
        if hasattr(self.dropout, "draw"):
            return self.dropout.draw(site, n, c, self.device)
        m = self.dropout.get(site)
        if m is None:
            return None
        return m.to(device=self.device, dtype=torch.float32).reshape(n, c).contiguous()

    # ------------------------------------------------------------------------------------------ inputs / outputs
    def input_nchw(self, x, c_pad, n_shift=1, out=None, needs_grad=False, rnd=True):
        """NCHW (possibly strided) torch tensor -> NHWC Act (tf32-rounded, channel padded)."""
        n, c, h, w = x.shape
        if out is None:
            out = self.new(n, h, w, c_pad, needs_grad=needs_grad)
        assert out.t.stride(1) == w * out.t.stride(2) and out.t.stride(0) == h * out.t.stride(1)
        xd = x.detach()
        if xd.dtype != torch.float32:
            xd = xd.float()
        L.call("pmfb_pack_input", xd.data_ptr(), xd.stride(0), xd.stride(1), xd.stride(2), xd.stride(3), n, c, h, w, n_shift,
               out.t.data_ptr(), out.c, out.t.stride(2), self.R if rnd else 0, self.st)
        return out

    def to_nchw(self, t, c):
        """NHWC view (first c channels) -> dense NCHW torch tensor."""
        n, h, w, _ = t.shape
        out = torch.empty((n, c, h, w), device=self.device, dtype=torch.float32)
        L.call("pmfb_nhwc_to_nchw", C.byref(_view(t)), n, h, w, c, out.data_ptr(), self.st)
        return out

    # ------------------------------------------------------------------------------------------ raw conv launches
    @staticmethod
    def _tma_src(t, c, stride2=False):
        """TmaSrc over an NHWC view; c = channels exposed to the kernel (zero-filled beyond)."""
        s = TmaSrc()
        n, h, w, _ = t.shape
        sn, sy, sx = t.stride(0), t.stride(1), t.stride(2)
        es = t.element_size()
        s.ptr = t.data_ptr()
        s._keep = t
        if not stride2:
            s.dims[:] = [c, w, 1, h, n]
            s.strides[:] = [sx * es, sy * es, sy * es, sn * es]
        else:
            assert sx == c and h % 2 == 0 and w % 2 == 0, "stride-2 convs need a dense NHWC input"
            s.dims[:] = [2 * c, w // 2, 2, h // 2, n]
            s.strides[:] = [2 * sx * es, sy * es, 2 * sy * es, sn * es]
        return s

    def split(self, t, mode):
        """pmfb_split_tf32 of an NHWC view (precise mode).  mode 0 / 1: dense (n,h,w,3*roundup(c,32)) [hi|hi|lo] / [hi|lo|hi];
        mode 2 / 3: dense (n,h,w,c) hi / lo."""
        n, h, w, c = t.shape
        co = 3 * _rup(c, 32) if mode < 2 else c
        out = torch.empty((n, h, w, co), device=self.device, dtype=torch.float32)
        L.call("pmfb_split_tf32", C.byref(_view(t)), n, h, w, c, out.data_ptr(), out.stride(0), out.stride(1), out.stride(2),
               mode, self.st)
        return out

    def _conv_launch(self, x_t, c_in, stride2, w_packed, c_out, taps, n, out_h, out_w, out_t, epi, bn_stats=None, x16=None,
                     w16=None, dt16=0, out_half=False):
        """Returns True when ``bn_stats`` (2*c_out fp64 sums, zeroed) was accumulated by the conv's own epilogue.
        ``out_half``: out_t is an fp16 buffer (pmfb_conv_desc.out_half, fused statistics only); returns None WITHOUT
        launching when the library has no such epilogue for this layer.
        Precise mode: ``w_packed`` is the [hi|lo|hi] split packing over 3*roundup(c_in,32) channels and x is split here.
        "f16" mode: ``x16`` / ``w16`` are 16-bit shadows of x_t / w_packed (dt16 = DT_F16 or DT_BF16); they are used when the
        library accepts 16-bit operands for this geometry (pmfb_conv16_ok), else the fp32 operands."""
        if x_t is None:  # "f16" mode backward: only the bf16 copy of the output gradient was stored
            assert x16 is not None and w16 is not None and not stride2
        if self.precise:
            c3 = 3 * _rup(c_in, 32)
            x_t = self.split(x_t, 0)
            if stride2:  # the parity view interleaves two pixels along the channel axis: tap_dc counts in split channels
                taps = [((dc // c_in) * c3, dw, dp, dh, wi) for (dc, dw, dp, dh, wi) in taps]
            c_in = c3
        d = ConvDesc()
        d.x = self._tma_src(x_t if x_t is not None else x16, c_in, stride2)
        d.w = w_packed.data_ptr()
        d.c_in, d.c_out, d.n_taps = c_in, c_out, len(taps)
        for i, (dc, dw, dp, dh, wi) in enumerate(taps):
            d.tap_dc[i], d.tap_dw[i], d.tap_dp[i], d.tap_dh[i], d.tap_wi[i] = dc, dw, dp, dh, wi
        d.use_tap_wi = 1
        d.n_batch, d.out_h, d.out_w = n, out_h, out_w
        d.tile_w, d.tile_h = _pick_tile(out_h, out_w, 128)
        d.n_tile = min(256, _rup(c_out, 16))
        d.out = out_t.data_ptr()
        d.o_sn, d.o_sy, d.o_sx = out_t.stride(0), out_t.stride(1), out_t.stride(2)
        d.epi = epi
        d.bn_stats = None
        # 16-bit operands from H16_MIN_K channels on (thinner layers stay on the tf32 path)
        if (x16 is not None and w16 is not None and not stride2 and c_in % 8 == 0 and c_in >= H16_MIN_K
                and L.query("pmfb_conv16_ok", C.byref(d)) == 1):
            d.x = self._tma_src(x16, c_in, False)
            d.w = w16.data_ptr()
            d.dtype = dt16
        elif x_t is None:
            raise L.PmfbError("pmf_b200: the library refused 16-bit operands for a layer planned on the bf16 backward path")
        fused = False
        d.out_half = 1 if out_half else 0
        if bn_stats is not None and L.query("pmfb_conv_fused_stats_ok", C.byref(d)) == 1:
            d.bn_stats = bn_stats.data_ptr()
            fused = True
        if out_half and not fused:
            return None
        L.call("pmfb_conv_fwd", C.byref(d), self.st)
        return fused

    def _conv_fwd(self, x, cp, out_t, epi, bn_stats=None, out_half=False):
        """out = epi(conv(x)); x: Act whose channel count equals cp.c_in_p."""
        e = self.cache.get(cp, self.record, self.st)
        n, h, w, cx = x.shape
        assert cx == cp.c_in_p, (cp.name, cx, cp.c_in_p)
        if cp.stem:
            oh, ow = h, w
        else:
            oh = (h + 2 * cp.pad - cp.dil * (cp.kh - 1) - 1) // cp.stride + 1
            ow = (w + 2 * cp.pad - cp.dil * (cp.kw - 1) - 1) // cp.stride + 1
        assert tuple(out_t.shape) == (n, oh, ow, cp.c_out_p), (cp.name, tuple(out_t.shape), (n, oh, ow, cp.c_out_p))
        x16 = self._ensure_shadow(x) if (self.h16 and cp.stride == 1 and cp.c_in_p % 8 == 0 and cp.c_in_p >= H16_MIN_K
                                         and e.get("fwd16") is not None) else None
        fused = self._conv_launch(x.t, cp.c_in_p, cp.stride == 2, e["fwd3" if self.precise else "fwd"], cp.c_out_p, cp.fwd_taps(),
                                  n, oh, ow, out_t, epi, bn_stats=bn_stats, x16=x16, w16=e.get("fwd16"), dt16=L.DT_F16,
                                  out_half=out_half)
        if fused is None:
            return None
        if self.record:
            self._wg_list.append(cp)
        return fused if bn_stats is not None else e

    # ---- backward segments (graph mode): the reversed tape is cut into K contiguous pieces of roughly equal convolution
    # work; each piece becomes its own CUDA graph and autograd node, so that DDP's bucketed all-reduce of a piece's
    # parameter gradients (tasks/pmf/trainer.py:38-39) overlaps the backward of the next pieces.
    def _backward_order(self):
        """Execution order of the backward: entries ("fn", tape index, "main" | "cam"), ("rec", tag, key), ("wait", tag, key).

        Without a forward branch: the reversed tape on the main stream.  With one (PMF: camera stream = tape [b0, b1), LiDAR
        stream = the rest) the camera stream's closures run on a second auxiliary stream next to the LiDAR stream's.  Both
        chains write the gradients of the J shared feature maps feats[0..J-1], so every writer of one buffer is ordered by
        an event:
          * the decoder's backward (it only needs the camera logits' gradient) goes first and is the FIRST writer of every
            feats[i].grad; the LiDAR chain waits for it before its first feature-map copy backward W[J-1];
          * encoder layer j (reads feats[j-1].grad, accumulates into feats[j-2].grad) is released after W[j-2], i.e. after
            the LiDAR chain's own accumulation into the buffer layer j writes, one stage later than its input would allow;
          * layer 1 (and the stem) follow layer 2 in stream order.
        Segments (plan_segments) end with a join of both auxiliary streams, so a wait whose record fell into an earlier
        segment is simply dropped."""
        n = len(self.tape)
        plain = [("fn", i, "main") for i in range(n - 1, -1, -1)]
        span = self._branch_span
        if not (BWD_BRANCH and self.use_branch and span and span[1] is not None and len(self._marks) >= 2
                and len(self._marks) == len(self._waits)):
            return plain
        b0, b1 = span
        marks = [m for (_r, m) in self._marks]
        W = []
        for (root, _m), (wroot, wi) in zip(self._marks, self._waits):
            if wroot is not root or not (b1 <= wi < n) or getattr(self.tape[wi], "copy_src_root", None) is not root:
                return plain
            W.append(wi)
        if b0 != 0 or sorted(W) != W or sorted(marks) != marks or marks[0] <= b0 or marks[-1] > b1:
            return plain
        J = len(marks)
        groups, lo = [], b0
        for m in marks:  # groups[j-1] = closures of encoder layer j (the stem rides with layer 1), in backward order
            groups.append(list(range(m - 1, lo - 1, -1)))
            lo = m
        order = [("fn", i, "cam") for i in range(b1 - 1, marks[-1] - 1, -1)] + [("rec", "cam", "D")]
        release = {W[j - 2]: j for j in range(J, 1, -1)}
        first_w = max(W)
        for i in range(n - 1, b1 - 1, -1):
            if i == first_w:
                order.append(("wait", "main", "D"))
            order.append(("fn", i, "main"))
            j = release.get(i)
            if j is not None:
                key = "W%d" % i
                order += [("rec", "main", key), ("wait", "cam", key)] + [("fn", c, "cam") for c in groups[j - 1]]
        order += [("fn", c, "cam") for c in groups[0]]
        return order

    def _run_entries(self, entries):
        """Executes entries of _backward_order (the current stream is the main stream)."""
        main = torch.cuda.current_stream(self.device)
        main_st = self.st
        has_cam = self.use_branch and any(e[0] == "fn" and e[2] == "cam" for e in entries)
        bev = {}
        if has_cam:
            if self.side2 is None:
                self.side2 = _side_stream(self.device, 1)
            ev = torch.cuda.Event()
            ev.record(main)
            self.side2.wait_event(ev)
        for e in entries:
            if e[0] == "fn":
                fn = self.tape[e[1]]
                if has_cam and e[2] == "cam":
                    with torch.cuda.stream(self.side2):  # allocations from this stream's pool (see Engine.branch)
                        self.st = self.side2.cuda_stream
                        try:
                            fn()
                        finally:
                            self.st = main_st
                else:
                    fn()
            elif has_cam:
                st = self.side2 if e[1] == "cam" else main
                if e[0] == "rec":
                    bev[e[2]] = torch.cuda.Event()
                    bev[e[2]].record(st)
                elif e[2] in bev:
                    st.wait_event(bev[e[2]])
        if has_cam:
            ev = torch.cuda.Event()
            ev.record(self.side2)
            main.wait_event(ev)

    def plan_segments(self, k):
        """Returns [(entries in execution order, parameter names produced)] * <= k for the recorded tape."""
        order = self._backward_order()

        def fn_of(e):
            return self.tape[e[1]] if e[0] == "fn" else None

        def cost_of(e):
            f = fn_of(e)
            if f is None or getattr(f, "cp", None) is None:
                return 0.0
            return 2.0 * getattr(f, "px", 0) * f.cp.taps * f.cp.c_in_p * f.cp.c_out_p
        cost = [cost_of(e) for e in order]
        total = sum(cost) or 1.0
        segs, cur, acc, done = [], [], 0.0, 0.0
        for e, c in zip(order, cost):
            cur.append(e)
            acc += c
            if len(segs) < k - 1 and done + acc >= total * (len(segs) + 1) / k:
                segs.append(cur)
                done += acc
                cur, acc = [], 0.0
        if cur:
            segs.append(cur)
        out = []
        for es in segs:
            names = []
            for e in es:
                f = fn_of(e)
                cp, bn = getattr(f, "cp", None), getattr(f, "bn", None)
                if cp is not None:
                    names.append(cp.name + ".weight")
                    if cp.bias is not None:
                        names.append(cp.name + ".bias")
                if bn is not None:
                    names += [bn.name + ".weight", bn.name + ".bias"]
            out.append((es, names))
        self.segments = out
        return out

    def prepare_backward(self):
        """Graph mode, called once BEFORE the backward capture: one arena for every packed weight gradient (zeroed by a
        single memset per pass) and, per backward segment, the device job table that unpacks that segment's weight
        gradients into the flat gradient buffer."""
        if self.flat_views is None or not self._wg_list:
            return
        if getattr(self, "segments", None) is None:
            self.plan_segments(1)
        total = sum(cp.taps * cp.c_in_p * cp.c_out_p for cp in self._wg_list)
        self._wg_arena = torch.empty(total, device=self.device, dtype=torch.float32)
        off = 0
        self._unpack_tables = []
        for es, _names in self.segments:
            jobs, start = [], 0
            for e in es:
                cp = getattr(self.tape[e[1]], "cp", None) if e[0] == "fn" else None
                if cp is None:
                    continue
                size = cp.taps * cp.c_in_p * cp.c_out_p
                self._wg_off[cp.name] = (off, size)
                gw = self.flat_views[cp.name + ".weight"]
                j = WeightJob()
                j.src, j.dst, j.dst2 = self._wg_arena.data_ptr() + 4 * off, gw.data_ptr(), None
                j.c_out, j.c_in, j.kh, j.kw, j.stem = cp.c_out, cp.c_in, cp.kh, cp.kw, 1 if cp.stem else 0
                j.c_out_p, j.c_in_p, j.accumulate, j.start = cp.c_out_p, cp.c_in_p, 0, start
                jobs.append(j)
                off += size
                start += cp.c_out * cp.c_in * cp.kh * cp.kw
            self._unpack_tables.append((_job_table(jobs, self.device), len(jobs), start) if jobs else None)

    def _conv_bwd(self, x, cp, d_pre):
        """wgrad (+ dgrad into x's gradient) of out = conv(x) given d_pre = dL/d(conv output), tf32-rounded."""
        d16, self._dpre16 = self._dpre16, None  # bf16 shadow of d_pre left by _bn_backward for THIS layer (or None)
        if d16 is not None and d_pre is not None and tuple(d16.shape) != tuple(d_pre.shape):
            d16 = None
        e = self.cache.get(cp, True, self.st)
        n, h, w, _ = x.shape
        _, oh, ow, co = (d_pre if d_pre is not None else d16).shape
        assert co == cp.c_out_p
        # Order of enqueue: the event "d_pre is ready" is recorded first, then the dgrad goes to the main stream and only then
        # the wgrad to the side stream.  Both become runnable at the same moment and their CTAs cannot share an SM (180 KB +
        # 190 KB of shared memory): the dgrad is on the critical path, the wgrad can fill in next to the BatchNorm-backward
        # kernels of the following layer, so the dgrad should win the SMs.
        ev_ready = None
        if self.use_side:
            if self.side is None:
                self.side = _side_stream(self.device)
            ev_ready = torch.cuda.Event()
            ev_ready.record(torch.cuda.current_stream(self.device))   # d_pre (and, the first time, the zeroed arena) are ready
        main_st = self.st

        def launch_wgrad():
            if not self.use_side:
                return _wgrad_body()
            # Everything the wgrad allocates (split operands of the precise mode, lazily converted shadows, the packed
            # gradient) must come from the SIDE stream's allocator pool: a block the main stream has just released (the
            # dgrad's temporaries above) may still be read by a kernel in flight there.
            with torch.cuda.stream(self.side):
                return _wgrad_body()

        def _wgrad_body():
            # ---- wgrad -> packed [taps][c_in_p][c_out_p] (split-K atomics; buffer zeroed here) -> OIHW
            if self.use_side:
                self.side.wait_event(ev_ready)
                self.st = self.side.cuda_stream
            batched = self._wg_arena is not None and cp.name in self._wg_off
            if batched:
                off, size = self._wg_off[cp.name]
                packed = self._wg_arena[off:off + size]
            else:
                packed = torch.empty(cp.taps * cp.c_in_p * cp.c_out_p, device=self.device, dtype=torch.float32)
                L.call("pmfb_memset_zero", packed.data_ptr(), packed.numel() * 4, self.st)
            d = WgradDesc()
            d.c_in, d.c_out, d.n_taps = cp.c_in_p, cp.c_out_p, cp.taps
            for i, (dc, dw, dp, dh, _wi) in enumerate(cp.fwd_taps()):
                d.tap_dc[i], d.tap_dw[i], d.tap_dp[i], d.tap_dh[i] = dc, dw, dp, dh
            d.n_batch, d.out_h, d.out_w = n, oh, ow
            d.ptile_w, d.ptile_h = _pick_tile(oh, ow, 32)
            d.n_tile = min(256, _rup(cp.c_out_p, 32))
            total_pt = (_rup(ow, d.ptile_w) // d.ptile_w) * (_rup(oh, d.ptile_h) // d.ptile_h) * n
            base = ((cp.taps * (_rup(cp.c_in_p, 32) // 32) + 3) // 4) * ((cp.c_out_p + d.n_tile - 1) // d.n_tile)
            d.ksplit = max(1, min(max(1, total_pt // 4), (2 * self.n_sm + base - 1) // base))
            d.dw = packed.data_ptr()
            if self.use_side:
                self._side_keep.append((x.t, d_pre, packed))
            if self.precise:
                # dw = x_hi*dy_hi + x_hi*dy_lo + x_lo*dy_hi: three launches accumulating into the same packed gradient
                x_hi, x_lo = self.split(x.t, 2), self.split(x.t, 3)
                dy_hi, dy_lo = self.split(d_pre, 2), self.split(d_pre, 3)
                pairs = ((x_hi, dy_hi), (x_hi, dy_lo), (x_lo, dy_hi))
            else:
                pairs = ((x.t, d_pre),)
                # "f16" mode: bf16 shadows of both operands (kind::f16, K = 16 pixels per UMMA) where the library takes them
                if (self.h16 and H16_WGRAD and d16 is not None and cp.stride == 1 and cp.c_in_p % 8 == 0
                        and cp.c_in_p >= H16_WGRAD_MIN and cp.c_out_p >= H16_WGRAD_MIN):
                    d.x = self._tma_src(x.t, cp.c_in_p, False)
                    d.dy = self._tma_src(d16, cp.c_out_p)
                    if L.query("pmfb_wgrad16_ok", C.byref(d)) == 1:
                        xb = self._ensure_shadow(x, bf16=True)  # written by x's producer, else converted here once
                        if xb is None:
                            xb = torch.empty(x.t.shape, device=self.device, dtype=torch.bfloat16)
                            L.call("pmfb_convert16", C.byref(_view(x.t)), n, h, w, cp.c_in_p, xb.data_ptr(), xb.stride(0),
                                   xb.stride(1), xb.stride(2), L.DT_BF16, self.st)
                        pairs = ((xb, d16),)
                        d.dtype = L.DT_BF16
                assert pairs[0][1] is not None, "the fp32 output gradient was skipped for a layer whose wgrad needs it: " + cp.name
            for xa, dya in pairs:
                d.x = self._tma_src(xa, cp.c_in_p, cp.stride == 2)
                d.dy = self._tma_src(dya, cp.c_out_p)
                L.call("pmfb_conv_wgrad", C.byref(d), self.st)
                if self.use_side:
                    self._side_keep.append((xa, dya))
            gw = self._pgrad(cp.name + ".weight", cp.weight)
            if not batched:
                L.call("pmfb_unpack_wgrad", packed.data_ptr(), cp.c_out, cp.c_in, cp.kh, cp.kw, 1 if cp.stem else 0, cp.c_out_p,
                       cp.c_in_p, gw.data_ptr(), 0, self.st)
            self.st = main_st
            self.param_grads[cp.name + ".weight"] = gw

        # ---- dgrad
        if not x.needs_grad:
            launch_wgrad()
            return
        assert not cp.stem and cp.c_in_p == cp.c_in
        gx, acc = x.grad_target()
        rnd = 1 if x.round_grad else 0
        w_dgrad = e["dgrad3" if self.precise else "dgrad"]
        if cp.stride == 1:
            taps = [(0, -dw, 0, -dh, wi) for (_dc, dw, _dp, dh, wi) in cp.fwd_taps()]
            self._conv_launch(d_pre, cp.c_out_p, False, w_dgrad, cp.c_in_p, taps, n, h, w, gx,
                              self._epi(r1=gx if acc else None, rnd=rnd), x16=d16, w16=e.get("dgrad16"), dt16=L.DT_BF16)
            launch_wgrad()
            return
        # stride 2: one stride-1 convolution over dy per input parity class
        for py in (0, 1):
            for px in (0, 1):
                taps = []
                for i in range(cp.kh):
                    for j in range(cp.kw):
                        dh, dw = i * cp.dil - cp.pad, j * cp.dil - cp.pad
                        if (py - dh) % 2 == 0 and (px - dw) % 2 == 0:
                            taps.append((0, (px - dw) // 2, 0, (py - dh) // 2, i * cp.kw + j))
                sub = gx[:, py::2, px::2, :]
                if not taps:
                    if not acc:
                        self.pointwise(None, sub)
                    continue
                self._conv_launch(d_pre, cp.c_out_p, False, w_dgrad, cp.c_in_p, taps, n, h // 2, w // 2, sub,
                                  self._epi(r1=sub if acc else None, rnd=rnd), x16=d16, w16=e.get("dgrad16"), dt16=L.DT_BF16)
        launch_wgrad()

    def _bias_grad(self, cp, colsum64):
        gb = self._pgrad(cp.name + ".bias", cp.bias)
        L.call("pmfb_d2f", colsum64.data_ptr(), gb.data_ptr(), cp.c_out, 1.0, 0, 0, self.st)
        self.param_grads[cp.name + ".bias"] = gb

    # ------------------------------------------------------------------------------------------ BatchNorm pieces
    def _bn_eval_affine(self, bn):
        if self.cache.frozen:
            hit = self.cache.folded.get(bn.name)
            if hit is not None:
                return hit
            ab = torch.empty(2 * bn.c, device=self.device, dtype=torch.float32)
            self.cache.folded[bn.name] = (ab[:bn.c], ab[bn.c:])
        else:
            ab = self.f32.take(2 * bn.c)
        alpha, beta = ab[:bn.c], ab[bn.c:]
        L.call("pmfb_bn_finalize", None, 0, bn.c, _p(bn.weight.detach()), _p(bn.bias.detach()), bn.running_mean.data_ptr(),
               bn.running_var.data_ptr(), bn.momentum, bn.eps, alpha.data_ptr(), beta.data_ptr(), None, None, self.st)
        return alpha, beta

    def _conv_fwd_with_stats(self, x, cp, bn, shp, epi):
        """(a, stats, fuse): a = epi(conv(x)), the pre-BatchNorm activation, and the training-mode BatchNorm statistics of it
        -- fused into the conv's epilogue where the library supports it (pmfb_conv_fused_stats_ok), else by a separate
        pmfb_bn_stats pass.  "f16" mode: a is stored as fp16 where the fused epilogue exists (statistics of the stored
        values); the three BatchNorm passes that read it take fp16, and the finalisation (sums -> alpha / beta / mean /
        invstd, running statistics) is deferred into the BN-apply launch: ``fuse`` is then the BnFuse the caller hands to
        ``pointwise`` -- stats' vectors are filled by that launch -- and alpha1 / beta1 must not be passed."""
        assert cp.c_out_p == bn.c
        sums = self.d64.take(2 * bn.c)
        if self.h16 and PRE_BN_HALF and shp[3] % 8 == 0:
            a = torch.empty(shp, device=self.device, dtype=torch.float16)
            if self._conv_fwd(x, cp, a, epi, bn_stats=sums, out_half=True):
                if not BN_FUSE_FINALIZE:
                    return a, self._bn_train_affine(bn, a, sums=sums, have_sums=True), None
                c = bn.c
                v = self.f32.take(4 * c)
                stats = (v[:c], v[c:2 * c], v[2 * c:3 * c], v[3 * c:])
                f = BnFuse()
                f.sums, f.count = sums.data_ptr(), shp[0] * shp[1] * shp[2]
                f.gamma, f.beta = _p(bn.weight.detach()), _p(bn.bias.detach())
                f.running_mean, f.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                f.momentum, f.eps = bn.momentum, bn.eps
                f.alpha_out, f.beta_out, f.mean_out, f.invstd_out = (t.data_ptr() for t in stats)
                f._keep = (sums, v, bn)
                if bn.nbt is not None:
                    self.nbt_list.append(bn.nbt)
                return a, stats, f
        a = torch.empty(shp, device=self.device, dtype=torch.float32)
        fused = self._conv_fwd(x, cp, a, epi, bn_stats=sums)
        return a, self._bn_train_affine(bn, a, sums=sums, have_sums=fused), None

    def _bn_train_affine(self, bn, a_t, sums=None, have_sums=False):
        n, h, w, c = a_t.shape
        assert c == bn.c
        if sums is None:
            sums = self.d64.take(2 * c)
        if not have_sums:
            L.call("pmfb_bn_stats", C.byref(_view(a_t)), n, h, w, c, sums.data_ptr(), self.st)
        v = self.f32.take(4 * c)
        alpha, beta, mean, invstd = v[:c], v[c:2 * c], v[2 * c:3 * c], v[3 * c:]
        L.call("pmfb_bn_finalize", sums.data_ptr(), n * h * w, c, _p(bn.weight.detach()), _p(bn.bias.detach()),
               bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.momentum, bn.eps, alpha.data_ptr(), beta.data_ptr(),
               mean.data_ptr(), invstd.data_ptr(), self.st)
        if bn.nbt is not None:
            self.nbt_list.append(bn.nbt)
        return alpha, beta, mean, invstd

    def _layer16(self, x, cp):
        """True when BOTH backward kernels of this layer run from bf16 operands in "f16" mode (the fp32 copy of the
        gradient with respect to the conv output is then never read and is not stored)."""
        if not (self.h16 and H16_WGRAD and cp.stride == 1 and not cp.stem and cp.c_in_p % 8 == 0 and cp.c_out_p % 8 == 0
                and cp.c_in_p >= H16_WGRAD_MIN and cp.c_out_p >= H16_WGRAD_MIN and x.shadow(True) is not None
                and cp.c_in_p <= 512 and cp.c_out_p <= 512):  # the halo kernels stage <= 512 output channels
            return False
        reach = cp.dil * (cp.kh - 1) - cp.pad if cp.kh > 1 else 0
        return max(abs(reach), abs(cp.pad)) <= 2  # taps inside the halo kernels' +-2 window

    def _bn_backward(self, bn, dy, a_t, stats, mul=None, z=None, act_z=ACT_NONE, leaky_x=0, want_colsum=False,
                     g_out=None, g_acc=False, need32=True):
        """Returns (d_pre [tf32-rounded gradient w.r.t. the conv output], colsum64 or None)."""
        alpha, beta, mean, invstd = stats
        n, h, w, c = a_t.shape
        red = self.d64.take(2 * c)
        dyv, mulv, zv, xv = _view(dy), (mul if isinstance(mul, View) else _view(mul)), _view(z), _view(a_t)
        x_half = 1 if a_t.dtype == torch.float16 else 0
        L.call("pmfb_bn_bwd_reduce16", C.byref(dyv), C.byref(mulv), C.byref(zv), act_z, C.byref(xv), mean.data_ptr(),
               invstd.data_ptr(), alpha.data_ptr(), beta.data_ptr(), n, h, w, c, red.data_ptr(), x_half, self.st)
        # "f16" mode: the same pass also stores d_pre as bf16, the operand of the kind::f16 dgrad / wgrad; the fp32 copy is
        # skipped when nothing reads it (need32 False)
        self._dpre16 = torch.empty((n, h, w, c), device=self.device, dtype=torch.bfloat16) if (self.h16 and c % 8 == 0 and c >= H16_MIN_K) else None
        if self._dpre16 is None:
            need32 = True
        d_pre = torch.empty((n, h, w, c), device=self.device, dtype=torch.float32) if need32 else None
        cs = self.d64.take(c) if want_colsum else None
        gw, gb = self._pgrad(bn.name + ".weight", bn.weight), self._pgrad(bn.name + ".bias", bn.bias)
        L.call("pmfb_bn_bwd_apply16", C.byref(dyv), C.byref(mulv), C.byref(zv), act_z, C.byref(xv), mean.data_ptr(),
               invstd.data_ptr(), alpha.data_ptr(), beta.data_ptr(), bn.weight.detach().data_ptr(), red.data_ptr(), leaky_x,
               n, h, w, c, _p(d_pre), c * h * w, c * w, c, self.R, gw.data_ptr(),
               gb.data_ptr(), _p(cs), _p(g_out), *( (g_out.stride(0), g_out.stride(1), g_out.stride(2)) if g_out is not None
                                                    else (0, 0, 0)), 1 if g_acc else 0, _p(self._dpre16), x_half, self.st)
        self.param_grads[bn.name + ".weight"] = gw
        self.param_grads[bn.name + ".bias"] = gb
        return d_pre, cs

    def _out_for(self, x, cp, out):
        n, h, w, _ = x.shape
        if cp.stem:
            oh, ow = h, w
        else:
            oh = (h + 2 * cp.pad - cp.dil * (cp.kh - 1) - 1) // cp.stride + 1
            ow = (w + 2 * cp.pad - cp.dil * (cp.kw - 1) - 1) // cp.stride + 1
        if out is None:
            out = self.new(n, oh, ow, cp.c_out_p)
        return out, (n, oh, ow, cp.c_out_p)

    # ------------------------------------------------------------------------------------------ fused layer patterns
    def conv_act(self, x, cname, act=ACT_NONE, out=None, rnd=True):
        """y = act(conv(x) + bias), act in {NONE, LEAKY}."""
        cp = self.P.conv(cname)
        y, _ = self._out_for(x, cp, out)
        e = self.cache.get(cp, self.record, self.st)
        self._conv_fwd(x, cp, y.t, self._epi(beta1=e["bias"], act=act, rnd=1 if rnd else 0))
        if self.record:
            if act == ACT_NONE:
                y.round_grad = True

            def bwd():
                dy = y.grad_read()
                n, h, w, c = dy.shape
                cs = self.d64.take(c) if cp.bias is not None else None
                if act == ACT_NONE:
                    d_pre = dy
                    if cs is not None:
                        L.call("pmfb_colsum", C.byref(_view(dy)), n, h, w, c, 0, cs.data_ptr(), self.st)
                else:
                    d_pre = torch.empty((n, h, w, c), device=self.device, dtype=torch.float32)
                    nv = View()
                    L.call("pmfb_bn_bwd_apply", C.byref(_view(dy)), C.byref(nv), C.byref(_view(y.t)), act, C.byref(nv), None,
                           None, None, None, None, None, 0, n, h, w, c, d_pre.data_ptr(), d_pre.stride(0), d_pre.stride(1),
                           d_pre.stride(2), self.R, None, None, _p(cs), None, 0, 0, 0, 0, self.st)
                if cs is not None:
                    self._bias_grad(cp, cs)
                self._conv_bwd(x, cp, d_pre)

            bwd.cp, bwd.bn = cp, None
            bwd.px = y.shape[0] * y.shape[1] * y.shape[2]
            self.tape.append(bwd)
        return y

    def conv_act_bn(self, x, cname, bnname, out=None, shortcut=None, mask=None):
        """y = (BN(LeakyReLU(conv(x) + bias)) + shortcut) * mask   (SalsaNext / fusion / decoder order)."""
        cp, bn = self.P.conv(cname), self.P.bn(bnname)
        y, shp = self._out_for(x, cp, out)
        e = self.cache.get(cp, self.record, self.st)
        sc_t = None if shortcut is None else shortcut.t
        if not self.train:
            alpha, beta = self._bn_eval_affine(bn)
            self._conv_fwd(x, cp, y.t, self._epi(beta1=e["bias"], act=ACT_LEAKY, alpha2=alpha, beta2=beta, r2=sc_t, rnd=1))
            return y
        a, stats, fuse = self._conv_fwd_with_stats(x, cp, bn, shp, self._epi(beta1=e["bias"], act=ACT_LEAKY))
        mv = None if mask is None else _chan_view(mask)
        ab = {} if fuse is not None else dict(alpha1=stats[0], beta1=stats[1])
        self.pointwise(a, y.t, shadow=y, bn_fuse=fuse, r1=sc_t, mul=mv, rnd=1, **ab)
        if self.record:
            def bwd():
                dy = y.grad_read()
                g_out, g_acc = (None, False)
                if shortcut is not None and shortcut.needs_grad:
                    g_out, g_acc = shortcut.grad_target()
                d_pre, cs = self._bn_backward(bn, dy, a, stats, mul=mv, leaky_x=1, want_colsum=cp.bias is not None,
                                              g_out=g_out, g_acc=g_acc, need32=not self._layer16(x, cp))
                if cs is not None:
                    self._bias_grad(cp, cs)
                self._conv_bwd(x, cp, d_pre)

            bwd.cp, bwd.bn = cp, bn
            bwd.px = shp[0] * shp[1] * shp[2]
            self.tape.append(bwd)
        return y

    def conv_bn(self, x, cname, bnname, post=ACT_NONE, out=None, identity=None, mask=None, gate=None, rnd=True):
        """y = post(BN(conv(x) + bias) + identity) * mask          (torchvision blocks, attention.0/1/2)
        gate=(f, pcd):  y = sigmoid(BN(conv(x) + bias)) * f + pcd    (attention.3/4/5 + pmf_net.py:35)."""
        cp, bn = self.P.conv(cname), self.P.bn(bnname)
        y, shp = self._out_for(x, cp, out)
        e = self.cache.get(cp, self.record, self.st)
        id_t = None if identity is None else identity.t
        f_t, pcd_t = (gate[0].t, gate[1].t) if gate is not None else (None, None)
        if not self.train:
            alpha, beta = self._bn_eval_affine(bn)
            if e["bias"] is not None:  # beta' = alpha*bias + beta  (tiny per-channel vector op on the device)
                hit = self.cache.folded.get(bn.name + "|bias") if self.cache.frozen else None
                if hit is not None:
                    beta = hit
                else:
                    b2 = torch.empty(bn.c, device=self.device, dtype=torch.float32) if self.cache.frozen else self.f32.take(bn.c)
                    self.pointwise(e["bias"].view(1, 1, 1, -1), b2.view(1, 1, 1, -1), alpha1=alpha, beta1=beta)
                    if self.cache.frozen:
                        self.cache.folded[bn.name + "|bias"] = b2
                    beta = b2
            self._conv_fwd(x, cp, y.t, self._epi(alpha1=alpha, beta1=beta, r1=id_t, act=post, mul=f_t, r2=pcd_t,
                                                 rnd=1 if rnd else 0))
            return y
        c_t, stats, fuse = self._conv_fwd_with_stats(x, cp, bn, shp, self._epi(beta1=e["bias"]))
        mv = None if mask is None else _chan_view(mask)
        ab = {} if fuse is not None else dict(alpha1=stats[0], beta1=stats[1])
        self.pointwise(c_t, y.t, shadow=y if rnd else None, bn_fuse=fuse, r1=id_t, act=post,
                       mul=f_t if gate is not None else mv, r2=pcd_t, rnd=1 if rnd else 0, **ab)
        if self.record:
            def bwd():
                dy = y.grad_read()
                if gate is not None:
                    f, pcd = gate
                    gf, acc = f.grad_target()  # d f = dy * sigmoid(BN(c))
                    self.pointwise(c_t, gf, alpha1=stats[0], beta1=stats[1], act=ACT_SIGMOID, mul=dy, r2=gf if acc else None)
                    if pcd.needs_grad:
                        gp, acc = pcd.grad_target()  # d pcd = dy
                        self.pointwise(dy, gp, r1=gp if acc else None)
                    d_pre, cs = self._bn_backward(bn, dy, c_t, stats, mul=f.t, z=None, act_z=ACT_SIGMOID,
                                                  want_colsum=cp.bias is not None, need32=not self._layer16(x, cp))
                else:
                    g_out, g_acc = (None, False)
                    if identity is not None and identity.needs_grad:
                        g_out, g_acc = identity.grad_target()
                    d_pre, cs = self._bn_backward(bn, dy, c_t, stats, mul=mv, z=y.t if post != ACT_NONE else None, act_z=post,
                                                  want_colsum=cp.bias is not None, g_out=g_out, g_acc=g_acc,
                                                  need32=not self._layer16(x, cp))
                if cs is not None:
                    self._bias_grad(cp, cs)
                self._conv_bwd(x, cp, d_pre)

            bwd.cp, bwd.bn = cp, bn
            bwd.px = shp[0] * shp[1] * shp[2]
            self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------------------------------ EPMF sparse-conv ops (eval)
    def pixel_mask(self, x):
        """(N,H,W) fp32 map: 1 where any channel of the pixel is non-zero (epmf_net.py:67)."""
        n, h, w, c = x.shape
        m = torch.empty((n, h, w), device=self.device, dtype=torch.float32)
        L.call("pmfb_pixel_mask", C.byref(_view(x.t)), n, h, w, c, m.data_ptr(), self.st)
        return m

    def mask_maxpool(self, m, k, stride, dil, pad):
        """MaxPool2d(k, stride, 0, dil)(F.pad(m, pad)) of a pixel mask (epmf_net.py:43-44)."""
        n, h, w = m.shape
        oh = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
        ow = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
        out = torch.empty((n, oh, ow), device=self.device, dtype=torch.float32)
        L.call("pmfb_mask_maxpool", m.data_ptr(), n, h, w, k, stride, dil, pad, out.data_ptr(), self.st)
        return out

    def sparse_conv(self, x, mask, name):
        """SparseVariantConv (epmf_net.py:10-50) up to, not including, the final ``* mask``: returns
        (conv(x) + conv.bias + bias  [NHWC tensor, NOT yet masked],  dilated mask).  The caller guarantees x == x*mask
        (true whenever x was produced with ``post=mask`` or the mask was derived from x itself)."""
        if self.record:
            raise NotImplementedError("pmf_b200: the EPMF sparse convolution has no backward yet (inference only)")
        cp = self.P.conv(name + ".conv")
        extra = self.P.mods[name.lstrip(".")].bias
        e = self.cache.get(cp, False, self.st)
        bias = e["bias"]
        if extra is not None:
            b2 = self.f32.take(cp.c_out_p)
            src = bias if bias is not None else torch.zeros(cp.c_out_p, device=self.device)
            ex = extra.detach()
            if cp.c_out_p != cp.c_out:
                exp = torch.zeros(cp.c_out_p, device=self.device, dtype=torch.float32)
                exp[:cp.c_out] = ex
                ex = exp
            self.pointwise(src.view(1, 1, 1, -1), b2.view(1, 1, 1, -1), r1=ex.view(1, 1, 1, -1))
            bias = b2
        n, h, w, _ = x.shape
        oh = (h + 2 * cp.pad - cp.dil * (cp.kh - 1) - 1) // cp.stride + 1
        ow = (w + 2 * cp.pad - cp.dil * (cp.kw - 1) - 1) // cp.stride + 1
        y = torch.empty((n, oh, ow, cp.c_out_p), device=self.device, dtype=torch.float32)
        self._conv_fwd(x, cp, y, self._epi(beta1=bias))
        return y, self.mask_maxpool(mask, cp.kh, cp.stride, cp.dil, cp.pad)

    def pixel_scale(self, src_t, out=None, pre=None, act=ACT_NONE, bn=None, r=None, post=None, rnd=True):
        """out = (BN_eval(act(src * pre[pixel])) + r) * post[pixel]  (epmf_net.py:31,49,69-82)."""
        n, h, w, c = src_t.shape
        if out is None:
            out = self.new(n, h, w, c, needs_grad=False)
        alpha = beta = None
        if bn is not None:
            alpha, beta = self._bn_eval_affine(self.P.bn(bn))
        rv = _view(r.t) if r is not None else View()
        L.call("pmfb_pixel_scale", C.byref(_view(src_t)), n, h, w, c, _p(pre), act, _p(alpha), _p(beta), C.byref(rv), _p(post),
               out.t.data_ptr(), out.t.stride(0), out.t.stride(1), out.t.stride(2), self.R if rnd else 0, self.st)
        return out

    # ------------------------------------------------------------------------------------------ data movement ops
    def copy(self, src, dst, mask=None):
        """dst = src * mask (a channel slice of a concat buffer)."""
        mv = None if mask is None else _chan_view(mask)
        self.pointwise(src.t, dst.t, shadow=dst, mul=mv, rnd=1 if mask is not None else 0)
        if self.record and src.needs_grad:
            def bwd():
                g = dst.grad_read()
                gs, acc = src.grad_target()
                self.pointwise(g, gs, mul=mv, r2=gs if acc else None, rnd=1 if src.round_grad else 0)

            bwd.copy_src_root = src.root
            self.tape.append(bwd)
        return dst

    def pool(self, x, kind, out=None, mask=None):
        """3x3 stride-2 pad-1 pooling; kind 'avg' (salsanext.py:65) or 'max' (torchvision stem)."""
        n, h, w, c = x.shape
        k = 0 if kind == "avg" else 1
        if out is None:
            out = self.new(n, h // 2, w // 2, c)
        idx = torch.empty((n, h // 2, w // 2, c), device=self.device, dtype=torch.uint8) if (k == 1 and self.record) else None
        s16, s16b = self._shadow_ptrs(out)
        L.call("pmfb_pool3s2", k, C.byref(_view(x.t)), n, h, w, c, _p(mask), out.t.data_ptr(), out.t.stride(0), out.t.stride(1),
               out.t.stride(2), _p(idx), self.R, s16, s16b, self.st)
        if self.record and x.needs_grad:
            def bwd():
                g = out.grad_read()
                gx, acc = x.grad_target()
                L.call("pmfb_pool3s2_bwd", k, C.byref(_view(g)), n, h, w, c, _p(mask), gx.data_ptr(), gx.stride(0), gx.stride(1),
                       gx.stride(2), _p(idx), 1 if acc else 0, self.st)

            self.tape.append(bwd)
        return out

    def pixel_shuffle(self, x, out, mask=None):
        """nn.PixelShuffle(2) (+ Dropout2d scale): out (N,2h,2w,c/4)."""
        n, h, w, c4 = x.shape
        c = c4 // 4
        s16, s16b = self._shadow_ptrs(out)
        L.call("pmfb_pixel_shuffle", C.byref(_view(x.t)), n, h, w, c, _p(mask), out.t.data_ptr(), out.t.stride(0),
               out.t.stride(1), out.t.stride(2), self.R, s16, s16b, self.st)
        if self.record and x.needs_grad:
            def bwd():
                g = out.grad_read()
                gx, acc = x.grad_target()
                L.call("pmfb_pixel_shuffle_bwd", C.byref(_view(g)), n, h, w, c, _p(mask), gx.data_ptr(), gx.stride(0),
                       gx.stride(1), gx.stride(2), 1 if acc else 0, self.R if x.round_grad else 0, self.st)

            self.tape.append(bwd)
        return out

    def upsample2x(self, x, out):
        n, h, w, c = x.shape
        s16, s16b = self._shadow_ptrs(out)
        L.call("pmfb_upsample2x", C.byref(_view(x.t)), n, h, w, c, out.t.data_ptr(), out.t.stride(0), out.t.stride(1),
               out.t.stride(2), self.R, s16, s16b, self.st)
        if self.record and x.needs_grad:
            def bwd():
                g = out.grad_read()
                gx, acc = x.grad_target()
                L.call("pmfb_upsample2x_bwd", C.byref(_view(g)), n, h, w, c, gx.data_ptr(), gx.stride(0), gx.stride(1),
                       gx.stride(2), 1 if acc else 0, self.st)

            self.tape.append(bwd)
        return out

    def global_avg(self, x):
        """AdaptiveAvgPool2d(1): (N,h,w,C) -> (N,1,1,C)."""
        n, h, w, c = x.shape
        s = self.d64.take(n * c)
        L.call("pmfb_colsum", C.byref(_view(x.t)), n, h, w, c, 1, s.data_ptr(), self.st)
        out = self.new(n, 1, 1, c)
        L.call("pmfb_d2f", s.data_ptr(), out.t.data_ptr(), n * c, 1.0 / (h * w), 0, self.R, self.st)
        if self.record and x.needs_grad:
            def bwd():
                g = out.grad_read()  # (N,1,1,C)
                gx, acc = x.grad_target()
                scale = torch.full((c,), 1.0 / (h * w), device=self.device, dtype=torch.float32)
                bv = View()
                bv.ptr, bv.sn, bv.sy, bv.sx = g.data_ptr(), g.stride(0), 0, 0
                self.pointwise(bv, gx, alpha1=scale, r1=gx if acc else None, rnd=1 if x.round_grad else 0)

            self.tape.append(bwd)
        return out

    def broadcast(self, v, out):
        """(N,1,1,C) -> (N,h,w,C) (F.interpolate of a 1x1 map, pmf_net.py:124-125)."""
        n, h, w, c = out.shape
        bv = View()
        bv.ptr, bv.sn, bv.sy, bv.sx = v.t.data_ptr(), v.t.stride(0), 0, 0
        self.pointwise(bv, out.t)
        if self.record and v.needs_grad:
            def bwd():
                g = out.grad_read()
                s = self.d64.take(n * c)
                L.call("pmfb_colsum", C.byref(_view(g)), n, h, w, c, 1, s.data_ptr(), self.st)
                gv, acc = v.grad_target()
                L.call("pmfb_d2f", s.data_ptr(), gv.data_ptr(), n * c, 1.0, 1 if acc else 0, self.R if v.round_grad else 0, self.st)

            self.tape.append(bwd)
        return out

    def softmax_nchw(self, logits, nclasses):
        """F.softmax(dim=1) -> dense NCHW probabilities (the module's return value)."""
        n, h, w, _ = logits.shape
        out = torch.empty((n, nclasses, h, w), device=self.device, dtype=torch.float32)
        L.call("pmfb_softmax_nchw", C.byref(_view(logits.t)), n, h, w, nclasses, out.data_ptr(), self.st)
        return out

    def softmax_backward(self, logits, probs, dprobs):
        """d logits from d probs (both dense NCHW) into the logits' gradient buffer; returns that buffer."""
        g, acc = logits.grad_target()
        assert not acc
        self.softmax_backward_into(g, probs, dprobs)
        return g

    def softmax_backward_into(self, g, probs, dprobs, stream=None):
        n, h, w, _ = g.shape
        if not dprobs.is_contiguous():
            dprobs = dprobs.contiguous()
        L.call("pmfb_softmax_nchw_bwd", probs.data_ptr(), dprobs.data_ptr(), n, h, w, probs.shape[1], g.data_ptr(), g.stride(0),
               g.stride(1), g.stride(2), self.R, self.st if stream is None else stream)

    # ------------------------------------------------------------------------------------------ tape
    def finish_forward(self):
        if self.train and self.nbt_list:
            torch._foreach_add_(self.nbt_list, 1)  # num_batches_tracked bookkeeping (host-side plumbing)
            self.nbt_list = []

    def join_side(self):
        """The main stream waits for every wgrad launched on the side stream so far; their operands may be released."""
        if self.side is not None and self._side_keep:
            ev = torch.cuda.Event()
            ev.record(self.side)
            torch.cuda.current_stream(self.device).wait_event(ev)
        self._side_keep = []

    def run_backward(self):
        self.begin_backward()
        self._run_entries(self._backward_order())
        self.join_side()
        for tab in (self._unpack_tables or []):
            if tab is not None:
                L.call("pmfb_weight_jobs", 1, tab[0].data_ptr(), tab[1], tab[2], self.st)
        self.tape = []
        return self.param_grads

    def begin_backward(self):
        self.st = torch.cuda.current_stream(self.device).cuda_stream
        self.d64 = _Scratch(torch.float64, 1 << 18, self.device, lambda: self.st, zero=True)
        if self._wg_arena is not None:
            L.call("pmfb_memset_zero", self._wg_arena.data_ptr(), self._wg_arena.numel() * 4, self.st)

    def run_backward_segment(self, k):
        """Segment k of plan_segments (k = 0 runs first): its closures, the join of its side-stream wgrads and the unpack
        of its weight gradients.  Segment 0 also does the per-pass set-up."""
        self.st = torch.cuda.current_stream(self.device).cuda_stream
        if k == 0:
            self.begin_backward()
        self._run_entries(self.segments[k][0])
        self.join_side()
        tab = self._unpack_tables[k] if self._unpack_tables else None
        if tab is not None:
            L.call("pmfb_weight_jobs", 1, tab[0].data_ptr(), tab[1], tab[2], self.st)
        return self.param_grads
