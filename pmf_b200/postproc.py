"""KNN back-projection and perspective projection on the B200 (C-ABI calls into libpmf_b200.so).

  KNN(params, nclasses).forward(proj_range, unproj_range, proj_argmax, px, py) -> LongTensor (P,)
      mirrors pc_processor/postproc/knn.py:38-143 (same ctor banner, same argument meaning: px = column index,
      py = row index, proj_range holds -1 at empty pixels, un-batched).
  project_scatter(points, labels, proj_matrix, H, W) -> dict
      the per-frame scatter of pc_processor/dataset/perspective_view_loader.py:87-131 with the projection of
      pc_processor/dataset/semantic_kitti/parser.py:209-227, done on the device.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L


def get_gaussian_kernel(kernel_size=3, sigma=2, channels=1):
    """knn.py:12-34 — a (S,S) normalised Gaussian, fp32, built on the host with the reference's operation order
    (the table has S*S <= 225 entries; it is an input of the device kernel, not part of the hot loop)."""
    x_coord = torch.arange(kernel_size)
    x_grid = x_coord.repeat(kernel_size).view(kernel_size, kernel_size)
    y_grid = x_grid.t()
    xy_grid = torch.stack([x_grid, y_grid], dim=-1).float()
    mean = (kernel_size - 1) / 2.
    variance = sigma ** 2.
    g = (1. / (2. * math.pi * variance)) * torch.exp(-torch.sum((xy_grid - mean) ** 2., dim=-1) / (2 * variance))
    g = g / torch.sum(g)
    return g.view(kernel_size, kernel_size)


class KNN(nn.Module):
    def __init__(self, params, nclasses):
        super().__init__()
        print("*" * 80)
        print("Cleaning point-clouds with kNN post-processing")
        self.knn = params["knn"]
        self.search = params["search"]
        self.sigma = params["sigma"]
        self.cutoff = params["cutoff"]
        self.nclasses = nclasses
        print("kNN parameters:")
        print("knn:", self.knn)
        print("search:", self.search)
        print("sigma:", self.sigma)
        print("cutoff:", self.cutoff)
        print("nclasses:", self.nclasses)
        print("*" * 80)
        self._tables = {}

    def _inv_gauss(self, device):
        key = (self.search, float(self.sigma), str(device))
        t = self._tables.get(key)
        if t is None:
            t = (1 - get_gaussian_kernel(self.search, self.sigma, 1)).reshape(-1).float().contiguous().to(device)
            self._tables[key] = t
        return t

    def forward(self, proj_range, unproj_range, proj_argmax, px, py):
        """Un-batched, like the reference (knn.py:56-59).  All tensors must live on the CUDA device."""
        if self.search % 2 == 0:
            raise ValueError("Nearest neighbor kernel must be odd number")  # knn.py:73-74
        for t in (proj_range, unproj_range, proj_argmax, px, py):
            if not t.is_cuda:
                raise RuntimeError("pmf_b200.KNN runs on a B200 only (no CPU fallback); got a %s tensor" % t.device)
        L.require_device()
        H, W = proj_range.shape
        P = unproj_range.shape[0]
        dev = proj_range.device
        rng = proj_range.detach().float().contiguous()
        unp = unproj_range.detach().float().contiguous()
        lab = proj_argmax.detach().long().contiguous()
        pxl = px.detach().long().contiguous()
        pyl = py.detach().long().contiguous()
        out = torch.empty((P,), device=dev, dtype=torch.int64)
        st = torch.cuda.current_stream(dev).cuda_stream
        L.call("pmfb_knn_vote", rng.data_ptr(), lab.data_ptr(), H, W, unp.data_ptr(), pxl.data_ptr(), pyl.data_ptr(), P,
               self._inv_gauss(dev).data_ptr(), int(self.search), int(self.knn), float(self.cutoff), int(self.nclasses),
               out.data_ptr(), st)
        return out


def project_scatter(points, labels, proj_matrix, H, W):
    """points: CUDA (N,4) fp32 [x,y,z,intensity]; labels: CUDA (N,) int32 (already mapped to train ids) or None;
    proj_matrix: 3x4 float64 P2 @ Tr (numpy / list / CPU tensor).  Returns a dict of CUDA tensors:
      feat (5,H,W) f32 [depth,x,y,z,i], mask (H,W) f32, label (H,W) f32, rows/cols (N,) i32 (-1 = dropped),
      depth (N,) f32, keep (N,) bool.  Collisions: the highest point index wins (numpy fancy-assignment order)."""
    if not points.is_cuda:
        raise RuntimeError("pmf_b200.project_scatter runs on a B200 only (no CPU fallback)")
    L.require_device()
    dev = points.device
    pts = points.detach().float().contiguous()
    n = pts.shape[0]
    assert pts.dim() == 2 and pts.shape[1] == 4
    lab = None if labels is None else labels.detach().to(torch.int32).contiguous()
    m = np.ascontiguousarray(np.asarray(proj_matrix, dtype=np.float64)[:3, :4]).reshape(12)
    mbuf = (C.c_double * 12)(*m.tolist())
    winner = torch.empty((H, W), device=dev, dtype=torch.int32)
    feat = torch.empty((5, H, W), device=dev, dtype=torch.float32)
    mask = torch.empty((H, W), device=dev, dtype=torch.float32)
    limg = torch.empty((H, W), device=dev, dtype=torch.float32)
    rows = torch.empty((n,), device=dev, dtype=torch.int32)
    cols = torch.empty((n,), device=dev, dtype=torch.int32)
    depth = torch.empty((n,), device=dev, dtype=torch.float32)
    st = torch.cuda.current_stream(dev).cuda_stream
    L.call("pmfb_project_scatter", pts.data_ptr(), None if lab is None else lab.data_ptr(), n, mbuf, H, W, winner.data_ptr(),
           feat.data_ptr(), mask.data_ptr(), limg.data_ptr(), rows.data_ptr(), cols.data_ptr(), depth.data_ptr(), st)
    return dict(feat=feat, mask=mask, label=limg, rows=rows, cols=cols, depth=depth, keep=rows >= 0, winner=winner)


# ------------------------------------------------------------------------------------------ batched / on-device tail
def knn_batched(knn, proj_range, unproj_range, proj_argmax, px, py, point_offsets):
    """KNN.forward for several frames in ONE launch (the reference's forward is un-batched,
    knn.py:56-59).  proj_range / proj_argmax: (F, H, W); unproj_range / px / py: the frames' point arrays concatenated
    (sum of P_f entries, same meaning as in KNN.forward: px = column, py = row); point_offsets: (F+1,) int64 with frame
    f's points at [offsets[f], offsets[f+1]).  Returns the int64 labels of all points, concatenated."""
    if knn.search % 2 == 0:
        raise ValueError("Nearest neighbor kernel must be odd number")  # knn.py:73-74
    for t in (proj_range, unproj_range, proj_argmax, px, py, point_offsets):
        if not t.is_cuda:
            raise RuntimeError("pmf_b200.knn_batched runs on a B200 only (no CPU fallback); got a %s tensor" % t.device)
    L.require_device()
    F, H, W = proj_range.shape
    dev = proj_range.device
    rng = proj_range.detach().float().contiguous()
    lab = proj_argmax.detach().long().contiguous()
    unp = unproj_range.detach().float().contiguous()
    pxl, pyl = px.detach().long().contiguous(), py.detach().long().contiguous()
    offs = point_offsets.detach().long().contiguous()
    assert offs.numel() == F + 1 and tuple(lab.shape) == (F, H, W)
    P = unp.shape[0]
    out = torch.empty((P,), device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        L.call("pmfb_knn_vote_batched", rng.data_ptr(), lab.data_ptr(), F, H, W, unp.data_ptr(), pxl.data_ptr(), pyl.data_ptr(),
               offs.data_ptr(), P, knn._inv_gauss(dev).data_ptr(), int(knn.search), int(knn.knn), float(knn.cutoff),
               int(knn.nclasses), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return out


def argmax_nchw(probs, crop=None, with_conf=False):
    """argmax over the class dimension of a dense (B, C, H, W) probability map inside ``crop = (y0, x0, h, w)`` (the
    un-padding of infer.py:107-110); first maximum wins like torch.argmax.  Returns int64 (B, h, w) [, fp32 confidence]."""
    if not probs.is_cuda:
        raise RuntimeError("pmf_b200.argmax_nchw runs on a B200 only (no CPU fallback)")
    L.require_device()
    p = probs.detach().float().contiguous()
    B, Cc, H, W = p.shape
    y0, x0, oh, ow = crop if crop is not None else (0, 0, H, W)
    label = torch.empty((B, oh, ow), device=p.device, dtype=torch.int64)
    conf = torch.empty((B, oh, ow), device=p.device, dtype=torch.float32) if with_conf else None
    with torch.cuda.device(p.device):
        L.call("pmfb_argmax_nchw", p.data_ptr(), B, Cc, H, W, int(y0), int(x0), int(oh), int(ow), label.data_ptr(),
               None if conf is None else conf.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream)
    return (label, conf) if with_conf else label


def lut_remap(labels, lut):
    """class_map_lut_inv[labels] on the device (infer.py:129): int64 training ids -> int32 dataset ids."""
    if not labels.is_cuda:
        raise RuntimeError("pmf_b200.lut_remap runs on a B200 only (no CPU fallback)")
    L.require_device()
    lab = labels.detach().long().contiguous()
    table = torch.as_tensor(np.asarray(lut), dtype=torch.int32).to(lab.device).contiguous() if not torch.is_tensor(lut) else \
        lut.to(device=lab.device, dtype=torch.int32).contiguous()
    out = torch.empty(lab.shape, device=lab.device, dtype=torch.int32)
    with torch.cuda.device(lab.device):
        L.call("pmfb_lut_remap", lab.data_ptr(), lab.numel(), table.data_ptr(), table.numel(), out.data_ptr(),
               torch.cuda.current_stream(lab.device).cuda_stream)
    return out


def merge_cameras(point_idx_list, pred_conf_list, pred_argmax_list, pc_size):
    """getMergePred of tasks/pmf_eval_nuscenes/infer.py:18-38 (a per-point Python loop in the reference): per LiDAR point
    the prediction of the camera that saw it with the highest confidence; points no camera saw get -1."""
    dev = pred_conf_list[0].device
    if dev.type != "cuda":
        raise RuntimeError("pmf_b200.merge_cameras runs on a B200 only (no CPU fallback)")
    L.require_device()
    idx = torch.cat([t.detach().long().reshape(-1) for t in point_idx_list]).contiguous()
    conf = torch.cat([t.detach().float().reshape(-1) for t in pred_conf_list]).contiguous()
    arg = torch.cat([t.detach().long().reshape(-1) for t in pred_argmax_list]).contiguous()
    cam = torch.cat([torch.full((t.numel(),), j, dtype=torch.int32, device=dev) for j, t in enumerate(point_idx_list)]).contiguous()
    scratch = torch.empty((max(int(pc_size), 1),), device=dev, dtype=torch.int64)
    merged = torch.empty((int(pc_size),), device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        L.call("pmfb_merge_cameras", idx.data_ptr(), conf.data_ptr(), arg.data_ptr(), cam.data_ptr(), idx.numel(), int(pc_size),
               scratch.data_ptr(), merged.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return merged


class InferenceTail(nn.Module):
    """Everything tasks/pmf_eval_semantickitti/infer.py:107-146 does after the forward, on the device and for a whole
    batch of frames: crop -> argmax -> KNN back-projection (or the plain gather of the default ``KNN.use: false``) ->
    class_map_lut_inv remap.  ``forward`` returns the int32 dataset-id label of every point of every frame."""

    def __init__(self, knn_params, nclasses, lut_inv, use_knn=True):
        super().__init__()
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            self.knn = KNN(knn_params, nclasses)
        self.use_knn = use_knn
        self.register_buffer("lut", torch.as_tensor(np.asarray(lut_inv), dtype=torch.int32))

    def forward(self, probs, proj_depth, unproj_range, rows, cols, point_offsets, crop=None):
        """probs: (F, C, Hp, Wp) network output; proj_depth: (F, H, W) un-normalised depth channel with 0 at empty pixels
        (turned into -1 like infer.py:85-86); unproj_range / rows / cols: concatenated per-point depth, row and column
        (uproj_depth, uproj_x_idx, uproj_y_idx of the reference loader); point_offsets: (F+1,) int64."""
        argmax = argmax_nchw(probs, crop)
        if self.use_knn:
            rng = proj_depth - proj_depth.eq(0).float()
            labels = knn_batched(self.knn, rng, unproj_range, argmax, cols, rows, point_offsets)
        else:
            F, H, W = argmax.shape
            frame = torch.bucketize(torch.arange(rows.numel(), device=rows.device), point_offsets[1:], right=True)
            labels = argmax.reshape(-1)[frame * (H * W) + rows.long() * W + cols.long()]
        return lut_remap(labels, self.lut)
