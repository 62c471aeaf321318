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
