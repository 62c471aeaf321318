"""numpy restatement of the perspective projection + scatter.  TEST INFRASTRUCTURE.

Follows pc_processor/dataset/semantic_kitti/parser.py:209-227 (``SemanticKitti.mapLidar2Camera``) and
pc_processor/dataset/perspective_view_loader.py:87-131 (``PerspectiveViewLoader.__getitem__`` scatter).
Pinned against those two functions run from /root/reference (tests/test_oracle_pinning.py,
tests/golden/project_*.npz).
"""
import numpy as np


def map_lidar_to_camera(proj_matrix, xyz, img_h_px, img_w_px):
    """parser.py:209-227.  proj_matrix (3,4) float64 = P2 @ Tr; xyz (N,3) float32.

    NOTE the reference's argument order quirk (parser.py:222-223 vs the call site at
    perspective_view_loader.py:89-90): the caller passes (image.shape[1], image.shape[0]) = (W, H) into
    parameters named (img_h, img_w); column u is tested against the first, row v against the second.
    Here the parameters are named for what they bound: u < img_w_px is expressed as ``u < first arg``.
    Returns (rows_cols (M,2) float64 [row=v, col=u], keep_mask (N,) bool)."""
    keep = xyz[:, 0] > 0.5
    hc = np.concatenate([xyz[keep], np.ones([keep.sum(), 1], dtype=np.float32)], axis=1)
    mapped = (proj_matrix @ hc.T).T  # float64 because proj_matrix is float64
    mapped = mapped[:, :2] / np.expand_dims(mapped[:, 2], axis=1)
    ok = (mapped[:, 0] > 0) * (mapped[:, 0] < img_h_px) * (mapped[:, 1] > 0) * (mapped[:, 1] < img_w_px)
    keep = keep.copy()
    keep[keep] = ok
    mapped = np.fliplr(mapped)
    return mapped[ok], keep


def project_scatter(proj_matrix, pointcloud, labels, H, W):
    """perspective_view_loader.py:87-131 (without RGB / augmentation).

    pointcloud (N,4) float32 [x,y,z,intensity]; labels (N,) int32 already mapped to train ids.
    Returns dict: depth (H,W) f32, xyzi (H,W,4) f32, label (H,W) i32, mask (H,W) i32,
    rows (M,) i32, cols (M,) i32, keep (N,) bool, point_depth (N,) f32.
    Collisions: numpy fancy assignment — the LAST point in input order wins."""
    mapped, keep = map_lidar_to_camera(proj_matrix, pointcloud[:, :3], W, H)
    rows = mapped[:, 0].astype(np.int32)  # truncation
    cols = mapped[:, 1].astype(np.int32)
    depth_pts = np.linalg.norm(pointcloud[:, :3], 2, axis=1)
    kept = pointcloud[keep]
    xyzi = np.zeros((H, W, kept.shape[1]), dtype=np.float32)
    xyzi[rows, cols] = kept
    depth = np.zeros((H, W), dtype=np.float32)
    depth[rows, cols] = depth_pts[keep]
    label = np.zeros((H, W), dtype=np.int32)
    label[rows, cols] = labels[keep]
    mask = np.zeros((H, W), dtype=np.int32)
    mask[rows, cols] = 1
    return dict(depth=depth, xyzi=xyzi, label=label, mask=mask, rows=rows, cols=cols, keep=keep,
                point_depth=depth_pts.astype(np.float32))
