"""Deterministic synthetic inputs shared by the fixture generator, the tests and bench.py.

Everything is drawn from ``numpy.random.RandomState`` (frozen legacy generator) so the same arrays are
produced in the build container and on the GPU box.
"""
import numpy as np
import torch

# ----------------------------------------------------------------------------- KNN (knn.py:55-143)
KNN_CASES = [
    dict(name="rand_s5", H=32, W=128, P=3000, knn=5, search=5, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.3, seed=11, kind="rand"),
    dict(name="rand_s11", H=32, W=96, P=2000, knn=5, search=11, sigma=1.0, cutoff=1.0, nclasses=17, empty=0.5, seed=12, kind="rand"),
    dict(name="sparse", H=24, W=64, P=1500, knn=5, search=5, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.95, seed=13, kind="rand"),
    dict(name="border", H=16, W=48, P=256, knn=5, search=5, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.2, seed=14, kind="border"),
    dict(name="cutoff0", H=32, W=64, P=1000, knn=7, search=7, sigma=2.0, cutoff=0.0, nclasses=20, empty=0.3, seed=15, kind="rand"),
    dict(name="ties", H=16, W=32, P=512, knn=5, search=5, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.3, seed=16, kind="ties"),
    dict(name="k1", H=16, W=32, P=300, knn=1, search=3, sigma=1.0, cutoff=0.5, nclasses=5, empty=0.1, seed=17, kind="rand"),
]


def knn_inputs(case):
    rs = np.random.RandomState(case["seed"])
    H, W, P, C = case["H"], case["W"], case["P"], case["nclasses"]
    yy, xx = np.mgrid[0:H, 0:W]
    base = 10.0 + 6.0 * np.sin(xx / 9.0) + 3.0 * np.cos(yy / 5.0)
    if case["kind"] == "ties":
        rng = np.round(base).astype(np.float32)  # piecewise-constant -> many exact distance ties
    else:
        rng = (base + rs.normal(0, 0.4, (H, W))).astype(np.float32)
    empty = rs.rand(H, W) < case["empty"]
    rng[empty] = -1.0  # infer.py:85-86: empty pixels are -1
    lab = rs.randint(0, C, (H, W)).astype(np.int64)
    lab[empty] = 0
    if case["kind"] == "border":
        side = rs.randint(0, 4, P)
        py = np.where(side == 0, 0, np.where(side == 1, H - 1, rs.randint(0, H, P)))
        px = np.where(side == 2, 0, np.where(side == 3, W - 1, rs.randint(0, W, P)))
        corners = np.array([[0, 0], [0, W - 1], [H - 1, 0], [H - 1, W - 1]])
        py[:4], px[:4] = corners[:, 0], corners[:, 1]
    else:
        py = rs.randint(0, H, P)
        px = rs.randint(0, W, P)
    pix = np.where(rng[py, px] > 0, rng[py, px], base[py, px].astype(np.float32))
    if case["kind"] == "ties":
        unproj = pix.astype(np.float32)
    else:
        unproj = (pix + rs.normal(0, 0.3, P)).astype(np.float32)
    return dict(proj_range=rng, unproj_range=unproj, proj_argmax=lab, px=px.astype(np.int64), py=py.astype(np.int64))


# ----------------------------------------------------------------------------- synthetic LiDAR sweep + camera
def lidar_sweep(n_rows=64, n_cols=2048, seed=1, elev_deg=(-25.0, 3.0)):
    """Organised scan: azimuth uniform in [-pi,pi), n_rows elevations, range from a ground plane + boxes.
    Returns (N,4) float32 [x,y,z,intensity] and (N,) int32 labels in 1..19."""
    rs = np.random.RandomState(seed)
    az = -np.pi + 2 * np.pi * (np.arange(n_cols) + 0.5) / n_cols
    el = np.deg2rad(np.linspace(elev_deg[1], elev_deg[0], n_rows))
    azg, elg = np.meshgrid(az, el)
    dx, dy, dz = np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)
    with np.errstate(divide="ignore"):
        t_ground = np.where(dz < -1e-3, -1.73 / dz, 80.0)
    rng = np.clip(t_ground, 1.0, 80.0)
    lab = np.where(rng < 80.0, 9, 15).astype(np.int32)  # road / vegetation-ish background
    K = 32
    centers = np.stack([rs.uniform(4, 45, K) * np.cos(a) for a in [rs.uniform(-np.pi, np.pi, K)]], 0)[0]
    ang = rs.uniform(-np.pi, np.pi, K)
    dist = rs.uniform(4, 45, K)
    cx, cy = dist * np.cos(ang), dist * np.sin(ang)
    half = rs.uniform(0.6, 2.5, K)
    height = rs.uniform(0.5, 3.0, K)
    cls = rs.randint(1, 20, K)
    for k in range(K):
        # ray / vertical-cylinder intersection (cheap box stand-in), closest hit wins
        bx, by = cx[k], cy[k]
        a = dx ** 2 + dy ** 2
        b = -2 * (dx * bx + dy * by)
        c = bx ** 2 + by ** 2 - half[k] ** 2
        disc = b ** 2 - 4 * a * c
        with np.errstate(invalid="ignore"):
            t = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a)
        z_hit = t * dz
        hit = (disc > 0) & (t > 1.0) & (t < rng) & (z_hit > -1.73) & (z_hit < -1.73 + height[k])
        rng = np.where(hit, t, rng)
        lab = np.where(hit, cls[k], lab)
    del centers
    rng = np.clip(rng + rs.normal(0, 0.02, rng.shape), 1.0, 80.0)
    pts = np.stack([rng * dx, rng * dy, rng * dz, rs.uniform(0, 1, rng.shape)], -1).reshape(-1, 4).astype(np.float32)
    return pts, lab.reshape(-1).astype(np.int32)


def camera_calibration(H, W, fov_scale=0.5625):
    """KITTI-like calibration scaled to an HxW image: (P2 3x4, Tr 4x4), float64.  fx=fy=fov_scale*W, principal point
    centred; Tr maps velodyne (x fwd, y left, z up) to camera (x right, y down, z fwd) with a small translation."""
    fx = fy = fov_scale * W
    P2 = np.array([[fx, 0, W / 2.0, 4.5e1 * W / 1242.0], [0, fy, H / 2.0, -0.3], [0, 0, 1.0, 0.003]], np.float64)
    Tr = np.array([[0, -1, 0, 0.004], [0, 0, -1, -0.076], [1, 0, 0, -0.272], [0, 0, 0, 1]], np.float64)
    return P2, Tr


def camera_matrix(H, W, fov_scale=0.5625):
    """P2 @ Tr (3x4, float64) of camera_calibration: what parser.py:73-75 hands to mapLidar2Camera."""
    P2, Tr = camera_calibration(H, W, fov_scale)
    return (P2 @ Tr)[:3]


PROJECT_CASES = [
    dict(name="small", H=48, W=96, rows=32, cols=512, seed=3),
    dict(name="wide", H=64, W=208, rows=64, cols=1024, seed=4),
]


def project_inputs(case):
    pts, lab = lidar_sweep(case["rows"], case["cols"], seed=case["seed"])
    return dict(pointcloud=pts, labels=lab, proj_matrix=camera_matrix(case["H"], case["W"]))


# ----------------------------------------------------------------------------- network inputs
FEATURE_MEAN = np.array([12.12, 10.88, 0.23, -1.04, 0.21], np.float32)  # config_server_kitti.yaml:80-91
FEATURE_STD = np.array([12.32, 11.47, 6.91, 0.86, 0.16], np.float32)


def frame_tensor(B, H, W, seed, density=0.35):
    """(B,8,H,W) input_feature as trainer.py:291-297 sees it AFTER normalisation, plus mask and labels.
    ch0:5 = normalised [depth,x,y,z,i] * mask (exactly 0 on empty pixels), ch5:8 = RGB in [0,1]."""
    rs = np.random.RandomState(seed)
    mask = (rs.rand(B, H, W) < density).astype(np.float32)
    depth = rs.uniform(2, 60, (B, H, W)).astype(np.float32)
    xyz = rs.normal(0, 1, (B, 3, H, W)).astype(np.float32) * np.array([20, 12, 1.0], np.float32)[None, :, None, None]
    inten = rs.uniform(0, 1, (B, 1, H, W)).astype(np.float32)
    pcd = np.concatenate([depth[:, None], xyz, inten], 1)
    pcd = (pcd - FEATURE_MEAN[None, :, None, None]) / FEATURE_STD[None, :, None, None] * mask[:, None]
    rgb = rs.uniform(0, 1, (B, 3, H + 2, W + 2)).astype(np.float32)
    rgb = sum(rgb[:, :, i:i + H, j:j + W] for i in range(3) for j in range(3)) / 9.0  # 3x3 box low-pass
    feat = np.concatenate([pcd, rgb], 1).astype(np.float32)
    label = (rs.randint(1, 20, (B, H, W)) * mask).astype(np.int64)
    return torch.from_numpy(feat), torch.from_numpy(mask), torch.from_numpy(label)


FUSION_CASES = [
    dict(name="c64", pcd_c=64, img_c=64, B=2, H=16, W=32, seed=21),
    dict(name="c256x512", pcd_c=256, img_c=512, B=1, H=4, W=24, seed=22),
]


def fusion_inputs(case):
    rs = np.random.RandomState(case["seed"])
    pcd = rs.normal(0, 1, (case["B"], case["pcd_c"], case["H"], case["W"])).astype(np.float32)
    img = np.maximum(rs.normal(0, 1, (case["B"], case["img_c"], case["H"], case["W"])), 0).astype(np.float32)
    return torch.from_numpy(pcd), torch.from_numpy(img)


PMF_CASES = [
    dict(name="r34_small", backbone="resnet34", nclasses=20, B=2, H=32, W=64, seed=1),
]
EPMF_CASES = [
    dict(name="r34_small", backbone="resnet34", nclasses=20, B=2, H=32, W=64, seed=4, density=0.15),
]


def epmf_inputs(case):
    feat, _, _ = frame_tensor(case["B"], case["H"], case["W"], seed=case["seed"] + 200, density=case["density"])
    return feat[:, 0:5].contiguous(), feat[:, 5:8].contiguous()


PMF_GRAD_PICKS = [
    "lidar_stream.downCntx.conv1.weight", "lidar_stream.downCntx.conv2.weight", "camera_stream_encoder.conv1.weight",
    "lidar_stream.fusionblock_1.fuse_conv.2.weight", "lidar_stream.resBlock1.conv4.weight",
    "camera_stream_encoder.layer2.0.downsample.0.weight", "lidar_stream.upBlock4.conv1.weight",
    "camera_stream_decoder.conv.bias", "lidar_stream.logits.weight", "lidar_stream.aspp.conv.bias",
]
PMF_STAT_PICKS = [
    "lidar_stream.downCntx.bn1.running_mean", "lidar_stream.downCntx.bn1.running_var",
    "camera_stream_encoder.bn1.running_var", "lidar_stream.fusionblock_4.attention.4.running_mean",
    "camera_stream_decoder.up_1a.2.running_var",
]


def pmf_inputs(case):
    feat, _, _ = frame_tensor(case["B"], case["H"], case["W"], seed=100 + case["seed"])
    # channel-slice views of one tensor, as trainer.py:296-297 passes them
    return feat[:, 0:5], feat[:, 5:8]


def pmf_loss_weights(case):
    rs = np.random.RandomState(200 + case["seed"])
    shp = (case["B"], case["nclasses"], case["H"], case["W"])
    return torch.from_numpy(rs.normal(0, 1, shp).astype(np.float32)), torch.from_numpy(rs.normal(0, 1, shp).astype(np.float32))
