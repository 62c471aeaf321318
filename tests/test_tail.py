"""On-device inference tail: batched KNN, crop + argmax, LUT remap, 6-camera merge — bit-exact against
the oracle restatements (oracle/knn_oracle.py, oracle/tail_oracle.py).  CPU part: the merge oracle against the
reference's own getMergePred when the reference tree is present."""
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle, tail_oracle
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _merge_inputs(seed, pc_size=5000, n_cam=6, tie=False):
    rs = np.random.RandomState(seed)
    idx, conf, arg = [], [], []
    for j in range(n_cam):
        k = rs.randint(pc_size // 10, pc_size // 3)
        idx.append(np.sort(rs.choice(pc_size, k, replace=False)).astype(np.int64))
        c = rs.uniform(0.05, 1.0, k).astype(np.float32)
        if tie:
            c = np.round(c * 4) / 4  # many exact ties between cameras: the lowest camera must win
        conf.append(c)
        arg.append(rs.randint(1, 17, k).astype(np.int64))
    return idx, conf, arg, pc_size


def test_merge_oracle_equals_reference_function():
    ref = None
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        f = os.path.join(cand, "tasks", "pmf_eval_nuscenes", "infer.py")
        if os.path.exists(f):
            ref = f
            break
    if ref is None:
        pytest.skip("reference tree not present")
    # the function is pure torch apart from .cuda(): take its source text and run it on the CPU
    src = open(ref).read()
    start = src.index("def getMergePred")
    end = src.index("class Inference")
    ns = {"torch": torch}
    exec(src[start:end].replace(".cuda()", ""), ns)  # noqa: S102  (test-only: the reference's own function body)
    for seed, tie in ((1, False), (2, True)):
        idx, conf, arg, n = _merge_inputs(seed, 800, tie=tie)
        want = ns["getMergePred"]([torch.from_numpy(i) for i in idx], [torch.from_numpy(c) for c in conf],
                                  [torch.from_numpy(a) for a in arg], n).numpy()
        assert np.array_equal(tail_oracle.merge_cameras(idx, conf, arg, n), want)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_batched_knn_equals_per_frame_oracle(dev):
    import pmf_b200
    for search, knn_k, frames in ((5, 5, 4), (11, 5, 3), (3, 1, 2), (7, 8, 2)):
        H, W = 64, 96
        pr, am, ur, px, py, offs, want = [], [], [], [], [], [0], []
        for f in range(frames):
            case = dict(name="b", H=H, W=W, P=700 + 333 * f, knn=knn_k, search=search, sigma=1.0, cutoff=1.0, nclasses=20, empty=0.7,
                        seed=100 + f + search, kind="rand" if f % 2 == 0 else "ties")
            inp = synth.knn_inputs(case)
            pr.append(inp["proj_range"]); am.append(inp["proj_argmax"]); ur.append(inp["unproj_range"])
            px.append(inp["px"]); py.append(inp["py"]); offs.append(offs[-1] + case["P"])
            want.append(knn_oracle.knn_vote(inp["proj_range"], inp["unproj_range"], inp["proj_argmax"], inp["px"], inp["py"], knn_k,
                                            search, 1.0, 1.0, 20))
        knn = pmf_b200.KNN(dict(knn=knn_k, search=search, sigma=1.0, cutoff=1.0), 20)
        t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)  # noqa: E731
        got = pmf_b200.knn_batched(knn, t(np.stack(pr)), t(np.concatenate(ur)), t(np.stack(am)), t(np.concatenate(px)),
                                   t(np.concatenate(py)), t(np.asarray(offs, np.int64)))
        assert got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), np.concatenate(want)), (search, knn_k)
        # the single-frame entry point runs the same kernel
        one = knn(t(pr[0]), t(ur[0]), t(am[0]), t(px[0]), t(py[0]))
        assert np.array_equal(one.cpu().numpy(), want[0])
    # cutoff disabled (labels of infinite-distance neighbours vote) and k > 8 (warp kernel)
    for knn_k, cutoff in ((5, 0.0), (12, 1.0)):
        case = dict(name="c", H=48, W=64, P=900, knn=knn_k, search=5, sigma=1.0, cutoff=cutoff, nclasses=20, empty=0.8, seed=7, kind="rand")
        inp = synth.knn_inputs(case)
        knn = pmf_b200.KNN(dict(knn=knn_k, search=5, sigma=1.0, cutoff=cutoff), 20)
        got = knn(*[torch.from_numpy(inp[k]).to(dev) for k in ("proj_range", "unproj_range", "proj_argmax", "px", "py")])
        want = knn_oracle.knn_vote(inp["proj_range"], inp["unproj_range"], inp["proj_argmax"], inp["px"], inp["py"], knn_k, 5, 1.0,
                                   cutoff, 20)
        assert np.array_equal(got.cpu().numpy(), want), (knn_k, cutoff)


@pytest.mark.gpu
def test_argmax_lut_merge_and_tail(dev):
    import pmf_b200
    rs = np.random.RandomState(3)
    probs = rs.rand(3, 20, 40, 56).astype(np.float32)
    probs[0, 4] = probs[0, 9]  # exact ties between classes: the first maximum wins
    p = torch.from_numpy(probs).to(dev)
    lab, conf = pmf_b200.argmax_nchw(p, crop=(4, 8, 32, 40), with_conf=True)
    crop = probs[:, :, 4:36, 8:48]
    assert np.array_equal(lab.cpu().numpy(), crop.argmax(1)) and np.array_equal(conf.cpu().numpy(), crop.max(1))
    assert np.array_equal(pmf_b200.argmax_nchw(p).cpu().numpy(), probs.argmax(1))
    lut = rs.randint(0, 260, 20).astype(np.int32)
    got = pmf_b200.lut_remap(lab, lut)
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), lut[crop.argmax(1)])
    for seed, tie in ((5, False), (6, True)):
        idx, cf, arg, n = _merge_inputs(seed, tie=tie)
        got = pmf_b200.merge_cameras([torch.from_numpy(i).to(dev) for i in idx], [torch.from_numpy(c).to(dev) for c in cf],
                                     [torch.from_numpy(a).to(dev) for a in arg], n)
        assert np.array_equal(got.cpu().numpy(), tail_oracle.merge_cameras(idx, cf, arg, n))
    # whole tail, 3 frames, against the per-frame oracle (crop of a padded output, KNN and plain-gather variants)
    H, W, hp, wp = 32, 40, 4, 8
    params = dict(knn=5, search=5, sigma=1.0, cutoff=1.0)
    depth, ur, rows, cols, offs = [], [], [], [], [0]
    for f in range(3):
        d = rs.uniform(2, 50, (H, W)).astype(np.float32) * (rs.rand(H, W) < 0.4)
        n_pts = 300 + 50 * f
        r_, c_ = rs.randint(0, H, n_pts), rs.randint(0, W, n_pts)
        depth.append(d.astype(np.float32)); rows.append(r_.astype(np.int32)); cols.append(c_.astype(np.int32))
        ur.append((d[r_, c_] + rs.normal(0, 0.05, n_pts) + (d[r_, c_] == 0) * 10).astype(np.float32)); offs.append(offs[-1] + n_pts)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dev)  # noqa: E731
    for use_knn in (True, False):
        tail = pmf_b200.InferenceTail(params, 20, lut, use_knn=use_knn).to(dev)
        got = tail(p, t(np.stack(depth)), t(np.concatenate(ur)), t(np.concatenate(rows)), t(np.concatenate(cols)),
                   t(np.asarray(offs, np.int64)), crop=(hp, wp, H, W)).cpu().numpy()
        want = np.concatenate([tail_oracle.semantic_tail(probs[f], depth[f], ur[f], rows[f], cols[f], lut, params, 20,
                                                         crop=(hp, wp, H, W), use_knn=use_knn) for f in range(3)])
        assert np.array_equal(got, want), use_knn
