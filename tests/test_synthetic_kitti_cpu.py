"""the synthetic SemanticKITTI-layout dataset (tools/make_synthetic_kitti.py) is readable by the
reference's own loader, and the loader's projected frame equals the oracle / device projection contract."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import make_synthetic_kitti as mk  # noqa: E402
from oracle import project_oracle  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_reference_file, reference_available  # noqa: E402
from tests import synth  # noqa: E402


def test_layout_round_trip(tmp_path):
    H, W = 48, 96
    frames = mk.write_sequence(str(tmp_path), 3, 2, H, W, rows=16, cols=256, seed=2)
    sdir = tmp_path / "03"
    assert sorted(os.listdir(sdir)) == ["calib.txt", "image_2", "labels", "velodyne"]
    for f, (pts, lab) in enumerate(frames):
        p = np.fromfile(sdir / "velodyne" / ("%06d.bin" % f), dtype=np.float32).reshape(-1, 4)
        raw = np.fromfile(sdir / "labels" / ("%06d.label" % f), dtype=np.int32)
        assert np.array_equal(p, pts)
        assert np.array_equal(raw & 0xFFFF, mk.RAW_ID[lab].astype(np.int32)) and not (raw >> 16).any()
    calib = {ln.split(":")[0]: np.array([float(x) for x in ln.split(":", 1)[1].split()]) for ln in open(sdir / "calib.txt")}
    Tr = np.identity(4)
    Tr[:3] = calib["Tr"].reshape(3, 4)
    assert np.allclose((calib["P2"].reshape(3, 4) @ Tr)[:3], synth.camera_matrix(H, W), rtol=1e-11, atol=1e-11)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_reference_parser_reads_it_and_projection_matches_oracle(tmp_path):
    """The UNMODIFIED reference parser opens the synthetic sequence; its mapLidar2Camera on our files gives the pixel
    coordinates the projection oracle (and therefore pmfb_project_scatter, bit-exact to it) produces."""
    H, W = 48, 96
    frames = mk.write_sequence(str(tmp_path), 0, 2, H, W, rows=32, cols=512, seed=5)
    parser = load_reference_file("pc_processor/dataset/semantic_kitti/parser.py", "ref_parser_synth")
    cfg = os.path.join(REFERENCE_ROOT, "pc_processor/dataset/semantic_kitti/semantic-kitti.yaml")
    ds = parser.SemanticKitti(str(tmp_path), [0], cfg, has_image=True, has_pcd=True, has_label=True)
    assert len(ds.pointcloud_files) == 2 and len(ds.label_files) == 2 and len(ds.image_files) == 2
    assert np.allclose(ds.proj_matrix["00"][:3], synth.camera_matrix(H, W), rtol=1e-11, atol=1e-11)
    for idx, (pts, lab) in enumerate(frames):
        pc, sem, inst = ds.loadDataByIndex(idx)
        assert np.array_equal(pc, pts)
        assert np.array_equal(ds.labelMapping(sem), lab)  # raw ids map back to the 20 training classes
        assert ds.loadImage(idx).size == (W, H)
        seq_id, frame_id = ds.parsePathInfoByIndex(idx)
        assert seq_id == "00" and frame_id == "%06d" % idx
        # the loader's call (perspective_view_loader.py:89-90) passes (image width, image height)
        mapped, keep = ds.mapLidar2Camera(seq_id, pc[:, :3], W, H)
        ref = project_oracle.project_scatter(ds.proj_matrix[seq_id][:3], pc, ds.labelMapping(sem), H, W)
        rows, cols = mapped[:, 0].astype(np.int32), mapped[:, 1].astype(np.int32)
        assert np.array_equal(keep, ref["keep"]) and int(keep.sum()) > 100
        assert np.array_equal(rows, ref["rows"]) and np.array_equal(cols, ref["cols"])
