"""SURVEY.md §8f-4 (-m gpu): IOUEval on the device against the reference's formulas, and the frozen / folded inference
export against the live eval forward."""
import os

import numpy as np
import pytest
import torch

from tests import synth
from tests.test_gpu_parity import _miou

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    return torch.device("cuda:0")


def test_iou_eval_on_device_matches_reference_formulas(dev):
    from pmf_b200.metrics import IOUEval
    rs = np.random.RandomState(0)
    ev = IOUEval(n_classes=20, device=dev, ignore=[0], is_distributed=False)
    conf = np.zeros((20, 20), np.int64)
    preds, labels = [], []
    for _ in range(3):
        p, t = rs.randint(0, 20, (2, 48, 64)), rs.randint(0, 20, (2, 48, 64))
        ev.addBatch(torch.from_numpy(p).to(dev), torch.from_numpy(t).to(dev))
        np.add.at(conf, (p.reshape(-1), t.reshape(-1)), 1)
        preds.append(p)
        labels.append(t)
    assert np.array_equal(ev.conf_matrix.cpu().numpy(), conf)  # iou_eval.py:52-53 (rows = prediction, columns = target)
    miou, iou = ev.getIoU()
    assert abs(float(miou) - _miou(np.concatenate(preds), np.concatenate(labels), 20)) < 1e-12
    c = conf.astype(np.float64)
    c[0, :] = 0
    c[:, 0] = 0
    tp, fp, fn = np.diag(c), c.sum(1) - np.diag(c), c.sum(0) - np.diag(c)
    macc, acc = ev.getAcc()
    mrec, rec = ev.getRecall()
    assert np.allclose(acc.cpu().numpy(), tp / (tp + fp + 1e-15)) and np.allclose(rec.cpu().numpy(), tp / (tp + fn + 1e-15))
    assert abs(float(macc) - (tp / (tp + fp + 1e-15))[1:].mean()) < 1e-12 and abs(float(mrec) - (tp / (tp + fn + 1e-15))[1:].mean()) < 1e-12
    ev.reset()
    assert int(ev.conf_matrix.sum()) == 0
    ev.addBatch(torch.full((7,), 25, device=dev), torch.zeros(7, dtype=torch.long, device=dev))  # out-of-range ids are skipped
    assert int(ev.conf_matrix.sum()) == 0


def test_frozen_inference_and_folded_export(dev, tmp_path):
    import pmf_b200
    from pmf_b200 import export
    torch.manual_seed(3)
    m = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34")
    # non-trivial running statistics so that the folded affines matter
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.uniform_(-0.2, 0.2)
            mod.running_var.uniform_(0.5, 1.5)
    m.to(dev).eval()
    feat, _, _ = synth.frame_tensor(2, 64, 96, seed=8)
    x = feat.to(dev)
    with torch.no_grad():
        live = [m(x[:, 0:5], x[:, 5:8]) for _ in range(2)][1]      # eager, then captured
        m.freeze()
        frozen = [m(x[:, 0:5], x[:, 5:8]) for _ in range(3)]         # eager (packs + folds once), capture, replay
    for f in frozen:
        assert torch.equal(f[0], live[0]) and torch.equal(f[1], live[1])
    live_calls = None
    frozen_calls = [r.n_fwd_calls for r in m._graphs.values()][0]
    m.freeze(False)
    with torch.no_grad():
        m(x[:, 0:5], x[:, 5:8])
        m(x[:, 0:5], x[:, 5:8])
    live_calls = [r.n_fwd_calls for r in m._graphs.values()][0]
    assert frozen_calls < live_calls - 90, (frozen_calls, live_calls)  # no weight packing / 94 BatchNorm finalisations in the graph
    # frozen really is frozen: a parameter change is not seen until the module is re-frozen
    m.freeze()
    with torch.no_grad():
        a = m(x[:, 0:5], x[:, 5:8])[0]
        m.lidar_stream.logits.weight.data.mul_(1.5)
        b = m(x[:, 0:5], x[:, 5:8])[0]
        assert torch.equal(a, b)
        m.freeze()
        c = m(x[:, 0:5], x[:, 5:8])[0]
        assert not torch.equal(a, c)
        m.lidar_stream.logits.weight.data.div_(1.5)
    # on-disk form: save, load into a fresh process-independent module, same outputs; state_dict stays reference-shaped
    m.freeze(False)
    path = os.path.join(str(tmp_path), "pmf_folded.pt")
    blob = export.save_folded(m, path)
    assert blob["format"] == export.FORMAT and len(blob["state_dict"]) == 654 and len(blob["conv"]) == 110
    m2 = export.load_folded(path, dev)
    with torch.no_grad():
        out2 = m2(x[:, 0:5], x[:, 5:8])
        ref = m(x[:, 0:5], x[:, 5:8])
    assert torch.equal(out2[0], ref[0]) and torch.equal(out2[1], ref[1])
