"""INTEGRATION.md section 1, exercised on the CPU box: a merged tree — our ``pc_processor`` shim (models, postproc) next to
the reference's own host-side sub-packages (loss, metrics, layers, utils) — imports the way the unchanged task scripts
import it, the module trees the trainer touches exist under the reference's names, and the reference's loss block
(tasks/pmf/trainer.py:305-332 ingredients: FocalSoftmaxLoss, Lovasz_softmax, IOUEval) accepts what our modules return."""
import os
import subprocess
import sys
import textwrap

import pytest

from oracle.ref_loader import REFERENCE_ROOT, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_merged_tree_imports_like_the_task_scripts(tmp_path):
    pkg = tmp_path / "pc_processor"
    pkg.mkdir()
    for name in ("__init__.py", "models", "postproc"):          # ours
        os.symlink(os.path.join(ROOT, "pc_processor", name), pkg / name)
    for name in ("loss", "metrics", "layers", "utils"):          # the reference's, unchanged
        os.symlink(os.path.join(REFERENCE_ROOT, "pc_processor", name), pkg / name)
    os.symlink(os.path.join(ROOT, "pmf_b200"), tmp_path / "pmf_b200")
    code = textwrap.dedent("""
        import os, sys, torch
        sys.path.insert(0, %r)
        import pc_processor
        import pc_processor.loss, pc_processor.metrics, pc_processor.layers, pc_processor.utils
        from pc_processor.models import PMFNet, EPMFNet                      # tasks/pmf/main.py:34, tasks/epmf/main.py:44
        from pc_processor.models.pmf_net import ResidualBasedFusionBlock      # pc_processor/models/epmf_net.py:8
        from pc_processor.postproc import KNN                                 # tasks/pmf_eval_semantickitti/infer.py:24
        assert os.path.realpath(pc_processor.models.__file__).startswith(%r)
        assert "reference" in os.path.realpath(pc_processor.loss.__file__)
        m = PMFNet(pcd_channels=5, img_channels=3, nclasses=20, base_channels=32, image_backbone="resnet34",
                   imagenet_pretrained=False)
        # trainer.py:82-89: the three parameter groups of the two optimisers
        n = [sum(p.numel() for p in g.parameters()) for g in (m.lidar_stream, m.camera_stream_encoder, m.camera_stream_decoder)]
        assert 36.3e6 < sum(n) < 36.5e6, n   # 36.416 M parameters (SURVEY.md Appendix B)
        # the reference's loss ingredients run on (B, C, H, W) probability maps such as our modules return
        p = torch.softmax(torch.randn(2, 20, 16, 32), 1).requires_grad_(True)
        y = torch.randint(0, 20, (2, 16, 32))
        focal = pc_processor.loss.FocalSoftmaxLoss(20, gamma=2, alpha=[1.0] * 20, softmax=False)   # trainer.py:202-204
        lovasz = pc_processor.loss.Lovasz_softmax(ignore=0)
        loss = focal(p, y, mask=(y > 0).float()) + lovasz(p, y)
        loss.backward()
        assert torch.isfinite(p.grad).all()
        ev = pc_processor.metrics.IOUEval(n_classes=20, device=torch.device("cpu"), ignore=[0])
        ev.addBatch(p.argmax(1), y)
        assert 0.0 <= float(ev.getIoU()[0]) <= 1.0
        print("merged tree ok")
    """) % (str(tmp_path), os.path.join(ROOT, "pc_processor"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "merged tree ok" in r.stdout, r.stderr[-3000:]
