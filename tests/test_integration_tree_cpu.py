"""INTEGRATION.md section 1, exercised on the CPU box: the merged tree — our ``pc_processor`` shim (models, postproc,
dataset/perspective_view_loader) over the reference's own host-side files — is imported THE WAY THE TASK SCRIPTS DO IT: a
bare ``import pc_processor`` followed by attribute access (tasks/pmf/main.py:9,16,27,34; trainer.py:19,36,50,61,103,139,
190,202), the module tree the trainer touches exists under the reference's names, and the reference's loss block
ingredients accept what our modules return."""
import os
import subprocess
import sys
import textwrap

import pytest

from tests import merged_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(merged_tree.reference_root() is None, reason="no reference tree (/root/reference or baseline/_ref)")
def test_merged_tree_imports_like_the_task_scripts(tmp_path):
    top = merged_tree.build(tmp_path)
    code = textwrap.dedent("""
        import os, sys, torch
        sys.path.insert(0, %r)           # third-party stubs (tensorboardX, nuscenes, ...): not installed in this image
        os.chdir(%r)
        sys.path.insert(0, "../../")     # tasks/pmf/option.py:6
        import pc_processor              # bare import, as main.py:9 / trainer.py:8 / option.py:8
        # every attribute the unchanged task scripts reach through the bare import
        pc_processor.utils.init_distributed_mode; pc_processor.utils.is_main_process; pc_processor.utils.RemainTime
        pc_processor.utils.WarmupCosineLR; pc_processor.utils.AverageMeter
        pc_processor.checkpoint.Recorder
        pc_processor.layers.sync_bn.replaceBN
        pc_processor.loss.Lovasz_softmax; pc_processor.loss.FocalSoftmaxLoss
        pc_processor.metrics.IOUEval
        pc_processor.dataset.semantic_kitti.SemanticKitti; pc_processor.dataset.nuScenes.Nuscenes
        pc_processor.dataset.PerspectiveViewLoader
        pc_processor.models.SalsaNext; pc_processor.models.EPMFNet
        pc_processor.postproc.KNN
        from pc_processor.models.pmf_net import ResidualBasedFusionBlock      # pc_processor/models/epmf_net.py:8
        ours = %r
        for mod in (pc_processor, pc_processor.models, pc_processor.models.pmf_net, pc_processor.postproc.knn,
                    pc_processor.dataset, pc_processor.dataset.perspective_view_loader):
            assert os.path.realpath(mod.__file__).startswith(ours), mod.__file__
        for mod in (pc_processor.loss, pc_processor.metrics, pc_processor.layers, pc_processor.utils, pc_processor.checkpoint,
                    pc_processor.dataset.semantic_kitti, pc_processor.models.salsanext):
            assert not os.path.realpath(mod.__file__).startswith(ours), mod.__file__
        assert pc_processor.dataset.PerspectiveViewLoader.__module__ == "pmf_b200.loader"
        m = pc_processor.models.PMFNet(pcd_channels=5, img_channels=3, nclasses=20, base_channels=32,
                                       image_backbone="resnet34", imagenet_pretrained=False)   # main.py:34-41
        # trainer.py:82-89: the three parameter groups of the two optimisers
        n = [sum(p.numel() for p in g.parameters()) for g in (m.lidar_stream, m.camera_stream_encoder, m.camera_stream_decoder)]
        assert 36.3e6 < sum(n) < 36.5e6, n   # 36.416 M parameters (SURVEY.md Appendix B)
        # trainer.py:36: replaceBN walks named_children / add_module; key set and strict load survive it
        keys = list(m.state_dict().keys())
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        m2 = pc_processor.layers.sync_bn.replaceBN(m)
        assert list(m2.state_dict().keys()) == keys
        m2.load_state_dict(sd, strict=True)
        assert type(m2.lidar_stream.downCntx.bn1).__name__ == "SynchronizedBatchNorm2d"
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
    """) % (os.path.join(top, "stubs"), os.path.join(top, "tasks", "pmf"), ROOT + os.sep)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "merged tree ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
