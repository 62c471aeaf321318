"""N>1 path on CPU: two gloo ranks drive pmf_b200.PMFNet (numpy C-ABI model underneath, tests/cabi_mock.py) wrapped in
DistributedDataParallel as tasks/pmf/trainer.py:38-39 does; the all-reduced gradients must equal the mean of the
per-rank gradients computed in one process, and sharding helpers must partition the frames."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pmf_b200 import dist as pdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(seed_weights=1):
    from oracle import pmf_oracle as po
    from pmf_b200 import modules as M
    torch.manual_seed(0)
    m = M.PMFNet(5, 3, 20, 32, False, "resnet34")
    m.load_state_dict(po.synth_state_dict(po.pmf_param_shapes(20, 32, "resnet34"), seed=seed_weights))
    m.train()
    m._dropout_override = False
    return m


def _frames(rank):
    from tests import synth
    feat, _, label = synth.frame_tensor(1, 16, 32, seed=pdist.shard_seed(1, rank))
    return feat, label


def _loss(lid, cam, label):
    t = label.unsqueeze(1)
    return -(torch.log(lid.gather(1, t).clamp_min(1e-8)).mean() + torch.log(cam.gather(1, t).clamp_min(1e-8)).mean())


def _install_mock():
    from _pytest.monkeypatch import MonkeyPatch
    from pmf_b200 import modules as M
    from tests import cabi_mock
    mpatch = MonkeyPatch()
    cabi_mock.install(mpatch, exact=True)
    mpatch.setattr(M, "_require_cuda", lambda *t: None)
    return mpatch


def _worker(rank, world, port, out_dir, flat=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpatch = _install_mock()
    try:
        m = _build()
        ddp = pdist.wrap_ddp(m, rank, flat=flat)
        feat, label = _frames(rank)
        lid, cam = ddp(feat[:, 0:5], feat[:, 5:8])
        _loss(lid, cam, label).backward()
        g = {n: p.grad.clone() for n, p in m.named_parameters()}
        torch.save(g, os.path.join(out_dir, "grads_%d.pt" % rank))
        assert pdist.max_over_ranks(float(rank + 1), torch.device("cpu")) == float(world)
    finally:
        mpatch.undo()
        dist.destroy_process_group()


def test_shard_helpers():
    assert pdist.shard_seed(1, 0) != pdist.shard_seed(1, 1)
    got = sorted(i for r in range(4) for i in pdist.shard_indices(10, r, 4))
    assert len(got) == 12 and set(got) == set(range(10))  # padded by wrap-around like DistributedSampler
    assert all(len(pdist.shard_indices(10, r, 4)) == 3 for r in range(4))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("flat", [False, True], ids=["torch_ddp", "frame_parallel"])
def test_ddp_two_ranks_gloo(tmp_path, flat):
    """flat=False: stock DistributedDataParallel (what the unchanged trainer builds); flat=True: pmf_b200.dist.FrameParallel
    (parameters / buffers broadcast from rank 0, gradients averaged by the module's own all-reduce)."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), flat), nprocs=2, join=True)
    g0 = torch.load(os.path.join(tmp_path, "grads_0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "grads_1.pt"))
    # single-process reference: mean of the two ranks' gradients
    mpatch = _install_mock()
    try:
        per_rank = []
        for rank in range(2):
            m = _build()
            feat, label = _frames(rank)
            lid, cam = m(feat[:, 0:5], feat[:, 5:8])
            _loss(lid, cam, label).backward()
            per_rank.append({n: p.grad.clone() for n, p in m.named_parameters()})
    finally:
        mpatch.undo()
    for n in g0:
        assert torch.equal(g0[n], g1[n]), n  # all-reduced: identical on both ranks
        ref = 0.5 * (per_rank[0][n] + per_rank[1][n])
        assert torch.allclose(g0[n], ref, rtol=1e-4, atol=1e-6 * float(ref.abs().max()) + 1e-12), n
