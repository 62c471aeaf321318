"""Worker of tests/test_frame_parallel_gpu.py (launched by torch.distributed.run, one rank per GPU): the same three
training steps (eager, capture, replay) under pmf_b200.dist.FrameParallel and under stock DistributedDataParallel must give
the same averaged gradients, the same BatchNorm buffers and the same parameters after the optimiser steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run(kind, rank, local):
    import pmf_b200
    from pmf_b200 import dist as pdist
    from pmf_b200.loss import TrainerLoss
    from tests import synth
    dev = torch.device("cuda", local)
    torch.manual_seed(5 + rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
    m = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34").to(dev)
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.eval()
    net = pdist.FrameParallel(m) if kind == "flat" else torch.nn.parallel.DistributedDataParallel(m, device_ids=[local])
    # lr = 0: the three iterations (eager pass, graph capture, graph replay) see the same weights, so that each one's
    # gradients can be compared with torch DDP's (batch-statistics BN on random weights amplifies any parameter difference)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    crit = TrainerLoss(20, impl="torch").to(dev)
    feat, _, label = synth.frame_tensor(2, 64, 96, seed=40 + rank, density=0.3)
    x, y = feat.to(dev), label.to(dev)
    grads = []
    for it in range(3):
        lid, cam = net(x[:, 0:5], x[:, 5:8])
        loss = crit(lid, cam, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.cuda.synchronize()
        grads.append({n: p.grad.detach().clone() for n, p in m.named_parameters()})
        opt.step()
    return m, grads


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ma, ga = run("flat", rank, local)
    mb, gb = run("torch", rank, local)
    worst = 0.0
    for it in range(3):
        for n in ga[it]:
            # conv biases in front of a BatchNorm have analytically zero gradients: compare on the layer's weight-gradient scale
            wn = n.rsplit(".", 1)[0] + ".weight"
            den = max(float(gb[it][n].abs().max()), 1e-2 * float(gb[it][wn].abs().max())) + 1e-20
            worst = max(worst, float((ga[it][n] - gb[it][n]).abs().max()) / den)
    # wgrad flushes with fp32 atomics and (f16 mode) bf16 operands: a few 1e-3 of a tensor's scale is the kernels' own noise
    assert worst < 2e-2, worst
    for (n, a), (_, b) in zip(ma.state_dict().items(), mb.state_dict().items()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-3, atol=1e-4), n
    # every rank holds the same parameters (rank 0's initialisation + the same averaged updates)
    flat = torch.cat([p.detach().reshape(-1) for p in ma.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref)
    if rank == 0:
        print("frame parallel ok: worst relative gradient difference vs torch DDP %.2e" % worst)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
