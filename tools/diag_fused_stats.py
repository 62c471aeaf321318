"""Diagnostic: train-mode forward with / without fused BN statistics, eager vs CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmf_b200
from tests import synth

dev = torch.device("cuda:0")
torch.manual_seed(1)
m = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34")
sd = {k: v.clone() for k, v in m.state_dict().items()}
m.to(dev).train()
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout2d):
        mod.eval()
feat, _, label = synth.frame_tensor(2, 64, 96, seed=9)
x = feat.to(dev)

def run():
    m.load_state_dict(sd)
    with torch.no_grad():
        pass
    lid, cam = m(x[:, 0:5], x[:, 5:8])
    return lid.detach().clone(), {k: v.clone() for k, v in m.state_dict().items() if "running" in k}

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())

os.environ["PMFB_CUDA_GRAPH"] = "0"
os.environ["PMFB_FUSED_BN_STATS"] = "0"
a0, s0 = run(); a1, s1 = run()
os.environ["PMFB_FUSED_BN_STATS"] = "1"
b0, t0 = run(); b1, t1 = run()
print("unfused eager run-to-run", rel(a1, a0))
print("fused eager run-to-run  ", rel(b1, b0))
print("fused vs unfused (eager)", rel(b0, a0))
worst = max(((rel(t0[k], s0[k]), k) for k in s0))
print("running stats fused vs unfused worst", worst)
ws = sorted(((rel(t0[k], s0[k]), k) for k in s0), reverse=True)[:6]
print(ws)

keys = [k for k in s0 if k.endswith("running_var")]
print("per-layer running_var: fused vs unfused | fused run-to-run  (state_dict order)")
for k in keys:
    d1, d2 = rel(t0[k], s0[k]), rel(t1[k], t0[k])
    if d1 > 1e-6 or d2 > 1e-6:
        print("  %-55s %.2e %.2e  C=%d" % (k, d1, d2, s0[k].numel()))
