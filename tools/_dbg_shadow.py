import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import pmf_b200
from pmf_b200 import engine as eng, _lib as L
cnt = collections.Counter()
orig = L.call
def call(name, *a):
    if name == "pmfb_convert16":
        import traceback
        fr = [f.name for f in traceback.extract_stack()[-6:-1]]
        cnt[(tuple(a[1:5]), a[-2], "/".join(fr[-3:]))] += 1
    return orig(name, *a)
L.call = call
eng.L.call = call
os.environ["PMFB_CUDA_GRAPH"] = "0"
torch.manual_seed(0)
m = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34").cuda().train()
x = torch.randn(2, 8, 64, 96, device="cuda")
lid, cam = m(x[:, :5], x[:, 5:])
(lid.sum() + cam.sum()).backward()
for k, v in sorted(cnt.items(), key=lambda kv: -kv[0][0][0] * kv[0][0][1] * kv[0][0][2] * kv[0][0][3]):
    if not (k[0][0] == 1 and k[0][1] == 1): print(v, k)
