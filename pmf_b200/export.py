"""Frozen-inference export: the eval-mode network with its BatchNorm layers folded and its weights in
the kernels' packed operand layout, as ONE file.

    save_folded(model, path)        model: pmf_b200.PMFNet in eval mode on the GPU
    load_folded(path, device)    -> a frozen pmf_b200.PMFNet whose first forward neither packs nor folds anything

File contents (torch.save of a dict): the constructor arguments, the reference-compatible ``state_dict`` (654 keys: a
released checkpoint loads into it and the export stays loadable by the reference), and the folded inference parameters —
per convolution the packed tf32-rounded weights ``[taps][c_out_p][c_in_p]`` + bias, per BatchNorm the folded
``(alpha, beta) = (gamma * invstd, beta - mean * gamma * invstd)`` and, where a conv bias precedes the BatchNorm,
``alpha * bias + beta``.  Folding keeps the reference's operator order (pmf_net.py:13-29, salsanext.py:27-33): conv -> BN
-> ReLU folds in FRONT of the activation, conv -> LeakyReLU -> BN BEHIND it — it cannot be folded into the same conv's
weights, and folding it into the next conv would break at the zero-padded borders.
"""
import torch

from . import _lib as L
from .engine import WeightCache
from .modules import PMFNet

FORMAT = "pmfb-folded-v1"


def fold(model):
    """Runs one eval forward of a frozen copy of the caches and returns the folded parameter dict (device tensors)."""
    if not isinstance(model, PMFNet):
        raise TypeError("fold() takes a pmf_b200.PMFNet")
    if model.training:
        raise RuntimeError("fold() needs model.eval(): it folds the RUNNING BatchNorm statistics")
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("fold() runs on a B200 (the packing and folding kernels are device code)")
    was = bool(getattr(model, "_frozen", False))
    model.freeze(True)
    with torch.no_grad(), L_precision("tf32"):
        pcd = torch.zeros(1, model.lidar_stream.downCntx.conv1.in_channels, 16, 16, device=dev)
        img = torch.zeros(1, model.camera_stream_encoder.conv1.in_channels, 16, 16, device=dev)
        model(pcd, img)
    cache = model._cache
    out = {"conv": {k: {"fwd": e["fwd"].clone(), "bias": None if e["bias"] is None else e["bias"].clone()}
                    for k, e in cache.entries.items()},
           "bn": {k: (v[0].clone(), v[1].clone()) if isinstance(v, tuple) else v.clone() for k, v in cache.folded.items()}}
    model.freeze(was)
    return out


class L_precision:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = L.set_precision(self.mode)

    def __exit__(self, *a):
        L.set_precision(self.prev)


def save_folded(model, path):
    folded = fold(model)
    cpu = lambda t: None if t is None else t.detach().cpu()  # noqa: E731
    blob = {"format": FORMAT,
            "ctor": dict(pcd_channels=model.lidar_stream.downCntx.conv1.in_channels,
                         img_channels=model.camera_stream_encoder.conv1.in_channels, nclasses=model.nclasses,
                         base_channels=model.lidar_stream.base_channels, imagenet_pretrained=False,
                         image_backbone=model.image_backbone),
            "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "conv": {k: {"fwd": cpu(e["fwd"]), "bias": cpu(e["bias"])} for k, e in folded["conv"].items()},
            "bn": {k: (cpu(v[0]), cpu(v[1])) if isinstance(v, tuple) else cpu(v) for k, v in folded["bn"].items()}}
    torch.save(blob, path)
    return blob


def load_folded(path, device="cuda"):
    blob = torch.load(path, map_location="cpu")
    if blob.get("format") != FORMAT:
        raise ValueError("not a %s file: %r" % (FORMAT, path))
    model = PMFNet(**blob["ctor"])
    model.load_state_dict(blob["state_dict"], strict=True)
    model.to(device).eval()
    model.freeze(True)
    dev = torch.device(device)
    cache = model._cache
    for k, e in blob["conv"].items():
        cache.entries[k] = {"fwd": e["fwd"].to(dev), "dgrad": None, "bias": None if e["bias"] is None else e["bias"].to(dev),
                            "tag": "folded"}
    for k, v in blob["bn"].items():
        cache.folded[k] = (v[0].to(dev), v[1].to(dev)) if isinstance(v, tuple) else v.to(dev)
    return model
