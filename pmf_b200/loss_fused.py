"""Fused implementation of pmf_b200.loss.TrainerLoss on libpmf_b200.so (pmfb_loss_head + pmfb_lovasz, include/pmfb.h).

The forward computes the loss value AND its gradient with respect to the two probability maps in the same kernels (the
gradient of every term is local to a pixel once three scalars are known: the labelled-pixel count, the number of present
classes and a sorted rank), so backward only scales the stored gradient maps by the incoming scalar.
"""
import torch

from . import _lib as L

_ws_cache = {}


def available():
    try:
        return hasattr(L.lib(), "pmfb_loss_head")
    except L.PmfbError:
        return False


def _workspace(dev, n_pix, c, heads):
    need = int(L.lib().pmfb_lovasz_workspace_bytes(n_pix, c, heads))
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws, need


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lidar, camera, label, mod):
        if not (lidar.is_cuda and camera.is_cuda and label.is_cuda):
            raise RuntimeError("pmf_b200 fused TrainerLoss runs on a B200 only (use impl='torch' elsewhere)")
        L.require_device()
        n, c, h, w = lidar.shape
        assert camera.shape == lidar.shape and tuple(label.shape) == (n, h, w), (lidar.shape, camera.shape, label.shape)
        dev = lidar.device
        pl = lidar.detach().float().contiguous()
        pc = camera.detach().float().contiguous()
        lab = label.detach().long().contiguous()
        alpha = mod.alpha.to(device=dev, dtype=torch.float32).contiguous()
        need_grad = any(ctx.needs_input_grad[:2])
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            sums = torch.zeros(10, dtype=torch.float64, device=dev)  # [0:8] pmfb_loss_head sums, [8:10] Lovasz per head
            dl = torch.empty_like(pl) if need_grad else None
            dc = torch.empty_like(pc) if need_grad else None
            L.call("pmfb_loss_head", pl.data_ptr(), pc.data_ptr(), lab.data_ptr(), n, c, h, w, alpha.data_ptr(), mod.focal_gamma,
                   mod.tau, 1.0, mod.gamma, None if dl is None else dl.data_ptr(), None if dc is None else dc.data_ptr(),
                   sums.data_ptr(), st)
            ws, need = _workspace(dev, n * h * w, c, 2)
            L.call("pmfb_lovasz", pl.data_ptr(), pc.data_ptr(), lab.data_ptr(), n, c, h, w, mod.ignore, mod.lambda_,
                   None if dl is None else dl.data_ptr(), None if dc is None else dc.data_ptr(), sums[8:].data_ptr(), ws.data_ptr(),
                   need, st)
        # the handful of scalar operations that combine the device-side sums (plumbing on 10 doubles)
        n_lab = sums[2].clamp_min(1.0)
        foc, foc_c = sums[0] / n_lab, sums[1] / n_lab
        per = (sums[3] + sums[4]) / float(n * c * h * w)
        total = foc + foc_c + mod.lambda_ * (sums[8] + sums[9]) + mod.gamma * per
        mod.last = dict(focal=foc.float(), lovasz=sums[8].float(), focal_cam=foc_c.float(), lovasz_cam=sums[9].float(),
                        perception=per.float(), entropy=(sums[5] / float(n * h * w)).float(),
                        entropy_cam=(sums[6] / float(n * h * w)).float())
        if need_grad:
            ctx.save_for_backward(dl, dc)
        return total.float()

    @staticmethod
    def backward(ctx, g):
        dl, dc = ctx.saved_tensors
        return (dl * g if ctx.needs_input_grad[0] else None, dc * g if ctx.needs_input_grad[1] else None, None, None)


def trainer_loss(mod, lidar_pred, camera_pred, label):
    return _Fn.apply(lidar_pred, camera_pred, label, mod)
