"""The loss block the reference trainer applies to the two probability maps, as ONE module.

    TrainerLoss(nclasses, alpha, lambda_, gamma, tau)(lidar_pred, camera_pred, label) -> total loss (scalar)

Semantics (paths relative to the reference tree):
  * tasks/pmf/trainer.py:305-332   total = focal + lambda*lovasz (LiDAR head) + focal + lambda*lovasz (camera head)
                                   + gamma * perception-aware loss
  * tasks/pmf/trainer.py:231-252   perception-aware loss: confidence = 1 - H(p)/log C per head; each head is pulled
                                   towards the other (KLDivLoss(reduction="none"), mean over ALL B*C*H*W elements)
                                   where the other is the more confident one AND its confidence >= tau
  * pc_processor/loss/focal_softmax.py:28-62   FocalSoftmaxLoss(gamma=2, alpha, softmax=False)(pred, label, mask=label>0)
  * pc_processor/loss/lovasz_softmax.py:89-145 Lovasz_softmax(classes="present", ignore=0)

``impl="torch"``: the block as plain PyTorch autograd ops on the module's outputs — exactly what the UNCHANGED
tasks/pmf/trainer.py runs on our outputs (kept as the comparison arm; ~40 elementwise kernels over 2x B*C*H*W tensors
and 2x19 full-tensor sorts per step).
``impl="fused"`` (default on CUDA tensors): libpmf_b200.so — pmfb_loss_head (one pass over the two maps: log, entropy,
focal gather, both KL directions and guide weights, forward value AND d loss / d probabilities) and pmfb_lovasz
(compaction of the labelled pixels, one segmented sort over (class, error) keys, Jaccard-gradient scan, gradient
scatter).  Opt-in add-on: it needs a trainer that calls this module instead of the inline block.
"""
import math

import torch
import torch.nn as nn


def focal_alpha_from_freq(cls_freq):
    """trainer.py:108,195-199 (SemanticKitti): w = 1/(freq + 1e-3); alpha = log(1 + w) / max; alpha[0] = 0."""
    w = 1.0 / (torch.as_tensor(cls_freq, dtype=torch.float64) + 1e-3)
    a = torch.log(1 + w)
    a = a / a.max()
    a[0] = 0
    return a.float()


def _lovasz_grad(gt_sorted):
    p = gt_sorted.shape[0]
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.cumsum(0)
    union = gts + (1 - gt_sorted).cumsum(0)
    jac = 1.0 - inter / union
    if p > 1:
        jac[1:p] = jac[1:p] - jac[0:-1]
    return jac


def _lovasz_torch(pred, label, ignore):
    c = pred.shape[1]
    p = pred.permute(0, 2, 3, 1).reshape(-1, c)
    t = label.reshape(-1)
    keep = t != ignore
    p, t = p[keep], t[keep]
    if p.numel() == 0:
        return p.sum() * 0.0
    losses = []
    for k in range(c):
        fg = (t == k).float()
        if fg.sum() == 0:
            continue
        err = (fg - p[:, k]).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)
        losses.append(torch.dot(err_sorted, _lovasz_grad(fg[perm])))
    return sum(losses) / len(losses)


def _focal_torch(pred, label, alpha, gamma):
    c = pred.shape[1]
    p = pred.permute(0, 2, 3, 1).reshape(-1, c)
    t = label.reshape(-1, 1)
    pt = p.gather(1, t).view(-1)
    loss = -(1 - pt).pow(gamma) * pt.clamp(1e-6).log() * alpha.gather(0, t.squeeze(1))
    mask = (label > 0).float().view(-1)
    return (loss * mask).sum() / mask.sum()


def _perception_torch(pcd_pred, img_pred, nclasses, tau):
    pcd_log = torch.log(pcd_pred.clamp(min=1e-8))
    img_log = torch.log(img_pred.clamp(min=1e-8))
    pcd_conf = 1 + (pcd_pred * pcd_log).sum(1) / math.log(nclasses)
    img_conf = 1 + (img_pred * img_log).sum(1) / math.log(nclasses)
    imp = pcd_conf - img_conf
    pcd_w = imp.gt(0).float() * imp.abs() * pcd_conf.ge(tau).float()
    img_w = imp.lt(0).float() * imp.abs() * img_conf.ge(tau).float()
    kl = nn.functional.kl_div
    loss_pcd = (kl(pcd_log, img_pred, reduction="none") * img_w.unsqueeze(1)).mean()
    loss_img = (kl(img_log, pcd_pred, reduction="none") * pcd_w.unsqueeze(1)).mean()
    return loss_pcd + loss_img


class TrainerLoss(nn.Module):
    """The trainer's loss block (trainer.py:305-332) on the two (B, C, H, W) probability maps and (B, H, W) int64 labels.
    Defaults are the shipped SemanticKITTI configuration (config_server_kitti.yaml:29-31: lambda 1.0, gamma 0.5, tau 0.7;
    focal gamma 2, trainer.py:202-204).  ``alpha``: per-class focal weights (alpha[0] = 0); None = the nuScenes branch
    (trainer.py:200-201: ones with alpha[0] = 0)."""

    def __init__(self, nclasses=20, alpha=None, lambda_=1.0, gamma=0.5, tau=0.7, focal_gamma=2.0, ignore=0, impl="auto"):
        super().__init__()
        if alpha is None:
            alpha = torch.ones(nclasses)
            alpha[0] = 0
        self.register_buffer("alpha", torch.as_tensor(alpha, dtype=torch.float32).clone())
        self.nclasses, self.lambda_, self.gamma, self.tau = int(nclasses), float(lambda_), float(gamma), float(tau)
        self.focal_gamma, self.ignore = float(focal_gamma), int(ignore)
        if impl not in ("auto", "torch", "fused"):
            raise ValueError("impl must be 'auto', 'torch' or 'fused'")
        self.impl = impl
        self.last = {}

    def forward(self, lidar_pred, camera_pred, label):
        impl = self.impl
        if impl == "auto":
            impl = "fused" if (lidar_pred.is_cuda and _fused_available()) else "torch"
        if impl == "fused":
            from . import loss_fused
            return loss_fused.trainer_loss(self, lidar_pred, camera_pred, label)
        alpha = self.alpha.to(lidar_pred.device)
        foc = _focal_torch(lidar_pred, label, alpha, self.focal_gamma)
        lov = _lovasz_torch(lidar_pred, label, self.ignore)
        foc_c = _focal_torch(camera_pred, label, alpha, self.focal_gamma)
        lov_c = _lovasz_torch(camera_pred, label, self.ignore)
        per = _perception_torch(lidar_pred, camera_pred, self.nclasses, self.tau)
        self.last = dict(focal=foc.detach(), lovasz=lov.detach(), focal_cam=foc_c.detach(), lovasz_cam=lov_c.detach(),
                         perception=per.detach())
        return foc + lov * self.lambda_ + foc_c + lov_c * self.lambda_ + per * self.gamma


def _fused_available():
    try:
        from . import loss_fused  # noqa: F401
        return loss_fused.available()
    except ImportError:
        return False
