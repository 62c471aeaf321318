"""CPU restatement of the loss block the reference trainer applies to the two probability maps.
TEST INFRASTRUCTURE: the oracle for a future fused loss head (8f-1); today the loss runs as unchanged PyTorch in the
caller.  Follows tasks/pmf/trainer.py:188-252, 305-332, pc_processor/loss/focal_softmax.py:28-62 and
pc_processor/loss/lovasz_softmax.py:55-145 (paths relative to /root/reference).  Pinned by
tests/test_oracle_pinning.py::test_loss_oracle_equals_reference_modules against the reference's own classes composed
the way the trainer composes them.
"""
import math

import torch


def focal_alpha(cls_freq):
    """trainer.py:108,195-199 (SemanticKitti): w = 1/(freq + 1e-3); alpha = log(1 + w) / max; alpha[0] = 0."""
    w = 1.0 / (torch.as_tensor(cls_freq, dtype=torch.float64) + 1e-3)
    a = torch.log(1 + w)
    a = a / a.max()
    a[0] = 0
    return a.float()


def focal_loss(pred, label, alpha, gamma=2.0):
    """FocalSoftmaxLoss(softmax=False).forward(pred, label, mask=label > 0), focal_softmax.py:28-62:
    -(1 - p_t)^gamma * log(clamp(p_t, 1e-6)) * alpha[t], averaged over the labelled (label > 0) pixels."""
    c = pred.shape[1]
    p = pred.permute(0, 2, 3, 1).reshape(-1, c)
    t = label.reshape(-1, 1)
    pt = p.gather(1, t).view(-1)
    loss = -(1 - pt).pow(gamma) * pt.clamp(1e-6).log() * alpha.to(pred.device).gather(0, t.squeeze(1))
    mask = (label > 0).float().view(-1)
    return (loss * mask).sum() / mask.sum()


def lovasz_grad(gt_sorted):
    """lovasz_softmax.py:55-66."""
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jaccard = 1.0 - intersection / union
    if p > 1:
        jaccard[1:p] = jaccard[1:p] - jaccard[0:-1]
    return jaccard


def lovasz_softmax(pred, label, ignore=0):
    """Lovasz_softmax(classes='present', per_image=False, ignore=0), lovasz_softmax.py:69-145: pixels with the ignored
    label dropped, one sorted-error / Jaccard-gradient dot product per class PRESENT in the labels, mean over them."""
    c = pred.shape[1]
    p = pred.permute(0, 2, 3, 1).reshape(-1, c)
    t = label.reshape(-1)
    keep = t != ignore
    p, t = p[keep], t[keep]
    if p.numel() == 0:
        return p * 0.0
    losses = []
    for k in range(c):
        fg = (t == k).float()
        if fg.sum() == 0:
            continue
        errors = (fg - p[:, k]).abs()
        errors_sorted, perm = torch.sort(errors, 0, descending=True)
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg[perm])))
    return sum(losses) / len(losses)


def perception_aware_loss(pcd_pred, img_pred, nclasses, tau=0.7):
    """trainer.py:231-252 with the entropies of :308-321: confidence = 1 - H(p)/log C; each modality is pulled towards the
    other where the other is the more confident one AND above tau; KLDivLoss(reduction='none')(log p_a, p_b) =
    p_b * (log p_b - log p_a) (0 where p_b = 0), weighted per pixel, mean over ALL B*C*H*W elements."""
    pcd_log = torch.log(pcd_pred.clamp(min=1e-8))
    img_log = torch.log(img_pred.clamp(min=1e-8))
    pcd_conf = 1 + (pcd_pred * pcd_log).sum(1) / math.log(nclasses)
    img_conf = 1 + (img_pred * img_log).sum(1) / math.log(nclasses)
    imp = pcd_conf - img_conf
    pcd_w = imp.gt(0).float() * imp.abs() * pcd_conf.ge(tau).float()
    img_w = imp.lt(0).float() * imp.abs() * img_conf.ge(tau).float()

    def kl(log_a, b):
        return torch.xlogy(b, b) - b * log_a

    loss_pcd = (kl(pcd_log, img_pred) * img_w.unsqueeze(1)).mean()
    loss_img = (kl(img_log, pcd_pred) * pcd_w.unsqueeze(1)).mean()
    return loss_pcd + loss_img, pcd_w, img_w


def total_loss(lidar_pred, camera_pred, label, alpha, nclasses=20, lambda_=1.0, gamma=0.5, tau=0.7):
    """trainer.py:305-332: focal + lambda * lovasz on both heads + gamma * perception-aware loss
    (config_server_kitti.yaml:29-31: lambda 1.0, gamma 0.5, tau 0.7)."""
    per, _, _ = perception_aware_loss(lidar_pred, camera_pred, nclasses, tau)
    return (focal_loss(lidar_pred, label, alpha) + lambda_ * lovasz_softmax(lidar_pred, label) +
            focal_loss(camera_pred, label, alpha) + lambda_ * lovasz_softmax(camera_pred, label) + gamma * per)
