"""CPU restatement of the inference tail (TEST INFRASTRUCTURE).

  merge_cameras : getMergePred, tasks/pmf_eval_nuscenes/infer.py:18-38 (per point, the prediction of the camera with the
                  highest confidence; torch.argmax over the 6 camera rows takes the FIRST maximum; unseen points -> -1)
  semantic_tail : tasks/pmf_eval_semantickitti/infer.py:85-86,107-129 (crop -> argmax -> KNN or gather -> class_map_lut_inv)
Pinned against the reference function itself in tests/test_oracle_pinning.py when /root/reference is present.
"""
import numpy as np

from oracle import knn_oracle


def merge_cameras(point_idx_list, pred_conf_list, pred_argmax_list, pc_size):
    n_cam = len(point_idx_list)
    conf = np.zeros((n_cam, pc_size), np.float32)
    arg = np.full((n_cam, pc_size), -1, np.int64)
    for j in range(n_cam):
        conf[j, point_idx_list[j]] = pred_conf_list[j]
        arg[j, point_idx_list[j]] = pred_argmax_list[j]
    best = conf.argmax(0)  # first maximum
    return arg[best, np.arange(pc_size)]


def semantic_tail(probs, proj_depth, unproj_range, rows, cols, lut_inv, knn_params, nclasses, crop=None, use_knn=True):
    """One frame.  probs (C, Hp, Wp); proj_depth (H, W) with 0 at empty pixels; rows / cols / unproj_range per point."""
    if crop is not None:
        y0, x0, h, w = crop
        probs = probs[:, y0:y0 + h, x0:x0 + w]
    argmax = probs.argmax(0).astype(np.int64)
    if use_knn:
        rng = (proj_depth - (proj_depth == 0)).astype(np.float32)
        labels = knn_oracle.knn_vote(rng, unproj_range.astype(np.float32), argmax, cols.astype(np.int64), rows.astype(np.int64),
                                     knn_params["knn"], knn_params["search"], knn_params["sigma"], knn_params["cutoff"], nclasses)
    else:
        labels = argmax[rows, cols]
    return np.asarray(lut_inv)[labels].astype(np.int32)
