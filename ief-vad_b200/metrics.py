"""Frame-level AUC / AP on the GPU - what the reference computes with scikit-learn on 16x-repeated segment
scores (train/ucf_test.py:151-152, :164-178, :336-353)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


def segment_positives(gt, repeat: int = 16, device=None) -> torch.Tensor:
    """gt: 0/1 per raw frame (numpy or tensor), `repeat` frames per embedding row -> int32 positives per row."""
    g = torch.as_tensor(gt)
    if device is not None:
        g = g.to(device)
    if g.numel() % repeat:
        raise ValueError(f"gt length {g.numel()} is not a multiple of repeat={repeat}")
    return g.reshape(-1, repeat).to(torch.int32).sum(dim=1, dtype=torch.int32)


def frame_auc_ap(scores: torch.Tensor, pos: torch.Tensor, repeat: int = 16) -> Tuple[float, float]:
    """(roc_auc_score, average_precision_score) of np.repeat(scores, repeat) vs the frame labels summarised by
    `pos` (see segment_positives).  One device->host read of 4 doubles."""
    out = ops.auc_ap(scores, pos, repeat).cpu()
    return float(out[0]), float(out[1])


# ------------------------------------------------------------------------------------------------ localisation mAP (row N5)
UCF_CLASSLIST = ['Normal', 'Abuse', 'Arrest', 'Arson', 'Assault', 'Burglary', 'Explosion', 'Fighting', 'RoadAccidents',
                 'Robbery', 'Shooting', 'Shoplifting', 'Stealing', 'Vandalism']          # train/metrics.py:53


def getLocMAP(predictions, th, gtsegments, gtlabels, excludeNormal, _cache=None):
    """train/metrics.py:44-126 on the GPU.  predictions: list of per-video [T_v, 14] arrays / tensors; gtsegments[i] =
    [[start, end], ...] and gtlabels[i] = [class name, ...] of video i.  Returns 100 * mean AP over the 14 classes, or 0
    as soon as one class has no proposal at all (:92-93), like the reference."""
    import ctypes as C

    import numpy as np

    from . import _lib
    if excludeNormal is True:                                                # :45-48
        predictions = predictions[:140]
    prop = _cache if _cache is not None else _locmap_proposals(predictions)
    V, ncls, dev = prop["V"], prop["C"], prop["device"]
    rows = [[] for _ in range(ncls)]
    for i in range(len(gtsegments)):                                          # :99-100, grouped by class
        for j in range(len(gtsegments[i])):
            name = gtlabels[i][j]
            if name in UCF_CLASSLIST[:ncls]:
                rows[UCF_CLASSLIST.index(name)].append((i, int(gtsegments[i][j][0]), int(gtsegments[i][j][1])))
    gt_off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    flat = np.array([x for r in rows for x in r], dtype=np.int32).reshape(-1, 3)
    gt_d = torch.as_tensor(flat if flat.size else np.zeros((1, 3), np.int32), device=dev)
    off_d = torch.as_tensor(gt_off, device=dev)
    ap = torch.empty(ncls, dtype=torch.float64, device=dev)
    n_pred = torch.empty(ncls, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.iefvad_locmap_match(prop["count"].data_ptr(), prop["se"].data_ptr(), prop["score"].data_ptr(), V, ncls,
                                                gt_d.data_ptr(), off_d.data_ptr(), int(flat.shape[0]), C.c_double(float(th)),
                                                ap.data_ptr(), n_pred.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    ap_h, n_h = ap.cpu().numpy(), n_pred.cpu().numpy()
    for c in range(ncls):                                                     # :92-93 `return 0` inside the class loop
        if n_h[c] == 0:
            return 0
    return 100 * np.mean(ap_h)


def _locmap_proposals(predictions):
    """Threshold-independent half of getLocMAP: per (video, class) proposals after NMS, on the device."""
    import numpy as np

    from . import _lib
    preds = [torch.as_tensor(p) for p in predictions]
    dev = next((p.device for p in preds if p.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    ncls = int(preds[0].shape[1]) if preds else 14
    lens = np.array([int(p.shape[0]) for p in preds], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(lens)])[:-1].astype(np.int64)
    allp = (torch.cat([p.to(dev, torch.float32).reshape(-1, ncls) for p in preds]).contiguous() if preds
            else torch.zeros((1, ncls), dtype=torch.float32, device=dev))
    V = len(preds)
    count = torch.zeros((max(V, 1), ncls), dtype=torch.int32, device=dev)
    se = torch.zeros((max(V, 1), ncls, 512, 2), dtype=torch.int32, device=dev)
    score = torch.zeros((max(V, 1), ncls, 512), dtype=torch.float32, device=dev)
    cscore = torch.zeros((max(V, 1), ncls), dtype=torch.float32, device=dev)
    off_d, len_d = torch.as_tensor(off, device=dev), torch.as_tensor(lens, device=dev)     # (kept alive across the launch)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.iefvad_locmap_proposals(allp.data_ptr(), off_d.data_ptr(), len_d.data_ptr(), V, ncls,
                                                    int(lens.max()) if V else 0, count.data_ptr(), se.data_ptr(), score.data_ptr(),
                                                    cscore.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    if V and int(count.min()) < 0:
        raise RuntimeError("getLocMAP: a (video, class) column produced more than 512 proposals")
    return {"V": V, "C": ncls, "device": dev, "count": count, "se": se, "score": score, "class_score": cscore}


def getDetectionMAP(predictions, segments, labels, excludeNormal=False):
    """train/metrics.py:129-136: localisation mAP at IoU 0.1 ... 0.5 -> (dmap_list, iou_list).  The proposals do not depend
    on the IoU threshold, so they are built once and matched five times."""
    iou_list = [0.1, 0.2, 0.3, 0.4, 0.5]
    preds = predictions[:140] if excludeNormal is True else predictions
    cache = _locmap_proposals(preds)
    return [getLocMAP(preds, iou, segments, labels, False, _cache=cache) for iou in iou_list], iou_list
