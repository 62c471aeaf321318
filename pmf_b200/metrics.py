"""IOUEval on the device: the reference keeps its confusion matrix on the CPU, is fed CUDA tensors
(tasks/pmf/trainer.py:384-396) and, when distributed, runs a barrier + all_reduce inside EVERY getIoU / getAcc / getRecall
call — six pairs per training iteration (pc_processor/metrics/iou_eval.py:59-74).

Same interface and the same numbers (``addBatch``, ``getStats``, ``getIoU``, ``getAcc``, ``getRecall``, ``reset``):
  * ``addBatch`` is one kernel (pmfb_confusion_add: shared-memory histogram of (prediction, target) pairs) on the int64
    device matrix, no host synchronisation;
  * the distributed reduction runs at most ONCE per batch of queries (the reduced matrix is cached until the next addBatch),
    without barriers.
"""
import torch
import torch.distributed as dist

from . import _lib as L


class IOUEval:
    def __init__(self, n_classes, device=None, ignore=None, is_distributed=False):
        self.n_classes = int(n_classes)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None or torch.device(device).type != "cuda" \
            else torch.device(device)
        self.ignore = torch.tensor(ignore if ignore is not None else []).long()
        self.include = torch.tensor([n for n in range(self.n_classes) if n not in self.ignore]).long()
        self.is_distributed = bool(is_distributed)
        self.reset()

    def num_classes(self):
        return self.n_classes

    def reset(self):
        self.conf_matrix = torch.zeros((self.n_classes, self.n_classes), device=self.device, dtype=torch.int64)
        self._reduced = None

    def addBatch(self, x, y):  # x = predictions, y = targets
        if not torch.is_tensor(x):
            x = torch.as_tensor(x)
        if not torch.is_tensor(y):
            y = torch.as_tensor(y)
        x = x.to(self.device).long().reshape(-1).contiguous()
        y = y.to(self.device).long().reshape(-1).contiguous()
        assert x.numel() == y.numel()
        L.require_device()
        with torch.cuda.device(self.device):
            L.call("pmfb_confusion_add", x.data_ptr(), y.data_ptr(), x.numel(), self.n_classes, self.conf_matrix.data_ptr(),
                   torch.cuda.current_stream(self.device).cuda_stream)
        self._reduced = None

    def getStats(self):
        if self._reduced is None:
            conf = self.conf_matrix.clone().double()
            if self.is_distributed and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(conf)
            ig = self.ignore.to(self.device)
            conf[ig] = 0
            conf[:, ig] = 0
            self._reduced = conf
        conf = self._reduced
        tp = conf.diag()
        fp = conf.sum(dim=1) - tp
        fn = conf.sum(dim=0) - tp
        return tp, fp, fn

    def getIoU(self):
        tp, fp, fn = self.getStats()
        union = tp + fp + fn + 1e-15
        iou = tp / union
        inc = self.include.to(self.device)
        return (tp[inc] / union[inc]).mean(), iou

    def getAcc(self):
        tp, fp, fn = self.getStats()
        acc = tp / (tp + fp + 1e-15)
        return acc[self.include.to(self.device)].mean(), acc

    def getRecall(self):
        tp, fp, fn = self.getStats()
        recall = tp / (tp + fn + 1e-15)
        return recall[self.include.to(self.device)].mean(), recall
