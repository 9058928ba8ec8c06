"""ARIEvaluator -- host-side mirror of the reference's ``lib/eval/ari_eval.py:6-46`` (and ``make_evaluator``,
``lib/eval/build.py:6-7``): same constructor, ``evaluate(model, data)``, ``reset()``, ``get_results()``.

The reference binarises the predicted masks by argmax, moves ``[K,H,W]`` and ``[N,H,W]`` masks of every image to
the host and builds the contingency table with numpy/scipy (``lib/utils/ari.py``).  Here argmax, table and ARI run
on the device (``iodine_ari`` in ``include/iodine_b200.h``, ``csrc/ari.cu``); only B doubles come back.
"""
import ctypes as C

import numpy as np
import torch

from .. import _cabi


def device_ari(pred_mask, gt_masks):
    """pred_mask: [B,K,1,H,W] or [B,K,H,W] float CUDA tensor (``IODINE.reconstruct``'s mask);
    gt_masks: list of B tensors [N_b,H,W] (any dtype; converted with ``.byte()`` like ``lib/utils/ari.py:44``).
    Returns (ari[B] float64 CUDA tensor, table[B,G,K] int64 CUDA tensor)."""
    lib = _cabi.load()
    if pred_mask.dim() == 5:
        pred_mask = pred_mask[:, :, 0]
    if pred_mask.device.type != 'cuda':
        raise _cabi.IodineError('device_ari runs on CUDA tensors only (no CPU fallback)')
    pred_mask = pred_mask.to(torch.float32).contiguous()
    B, K, H, W = pred_mask.shape
    assert len(gt_masks) == B, 'one ground-truth mask stack per image'
    G = max(1, max(int(m.shape[0]) for m in gt_masks))
    gt = torch.zeros(B, G, H, W, dtype=torch.uint8, device=pred_mask.device)
    n_gt = torch.zeros(B, dtype=torch.int32)
    for b, m in enumerate(gt_masks):
        n_gt[b] = m.shape[0]
        if m.shape[0]:
            gt[b, :m.shape[0]] = m.byte().to(pred_mask.device)
    n_gt = n_gt.to(pred_mask.device)
    table = torch.empty(B, G, K, dtype=torch.int64, device=pred_mask.device)
    ari = torch.empty(B, dtype=torch.float64, device=pred_mask.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(pred_mask.device):
        _cabi.check(lib.iodine_ari(p(pred_mask), p(gt), p(n_gt), B, K, G, H, W, p(table), p(ari),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return ari, table


class ARIEvaluator:
    def __init__(self):
        self.aris = []

    def evaluate(self, model, data):
        """data = (image [B,3,H,W], mask: list of [N,H,W]) -- reference ari_eval.py:13-39."""
        image, mask = data
        pred, pred_mask, mean = model.reconstruct(image)
        ari, _ = device_ari(pred_mask, list(mask))
        self.aris.extend(ari.cpu().tolist())

    def reset(self):
        self.aris = []

    def get_results(self):
        return 'Ari: {}'.format(np.mean(self.aris) if self.aris else 0)


def make_evaluator(cfg):
    return ARIEvaluator()
