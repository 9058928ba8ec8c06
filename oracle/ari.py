"""CPU restatement of the reference's ARI evaluator tail.  TEST INFRASTRUCTURE.

Follows ``/root/reference/lib/utils/ari.py`` (``compute_ari`` 6-33, ``compute_mask_ari`` 36-54) and the
argmax / one-hot step of ``/root/reference/lib/eval/ari_eval.py:32-35`` in plain numpy.

Pinned: by the only known answer the reference holds for this code -- the table in ``ari.py:56-63``
(``[[3,0,1],[1,2,1],[0,2,2]]`` -> 0.08333333333333333) -- and, when ``/root/reference`` is
mounted, against ``lib.utils.ari`` itself on random masks (``tests/test_ari.py``).
"""
import numpy as np

KNOWN_TABLE = np.array([[3, 0, 1], [1, 2, 1], [0, 2, 2]])
KNOWN_ARI = 0.08333333333333333


def _comb2(x):
    """scipy.special.comb(x, 2) for non-negative integer-valued x (exact in f64 at these magnitudes)."""
    x = np.asarray(x, dtype=np.float64)
    return x * (x - 1.0) * 0.5


def compute_ari(table):
    """ari.py:6-33."""
    table = np.asarray(table)
    a = table.sum(axis=1)
    b = table.sum(axis=0)
    n = a.sum()
    comb_a = _comb2(a).sum()
    comb_b = _comb2(b).sum()
    comb_n = _comb2(n)
    comb_table = _comb2(table).sum()
    if comb_b == comb_a == comb_n == comb_table:          # "the perfect case" (23-25)
        return 1.0
    with np.errstate(divide='ignore', invalid='ignore'):
        return float((comb_table - comb_a * comb_b / comb_n) /
                     (0.5 * (comb_a + comb_b) - (comb_a * comb_b) / comb_n))


def one_hot_argmax(pred_mask):
    """ari_eval.py:32-35: pred_mask [K,H,W] float -> [K,H,W] 0/1 (first maximal index, as torch.argmax)."""
    idx = np.argmax(pred_mask, axis=0)
    out = np.zeros_like(pred_mask)
    np.put_along_axis(out, idx[None], 1.0, axis=0)
    return out


def mask_table(gt_masks, pred_onehot):
    """ari.py:44-52: table[g][k] = sum over pixels of (byte(gt[g]) & byte(pred[k]))."""
    m0 = np.asarray(gt_masks).astype(np.uint8)[:, None]
    m1 = np.asarray(pred_onehot).astype(np.uint8)[None, :]
    return (m0 & m1).astype(np.int64).sum(axis=-1).sum(axis=-1)


def compute_mask_ari(gt_masks, pred_mask):
    """ari_eval.py:32-39 + ari.py:36-54 for one image: gt_masks [N,H,W], pred_mask [K,H,W] (soft)."""
    return compute_ari(mask_table(gt_masks, one_hot_argmax(pred_mask)))
