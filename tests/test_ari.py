"""ARI evaluator tail: oracle vs the reference's known answer / live reference (CPU), and
the device kernel vs the oracle, bit-exact on the contingency tables (GPU)."""
import sys

import numpy as np
import pytest
import torch

from oracle import ari as OA
from oracle import ref_loader as R


def _case(seed, G, K, H, W, soft=True):
    rng = np.random.RandomState(seed)
    label = rng.randint(0, G, size=(H, W))
    gt = np.stack([(label == g) for g in range(G)]).astype(np.float32)
    pred = rng.rand(K, H, W).astype(np.float32)
    if not soft:
        pred = OA.one_hot_argmax(pred)
    return gt, pred


def test_oracle_known_answer():
    assert OA.compute_ari(OA.KNOWN_TABLE) == pytest.approx(OA.KNOWN_ARI, abs=1e-15)
    eye = np.diag([5, 7, 9])
    assert OA.compute_ari(eye) != OA.compute_ari(OA.KNOWN_TABLE)
    gt, _ = _case(0, 3, 3, 8, 8)
    assert OA.compute_mask_ari(gt, gt) == pytest.approx(1.0)          # identical partitions


@pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('seed,G,K', [(1, 3, 4), (2, 5, 7), (3, 1, 2), (4, 6, 6)])
def test_oracle_matches_live_reference(seed, G, K):
    if R.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, R.REFERENCE_ROOT)
    from lib.utils import ari as ref_ari
    gt, pred = _case(seed, G, K, 16, 12)
    onehot = OA.one_hot_argmax(pred)
    ref = ref_ari.compute_mask_ari(torch.from_numpy(gt), torch.from_numpy(onehot))
    assert OA.compute_mask_ari(gt, pred) == pytest.approx(ref, abs=1e-14)
    assert ref_ari.compute_ari(OA.KNOWN_TABLE) == pytest.approx(OA.KNOWN_ARI, abs=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize('B,K,H,W,gs', [(3, 4, 16, 12, [3, 1, 5]), (2, 7, 128, 128, [7, 4]), (1, 16, 9, 7, [16]),
                                        (2, 3, 8, 8, [0, 2])])
def test_device_ari_matches_oracle(B, K, H, W, gs):
    from iodine_b200.eval.ari_eval import device_ari
    preds, gts = [], []
    for b in range(B):
        gt, pred = _case(10 + b, max(gs[b], 1), K, H, W)
        gts.append(torch.from_numpy(gt[:gs[b]]))
        preds.append(pred)
    pred = torch.from_numpy(np.stack(preds)).cuda()
    ari, table = device_ari(pred[:, :, None], gts)
    ari, table = ari.cpu().numpy(), table.cpu().numpy()
    for b in range(B):
        t = OA.mask_table(gts[b].numpy(), OA.one_hot_argmax(preds[b])) if gs[b] else np.zeros((0, K), np.int64)
        assert np.array_equal(table[b, :gs[b]], t), b                   # integer work: bit-exact
        assert not table[b, gs[b]:].any()
        want = OA.compute_ari(t) if gs[b] else OA.compute_ari(np.zeros((1, K), np.int64))
        assert (np.isnan(want) and np.isnan(ari[b])) or ari[b] == pytest.approx(want, abs=1e-13), (b, ari[b], want)


@pytest.mark.gpu
def test_device_ari_ties_nan_and_perfect_case():
    from iodine_b200.eval.ari_eval import device_ari
    H = W = 8
    pred = torch.zeros(1, 3, H, W)
    pred[0, 1, :4] = 1.0                         # rows 0-3 -> slot 1, rows 4-7: three-way tie -> slot 0 (first max)
    pred[0, 2, 7, 7] = float('nan')              # NaN counts as maximal (torch.argmax)
    gt = torch.zeros(2, H, W)
    gt[0, :4] = 1
    gt[1, 4:] = 1
    ari, table = device_ari(pred.cuda(), [gt])
    want = np.array([[0, 32, 0], [31, 0, 1]])
    assert np.array_equal(table[0].cpu().numpy(), want)
    assert ari[0].item() == pytest.approx(OA.compute_ari(want), abs=1e-13)
    pred2 = torch.stack([gt[0], gt[1]])[None]    # identical partitions -> exactly 1.0
    ari2, _ = device_ari(pred2.cuda(), [gt])
    assert ari2[0].item() == 1.0


@pytest.mark.gpu
def test_evaluator_runs_on_the_native_model():
    from helpers import seeded_model
    from iodine_b200.eval import ARIEvaluator
    from oracle import arch as A
    arch = A.arch_by_name('tiny')
    model = seeded_model(arch, 4.0, precision='fp16').to('cuda:0')
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 16, 16, generator=g).cuda()
    masks = [torch.from_numpy(_case(5, 3, 3, 16, 16)[0]), torch.from_numpy(_case(6, 2, 3, 16, 16)[0])]
    ev = ARIEvaluator()
    ev.evaluate(model, (x, masks))
    assert len(ev.aris) == 2 and ev.get_results().startswith('Ari: ')
    pm = model.mask[:, :, 0].cpu().numpy()
    for b in range(2):
        assert ev.aris[b] == pytest.approx(OA.compute_mask_ari(masks[b].numpy(), pm[b]), abs=1e-13)
