"""GPU: the whole evaluation flow of the reference's ``tools/test_net.py`` on the native pieces -- dataset ->
``make_dataloader`` -> ``DevicePrefetcher`` -> ``IODINE.reconstruct`` (CUDA loop) -> device ARI -- against the same
samples pushed through the model and the ARI oracle one by one."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import arch as A
from oracle import ari as OA
from oracle.make_data_golden import synthetic_sample

from helpers import seeded_model
from iodine_b200.data import CLEVR, make_dataloader
from iodine_b200.eval import ARIEvaluator, evaluate

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_evaluate_over_a_clevr_directory(tmp_path, capsys):
    root = tmp_path / 'CLEVR'
    (root / 'images').mkdir(parents=True)
    (root / 'masks').mkdir()
    for i in range(3):
        img, mask = synthetic_sample(seed=30 + i, H=200, W=320)
        Image.fromarray(img).save(root / 'images' / ('%d.png' % i))
        Image.fromarray(mask).save(root / 'masks' / ('%d.png' % i))
    cfg = NS(DATASET=NS(TRAIN='CLEVR'), TEST=NS(BATCH_SIZE=2), DATALOADER=NS(NUM_WORKERS=0))
    dl = make_dataloader(cfg, 'test', root=str(root))
    arch = A.arch_by_name('tiny', img_size=128, slots=4, iters=2)
    model = seeded_model(arch, 4.0, precision='fp16')

    class FixedNoise(torch.nn.Module):          # reconstruct(image) with reproducible noise, like the evaluator calls it
        def __init__(self, m):
            super().__init__()
            self.m = m

        def reconstruct(self, x):
            g = torch.Generator().manual_seed(int(x.shape[0]))
            eps = torch.randn(arch.ITERS + 1, x.shape[0], arch.SLOTS, arch.DIM_LATENT, generator=g)
            return self.m.reconstruct(x, eps=eps.to(x.device))

    wrapped = FixedNoise(model)
    ev = ARIEvaluator()
    out = evaluate(wrapped, DEV, dl, ev)
    assert len(ev.aris) == 3 and out == ev.get_results() and 'Final:  Ari:' in capsys.readouterr().out
    ds = CLEVR(str(root))
    want = []
    for b0 in (0, 2):                            # the loader's batches: [0, 1], [2]
        xs = torch.stack([ds[i][0] for i in range(b0, min(b0 + 2, 3))]).to(DEV)
        _, mask, _ = wrapped.reconstruct(xs)
        for j in range(xs.shape[0]):
            want.append(OA.compute_mask_ari(ds[b0 + j][1].numpy(), mask[j, :, 0].cpu().numpy()))
    assert np.allclose(ev.aris, want, atol=1e-12)
