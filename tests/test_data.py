"""CPU: the input pipeline (iodine_b200.data) against

* the committed fixture ``tests/golden/data_clevr.npz`` = what the UNMODIFIED reference dataset returned for a
  synthetic CLEVR-shaped sample (``oracle/make_data_golden.py``),
* the live reference datasets where ``/root/reference`` is mounted (``oracle/data_ref.py`` shims),
* the torchvision transforms the reference composes (``lib/data/clevr.py:26-38``), bit for bit,
and the loader / prefetcher / evaluation-loop plumbing with stand-in models.
"""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import data_ref
from oracle.make_data_golden import canonical, synthetic_sample

from iodine_b200.data import CLEVR, DevicePrefetcher, MultiDSprites, collate_fn, make_dataloader, make_dataset
from iodine_b200.data import transforms as T
from iodine_b200.eval import evaluate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'data_clevr.npz')


def _clevr_dir(tmp_path, samples, with_masks=True):
    root = tmp_path / 'CLEVR'
    (root / 'images').mkdir(parents=True)
    (root / 'masks').mkdir()
    for name, (img, mask) in samples.items():
        Image.fromarray(img).save(root / 'images' / name)
        if with_masks and mask is not None:
            Image.fromarray(mask).save(root / 'masks' / name)
    return str(root)


def test_clevr_matches_the_reference_fixture(tmp_path):
    g = np.load(GOLDEN)
    root = tmp_path / 'CLEVR'
    (root / 'images').mkdir(parents=True)
    (root / 'masks').mkdir()
    (root / 'images' / 'a.png').write_bytes(g['image_png'].tobytes())
    (root / 'masks' / 'a.png').write_bytes(g['mask_png'].tobytes())
    ds = CLEVR(str(root), 'test')
    assert len(ds) == 1
    x, m = ds[0]
    assert x.dtype == torch.float32 and tuple(x.shape) == (3, 128, 128)
    assert m.dtype == torch.float32 and tuple(m.shape) == tuple(g['masks'].shape)
    assert np.array_equal(np.round(x.numpy() * 255).astype(np.uint8), g['image'])
    assert x.double().sum().item() == float(g['image_f32_checksum'])          # same float32 values, not just same bytes
    assert np.array_equal(canonical(m.numpy()), g['masks'])                   # same masks up to their order


@pytest.mark.skipif(not data_ref.data_reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('shape,rgba', [((320, 480), False), ((320, 480), True), ((200, 333), False), ((150, 170), False)])
def test_clevr_matches_live_reference(tmp_path, shape, rgba):
    """full-size, RGBA, odd-sized and smaller-than-crop images (torchvision pads those)"""
    img, mask = synthetic_sample(seed=3, H=max(shape[0], 200), W=max(shape[1], 320))
    img, mask = img[:shape[0], :shape[1]], mask[:shape[0], :shape[1]]
    mask[0, 0] = 64                                                           # the background colour must exist
    if rgba:
        img = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], -1)
    root = _clevr_dir(tmp_path, {'b.png': (img, mask), 'a.png': (img[::-1].copy(), None)})
    ref = data_ref.load_reference_dataset_module('clevr').CLEVR(root, 'test')
    ours = CLEVR(root, 'test')
    assert ours.img_paths == ref.img_paths and len(ours) == len(ref) == 2
    with data_ref.legacy_numpy():
        for i in range(2):
            xr, mr = ref[i]
            xo, mo = ours[i]
            assert torch.equal(xo, xr)
            assert (mo is None) == (mr is None)
            if mr is not None:
                assert mo.dtype == mr.dtype and np.array_equal(canonical(mo.numpy()), canonical(mr.numpy()))


@pytest.mark.skipif(not data_ref.data_reference_available(), reason='reference tree not mounted')
def test_dsprites_matches_live_reference(tmp_path):
    root = tmp_path / 'DSPRITES'
    (root / 'images').mkdir(parents=True)
    (root / 'masks').mkdir()
    rng = np.random.RandomState(0)
    for i in range(3):
        Image.fromarray(rng.randint(0, 256, size=(64, 64, 3)).astype(np.uint8)).save(root / 'images' / ('%d.png' % i))
        np.save(root / 'masks' / ('%d.npy' % i), rng.randint(0, 2, size=(2 + i, 64, 64)).astype(np.uint8))
    ref = data_ref.load_reference_dataset_module('dsprite').MultiDSprites(str(root), 'test')
    ours = MultiDSprites(str(root), 'test')
    assert len(ours) == 3                         # the reference hard-codes 60000 (dsprite.py:31)
    with data_ref.legacy_numpy():
        for i in range(3):
            xr, mr = ref[i]
            xo, mo = ours[i]
            assert xo.dtype == xr.dtype and torch.equal(xo, xr)
            assert mo.dtype == mr.dtype and torch.equal(mo, mr)


def test_transforms_are_bit_exact_with_torchvision():
    tv = pytest.importorskip('torchvision.transforms')
    rng = np.random.RandomState(1)
    for H, W in [(320, 480), (192, 192), (200, 191), (100, 260), (128, 128)]:
        a = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        want = tv.Compose([tv.ToPILImage(), tv.CenterCrop(192), tv.Resize(128), tv.ToTensor()])(a)
        got = T.to_tensor(T.resize_shorter(T.center_crop(Image.fromarray(a), 192), 128, Image.BILINEAR))
        assert torch.equal(got, want), (H, W)
        m = (rng.rand(H, W) > 0.5).astype(np.uint8)
        want_m = np.array(tv.Compose([tv.ToPILImage(), tv.CenterCrop(192),
                                      tv.Resize(128, interpolation=tv.InterpolationMode.NEAREST)])(m[:, :, None]))
        got_m = np.asarray(T.resize_shorter(T.center_crop(Image.fromarray(m), 192), 128, Image.NEAREST))
        assert np.array_equal(got_m, want_m), (H, W)


def test_sep_against_per_pixel_enumeration_and_missing_background():
    _, mask = synthetic_sample(seed=5, H=200, W=320)
    got = CLEVR.sep(mask)
    colours = {tuple(p) for p in mask.reshape(-1, 3)} - {(64, 64, 64)}
    want = [np.all(mask == np.array(c, np.uint8), axis=2) for c in colours]
    assert len(got) == len(want) == len(colours)
    assert np.array_equal(canonical(got), canonical(want))
    assert all(m.dtype == np.bool_ for m in got)
    rgba = np.concatenate([mask, np.full(mask.shape[:2] + (1,), 255, np.uint8)], -1)      # alpha is ignored (clevr.py:63)
    assert np.array_equal(canonical(CLEVR.sep(rgba)), canonical(want))
    with pytest.raises(KeyError):                                              # set.remove in the reference (clevr.py:72)
        CLEVR.sep(np.zeros((4, 4, 3), np.uint8))


def test_imread_expands_palette_images(tmp_path):
    a = np.zeros((8, 8, 3), np.uint8)
    a[2:5, 3:6] = (255, 0, 0)
    p = tmp_path / 'p.png'
    Image.fromarray(a).convert('P', palette=Image.ADAPTIVE, colors=4).save(p)
    assert np.array_equal(T.imread(str(p)), a)


def _cfg(name, bs=2, workers=0):
    return NS(DATASET=NS(TRAIN=name), TRAIN=NS(BATCH_SIZE=bs), VAL=NS(BATCH_SIZE=bs), TEST=NS(BATCH_SIZE=bs),
              DATALOADER=NS(NUM_WORKERS=workers))


def test_make_dataloader_contract(tmp_path):
    samples = {}
    for i in range(5):
        img, mask = synthetic_sample(seed=10 + i, H=200, W=320)
        samples['%02d.png' % i] = (img, mask if i != 3 else None)             # one image without a mask file
    root = _clevr_dir(tmp_path, samples)
    dl = make_dataloader(_cfg('CLEVR', bs=2), 'test', root=root)
    batches = list(dl)
    assert [b[0].shape[0] for b in batches] == [2, 2, 1]                      # ragged tail kept, order kept
    ds = CLEVR(root)
    for bi, (data, mask) in enumerate(batches):
        assert isinstance(mask, tuple) and len(mask) == data.shape[0] and data.dtype == torch.float32
        for j in range(data.shape[0]):
            x, m = ds[2 * bi + j]
            assert torch.equal(data[j], x)
            assert (mask[j] is None) == (m is None) and (m is None or torch.equal(mask[j], m))
    assert batches[1][1][1] is None
    assert isinstance(make_dataset(_cfg('DSPRITES'), 'train', root=str(tmp_path)), MultiDSprites)
    with pytest.raises(ValueError):
        make_dataset(_cfg('MNIST'), 'train')
    with pytest.raises(ValueError):
        make_dataloader(_cfg('CLEVR'), 'predict', root=root)
    d, m = collate_fn([(torch.zeros(3, 4, 4), None), (torch.ones(3, 4, 4), torch.ones(2, 4, 4))])
    assert tuple(d.shape) == (2, 3, 4, 4) and m[0] is None and m[1].shape[0] == 2


def test_prefetcher_and_evaluate_loop_on_cpu(tmp_path, capsys):
    samples = {'%d.png' % i: synthetic_sample(seed=20 + i, H=200, W=320) for i in range(3)}
    root = _clevr_dir(tmp_path, samples)
    dl = make_dataloader(_cfg('CLEVR', bs=2), 'val', root=root)
    plain = list(dl)
    staged = list(DevicePrefetcher(dl, 'cpu'))
    assert len(staged) == len(plain) == len(DevicePrefetcher(dl, 'cpu'))
    for a, b in zip(staged, plain):
        assert torch.equal(a[0], b[0]) and len(a[1]) == len(b[1])
    assert list(DevicePrefetcher([], 'cpu')) == []

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

    class Evaluator:
        def __init__(self):
            self.calls, self.resets, self.train_flags = [], 0, []

        def reset(self):
            self.resets += 1
            self.calls = []

        def evaluate(self, model, data):
            self.train_flags.append(model.training)
            self.calls.append((data[0].shape[0], len(data[1])))

        def get_results(self):
            return 'Ari: {}'.format(len(self.calls))

    ev = Evaluator()
    out = evaluate(torch.nn.DataParallel(Model()), 'cpu', dl, ev)              # DataParallel is unwrapped (eval.py:15-16)
    assert ev.resets == 1 and ev.calls == [(2, 2), (1, 1)] and ev.train_flags == [False, False]
    assert out == 'Ari: 2' and 'Final:  Ari: 2' in capsys.readouterr().out
