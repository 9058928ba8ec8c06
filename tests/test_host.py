"""CPU: the host mirror keeps the reference's Python boundary: state_dict keys / shapes, seeded
default init, checkpoint loading incl. the DataParallel ``module.`` prefix, ``make_model(cfg)``."""
import io
from types import SimpleNamespace

import pytest
import torch

from oracle import arch as A
from oracle import ref_loader as R

from helpers import seeded_model
from iodine_b200.modeling import IODINE, make_model


def _cfg(arch, device='cpu', parallel=False, precision=None):
    m = SimpleNamespace(NAME='IODINE', DEVICE=device, PARALLEL=parallel)
    if precision:
        m.PRECISION = precision
    return SimpleNamespace(MODEL=m, ARCH=arch)


def test_state_dict_keys_match_survey_listing():
    m = IODINE(A.arch_by_name('clevr6'))
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 1109956          # SURVEY.md 8b [measured on the reference]
    assert tuple(sd['refine.mlc.layers.0.weight'].shape) == (64, 17, 3, 3)
    assert tuple(sd['decoder.mlc.layers.0.weight'].shape) == (64, 66, 3, 3)
    assert tuple(sd['decoder.conv.weight'].shape) == (4, 64, 3, 3)
    assert tuple(sd['refine.lstm.weight_ih'].shape) == (1024, 512)
    assert tuple(sd['posterior.init_mean'].shape) == (64,)


@pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('name', ['tiny', 'dsprites', 'test5x5'])
def test_same_keys_shapes_and_seeded_init_as_live_reference(name):
    arch = A.arch_by_name(name)
    ref = R.build_reference_model(arch, seed=0, sharpen=1.0)
    ours = seeded_model(arch)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k          # same creation order -> bit-identical default init


@pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')
def test_reference_checkpoint_loads_with_and_without_module_prefix():
    arch = A.arch_by_name('tiny')
    ref = R.build_reference_model(arch, seed=3, sharpen=2.0)
    buf = io.BytesIO()
    torch.save({'model': ref.state_dict(), 'epoch': 7}, buf)                 # lib/utils/checkpoint.py:43-54
    buf.seek(0)
    ck = torch.load(buf)
    plain = make_model(_cfg(arch))
    plain.load_state_dict(ck['model'])                                       # checkpoint.py:68
    assert torch.equal(plain.decoder.conv.weight, ref.decoder.conv.weight)
    dp = make_model(_cfg(arch, parallel=True))                               # MODEL.PARALLEL: keys carry 'module.'
    dp.load_state_dict({'module.' + k: v for k, v in ck['model'].items()})
    assert torch.equal(dp.module.refine.lstm.weight_hh, ref.refine.lstm.weight_hh)
    assert dp.module.sigma == arch.SIGMA                                     # lib/engine/train.py:97


def test_make_model_contract():
    arch = A.arch_by_name('tiny')
    m = make_model(_cfg(arch, precision='fp16'))
    assert isinstance(m, IODINE) and m.precision == 'fp16'
    with pytest.raises(ValueError):
        make_model(SimpleNamespace(MODEL=SimpleNamespace(NAME='VAE', DEVICE='cpu', PARALLEL=False), ARCH=arch))
    # partial ARCH.ENCODING lists (reference iodine.py:345-374): fewer input channels of the first refinement conv,
    # scattered into the engine's full 17-channel layout with zero weights for the rest
    sub = A.arch_by_name('tiny')
    sub.ENCODING = [e for e in sub.ENCODING if e not in ('coordinate', 'mask_posterior')]
    ms = IODINE(sub)
    assert ms.get_input_size() == (14, 4 * sub.DIM_LATENT)
    assert tuple(ms.refine.mlc.layers[0].weight.shape[:2]) == (sub.REF.CONV_CHAN, 14)
    full = ms._engine_state_dict()['refine.mlc.layers.0.weight']
    assert tuple(full.shape[:2]) == (sub.REF.CONV_CHAN, 17)
    assert full[:, [8, 15, 16]].abs().max() == 0 and torch.equal(full[:, :8], ms.refine.mlc.layers[0].weight[:, :8])
    assert torch.equal(full[:, 9:15], ms.refine.mlc.layers[0].weight[:, 8:14])
    bad = A.arch_by_name('tiny')
    bad.ENCODING = [e for e in bad.ENCODING if e != 'grad_post']
    with pytest.raises(ValueError):
        IODINE(bad)
    bad.ENCODING = ['posterior', 'grad_post', 'not_an_encoding']
    with pytest.raises(ValueError):
        IODINE(bad)


def test_bench_configs_follow_baseline_json():
    """bench.py --config N: the five BASELINE.json configurations, with the per-unit FLOP figures of SURVEY.md 8(d)"""
    import argparse
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    base = json.load(open(os.path.join(root, 'BASELINE.json')))
    assert sorted(bench.CONFIGS) == list(range(1, len(base['configs']) + 1))
    want = {1: (64, 6, 3, 4), 2: (128, 7, 5, 32), 3: (128, 7, 5, 256), 4: (128, 11, 7, 32), 5: (256, 16, 8, 64)}
    for n, (S_, K, T, B) in want.items():
        args = argparse.Namespace(config=n, slots=0, iters=0, img_size=0, batch=0, precision='')
        cfg, arch, b, prec = bench.resolve_config(args)
        assert (arch.IMG_SIZE, arch.SLOTS, arch.ITERS, b) == (S_, K, T, B), n
        assert prec in ('tf32', 'fp16')
        w = bench.workload_string(cfg, S_, K, T, b, prec)
        assert 'configs[%d]' % (n - 1) in w
    # the default line is labelled configs[1] only for fp32-class arithmetic
    cfg = bench.CONFIGS[2]
    assert 'configs[1])' in bench.workload_string(cfg, 128, 7, 5, 32, 'tf32')
    assert 'fp16 operands' in bench.workload_string(cfg, 128, 7, 5, 32, 'fp16')
    assert 'OUTSIDE' in bench.workload_string(cfg, 128, 7, 5, 32, 'bf16')
    a = bench.clevr6_arch()
    f_dec, f_cc = bench.flops_per_unit(a)
    assert abs(2 * f_dec + bench.refine_flops_per_unit(a) - 10.071e9) < 1e6          # F_unit
    assert abs(2 * bench.flops_per_unit(bench.bench_arch('dsprites'))[0]
               + bench.refine_flops_per_unit(bench.bench_arch('dsprites')) - 0.724e9) < 1e6
