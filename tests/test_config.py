"""CPU: the yacs-free configuration loader (iodine_b200.config.load_config) on the reference's own experiment files
(where the reference tree is mounted) and on a synthetic one."""
import os

import pytest

from oracle import ref_loader as R

from iodine_b200.config import arch_by_name, default_config, load_config
from iodine_b200.modeling import IODINE, make_model

CONFIGS = os.path.join(R.REFERENCE_ROOT, 'configs')


@pytest.mark.skipif(not os.path.isdir(CONFIGS), reason='reference tree not mounted')
@pytest.mark.parametrize('fname,arch_name,over', [('clevr6_prop.yaml', 'clevr6', {}), ('dsprites_noclip.yaml', 'dsprites', None)])
def test_reference_experiment_files_load(fname, arch_name, over):
    cfg = load_config(os.path.join(CONFIGS, fname), ['MODEL.DEVICE', 'cpu', 'MODEL.PARALLEL', 'False', 'ARCH.ITERS', '3'])
    assert cfg.MODEL.NAME == 'IODINE' and cfg.MODEL.DEVICE == 'cpu' and cfg.MODEL.PARALLEL is False and cfg.ARCH.ITERS == 3
    want = arch_by_name(arch_name, iters=3)
    for k in ('SLOTS', 'SIGMA', 'DIM_LATENT', 'IMG_SIZE', 'IMG_CHANNELS', 'LAYERNORM'):
        assert getattr(cfg.ARCH, k) == getattr(want, k), k
    for sub in ('REF', 'DEC'):
        for k, v in vars(getattr(want, sub)).items():
            assert getattr(getattr(cfg.ARCH, sub), k) == v, (sub, k)
    assert sorted(cfg.ARCH.ENCODING) == sorted(want.ENCODING)
    assert cfg.DATASET.TEST in ('CLEVR', 'DSPRITES') and isinstance(cfg.TENSORBOARD.TARGETS.IMAGE, list)
    model = make_model(cfg)                               # the loader's tree is what make_model(cfg) reads
    assert isinstance(model, IODINE) and model.K == want.SLOTS


def test_merge_rules(tmp_path):
    p = tmp_path / 'exp.yaml'
    p.write_text('EXP:\n  NAME: mine\nARCH:\n  SLOTS: 4\n  REF:\n    CONV_CHAN: 16\nTRAIN:\n  BASE_LR: 3e-4\n')
    cfg = load_config(str(p), ['ARCH.SIGMA', '0.2', 'MODEL.NAME', 'IODINE', 'ARCH.ENCODING', "['image', 'mask']"])
    assert cfg.EXP.NAME == 'mine' and cfg.ARCH.SLOTS == 4 and cfg.ARCH.REF.CONV_CHAN == 16 and cfg.ARCH.REF.STRIDE == 2
    assert cfg.TRAIN.BASE_LR == pytest.approx(3e-4) and cfg.ARCH.SIGMA == 0.2 and cfg.ARCH.ENCODING == ['image', 'mask']
    assert cfg.MODEL.NAME == 'IODINE' and cfg.DATALOADER.NUM_WORKERS == 4          # untouched default
    with pytest.raises(AttributeError):
        cfg.merge_from_list(['ARCH.SLOTS', '9'])                                    # frozen
    bad = tmp_path / 'bad.yaml'
    bad.write_text('ARCH:\n  NO_SUCH_KEY: 1\n')
    with pytest.raises(KeyError):
        load_config(str(bad))
    with pytest.raises(KeyError):
        load_config('', ['NOPE.X', '1'])
    d = default_config()
    assert d.ARCH.DEC.KERNEL_SIZE == 5 and d.MODEL.NAME == 'VAE' and d.GETTER == 'VAE'
