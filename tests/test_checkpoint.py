"""CPU: ``iodine_b200.utils.checkpoint.Checkpointer`` -- same contract and on-disk format as
``lib/utils/checkpoint.py``; where the reference tree is mounted, directories written by one are resumed by the
other and the rotation of old files is compared step by step."""
import os
import pickle
import sys

import pytest
import torch

from oracle import arch as A
from oracle import ref_loader as R

from helpers import seeded_model
from iodine_b200.modeling import make_model
from iodine_b200.utils.checkpoint import Checkpointer
from types import SimpleNamespace as NS


def _opt(model):
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    return opt, sched


def test_save_load_round_trip_and_rotation(tmp_path, capsys):
    arch = A.arch_by_name('tiny')
    m = seeded_model(arch, 2.0)
    opt, sched = _opt(m)
    d = str(tmp_path / 'run' / 'exp')                     # parents are created
    ck = Checkpointer(m, opt, sched, args={'epoch': 0}, max_checkpoints=2, save_dir=d)
    assert not ck.has_checkpoint() and ck.load() == {} and 'No checkpoint found.' in capsys.readouterr().out
    for e in range(4):
        ck.args['epoch'] = e
        sched.step()
        ck.save('model_{:03d}'.format(e))
    with open(os.path.join(d, 'checkpoint.pkl'), 'rb') as f:
        assert pickle.load(f) == ['model_002.pth', 'model_003.pth']
    assert sorted(os.listdir(d)) == ['checkpoint.pkl', 'model_002.pth', 'model_003.pth']
    assert ck.get_checkpoint_file() == os.path.join(d, 'model_003.pth')
    raw = torch.load(os.path.join(d, 'model_003.pth'))
    assert set(raw) == {'model', 'optimizer', 'scheduler', 'epoch'} and raw['epoch'] == 3
    # resume into a fresh model / optimizer; f= is ignored when the index exists (checkpoint.py:57-59)
    m2 = seeded_model(arch, 1.0)
    opt2, sched2 = _opt(m2)
    args2 = {}
    Checkpointer(m2, opt2, sched2, args=args2, save_dir=d).load(f='/nonexistent.pth')
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert args2 == {'epoch': 3} and sched2.state_dict() == sched.state_dict()
    assert opt2.state_dict()['param_groups'][0]['lr'] == opt.state_dict()['param_groups'][0]['lr']
    # saving the same name twice must not delete the live file when the older entry rotates out
    ck1 = Checkpointer(m, save_dir=str(tmp_path / 'same'), max_checkpoints=1)
    ck1.save('model')
    ck1.save('model')
    assert os.path.exists(os.path.join(str(tmp_path / 'same'), 'model.pth'))


def test_module_prefix_is_repaired_both_ways(tmp_path):
    arch = A.arch_by_name('tiny')
    cfg = lambda par: NS(MODEL=NS(NAME='IODINE', DEVICE='cpu', PARALLEL=par), ARCH=arch)
    plain, wrapped = make_model(cfg(False)), make_model(cfg(True))
    with torch.no_grad():
        plain.decoder.conv.weight.add_(1.0)
    Checkpointer(plain, save_dir=str(tmp_path / 'a')).save('m')
    Checkpointer(wrapped, save_dir=str(tmp_path / 'a')).load()
    assert torch.equal(wrapped.module.decoder.conv.weight, plain.decoder.conv.weight)
    with torch.no_grad():
        wrapped.module.posterior.init_mean.add_(2.0)
    Checkpointer(wrapped, save_dir=str(tmp_path / 'b')).save('m')
    assert next(iter(torch.load(str(tmp_path / 'b' / 'm.pth'))['model'])).startswith('module.')
    Checkpointer(plain, save_dir=str(tmp_path / 'b')).load()
    assert torch.equal(plain.posterior.init_mean, wrapped.module.posterior.init_mean)


@pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')
def test_interoperates_with_the_live_reference_checkpointer(tmp_path):
    if R.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, R.REFERENCE_ROOT)
    from lib.utils.checkpoint import Checkpointer as RefCheckpointer
    arch = A.arch_by_name('tiny')
    ref_model = R.build_reference_model(arch, seed=3, sharpen=2.0)
    ours = seeded_model(arch, 1.0)
    # reference writes, we resume
    d1 = str(tmp_path / 'ref_writes')
    rc = RefCheckpointer(ref_model, *_opt(ref_model), args={'epoch': 5}, max_checkpoints=2, save_dir=d1)
    for n in ('a', 'b', 'c'):
        rc.save(n)
    args = {}
    Checkpointer(ours, *_opt(ours), args=args, save_dir=d1).load()
    assert args == {'epoch': 5}
    for (ka, va), (kb, vb) in zip(ref_model.state_dict().items(), ours.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    # we write, the reference resumes; and the two rotate files identically
    d2, d3 = str(tmp_path / 'we_write'), str(tmp_path / 'ref_mirror')
    with torch.no_grad():
        ours.refine.lstm.bias_hh.add_(0.5)
    oc = Checkpointer(ours, args={'it': 1}, max_checkpoints=2, save_dir=d2)
    rc2 = RefCheckpointer(ours, args={'it': 1}, max_checkpoints=2, save_dir=d3)
    for n in ('x', 'y', 'x', 'z', 'z'):
        oc.save(n)
        rc2.save(n)
        assert sorted(os.listdir(d2)) == sorted(os.listdir(d3)), n
        with open(os.path.join(d2, 'checkpoint.pkl'), 'rb') as f2, open(os.path.join(d3, 'checkpoint.pkl'), 'rb') as f3:
            assert pickle.load(f2) == pickle.load(f3), n
    fresh = R.build_reference_model(arch, seed=0, sharpen=1.0)
    rargs = {}
    RefCheckpointer(fresh, args=rargs, save_dir=d2).load()
    assert rargs == {'it': 1} and torch.equal(fresh.refine.lstm.bias_hh, ours.refine.lstm.bias_hh)
