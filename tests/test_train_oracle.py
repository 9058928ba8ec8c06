"""CPU: the explicit training gradients of oracle/train_restatement.py (the specification of the
weight-gradient kernels) against the UNMODIFIED reference's autograd -- ``loss = model(x); loss.backward()`` as
``lib/engine/train.py:60-64`` runs it -- in float64 (tight) and float32 (the reference's own precision)."""
import pytest
import torch

from oracle import arch as A
from oracle import ref_loader as R
from oracle import restatement as S
from oracle import train_restatement as TR

needs_reference = pytest.mark.skipif(not R.reference_available(), reason='reference tree not mounted')


def _rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


@needs_reference
@pytest.mark.parametrize('name,over,B,sharpen,dtype,tol', [
    ('tiny', {}, 3, 1.0, torch.float64, 1e-6),
    ('tiny', {}, 2, 5.0, torch.float64, 1e-6),
    ('test5x5', {}, 2, 3.0, torch.float64, 1e-6),                 # 5x5 kernels, other channel widths
    ('tiny', dict(iters=1), 2, 2.0, torch.float64, 1e-6),         # a single refiner call: no LSTM chain
    ('tiny', dict(img_size=20, ref_layers=3), 2, 2.0, torch.float64, 1e-6),   # odd stride-2 sizes (output_padding)
    ('tiny', {}, 3, 3.0, torch.float32, 2e-3),
])
def test_training_gradients_match_reference_autograd(name, over, B, sharpen, dtype, tol):
    arch = A.arch_by_name(name, **over)
    model = R.build_reference_model(arch, seed=0, sharpen=sharpen, dtype=dtype)
    x, eps = R.make_inputs(arch, B, dtype=dtype)
    ref_loss, ref_grads = R.run_reference_training_step(model, x, eps)
    sd = S.state_dict_to(model.state_dict(), dtype)
    loss, grads, elbos = TR.loss_and_grads(sd, arch, x, eps)
    assert abs(loss.item() - ref_loss.item()) <= tol * abs(ref_loss.item())
    assert set(grads) == set(ref_grads) == set(sd)
    for k in sd:
        assert grads[k].shape == ref_grads[k].shape, k
        assert _rel(grads[k], ref_grads[k]) < tol, (k, _rel(grads[k], ref_grads[k]))
    # every parameter takes part: no gradient of the reference is identically zero here (with a single refiner call
    # the LSTM starts from h = 0, so weight_hh alone gets none)
    assert all(ref_grads[k].abs().max() > 0 for k in sd if not (arch.ITERS == 1 and k == 'refine.lstm.weight_hh'))


@needs_reference
def test_forward_loss_matches_inference_restatement_elbos():
    """the T+1 ELBOs of the training step are the inference loop's ELBOs plus the final one"""
    arch = A.arch_by_name('tiny')
    model = R.build_reference_model(arch, seed=0, sharpen=2.0, dtype=torch.float64)
    x, eps = R.make_inputs(arch, 2, dtype=torch.float64)
    sd = S.state_dict_to(model.state_dict(), torch.float64)
    _, _, elbos = TR.loss_and_grads(sd, arch, x, eps)
    tr = S.encode_trace(sd, arch, x, eps)
    for i, st in enumerate(tr['steps']):
        assert abs(elbos[i].item() - st['elbo'].item()) <= 1e-12 * abs(st['elbo'].item())


def _shard_worker(rank, world, port, B, out_dir):
    import os
    import torch.distributed as dist
    from iodine_b200.parallel import shard_bounds
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        arch = A.arch_by_name('tiny')
        model = R.build_reference_model(arch, seed=0, sharpen=2.0, dtype=torch.float64)
        sd = S.state_dict_to(model.state_dict(), torch.float64)
        x, eps = R.make_inputs(arch, B, dtype=torch.float64)
        b0, b1 = shard_bounds(B, world, rank)
        loss, grads, _ = TR.loss_and_grads(sd, arch, x[b0:b1], eps[:, b0:b1], global_batch=B)
        flat = torch.cat([loss.reshape(1)] + [grads[k].reshape(-1) for k in sorted(grads)])
        dist.all_reduce(flat)                              # the training exchange: one sum over ranks
        torch.save(flat, os.path.join(out_dir, 'r%d.pt' % rank))
    finally:
        dist.destroy_process_group()


@needs_reference
@pytest.mark.parametrize('B', [4, 3])                      # equal shards and a ragged split
def test_sharded_gradients_sum_to_the_full_batch(B, tmp_path):
    """SURVEY.md 8e 'Training': whole-image shards + ONE sum all-reduce of the gradients (gloo, 2 ranks) reproduce
    the single-process gradients of the global batch."""
    import os
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_shard_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    arch = A.arch_by_name('tiny')
    model = R.build_reference_model(arch, seed=0, sharpen=2.0, dtype=torch.float64)
    sd = S.state_dict_to(model.state_dict(), torch.float64)
    x, eps = R.make_inputs(arch, B, dtype=torch.float64)
    loss, grads, _ = TR.loss_and_grads(sd, arch, x, eps)
    want = torch.cat([loss.reshape(1)] + [grads[k].reshape(-1) for k in sorted(grads)])
    for r in range(2):
        got = torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r))
        assert _rel(got, want) < 1e-12, r


# --------------------------------------------------------------------------- committed fixtures (no reference tree needed)
def _golden_case(name):
    import os
    import numpy as np
    from oracle import make_train_golden as MTG
    from helpers import GOLDEN, seeded_model
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    arch_name, over, B, sharpen = MTG.CASES[name]
    arch = A.arch_by_name(arch_name, **over)
    sd = S.state_dict_to(seeded_model(arch, sharpen).state_dict(), torch.float32)
    return g, arch, sd


@pytest.mark.parametrize('name', ['train_tiny_b2_sharp', 'train_test5x5_b2_sharp'])
def test_training_gradients_match_committed_reference_vectors(name):
    """the fixtures hold the reference's fp32 autograd gradients; the restatement runs in fp64 on the same fp32
    weights / inputs, so the residual is the reference's own fp32 rounding"""
    import numpy as np
    from oracle import make_golden as MG
    g, arch, sd = _golden_case(name)
    cs = MG.weights_checksum(sd)
    assert abs(cs - float(g['weights_checksum'])) <= 1e-9 * abs(cs)
    sd64 = S.state_dict_to(sd, torch.float64)
    x, eps = torch.from_numpy(g['x']).double(), torch.from_numpy(g['eps']).double()
    loss, grads, _ = TR.loss_and_grads(sd64, arch, x, eps)
    assert abs(loss.item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    for k in sd:
        ref = torch.from_numpy(np.asarray(g['grad/' + k])).double()
        assert _rel(grads[k], ref) < 2e-3, (k, _rel(grads[k], ref))


def test_training_restatement_against_the_clevr6_size_fixture():
    """tests/golden/train_clevr6_b2_t2.npz (oracle/make_train_golden_size.py): samples + sums of the reference's
    gradients at the CLEVR6 layer sizes; the restatement must reproduce them (the GPU tests use the same fixture)."""
    import os
    import numpy as np
    from oracle import make_golden as MG
    from oracle import make_golden_size as MS
    from oracle import make_train_golden_size as TS
    from helpers import GOLDEN, seeded_model
    g = dict(np.load(os.path.join(GOLDEN, TS.NAME + '.npz')))
    arch, x, eps = TS.case_inputs()
    assert abs(x.double().sum().item() - float(g['x_checksum'])) <= 1e-9 * abs(float(g['x_checksum']))
    sd = S.state_dict_to(seeded_model(arch, TS.SHARPEN).state_dict(), torch.float32)
    cs = MG.weights_checksum(sd)
    assert abs(cs - float(g['weights_checksum'])) <= 1e-9 * abs(cs)
    loss, grads, _ = TR.loss_and_grads(sd, arch, x, eps)
    assert abs(loss.item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    for k in sd:
        flat = grads[k].reshape(-1).double()
        want = torch.from_numpy(g['grad_s/' + k]).double()
        gmax = float(g['grad_max/' + k])
        assert ((flat[MS.sample_index(flat.numel(), TS.NS)] - want).abs().max() / gmax).item() < 2e-3, k
        assert abs(flat.sum().item() - float(g['grad_sum/' + k])) / float(g['grad_abs/' + k]) < 2e-3, k
