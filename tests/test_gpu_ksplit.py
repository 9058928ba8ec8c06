"""GPU, two ranks: K-split (SURVEY.md 8e fallback; include/iodine_b200.h IodineShape.slot_ranks).

For a batch smaller than the number of GPUs the ranks own SLOTS instead of images; the K-way reductions of
IODINE.elbo (reference iodine.py:185, 213-216, 292, 324) then cross ranks through one in-library ncclAllGather of the
decoder's 4-channel output per elbo() evaluation.  A 2-rank run must reproduce the single-GPU result of the same
image(s): B=1 with K=16 on the CLEVR6 layer sizes (config #5's slot count), and B=2 K=4 on the tiny arch (the
rank-major layout of the gathered buffer differs from the image-major one as soon as B > 1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import arch as A

from helpers import rel_err, seeded_model
from iodine_b200 import _cabi
from iodine_b200.parallel import KSplit

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
CASES = {'clevr6_b1_k16': ('clevr6', dict(slots=16, iters=2), 1, 'fp32'),
         'clevr6_b1_k16_fp16': ('clevr6', dict(slots=16, iters=2), 1, 'fp16'),
         'tiny_b2_k4': ('tiny', dict(slots=4), 2, 'fp32')}


def _inputs(arch, B):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=g)
    eps = torch.randn(arch.ITERS + 1, B, arch.SLOTS, arch.DIM_LATENT, generator=g)
    return x, eps


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        name, over, B, prec = CASES[case]
        arch = A.arch_by_name(name, **over)
        model = seeded_model(arch, 3.0, precision=prec).to(dev)
        x, eps = _inputs(arch, B)
        ks = KSplit(model)
        assert ks.slots() == (rank * arch.SLOTS // world, (rank + 1) * arch.SLOTS // world)
        pred, mask, mean = ks.reconstruct(x.to(dev), eps.to(dev))
        res = {'pred': pred.cpu(), 'mask': mask.cpu(), 'mean': mean.cpu(), 'elbo': ks.elbo_per_step().cpu(),
               'terms': ks.elbo_terms.cpu(), 'z_local': model.z.cpu()}
        res['z'] = ks.encode(x.to(dev), eps.to(dev), gather=True).cpu()
        torch.save(res, os.path.join(out_dir, 'r%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('case', sorted(CASES))
def test_two_rank_k_split_matches_single_gpu(case, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    name, over, B, prec = CASES[case]
    arch = A.arch_by_name(name, **over)
    model = seeded_model(arch, 3.0, precision=prec).to(DEV)
    x, eps = _inputs(arch, B)
    pred, mask, mean = model.reconstruct(x.to(DEV), eps=eps.to(DEV))
    elbo, z = model.elbo_per_step(B).cpu(), model.z.cpu()
    outs = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r)) for r in range(world)]
    kl = arch.SLOTS // world
    tol = 1e-5 if prec == 'fp32' else 2e-4          # 16-bit seeds: atomics in another order move the last bits
    for r, o in enumerate(outs):
        assert tuple(o['mask'].shape) == tuple(mask.shape) and tuple(o['z_local'].shape) == (B, kl, arch.DIM_LATENT)
        assert rel_err(o['pred'], pred.cpu()) < tol, r
        assert rel_err(o['mask'], mask.cpu()) < tol, r
        assert rel_err(o['mean'], mean.cpu()) < tol, r
        assert rel_err(o['elbo'], elbo) < tol, r
        assert rel_err(o['z_local'], z[:, r * kl:(r + 1) * kl]) < tol, r
        assert rel_err(o['z'], z) < tol, r
    assert torch.equal(outs[0]['terms'], outs[1]['terms'])      # all-reduced: identical everywhere
    assert torch.equal(outs[0]['pred'], outs[1]['pred'])        # same gathered buffer, same kernel


def test_k_split_plan_validation():
    """one GPU: the plan refuses slot counts that do not divide and steps without a communicator"""
    from iodine_b200.engine import RefinementEngine
    arch = A.arch_by_name('tiny', slots=4)
    with pytest.raises(_cabi.IodineError):
        RefinementEngine(A.arch_by_name('tiny', slots=3), 1, DEV, 'fp32', slot_split=(0, 2))
    eng = RefinementEngine(arch, 1, DEV, 'fp32', slot_split=(1, 2))
    assert eng.K == 2 and eng.K_total == 4
    eng.set_weights(seeded_model(arch).to(DEV).state_dict())
    x, eps = _inputs(arch, 1)
    with pytest.raises(_cabi.IodineError, match='communicator'):
        eng.encode(x.to(DEV), eps[:, :, 2:4].to(DEV))
