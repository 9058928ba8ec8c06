"""GPU: the in-library exchange of the path (iodine_plan_set_comm, include/iodine_b200.h).

* one rank: a single-rank NCCL communicator makes the all-reduce the identity, so encode()/reconstruct()/
  reconstruct_host() must return exactly what they return without a communicator -- this exercises the run-time
  NCCL resolution and the stream-ordered call on any one-GPU box;
* two ranks (skipped with fewer than two GPUs): SlotShard in native mode on NCCL must reproduce the single-process
  result of the same global batch, for equal and ragged shards.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import arch as A

from helpers import rel_err, seeded_model
from iodine_b200 import _cabi
from iodine_b200.parallel import NcclComm, SlotShard, shard_bounds

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _inputs(arch, B):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=g)
    eps = torch.randn(arch.ITERS + 1, B, arch.SLOTS, arch.DIM_LATENT, generator=g)
    return x, eps


def test_single_rank_communicator_is_the_identity():
    arch = A.arch_by_name('tiny')
    B = 3
    model = seeded_model(arch, 3.0).to(DEV)
    x, eps = _inputs(arch, B)
    x, eps = x.to(DEV), eps.to(DEV)
    eng = model.state_for_debug(B)
    ref = eng.reconstruct(x, eps)
    ref_terms = ref[4].clone()
    comm = NcclComm(rank=0, world=1)
    eng.set_comm(comm, 0, 1)
    for _ in range(3):                                   # eager call, graph capture, graph replay
        out = eng.reconstruct(x, eps)
        torch.cuda.synchronize()
        assert rel_err(out[4], ref_terms) < 1e-5
        assert rel_err(out[0], ref[0]) < 1e-5
    z, terms, _ = eng.encode(x, eps)
    assert rel_err(terms, ref_terms) < 1e-5
    host = eng.reconstruct_host(x.cpu().pin_memory(), eps.cpu().pin_memory())
    assert rel_err(host['terms'], ref_terms.cpu()) < 1e-5
    eng.set_comm(None)
    assert rel_err(eng.encode(x, eps)[1], ref_terms) < 1e-5
    comm.close()


def test_set_comm_rejects_bad_ranks():
    arch = A.arch_by_name('tiny')
    eng = seeded_model(arch).to(DEV).state_for_debug(1)
    comm = NcclComm(rank=0, world=1)
    with pytest.raises(_cabi.IodineError):
        eng.set_comm(comm, 2, 2)
    with pytest.raises(_cabi.IodineError):
        eng.set_comm(comm, 0, 0)
    comm.close()


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        arch = A.arch_by_name('tiny')
        model = seeded_model(arch, 3.0).to(dev)
        x, eps = _inputs(arch, B)
        sh = SlotShard(model)
        assert sh.native, 'NCCL process group: the library should own the all-reduce'
        pred, mask, mean = sh.reconstruct(x.to(dev), eps.to(dev), gather=True)
        res = {'pred': pred.cpu(), 'mask': mask.cpu(), 'elbo': sh.elbo_per_step().cpu(), 'terms': sh.elbo_terms.cpu()}
        b0, b1 = shard_bounds(B, world, rank)
        sh.encode(x[b0:b1].to(dev), eps[:, b0:b1].to(dev), local=True)      # replayed graph + all-reduce
        res['elbo_local'] = sh.elbo_per_step().cpu()
        torch.save(res, os.path.join(out_dir, 'r%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('B', [4, 3])          # equal shards and a ragged split
def test_two_rank_nccl_in_library_allreduce_matches_single_process(B, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    arch = A.arch_by_name('tiny')
    model = seeded_model(arch, 3.0).to(DEV)
    x, eps = _inputs(arch, B)
    single = SlotShard(model)
    pred, mask, _ = single.reconstruct(x.to(DEV), eps.to(DEV))
    elbo = single.elbo_per_step().cpu()
    outs = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r)) for r in range(world)]
    for r, o in enumerate(outs):
        assert rel_err(o['pred'], pred.cpu()) < 1e-5, r
        assert rel_err(o['mask'], mask.cpu()) < 1e-5, r
        assert rel_err(o['elbo'], elbo) < 1e-5, r
        assert rel_err(o['elbo_local'], elbo) < 1e-5, r
    assert torch.equal(outs[0]['terms'], outs[1]['terms'])      # all-reduced: identical everywhere


# --------------------------------------------------------------------------- training: ONE flat gradient all-reduce
def _train_worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        arch = A.arch_by_name('tiny')
        model = seeded_model(arch, 3.0).to(dev)
        x, eps = _inputs(arch, B)
        sh = SlotShard(model)                       # installs the communicator on the model's engines
        b0, b1 = shard_bounds(B, world, rank)
        model.global_batch = B
        loss = model(x[b0:b1].to(dev), eps=eps[:, b0:b1].to(dev))
        loss.mean().backward()
        torch.cuda.synchronize()
        res = {'loss': loss.detach().cpu(), 'grads': {k: p.grad.cpu() for k, p in model.named_parameters()}}
        torch.save(res, os.path.join(out_dir, 'r%d.pt' % rank))
        del sh
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('B', [4, 3])
def test_two_rank_training_step_all_reduces_the_gradients(B, tmp_path):
    """slot-shard training (SURVEY.md 8e "Training"): every rank ends with the FULL-batch loss and gradients -- what
    DataParallel's gather + reduce-to-GPU-0 gives the reference (lib/engine/train.py:60-65)"""
    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    arch = A.arch_by_name('tiny')
    model = seeded_model(arch, 3.0).to(DEV)
    x, eps = _inputs(arch, B)
    loss = model(x.to(DEV), eps=eps.to(DEV))
    loss.mean().backward()
    outs = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r)) for r in range(world)]
    for r, o in enumerate(outs):
        assert rel_err(o['loss'], loss.detach().cpu()) < 1e-5, r
        for k, p in model.named_parameters():
            assert rel_err(o['grads'][k], p.grad.cpu()) < 2e-4, (r, k)
    for k in outs[0]['grads']:
        assert torch.equal(outs[0]['grads'][k], outs[1]['grads'][k]), k      # all-reduced: identical everywhere
