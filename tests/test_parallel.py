"""CPU (gloo, world_size 2): the slot-shard host logic of iodine_b200.parallel.

The native engine is CUDA-only, so the per-rank "model" here is the oracle restatement wrapped in
the IODINE method surface (test infrastructure standing in for the engine); what is under test is
the sharding, the single [T,2] all-reduce and the gathers -- a 2-rank run must reproduce the
single-process result of the same global batch.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import arch as A
from oracle import restatement as S

from helpers import seeded_model
from iodine_b200.parallel import SlotShard, shard_bounds, shard_table


class OracleModel:
    """IODINE-shaped stand-in: reconstruct/encode through oracle.restatement on CPU."""

    def __init__(self, arch, sd):
        self.arch, self.sd = arch, sd
        self.elbo_terms = None

    def _run(self, x, eps):
        B = x.shape[0]
        tr = S.encode_trace(self.sd, self.arch, x, eps)
        self.elbo_terms = torch.stack([torch.stack((s['ll'] * B, s['kl'] * B)) for s in tr['steps']])
        return tr

    def reconstruct(self, x, eps=None):
        tr = self._run(x, eps)
        return tr['pred'], tr['mask'], tr['mean']

    def encode(self, x, eps=None):
        return self._run(x, eps)['z']


def _inputs(arch, B):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, arch.IMG_SIZE, arch.IMG_SIZE, generator=g)
    eps = torch.randn(arch.ITERS + 1, B, arch.SLOTS, arch.DIM_LATENT, generator=g)
    return x, eps


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        arch = A.arch_by_name('tiny')
        sd = S.state_dict_to(seeded_model(arch, 3.0).state_dict(), torch.float32)
        x, eps = _inputs(arch, B)
        sh = SlotShard(OracleModel(arch, sd))
        assert not sh.native                      # gloo: the [T,2] table goes through dist.all_reduce
        pred, mask, mean = sh.reconstruct(x, eps, gather=True)
        res = {'pred': pred, 'mask': mask, 'elbo': sh.elbo_per_step(), 'terms': sh.elbo_terms,
               'gb': sh.global_batch}
        # local=True: the caller already holds its shard
        b0, b1 = shard_bounds(B, world, rank)
        z_local = sh.encode(x[b0:b1], eps[:, b0:b1], local=True)
        res['z_rows'] = z_local.shape[0]
        res['gb_local'] = sh.global_batch
        res['elbo_local'] = sh.elbo_per_step()
        torch.save(res, os.path.join(out_dir, 'r%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_the_batch():
    for B in (1, 2, 5, 8, 33):
        for w in (1, 2, 3, 8):
            tab = shard_table(B, w)
            assert tab[0][0] == 0 and tab[-1][1] == B
            for (a0, a1), (b0, b1) in zip(tab, tab[1:]):
                assert a1 == b0 and a1 - a0 >= b1 - b0 >= 0
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


@pytest.mark.parametrize('B', [4, 3])          # equal shards and a ragged split
def test_two_rank_gloo_matches_single_process(B, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    arch = A.arch_by_name('tiny')
    sd = S.state_dict_to(seeded_model(arch, 3.0).state_dict(), torch.float32)
    x, eps = _inputs(arch, B)
    single = SlotShard(OracleModel(arch, sd))
    pred, mask, _ = single.reconstruct(x, eps)
    elbo = single.elbo_per_step()
    outs = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r)) for r in range(world)]
    for r, o in enumerate(outs):
        assert o['gb'] == B and o['gb_local'] == B
        assert torch.allclose(o['pred'], pred, atol=1e-6), r
        assert torch.allclose(o['mask'], mask, atol=1e-6), r
        assert torch.allclose(o['elbo'], elbo, rtol=1e-5), r
        assert torch.allclose(o['elbo_local'], elbo, rtol=1e-5), r
        b0, b1 = shard_bounds(B, world, r)
        assert o['z_rows'] == b1 - b0
    assert torch.equal(outs[0]['terms'], outs[1]['terms'])      # all-reduced: identical everywhere
