"""GPU: the logger side channel (reference lib/modeling/iodine.py:225-239).

The reference writes image / pred / kl / likelihood / mask_i / pred_i into the process-global logger inside EVERY
elbo() call, each write replacing the last one: after reconstruct() a consumer finds the quantities of the LAST
in-loop elbo() (refinement step T-1, image 0), not those of the final decode."""
import pytest
import torch

from iodine_b200.utils.vis_logger import logger

from helpers import golden_state_dict, rel_err, t

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('name', ['tiny_b2_sharp', 'test5x5_b2_sharp'])
def test_logger_holds_last_in_loop_elbo_after_reconstruct(name):
    g, arch, B, sd, model = golden_state_dict(name)
    model.to(DEV)
    logger.things.clear()
    x = t(g['x']).to(DEV)
    model.reconstruct(x, eps=t(g['eps']).to(DEV))
    torch.cuda.synchronize()
    K, T = arch.SLOTS, arch.ITERS
    keys = ['image', 'pred', 'kl', 'likelihood'] + ['mask_%d' % i for i in range(K)] + ['pred_%d' % i for i in range(K)]
    for k in keys:
        assert k in logger, k
    last = lambda k: t(g['s%d_%s' % (T - 1, k)])
    mask, mean = last('mask'), last('mean')                    # [B,K,1,H,W], [B,K,3,H,W] of step T-1
    assert rel_err(logger['image'], g['x'][0]) == 0
    assert rel_err(logger['pred'], (mask * mean).sum(dim=1)[0]) < 1e-3
    for i in range(K):
        assert rel_err(logger['mask_%d' % i], mask[0, i, 0]) < 1e-3
        assert rel_err(logger['pred_%d' % i], mean[0, i]) < 1e-3
        assert tuple(logger['mask_%d' % i].shape) == tuple(mask[0, i, 0].shape)
        assert tuple(logger['pred_%d' % i].shape) == tuple(mean[0, i].shape)
    assert abs(float(logger['kl']) - float(last('kl'))) < 1e-3 * max(1.0, abs(float(last('kl'))))
    assert abs(float(logger['likelihood']) - float(last('ll'))) < 1e-3 * abs(float(last('ll')))


def test_logger_after_forward_holds_the_final_elbo():
    """forward() ends with one more elbo() on the final posterior (iodine.py:146-147): z = sample(eps[T]), which is
    what decode() sees in reconstruct() -- so the logged masks equal the final masks of the golden"""
    g, arch, B, sd, model = golden_state_dict('tiny_b2_sharp')
    model.to(DEV)
    logger.things.clear()
    model(t(g['x']).to(DEV), eps=t(g['eps']).to(DEV))
    torch.cuda.synchronize()
    for i in range(arch.SLOTS):
        assert rel_err(logger['mask_%d' % i], t(g['final_mask'])[0, i, 0]) < 1e-3
        assert rel_err(logger['pred_%d' % i], t(g['final_mean'])[0, i]) < 1e-3
    assert rel_err(logger['pred'], t(g['final_pred'])[0]) < 1e-3
    assert 'init_mean' in logger and 'init_logvar' in logger
