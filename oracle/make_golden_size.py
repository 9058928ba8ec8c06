"""Golden vectors of the UNMODIFIED reference at the sizes bench.py publishes (build container only).

    python -m oracle.make_golden_size [case ...]

TEST INFRASTRUCTURE.  Same recipe as oracle/make_golden.py (seeded default weights, optional "sharpen" edit,
injected noise), but the cases are large (CLEVR6 B=32; K=11/T=7; 256x256 K=16), so the fixtures are COMPACT:
inputs and weights are regenerated from their seeds by the tests (checksums stored), and of the big outputs only
  * the per-step scalars (elbo, ll, kl) and the posterior of every step,
  * final z / posterior,
  * final pred, mask and mean at NS fixed pseudo-random positions (seeded permutation),
  * exact float64 sums of pred / mask / mean per image (slot, channel),
  * the per-pixel argmax of the final masks (uint8) together with the top-2 margin class (so that near-ties
    can be excluded from an argmax comparison)
are stored.
"""
import os
import sys

import numpy as np
import torch

from . import arch as A
from . import make_golden as MG
from . import ref_loader as R

OUT = MG.OUT
NS = 65536          # sampled positions per tensor
SAMPLE_SEED = 4242

# name -> dict(arch, over, B, sharpen, inputs)
#   inputs 'rand'   : x ~ U[0,1) (oracle.ref_loader.make_inputs)
#   inputs 'blocks' : block-structured image around the decoder's resting output (trained-like stress: the
#                     reconstruction error is small where the model is right, the likelihoods are peaked)
CASES = {
    # BASELINE config #2 at its full batch -- the configuration bench.py times
    'clevr6_b32_sharp': dict(arch='clevr6', over={}, B=32, sharpen=4.0, inputs='rand'),
    # BASELINE config #4 geometry: K=11, T=7 on the CLEVR6 architecture
    'clevr6_k11t7_b1_sharp': dict(arch='clevr6', over=dict(slots=11, iters=7), B=1, sharpen=4.0, inputs='rand'),
    # BASELINE config #5 geometry: 256x256, K=16 (T reduced to 2 to keep the reference run short)
    'clevr6_256_k16_t2_b1_sharp': dict(arch='clevr6', over=dict(slots=16, iters=2, img_size=256), B=1, sharpen=4.0,
                                       inputs='rand'),
    # trained-like stress for the 16-bit modes: structured image, strongly peaked masks, small sigma
    'clevr6_b1_stress_s10': dict(arch='clevr6', over=dict(sigma=0.10), B=1, sharpen=10.0, inputs='blocks'),
    'clevr6_b1_stress_s05': dict(arch='clevr6', over=dict(sigma=0.05), B=1, sharpen=10.0, inputs='blocks'),
}


def block_image(B, S, seed=7, lo=0.35, hi=0.65):
    """Piecewise-constant image: a coarse random grid of coloured blocks (values in [lo, hi]) plus a few
    rectangles, so that per-pixel likelihoods are structured instead of white noise."""
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(B, 3, S, S)
    cell = S // 8
    grid = lo + (hi - lo) * torch.rand(B, 3, 8, 8, generator=g)
    x.copy_(grid.repeat_interleave(cell, 2).repeat_interleave(cell, 3))
    for b in range(B):
        for _ in range(5):
            y0, x0 = [int(v) for v in torch.randint(0, S - S // 4, (2,), generator=g)]
            h, w = [int(v) for v in torch.randint(S // 16, S // 4, (2,), generator=g)]
            col = lo + (hi - lo) * torch.rand(3, generator=g)
            x[b, :, y0:y0 + h, x0:x0 + w] = col[:, None, None]
    return x


def case_inputs(name):
    c = CASES[name]
    arch = A.arch_by_name(c['arch'], **c['over'])
    x, eps = R.make_inputs(arch, c['B'])
    if c['inputs'] == 'blocks':
        x = block_image(c['B'], arch.IMG_SIZE)
    return arch, x, eps


def sample_index(n, ns=NS):
    g = torch.Generator().manual_seed(SAMPLE_SEED)
    if n <= ns:
        return torch.arange(n)
    return torch.randperm(n, generator=g)[:ns].sort().values


def compact(tr, arch, B):
    out = {}
    for t, st in enumerate(tr['steps']):
        for k in ('elbo', 'kl', 'll', 'post_mean', 'post_logvar'):
            out['s%d_%s' % (t, k)] = st[k].numpy()
    for k in ('post_mean', 'post_logvar', 'z'):
        out['final_' + k] = tr[k].numpy()
    pred, mask, mean = tr['pred'], tr['mask'], tr['mean']
    for nm, v in (('pred', pred), ('mask', mask), ('mean', mean)):
        flat = v.reshape(-1)
        out['final_%s_s' % nm] = flat[sample_index(flat.numel())].numpy()
    out['final_pred_sum'] = pred.double().sum(dim=(2, 3)).numpy()             # [B,3]
    out['final_mask_sum'] = mask.double().sum(dim=(2, 3, 4)).numpy()          # [B,K]
    out['final_mean_sum'] = mean.double().sum(dim=(3, 4)).numpy()             # [B,K,3]
    m = mask[:, :, 0]                                                         # [B,K,H,W]
    top2 = m.topk(2, dim=1)
    out['final_mask_argmax'] = top2.indices[:, 0].to(torch.uint8).numpy()
    out['final_mask_margin'] = (top2.values[:, 0] - top2.values[:, 1]).to(torch.float16).numpy()
    return out


def make_case(name):
    c = CASES[name]
    arch, x, eps = case_inputs(name)
    model = R.build_reference_model(arch, seed=0, sharpen=c['sharpen'])
    tr = R.run_reference_trace(model, x, eps, keep_aux=False)
    out = compact(tr, arch, c['B'])
    out['weights_checksum'] = np.float64(MG.weights_checksum(model.state_dict()))
    out['x_checksum'] = np.float64(x.double().sum().item())
    out['eps_checksum'] = np.float64(eps.double().abs().sum().item())
    finite = all(np.isfinite(v).all() for k, v in out.items() if v.dtype.kind == 'f')
    out['reference_finite'] = np.bool_(finite)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    return path, finite


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count() or 1)
    for n in (sys.argv[1:] or list(CASES)):
        import time
        t0 = time.time()
        p, finite = make_case(n)
        print('%s  %.1f KB  %.1f s  finite=%s' % (p, os.path.getsize(p) / 1024, time.time() - t0, finite), flush=True)
