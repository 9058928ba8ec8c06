"""GPU tests of the tcgen05 tensor-core decoder path (precision='fp16' / 'bf16', conv_tc.cu).

Two bars:
  * kernel-level: every decoder buffer of one refinement step (activations, 4-channel output,
    data-gradient, class sums, dz) against the exact-fp32 FFMA path of the same library on the same
    inputs, within bf16 rounding accumulated over the layer stack;
  * path-level (the north-star bar): recon / masks / ELBO of a whole reconstruct() within 1e-3 of
    the committed golden vectors of the unmodified reference.
"""
import pytest
import torch

from oracle import arch as A


from helpers import load_golden, rel_err, seeded_model, t

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL_OUT = 1e-3


def _pair(arch, B, sharpen=1.0, prec='bf16'):
    m32 = seeded_model(arch, sharpen).to(DEV)
    m16 = seeded_model(arch, sharpen, precision=prec).to(DEV)
    return m32, m16, m32.state_for_debug(B), m16.state_for_debug(B)


def _nrm(a, b):
    """relative L2 error (bf16 noise is judged in the mean, not on the worst element)."""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize('name,over,B', [
    ('tiny', dict(dec_layers=3), 2),              # C=16, W=16: flat tiling, several rows per tile
    ('dsprites', dict(iters=1), 1),               # C=32, W=64, 5 layers
    ('test5x5', dict(iters=1), 2),                # 5x5 taps, C=32, W=32
    ('clevr6', dict(iters=1, slots=2), 1),        # C=64, W=128: row-aligned tiles
    ('tiny', dict(img_size=24, dec_chan=32), 3),  # odd width
    ('tiny', dict(dec_layers=1), 2),              # no C->C layer: only the 4-channel ends
    ('tiny', dict(img_size=128, dec_chan=32, dec_layers=3, slots=2, iters=1), 1),   # row-streaming, N = 3*32
    ('tiny', dict(img_size=128, dec_chan=16, dec_layers=3, slots=2, iters=1), 1),   # row-streaming, N = 3*16
    ('tiny', dict(img_size=128, dec_chan=64, dec_layers=2, slots=3, iters=1), 2),   # row-streaming, odd item counts
    ('tiny', dict(img_size=256, dec_chan=16, dec_layers=3, slots=2, iters=1), 1),   # row-streaming, two 128-column segments
    ('tiny', dict(img_size=256, dec_chan=64, dec_layers=2, slots=1, iters=1), 1),   # BASELINE config #5 geometry (W=256, C=64)
])
@pytest.mark.parametrize('prec', ['bf16', 'fp16'])
def test_decoder_buffers_against_fp32_path(name, over, B, prec):
    arch = A.arch_by_name(name, **over)
    m32, m16, e32, e16 = _pair(arch, B, sharpen=2.0, prec=prec)
    K, L, H, C = arch.SLOTS, arch.DIM_LATENT, arch.IMG_SIZE, arch.DEC.CONV_CHAN
    g = torch.Generator().manual_seed(11)
    x = torch.rand(B, 3, H, H, generator=g).to(DEV)
    eps = torch.randn(B, K, L, generator=g).to(DEV)
    outs = []
    for eng in (e32, e16):
        mu, lv, h, c = eng.init_state()
        terms, _, _ = eng.refine_step(x, eps, mu, lv, h, c)
        torch.cuda.synchronize()
        d = {'terms': terms.clone(), 'mu': mu.clone(), 'lv': lv.clone()}
        for i in range(arch.DEC.CONV_LAYERS):
            d['act%d' % i] = eng.debug_read('act%d' % i).clone()
        for nm in ('out4', 'seed4', 'G', 'dz', 'pool'):
            d[nm] = eng.debug_read(nm).clone()
        outs.append(d)
    ref, got = outs
    errs = {k: _nrm(got[k], ref[k]) for k in ref}
    print(prec, name, over, {k: '%.2e' % v for k, v in errs.items()})
    n = arch.DEC.CONV_LAYERS
    u = 4e-3 if prec == 'bf16' else 5e-4             # unit roundoff of the 16-bit format (2^-8 / 2^-11)
    assert errs['act0'] < u                          # rounding of an exact value
    for i in range(1, n):
        assert errs['act%d' % i] < u * (i + 2), ('act', i, errs)
    assert errs['out4'] < u * (n + 2), errs
    assert errs['seed4'] < 12 * u, errs              # gradient seeds amplify by 1/sigma^2
    assert errs['G'] < 12 * u and errs['dz'] < 12 * u, errs
    assert errs['pool'] < 12 * u, errs             # tcgen05 refinement encoder (refine_tc.cu) vs FFMA
    assert errs['terms'] < 1e-3, errs


@pytest.mark.parametrize('name', ['tiny_b2', 'tiny_b2_sharp', 'dsprites_b2', 'dsprites_b2_sharp',
                                  'clevr6_b1', 'clevr6_b1_sharp', 'test5x5_b2_sharp'])
@pytest.mark.parametrize('prec', ['fp16', 'bf16'])
def test_reconstruct_16bit_against_golden(name, prec):
    g, arch, B, sharpen, detail = load_golden(name)
    model = seeded_model(arch, sharpen, precision=prec).to(DEV)
    pred, mask, mean = model.reconstruct(t(g['x']).to(DEV), eps=t(g['eps']).to(DEV))
    torch.cuda.synchronize()
    e = {'pred': rel_err(pred, g['final_pred']), 'mask': rel_err(mask, g['final_mask']),
         'mean': rel_err(mean, g['final_mean'])}
    elbo = model.elbo_per_step(B).cpu()
    e['elbo'] = max(abs(elbo[i].item() - float(g['s%d_elbo' % i])) / abs(float(g['s%d_elbo' % i]))
                    for i in range(arch.ITERS))
    e['z_l2'] = _nrm(model.z, t(g['final_z']))
    print(prec, name, {k: '%.2e' % v for k, v in e.items()})
    # fp16 (10-bit mantissa, like TF32) has to meet the north-star bar of 1e-3 on recon / masks /
    # ELBO; bf16 (7-bit mantissa) is the throughput mode of BASELINE config #3 and is held to 1e-2
    # (its measured errors are printed and recorded in DESIGN.md)
    bar = TOL_OUT if prec == 'fp16' else 1e-2
    assert e['pred'] < bar and e['mask'] < bar and e['elbo'] < bar, e
    assert e['mean'] < 5 * bar and e['z_l2'] < 50 * bar, e


@pytest.mark.parametrize('prec,bar', [('fp16', TOL_OUT), ('bf16', 1e-2)])
def test_16bit_full_size_matches_fp32_path_clevr6_b4(prec, bar):
    """BASELINE config #2 architecture at B=4: tensor-core path vs the exact path."""
    arch = A.arch_by_name('clevr6')
    B = 4
    m32, m16, _, _ = _pair(arch, B, sharpen=1.0, prec=prec)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 3, 128, 128, generator=g).to(DEV)
    eps = torch.randn(arch.ITERS + 1, B, arch.SLOTS, arch.DIM_LATENT, generator=g).to(DEV)
    p32, k32, _ = m32.reconstruct(x, eps=eps)
    e32 = m32.elbo_per_step(B).clone()
    p16, k16, _ = m16.reconstruct(x, eps=eps)
    e16 = m16.elbo_per_step(B).clone()
    errs = {'pred': rel_err(p16, p32), 'mask': rel_err(k16, k32), 'elbo': rel_err(e16, e32)}
    print(prec, {k: '%.2e' % v for k, v in errs.items()})
    assert max(errs.values()) < bar, errs


@pytest.mark.parametrize('name,over,B,prec', [
    ('clevr6', dict(iters=2, slots=3), 2, 'fp16'),                          # the fused kernel's flagship shape
    ('dsprites', dict(iters=2), 2, 'tf32'),                                  # 64x64: Wo = 32, one narrow segment
    ('tiny', dict(img_size=256, dec_chan=16, dec_layers=2, slots=9, iters=1, ref_chan=32), 1, 'fp16'),   # two segments, K > 8
    ('tiny', dict(iters=2), 3, 'bf16'),                                      # 16x16: tiny items (TR = 1)
    ('tiny', dict(img_size=144, slots=2, iters=1, ref_chan=16), 1, 'fp16'),  # Wo = 72: a full and a partial column segment
    ('tiny', dict(img_size=20, slots=2, iters=2), 2, 'tf32'),                # Ho = 10: a short last strip
])
def test_fused_aux_path_matches_separate_assembly(name, over, B, prec, monkeypatch):
    """mixture_fast_kernel<FUSED> + refine_l0f_kernel (no assembled refinement input in HBM) against the same library with
    IODINE_NO_AUX_FUSE=1 (raw fp32 aux channels -> assemble16_kernel -> 16-byte gather kernel): the same reconstruction
    up to the 16-bit rounding of the refinement encoder's first layer (the x-coordinate plane is a rounded operand
    channel in the fused kernel and an fp32 table entry in the other)."""
    arch = A.arch_by_name(name, **over)
    K, L, H = arch.SLOTS, arch.DIM_LATENT, arch.IMG_SIZE
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, H, H, generator=g).to(DEV)
    eps = torch.randn(arch.ITERS + 1, B, K, L, generator=g).to(DEV)
    res = []
    for nofuse in (False, True):
        if nofuse:
            monkeypatch.setenv('IODINE_NO_AUX_FUSE', '1')
        else:
            monkeypatch.delenv('IODINE_NO_AUX_FUSE', raising=False)
        m = seeded_model(arch, 2.0, precision=prec).to(DEV)
        pred, mask, mean = m.reconstruct(x, eps=eps)
        torch.cuda.synchronize()
        res.append((pred.clone(), mask.clone(), m.elbo_per_step(B).clone()))
    tol = 6e-3 if prec == 'bf16' else 1e-3
    assert rel_err(res[0][0], res[1][0]) < tol
    assert rel_err(res[0][1], res[1][1]) < tol
    assert rel_err(res[0][2], res[1][2]) < tol
