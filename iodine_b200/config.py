"""Architecture descriptions: the ``cfg.ARCH`` fields the model reads (reference ``lib/config/defaults.py:35-100``,
``lib/modeling/iodine.py:10-32``) as plain ``SimpleNamespace`` objects, for callers without a yacs config
(``bench.py``, scripts).  ``make_model(cfg)`` accepts a yacs node or anything with the same attributes.
"""
from types import SimpleNamespace

ALL_ENCODINGS = [
    'posterior', 'grad_post', 'image', 'means', 'mask', 'mask_logits', 'mask_posterior',
    'grad_means', 'grad_mask', 'likelihood', 'leave_one_out_likelihood', 'coordinate',
]


def make_arch(iters, slots, dim_latent, img_size, ref_chan, ref_layers, mlp_units,
              dec_chan, dec_layers, ref_k=3, dec_k=3, sigma=0.10, ref_stride=2,
              layernorm=True):
    return SimpleNamespace(
        ITERS=iters, SLOTS=slots, SIGMA=sigma, DIM_LATENT=dim_latent, IMG_SIZE=img_size,
        IMG_CHANNELS=3, LAYERNORM=layernorm, STOP_GRADIENT=False,
        ENCODING=list(ALL_ENCODINGS),
        REF=SimpleNamespace(CONV_CHAN=ref_chan, CONV_LAYERS=ref_layers, MLP_UNITS=mlp_units,
                            KERNEL_SIZE=ref_k, STRIDE=ref_stride),
        DEC=SimpleNamespace(CONV_CHAN=dec_chan, CONV_LAYERS=dec_layers, KERNEL_SIZE=dec_k),
    )


def arch_by_name(name, **over):
    """Named architectures.

    ``dsprites``  = configs/dsprites_noclip.yaml:26-45 (BASELINE config #1 uses ITERS=3)
    ``clevr6``    = configs/clevr6_prop.yaml:26-45     (BASELINE configs #2-#4)
    ``test5x5``   = configs/test.yaml flavour: 5x5 kernels (small sizes for tests)
    ``tiny``      = a miniature with every code path (odd sizes on purpose)
    """
    if name == 'dsprites':
        kw = dict(iters=3, slots=6, dim_latent=16, img_size=64, ref_chan=32, ref_layers=3,
                  mlp_units=128, dec_chan=32, dec_layers=5)
    elif name == 'clevr6':
        kw = dict(iters=5, slots=7, dim_latent=64, img_size=128, ref_chan=64, ref_layers=4,
                  mlp_units=256, dec_chan=64, dec_layers=4)
    elif name == 'test5x5':
        kw = dict(iters=2, slots=3, dim_latent=16, img_size=32, ref_chan=32, ref_layers=3,
                  mlp_units=64, dec_chan=32, dec_layers=3, ref_k=5, dec_k=5)
    elif name == 'tiny':
        kw = dict(iters=3, slots=3, dim_latent=8, img_size=16, ref_chan=16, ref_layers=2,
                  mlp_units=32, dec_chan=16, dec_layers=2)
    else:
        raise KeyError(name)
    kw.update(over)
    return make_arch(**kw)
