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


# ------------------------------------------------------------------------------------------------
# Configuration files: the reference's YAML experiment files (configs/*.yaml) read without yacs
# ------------------------------------------------------------------------------------------------
class Cfg(SimpleNamespace):
    """A yacs-``CfgNode``-shaped tree (attribute access, ``merge_from_file``, ``merge_from_list``, ``freeze``) that the
    host mirrors accept wherever the reference passes its ``cfg`` (``make_model``, ``make_dataloader``,
    ``make_evaluator``, ``make_getter``).  Like yacs it refuses keys that the defaults do not declare."""

    def _set(self, path, value, create=False):
        node = self
        for k in path[:-1]:
            if not hasattr(node, k):
                raise KeyError('Non-existent config key: %s' % '.'.join(path))
            node = getattr(node, k)
        if not create and not hasattr(node, path[-1]):
            raise KeyError('Non-existent config key: %s' % '.'.join(path))
        if getattr(self, '_frozen', False):
            raise AttributeError('Attempted to set %s on a frozen config' % '.'.join(path))
        setattr(node, path[-1], value)

    def _merge(self, d, prefix=()):
        for k, v in d.items():
            if isinstance(v, dict):
                if not hasattr(self._at(prefix), k):
                    raise KeyError('Non-existent config key: %s' % '.'.join(prefix + (k,)))
                self._merge(v, prefix + (k,))
            else:
                self._set(prefix + (k,), _decode(v))

    def _at(self, path):
        node = self
        for k in path:
            node = getattr(node, k)
        return node

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        """``['ARCH.ITERS', '3', 'MODEL.DEVICE', 'cpu']`` -- values parsed as Python literals when they are ones"""
        opts = list(opts or [])
        assert len(opts) % 2 == 0, 'override list must be KEY VALUE pairs'
        for key, val in zip(opts[0::2], opts[1::2]):
            self._set(tuple(key.split('.')), _decode(val))

    def freeze(self):
        object.__setattr__(self, '_frozen', True)


def _decode(v):
    """strings that are Python literals become those (yacs does the same: YAML 1.1 reads ``3e-4`` as a string)"""
    if isinstance(v, str):
        import ast
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _tree(d):
    return Cfg(**{k: _tree(v) if isinstance(v, dict) else v for k, v in d.items()})


def default_config():
    """The reference's defaults (``lib/config/defaults.py:4-172``) as far as the IODINE path and its callers read them
    -- values restated, none of that module's import-time side effects (creating ``data/model``, deleting the log
    directory).  Paths are relative to the working directory, like the reference's ``data/...``."""
    return _tree({
        'EXP': {'NAME': 'test'},
        'MODEL': {'NAME': 'VAE', 'DEVICE': 'cpu', 'PARALLEL': False, 'DEVICE_IDS': [], 'PRECISION': 'fp32'},
        'ARCH': {
            'ITERS': 5, 'SLOTS': 7, 'SIGMA': 0.13, 'DIM_LATENT': 128, 'IMG_SIZE': 32, 'IMG_CHANNELS': 3,
            'LAYERNORM': True, 'STOP_GRADIENT': False,
            'ENCODING': ['image', 'means', 'mask', 'mask_logits', 'grad_means', 'grad_mask', 'grad_post', 'posterior',
                         'mask_posterior', 'likelihood', 'leave_one_out_likelihood'],
            'REF': {'CONV_CHAN': 32, 'CONV_LAYERS': 3, 'MLP_UNITS': 256, 'KERNEL_SIZE': 3, 'STRIDE': 2},
            'DEC': {'CONV_CHAN': 64, 'CONV_LAYERS': 5, 'KERNEL_SIZE': 5},
        },
        'DATASET': {'TRAIN': 'MNIST', 'VAL': 'MNIST', 'TEST': 'MNIST'},
        'DATALOADER': {'NUM_WORKERS': 4},
        'TRAIN': {'RESUME': False, 'MAX_EPOCHS': 30, 'BATCH_SIZE': 128, 'BASE_LR': 0.001, 'WEIGHT_DECAY': 0.0005,
                  'CHECKPOINT_PERIOD': 2500, 'NUM_CHECKPOINTS': 3, 'PRINT_EVERY': 100, 'VAL_EVERY': 1000},
        'VAL': {'IS_ON': False, 'BATCH_SIZE': 1, 'EVALUATOR': ''},
        'TEST': {'BATCH_SIZE': 32, 'EVALUATOR': ''},
        'TENSORBOARD': {'IS_ON': True, 'TARGETS': {'SCALAR': ['loss'], 'IMAGE': ['image', 'pred']}, 'LOG_DIR': 'logs'},
        'MODEL_DIR': 'data/model',
        'GETTER': 'VAE',
    })


def load_config(config_file='', opts=()):
    """What ``lib/config/parse.py:3-28`` does with ``--config-file`` and the trailing ``KEY VALUE`` overrides:
    defaults <- YAML file <- override list, then frozen.  ``MODEL.PRECISION`` (``fp32|fp16|bf16``) is this
    implementation's one additional key."""
    cfg = default_config()
    if config_file:
        cfg.merge_from_file(config_file)
    cfg.merge_from_list(opts)
    cfg.freeze()
    return cfg
