"""``make_dataloader`` / ``make_dataset`` / ``collate_fn`` of the reference (lib/data/build.py:6-45): same config
keys (``cfg.{TRAIN,VAL,TEST}.BATCH_SIZE``, ``cfg.DATALOADER.NUM_WORKERS``, ``cfg.DATASET.TRAIN``), same dataset
roots (``data/CLEVR``, ``data/DSPRITES``), same batch structure ``(images [B,3,H,W], tuple of per-image masks)``.

Added for the GPU loop: batches are collated into PINNED memory when CUDA is present (so that
``DevicePrefetcher`` / ``.to(device, non_blocking=True)`` overlap the copy with the kernels) and the workers stay
alive between epochs.  ``MNIST`` belongs to the reference's VAE model, not to the IODINE path, and is not provided.
"""
import torch
from torch.utils.data import DataLoader

from .clevr import CLEVR
from .dsprites import MultiDSprites

ROOTS = {'CLEVR': 'data/CLEVR', 'DSPRITES': 'data/DSPRITES'}


def collate_fn(batch):
    """list of ``(image, mask)`` -> ``(images stacked on dim 0, tuple of masks)`` -- masks stay a per-image sequence
    because their object count differs (build.py:26-37)."""
    data, mask = zip(*batch)
    return torch.stack(data, dim=0), mask


def make_dataset(cfg, mode, root=None):
    name = cfg.DATASET.TRAIN
    if name == 'CLEVR':
        return CLEVR(root or ROOTS[name], mode)
    if name == 'DSPRITES':
        return MultiDSprites(root or ROOTS[name], mode)
    raise ValueError('dataset %r is not part of the IODINE path (CLEVR, DSPRITES)' % (name,))


def make_dataloader(cfg, mode, root=None):
    if mode == 'train':
        batch_size, shuffle = cfg.TRAIN.BATCH_SIZE, True
    elif mode == 'val':
        batch_size, shuffle = cfg.VAL.BATCH_SIZE, False
    elif mode == 'test':
        batch_size, shuffle = cfg.TEST.BATCH_SIZE, False
    else:
        raise ValueError("mode must be 'train', 'val' or 'test', got %r" % (mode,))
    workers = int(cfg.DATALOADER.NUM_WORKERS)
    return DataLoader(make_dataset(cfg, mode, root), batch_size=batch_size, collate_fn=collate_fn, shuffle=shuffle,
                      num_workers=workers, pin_memory=torch.cuda.is_available(),
                      persistent_workers=workers > 0)
