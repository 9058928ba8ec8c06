"""Input pipeline of the refinement loop: the reference's ``lib/data`` datasets and
``make_dataloader`` with the same directory layout, transforms and return types, plus a device prefetcher that
overlaps the host->device copy of the next batch with the loop running on the current one."""
from .build import collate_fn, make_dataloader, make_dataset  # noqa: F401
from .clevr import CLEVR  # noqa: F401
from .dsprites import MultiDSprites  # noqa: F401
from .prefetch import DevicePrefetcher  # noqa: F401
