"""The three torchvision transforms the reference composes (lib/data/clevr.py:26-38), written directly on PIL so
that the pipeline needs neither torchvision nor skimage: ``CenterCrop`` -> ``Resize`` -> ``ToTensor``.

Bit-exact with torchvision's PIL code path (tests/test_data.py compares against it): the crop offsets follow
``torchvision.transforms.functional.center_crop`` (``int(round((h - th) / 2.0))``, zero padding when the image is
smaller than the crop), ``Resize(int)`` maps the SHORTER side to the size keeping the aspect ratio with PIL's
antialiased ``BILINEAR`` (or ``NEAREST`` for masks), ``ToTensor`` is ``uint8 HWC -> float32 CHW / 255``.
"""
import numpy as np
import torch
from PIL import Image


def imread(path):
    """``skimage.io.imread`` as the reference uses it (clevr.py:24, dsprite.py:18): the decoded image as a uint8
    ``H x W x C`` (or ``H x W``) array, palette images expanded."""
    with Image.open(path) as im:
        if im.mode == 'P':
            im = im.convert('RGBA' if 'transparency' in im.info else 'RGB')
        elif im.mode not in ('L', 'RGB', 'RGBA'):
            im = im.convert('RGB')
        return np.asarray(im)


def center_crop(img, size, fill=0):
    w, h = img.size
    th = tw = int(size)
    if tw > w or th > h:                         # torchvision pads with zeros first
        pl, pt = max((tw - w) // 2, 0), max((th - h) // 2, 0)
        pr, pb = max((tw - w + 1) // 2, 0), max((th - h + 1) // 2, 0)
        canvas = Image.new(img.mode, (w + pl + pr, h + pt + pb), fill)
        canvas.paste(img, (pl, pt))
        img = canvas
        w, h = img.size
    top = int(round((h - th) / 2.0))
    left = int(round((w - tw) / 2.0))
    return img.crop((left, top, left + tw, top + th))


def resize_shorter(img, size, resample=Image.BILINEAR):
    w, h = img.size
    short, long_ = (w, h) if w <= h else (h, w)
    if short == size:
        return img
    new_short, new_long = size, int(size * long_ / short)
    nw, nh = (new_short, new_long) if w <= h else (new_long, new_short)
    return img.resize((nw, nh), resample)


def to_tensor(img):
    """PIL image or uint8 ``H x W [x C]`` array -> float32 ``C x H x W`` in [0, 1]."""
    a = np.asarray(img)
    if a.ndim == 2:
        a = a[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))
    return t.float().div(255) if t.dtype == torch.uint8 else t.float()
