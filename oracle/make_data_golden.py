"""Generate ``tests/golden/data_clevr.npz``: one synthetic CLEVR-shaped sample (320 x 480 PNG image + colour-coded
mask PNG) and what the UNMODIFIED reference dataset (``lib/data/clevr.py``, loaded through ``oracle.data_ref``)
returns for it.  Run in the build container (needs ``/root/reference``):  python -m oracle.make_data_golden
TEST INFRASTRUCTURE.
"""
import io
import os
import tempfile

import numpy as np
from PIL import Image

from . import data_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'data_clevr.npz')


def synthetic_sample(seed=7, H=320, W=480):
    """smooth random image + 4 coloured rectangles/ellipses on the CLEVR background grey"""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([127 + 120 * np.sin(xx / (11.0 + 3 * c) + c) * np.cos(yy / (7.0 + 2 * c)) for c in range(3)], -1)
    img = np.clip(img + rng.randint(-6, 7, size=((H + 15) // 16, (W + 15) // 16, 3)).repeat(16, 0).repeat(16, 1)[:H, :W], 0, 255).astype(np.uint8)
    mask = np.full((H, W, 3), 64, np.uint8)
    for i, col in enumerate([(255, 0, 0), (0, 255, 0), (0, 0, 255), (10, 200, 120)]):
        cy, cx = rng.randint(90, H - 90), rng.randint(150, W - 150)
        ry, rx = rng.randint(12, 45), rng.randint(12, 45)
        sel = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0 if i % 2 else (abs(yy - cy) <= ry) & (abs(xx - cx) <= rx)
        mask[sel] = col
    return img, mask


def png_bytes(a):
    b = io.BytesIO()
    Image.fromarray(a).save(b, format='PNG')
    return np.frombuffer(b.getvalue(), dtype=np.uint8)


def canonical(masks):
    """order-free form of a mask stack: rows sorted by their bytes"""
    m = np.asarray(masks, dtype=np.uint8)
    order = sorted(range(m.shape[0]), key=lambda i: m[i].tobytes())
    return m[order]


def main():
    img, mask = synthetic_sample()
    ib, mb = png_bytes(img), png_bytes(mask)
    with tempfile.TemporaryDirectory() as root:
        os.makedirs(os.path.join(root, 'images'))
        os.makedirs(os.path.join(root, 'masks'))
        for d, b in (('images', ib), ('masks', mb)):
            with open(os.path.join(root, d, 'a.png'), 'wb') as f:
                f.write(b.tobytes())
        ref = data_ref.load_reference_dataset_module('clevr')
        with data_ref.legacy_numpy():
            x, m = ref.CLEVR(root, 'test')[0]
    np.savez_compressed(OUT, image_png=ib, mask_png=mb, image=np.round(x.numpy() * 255).astype(np.uint8),
                        image_f32_checksum=np.float64(x.double().sum().item()), masks=canonical(m.numpy()))
    print('wrote', OUT, x.shape, m.shape)


if __name__ == '__main__':
    main()
