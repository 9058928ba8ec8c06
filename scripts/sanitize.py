"""compute-sanitizer target: one tiny reconstruct per kernel family (run: compute-sanitizer --tool memcheck python scripts/sanitize.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
from helpers import seeded_model  # noqa: E402
from iodine_b200.config import arch_by_name  # noqa: E402

CASES = [('tiny', {}, 'fp32'), ('tiny', {}, 'fp16'), ('test5x5', dict(iters=1), 'bf16'),
         ('tiny', dict(img_size=128, dec_chan=16, dec_layers=3, slots=2, iters=1), 'fp16'),
         ('tiny', dict(img_size=256, dec_chan=16, dec_layers=2, slots=1, iters=1), 'fp16'),
         # round 2: tf32 row-streaming kernels (scout warp, two issuers, 16 accumulator slots), fused aux path with a
         # 64-channel refinement encoder (refine_l0f_kernel<64>), 9 slots (mixture_fast_kernel<12>)
         ('tiny', dict(img_size=128, dec_chan=32, dec_layers=3, slots=2, iters=1), 'tf32'),
         ('tiny', dict(img_size=128, dec_chan=16, dec_layers=2, slots=9, iters=1, ref_chan=64, ref_layers=3), 'fp16')]
for name, over, prec in CASES:
    arch = arch_by_name(name, **over)
    m = seeded_model(arch, 2.0, precision=prec).to('cuda:0')
    x = torch.rand(1, 3, arch.IMG_SIZE, arch.IMG_SIZE, device='cuda:0')
    pred, mask, mean = m.reconstruct(x)
    torch.cuda.synchronize()
    print(name, over, prec, 'ok', float(pred.mean()))
