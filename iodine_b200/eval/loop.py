"""``evaluate`` -- the reference's test engine (``lib/engine/eval.py:5-34``): ``evaluator.reset()``, ``model.eval()``,
one ``evaluator.evaluate(model, data)`` per batch with the images moved to the device, results printed at the end.

Same call and same printed line; the images of batch i+1 are copied host->device on a side stream while the
refinement loop runs on batch i (``iodine_b200.data.DevicePrefetcher``) instead of a synchronous ``.to(device)``
inside the loop, and the results string is also returned (the reference leaves its ``return`` commented out).
"""
import torch

from ..data.prefetch import DevicePrefetcher


def evaluate(model, device, dataloader, evaluator, progress=False):
    model = model.to(device)
    if isinstance(model, torch.nn.DataParallel):        # eval.py:15-16: evaluate the wrapped module
        model = model.module
    evaluator.reset()                                   # "Very important!" (eval.py:19-20)
    model.eval()
    batches = DevicePrefetcher(dataloader, device)
    if progress:
        from tqdm import tqdm
        batches = tqdm(batches, total=len(dataloader))
    for data in batches:
        evaluator.evaluate(model, data)
        if progress:
            batches.set_description(evaluator.get_results())
    results = evaluator.get_results()
    print('Final: ', results)
    return results
