"""Device prefetcher for the evaluation / training loops (``lib/engine/eval.py:24-28``: ``data[0] = data[0].to(device)``
inside the loop, synchronous).  Here the NEXT batch's images are copied host->device on a side stream while the
refinement loop runs on the current batch; the consumer's stream waits on the copy's event, so no host
synchronisation is added.  Masks (ragged, evaluator-side) stay on the host exactly as the reference leaves them.
"""
import torch


class DevicePrefetcher:
    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        batch = list(batch)
        if not self.cuda:
            batch[0] = batch[0].to(self.device)
            return batch, None
        with torch.cuda.stream(self.stream):
            batch[0] = batch[0].to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return batch, ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            batch, ev = nxt
            try:
                nxt = self._stage(next(it))          # enqueue the next copy before handing out the current batch
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                batch[0].record_stream(torch.cuda.current_stream(self.device))
            yield batch
