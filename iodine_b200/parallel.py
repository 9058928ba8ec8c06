"""Slot-shard data parallelism: one process per GPU, each rank owns a contiguous block of WHOLE
images (all K slots of an image stay on one GPU).

Replaces the reference's only parallelism strategy, ``torch.nn.DataParallel``
(``lib/modeling/build.py:11-12``: scatter the batch on dim 0, replicate the parameters, gather
the per-replica losses; ``lib/engine/train.py:61`` then averages them).  Because every K-way
reduction of the refinement loop (mask softmax ``iodine.py:185``, mixture logsumexp ``213-216``,
mask-posterior / leave-one-out sums ``292, 324``) is over the slots of ONE image, whole-image
sharding needs no data-path collective inside a step.  The only cross-rank quantity is the pair
of batch sums behind the ELBO -- ``[sum_b log-likelihood, sum_b KL]`` per refinement step
(``iodine.py:193, 220``: means over the batch) -- and nothing inside the loop consumes the global
value (the backward seed is ``B * elbo``, a per-image sum, ``iodine.py:86-90``).  So the exchange
is ONE all-reduce of the ``[T, 2]`` table per call (NCCL over NVLink on GPUs, gloo in the CPU
tests), issued on the stream the loop ran on.

Equal shards reproduce DataParallel's "mean of replica means" exactly; unequal shards (B not a
multiple of the world size) still give the exact GLOBAL mean here, which is the better-defined
quantity (DataParallel's differs in that case).
"""
import torch
import torch.distributed as dist


def shard_bounds(B, world, rank):
    """Contiguous block of images for ``rank``: the first ``B % world`` ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank %d / world %d' % (rank, world))
    q, r = divmod(int(B), int(world))
    b0 = rank * q + min(rank, r)
    return b0, b0 + q + (1 if rank < r else 0)


def shard_table(B, world):
    return [shard_bounds(B, world, r) for r in range(world)]


class SlotShard:
    """Wraps an ``IODINE``-like model (anything with ``reconstruct(x, eps=)`` / ``encode(x, eps=)``
    that leaves the per-step ``[T,2]`` sums in ``.elbo_terms``) for one-process-per-GPU use.

    ``x`` / ``eps`` passed to the methods are the GLOBAL batch (as DataParallel's callers pass
    it); each rank slices out its own images.  Pass ``local=True`` when the caller already holds
    only its shard (e.g. a DistributedSampler-fed loader).
    """

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.global_batch = None
        self.elbo_terms = None          # [T,2] global sums after the all-reduce

    # ------------------------------------------------------------------ plumbing
    def _local(self, x, eps, local):
        if local or self.world == 1:
            n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
            if self.world > 1:
                dist.all_reduce(n, group=self.group)
            self.global_batch = int(n.item())
            return x, eps
        B = x.shape[0]
        self.global_batch = B
        b0, b1 = shard_bounds(B, self.world, self.rank)
        if b1 <= b0:
            raise ValueError('rank %d of %d has no image: global batch %d is smaller than the '
                             'world size (K-split mode is not implemented)' % (self.rank, self.world, B))
        return x[b0:b1], (None if eps is None else eps[:, b0:b1])

    def _reduce_terms(self):
        t = self.model.elbo_terms
        t = t.clone() if isinstance(t, torch.Tensor) else torch.as_tensor(t)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)     # the path's single exchange: [T,2] sums
        self.elbo_terms = t
        return t

    def _gather(self, t):
        """all-gather along dim 0 (ragged-safe); used only when the caller asks for global outputs."""
        if self.world == 1:
            return t
        sizes = [b1 - b0 for b0, b1 in shard_table(self.global_batch, self.world)]
        if len(set(sizes)) == 1:
            out = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(out, t.contiguous(), group=self.group)
        else:
            out = [t.new_empty((s,) + tuple(t.shape[1:])) for s in sizes]
            for r in range(self.world):      # ragged: one broadcast per owner
                if r == self.rank:
                    out[r].copy_(t)
                dist.broadcast(out[r], src=dist.get_global_rank(self.group, r) if self.group else r,
                               group=self.group)
        return torch.cat(out, dim=0)

    # ------------------------------------------------------------------ API (mirrors IODINE)
    def encode(self, x, eps=None, local=False, gather=False):
        xl, el = self._local(x, eps, local)
        z = self.model.encode(xl, eps=el)
        self._reduce_terms()
        return self._gather(z) if gather else z

    def reconstruct(self, x, eps=None, local=False, gather=False):
        """Returns this rank's (pred, mask, mean) -- or the global ones with ``gather=True``."""
        xl, el = self._local(x, eps, local)
        pred, mask, mean = self.model.reconstruct(xl, eps=el)
        self._reduce_terms()
        if gather:
            return self._gather(pred), self._gather(mask), self._gather(mean)
        return pred, mask, mean

    def elbo_per_step(self):
        """Global ELBO of every refinement step (mean over the GLOBAL batch), identical on all ranks."""
        t = self.elbo_terms
        return (t[:, 0] - t[:, 1]) / self.global_batch
