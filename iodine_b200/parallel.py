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
import ctypes as C

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


class NcclComm:
    """An ``ncclComm_t`` of this process for ``iodine_plan_set_comm`` (include/iodine_b200.h): the library then sums
    the ``[T,2]`` ELBO table with one ``ncclAllReduce`` on the stream of the call -- no host round trip, and no torch
    on the data path.  torch does not expose the communicators behind ``torch.distributed``, so this creates one
    with the same NCCL library the process already loaded (torch's ``libnccl.so.2``): rank 0 draws the unique id,
    ``torch.distributed`` (any backend) carries its 128 bytes to the other ranks, every rank joins with
    ``ncclCommInitRank`` on its current CUDA device.  A C/C++ host does the same with its own bootstrap.
    """

    class _UniqueId(C.Structure):                  # ncclUniqueId: 128 opaque bytes (nccl.h)
        _fields_ = [('internal', C.c_byte * 128)]

    def __init__(self, rank=None, world=None, group=None):
        distributed = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank(group) if distributed else 0) if rank is None else int(rank)
        self.world = (dist.get_world_size(group) if distributed else 1) if world is None else int(world)
        if not torch.cuda.is_available():
            raise RuntimeError('NcclComm needs a CUDA device (NCCL communicators are per GPU)')
        torch.cuda.current_device()                # make sure the CUDA context of this rank's device exists
        try:
            self._nccl = C.CDLL('libnccl.so.2')
        except OSError as e:
            raise RuntimeError('cannot load libnccl.so.2: %s' % e)
        self._nccl.ncclGetErrorString.restype = C.c_char_p
        uid = NcclComm._UniqueId()
        if self.rank == 0:
            self._ok(self._nccl.ncclGetUniqueId(C.byref(uid)), 'ncclGetUniqueId')
        if self.world > 1:
            if not distributed:
                raise RuntimeError('NcclComm with world > 1 needs torch.distributed to carry the unique id')
            box = [bytes(uid.internal) if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group else 0, group=group)
            C.memmove(C.byref(uid), box[0], 128)
        comm = C.c_void_p()
        self._nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclComm._UniqueId, C.c_int]
        self._ok(self._nccl.ncclCommInitRank(C.byref(comm), self.world, uid, self.rank), 'ncclCommInitRank')
        self.handle = comm.value

    def _ok(self, rc, what):
        if rc != 0:
            raise RuntimeError('%s failed: %s' % (what, self._nccl.ncclGetErrorString(rc).decode()))

    def close(self):
        if getattr(self, 'handle', None):
            self._nccl.ncclCommDestroy.argtypes = [C.c_void_p]
            self._nccl.ncclCommDestroy(C.c_void_p(self.handle))
            self.handle = None


class SlotShard:
    """Wraps an ``IODINE``-like model (anything with ``reconstruct(x, eps=)`` / ``encode(x, eps=)``
    that leaves the per-step ``[T,2]`` sums in ``.elbo_terms``) for one-process-per-GPU use.

    ``x`` / ``eps`` passed to the methods are the GLOBAL batch (as DataParallel's callers pass
    it); each rank slices out its own images.  Pass ``local=True`` when the caller already holds
    only its shard (e.g. a DistributedSampler-fed loader).
    """

    def __init__(self, model, group=None, native_comm='auto'):
        """``native_comm``: let the library all-reduce the ``[T,2]`` table itself on the stream of the call
        (``iodine_plan_set_comm`` + ``NcclComm``) instead of a ``torch.distributed`` all-reduce afterwards.
        ``'auto'`` = when the process group runs on NCCL and the model offers ``set_comm``."""
        self.model = model
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.global_batch = None
        self.elbo_terms = None          # [T,2] global sums after the all-reduce
        if native_comm == 'auto':
            native_comm = (self.world > 1 and hasattr(model, 'set_comm')
                           and 'nccl' in str(dist.get_backend(group)).lower())
        self.native = bool(native_comm) and self.world > 1
        self._comm = None
        self._calls_checked = set()
        self._local_mode = False
        if self.native:
            self._comm = NcclComm(group=group)
            model.set_comm(self._comm, self.rank, self.world)

    # ------------------------------------------------------------------ plumbing
    def _local(self, x, eps, local):
        self._local_mode = bool(local)
        if local or self.world == 1:
            n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
            if self.world > 1:
                dist.all_reduce(n, group=self.group)
            self._set_global_batch(int(n.item()))
            return x, eps
        B = x.shape[0]
        self._set_global_batch(B)
        b0, b1 = shard_bounds(B, self.world, self.rank)
        if b1 <= b0:
            raise ValueError('rank %d of %d has no image: global batch %d is smaller than the world size -- '
                             'use iodine_b200.parallel.KSplit (ranks own slots instead of images)'
                             % (self.rank, self.world, B))
        return x[b0:b1], (None if eps is None else eps[:, b0:b1])

    def _set_global_batch(self, n):
        self.global_batch = n
        if self.native:
            # the engine now returns sums over all ranks' images: the model's means (elbo_per_step, the logged
            # kl / likelihood, forward()'s loss) are over the global batch
            self.model.global_batch = n

    def _check_calls(self, n_local):
        """native mode: the library all-reduces once per engine call, so every rank must make the same number of
        calls (the model chunks a shard that does not fit HBM).  The decision to run this check is keyed on the
        GLOBAL batch size, which is identical on every rank -- never on the local shard size, which is not when the
        shards are ragged (a rank-dependent skip would leave some ranks inside this all-reduce and the others
        inside the library's).  The chunk size itself is agreed once (minimum over ranks) and then frozen in
        ``model.max_images_per_call`` so that later calls cannot drift apart with each rank's free memory."""
        if not self.native or not hasattr(self.model, '_spans'):
            return
        key = int(self.global_batch)
        if key in self._calls_checked:
            return
        dev = self.model._device()
        if not getattr(self.model, 'max_images_per_call', None):
            biggest = max(b1 - b0 for b0, b1 in shard_table(self.global_batch, self.world)) if not self._local_mode \
                else n_local
            c = torch.tensor([self.model._chunk(biggest)], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.MIN, group=self.group)
            self.model.max_images_per_call = int(c.item())
        n = len(self.model._spans(n_local))
        t = torch.tensor([n, -n], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        if int(t[0]) != -int(t[1]):
            raise RuntimeError('ranks would split their shards into different numbers of engine calls (%d..%d, chunk '
                               '%s images); choose model.max_images_per_call so that they agree, or use '
                               'native_comm=False' % (-int(t[1]), int(t[0]), self.model.max_images_per_call))
        self._calls_checked.add(key)

    def _reduce_terms(self):
        t = self.model.elbo_terms
        t = t.clone() if isinstance(t, torch.Tensor) else torch.as_tensor(t)
        if self.world > 1 and not self.native:
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
        self._check_calls(xl.shape[0])
        z = self.model.encode(xl, eps=el)
        self._reduce_terms()
        return self._gather(z) if gather else z

    def reconstruct(self, x, eps=None, local=False, gather=False):
        """Returns this rank's (pred, mask, mean) -- or the global ones with ``gather=True``."""
        xl, el = self._local(x, eps, local)
        self._check_calls(xl.shape[0])
        pred, mask, mean = self.model.reconstruct(xl, eps=el)
        self._reduce_terms()
        if gather:
            return self._gather(pred), self._gather(mask), self._gather(mean)
        return pred, mask, mean

    def elbo_per_step(self):
        """Global ELBO of every refinement step (mean over the GLOBAL batch), identical on all ranks."""
        t = self.elbo_terms
        return (t[:, 0] - t[:, 1]) / self.global_batch


class KSplit:
    """K-split: the fallback partition for a batch SMALLER than the number of GPUs (e.g. one image
    with K = 16 on 2-8 GPUs), again replacing ``torch.nn.DataParallel`` (``lib/modeling/build.py:11-12``), which can
    only scatter the batch axis.  Every rank holds ALL images and ``K / world`` of their slots: decoder, data-gradient,
    refinement network and posterior state are per slot, so they split; the K-way reductions of ``IODINE.elbo``
    (``iodine.py:185, 213-216, 292, 324``) need every slot's decoder output, which the library all-gathers once per
    elbo() evaluation (``ncclAllGather`` of ``[B,K,H,W,4]`` fp32 on the stream of the call, ``iodine_plan_set_comm`` on
    a K-split plan).  ``pred / mask / mean`` come back complete on every rank, ``z`` as this rank's slots (or all of
    them with ``gather=True``); the ``[T,2]`` ELBO table is all-reduced by the library as in ``SlotShard``.

    NCCL only (the exchange sits between two kernels of a step): needs a CUDA model and an initialised
    ``torch.distributed`` NCCL group to carry the communicator's unique id.
    """

    def __init__(self, model, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('KSplit needs torch.distributed (one process per GPU)')
        self.model, self.group = model, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if model.K % self.world:
            raise ValueError('K-split: SLOTS=%d is not a multiple of the world size %d' % (model.K, self.world))
        self.k_local = model.K // self.world
        self.k0 = self.rank * self.k_local
        self._comm = NcclComm(group=group)
        model.set_slot_split(self._comm, self.rank, self.world)
        self.elbo_terms = None
        self.global_batch = None

    def slots(self):
        """this rank's slot range [k0, k1)"""
        return self.k0, self.k0 + self.k_local

    def _eps(self, eps):
        return None if eps is None else eps[:, :, self.k0:self.k0 + self.k_local].contiguous()

    def _gather_slots(self, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return torch.cat(out, dim=1)

    def _finish(self, B):
        self.global_batch = B
        self.elbo_terms = self.model.elbo_terms.clone()            # already summed over the ranks by the library

    def encode(self, x, eps=None, gather=False):
        """x[B,3,H,W] (the same on every rank); eps[T+1,B,K,L] for ALL slots (each rank slices its own)."""
        z = self.model.encode(x, eps=self._eps(eps))
        self._finish(x.shape[0])
        return self._gather_slots(z) if gather else z

    def reconstruct(self, x, eps=None):
        pred, mask, mean = self.model.reconstruct(x, eps=self._eps(eps))
        self._finish(x.shape[0])
        return pred, mask, mean

    def elbo_per_step(self):
        t = self.elbo_terms
        return (t[:, 0] - t[:, 1]) / self.global_batch
