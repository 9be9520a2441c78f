"""Training-loop data feed: device-side, asynchronous ``get_batch`` with a prefetch queue.

The reference's training loops call ``sampler.get_batch(batch_size=128, guide=...)`` synchronously every
iteration (``training/train_CDNS.py:128-141``, ``train_NWP.py:128-141``, ``train_CLIP.py:145``): the batch is
sampled and belief-propagated on the CPU, then copied to the GPU.  Here the batch is produced on the GPU by the
Philox sampler and the BP kernels; ``BatchPrefetcher`` issues ``get_batch(..., async_=True)`` for the next
``depth`` batches on a side stream so that they overlap the consumer's forward/backward kernels, and hands a batch
over with a stream-level dependency (``wait_event``) instead of a host synchronisation.
"""
import collections

import torch

__all__ = ["BatchPrefetcher"]


def _tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            yield from _tensors(o)


class BatchPrefetcher:
    """Iterator over ``sampler.get_batch(batch_size=..., guide=..., device=<the sampler's GPU>, async_=True)``.

    ``depth`` batches are in flight on the prefetch stream.  ``next()`` makes the CURRENT stream wait for the
    batch's event and marks its tensors as used on that stream (so the caching allocator does not recycle them
    early); no host-side wait happens anywhere.  The sampler must be in Philox mode (``rng="philox"``): the NumPy
    parity mode draws on the host.  Sampler state (``seed``, ``tree_offset``) advances per issued batch, so a run is
    resumable from ``sampler.tree_offset - depth * trees_per_batch``.
    """

    def __init__(self, sampler, batch_size=128, guide=True, depth=2, **get_batch_kwargs):
        if getattr(sampler, "rng", None) != "philox":
            raise ValueError("BatchPrefetcher needs a sampler constructed with rng='philox'")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.sampler, self.batch_size, self.guide, self.depth = sampler, int(batch_size), bool(guide), int(depth)
        self.kwargs = get_batch_kwargs
        self.device = sampler.device
        self.stream = torch.cuda.Stream(device=self.device)
        self.queue = collections.deque()
        self.issued = 0
        for _ in range(self.depth):
            self._issue()

    def _issue(self):
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            batch = self.sampler.get_batch(batch_size=self.batch_size, guide=self.guide, device=self.device,
                                           async_=True, **self.kwargs)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((batch, ev))
        self.issued += 1

    def __iter__(self):
        return self

    def __next__(self):
        batch, ev = self.queue.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in _tensors(batch):
            if t.is_cuda:
                t.record_stream(cur)
        self._issue()
        return batch
