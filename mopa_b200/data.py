"""Host <-> device plumbing of a training loop around the 3D branch.

The reference moves a batch to the GPU inside the step (`data_batch['x'][1] = data_batch['x'][1].cuda()`,
mopa/train/train_xmuda_mopa.py:233-242; the int64 coordinates stay on the host and SparseConvNet's InputLayer reads them
there) and reads every loss back with `.item()` for its meters (train_xmuda_mopa.py:364, 399;
mopa/common/utils/metric_logger.py:68-80): synchronous stalls in every step. `DevicePrefetcher` and `LaggedScalar`
keep the same data flow, one step deep:

* DevicePrefetcher copies batch i+1 (pinned host memory -> device, side stream) while batch i is being processed;
* LaggedScalar.push(loss_i) starts the device -> host copy of step i's loss and hands back the value of step i-1, whose
  copy finished long ago: the host never waits for the step it has just launched.

Every step's inputs still cross PCIe and every step's loss still reaches the host; only the waiting is gone.

Device tensors handed out by the prefetcher (or passed through `mark_ready`) carry the event of the copy that produced
them. `scn.InputLayer` / `UNetSCN` then know the coordinates are complete and start the voxel hashing of step i+1 on the
library's geometry stream without waiting for the caller's stream, i.e. while the backward pass of step i is still running
(C ABI: coords_on_device = 2, include/mopa_scn.h).
"""
from collections import deque

import torch


def mark_ready(tensor):
    """Declare a device tensor complete as of now (everything queued on the current stream so far) and immutable from here
    on: records an event and attaches it. For inputs that stay resident on the device across steps."""
    ev = torch.cuda.Event()
    ev.record()
    tensor._mopa_ready = ev
    return tensor


def map_tensors(obj, fn):
    """Apply fn to every tensor inside nested dicts / lists / tuples (MoPA's collate output: a dict with 'x': [coords,
    feats], 'seg_label', 'img', 'img_indices': [ndarray, ...], ... -- mopa/data/collate.py:170-260); everything else is
    passed through untouched."""
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, tuple):
        return tuple(map_tensors(v, fn) for v in obj)
    if isinstance(obj, list):
        return [map_tensors(v, fn) for v in obj]
    return obj


class DevicePrefetcher:
    """Iterate over host batches, handing out device copies `depth` batches ahead.

    batches: iterable of `[coords, feats]` (coords int64 (N, 3|4), feats float32 (N, C)) or of any nesting of dicts /
    lists / tuples holding tensors (a MoPA `data_batch`): every tensor is copied, everything else is passed through.
    Tensors that are not pinned are pinned first (an extra host copy: pin them in the DataLoader, `pin_memory=True`, to
    avoid it). The yielded `[coords_dev, feats_dev]` / `data_batch['x']` goes straight into `UNetSCN.forward`.
    """

    def __init__(self, batches, device=None, depth=2):
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePrefetcher needs a CUDA device: mopa_b200 has no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = max(1, int(depth))
        self._it = iter(batches)
        self._stream = torch.cuda.Stream(device=self.device)
        self._queue = deque()
        self._done = False

    def _pin(self, t):
        return t if t.is_pinned() else t.pin_memory()

    def _issue(self):
        if self._done:
            return
        try:
            batch = next(self._it)
        except StopIteration:
            self._done = True
            return
        with torch.cuda.stream(self._stream):
            out = map_tensors(batch, lambda t: self._pin(t).to(self.device, non_blocking=True))
            ready = torch.cuda.Event()
            ready.record(self._stream)

        def tag(t):
            t._mopa_ready = ready
            return t
        map_tensors(out, tag)
        self._queue.append((out, ready))

    def __iter__(self):
        return self

    def __next__(self):
        while len(self._queue) < self.depth and not self._done:
            self._issue()
        if not self._queue:
            raise StopIteration
        out, ready = self._queue.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)

        def used_on_caller(t):
            t.record_stream(cur)  # allocated on the copy stream, consumed on the caller's
            return t
        map_tensors(out, used_on_caller)
        self._issue()  # keep `depth` copies in flight while the caller works on this batch
        return out


class LaggedScalar:
    """Device scalars read on the host one step late: push(x_i) returns float(x_{i-1}) (None on the first call)."""

    def __init__(self):
        self._slots = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._events = [None, None]
        self._n = 0

    def push(self, value):
        prev = self.last()
        k = self._n & 1
        self._slots[k].copy_(value.detach().reshape(1).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[k] = ev
        self._n += 1
        return prev

    def last(self):
        """The most recently pushed value (waits for its copy)."""
        if self._n == 0:
            return None
        k = (self._n - 1) & 1
        self._events[k].synchronize()
        return float(self._slots[k][0])
