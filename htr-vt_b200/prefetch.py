"""Host -> device input staging for the drop-in model: double-buffered copies on a side stream.

The reference moves every batch right before it uses it (`image.cuda()` in model_v1/train.py:100-104, valid.py:26-29):
16.8 MB of fp32 line images per 128-line training batch, 67 MB per 512-line inference batch, serialised with the step on
one stream.  `HostPrefetcher.put(...)` enqueues the copies of the NEXT batch on its own stream while the current step's
kernels run; `get()` makes the compute stream wait for them.  Every batch is still copied once, from (preferably pinned)
host memory; only the copy engine and the SMs now work at the same time.

    pf = HostPrefetcher(device)
    pf.put(image, text, length)                 # first batch
    for next_batch in loader:
        image, text, length = pf.get()          # device tensors of the batch put last
        pf.put(*next_batch)                      # its copies run under this step
        loss = compute_loss(image, text, length) ...
"""
import torch


class HostPrefetcher(object):
    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("HostPrefetcher stages batches for a CUDA device")
        self.stream = torch.cuda.Stream(self.device)
        self._slot = 0
        self._buffers = [None, None]          # device tensors of the two slots (reused while shapes / dtypes match)
        self._released = [None, None]         # event on the compute stream: the slot's previous batch is no longer needed
        self._pending = None

    def put(self, *tensors):
        """Enqueue the H2D copies of one batch (host tensors; device tensors pass through) on the side stream."""
        if self._pending is not None:
            raise RuntimeError("HostPrefetcher.put: the previous batch was not taken with get()")
        slot = self._slot
        self._slot ^= 1
        bufs = self._buffers[slot]
        if bufs is None or len(bufs) != len(tensors) or any(
                b is None or b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, tensors)):
            bufs = [None if t.is_cuda else torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in tensors]
            self._buffers[slot] = bufs
        with torch.cuda.stream(self.stream):
            if self._released[slot] is not None:
                self.stream.wait_event(self._released[slot])          # the compute stream is done with this slot's tensors
            outs = []
            for b, t in zip(bufs, tensors):
                if t.is_cuda:
                    outs.append(t)
                else:
                    b.copy_(t, non_blocking=True)
                    outs.append(b)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self._pending = (outs, ready, slot)

    def get(self):
        """Device tensors of the batch passed to the last put(); the current stream waits for their copies."""
        if self._pending is None:
            raise RuntimeError("HostPrefetcher.get: nothing was put")
        outs, ready, slot = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        # whatever the compute stream had enqueued BEFORE this point used the other slot: it may be refilled once the
        # stream gets here
        other = slot ^ 1
        ev = torch.cuda.Event()
        ev.record(cur)
        self._released[other] = ev
        return outs if len(outs) != 1 else outs[0]


class RunningLoss(object):
    """`train_loss += loss.item()` (model_v1/train.py:129) without the device synchronisation of every step.

    `.item()` stalls the host until the step has finished; the next step's ~200 launches are then enqueued into an empty
    queue, and wherever kernels are shorter than their enqueue the GPU waits for the host (0.5-0.7 ms of a 17 ms step,
    tools/e2e_probe.py).  `add(loss)` copies the scalar into a pinned ring on the current stream and returns at once;
    every value is still read back - when its copy has landed (harvested by later add() calls) or, at the latest, by
    `total()` / `mean()`, which wait for what is outstanding (train.py:133 reads the sum every print_iter steps)."""

    def __init__(self, slots=64):
        self._buf = torch.zeros(slots, dtype=torch.float32, pin_memory=True)
        self._pending = []                   # (slot, event), oldest first
        self._sum, self._count, self._next = 0.0, 0, 0

    def _harvest(self, wait_all=False, make_room=False):
        while self._pending:
            slot, ev = self._pending[0]
            if not (wait_all or (make_room and len(self._pending) >= self._buf.numel()) or ev.query()):
                break
            ev.synchronize()
            self._sum += float(self._buf[slot])
            self._pending.pop(0)

    def add(self, loss):
        self._harvest(make_room=True)
        slot = self._next
        self._next = (slot + 1) % self._buf.numel()
        self._buf[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(loss.device))
        self._pending.append((slot, ev))
        self._count += 1

    def total(self):
        self._harvest(wait_all=True)
        return self._sum

    @property
    def count(self):
        return self._count

    def mean(self):
        return self.total() / self._count if self._count else 0.0

    def reset(self):
        self._harvest(wait_all=True)
        self._sum, self._count = 0.0, 0
