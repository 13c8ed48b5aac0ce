"""Replay -> device input pipeline.

The reference trains from ``iter(PyTorchReplayBuffer.dataset())`` (YARR yarr/replay_buffer/wrappers/pytorch_replay_buffer.py:
75-82: a DataLoader over the uniform replay buffer, uniform_replay_buffer.py:351-386, ``pin_memory=True``) and moves every
sampled batch to the GPU inside the training loop with blocking ``.to(device)`` calls, so the ~28 MB of RGB-D per batch cross
PCIe while the GPU idles.  ``DeviceFeeder`` wraps ANY such iterator of batch dicts (numpy arrays or CPU tensors, the replay
buffer's own storage format): a background thread pulls the next batch, packs it into reusable pinned staging buffers and
enqueues its host->device copies on a private copy stream; the training step acquires a batch whose copies have (usually)
already finished.  The data, its dtypes and the sampling order are untouched -- this is plumbing, not a different sampler.
"""
import queue
import threading

import numpy as np
import torch


class DeviceFeeder:
    def __init__(self, batches, device, depth=2, keys=None):
        """batches: iterable of dicts; device: torch device; depth: batches in flight; keys: subset of entries to move
        (others are passed through untouched on the host)."""
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        self.depth = max(1, int(depth))
        self.keys = set(keys) if keys is not None else None
        self._it = iter(batches)
        self._q = queue.Queue(maxsize=self.depth)
        self._free = queue.Queue()
        for _ in range(self.depth + 1):
            self._free.put({'pinned': {}, 'dev': {}, 'ready': torch.cuda.Event() if self.cuda else None,
                            'done': torch.cuda.Event() if self.cuda else None, 'used': False})
        self._copy_stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._cur = None
        self._err = None
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()

    @staticmethod
    def _as_tensor(v):
        if isinstance(v, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(v))
        return v

    def _stage(self, slot, batch):
        out = {}
        if self.cuda and slot['used']:
            slot['done'].synchronize()                      # the step that consumed this slot's device buffers has finished
        ctx = torch.cuda.stream(self._copy_stream) if self.cuda else None
        if ctx is not None:
            ctx.__enter__()
        try:
            for k, v in batch.items():
                t = self._as_tensor(v)
                if not torch.is_tensor(t) or (self.keys is not None and k not in self.keys):
                    out[k] = v
                    continue
                if not self.cuda:
                    out[k] = t
                    continue
                pin = slot['pinned'].get(k)
                if pin is None or pin.shape != t.shape or pin.dtype != t.dtype:
                    pin = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    slot['pinned'][k] = pin
                    slot['dev'][k] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                pin.copy_(t)
                slot['dev'][k].copy_(pin, non_blocking=True)
                out[k] = slot['dev'][k]
            if self.cuda:
                slot['ready'].record(self._copy_stream)
        finally:
            if ctx is not None:
                ctx.__exit__(None, None, None)
        slot['used'] = True
        return out

    def _work(self):
        try:
            for batch in self._it:
                slot = self._free.get()
                self._q.put((slot, self._stage(slot, batch)))
        except BaseException as e:          # surfaced in the consumer thread
            self._err = e
        self._q.put(None)

    def __iter__(self):
        return self

    def __next__(self):
        if self._cur is not None:
            # the previous batch's consumers have been enqueued on the compute stream: its buffers may be reused after them
            if self.cuda:
                self._cur['done'].record(torch.cuda.current_stream(self.device))
            self._free.put(self._cur)
            self._cur = None
        item = self._q.get()
        if item is None:
            if self._err is not None:
                raise self._err
            raise StopIteration
        slot, out = item
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_event(slot['ready'])
        self._cur = slot
        return out
