"""Host <-> device pipelining around the per-click forward.

`HostPipeline` overlaps, on three CUDA streams, the host->device copy of batch i+1, the forward of batch i and the
device->host copy of the result of batch i-1.  Inputs are pinned host tensors, results land in pinned host buffers
owned by the pipeline; `submit` returns a ticket whose `.result()` blocks until that batch's logits are on the host.
No arithmetic happens here: the forward is the module call, the copies are cudaMemcpyAsync.
"""
import collections

import torch


class _Ticket:
    def __init__(self, event, out_host):
        self._event, self._out = event, out_host

    def result(self):
        self._event.synchronize()
        return self._out


class HostPipeline:
    def __init__(self, model, device, depth=3, output="instances"):
        self.model, self.device, self.depth, self.output = model, device, depth, output
        self.h2d, self.compute, self.d2h = (torch.cuda.Stream(device) for _ in range(3))
        self._slots = collections.deque()           # in-flight (ticket, keep-alive tensors)
        self._host_out = {}                         # (slot index, shape) -> pinned buffer
        self._n = 0

    def _out_buffer(self, shape, dtype):
        key = (self._n % self.depth, tuple(shape), dtype)
        if key not in self._host_out:
            self._host_out[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return self._host_out[key]

    def submit(self, image_host, points_host, prompts=None, as_prompt_type=0, prev_mask_host=None):
        """image_host / points_host: pinned CPU tensors.  Returns a ticket; at most `depth` batches are in flight.

        image_host is either the network operand itself, fp32 [B,4,H,W] (RGB in [0,1] + previous mask), or the decoded images as
        they come from a dataset, uint8 [B,H,W,3], with the previous masks in `prev_mask_host` (fp32 [B,H,W] / [B,1,H,W], None =
        zeros): then 3 + 4 bytes per pixel cross PCIe instead of 16 and the predictor's ToTensor (x / 255) runs on the device
        (ops.image_from_u8, bit-identical with the host division)."""
        while len(self._slots) >= self.depth:
            self._slots.popleft()[0].result()
        u8 = image_host.dtype == torch.uint8
        with torch.cuda.stream(self.h2d):
            img = image_host.to(self.device, non_blocking=True)
            prev = prev_mask_host.to(self.device, non_blocking=True) if (u8 and prev_mask_host is not None) else None
            pts = points_host.to(self.device, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.h2d)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(copied)
            if u8:
                from . import ops
                raw = (img, prev)
                img = ops.image_from_u8(img, prev)
                for t in raw:
                    if t is not None:
                        t.record_stream(self.compute)
            out = self.model(img, pts, prompts, as_prompt_type)[self.output]
            done = torch.cuda.Event()
            done.record(self.compute)
            img.record_stream(self.compute)
            pts.record_stream(self.compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(done)
            host = self._out_buffer(out.shape, out.dtype)
            host.copy_(out, non_blocking=True)
            out.record_stream(self.d2h)
            landed = torch.cuda.Event()
            landed.record(self.d2h)
        self._n += 1
        t = _Ticket(landed, host)
        self._slots.append((t, (img, pts, out)))
        return t

    def drain(self):
        while self._slots:
            self._slots.popleft()[0].result()
