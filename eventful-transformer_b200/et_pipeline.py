"""
Host-to-host streaming of a video through an Eventful backbone with copies overlapped on a second CUDA stream.

    pipe = FramePipeline(model, frame_shape, dtype, device)
    for host_frame in video:                  # pinned host tensors
        host_out = pipe.step(host_frame)      # pinned host tensor holding the PREVIOUS frame's output (None at first)
    last = pipe.flush()

Frame t+1 is uploaded while frame t is computed, and frame t's feature map is downloaded while frame t+1 is
computed (double-buffered staging on both sides), so the per-frame cost is max(compute, copies) instead of their sum.
The model call itself is the public `ViTBackbone.forward` (CUDA-graph replay when enabled).
"""

import torch


class FramePipeline:
    def __init__(self, model, frame_shape, dtype, device):
        self.model = model
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.stage = [torch.empty(frame_shape, dtype=dtype, device=device) for _ in range(2)]
        self.out_host = [torch.empty(frame_shape, dtype=dtype).pin_memory() for _ in range(2)]
        self.in_ready = [torch.cuda.Event() for _ in range(2)]
        self.in_free = [torch.cuda.Event() for _ in range(2)]
        self.out_ready = [torch.cuda.Event() for _ in range(2)]
        self.out_done = [torch.cuda.Event() for _ in range(2)]
        self.t = 0
        self._pending = None  # (slot, device output) of the frame whose download has been queued
        self._uploaded = False

    def _upload(self, host_frame, slot):
        with torch.cuda.stream(self.copy_stream):
            if self.t >= 2:
                self.copy_stream.wait_event(self.in_free[slot])  # compute has consumed this staging buffer
            self.stage[slot].copy_(host_frame, non_blocking=True)
            self.in_ready[slot].record(self.copy_stream)

    def step(self, host_frame, next_host_frame=None):
        """Computes `host_frame`; optionally starts uploading `next_host_frame`. Returns the previous output (host)."""
        main = torch.cuda.current_stream(self.device)
        slot = self.t & 1
        if not self._uploaded:
            self._upload(host_frame, slot)
        main.wait_event(self.in_ready[slot])
        out = self.model(self.stage[slot])
        self.in_free[slot].record(main)
        self.out_ready[slot].record(main)
        self.t += 1
        self._uploaded = False
        if next_host_frame is not None:  # overlap the next upload with this frame's compute
            self._upload(next_host_frame, self.t & 1)
            self._uploaded = True
        previous = None
        if self._pending is not None:
            pslot, _ = self._pending
            self.out_done[pslot].synchronize()  # previous frame's feature map is on the host
            previous = self.out_host[pslot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.out_ready[slot])
            self.out_host[slot].copy_(out, non_blocking=True)
            self.out_done[slot].record(self.copy_stream)
        out.record_stream(self.copy_stream)
        self._pending = (slot, out)
        return previous

    def flush(self):
        if self._pending is None:
            return None
        pslot, _ = self._pending
        self.out_done[pslot].synchronize()
        self._pending = None
        return self.out_host[pslot]
