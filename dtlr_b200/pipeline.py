"""Host <-> device streaming around the model: the user-facing way to push batches of line images that live in (pinned)
host memory through DINO.forward + the fused CTC-view decode and get the decoded frame ids back on the host.

Three CUDA streams: uploads (H2D of batch i+1), compute (forward + decode of batch i, the caller's current stream) and
downloads (D2H of the frames of batch i-1) overlap; a 2-slot ring of device input buffers and pinned result buffers
bounds memory.  Every batch still pays its own H2D and D2H -- they are just off the critical path.
"""
import torch

from . import dino


class HostPipeline:
    def __init__(self, model, device=None, eps=0.003, slots=2, copy=True):
        """copy=True (default): every yielded tensor is the caller's own.  copy=False yields the internal pinned result buffer of
        the slot, which is OVERWRITTEN `slots` batches later -- only for callers that consume each result before pulling the next."""
        self.model = model
        self.copy = copy
        self.device = device or next(model.parameters()).device
        self.eps = eps
        self.slots = slots
        self.up = torch.cuda.Stream(device=self.device)
        self.down = torch.cuda.Stream(device=self.device)
        self._dev_in = [None] * slots
        self._host_out = [None] * slots
        self._consumed = [None] * slots       # compute-stream event: the forward that read this input slot has been enqueued

    @torch.no_grad()
    def run(self, host_batches):
        """host_batches: iterable of pinned fp32 CPU tensors (B,3,H,W).  Yields one int32 CPU tensor (B,Q) of frame labels
        per batch (0 = blank, c+1 = class c), in order."""
        compute = torch.cuda.current_stream(self.device)
        pending = []          # (slot, download-done event)
        for i, hb in enumerate(host_batches):
            slot = i % self.slots
            if len(pending) >= self.slots:                      # the slot's previous result must have left the device
                s, ev = pending.pop(0)
                ev.synchronize()
                yield self._host_out[s].clone() if self.copy else self._host_out[s]
            if self._dev_in[slot] is None or self._dev_in[slot].shape != hb.shape:
                self._dev_in[slot] = torch.empty(hb.shape, dtype=hb.dtype, device=self.device)
                self._consumed[slot] = None                     # recycled allocator block: order after all compute work
            if self._consumed[slot] is not None:                # only the forward that last read THIS slot (batch i - slots):
                self.up.wait_event(self._consumed[slot])        # the upload of batch i overlaps the forward of batch i-1
            else:
                self.up.wait_stream(compute)                    # first use: order after the allocation / caller's work
            with torch.cuda.stream(self.up):
                self._dev_in[slot].copy_(hb, non_blocking=True)
            compute.wait_stream(self.up)
            out = self.model(self._dev_in[slot])
            self._consumed[slot] = torch.cuda.Event()
            self._consumed[slot].record(compute)
            frames = dino.decode_frames(out, self.eps)
            if self._host_out[slot] is None or self._host_out[slot].shape != frames.shape:
                self._host_out[slot] = torch.empty(frames.shape, dtype=frames.dtype).pin_memory()
            self.down.wait_stream(compute)
            with torch.cuda.stream(self.down):
                self._host_out[slot].copy_(frames, non_blocking=True)
                frames.record_stream(self.down)
                ev = torch.cuda.Event()
                ev.record(self.down)
            pending.append((slot, ev))
        for s, ev in pending:
            ev.synchronize()
            yield self._host_out[s].clone() if self.copy else self._host_out[s]
