"""Host -> device input feed for the training step: the overlapped counterpart of
``move_data_to_device`` (/root/reference/pytorch/pytorch_utils.py:6-15, called per key at
/root/reference/pytorch/main.py:238-239).

The reference converts every batch to a pageable fp32 tensor and copies it synchronously on the
compute stream (655 MB per step at batch_size 256 + mixup: longer than many of the kernels it
feeds).  Here the batch for step i+1 is copied from PINNED host memory on a dedicated copy stream
while step i computes; CUDA events order (a) the step after its copy and (b) the reuse of a device
slot after the step that read it.  Two device slots => one batch in flight.
"""
import torch


class DeviceFeed(object):
    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * slots
        self.ready = [None] * slots          # copy finished   (recorded on the copy stream)
        self.free = [None] * slots           # consumer finished (recorded on the compute stream)

    def submit(self, slot, host_tensors):
        """Start copying a dict of pinned host tensors into device slot ``slot``."""
        with torch.cuda.device(self.device):
            if self.free[slot] is not None:
                self.stream.wait_event(self.free[slot])
            with torch.cuda.stream(self.stream):
                dst = self.slots[slot]
                if dst is None or any(dst[k].shape != v.shape or dst[k].dtype != v.dtype for k, v in host_tensors.items()):
                    dst = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_tensors.items()}
                    self.slots[slot] = dst
                for k, v in host_tensors.items():
                    dst[k].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.ready[slot] = ev

    def wait_copied(self, slot):
        """Block the HOST until the last copy submitted into ``slot`` has finished reading its pinned source (call
        before overwriting that host buffer; data_generator.TrainBatcher(before_refill=feed.wait_copied))."""
        if slot < len(self.ready) and self.ready[slot] is not None:
            self.ready[slot].synchronize()

    def acquire(self, slot):
        """Device tensors of ``slot``; the current stream waits for the copy."""
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        return self.slots[slot]

    def release(self, slot):
        """Call after enqueueing the work that reads ``slot``."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev
