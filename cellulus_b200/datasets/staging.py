"""Host-side staging of the pair lists on their way to the device (`cellulus/train.py:162-166`).

The reference's DataLoader delivers the anchor / reference lists as int64 `(B, P, D)` host tensors -- 180 MB per
step at BASELINE configs[1], 85 % of what the loss step moves, and a PCIe copy of 3.5 ms against 46 us of kernels.
Coordinates are pixel indices: below 32768 pixels per axis they fit int16, which the loss kernels read directly.
`PairListStager` converts the lists on the HOST (a dtype-converting copy over all host threads) into pinned int16
staging, so that a quarter of the bytes cross PCIe; two slots let the conversion of step i + 1 run while step i's
copy and kernels are in flight.  Plumbing (no CUDA kernels here).
"""

from __future__ import annotations

import torch


class PairListStager:
    def __init__(self, shape, device, slots: int = 2, max_extent: int | None = None):
        """`shape` = (B, P, D) of one list; `device` = the CUDA device the lists go to; `max_extent` = the largest
        spatial extent the coordinates index (checked against the int16 range when given)."""
        if max_extent is not None and int(max_extent) > 32767:
            raise ValueError(f"coordinates up to {max_extent} do not fit int16: copy the int64 lists as they are")
        self.device = torch.device(device)
        self.stage = [[torch.empty(tuple(shape), dtype=torch.int16).pin_memory() for _ in range(2)] for _ in range(slots)]
        self.free = [None] * slots  # event recorded after the last device copy out of the slot

    def narrow(self, anchors: torch.Tensor, refs: torch.Tensor, slot: int):
        """int64 (or int32) host lists -> this slot's pinned int16 pair.  The caller guarantees coordinates in
        [-32768, 32767] (any list that is valid for an output below 32768 pixels per axis); values outside wrap."""
        if self.free[slot] is not None:
            self.free[slot].synchronize()  # the previous copy out of this slot has finished
        a16, r16 = self.stage[slot]
        a16.copy_(anchors)
        r16.copy_(refs)
        return a16, r16

    def copied(self, slot: int, stream=None) -> None:
        """Call after enqueuing the host -> device copies of `narrow(..., slot)`'s result."""
        ev = torch.cuda.Event()
        ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
        self.free[slot] = ev

    def upload(self, anchors: torch.Tensor, refs: torch.Tensor, slot: int = 0, stream=None):
        """Convenience: narrow + asynchronous copy; returns the device int16 lists."""
        a16, r16 = self.narrow(anchors, refs, slot)
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            out = a16.to(self.device, non_blocking=True), r16.to(self.device, non_blocking=True)
        self.copied(slot, st)
        return out
