"""Host-buffer front end of the rasterizer: fwd+bwd steps whose inputs and results live in (pinned)
HOST memory, software-pipelined over three CUDA streams.

This is the end-to-end call measured by bench.py's `e2e` leg: every step copies ALL of its inputs
host->device (splat parameters + output cotangents), runs GaussianRasterizer forward and the autograd
backward on the device, and copies ALL results device->host (color, allmap, radii and every
gradient).  PCIe is full duplex and independent of the SMs, so step i's D2H, step i+1's compute and
step i+2's H2D overlap:

    s_in   : H2D(i+2)  ->
    s_comp :            compute(i+1) ->
    s_out  :                          D2H(i)

Device input buffers are double-buffered; events carry the dependencies; nothing is allocated on the
copy streams (results allocated by the op on s_comp are handed to s_out with record_stream).
"""
from typing import Dict

import torch

NAMES = ("means3D", "scales", "rotations", "opacities", "shs")


class HostStepPipeline:
    def __init__(self, rasterizer, host_in: Dict[str, torch.Tensor], host_gc: torch.Tensor,
                 host_go: torch.Tensor, device: torch.device):
        self.rast, self.dev = rasterizer, device
        self.P = host_in["means3D"].shape[0]
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(device) for _ in range(3))
        self.dev_in = [{k: torch.empty_like(host_in[k], device=device) for k in NAMES} for _ in range(2)]
        self.dev_gc = [torch.empty_like(host_gc, device=device) for _ in range(2)]
        self.dev_go = [torch.empty_like(host_go, device=device) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        for e in self.free:
            e.record(self.s_comp)
        self.done = torch.cuda.Event()
        self.n_in = 0
        self.n_comp = 0
        self.trace = None        # profiling: list of (kind, step, start_event, end_event) when set to []

    def _mark(self, stream):
        if self.trace is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def _close(self, kind, step, mark, stream):
        if mark is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            self.trace.append((kind, step, mark, e))

    def h2d(self, host_in, host_gc, host_go):
        """Enqueue the host->device copies of one step's inputs."""
        b = self.n_in % 2
        self.s_in.wait_event(self.free[b])            # compute that last read this buffer has finished
        with torch.cuda.stream(self.s_in):
            mark = self._mark(self.s_in)
            for k in NAMES:
                self.dev_in[b][k].copy_(host_in[k], non_blocking=True)
            self.dev_gc[b].copy_(host_gc, non_blocking=True)
            self.dev_go[b].copy_(host_go, non_blocking=True)
            self.ready[b].record(self.s_in)
            self._close("h2d", self.n_in, mark, self.s_in)
        self.n_in += 1

    def compute_and_d2h(self, host_out, host_grad):
        """Run fwd+bwd of the oldest uploaded step and enqueue the device->host copies of its results."""
        b = self.n_comp % 2
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(self.ready[b])
            mark = self._mark(self.s_comp)
            leaf = {k: self.dev_in[b][k].detach().requires_grad_(True) for k in NAMES}
            m2d = torch.zeros(self.P, 3, device=self.dev, requires_grad=True)
            color, radii, allmap = self.rast(means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"],
                                             opacities=leaf["opacities"], scales=leaf["scales"],
                                             rotations=leaf["rotations"])
            torch.autograd.backward([color, allmap], [self.dev_gc[b], self.dev_go[b]])
            self.done.record(self.s_comp)
            self.free[b].record(self.s_comp)
            self._close("compute", self.n_comp, mark, self.s_comp)
        results = [("color", color.detach()), ("allmap", allmap.detach()), ("radii", radii)]
        grads = [(k, leaf[k].grad) for k in NAMES] + [("means2D", m2d.grad)]
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.done)
            mark = self._mark(self.s_out)
            for k, t in results:
                t.record_stream(self.s_out)
                host_out[k].copy_(t, non_blocking=True)
            for k, t in grads:
                t.record_stream(self.s_out)
                host_grad[k].copy_(t, non_blocking=True)
            self._close("d2h", self.n_comp, mark, self.s_out)
        self.n_comp += 1

    def run(self, steps, host_in, host_gc, host_go, host_out, host_grad):
        """`steps` pipelined fwd+bwd steps on host buffers; returns after everything has landed."""
        for i in range(steps + 1):
            if i < steps:
                self.h2d(host_in, host_gc, host_go)
            if i >= 1:
                self.compute_and_d2h(host_out, host_grad)
        self.s_out.synchronize()
        self.s_comp.synchronize()
