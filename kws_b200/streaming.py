"""Host-facing inference pipeline: pinned host buffers in, pinned host buffers out.

The recurrence itself is one kernel per call; what a host caller pays for is PCIe.  This helper
splits the batch into row chunks (rows are independent) and runs a
three-stage pipeline on three CUDA streams -- H2D of chunk i+1, the recurrence of chunk i and
D2H of chunk i-1 overlap -- with double-buffered device staging.  It is the public
"host buffers" entry point measured as ``e2e`` by bench.py.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import engine


class HostPipeline:
    def __init__(self, params: Dict[str, torch.Tensor], *, layout: str, T: int, I: int, H: int,
                 chunk_rows: int = 2048, x_dtype: torch.dtype = torch.float32,
                 gate_nl="sigmoid", update_nl="tanh", device: Optional[torch.device] = None,
                 last_state_only: bool = False):
        self.params = params
        self.layout, self.gate_nl, self.update_nl = layout, gate_nl, update_nl
        self.T, self.I, self.H = T, I, H
        self.chunk_rows = int(chunk_rows)
        self.device = device or params["bias_gate"].device
        self.last_state_only = last_state_only
        dev = self.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_cmp = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.xbuf = [torch.empty((self.chunk_rows, T, I), dtype=x_dtype, device=dev) for _ in range(2)]
        oshape = (self.chunk_rows, H) if last_state_only else (self.chunk_rows, T, H)
        self.obuf = [torch.empty(oshape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, x_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """x_host [B,T,I] (pinned), out_host [B,T,H] or [B,H] (pinned).  Returns after the last
        device-to-host copy has completed (host-visible result)."""
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("HostPipeline needs pinned host tensors (torch.empty(..., pin_memory=True))")
        B = x_host.shape[0]
        n = (B + self.chunk_rows - 1) // self.chunk_rows
        self.h2d_bytes = x_host.numel() * x_host.element_size()
        self.d2h_bytes = out_host.numel() * out_host.element_size()
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        for i in range(n):
            b, e = i * self.chunk_rows, min(B, (i + 1) * self.chunk_rows)
            k = i & 1
            rows = e - b
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_cmp[k])          # buffer k's previous compute finished
                self.xbuf[k][:rows].copy_(x_host[b:e], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.ev_in[k])
                self.s_cmp.wait_event(self.ev_out[k])         # buffer k's previous D2H finished
                if self.last_state_only:
                    _, _, _, last = engine.forward(self.xbuf[k][:rows], self.params, None, layout=self.layout,
                                                   batch_first=True, gate_nl=self.gate_nl, update_nl=self.update_nl,
                                                   want_states=False, want_last=True)
                    self.obuf[k][:rows].copy_(last)
                else:
                    engine.forward(self.xbuf[k][:rows], self.params, None, layout=self.layout, batch_first=True,
                                   gate_nl=self.gate_nl, update_nl=self.update_nl, out=self.obuf[k][:rows])
                self.ev_cmp[k].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_cmp[k])
                out_host[b:e].copy_(self.obuf[k][:rows], non_blocking=True)
                self.ev_out[k].record(self.s_out)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_in)
        self.s_out.synchronize()
