#!/usr/bin/env python
"""Host<->device copy bandwidth of this box (pinned memory): H2D alone, D2H alone, both directions at once.
The end-to-end leg of bench.py is bound by the simultaneous figure."""
import json
import torch

n = 1 << 28  # 2 GiB of float64 per buffer
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return reps * n * 8 / (a.elapsed_time(b) * 1e-3) / 1e9


run(True, True, 1)
print(json.dumps({"h2d_alone_GBs": run(True, False), "d2h_alone_GBs": run(False, True),
                  "each_direction_when_simultaneous_GBs": run(True, True)}))
