#!/usr/bin/env python
"""Does a stream waiting on an event recorded in the MIDDLE of a train of queued H2D copies resume when that copy
is done, or when the train is done?  (side stream: 12 x 100 MB pinned copies, an event after each)"""
import sys, time
import numpy as np
import torch
torch.cuda.set_device(0)
n = 25_000_000
h = torch.empty((12, n), dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty_like(h, device="cuda")
small_host = np.arange(200_000, dtype=np.int32)
side = torch.cuda.Stream()
cur = torch.cuda.current_stream()
x = torch.zeros(1024, device="cuda")


def run(small_copy, delay_ms):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    side.wait_stream(cur)
    evs = []
    with torch.cuda.stream(side):
        for k in range(12):
            d[k].copy_(h[k], non_blocking=True)
            ev = torch.cuda.Event(enable_timing=True); ev.record(side); evs.append(ev)
    if delay_ms:
        time.sleep(delay_ms / 1e3)
    if small_copy:
        t = torch.from_numpy(small_host).to("cuda", non_blocking=True)     # pageable 0.8 MB on the current stream
    cur.wait_event(evs[2])
    x.add_(1.0)
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(evs[2]), e0.elapsed_time(e1), e0.elapsed_time(evs[-1])


for small in (False, True):
    for delay in (0, 3):
        r = [run(small, delay) for _ in range(4)]
        print(f"pageable small copy on the waiting stream: {small}, host delay {delay} ms -> (copy 3 done, waiter done, train done) ms:",
              " ".join(f"({a:.1f} {b:.1f} {c:.1f})" for a, b, c in r))
