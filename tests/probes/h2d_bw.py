"""Pinned H2D copy rate of this box: one stream vs two concurrent streams (the floor of bench.py's
e2e number is 158 MB / this rate)."""
import torch
n = 130 * 1024 * 1024 // 4
x = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
def run(nstreams, reps=10):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    parts = [(i * n // nstreams, (i + 1) * n // nstreams) for i in range(nstreams)]
    def once():
        for s, (a, b) in zip(streams, parts):
            with torch.cuda.stream(s):
                d[a:b].copy_(x[a:b], non_blocking=True)
    for _ in range(3): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for _ in range(reps): once()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"H2D pinned 130 MiB, {nstreams} stream(s): {ms:.3f} ms -> {n*4/ms/1e6:.1f} GB/s")
for k in (1, 2, 4):
    run(k)
