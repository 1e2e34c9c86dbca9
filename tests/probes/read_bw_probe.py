"""Read-only bandwidth over a distance-matrix-sized buffer with stock torch reductions (a row-wise one:
one CTA or warp per row, the access pattern of the counting kernels; a flat one: the grid sweeps the
buffer linearly), to tell the memory system's ceiling for the pattern from the counting kernel's own cost."""
import torch
Q, G = 3368, 15913
d = torch.randn(Q, G, device="cuda")
gb = d.numel() * 4 / 1e9
def bench(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, f in [("row sum  d.sum(1)", lambda: d.sum(1)), ("row max  d.amax(1)", lambda: d.amax(1)),
                ("flat sum d.sum()", lambda: d.sum()), ("flat max d.amax()", lambda: d.amax()),
                ("col sum  d.sum(0)", lambda: d.sum(0)),
                ("copy (r+w bytes)", lambda: torch.empty_like(d).copy_(d))]:
    ms = bench(f)
    mult = 2 if name.startswith("copy") else 1
    print(f"{name:22s} {ms:.4f} ms  {mult * gb / (ms * 1e-3):7.0f} GB/s", flush=True)
