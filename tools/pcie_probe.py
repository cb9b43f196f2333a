"""dev: PCIe copy bandwidth of the box, each direction alone and both at once (pinned host memory, two streams)."""
import time
import torch

n = 28_800_448 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for _ in range(2):
    run(True, True, 3)
tu, td, tb = run(True, False), run(False, True), run(True, True)
gb = n * 8 / 1e9
print(f"H2D alone {gb / tu:.1f} GB/s ({tu * 1e3:.3f} ms), D2H alone {gb / td:.1f} GB/s ({td * 1e3:.3f} ms), "
      f"both at once {2 * gb / tb:.1f} GB/s combined ({tb * 1e3:.3f} ms per pair)")
