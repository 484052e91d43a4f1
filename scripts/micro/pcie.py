"""Host<->device copy rates of this box (pinned memory): one direction alone, both at once."""
import torch, time
n = 65536 * 1000
h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
gb = n * 4 / 1e6
print(f"262 MB: H2D alone {a:.2f} ms ({gb/a:.1f} GB/s), D2H alone {b:.2f} ms ({gb/b:.1f} GB/s), both at once {c:.2f} ms "
      f"({2*gb/c:.1f} GB/s total)")
