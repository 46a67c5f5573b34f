"""PCIe floor for the host-buffer entry point: pinned H2D / D2H bandwidth alone and concurrently (full duplex)."""
import torch, time
n = 398131200
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
for _ in range(2): run(True, True)
for name, u, d in (("H2D", True, False), ("D2H", False, True), ("both", True, True)):
    dt = run(u, d); print("%s: %.2f ms per 398 MB (%.1f GB/s per direction)" % (name, dt * 1e3, n / dt / 1e9))
