import torch, time
n = 278_000_000 // 4
h_in = torch.empty(n).pin_memory(); h_out = torch.empty(n).pin_memory()
d_in = torch.empty(n, device="cuda"); d_out = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=10):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    return dt
for name, u, d in (("H2D", True, False), ("D2H", False, True), ("both", True, True)):
    run(u, d, 2); dt = run(u, d)
    print(f"{name}: {dt*1e3:.2f} ms per 278 MB -> {0.278/dt:.1f} GB/s per direction")
