import time, torch
dev = torch.device("cuda", 0)
h_in = torch.empty((640000, 6)).pin_memory(); d_in = torch.empty((640000, 6), device=dev)
d_out = torch.empty((640000, 4), device=dev); h_out = torch.empty((640000, 4)).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(n, both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
run(5, True)
print(f"H2D only 15.4MB: {run(50, False):.3f} ms; H2D+D2H concurrent (15.4+10.2MB): {run(50, True):.3f} ms")
