import torch, time
x = torch.empty(1_500_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(1_500_000_000, dtype=torch.uint8).pin_memory()
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def copy():
    with torch.cuda.stream(sa): h.copy_(x, non_blocking=True)
def kern(n=60):
    with torch.cuda.stream(sb):
        for _ in range(n): a @ a
for _ in range(2): copy(); kern(); torch.cuda.synchronize()
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
print("copy %.1f ms, kernels %.1f ms, both %.1f ms" % (t(copy), t(kern), t(lambda: (copy(), kern()))))
