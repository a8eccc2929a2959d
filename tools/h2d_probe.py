import torch, time
x = torch.empty(1358954496 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("H2D GB/s", 1.358954496 / dt)
