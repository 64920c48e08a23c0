"""How long does the H2D staging of one c2 batch take alone?  (ids 26 x 65536 x int64 = 13.6 MB, dense 13 x 65536 x fp32 = 3.4 MB,
labels 0.26 MB; pinned host memory, one cudaMemcpyAsync per block -- what Trainer.stage issues)"""
import torch
B, F, C = 65536, 26, 13
dev = torch.device("cuda", 0)
ids_h = torch.zeros((F, B), dtype=torch.int64).pin_memory()
den_h = torch.zeros((C, B), dtype=torch.float32).pin_memory()
lab_h = torch.zeros((B,), dtype=torch.float32).pin_memory()
ids_d, den_d, lab_d = ids_h.to(dev), den_h.to(dev), lab_h.to(dev)
ids32_h = torch.zeros((F, B), dtype=torch.int32).pin_memory()
ids32_d = ids32_h.to(dev)
def run(n, fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def c64():
    ids_d.copy_(ids_h, non_blocking=True); den_d.copy_(den_h, non_blocking=True); lab_d.copy_(lab_h, non_blocking=True)
def c32():
    ids32_d.copy_(ids32_h, non_blocking=True); den_d.copy_(den_h, non_blocking=True); lab_d.copy_(lab_h, non_blocking=True)
for name, fn, nbytes in (("int64 ids", c64, F * B * 8 + C * B * 4 + B * 4), ("int32 ids", c32, F * B * 4 + C * B * 4 + B * 4)):
    run(20, fn)
    ms = run(200, fn)
    print(f"H2D of one c2 batch, {name}: {nbytes / 1e6:.1f} MB in {ms * 1e3:.1f} us = {nbytes / ms / 1e6:.1f} GB/s")
big = torch.zeros((256 << 20,), dtype=torch.uint8).pin_memory()
big_d = big.to(dev)
ms = run(10, lambda: big_d.copy_(big, non_blocking=True))
print(f"H2D 256 MiB pinned: {big.numel() / ms / 1e6:.1f} GB/s")
