import torch
x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")   # 8 GiB
for _ in range(3):
    x.fill_(1.5)
torch.cuda.synchronize()
