import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gaussiansplatting.jl_b200"))
from gsrast import ssim
x = torch.rand((1, 3, 1088, 1920), device="cuda"); t = torch.rand_like(x)
for _ in range(3):
    m, d0, d1, d2 = ssim.ssim_forward(x, t, train=True)
    g = ssim.ssim_backward(x, t, torch.full_like(x, 1e-7), d0, d1, d2)
torch.cuda.synchronize()
