"""Launch the d=40 self-attention at the UNet's 64x64-level shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instancediffusion_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
C, d, N = 320, 40, 4096
qkv = (torch.randn(B * N, 3 * C, device=dev)).half()
for _ in range(3):
    ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], batch=B, heads=8, head_dim=d, nq=N, n0=N, scale=d ** -0.5)
torch.cuda.synchronize()
print("ok")
