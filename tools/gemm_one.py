import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcd_b200 as P
dev = torch.device("cuda"); M, N, K = 131328, 1536, 512
a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3): P.ops.linear(a, w, bias, out=out)
torch.cuda.synchronize(); print("ok")
