import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import ndrustfft_b200 as nb
x = torch.complex(torch.rand((8192, 8192), device="cuda"), torch.rand((8192, 8192), device="cuda")); y = torch.empty_like(x)
h = nb.FftHandler(8192, np.float32)
os.environ["NDFB_FS_L2_KB"] = sys.argv[1]
for _ in range(2): nb.ndfft(x, y, h, 0)
torch.cuda.synchronize()
