import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from quisk_b200.rx import Channelizer
h = bench.pfb_proto()
n = 1 << 24
x = bench.synth_block_torch(torch, 1, n, "cuda", 3)[0].contiguous()
nf = n // 512
y = torch.zeros((1024, nf), dtype=torch.complex128, device="cuda")
ch = Channelizer(1024, 512, h)
ch.set_option(1, int(sys.argv[1]) if len(sys.argv) > 1 else 32768)
for _ in range(3):
    ch.seek(0); ch.process(x.data_ptr(), n, y.data_ptr(), nf, int(sys.argv[2]) if len(sys.argv) > 2 else 0, 0)
torch.cuda.synchronize()
