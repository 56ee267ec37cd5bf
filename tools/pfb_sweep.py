import sys, numpy as np, torch, time
sys.path.insert(0, "/root/repo")
import bench
from quisk_b200.rx import Channelizer
h = bench.pfb_proto()
n = 1 << 24
x = bench.synth_block_torch(torch, 1, n, "cuda", 3)[0].contiguous()
nf = n // 512
y = torch.zeros((1024, nf), dtype=torch.complex128, device="cuda")
y2 = torch.zeros((nf, 1024), dtype=torch.complex128, device='cuda')
for sl, gen, lay, pipe, ff, rg, pfx in [(65536, 0, 0, 0, 2, 16, 0), (65536, 0, 0, 0, 2, 16, 1), (65536, 0, 1, 0, 2, 16, 1), (65536, 0, 0, 0, 4, 16, 1)]:
    ch = Channelizer(1024, 512, h)
    ch.set_option(1, sl); ch.set_option(2, gen); ch.set_option(3, pipe); ch.set_option(4, ff); ch.set_option(5, rg); ch.set_option(6, pfx)
    def step():
        ch.seek(0); ch.process(x.data_ptr(), n, (y2 if lay else y).data_ptr(), 1024 if lay else nf, lay, 0)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("prefetch %d fftframes %d ring %d pipe %d layout %d slice %5d generic %d: %.3f ms  %.1f GS/s  frac %.3f" % (pfx, ff, rg, pipe, lay, sl, gen, ms, n / ms / 1e6, n * 48 / ms / 1e6 / 6549.1), flush=True)
    ch.close()
