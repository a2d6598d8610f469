"""Times cuFFT decompositions of the 1024^3 D2Z (measurement tool, not product)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genpk_b200 as gp
from genpk_b200 import api
dims = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.cuda.set_device(0)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with gp.Context(dims, 0) as ctx:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.grid_zero()
    print("3d in-place D2Z      %.3f ms" % timeit(lambda: ctx.fft()))
    ptr = ctx.grid_ptr()
    def two():
        ctx.slab_fft_yz(); ctx.slab_fft_x(ptr)
    print("2d yz + 1d x strided %.3f ms" % timeit(two))
    print("   2d yz only        %.3f ms" % timeit(lambda: ctx.slab_fft_yz()))
    print("   1d x only         %.3f ms" % timeit(lambda: ctx.slab_fft_x(ptr)))
x = torch.zeros(dims, dims, dims, dtype=torch.float64, device="cuda")
print("torch rfftn (out of place) %.3f ms" % timeit(lambda: torch.fft.rfftn(x)))
del x
