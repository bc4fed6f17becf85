"""What HBM bandwidth do plain torch elementwise kernels reach at the tensor sizes of this network?"""
import torch
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mb in (66, 133, 265, 1060, 4000):
    n = mb * 1000 * 1000 // 2
    x = torch.randn(n, device="cuda", dtype=torch.float16); r = torch.randn_like(x); y = torch.empty_like(x)
    # rotate through several buffers so that nothing is L2-resident between iterations
    xs = [torch.randn_like(x) for _ in range(3)]; ys = [torch.empty_like(x) for _ in range(3)]
    i = [0]
    def copy():
        i[0] = (i[0] + 1) % 3; ys[i[0]].copy_(xs[i[0]])
    def add():
        i[0] = (i[0] + 1) % 3; torch.add(xs[i[0]], r, out=ys[i[0]])
    def read():
        i[0] = (i[0] + 1) % 3; xs[i[0]].max()
    def fill():
        i[0] = (i[0] + 1) % 3; ys[i[0]].fill_(1.0)
    tc, ta, tr, tf = timeit(copy), timeit(add), timeit(read), timeit(fill)
    print(f"{mb:5d} MB tensors: copy {2*mb/tc:7.1f} GB/s  add(3 streams) {3*mb/ta:7.1f} GB/s  read-only(max) {mb/tr:7.1f} GB/s  write-only(fill) {mb/tf:7.1f} GB/s", flush=True)
