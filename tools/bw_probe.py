import torch
dev = torch.device('cuda')
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
for mb in (64, 256, 1024):
    a = torch.empty(mb << 18, device=dev); b = torch.empty_like(a)
    nb = a.numel() * 4
    print(mb, 'MB  fill %.0f GB/s' % (nb / t(lambda: a.zero_()) / 1e9),
          ' copy(r+w) %.0f GB/s' % (2 * nb / t(lambda: b.copy_(a)) / 1e9),
          ' read(sum) %.0f GB/s' % (nb / t(lambda: a.sum()) / 1e9))
