import torch, time
dev = torch.device("cuda", 0)
n = 78 * (1 << 20) // 8
src = torch.empty(n, dtype=torch.float64, device=dev).normal_()
dst = torch.empty(n, dtype=torch.float64).pin_memory()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
one = t(lambda: dst.copy_(src, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    h = n // 2
    with torch.cuda.stream(s1): dst[:h].copy_(src[:h], non_blocking=True)
    with torch.cuda.stream(s2): dst[h:].copy_(src[h:], non_blocking=True)
two_t = t(two)
def chunks(k):
    c = n // k
    for i in range(k): dst[i*c:(i+1)*c].copy_(src[i*c:(i+1)*c], non_blocking=True)
print("one copy %.1f GB/s  two streams %.1f GB/s  16 chunks one stream %.1f GB/s  64 chunks %.1f GB/s" % (
    n*8/one/1e9, n*8/two_t/1e9, n*8/t(lambda: chunks(16))/1e9, n*8/t(lambda: chunks(64))/1e9))
src2 = torch.empty(11 * (1 << 20) // 8, dtype=torch.float64).pin_memory(); d2 = torch.empty_like(src2, device=dev)
both = t(lambda: (dst.copy_(src, non_blocking=True), None))
s3 = torch.cuda.Stream()
def bidir():
    with torch.cuda.stream(s3): d2.copy_(src2, non_blocking=True)
    dst.copy_(src, non_blocking=True)
print("D2H with concurrent H2D: %.1f GB/s" % (n*8/t(bidir)/1e9))
