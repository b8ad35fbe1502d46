"""Microbenchmark: weight-gradient GEMM gw = x^T g (K = N nodes) formulations on B200."""
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
N = 2_400_000
def bench(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): r = f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r
def tf32_hi(t):
    return ((t.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
for F, W in ((100, 1536), (64, 1536), (512, 64)):
    x = torch.randn(N, F, device=dev); g = torch.randn(N, W, device=dev)
    ref = (x.double().t() @ g.double()) if F * W < 200000 else None
    def err(r):
        return float(((r.double() - ref).abs().max() / ref.abs().max())) if ref is not None else -1
    t, r = bench(lambda: x.t() @ g); print(F, W, "x.t()@g            %.2f ms err %.2e" % (t, err(r)))
    t, r = bench(lambda: (g.t() @ x).t()); print(F, W, "(g.t()@x).t()      %.2f ms err %.2e" % (t, err(r)))
    for S in (16, 64, 256):
        xs, gs = x.view(S, N // S, F), g.view(S, N // S, W)
        t, r = bench(lambda: torch.bmm(xs.transpose(1, 2), gs).sum(0)); print(F, W, "bmm S=%-4d         %.2f ms err %.2e" % (S, t, err(r)))
    def three():
        xh = tf32_hi(x); xl = x - xh
        gh = tf32_hi(g); gl = g - gh
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            a2 = torch.cat([xh, xl], 1).t() @ gh      # xh^T gh, xl^T gh
            r = a2[:F] + a2[F:] + xh.t() @ gl
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
        return r
    t, r = bench(three); print(F, W, "3xTF32 (split g)   %.2f ms err %.2e" % (t, err(r)))
    torch.backends.cuda.matmul.allow_tf32 = True
    t, r = bench(lambda: x.t() @ g); print(F, W, "1xTF32             %.2f ms err %.2e" % (t, err(r)))
    torch.backends.cuda.matmul.allow_tf32 = False
    # gx-type GEMM: g @ w^T
    if W == 1536:
        w = torch.randn(F, W, device=dev)
        t, r = bench(lambda: g @ w.t()); print(F, W, "gx = g@w.t()       %.2f ms" % t)
        def gx3():
            wh = tf32_hi(w); wl = w - wh
            gh = tf32_hi(g); gl = g - gh
            torch.backends.cuda.matmul.allow_tf32 = True
            try:
                a2 = gh @ torch.cat([wh, wl], 0).t()
                r = a2[:, :F] + a2[:, F:] + gl @ wh.t()
            finally:
                torch.backends.cuda.matmul.allow_tf32 = False
            return r
        t, r = bench(gx3); print(F, W, "gx 3xTF32          %.2f ms" % t)
    del x, g
