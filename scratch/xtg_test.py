"""Weight-gradient GEMM x^T g with a very long reduction (K = nodes): fp32 SIMT (library), the chunked bmm used so far, and
a tensor-core form -- x split exactly into two TF32 terms (x is narrow: free), g rounded TO NEAREST onto the TF32 grid
(unbiased), one TF32 GEMM with M = 2F -- against a float64 evaluation."""
import time, torch
torch.manual_seed(0)
dev = "cuda"
def tf32_rn(t): return ((t.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
def timed(f, n=5):
    f(); torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n, out
def xtg_tf32(x, g, chunks=1):
    xh = tf32_rn(x); xl = x - xh
    gr = tf32_rn(g)
    a = torch.cat([xh, xl], 1)
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        if chunks == 1:
            o = a.t() @ gr
        else:
            n = x.shape[0]; m = (n // chunks) * chunks
            o = torch.bmm(a[:m].view(chunks, m // chunks, -1).transpose(1, 2), gr[:m].view(chunks, m // chunks, -1)).sum(0)
            if m < n: o += a[m:].t() @ gr[m:]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    F = x.shape[1]
    return o[:F] + o[F:]
def xtg_tf32_noround(x, g):
    xh = tf32_rn(x); xl = x - xh
    a = torch.cat([xh, xl], 1)
    torch.backends.cuda.matmul.allow_tf32 = True
    try: o = a.t() @ g
    finally: torch.backends.cuda.matmul.allow_tf32 = False
    return o[:x.shape[1]] + o[x.shape[1]:]
def chunked(x, g, chunks=256):
    n = x.shape[0]; m = (n // chunks) * chunks
    out = torch.bmm(x[:m].view(chunks, m // chunks, -1).transpose(1, 2), g[:m].view(chunks, m // chunks, -1)).sum(0)
    if m < n: out += x[m:].t() @ g[m:]
    return out
for S, F, W in ((2400000, 100, 1024), (2400000, 64, 1536), (1700000, 100, 1024), (2400000, 64, 512)):
    x = torch.randn(S, F, device=dev); g = torch.randn(S, W, device=dev) * torch.rand(S, 1, device=dev)
    g = g + 0.05                        # a non-zero mean: exposes a truncation bias
    ref = torch.zeros(F, 64, dtype=torch.float64, device=dev); sa = torch.zeros_like(ref)
    for s0 in range(0, S, 200000):
        xs = x[s0:s0 + 200000].double(); gs = g[s0:s0 + 200000, :64].double()
        ref += xs.t() @ gs; sa += xs.abs().t() @ gs.abs()
    for name, f in (("fp32 library", lambda: x.t() @ g), ("fp32 chunked bmm", lambda: chunked(x, g)), ("tf32 x-split, g RN", lambda: xtg_tf32(x, g)),
                    ("tf32 x-split, g RN, 64 slabs", lambda: xtg_tf32(x, g, 64)), ("tf32 x-split, g raw", lambda: xtg_tf32_noround(x, g))):
        ms, out = timed(f)
        err = (out[:, :64].double() - ref).abs()
        print("S=%d F=%d W=%d  %-30s %6.2f ms   max err / sum|terms| %.2e   max err / max|result| %.2e" % (
            S, F, W, name, ms, float((err / sa).max()), float(err.max() / ref.abs().max())))
    del x, g
