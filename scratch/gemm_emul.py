import ctypes, os, sys, time
mode = sys.argv[1]
if mode != "torch":
    os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
    if len(sys.argv) > 2: os.environ["CUBLAS_EMULATION_STRATEGY"] = sys.argv[2]
    for n in ("libcublasLt.so.12", "libcublas.so.12"):
        ctypes.CDLL("/usr/local/cuda/lib64/" + n, mode=ctypes.RTLD_GLOBAL)
import torch
print(mode, "cublas", torch.backends.cuda.preferred_blas_library(), os.environ.get("CUBLAS_EMULATE_SINGLE_PRECISION"))
import subprocess
print(subprocess.run("grep -E 'libcublas' /proc/%d/maps | awk '{print $6}' | sort -u" % os.getpid(), shell=True, capture_output=True, text=True).stdout)
dev = "cuda"
N = 2_400_000
for (M, K, O, tag) in ((N, 100, 1536, "proj1"), (N, 64, 1536, "proj2"), (N, 512, 64, "fuse")):
    a = torch.randn(M, K, device=dev); b = torch.randn(K, O, device=dev)
    for _ in range(3): c = a @ b
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): c = a @ b
    torch.cuda.synchronize(); ms = (time.perf_counter() - t) / 5 * 1e3
    ref = (a[:4096].double() @ b.double())
    err = ((c[:4096].double() - ref).abs().max() / ref.abs().max()).item()
    g = torch.randn(M, O, device=dev)
    for _ in range(2): dw = a.t() @ g
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): dw = a.t() @ g
    torch.cuda.synchronize(); ms2 = (time.perf_counter() - t) / 5 * 1e3
    refw = (a[:200000].double().t() @ g[:200000].double()); dw2 = a[:200000].t() @ g[:200000]
    err2 = ((dw2.double() - refw).abs().max() / refw.abs().max()).item()
    for _ in range(2): dx = g @ b.t()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): dx = g @ b.t()
    torch.cuda.synchronize(); ms3 = (time.perf_counter() - t) / 5 * 1e3
    print("%s %-6s fwd %.2f ms (%.1f TF/s) err %.1e | dW %.2f ms err %.1e | dX %.2f ms" % (mode, tag, ms, 2*M*K*O/ms/1e9, err, ms2, err2, ms3))
    del a, b, c, g, dw, dx
