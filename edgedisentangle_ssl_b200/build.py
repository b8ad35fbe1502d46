"""Build libedis.so (sm_100a) in-tree with nvcc.  `python -m edgedisentangle_ssl_b200.build`."""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libedis.so")


def sources():
    return sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + [os.path.join(ROOT, "include", "edis.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines: extra -D macros (tuning variants, see csrc/disga.cu); out: alternative .so path."""
    if out is None and not defines and not force and not needs_build():
        return LIB
    lib_out = out or LIB
    tag = "" if not defines else "_" + "_".join(d.replace("=", "") for d in defines)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(PKG, "build", os.path.basename(src)[:-3] + tag + ".o")
        cmd = [nvcc] + ["-D" + d for d in defines] + [ "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed building libedis.so")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_out] + objs)
    return lib_out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
