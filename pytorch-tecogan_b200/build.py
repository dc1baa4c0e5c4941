"""Build libtecogan_b200.so in-tree with nvcc for sm_100a (and only sm_100a).

    python pytorch-tecogan_b200/build.py [--force] [--verbose]

The library is plain CUDA C++ behind a C ABI (include/tecogan_b200.h); it does not link
against torch.  The built .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtecogan_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def deps():
    d = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    d.append(os.path.join(HERE, "..", "include", "tecogan_b200.h"))
    d.append(os.path.abspath(__file__))
    return d


def source_digest(defines=()):
    """sha256 over the contents of every source / header the library is built from (+ flags): file times do not survive a
    copy to another box or a `git stash`, contents do."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(deps()):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS + list(defines)).encode())
    return h.hexdigest()


def up_to_date(target=OUT, defines=()):
    stamp = target + ".srchash"
    if not (os.path.exists(target) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == source_digest(defines)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: measurement builds (same-box A/B through `bench.py --lib`): libtecogan_b200.<variant>.so compiled
    with extra -D flags next to the product library; never loaded by the package itself."""
    target = OUT if not variant else os.path.join(HERE, f"libtecogan_b200.{variant}.so")
    if not force and up_to_date(target, defines):
        return target
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build" if not variant else f"build_{variant}")
    os.makedirs(bdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libtecogan_b200 (see output above)")
    cmd = [NVCC, "-shared", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(target + ".srchash", "w") as f:
        f.write(source_digest(defines) + "\n")
    return target


if __name__ == "__main__":
    # python build.py [--force] [--verbose] [--variant NAME -DFOO=1 ...]
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=var,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
