"""In-tree build of libdist_b200.so with nvcc for sm_100a (no torch extension machinery: the library is plain C ABI)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdist_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + \
    os.environ.get("DIST_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DDIST_TC_TIMELINE for the per-layer timeline of mlp_tc.cu
# march.cu mirrors PyTorch's separately-rounded elementwise ops: no FMA contraction there
SOURCES = {"abi.cu": [], "warp.cu": [], "mesh.cu": ["-fmad=false"], "march.cu": ["-fmad=false"], "mlp_simt.cu": ["-Xptxas", "-v"], "mlp_tc.cu": ["-Xptxas", "-v"]}


def _stamp():
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h", ".inc")):
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def _mc_tables():
    """csrc/mc_tables.inc (marching-cubes case tables) is derived by mc_tables.py; (re)write it before hashing the sources."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_dist_mc_tables", os.path.join(HERE, "mc_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.write_header(os.path.join(CSRC, "mc_tables.inc"))


def build(force=False, verbose=False):
    _mc_tables()
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [NVCC] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        if verbose:
            sys.stderr.write(out)
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
