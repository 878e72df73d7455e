"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "_lib"
LIB_PATH = LIB_DIR / "libganslate_b200.so"
SOURCES = ["api.cu", "pack.cu", "pack_v2.cu", "layout.cu", "pad.cu", "loss.cu", "instnorm.cu", "instnorm_fast.cu", "instnorm_v2.cu", "instnorm_v3.cu", "igemm_data.cu", "igemm_tma.cu", "igemm_halo.cu", "igemm_halo_narrow.cu", "igemm_xsplit.cu", "igemm_pair.cu", "igemm_cg2.cu", "igemm_wgrad.cu", "igemm_wgrad_narrow.cu", "patchnce.cu", "patch_mlp.cu",
           "adam.cu", "ssim.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return str(Path(cuda_home) / "bin" / "nvcc")


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    deps = [src] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "ganslate_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link libganslate_b200.so. Returns the library path."""
    LIB_DIR.mkdir(exist_ok=True)
    objs = []
    procs = []
    for name in SOURCES:
        src = CSRC / name
        if not src.exists():
            continue
        obj = LIB_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd))
            procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    relink = force or not LIB_PATH.exists()
    for name, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{out}")
        relink = True
    if relink or any(o.stat().st_mtime > LIB_PATH.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))
