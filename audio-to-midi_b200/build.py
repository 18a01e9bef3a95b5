"""In-tree build of libaudio2midi_b200.so for sm_100a (explicit nvcc, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libaudio2midi_b200.so")
LIB_F16 = os.path.join(OUT_DIR, "libaudio2midi_b200_f16.so")     # -DA2M_OP_F16: binary16 tensor-core operands (inference variant)
LIBS = {"bf16": LIB, "f16": LIB_F16}
SOURCES = ["a2m_api.cu", "modelutil.cpp"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "cnn_kernels.cuh", "attention.cuh", "block_fused.cuh", "block_mid.cuh", "ffn_fused.cuh", "qkv_fused.cuh", "postattn_fused.cuh", "block256_fused.cuh", "gemm_pair.cuh", "gemm_tc2.cuh", "gemm_wgrad.cuh", "attention_bwd.cuh", "train_kernels.cuh", "block_mid_bwd.cuh", "a2m_train.inc", "audio_prep.cuh", "event_metrics.cuh", os.path.join("..", "..", "include", "a2m.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, precision: str = "bf16") -> str:
    """Compile the CUDA library if missing or older than its sources; returns its path.  Serialised by a file lock and
    written through a temporary name, so that the ranks of a torchrun launch never compile into (or load) a half-written
    library: the first rank builds, the others wait and find it up to date."""
    LIB = LIBS[precision]       # noqa: N806  (shadows the module constant for the rest of the function)
    if not force and not _stale(LIB):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libaudio2midi_b200.so (there is no fallback path)")
    os.makedirs(OUT_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(OUT_DIR, f".build.{precision}.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale(LIB):      # another process built it while this one waited
                return LIB
            extra = os.environ.get("A2M_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DA2M_FFN_TIMING for the in-kernel timeline
            if precision == "f16":
                extra.append("-DA2M_OP_F16")
            tmp = LIB + f".tmp{os.getpid()}"
            cmd = [nvcc, *NVCC_FLAGS, *extra, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl", "-o", tmp]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


XLA_LIB = os.path.join(OUT_DIR, "liba2m_xla_ffi.so")


def build_xla_ffi() -> str:
    """Compiles csrc/a2m_xla_ffi.cc (the XLA typed-FFI handlers) against jaxlib's headers.  Only possible where jax is
    installed -- not in this image: raises RuntimeError otherwise.  Host-only C++ (it forwards device pointers
    to the C ABI), linked against the in-tree libaudio2midi_b200.so."""
    try:
        import jax.ffi
        inc = jax.ffi.include_dir()
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(f"jax.ffi is not importable here ({e}): the XLA-FFI handlers cannot be built") from e
    build()
    src = os.path.join(CSRC, "a2m_xla_ffi.cc")
    if os.path.exists(XLA_LIB) and os.path.getmtime(XLA_LIB) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return XLA_LIB
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{inc}", f"-I{cuda_inc}", src, "-o", XLA_LIB,
           f"-L{OUT_DIR}", "-laudio2midi_b200", f"-Wl,-rpath,{OUT_DIR}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed on a2m_xla_ffi.cc:\n" + res.stdout + res.stderr)
    return XLA_LIB


def build_all(force: bool = False) -> dict:
    """Both operand-format variants, compiled concurrently (two nvcc processes)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as ex:
        futs = {p: ex.submit(build, force, False, p) for p in LIBS}
        return {p: f.result() for p, f in futs.items()}


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build(force=True, precision="f16"))
