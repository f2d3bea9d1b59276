"""oracle/build.py -- TEST INFRASTRUCTURE ONLY.

Builds the two CPU checkers:

* ``oracle/_build/libref_rules.so``  -- our plain-C restatement (ref_rules.c);
* ``oracle/_ref/src/cython/bitboard*.so`` -- the REFERENCE'S OWN Cython bitboard,
  compiled from the source where it lies under /root/reference (never copied
  into this repo; only the compiled module lands in the git-ignored
  ``oracle/_ref/``).  Skipped when /root/reference is absent (GPU box): the
  prebuilt file travels with the snapshot;
* ``oracle/_ref/src/**/*.rbc`` -- the REFERENCE'S OWN Python hot-path modules and
  their callers (MCTS, network, self-play workers, replay buffer, trainer, arena,
  players), byte-compiled with ``py_compile`` from the sources where they lie
  (sourceless byte-code next to the compiled bitboard: outputs only, no reference
  source text enters the repository or its history; a neutral extension because
  snapshots to the GPU box drop ``*.pyc`` -- they are copied to ``*.pyc`` in a
  temp directory when imported).  They let the GPU box run
  the unmodified reference: as the CPU arm of bench.py (``--impl reference``) and
  as the caller side of the drop-in tests (reference arena / self-play / trainer
  code driving this package's classes).

Only tests/, __graft_entry__ (build + smoke) and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("OTHELLO_REFERENCE_ROOT", "/root/reference")

ORACLE_LIB = os.path.join(BUILD_DIR, "libref_rules.so")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force: bool = False) -> str:
    """gcc -O3 -fopenmp the C restatement; returns the path of the .so."""
    src = os.path.join(HERE, "ref_rules.c")
    hdr = os.path.join(HERE, "ref_rules.h")
    if not force and _newer(ORACLE_LIB, src, hdr):
        return ORACLE_LIB
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-std=c11",
           "-Wall", "-Wextra", "-o", ORACLE_LIB, src, "-lm"]
    subprocess.run(cmd, check=True)
    return ORACLE_LIB


def ref_bitboard_path() -> str | None:
    d = os.path.join(REF_OUT, "src", "cython")
    if not os.path.isdir(d):
        return None
    for f in sorted(os.listdir(d)):
        if f.startswith("bitboard") and f.endswith(".so"):
            return os.path.join(d, f)
    return None


def build_ref(force: bool = False) -> str | None:
    """Compile the reference's Cython bitboard into oracle/_ref (outputs only).

    The .pyx/.pxd are read in place; the generated C goes to a temp dir and is
    discarded; only the extension module is kept.  Flags follow the reference's
    setup.py:11-30 (-O3, boundscheck/wraparound off, cdivision on).
    """
    have = ref_bitboard_path()
    pyx = os.path.join(REFERENCE_ROOT, "src", "cython", "bitboard.pyx")
    if not os.path.exists(pyx):
        return have                      # GPU box: use what travelled
    if have and not force and os.path.getmtime(have) >= os.path.getmtime(pyx):
        return have
    import numpy as np
    out_dir = os.path.join(REF_OUT, "src", "cython")
    os.makedirs(out_dir, exist_ok=True)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(out_dir, "bitboard" + ext)
    with tempfile.TemporaryDirectory(prefix="refbuild_") as tmp:
        c_file = os.path.join(tmp, "bitboard.c")
        subprocess.run([sys.executable, "-m", "cython", "-3",
                        "-X", "boundscheck=False", "-X", "wraparound=False", "-X", "cdivision=True",
                        pyx, "-o", c_file], check=True)
        inc = sysconfig.get_paths()["include"]
        subprocess.run(["gcc", "-O3", "-shared", "-fPIC", "-fwrapv", "-w",
                        "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                        "-I", inc, "-I", np.get_include(), c_file, "-o", target], check=True)
    return target


# the reference's Python modules on and around the hot path; package __init__ files
# included so that `import src.train.trainer` resolves exactly as it does in the reference tree
REF_PY_MODULES = [
    "src/__init__.py", "src/cython/__init__.py",
    "src/mcts/__init__.py", "src/mcts/node.py", "src/mcts/mcts.py",
    "src/model/__init__.py", "src/model/net.py",
    "src/train/__init__.py", "src/train/self_play.py", "src/train/parallel_self_play.py", "src/train/buffer.py",
    "src/train/trainer.py",
    "src/eval/__init__.py", "src/eval/arena.py", "src/eval/players.py",
    "benchmark.py",
]


REF_BYTECODE_EXT = ".rbc"      # python byte-code under a neutral extension: snapshots to the GPU box drop *.pyc


def _materialise_bytecode() -> str | None:
    """Copy oracle/_ref/**/*.rbc to <tmp>/oth_ref_py/**/*.pyc (importable, sourceless) and return that directory."""
    if not os.path.exists(os.path.join(REF_OUT, "src", "mcts", "mcts" + REF_BYTECODE_EXT)):
        return None
    dst_root = os.path.join(tempfile.gettempdir(), f"oth_ref_py_{os.getuid()}")
    for rel in REF_PY_MODULES:
        src = os.path.join(REF_OUT, rel[:-3] + REF_BYTECODE_EXT)
        dst = os.path.join(dst_root, rel[:-3] + ".pyc")
        if not os.path.exists(src):
            return None
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst + ".tmp")
            os.replace(dst + ".tmp", dst)
    return dst_root


def ref_python_root() -> str | None:
    """Directory to put on sys.path so that `import src.mcts.mcts` finds the byte-compiled reference."""
    return _materialise_bytecode()


def build_ref_python(force: bool = False) -> str | None:
    """Byte-compile the reference's Python modules into oracle/_ref (sourceless byte-code, outputs only)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "mcts")):
        return ref_python_root()            # GPU box: use what travelled
    import py_compile
    for rel in REF_PY_MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(REF_OUT, rel[:-3] + REF_BYTECODE_EXT)
        if not os.path.exists(src):
            raise RuntimeError(f"reference module {rel} not found under {REFERENCE_ROOT}")
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile = the path shown in tracebacks; UNCHECKED_HASH: the byte-code must load without its source
        py_compile.compile(src, cfile=dst, dfile=f"<reference>/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return ref_python_root()


def main() -> None:
    print("oracle:", build_oracle(force="--force" in sys.argv))
    print("oracle/_ref:", build_ref(force="--force" in sys.argv))
    print("oracle/_ref (python):", build_ref_python(force="--force" in sys.argv))


if __name__ == "__main__":
    main()
