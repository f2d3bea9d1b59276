"""Run the tcgen05 descriptor probes (csrc/umma_probe.cu) and print one JSON line per probe.
Executed in a subprocess by tests/test_gpu_e_umma_probe.py (and by tools/gpu_firstlight.sh) so that a
faulting kernel cannot take the test process down with it."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np


def bf16_bytes(vals: np.ndarray) -> np.ndarray:
    """small integers -> bf16 bit patterns (exact)"""
    f = vals.astype(np.float32)
    return (f.view(np.uint32) >> 16).astype(np.uint16)


def bf16_image_to_f32(img16: np.ndarray) -> np.ndarray:
    return (img16.astype(np.uint32) << 16).view(np.float32)


def expected(img16, n, k_steps, a_off, a_lbo, a_sbo, a_kstep, b_off, b_lbo, b_sbo, b_kstep, swap=False):
    """K-major no-swizzle model: element (row r, k) of an operand lives at
    off + (r//8)*SBO + (r%8)*16 + (k//8)*LBO + (k%8)*2   (swap=True exchanges the roles of LBO/SBO)."""
    f = bf16_image_to_f32(img16)
    D = np.zeros((128, n), np.float64)
    for ks in range(k_steps):
        def gather(rows, off, lbo, sbo):
            if swap:
                lbo, sbo = sbo, lbo
            r = np.arange(rows)[:, None]; k = np.arange(16)[None, :]
            addr = off + (r // 8) * sbo + (r % 8) * 16 + (k // 8) * lbo + (k % 8) * 2
            return f[addr // 2]
        A = gather(128, a_off + ks * a_kstep, a_lbo, a_sbo)
        B = gather(n, b_off + ks * b_kstep, b_lbo, b_sbo)
        D += A.astype(np.float64) @ B.astype(np.float64).T
    return D.astype(np.float32)


PROBES = {
    # name: (n, k_steps, a_off, a_lbo, a_sbo, a_kstep, b_off, b_lbo, b_sbo, b_kstep)
    "baseline_contiguous":   (128, 1, 0,            2048, 128, 0,    16384, 2048, 128, 0),
    "n64":                   (64,  1, 0,            2048, 128, 0,    16384, 1024, 128, 0),
    "k_chain_4_steps":       (128, 4, 0,            2048, 128, 4096, 16384, 2048, 128, 4096),
    "a_shift_plus_one_row":  (128, 1, 16,           2048, 128, 0,    16384, 2048, 128, 0),
    "a_shift_minus_one_row": (128, 1, 4096 - 16,    2048, 128, 0,    16384, 2048, 128, 0),
    "a_sbo_144":             (128, 1, 0,            2592, 144, 0,    16384, 2048, 128, 0),
    "conv_layout_tap_-1_-1": (128, 4, (8 + 18 - 19) * 16, 2592, 144, 2 * 2592, 49152, 2048, 128, 4096),
    "conv_layout_tap_+1_+1": (128, 4, (8 + 18 + 19) * 16, 2592, 144, 2 * 2592, 49152, 2048, 128, 4096),
}


def main():
    import othello_reinforcement_learning_test_b200 as pkg
    ctx = pkg.Context.default(0)
    only = sys.argv[1:] or list(PROBES)
    rng = np.random.default_rng(0)
    image_bytes = 96 * 1024
    img16 = bf16_bytes(rng.integers(-3, 4, image_bytes // 2))
    for name in only:
        n, ks, *rest = PROBES[name]
        out = np.empty((128, n), np.float32)
        rc = ctx.lib.oth_debug_umma_probe(ctx.handle, img16.ctypes.data, image_bytes, n, ks, *rest, out.ctypes.data)
        if rc != 0:
            print(json.dumps({"probe": name, "rc": rc, "error": pkg._lib.last_error()}), flush=True)
            break
        want = expected(img16, n, ks, *rest)
        alt = expected(img16, n, ks, *rest, swap=True)
        print(json.dumps({"probe": name, "rc": 0, "match": bool(np.array_equal(out, want)),
                          "match_swapped_lbo_sbo": bool(np.array_equal(out, alt)),
                          "max_abs_diff": float(np.abs(out - want).max()),
                          "rows_ok": int((out == want).all(axis=1).sum())}), flush=True)


if __name__ == "__main__":
    main()
