"""Generate tests/golden/compression.npz by running the REFERENCE's own export functions
(/root/reference/experiment_scripts/compression.py:16-106) on seeded small grids.

The reference file is a script (argparse + model loading at import time), so only the two function definitions are
extracted from its source with `ast` and executed unmodified, with the real cv2 writing into a temporary directory;
the PNGs are read back and stored.  Runs only where /root/reference exists; the fixture travels, this script documents
how it was made.  usage: python oracle/make_golden_compression.py
"""
import ast
import math
import os
import tempfile

import numpy as np
import torch

REF = "/root/reference/experiment_scripts/compression.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "compression.npz")


def reference_functions():
    import cv2
    tree = ast.parse(open(REF).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("compress_keyframes", "compress_sparse_grid")]
    assert len(keep) == 2
    ns = {"torch": torch, "np": np, "cv2": cv2, "math": math, "os": os, "unit_multiplier": 2.0 ** 8 - 1.0}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns["compress_keyframes"], ns["compress_sparse_grid"]


def main():
    import cv2
    ck, cs = reference_functions()
    kcfg = {"n_levels": 5, "n_features_per_level": 2, "per_level_scale": 1.35}
    scfg = {"n_features_per_level": 2}
    g = torch.Generator().manual_seed(123)
    n_cells = sum((int(math.ceil(math.exp(i * math.log(1.35)) * 16 - 1)) + 1) ** 2 for i in range(5))
    kparams = (torch.rand(n_cells * 2, generator=g) - 0.5) * 0.7
    sparams = (torch.rand(4, 6, 5, 2, generator=g) - 0.3) * 1.3
    out = {"kparams": kparams.numpy(), "sparams": sparams.numpy()}
    with tempfile.TemporaryDirectory() as tmp:
        ck(kparams, kcfg, os.path.join(tmp, "k"))
        cs(sparams, scfg, os.path.join(tmp, "s"))
        for d in range(2):
            for i in range(5):
                out[f"k_d{d}_l{i}"] = cv2.imread(os.path.join(tmp, "k", f"dim{d}", f"{i:02d}.png"), cv2.IMREAD_GRAYSCALE)
            for t in range(4):
                out[f"s_d{d}_t{t}"] = cv2.imread(os.path.join(tmp, "s", f"dim{d}", f"{t:05d}.png"), cv2.IMREAD_GRAYSCALE)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
