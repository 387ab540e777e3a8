"""CPU tests of the codec export / import mirror (nvp_b200/compression.py) against fixtures produced by the reference's
own compress_keyframes / compress_sparse_grid (oracle/make_golden_compression.py; compression.py:16-106)."""
import os
import types

import numpy as np
import torch

from nvp_b200 import compression, eval_utils

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "compression.npz")
KCFG = {"n_levels": 5, "n_features_per_level": 2, "per_level_scale": 1.35}
SCFG = {"n_features_per_level": 2}


def test_level_layout_matches_the_reference_constants():
    res, offs = compression.level_layout({"n_levels": 16, "per_level_scale": 1.35})
    assert res == [16, 22, 30, 40, 54, 72, 97, 131, 177, 239, 322, 435, 587, 792, 1069, 1443]      # eval.py:29-30, SURVEY A.2
    assert offs[-1] == 4616112


def test_export_images_are_bit_identical_to_the_reference_export(tmp_path):
    g = np.load(GOLD)
    kparams, sparams = torch.from_numpy(g["kparams"]), torch.from_numpy(g["sparams"])
    compression.compress_keyframes(kparams, KCFG, str(tmp_path / "k"))
    compression.compress_sparse_grid(sparams, SCFG, str(tmp_path / "s"))
    for d in range(2):
        for i in range(5):
            got = compression._imread_gray(str(tmp_path / "k" / f"dim{d}" / f"{i:02d}.png"))
            assert got.dtype == np.uint8 and np.array_equal(got, g[f"k_d{d}_l{i}"]), (d, i)
        for t in range(4):
            got = compression._imread_gray(str(tmp_path / "s" / f"dim{d}" / f"{t:05d}.png"))
            assert np.array_equal(got, g[f"s_d{d}_t{t}"]), (d, t)


def test_png_round_trip_equals_the_eval_quantiser(tmp_path):
    """export -> import through 8-bit PNGs == eval.py's in-memory quantise/de-quantise (eval_utils, eval.py:19-109)."""
    g = np.load(GOLD)
    kparams, sparams = torch.from_numpy(g["kparams"]), torch.from_numpy(g["sparams"])
    compression.compress_keyframes(kparams, KCFG, str(tmp_path / "k"))
    compression.compress_sparse_grid(sparams, SCFG, str(tmp_path / "s"))
    k2, kbits = compression.load_compressed_keyframes(kparams, KCFG, str(tmp_path / "k"))
    s2, sbits = compression.load_compressed_sparse_grid(sparams, SCFG, str(tmp_path / "s"))
    assert kbits > 0 and sbits > 0
    assert torch.equal(k2.data, eval_utils.quantize_keyframes(kparams, KCFG).data)
    assert torch.equal(s2.data, eval_utils.quantize_sparse_grid(sparams, SCFG).data)
    span = float(kparams.max() - kparams.min())
    assert float((k2.data - kparams).abs().max()) <= 0.5 * span / 255 + 1e-6
    # decoded-array entry point (what skvideo.io.vread hands to the reference for the HEVC file)
    frames, _ = compression.sparse_grid_images(sparams, SCFG)
    s3, _ = compression.load_compressed_sparse_grid(sparams, SCFG, frames=[f[..., 0] for f in frames])
    assert torch.equal(s3.data, s2.data)


def test_export_model_writes_the_reference_directory_layout(tmp_path):
    g = np.load(GOLD)
    kp = torch.from_numpy(g["kparams"])
    cfg = {"2d_encoding_xy": KCFG, "2d_encoding_xt": KCFG, "2d_encoding_yt": KCFG, "3d_encoding": SCFG}
    enc = lambda: types.SimpleNamespace(params=torch.nn.Parameter(kp.clone()))     # noqa: E731
    model = types.SimpleNamespace(encoding_config=cfg, keyframes_xy=enc(), keyframes_xt=enc(), keyframes_yt=enc(),
                                  sparse_grid=types.SimpleNamespace(embeddings=torch.nn.Parameter(torch.from_numpy(g["sparams"]))))
    compression.export_model(model, str(tmp_path))
    src = tmp_path / "compression" / "src"
    for plane in ("xy", "xt", "yt"):
        assert sorted(os.listdir(src / "keyframes" / plane / "dim1")) == [f"{i:02d}.png" for i in range(5)]
    assert sorted(os.listdir(src / "sparsegrid" / "dim0")) == [f"{t:05d}.png" for t in range(4)]
