"""CPU tests of the host-side mirror of modules.NVP (names, shapes, init stream, loud failure)."""
import contextlib
import io

import pytest
import torch

import nvp_b200
from oracle import check_vs_reference as R
from oracle import nvp_oracle as O

A1_KEYS = [
    "keyframes_xy.params", "keyframes_yt.params", "keyframes_xt.params", "sparse_grid.embeddings",
    "net.layers.0.weight", "net.layers.0.bias", "net.layers.1.weight", "net.layers.1.bias",
    "net.layers.2.weight", "net.layers.2.bias", "net.last_layer.weight", "net.last_layer.bias",
]
A1_KEYS += ["wrapper." + k for k in A1_KEYS if k.startswith("net.")]
A1_KEYS += [f"wrapper.modulator.layers.{i}.0.{w}" for i in range(3) for w in ("weight", "bias")]


def small_json(F=2):
    return O.NVPConfig(n_features=F, sparse_features=F, t_resolution=6, x_resolution=20, y_resolution=24).to_json()


@pytest.mark.parametrize("F", [2, 4])
def test_state_dict_keys_and_shapes(F):
    m = nvp_b200.NVP(type="nvp", in_features=2, out_features=3, encoding_config=small_json(F))
    sd = m.state_dict()
    assert list(sd.keys()) == A1_KEYS  # SURVEY A.1 order
    Z = 57 * F
    assert sd["keyframes_xy.params"].shape == (4616112 * F,)
    assert sd["sparse_grid.embeddings"].shape == (6, 20, 24, F)
    assert sd["wrapper.modulator.layers.0.0.weight"].shape == (128, Z)
    assert sd["wrapper.modulator.layers.2.0.weight"].shape == (128, 128 + Z)
    assert sd["net.layers.0.weight"].shape == (128, 1) and sd["net.last_layer.weight"].shape == (3, 128)
    assert m.wrapper.net is m.net
    assert m.keyframes_xy.dtype == torch.float32 and m.sparse_grid.level_dim == F
    assert len(list(m.parameters())) == 18
    # eval.py:170-179 assigns fresh Parameters onto these attributes
    m.keyframes_xy.params = torch.nn.Parameter(torch.zeros_like(m.keyframes_xy.params))
    m.sparse_grid.embeddings = torch.nn.Parameter(torch.zeros_like(m.sparse_grid.embeddings))
    assert m.hot_path_parameters()[0] is m.keyframes_xy.params


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_same_seed_gives_reference_initial_weights():
    ref = R.import_reference()
    cfg = small_json(2)
    torch.manual_seed(123)
    with contextlib.redirect_stdout(io.StringIO()):
        a = ref.modules.NVP(type="nvp", out_features=3, encoding_config=cfg)
    torch.manual_seed(123)
    b = nvp_b200.NVP(type="nvp", out_features=3, encoding_config=cfg)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape, k
        if not k.startswith("keyframes_"):  # tcnn's own RNG (stubbed) vs ours: layout only
            assert torch.equal(sa[k], sb[k]), k


def test_cpu_tensors_fail_loudly():
    m = nvp_b200.NVP(out_features=3, encoding_config=small_json())
    x = {"all_coords": torch.rand(1, 8, 3), "temporal_steps": torch.rand(1, 8)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(x, temporal_interp=True)
    with pytest.raises(ValueError):
        nvp_b200.NVP(out_features=3, encoding_config=small_json(), mode="triton")


def test_padded_tcnn_layout_is_remapped_or_rejected():
    """a checkpoint whose keyframe params use upstream tiny-cuda-nn's padded level layout (every level rounded
    up to 8 entries) is remapped on load; any other size gets an error naming both layouts.  (The remap itself is
    'parity unpinned': no tcnn build exists here to produce such a checkpoint.)"""
    import pytest
    from nvp_b200.encoding import Encoding
    cfg = {"otype": "DenseGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 24, "base_resolution": 16,
           "per_level_scale": 1.35}
    enc = Encoding(2, cfg)
    ref = enc.params.detach().clone()
    padded = enc.padded_level_offsets()
    assert padded[2] - padded[1] == 488 and enc.level_res[1] ** 2 == 484   # level 1: 22^2 = 484 -> 488
    big = torch.full((padded[-1], 2), 7.0)
    for l in range(16):
        n = enc.level_res[l] ** 2
        big[padded[l]: padded[l] + n] = ref.reshape(-1, 2)[enc.level_offsets[l]: enc.level_offsets[l] + n]
    enc2 = Encoding(2, cfg, seed=1)
    enc2.load_state_dict({"params": big.reshape(-1)})
    assert torch.equal(enc2.params.detach(), ref)
    with pytest.raises(RuntimeError, match="padded"):
        enc2.load_state_dict({"params": torch.zeros(123)})
