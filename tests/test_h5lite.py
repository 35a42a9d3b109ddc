"""tools/h5lite.py (a stand-alone conversion tool, NOT on the product path; PARITY UNPINNED -- see its header): the reader against files produced by the test
writer of the same classic HDF5 structures, laid out like a Keras ``save_weights`` file of the reference's network
(agent/model.py:28-72,95), and the whole chain h5 -> layers -> blob."""
import numpy as np
import pytest

from reversi_zero_b200.agent import model as M
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "support"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import h5lite  # noqa: E402
import h5_write  # noqa: E402


def keras_style_tree(mc, w, offset=1):
    """{layer: {layer: {"kernel:0": ...}}} with Keras' automatic layer names (creation order = numeric suffix)"""
    prefixes = ["conv0"] + [f"res{i}.conv{j}" for i in range(mc.res_layer_num) for j in (1, 2)] + ["policy_conv", "value_conv"]
    tree = {}
    for i, p in enumerate(prefixes):
        c, b = f"conv2d_{offset + i}", f"batch_normalization_{offset + i}"
        tree[c] = {c: {"kernel:0": w[f"{p}.kernel"], "bias:0": w[f"{p}.bias"]}}
        tree[b] = {b: {"gamma:0": w[f"{p}.bn_gamma"], "beta:0": w[f"{p}.bn_beta"], "moving_mean:0": w[f"{p}.bn_mean"],
                       "moving_variance:0": w[f"{p}.bn_var"]}}
        tree[f"activation_{offset + i}"] = {}                       # weightless layers are empty groups
    tree["dense_1"] = {"dense_1": {"kernel:0": w["value_fc1.kernel"], "bias:0": w["value_fc1.bias"]}}
    tree["policy_out"] = {"policy_out": {"kernel:0": w["policy_fc.kernel"], "bias:0": w["policy_fc.bias"]}}
    tree["value_out"] = {"value_out": {"kernel:0": w["value_fc2.kernel"], "bias:0": w["value_fc2.bias"]}}
    tree["input_1"] = {}
    tree["flatten_1"] = {}
    return tree


@pytest.mark.parametrize("mc", [M.ModelConfig(cnn_filter_num=16, res_layer_num=1, value_fc_size=16), M.ModelConfig(cnn_filter_num=32, res_layer_num=10, value_fc_size=64)])
def test_h5_to_blob_roundtrip(tmp_path, mc):
    w = M.build_random_weights(mc, seed=4, perturb_bn=True)
    path = str(tmp_path / "model_weight.h5")
    h5_write.write(path, keras_style_tree(mc, w, offset=12))
    ds = h5lite.read_datasets(path)
    assert "conv2d_12/conv2d_12/kernel:0" in ds and ds["conv2d_12/conv2d_12/kernel:0"].shape == (3, 3, 2, mc.cnn_filter_num)
    assert len(ds) == (3 + 2 * mc.res_layer_num) * 6 + 6            # conv (2) + BN (4) per convolution, three dense layers (2 each)
    got = M.weights_from_keras_layers(mc, h5lite.keras_layers_from_h5(ds))
    assert np.array_equal(M.weights_to_blob(mc, got), M.weights_to_blob(mc, w))


def test_rejects_what_it_does_not_understand(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"definitely not hdf5")
    with pytest.raises(h5lite.H5FormatError):
        h5lite.read_datasets(str(p))
    h5_write.write(str(p), {"g": {"d:0": np.arange(6, dtype=np.float32).reshape(2, 3)}})
    raw = bytearray(p.read_bytes())
    assert np.array_equal(h5lite.read_datasets(str(p))["g/d:0"], np.arange(6, dtype=np.float32).reshape(2, 3))
    raw[8] = 2                                                      # a superblock version the reader does not implement
    p.write_bytes(bytes(raw))
    with pytest.raises(h5lite.H5FormatError):
        h5lite.read_datasets(str(p))


def test_worker_refuses_a_model_dir_with_only_h5_files(tmp_path):
    """load_or_build_weights (agent/api.py:102-115): the product path takes exported blobs only; with the reference's
    model_best_weight.h5 in place but no blob it must stop with the exporter hint instead of silently random-initialising
    (or reading HDF5 with an unpinned parser) -- start-up and the 60 s hot reload look at the same files."""
    from reversi_zero_b200.config import Config
    from reversi_zero_b200.worker import self_play as sp
    cfg = Config(project_dir=str(tmp_path), data_dir=str(tmp_path / "data"))
    cfg.model.update(dict(cnn_filter_num=16, res_layer_num=1, value_fc_size=16))
    cfg.resource.create_directories()
    w = M.build_random_weights(cfg.model, seed=8, perturb_bn=True)
    h5_write.write(cfg.resource.model_best_weight_path, keras_style_tree(cfg.model, w))

    class FakeNet:
        def load_blob(self, blob):
            self.blob = np.array(blob)
    net = FakeNet()
    assert sp.weight_source_path(cfg) is None and sp.keras_h5_source_path(cfg) == cfg.resource.model_best_weight_path
    with pytest.raises(RuntimeError, match="export_keras_weights"):
        sp.load_or_build_weights(cfg, net)
    assert not os.path.exists(sp.blob_path_of(cfg))                 # nothing was invented or saved as "best"
    # the stand-alone tool converts the file; with the blob next to it the worker starts from it
    np.save(sp.blob_path_of(cfg), h5lite.blob_from_keras_h5(cfg.model, cfg.resource.model_best_weight_path))
    sp.load_or_build_weights(cfg, net)
    assert np.array_equal(net.blob, M.weights_to_blob(cfg.model, w))
