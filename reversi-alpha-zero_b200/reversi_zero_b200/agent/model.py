"""Policy/value network description + weights container (host side, numpy only).

Mirrors the *architecture* of the reference's ``ReversiModel`` (agent/model.py:28-72) so the same
``ModelConfig`` (config.py:187-193) drives it; the forward pass itself runs in the CUDA engine
(csrc/rz_net_*.cu).  Weights are held in Keras layouts (Conv2D kernel ``(kh, kw, Cin, Cout)``,
Dense kernel ``(in, out)``) and flattened into the float32 blob ``rz_net_load_weights`` documents
(include/rz_engine.h).  Training / h5 save-load stay in the reference's Keras ``opt`` worker.
"""
import hashlib

import numpy as np

BN_EPS = 1e-3  # Keras BatchNormalization default epsilon (agent/model.py:35 uses the default)


class ModelConfig:
    """Same fields/defaults as the reference's ModelConfig (config.py:187-193)."""

    def __init__(self, cnn_filter_num=256, cnn_filter_size=3, res_layer_num=10, l2_reg=1e-4, value_fc_size=256):
        self.cnn_filter_num = cnn_filter_num
        self.cnn_filter_size = cnn_filter_size
        self.res_layer_num = res_layer_num
        self.l2_reg = l2_reg
        self.value_fc_size = value_fc_size


def tensor_specs(mc):
    """Ordered (name, shape) list of every tensor of the network == the blob layout."""
    F, V, ks = mc.cnn_filter_num, mc.value_fc_size, mc.cnn_filter_size
    specs = []

    def conv_bn(name, cin, cout, k):
        specs.append((f"{name}.kernel", (k, k, cin, cout)))
        specs.append((f"{name}.bias", (cout,)))
        for p in ("gamma", "beta", "mean", "var"):
            specs.append((f"{name}.bn_{p}", (cout,)))

    conv_bn("conv0", 2, F, ks)
    for i in range(mc.res_layer_num):
        conv_bn(f"res{i}.conv1", F, F, ks)
        conv_bn(f"res{i}.conv2", F, F, ks)
    conv_bn("policy_conv", F, 2, 1)
    specs.append(("policy_fc.kernel", (128, 64)))
    specs.append(("policy_fc.bias", (64,)))
    conv_bn("value_conv", F, 1, 1)
    specs.append(("value_fc1.kernel", (64, V)))
    specs.append(("value_fc1.bias", (V,)))
    specs.append(("value_fc2.kernel", (V, 1)))
    specs.append(("value_fc2.bias", (1,)))
    return specs


def blob_size(mc):
    return int(sum(int(np.prod(s)) for _, s in tensor_specs(mc)))


def build_random_weights(mc, seed=0, perturb_bn=False):
    """What ``ReversiModel.build()`` + ``--new`` produces (agent/api.py:112-114): glorot_uniform
    kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1.  ``perturb_bn`` randomises biases and BN
    statistics instead (used by tests so that BN folding is actually exercised)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in tensor_specs(mc):
        if name.endswith(".kernel"):
            if len(shape) == 4:
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            else:
                fan_in, fan_out = shape
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif perturb_bn:
            if name.endswith("bn_gamma"):
                w[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
            elif name.endswith("bn_var"):
                w[name] = rng.uniform(0.5, 2.0, size=shape).astype(np.float32)
            else:
                w[name] = rng.uniform(-0.2, 0.2, size=shape).astype(np.float32)
        else:
            fill = 1.0 if (name.endswith("bn_gamma") or name.endswith("bn_var")) else 0.0
            w[name] = np.full(shape, fill, dtype=np.float32)
    return w


def weights_to_blob(mc, w):
    parts = []
    for name, shape in tensor_specs(mc):
        a = np.asarray(w[name], dtype=np.float32)
        if a.shape != tuple(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def blob_to_weights(mc, blob):
    blob = np.asarray(blob, dtype=np.float32).reshape(-1)
    w, off = {}, 0
    for name, shape in tensor_specs(mc):
        n = int(np.prod(shape))
        w[name] = blob[off:off + n].reshape(shape).copy()
        off += n
    if off != blob.size:
        raise ValueError(f"blob has {blob.size} floats, expected {off}")
    return w


def blob_digest(blob):
    """sha256 of the blob; plays the role of ReversiModel.fetch_digest (agent/model.py:74-80)."""
    return hashlib.sha256(np.ascontiguousarray(blob, dtype=np.float32).tobytes()).hexdigest()
