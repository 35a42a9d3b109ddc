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


def weights_from_keras_layers(mc, layers):
    """Weight hand-off from the reference's Keras model without relying on ``model.get_weights()``
    order (Keras sorts the layers of the two heads by graph depth): ``layers`` is an iterable of
    ``(layer_name, class_name, [arrays])`` -- e.g. ``[(l.name, l.__class__.__name__, l.get_weights()) for l in
    model.layers]`` -- in any order.  Layers are matched by CREATION order, which Keras encodes in the numeric suffix of
    its automatic names (``conv2d_7``, ``batch_normalization_7``, ``dense_1``; the counter may start anywhere when
    several models were built in the process), following ``ReversiModel.build`` (agent/model.py:28-72): conv0, the
    residual blocks' convolutions, the policy-head 1x1 convolution, the value-head 1x1 convolution; every convolution is
    followed by its BatchNormalization; ``policy_out`` / ``value_out`` are named, the remaining Dense layer is the
    value head's hidden layer.  Returns the ``{name: ndarray}`` dict ``weights_to_blob`` packs."""
    def suffix(name):
        tail = name.rsplit("_", 1)[-1]
        return int(tail) if tail.isdigit() else 0

    convs, bns, dense, named = [], [], [], {}
    for name, cls, arrays in layers:
        arrays = [np.asarray(a, dtype=np.float32) for a in arrays]
        if name in ("policy_out", "value_out"):
            named[name] = arrays
        elif cls == "Conv2D":
            convs.append((suffix(name), arrays))
        elif cls == "BatchNormalization":
            bns.append((suffix(name), arrays))
        elif cls == "Dense":
            dense.append((suffix(name), arrays))
    convs.sort(key=lambda t: t[0])
    bns.sort(key=lambda t: t[0])
    n_conv = 1 + 2 * mc.res_layer_num + 2
    if len(convs) != n_conv or len(bns) != n_conv or len(dense) != 1 or set(named) != {"policy_out", "value_out"}:
        raise ValueError(f"expected {n_conv} Conv2D + {n_conv} BatchNormalization + 1 hidden Dense + policy_out + value_out layers, got "
                         f"{len(convs)} / {len(bns)} / {len(dense)} / {sorted(named)}")
    prefixes = ["conv0"] + [f"res{i}.conv{j}" for i in range(mc.res_layer_num) for j in (1, 2)] + ["policy_conv", "value_conv"]
    w = {}
    for prefix, (_, conv), (_, bn) in zip(prefixes, convs, bns):
        if len(conv) != 2 or len(bn) != 4:
            raise ValueError(f"{prefix}: expected [kernel, bias] and [gamma, beta, moving_mean, moving_variance]")
        w[f"{prefix}.kernel"], w[f"{prefix}.bias"] = conv
        w[f"{prefix}.bn_gamma"], w[f"{prefix}.bn_beta"], w[f"{prefix}.bn_mean"], w[f"{prefix}.bn_var"] = bn
    w["policy_fc.kernel"], w["policy_fc.bias"] = named["policy_out"]
    w["value_fc1.kernel"], w["value_fc1.bias"] = dense[0][1]
    w["value_fc2.kernel"], w["value_fc2.bias"] = named["value_out"]
    for name, shape in tensor_specs(mc):
        if w[name].shape != tuple(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {w[name].shape}")
    return w
