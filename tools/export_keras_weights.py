"""Run this on the TRAINER side (inside the reference's Keras environment; it imports nothing from this repo's CUDA
code): converts a reference model (model_config.json + model_weight.h5, agent/model.py:82-101) into the float32
blob ``rz_net_load_weights`` takes, written as ``model_weight.rzblob.npy`` next to the h5 file -- the file the
self-play / evaluate workers of this repo poll by digest (worker/self_play.py try_reload_model).

    python tools/export_keras_weights.py data/model/model_best_config.json data/model/model_best_weight.h5

Not exercised in this repo's tests (no Keras / h5py in the build image); the layer matching it relies on,
``reversi_zero_b200.agent.model.weights_from_keras_layers``, is (tests/test_host_logic.py).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reversi-alpha-zero_b200"))


def main(config_path, weight_path, out_path=None):
    from keras.engine.topology import Input  # noqa: F401  (same Keras generation as the reference, agent/model.py:8-14)
    from keras.engine.training import Model
    from reversi_zero_b200.agent import model as M
    with open(config_path, "rt") as f:
        model = Model.from_config(json.load(f))
    model.load_weights(weight_path)
    convs = [l for l in model.layers if l.__class__.__name__ == "Conv2D"]
    filters = sorted(c.filters for c in convs)
    dense_hidden = [l for l in model.layers if l.__class__.__name__ == "Dense" and l.name not in ("policy_out", "value_out")][0]
    mc = M.ModelConfig(cnn_filter_num=filters[-1], cnn_filter_size=max(c.kernel_size[0] for c in convs), res_layer_num=(len(convs) - 3) // 2,
                       value_fc_size=dense_hidden.units)
    w = M.weights_from_keras_layers(mc, [(l.name, l.__class__.__name__, l.get_weights()) for l in model.layers])
    out_path = out_path or os.path.splitext(weight_path)[0] + ".rzblob.npy"
    np.save(out_path, M.weights_to_blob(mc, w))
    print("wrote", out_path)


if __name__ == "__main__":
    main(*sys.argv[1:])
