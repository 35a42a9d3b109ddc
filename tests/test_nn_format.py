"""CPU: the number-format findings DESIGN.md section 5 relies on (tools/nn_format_study.py, oracle/nn.py format model) on a
4-block 256-filter tower with trained-like weights: one fp16 rounding of the operands already breaks 1e-3 on the logits,
power-of-two activation scaling changes nothing, a hi/lo split of both operands (3 MMAs per product) restores it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_format_floor_on_trained_like_weights():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nn_format_study as fs
    from nn_diag import weights_of
    from test_net_gpu import selfplay_positions
    from oracle import nn as onn
    from reversi_zero_b200.agent import model as M
    R, n = 4, 16
    planes = onn.planes_from_bitboards(*selfplay_positions(n, 5))
    planes_cal = onn.planes_from_bitboards(*selfplay_positions(96, 11))
    mc = M.ModelConfig(cnn_filter_num=256, res_layer_num=R, value_fc_size=256)
    w = weights_of("calibrated", mc, 5, planes_cal)
    ref = onn.forward_logits(w, planes, R)

    def err(mode):
        got = fs.tower(w, planes, R, mode)
        return max(float(np.abs(got[2] - ref[2]).max()), float(np.abs(got[3] - ref[3]).max())), float(np.abs(got[0] - ref[0]).max())

    fp16, p16 = err("fp16")
    scaled, _ = err("fp16-scaled")
    split3, _ = err("3-mma")
    # the format model of this module and the one the GPU tests use are the same thing
    emu = onn.forward_fp16_operands(w, planes, R)
    assert np.array_equal(emu[2], fs.tower(w, planes, R, "fp16")[2])
    assert fp16 > 1e-3                       # a single-pass fp16-operand evaluation cannot hold the logit tolerance here
    assert abs(scaled - fp16) <= 0.25 * fp16  # relative rounding: scaling into the fp16 range does not help
    assert split3 < 1e-4                     # both operands as hi + lo: three tensor-core passes per product
    assert p16 < 1e-3                        # what MCTS consumes (the probabilities) stays within 1e-3 even so
