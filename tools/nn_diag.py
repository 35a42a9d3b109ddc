"""Diagnostic for the tcgen05 tower: compares the fp32 tower output with the torch oracle for 0/1/2/10
residual blocks and prints where the error sits (by pixel row/column, channel group, board parity)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reversi-alpha-zero_b200"))


def main():
    import torch
    from oracle import nn as onn
    from reversi_zero_b200.agent import model as M
    from reversi_zero_b200 import net as N, device as D
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_net_gpu import selfplay_positions
    for R in (0, 1, 2, 10):
        mc = M.ModelConfig(cnn_filter_num=256, res_layer_num=R, value_fc_size=256)
        w = M.build_random_weights(mc, 2, perturb_bn=True)
        net = N.Net(mc)
        net.load_weights(w)
        n = 7
        own, enemy = selfplay_positions(n, 2)
        planes = onn.planes_from_bitboards(own, enemy)
        p_ref, v_ref, _, tower_ref = onn.forward(w, planes, R, return_tower=True)
        d_own, d_en = D.to_device(own), D.to_device(enemy)
        d_pol, d_val, d_tow = D.empty(n * 64, np.float32), D.empty(n, np.float32), D.empty(n * 64 * 256, np.float32)
        try:
            net.debug_tower_dev(d_own, d_en, d_pol, d_val, d_tow, n)
            torch.cuda.synchronize()
        except Exception as ex:
            print(f"R={R}: FAILED {ex}")
            return
        tower = d_tow.cpu().numpy().reshape(n, 64, 256).transpose(0, 2, 1).reshape(n, 256, 8, 8)
        err = np.abs(tower - tower_ref)
        p = d_pol.cpu().numpy().reshape(n, 64); v = d_val.cpu().numpy()
        print(f"R={R}: tower max err {err.max():.4g} (ref absmax {np.abs(tower_ref).max():.4g}, mean {np.abs(tower_ref).mean():.4g}); "
              f"policy err {np.abs(p - p_ref).max():.3g} value err {np.abs(v - v_ref).max():.3g}")
        if err.max() > 1e-2:
            print("  err by board :", np.round(err.max(axis=(1, 2, 3)), 3))
            print("  err by y     :", np.round(err.max(axis=(0, 1, 3)), 3))
            print("  err by x     :", np.round(err.max(axis=(0, 1, 2)), 3))
            print("  err by ch/32 :", np.round(err.reshape(n, 8, 32, 8, 8).max(axis=(0, 2, 3, 4)), 3))
            print("  sample got/ref:", tower[0, :4, 3, 3], tower_ref[0, :4, 3, 3])
            print("  frac wrong   :", (err > 1e-2).mean())
        net.close()


if __name__ == "__main__":
    main()
