"""Development check of the CTA-pair tower kernel (csrc/rz_net_tc2.cu) against the single-CTA kernel (csrc/rz_net_tc.cu)
and the fp32 oracle: numerics on towers of 0 / 1 / 2 / 10 blocks and ragged batch sizes, then timing of both kernels on
32 768 positions.  Usage: python tools/tower_v2_check.py [check|time|all] [iters]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reversi-alpha-zero_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def heads(net, own, enemy):
    import torch
    from reversi_zero_b200 import device as D
    n = own.size
    d_own, d_en = D.to_device(own), D.to_device(enemy)
    d_pol, d_val, d_log, d_vl = D.empty(n * 64, np.float32), D.empty(n, np.float32), D.empty(n * 64, np.float32), D.empty(n, np.float32)
    d_tow = D.empty(n * 64 * 256, np.float32)
    net.debug_heads_dev(d_own, d_en, d_pol, d_val, d_log, d_vl, n, tower_t=d_tow)
    torch.cuda.synchronize()
    return dict(policy=d_pol.cpu().numpy().reshape(n, 64), value=d_val.cpu().numpy(), logits=d_log.cpu().numpy().reshape(n, 64),
                vlogit=d_vl.cpu().numpy(), tower=d_tow.cpu().numpy().reshape(n, 64, 256).transpose(0, 2, 1).reshape(n, 256, 8, 8))


def check():
    from oracle import nn as onn
    from reversi_zero_b200.agent import model as M
    from reversi_zero_b200 import net as N
    from test_net_gpu import selfplay_positions
    ok = True
    for R, n in ((0, 7), (1, 7), (2, 9), (10, 1), (10, 2), (10, 3), (10, 301), (10, 1000)):
        mc = M.ModelConfig(cnn_filter_num=256, res_layer_num=R, value_fc_size=256)
        w = M.build_random_weights(mc, 2, perturb_bn=R < 10)
        own, enemy = selfplay_positions(n, 2)
        planes = onn.planes_from_bitboards(own, enemy)
        net = N.Net(mc)
        net.load_weights(w)
        N.set_tower_kernel(1)
        a = heads(net, own, enemy)
        N.set_tower_kernel(2)
        b = heads(net, own, enemy)
        b2 = heads(net, own, enemy)
        ref = dict(zip(("policy", "value", "logits", "vlogit", "tower"), onn.forward_logits(w, planes, R))) if n <= 301 else None
        row = dict(R=R, n=n, deterministic=all(np.array_equal(b[k], b2[k]) for k in b))
        for k in ("tower", "logits", "vlogit", "policy", "value"):
            row[f"{k}_v2_vs_v1"] = float(np.abs(a[k] - b[k]).max())
            if ref is not None:
                row[f"{k}_v2_vs_fp32"] = float(np.abs(b[k] - ref[k]).max())
                row[f"{k}_v1_vs_fp32"] = float(np.abs(a[k] - ref[k]).max())
        row["finite"] = bool(all(np.isfinite(b[k]).all() for k in b))
        print(json.dumps(row), flush=True)
        ok = ok and row["finite"] and row["deterministic"] and row["logits_v2_vs_v1"] < 2e-3
        net.close()
    print("CHECK", "OK" if ok else "FAILED", flush=True)
    return ok


def timing(iters):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import nn_bench
    from reversi_zero_b200 import net as N
    for v in (1, 2, 1, 2):
        N.set_tower_kernel(v)
        r = nn_bench.run(32768, iters=iters, warmup=3)
        r["tower_kernel"] = v
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    good = check() if what in ("check", "all") else True
    if what in ("time", "all") and good:
        timing(iters)
