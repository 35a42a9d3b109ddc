"""GPU parity tests for the network kernels through the C ABI against the fp32 torch oracle
(oracle/nn.py).  Tolerance: 1e-3 max-abs on softmax probabilities and tanh value for the fp16/fp32-acc
tcgen05 tower (BASELINE.json north_star: "within 1e-3"); 2e-5 for the fp32 generic kernel."""
import numpy as np
import pytest

from oracle import nn as onn
from oracle import bitboard as ob
from reversi_zero_b200.agent import model as M
from reversi_zero_b200.agent.api import ReversiModelAPI
from reversi_zero_b200 import net as N

pytestmark = pytest.mark.gpu


def selfplay_positions(n, seed=3):
    """positions from random playouts (side-to-move frame), a few of them dihedral-transformed."""
    rng = np.random.default_rng(seed)
    own, enemy = [], []
    while len(own) < n:
        e = ob.Env().reset()
        while not e.done and len(own) < n:
            o, en = e.own_enemy()
            t = int(rng.integers(8))
            own.append(ob.dihedral(o, t)); enemy.append(ob.dihedral(en, t))
            legal = ob.find_correct_moves(o, en)
            ms = [i for i in range(64) if legal >> i & 1]
            e.step(ms[rng.integers(len(ms))])
    return np.array(own, np.uint64), np.array(enemy, np.uint64)


def run_case(mc, n, impl, seed, perturb, tol, check_tower=False):
    import torch
    from reversi_zero_b200 import device as D
    w = M.build_random_weights(mc, seed, perturb_bn=perturb)
    net = N.Net(mc)
    net.load_weights(w)
    own, enemy = selfplay_positions(n, seed)
    planes = onn.planes_from_bitboards(own, enemy)
    p_ref, v_ref, logits_ref, tower_ref = onn.forward(w, planes, mc.res_layer_num, return_tower=True)
    d_own, d_en = D.to_device(own), D.to_device(enemy)
    d_pol, d_val = D.empty(n * 64, np.float32), D.empty(n, np.float32)
    if check_tower:
        d_tow = D.empty(n * 64 * 256, np.float32)
        net.debug_tower_dev(d_own, d_en, d_pol, d_val, d_tow, n)
        torch.cuda.synchronize()
        tower = d_tow.cpu().numpy().reshape(n, 64, 256).transpose(0, 2, 1).reshape(n, 256, 8, 8)
        terr = np.abs(tower - tower_ref).max()
        scale = np.abs(tower_ref).max()
        assert terr <= 4e-3 * max(scale, 1.0), f"tower max-abs err {terr} (scale {scale})"
    else:
        net.predict_dev(d_own, d_en, d_pol, d_val, n, impl)
        torch.cuda.synchronize()
    p = d_pol.cpu().numpy().reshape(n, 64)
    v = d_val.cpu().numpy()
    assert np.isfinite(p).all() and np.isfinite(v).all()
    perr, verr = np.abs(p - p_ref).max(), np.abs(v - v_ref).max()
    assert perr <= tol and verr <= tol, f"policy err {perr}, value err {verr} (tol {tol})"
    # host-buffer path (ReversiModelAPI.predict contract, agent/api.py:30-45)
    api = ReversiModelAPI(None, net, impl)
    k = min(5, n)
    p2, v2 = api.predict(planes[:k])
    assert p2.shape == (k, 64) and v2.shape == (k, 1)
    assert np.abs(p2 - p_ref[:k]).max() <= tol and np.abs(v2[:, 0] - v_ref[:k]).max() <= tol
    p1, v1 = api.predict(planes[0])
    assert p1.shape == (64,) and v1.shape == (1,)
    net.close()
    return perr, verr


MINI = dict(cnn_filter_num=16, res_layer_num=1, value_fc_size=16)
CH5 = dict(cnn_filter_num=256, res_layer_num=10, value_fc_size=256)


@pytest.mark.parametrize("cfg,n", [(MINI, 33), (dict(cnn_filter_num=32, res_layer_num=2, value_fc_size=64), 9),
                                   (dict(cnn_filter_num=256, res_layer_num=1, value_fc_size=256), 5)])
def test_generic_kernel_vs_oracle(cfg, n):
    run_case(M.ModelConfig(**cfg), n, N.IMPL_GENERIC, seed=1, perturb=True, tol=2e-5)


@pytest.mark.parametrize("res_blocks", [0, 1, 2])
def test_tcgen05_small_towers(res_blocks):
    """0 blocks isolates the layer-0 im2col GEMM + heads; 1-2 blocks the shifted-operand convs and the
    TMEM residual."""
    mc = M.ModelConfig(cnn_filter_num=256, res_layer_num=res_blocks, value_fc_size=256)
    run_case(mc, 7, N.IMPL_TCGEN05, seed=2, perturb=True, tol=1e-3, check_tower=True)
    run_case(mc, 7, N.IMPL_TCGEN05, seed=2, perturb=True, tol=1e-3)


@pytest.mark.parametrize("n", [1, 2, 301, 1000])
def test_tcgen05_ch5_vs_oracle(n):
    """ch5 network (256 filters x 10 blocks), random-init as `--new` builds it; n = 1000 > 2 x 148 SMs so
    every CTA processes several tiles."""
    run_case(M.ModelConfig(**CH5), n, N.IMPL_TCGEN05, seed=0, perturb=False, tol=1e-3)


def test_tcgen05_deep_tower_config4():
    """BASELINE config 4: the same kernel with a deeper tower (19 residual blocks = 39 convolutions, the AlphaGo Zero
    depth), random-init as `--new` builds it; same 1e-3 bound on policy probabilities and value."""
    run_case(M.ModelConfig(cnn_filter_num=256, res_layer_num=19, value_fc_size=256), 96, N.IMPL_TCGEN05, seed=0, perturb=False, tol=1e-3)


def test_tcgen05_ch5_perturbed_bn_and_value_fc():
    """stress case, NOT the north-star configuration: random biases and BN statistics (gamma 0.5-1.5, var 0.5-2)
    compound over 21 layers into activations ~10x larger than with `--new` weights, which amplifies the fp16
    operand rounding; measured 1.4e-3 on the value, so the bound here is 2.5e-3 (the 1e-3 bound is asserted on the
    ch5 random-init network above, measured 3e-4)."""
    run_case(M.ModelConfig(cnn_filter_num=256, res_layer_num=10, value_fc_size=128), 64, N.IMPL_TCGEN05, seed=5, perturb=True, tol=2.5e-3)


def test_tcgen05_matches_generic_on_device():
    """the two CUDA implementations agree with each other (cross-check independent of torch)."""
    import torch
    from reversi_zero_b200 import device as D
    mc = M.ModelConfig(**CH5)
    net = N.Net(mc)
    net.load_weights(M.build_random_weights(mc, 9))
    own, enemy = selfplay_positions(40, 9)
    d_own, d_en = D.to_device(own), D.to_device(enemy)
    outs = []
    for impl in (N.IMPL_GENERIC, N.IMPL_TCGEN05):
        d_pol, d_val = D.empty(40 * 64, np.float32), D.empty(40, np.float32)
        net.predict_dev(d_own, d_en, d_pol, d_val, 40, impl)
        torch.cuda.synchronize()
        outs.append((d_pol.cpu().numpy(), d_val.cpu().numpy()))
    assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-3 and np.abs(outs[0][1] - outs[1][1]).max() <= 1e-3
