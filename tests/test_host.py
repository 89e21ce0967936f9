"""CPU tests of the host-side logic: synthetic data generator, reference initialisers,
flag defaults, feed-dict helpers, and the data-parallel gradient protocol over gloo
(world_size 2) with the oracle standing in for the device step."""
import os
import socket
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sndvae_oracle as O

data = import_module("snd-vae_b200.data")
params = import_module("snd-vae_b200.params")
flags = import_module("snd-vae_b200.flags")
prep = import_module("snd-vae_b200.preprocessing")
session = import_module("snd-vae_b200.session")


def test_synthetic_graphs_properties():
    N, B, S = 30, 5, 4
    d = data.synthetic_graphs(N, B, S, seed=3)
    A, As = d["adj_truth"], d["adj"].reshape(B, S, N, N)
    assert A.shape == (B, N, N) and d["rel"].shape == (B * S, N, N, 1) and d["features"].shape == (B * S, N, 1)
    assert np.array_equal(A, A.transpose(0, 2, 1)) and (A[:, np.arange(N), np.arange(N)] == 0).all()    # input_data.py:65-67
    for b in range(B):
        for s in range(S):
            T = As[b, s]
            assert np.array_equal(T, T.T) and (T <= A[b]).all()               # spanning forest of the truth graph
            assert T.sum() <= 2 * (N - 1)
    # aligned tiling: row b*S+s carries graph b's features / rel (SURVEY quirk Q6 fixed)
    assert np.array_equal(d["features"].reshape(B, S, N, 1)[:, 0], d["feature_truth"])
    assert np.array_equal(d["rel"].reshape(B, S, N, N)[:, 2], d["rel_truth"][..., 0])
    t = data.tile_pool(d, 12, B, S)
    assert t["adj"].shape[0] == 12 * S and t["adj_truth"].shape[0] == 12
    assert np.array_equal(t["adj_truth"][5], d["adj_truth"][0])


def test_reference_initialisers():
    cfg = O.Config(num_nodes=25)
    table = []
    off = 0
    for name, shape, _ in O.param_table(cfg):
        table.append((name, off, tuple(shape))); off += int(np.prod(shape))
    P = params.init_params(table, seed=7)
    assert set(P) == {n for n, _, _ in table}
    assert (P["encoder/g_bn_g0/gamma"] == 1).all() and (P["encoder/g_bn_g0/beta"] == 0).all()
    assert (P["decoder/e1_deconv/biases1"] == 0).all()
    w = P["decoder/e1_deconv/w1"]
    assert w.shape == (1, 25, 50, 20) and np.abs(w).max() <= 0.04 + 1e-9 and 0.015 < w.std() < 0.02   # truncated at 2 sigma
    m = P["encoder/g_sg1_lin/Matrix"]
    assert abs(m.std() - 0.02) < 1e-3 and np.abs(m).max() > 0.05
    k = P["decoder/n0_deconv/kernel"]
    lim = np.sqrt(6.0 / (5 * 40 + 5 * 50))
    assert np.abs(k).max() <= lim and np.abs(k).max() > 0.9 * lim


def test_flags_defaults_and_overrides():
    F = flags.FLAGS
    F.reset()
    assert F.sg_hidden_size == 200 and F.learning_rate == 0.001 and F.batch_size == 2 and F.sampling_num == 10   # main.py:42-103
    F.apply_dataset("synthetic2")
    assert (F.sg_hidden_size, F.sg_latent_size, F.node_h_size, F.batch_size) == (100, 100, 20, 10)              # main.py:181-217
    assert F.learning_rate == 0.0008 and F.num_edge_feature == 2
    F.apply_dataset("synthetic1")
    assert (F.sg_hidden_size, F.node_h_size, F.learning_rate) == (500, 50, 0.001)
    F.apply_dataset("protein")                                                                                 # main.py:218-236
    assert F.sg_conv_hidden == [[10, 10, 10, 10], [20, 20, 20, 20]] and (F.spatial_dim, F.node_h_size, F.batch_size) == (3, 5, 50)
    with pytest.raises(ValueError):
        F.apply_dataset("scene")
    F.reset()


def test_feed_dict_helpers():
    ph = session.make_placeholders(2, 3, 5, 1, 2)
    assert ph["adj"].shape == (6, 5, 5) and ph["rel_truth"].shape == (2, 5, 5, 1)            # main.py:253-264
    arrs = [np.zeros(1) for _ in range(8)]
    fd = prep.construct_feed_dict_train(*arrs, ph)
    assert set(k.name for k in fd) == {"features", "adj", "spatial", "rel", "adj_truth", "feature_truth", "spatial_truth",
                                       "rel_truth"}
    fd2 = prep.construct_feed_dict(*arrs[:4], ph)
    assert len(fd2) == 4


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = O.Config(num_nodes=6, sampling_num=2)
    P = O.init_params(cfg, 7, torch.float64)
    Bg = 4; Bl = Bg // world; S = cfg.S
    inp = O.synthetic_inputs(cfg, Bg, 5, torch.float64)
    noise = O.synthetic_noise(cfg, Bg, 9, torch.float64)
    sl = slice(rank * Bl, (rank + 1) * Bl); sls = slice(rank * Bl * S, (rank + 1) * Bl * S)
    si = {k: (v[sls] if k in ("adj", "features", "spatial", "rel") else v[sl]) for k, v in inp.items()}
    sn = {"eps_s": noise["eps_s"][sl], "eps_g": noise["eps_g"][sl], "eps_sg": noise["eps_sg"][sls]}
    # the device step produces LOCAL sums scaled by 1/global_batch (sndvae_grads' contract):
    # a shard-mean gradient times Bl/Bg
    g = O.loss_and_grads(P, si, sn, cfg)[4]
    names = [n for n, _, _ in O.param_table(cfg)]
    arena = torch.cat([g[n].reshape(-1) for n in names]) * (Bl / Bg)
    dist.all_reduce(arena)                                   # the one collective of the path
    if rank == 0:
        full = O.loss_and_grads(P, inp, noise, cfg)[4]
        ref = torch.cat([full[n].reshape(-1) for n in names])
        q.put(float((arena - ref).abs().max()))
    dist.destroy_process_group()


def test_data_parallel_allreduce_gloo_world2():
    """Shard the graphs over 2 ranks, all-reduce the flat gradient arena (gloo on CPU): the
    result equals the full-batch gradient (exact: no cross-graph op besides the loss means)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) < 1e-13


def _line_walk_at(cta, it, grid, lines, lines0, N, order):
    """Host restatement of LineWalk::at (snd-vae_b200/csrc/spectral.cuh): the line a persistent FFT CTA transforms at its
    it-th iteration (>= lines: past the end)."""
    if order == 0:
        return cta + it * grid
    o = 2 * (cta + (it >> 1) * grid) + (it & 1)
    if order == 1 or o >= lines:
        return o
    b, r = divmod(o, 2 * N)
    return b * N + r if r < N else lines0 + b * N + (r - N)


def test_fft_line_walk_is_a_permutation():
    """Every order of the persistent FFT grids (plain stride, pairs of neighbouring lines, graph-major pairs) visits each
    line exactly once for any grid size, odd / even N and ragged micro-batches, and a CTA's lines end with the first one
    past the end (the kernels stop there)."""
    for order in (0, 1, 2):
        for N in (7, 8, 25, 256):
            for graphs in (1, 2, 3, 5):
                lines0 = graphs * N
                lines = 2 * lines0
                for grid_sms in (1, 3, 148, 296):
                    units = (lines + 1) // 2 if order else lines
                    grid = min(grid_sms, units)
                    seen = []
                    for cta in range(grid):
                        it = 0
                        line = _line_walk_at(cta, 0, grid, lines, lines0, N, order)
                        while line < lines:
                            seen.append(line)
                            it += 1
                            line = _line_walk_at(cta, it, grid, lines, lines0, N, order)
                        # nothing valid may follow the first out-of-range position
                        assert all(_line_walk_at(cta, it + k, grid, lines, lines0, N, order) >= lines for k in range(1, 4))
                    assert sorted(seen) == list(range(lines)), (order, N, graphs, grid_sms)
    # graph-major order: the row lines and the column lines of a graph are walked within 2N consecutive positions
    N, lines0 = 8, 24
    pos = [_line_walk_at(0, it, 1, 48, lines0, N, 2) for it in range(48)]
    for b in range(3):
        blk = pos[2 * N * b: 2 * N * (b + 1)]
        assert sorted(blk) == list(range(b * N, (b + 1) * N)) + list(range(lines0 + b * N, lines0 + (b + 1) * N))


def test_bench_compulsory_bytes_of_the_spectral_stage():
    """bench.py's roofline numerator: per graph, 2N lines, each transform / GEMM reads its inputs and writes
    its outputs once; dO feeds the row and the column lines of one launch and counts once."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, C1, C2 = 256, 50, 20
    L = bench.spectral_len(N)
    assert L == 384 and L >= N + (N - 1 - (N - 1) // 2)
    F = L // 2 + 1
    lines = 2 * N
    fwd_y = lines * (4 * N * C1 + 8 * F * C1)                 # fp32 line in, bf16 hi + lo [re | im] rows out
    gemm_f = lines * (8 * F * C1 + 8 * F * C2)                # planes in, fp32 [re | im] rows out
    inv_o = lines * (8 * F * C2 + 4 * N * C2)
    fwd_do = N * 4 * N * C2 + lines * 8 * F * C2              # dO once
    gemm_d = lines * (8 * F * C2 + 8 * F * C1)
    inv_dy = lines * (8 * F * C1 + 4 * N * C1)
    wgrad = lines * (8 * F * C1 + 8 * F * C2)
    assert bench.spectral_bytes(N) == fwd_y + gemm_f + inv_o + fwd_do + gemm_d + inv_dy + wgrad


def test_make_config_selects_the_3hop_branch():
    """Four hidden sizes per SGC layer (the reference's protein / mnist configuration, main.py:225,241) select
    SpatialGraphConvolution_3D (sg_hops = 3); mixed or malformed tuples raise instead of being truncated."""
    sv = import_module("snd-vae_b200")
    cfg = sv.make_config(8, 2, "disentangled", sg_conv_hidden=((10, 10, 10, 10), (20, 20, 20, 20)))
    assert cfg.sg_hops == 3 and [list(r) for r in cfg.sg_conv_hidden3] == [[10, 10, 10, 10], [20, 20, 20, 20]]
    with pytest.raises(Exception, match="SpatialGraphConvolution"):
        sv.make_config(8, 2, "disentangled", sg_conv_hidden=((10, 10, 10, 10), (20, 20, 20)))
    cfg = sv.make_config(8, 2, "disentangled", sg_conv_hidden=((4, 5, 6), (7, 8, 9)))
    assert cfg.sg_hops == 0 and [list(r) for r in cfg.sg_conv_hidden] == [[4, 5, 6], [7, 8, 9]]

def test_session_initializer_fetch_is_a_noop():
    """main.py:301-302: `sess.run(tf.global_variables_initializer())` right after tf.Session(); the shim returns None."""
    with session.Session() as sess:
        assert sess.run(session.global_variables_initializer()) is None
        assert sess.run([session.global_variables_initializer()]) == [None]


def test_fft2p_host(tmp_path):
    """The two-pass transforms (snd-vae_b200/csrc/fft2p.cuh) are written as __host__ __device__ per-thread pieces; tests/fft2p_host.cu
    runs the kernels' phases thread by thread on the CPU against a double-precision DFT: the register DFTs (16, pruned 24 / 12), the
    paired radix-16 butterflies with the fused channel separation (every frequency <= L / 2 emitted), the Hermitian loads of the
    inverse, at L = 384 and 192 with 25 and 10 channel pairs."""
    import shutil, subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fft2p_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tests", "fft2p_host.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
