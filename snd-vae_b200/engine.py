"""Engine: the device-side state behind the reference-facing shims.

Thin host layer over the C ABI: owns a handle, moves feeds, returns torch
tensors.  PyTorch is used for device memory, streams and torch.distributed only;
all arithmetic runs in libsndvae.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

FEED_KEYS = ("features", "spatial", "adj", "rel", "adj_truth", "feature_truth", "spatial_truth", "rel_truth")
NOISE_KEYS = ("eps_s", "eps_sg", "eps_g")


class SndvaeError(RuntimeError):
    """Raised where TensorFlow would raise InvalidArgumentError / a runtime error."""


class _ArenaView:
    """Zero-copy torch view of a library-owned device arena (CUDA array interface)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def make_config(num_nodes: int, batch_size: int, model_type: str = "disentangled", **kw) -> _lib.Config:
    lib = _lib.load()
    cfg = _lib.Config()
    lib.sndvae_default_config(C.byref(cfg))
    cfg.num_nodes, cfg.batch_size = num_nodes, batch_size
    cfg.model_type = 1 if model_type == "base" else 0
    for k, v in kw.items():
        cur = getattr(cfg, k)
        if hasattr(cur, "__len__"):
            if k == "sg_conv_hidden":
                if len(v) == 2 and all(len(hs) == 4 for hs in v):
                    # 4 sizes per layer = SpatialGraphConvolution_3D (layers.py:200-277; FLAGS.dataset protein / mnist)
                    cfg.sg_hops = 3
                    for i in range(2):
                        for j in range(4):
                            cfg.sg_conv_hidden3[i][j] = int(v[i][j])
                    continue
                if len(v) != 2 or any(len(hs) != 3 for hs in v):
                    raise SndvaeError(f"sg_conv_hidden={v!r}: two layers of three (SpatialGraphConvolution) or four "
                                      "(SpatialGraphConvolution_3D) hidden sizes required")
                for i in range(2):
                    for j in range(3):
                        cur[i][j] = int(v[i][j])
            else:
                for i, x in enumerate(v):
                    cur[i] = int(x)
        else:
            setattr(cfg, k, v)
    if cfg.model_type == 1:
        cfg.sampling_num = 1
    return cfg


class Engine:
    def __init__(self, cfg: _lib.Config, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise SndvaeError("no CUDA device: the SND-VAE hot path is hand-written CUDA for sm_100a and has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        self.cfg = cfg
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.sndvae_create(C.byref(cfg), C.c_void_p(stream), C.byref(self._h))
        if rc != 0:
            msg = self.lib.sndvae_last_error(self._h).decode() if self._h else "create failed"
            if self._h:
                self.lib.sndvae_destroy(self._h)
            self._h = C.c_void_p()
            raise SndvaeError(f"sndvae_create: {msg} (code {rc})")
        self.dis = cfg.model_type == 0
        self.N, self.F, self.D = cfg.num_nodes, cfg.num_feature, cfg.spatial_dim
        self.S, self.B = cfg.sampling_num, cfg.batch_size
        self.nparam = int(self.lib.sndvae_param_count(self._h))
        n = int(self.lib.sndvae_num_params(self._h))
        tab = (_lib.ParamInfo * n)()
        self._check(self.lib.sndvae_param_table(self._h, tab, n))
        self.table = [(t.name.decode(), int(t.offset), tuple(int(t.shape[i]) for i in range(t.rank))) for t in tab]
        self._keep = None

    # -- plumbing -----------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise SndvaeError(f"{self.lib.sndvae_last_error(self._h).decode()} (code {rc})")

    def close(self):
        if self._h:
            self.lib.sndvae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape_of(self, key):
        B, S, N, F, D = self.B, self.S, self.N, self.F, self.D
        c = self.cfg
        return {
            "features": (B * S, N, F), "spatial": (B * S, N, D), "adj": (B * S, N, N), "rel": (B * S, N, N, 1),
            "adj_truth": (B, N, N), "feature_truth": (B, N, F), "spatial_truth": (B, N, D), "rel_truth": (B, N, N, 1),
            "eps_s": (B, c.s_latent_size), "eps_sg": (B * S, c.sg_latent_size), "eps_g": (B, c.g_latent_size),
            "z_mean_s": (B, c.s_latent_size), "z_std_s": (B, c.s_latent_size), "z_s": (B, c.s_latent_size),
            "z_mean_g": (B, c.g_latent_size), "z_std_g": (B, c.g_latent_size), "z_g": (B, c.g_latent_size),
            "z_mean_sg": (B * S, c.sg_latent_size), "z_std_sg": (B * S, c.sg_latent_size), "z_sg": (B * S, c.sg_latent_size),
            "generated_adj": (B, N, N), "generated_adj_prob": (B, N, N, 2),
            "generated_spatial": (B, N, D), "generated_node_feat": (B, N, F),
        }[key]

    def _dev(self, key, t):
        """Validate one feed: static shape as in main.py:253-264 (TF raises on mismatch)."""
        if t is None:
            return None
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t))
        want = self._shape_of(key)
        if tuple(t.shape) != want and not (key in ("rel", "rel_truth") and tuple(t.shape) == want[:-1]):
            raise SndvaeError(f"feed '{key}' has shape {tuple(t.shape)}, expected {want}")
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def _pack(self, feeds, noise):
        keep = []
        inp = _lib.Inputs()
        for k in FEED_KEYS:
            t = self._dev(k, feeds.get(k))
            if t is not None:
                keep.append(t)
                setattr(inp, k, t.data_ptr())
        nz = _lib.Noise()
        for k in NOISE_KEYS:
            t = self._dev(k, noise.get(k)) if (self.dis or k == "eps_sg") else None
            if t is not None:
                keep.append(t)
                setattr(nz, k, t.data_ptr())
        return inp, nz, keep

    def _outs(self, fetch):
        out = _lib.Outputs()
        res = {}
        for k in fetch:
            if k not in _lib.OUTPUT_FIELDS:
                raise SndvaeError(f"unknown fetch '{k}'")
            if not self.dis and k in ("z_mean_s", "z_std_s", "z_s", "z_mean_g", "z_std_g", "z_g"):
                continue
            dt = torch.int64 if k == "generated_adj" else torch.float32
            res[k] = torch.empty(self._shape_of(k), dtype=dt, device=self.device)
            setattr(out, k, res[k].data_ptr())
        return out, res

    @property
    def n_losses(self):
        return 7 if self.dis else 5

    # -- parameters ---------------------------------------------------------------------
    def set_params(self, params: Dict[str, torch.Tensor]):
        flat = np.zeros(self.nparam, dtype=np.float32)
        for name, off, shape in self.table:
            v = params[name].detach().cpu().to(torch.float32).numpy()
            if tuple(v.shape) != shape:
                raise SndvaeError(f"param '{name}' has shape {tuple(v.shape)}, expected {shape}")
            flat[off:off + v.size] = v.reshape(-1)
        self._check(self.lib.sndvae_set_params(self._h, flat.ctypes.data))

    def _unflatten(self, flat):
        return {name: torch.from_numpy(flat[off:off + int(np.prod(shape))].reshape(shape).copy())
                for name, off, shape in self.table}

    def get_params(self) -> Dict[str, torch.Tensor]:
        flat = np.empty(self.nparam, dtype=np.float32)
        self._check(self.lib.sndvae_get_params(self._h, flat.ctypes.data))
        return self._unflatten(flat)

    def get_adam(self):
        m = np.empty(self.nparam, dtype=np.float32)
        v = np.empty(self.nparam, dtype=np.float32)
        bp = np.empty(2, dtype=np.float32)
        self._check(self.lib.sndvae_get_adam(self._h, m.ctypes.data, v.ctypes.data, bp.ctypes.data))
        return self._unflatten(m), self._unflatten(v), bp

    def set_adam(self, m, v, bp):
        fm = np.zeros(self.nparam, dtype=np.float32)
        fv = np.zeros(self.nparam, dtype=np.float32)
        for name, off, shape in self.table:
            a = m[name].detach().cpu().numpy().reshape(-1); fm[off:off + a.size] = a
            b = v[name].detach().cpu().numpy().reshape(-1); fv[off:off + b.size] = b
        bp = np.asarray(bp, dtype=np.float32)
        self._check(self.lib.sndvae_set_adam(self._h, fm.ctypes.data, fv.ctypes.data, bp.ctypes.data))

    def grads_tensor(self) -> torch.Tensor:
        """Zero-copy view of the gradient arena (for torch.distributed.all_reduce)."""
        ptr = self.lib.sndvae_grads_device(self._h)
        return torch.as_tensor(_ArenaView(ptr, self.nparam), device=self.device)

    def params_tensor(self) -> torch.Tensor:
        ptr = self.lib.sndvae_params_device(self._h)
        return torch.as_tensor(_ArenaView(ptr, self.nparam), device=self.device)

    def get_grads(self) -> Dict[str, torch.Tensor]:
        return self._unflatten(self.grads_tensor().cpu().numpy())

    # -- the step -----------------------------------------------------------------------
    def forward(self, feeds, noise, fetch=_lib.OUTPUT_FIELDS):
        inp, nz, keep = self._pack(feeds, noise)
        out, res = self._outs(fetch)
        losses = np.zeros(8, dtype=np.float32)
        self._check(self.lib.sndvae_forward(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data))
        res["overall_loss"] = losses[:self.n_losses].copy()
        return res

    def grads(self, feeds, noise, fetch=("generated_adj",), global_batch=0):
        inp, nz, keep = self._pack(feeds, noise)
        out, res = self._outs(fetch)
        losses = np.zeros(8, dtype=np.float32)
        self._check(self.lib.sndvae_grads(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data, global_batch))
        res["overall_loss"] = losses[:self.n_losses].copy()
        return res

    def apply_adam(self):
        self._check(self.lib.sndvae_apply_adam(self._h))

    def zero_grads(self):
        self._check(self.lib.sndvae_zero_grads(self._h))

    def grads_accumulate(self, feeds, noise, fetch=("generated_adj",), global_batch=0):
        """One micro-batch of a larger batch: adds its gradient (scaled 1 / global_batch) to the arena."""
        inp, nz, keep = self._pack(feeds, noise)
        out, res = self._outs(fetch)
        losses = np.zeros(8, dtype=np.float32)
        self._check(self.lib.sndvae_grads_accumulate(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data, global_batch))
        res["overall_loss"] = losses[:self.n_losses].copy()
        return res

    def grads_accumulate_packed(self, inp, nz, out, losses, global_batch):
        self._check(self.lib.sndvae_grads_accumulate(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data, global_batch))

    def set_beta(self, beta: float):
        """OptimizerVAE(..., beta=...) (optimizer.py:124): the KL / regulariser weight."""
        self._check(self.lib.sndvae_set_beta(self._h, float(beta)))
        self.cfg.beta = float(beta)

    # -- data parallelism ---------------------------------------------------------------
    def comm_init(self, rank: int, world: int, broadcast=None):
        """Create the handle's NCCL communicator.  `broadcast(tensor_uint8[128])` must copy rank 0's tensor to every rank
        (default: torch.distributed.broadcast on the default process group)."""
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            rc = self.lib.sndvae_comm_unique_id(buf)
            if rc != 0:
                raise SndvaeError("sndvae_comm_unique_id failed: NCCL is not available in this process")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if broadcast is None:
            import torch.distributed as dist
            dev_ident = ident.to(self.device) if dist.get_backend() == "nccl" else ident
            dist.broadcast(dev_ident, src=0)
            ident = dev_ident.cpu()
        else:
            ident = broadcast(ident)
        raw = (C.c_uint8 * 128)(*[int(x) for x in ident.tolist()])
        self._check(self.lib.sndvae_comm_init(self._h, raw, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def allreduce_grads(self):
        self._check(self.lib.sndvae_allreduce_grads(self._h))

    def train_step(self, feeds, noise, fetch=("generated_adj",)):
        inp, nz, keep = self._pack(feeds, noise)
        out, res = self._outs(fetch)
        losses = np.zeros(8, dtype=np.float32)
        self._check(self.lib.sndvae_train_step(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data))
        res["overall_loss"] = losses[:self.n_losses].copy()
        return res

    def train_step_packed(self, inp, nz, out, losses):
        """Bench path: pre-packed structs, device-resident feeds."""
        self._check(self.lib.sndvae_train_step(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data))

    def grads_packed(self, inp, nz, out, losses, global_batch):
        self._check(self.lib.sndvae_grads(self._h, C.byref(inp), C.byref(nz), C.byref(out), losses.ctypes.data, global_batch))

    def train_step_host(self, feeds_np, noise_np, gen_adj_np, losses_np):
        """The reference-facing call: host (pinned) numpy in, host numpy out (main.py:327-331)."""
        inp = _lib.Inputs()
        for k in FEED_KEYS:
            a = feeds_np.get(k)
            if a is not None:
                setattr(inp, k, a.ctypes.data)
        nz = _lib.Noise()
        for k in NOISE_KEYS:
            a = noise_np.get(k)
            if a is not None:
                setattr(nz, k, a.ctypes.data)
        self._check(self.lib.sndvae_train_step_host(self._h, C.byref(inp), C.byref(nz), gen_adj_np.ctypes.data,
                                                    losses_np.ctypes.data))

    def grads_host(self, feeds_np, noise_np, gen_adj_np, losses_np, global_batch=0, accumulate=False):
        """Host feeds in, gradient arena (+ generated_adj, losses) out; no all-reduce, no update."""
        inp = _lib.Inputs()
        for k in FEED_KEYS:
            a = feeds_np.get(k)
            if a is not None:
                setattr(inp, k, a.ctypes.data)
        nz = _lib.Noise()
        for k in NOISE_KEYS:
            a = noise_np.get(k)
            if a is not None:
                setattr(nz, k, a.ctypes.data)
        self._check(self.lib.sndvae_grads_host(self._h, C.byref(inp), C.byref(nz), gen_adj_np.ctypes.data if gen_adj_np is not None else None,
                                               losses_np.ctypes.data, int(global_batch), 1 if accumulate else 0))

    def train_step_host_compact(self, compact_np, noise_np, gen_bits_np, losses_np):
        """sndvae_train_step_host_compact: packed host feeds (preprocessing.pack_feeds) in, bit-packed adjacency out."""
        inp = _lib.InputsCompact()
        for k, _ in _lib.InputsCompact._fields_:
            setattr(inp, k, compact_np[k].ctypes.data)
        nz = _lib.Noise()
        for k in NOISE_KEYS:
            a = noise_np.get(k)
            if a is not None:
                setattr(nz, k, a.ctypes.data)
        self._check(self.lib.sndvae_train_step_host_compact(self._h, C.byref(inp), C.byref(nz),
                                                            gen_bits_np.ctypes.data if gen_bits_np is not None else None,
                                                            losses_np.ctypes.data))

    def synth_inputs(self, seed: int) -> Dict[str, torch.Tensor]:
        """Device-side synthetic feeds (random-geometric graphs + spanning-forest samples) for this engine's batch:
        the eight arrays of construct_feed_dict_train as CUDA tensors (sndvae_synth_inputs)."""
        B, S, N, F, D = self.B, self.S, self.N, self.F, self.D
        shapes = {"features": (B * S, N, F), "spatial": (B * S, N, D), "adj": (B * S, N, N), "rel": (B * S, N, N, 1),
                  "adj_truth": (B, N, N), "feature_truth": (B, N, F), "spatial_truth": (B, N, D), "rel_truth": (B, N, N, 1)}
        t = {k: torch.empty(v, dtype=torch.float32, device=self.device) for k, v in shapes.items()}
        io = _lib.Inputs(**{k: v.data_ptr() for k, v in t.items()})
        self._check(self.lib.sndvae_synth_inputs(self._h, int(seed) & (2 ** 64 - 1), C.byref(io)))
        return t

    def set_global_iter(self, it: int):
        """The `global_iter` feed (main.py:329); read by the 'disentangled_C' loss only."""
        self._check(self.lib.sndvae_set_global_iter(self._h, int(it)))

    def generate(self, z_s, z_sg, z_g, fetch=("generated_adj", "generated_adj_prob", "generated_spatial", "generated_node_feat")):
        zs = self._dev("z_s", z_s) if self.dis else None
        zg = self._dev("z_g", z_g) if self.dis else None
        zsg = self._dev("z_sg", z_sg)
        out, res = self._outs(fetch)
        self._check(self.lib.sndvae_generate(self._h, zs.data_ptr() if zs is not None else None, zsg.data_ptr(),
                                             zg.data_ptr() if zg is not None else None, C.byref(out)))
        return res

    def threshold_logits(self, logits: torch.Tensor) -> torch.Tensor:
        lg = logits.to(device=self.device, dtype=torch.float32).contiguous()
        n = lg.numel() // 2
        out = torch.empty(lg.shape[:-1], dtype=torch.int64, device=self.device)
        self._check(self.lib.sndvae_threshold_logits(self._h, lg.data_ptr(), n, out.data_ptr()))
        return out

    def inner_product_decode(self, z: torch.Tensor) -> torch.Tensor:
        """InnerProductDecoder._call (layers.py:400-410): z [B, N, h] -> z z^T [B, N, N] (raw logits).  Not part of the
        reference's models; see include/sndvae.h."""
        zz = z.to(device=self.device, dtype=torch.float32).contiguous()
        if zz.dim() != 3:
            raise ValueError("inner_product_decode expects [batch, num_nodes, dim]")
        Bn, N, hd = zz.shape
        out = torch.empty((Bn, N, N), dtype=torch.float32, device=self.device)
        self._check(self.lib.sndvae_inner_product_decode(self._h, zz.data_ptr(), Bn, N, hd, out.data_ptr()))
        torch.cuda.synchronize(self.device)      # zz must outlive the kernels on the handle's stream
        return out

    def debug_gemm(self, A: torch.Tensor, B: torch.Tensor, tA=False, tB=False, alpha=1.0, beta=0.0, C0=None, bias=None):
        """C = alpha op(A) op(B) + beta C0 (+ bias) through the node-level contraction kernel (sndvae_debug_gemm)."""
        A = A.to(self.device, torch.float32).contiguous(); B = B.to(self.device, torch.float32).contiguous()
        K, M = (A.shape if tA else A.shape[::-1])
        N = B.shape[0] if tB else B.shape[1]
        Cm = (C0.to(self.device, torch.float32).clone().contiguous() if C0 is not None
              else torch.full((M, N), float("nan"), device=self.device))
        bt = bias.to(self.device, torch.float32).contiguous() if bias is not None else None
        self._check(self.lib.sndvae_debug_gemm(self._h, int(tA), int(tB), M, N, K, float(alpha), A.data_ptr(), A.shape[1],
                                               B.data_ptr(), B.shape[1], float(beta), Cm.data_ptr(), N,
                                               bt.data_ptr() if bt is not None else None))
        return Cm

    def debug_read(self, name: str, n: int) -> np.ndarray:
        buf = np.empty(n, dtype=np.float32)
        got = self.lib.sndvae_debug_read(self._h, name.encode(), buf.ctypes.data, n)
        if got < 0:
            self._check(int(got))
        return buf[:got]

    def stage_times(self, enable=True):
        """Per-stage CUDA-event totals since the previous call: ({stage: ms}, steps); `enable` arms the timers for the next steps."""
        cap = 64
        names = C.create_string_buffer(32 * cap)
        ms = (C.c_double * cap)()
        steps = C.c_int64()
        n = self.lib.sndvae_stage_times(self._h, 1 if enable else 0, names, ms, cap, C.byref(steps))
        if n < 0:
            self._check(int(n))
        out = {names.raw[32 * i:32 * (i + 1)].split(b"\0", 1)[0].decode(): ms[i] for i in range(n)}
        return out, int(steps.value)

    def graph_replays(self) -> int:
        return int(self.lib.sndvae_graph_replays(self._h))

    def launch_count(self) -> int:
        return int(self.lib.sndvae_launch_count(self._h))

    def gemm_timing(self, reset=True):
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.lib.sndvae_gemm_timing(self._h, 1 if reset else 0, C.byref(ms), C.byref(n), C.byref(fl)))
        return ms.value, n.value, fl.value
