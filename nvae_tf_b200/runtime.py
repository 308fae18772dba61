"""Host runtime under the Keras-API mirror: parameter arenas, the gradient tape and the op
wrappers that launch libnvae_b200 through the C-ABI (`_lib.py`).

What TensorFlow gives the reference implicitly is made explicit here:
  * variables live in two flat fp32 arenas (trainable `params` + `grads`, non-trainable `state`)
    so the optimizer, the gradient all-reduce, spectral normalisation and the BN-gamma
    regulariser are each ONE launch over the arena (reference: ~750 per-variable TF ops);
  * `tf.GradientTape` (models.py:116,127) becomes a list of backward closures recorded by the op
    wrappers below and replayed in reverse;
  * torch is used for device memory, streams and torch.distributed only -- every arithmetic
    operation on the path is a libnvae_b200 kernel.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from contextlib import contextmanager
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import NVAE_ACT_ELU, NVAE_ACT_NONE, NVAE_ACT_SWISH, NvaeConvDesc, NvaeSnLayer

BN_MOMENTUM = 0.05  # common.py:148 (Keras momentum = retain factor)
BN_EPS = 1e-5

_STACK: List["Runtime"] = []


def current() -> "Runtime":
    if not _STACK:
        raise RuntimeError("no active nvae_tf_b200 Runtime: construct layers inside `with Runtime():` or via NVAE(...)")
    return _STACK[-1]


class Variable:
    """A tf.Variable stand-in: a named view into the parameter (or state) arena."""

    def __init__(self, name: str, shape: Sequence[int], init: np.ndarray, trainable: bool):
        self.name, self.shape, self.trainable = name, tuple(int(s) for s in shape), trainable
        self.size = int(np.prod(self.shape))
        self.offset = -1
        self._init = np.asarray(init, dtype=np.float32).reshape(self.shape)
        self.value: Optional[torch.Tensor] = None
        self.grad: Optional[torch.Tensor] = None

    def ptr(self) -> int:
        return self.value.data_ptr()

    def gptr(self) -> int:
        return self.grad.data_ptr()

    def assign(self, arr) -> None:
        t = torch.as_tensor(np.asarray(arr, dtype=np.float32)).reshape(self.shape)
        self.value.copy_(t)

    def numpy(self) -> np.ndarray:
        return self.value.detach().cpu().numpy().copy()


class DeviceTensor:
    """An activation: fp32 NHWC device buffer plus its (lazily created) gradient buffer."""
    __slots__ = ("data", "grad", "needs_grad")

    def __init__(self, data: torch.Tensor, needs_grad: bool = True):
        self.data, self.grad, self.needs_grad = data, None, needs_grad

    @property
    def shape(self):
        return tuple(self.data.shape)

    def ptr(self) -> int:
        return self.data.data_ptr()

    def numpy(self) -> np.ndarray:
        return self.data.detach().cpu().numpy()


def same_pad(in_size: int, k: int, stride: int) -> Tuple[int, int]:
    """TF SAME: out = ceil(in/stride); the odd padding element goes AFTER."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2


class Runtime:
    def __init__(self, device: Optional[str] = None, precision: Optional[int] = None, seed: int = 1):
        self.lib = _lib.lib()  # raises if libnvae_b200.so is missing: no fallback
        if device is not None and torch.device(device).type == "cpu":
            # layout-only mode: variables can be created, named and counted (checkpoint mapping, tests of
            # the host logic) but every launcher raises -- there is no CPU arithmetic path.
            self.lib = _lib.NoDeviceLib()
            self.device = torch.device("cpu")
        elif not torch.cuda.is_available():
            raise _lib.NvaeError("nvae_tf_b200 needs a CUDA device (sm_100a); there is no CPU path")
        else:
            self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.precision = precision = _lib.default_precision() if precision is None else precision
        self.pack_exact = int(precision == _lib.NVAE_PREC_TF32X3)
        # which tensor-core conv operands get an explicit RN-to-TF32 pass when their producer did not round them
        self.tf32_round = os.environ.get("NVAE_TF32_ROUND", "dy")  # none | dy | all
        self.rng = np.random.default_rng(seed)
        self.variables: Dict[str, Variable] = {}
        self.sn_convs: List["object"] = []  # SpectralNormalization wrappers, registration order
        self.convs: List["object"] = []     # every Conv2D (owner of the TF32 operand copies in self.pack)
        self.pack: Optional[torch.Tensor] = None
        self.bn_loss_layers: List["object"] = []
        self._scope: List[str] = []
        self.tape: Optional[List[Callable[[], None]]] = None
        self.params = self.grads = self.state = None
        self._ws: Dict[str, torch.Tensor] = {}
        self._ws_retired: List[torch.Tensor] = []
        self._side: Optional[torch.cuda.Stream] = None
        self._on_side = False
        self._side_dirty = False
        self._keepalive: List[torch.Tensor] = []
        # single-pass TF32 rounds gradients in place before use, which a concurrent reader must not see half-done
        # BN-apply (+ activation) of the BN -> 1x1 conv pairs inside the convolution's operand path (bn_conv2d)
        # 0 (default): never; 1: a BN without activation (an FMA per element); 2: also BN + swish / ELU.  Off by default:
        # the converter warps that would apply it are the critical path of these small GEMMs
        self.fuse_bn_conv = int(os.environ.get("NVAE_FUSE_BN_CONV", "0"))
        self.use_side_stream = (os.environ.get("NVAE_WGRAD_STREAM", "1") != "0" and
                                self.precision != _lib.NVAE_PREC_TF32 and self.device.type == "cuda")
        self.sn_done = False
        self.eps_injected: Optional[List[torch.Tensor]] = None
        self.eps_i = 0
        self.philox_seed = seed
        self.finalized = False

    # ---- scoping -------------------------------------------------------------------------
    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        _STACK.pop()

    @contextmanager
    def scope(self, name: str):
        self._scope.append(name)
        try:
            yield
        finally:
            self._scope.pop()

    def full_name(self, name: str) -> str:
        return "/".join(self._scope + [name]) if name else "/".join(self._scope)

    # ---- variables -----------------------------------------------------------------------
    def add_variable(self, name: str, shape, init: np.ndarray, trainable: bool = True) -> Variable:
        if self.finalized:
            raise RuntimeError("Runtime already finalized; create all layers first")
        full = self.full_name(name)
        if full in self.variables:
            raise ValueError(f"duplicate variable {full}")
        v = Variable(full, shape, init, trainable)
        self.variables[full] = v
        return v

    def glorot(self, shape, fan_in, fan_out) -> np.ndarray:
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return self.rng.uniform(-lim, lim, size=shape)

    def finalize(self) -> None:
        """Lay the variables out in the arenas (16-byte aligned) and upload the initial values."""
        if self.finalized:
            return
        offs = {True: 0, False: 0}
        for v in self.variables.values():
            v.offset = offs[v.trainable]
            offs[v.trainable] += (v.size + 3) // 4 * 4
        n_p, n_s = max(offs[True], 4), max(offs[False], 4)
        hp, hs = np.zeros(n_p, np.float32), np.zeros(n_s, np.float32)
        for v in self.variables.values():
            (hp if v.trainable else hs)[v.offset:v.offset + v.size] = v._init.ravel()
        self.params = torch.from_numpy(hp).to(self.device)
        self.state = torch.from_numpy(hs).to(self.device)
        self.grads = torch.zeros_like(self.params)
        for v in self.variables.values():
            arena = self.params if v.trainable else self.state
            v.value = arena[v.offset:v.offset + v.size].view(v.shape)
            if v.trainable:
                v.grad = self.grads[v.offset:v.offset + v.size].view(v.shape)
            v._init = None
        self.finalized = True
        self._alloc_packs()
        self._build_sn_tables()
        self._build_bn_loss_tables()
        self.zero = torch.zeros(8, device=self.device)

    @property
    def trainable_variables(self) -> List[Variable]:
        return [v for v in self.variables.values() if v.trainable]

    def n_trainable(self) -> int:
        return sum(v.size for v in self.trainable_variables)

    def load_named(self, named: Dict[str, np.ndarray]) -> None:
        missing = [k for k in self.variables if k not in named]
        extra = [k for k in named if k not in self.variables]
        if missing or extra:
            raise KeyError(f"variable name mismatch: missing {missing[:5]} extra {extra[:5]}")
        for k, v in self.variables.items():
            v.assign(named[k])

    def named_values(self) -> Dict[str, np.ndarray]:
        return {k: v.numpy() for k, v in self.variables.items()}

    def named_grads(self) -> Dict[str, np.ndarray]:
        return {k: v.grad.detach().cpu().numpy().copy() for k, v in self.variables.items() if v.trainable}

    # ---- memory / launch helpers ------------------------------------------------------------
    @property
    def stream(self) -> int:
        if self.device.type != "cuda":
            raise _lib.NvaeError("layout-only Runtime (device='cpu') cannot launch kernels")
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, *shape) -> torch.Tensor:
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def zeros(self, *shape) -> torch.Tensor:
        t = self.empty(*shape)
        self.lib.fill(t.data_ptr(), t.numel(), 0.0, self.stream)
        return t

    def workspace(self, nbytes: int) -> Tuple[int, int]:
        """Scratch for the launch being issued on the CURRENT stream (one buffer per stream: kernels on the
        weight-gradient side stream must not share scratch with the main stream's)."""
        nbytes = int(nbytes)
        key = "side" if self._on_side else "main"
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.NvaeError("workspace must be sized by an eager warm-up step before graph capture")
            torch.cuda.synchronize(self.device)  # nothing in flight may still be using the old buffer
            if ws is not None:
                self._ws_retired.append(ws)  # an already-captured graph (another batch shape) may have its address baked in
            ws = self._ws[key] = torch.empty(max(nbytes, 32 << 20), dtype=torch.uint8, device=self.device)
        return ws.data_ptr(), ws.numel()

    # ---- weight-gradient side stream ------------------------------------------------------------------
    @contextmanager
    def side_stream(self):
        """Runs the enclosed launches (backward-filter + bias-gradient of one conv) on a second stream that waits
        for everything issued so far on the current one.  Independent of the rest of backward (they only write the
        gradient arena), these mostly small, latency-bound kernels then overlap the dgrad / BN-backward chain.
        Works eagerly and under CUDA-graph capture (fork/join become graph edges).  `join_side_stream` must run
        before anything reads the gradient arena."""
        if not self.use_side_stream:
            yield
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(main)
        self._on_side = True
        try:
            with torch.cuda.stream(self._side):
                yield
        finally:
            self._on_side = False
            self._side_dirty = True

    def keep_alive(self, t) -> None:
        """Holds a reference to a buffer a side-stream launch reads until the streams are joined."""
        if self.use_side_stream:
            self._keepalive.append(t)

    def join_side_stream(self) -> None:
        if self._side_dirty:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._side_dirty = False
        self._keepalive.clear()

    def tensor(self, data: torch.Tensor, needs_grad: bool = True) -> DeviceTensor:
        return DeviceTensor(data, needs_grad)

    def from_host(self, arr, needs_grad: bool = False) -> DeviceTensor:
        t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32)).to(self.device)
        return DeviceTensor(t, needs_grad)

    def record(self, fn: Callable[[], None]) -> None:
        if self.tape is not None:
            self.tape.append(fn)

    def grad_target(self, t: DeviceTensor) -> Tuple[torch.Tensor, int]:
        """Buffer to write d(loss)/d(t) into and whether to accumulate ('+=') or assign ('=')."""
        if t.grad is None:
            t.grad = torch.empty_like(t.data)
            return t.grad, 0
        return t.grad, 1

    def add_grad(self, t: DeviceTensor, g: torch.Tensor, take: bool = True) -> None:
        """t.grad += g; when t has no gradient yet the buffer g is adopted (take) or copied."""
        if not t.needs_grad:
            return
        if t.grad is None:
            if take:
                t.grad = g
            else:
                t.grad = torch.empty_like(g)
                self.lib.axpby(g.data_ptr(), 1.0, t.grad.data_ptr(), 0.0, g.numel(), self.stream)
        else:
            self.lib.axpby(g.data_ptr(), 1.0, t.grad.data_ptr(), 1.0, g.numel(), self.stream)

    # ---- tape ------------------------------------------------------------------------------
    @contextmanager
    def gradient_tape(self):
        """tf.GradientTape() of models.py:116."""
        prev, self.tape = self.tape, []
        try:
            yield self.tape
        finally:
            self._last_tape, self.tape = self.tape, prev

    def backward(self, tape: List[Callable[[], None]], split_at: Optional[int] = None,
                 on_split: Optional[Callable[[], None]] = None) -> None:
        """Replays the tape in reverse.  `split_at` (a tape index) / `on_split`: once every closure recorded at or after
        that index has run -- and the weight-gradient side stream has been joined, so those layers' gradients are final --
        `on_split()` is called; the data-parallel step uses it to end one CUDA graph and begin the next, so the first
        gradient bucket can be all-reduced while the rest of backward runs."""
        for i in range(len(tape) - 1, -1, -1):
            if split_at is not None and on_split is not None and i == split_at - 1:
                self.join_side_stream()
                on_split()
            tape[i]()
        self.join_side_stream()
        tape.clear()

    # ---- epsilon source (common.py:67) -----------------------------------------------------------
    def inject_eps(self, eps: Optional[Sequence]) -> None:
        self.eps_injected = None if eps is None else [
            e if isinstance(e, torch.Tensor) else torch.as_tensor(np.asarray(e, dtype=np.float32)).to(self.device)
            for e in eps]
        self.eps_i = 0

    def next_eps(self, shape) -> torch.Tensor:
        if self.eps_injected is not None:
            e = self.eps_injected[self.eps_i % len(self.eps_injected)]
            self.eps_i += 1
            if tuple(e.shape) != tuple(shape):
                raise ValueError(f"injected epsilon {tuple(e.shape)} != {tuple(shape)}")
            return e
        out = self.empty(*shape)
        cnt = getattr(self, "counters", None)
        # device-side iteration counter (counters[1]) keys the stream so graph replays draw fresh noise
        self.lib.philox_normal(out.data_ptr(), out.numel(), self.philox_seed,
                               cnt[1:].data_ptr() if cnt is not None else None, self.eps_i, self.stream)
        self.eps_i += 1
        return out

    # ---- TF32 operand copies of the conv kernels (tensor-core modes) -----------------------------------
    def _alloc_packs(self) -> None:
        """Two operand copies per Conv2D kernel: HWIO (dgrad B operand) and [Cout][tap][Cin] (forward B operand),
        TF32-rounded in NVAE_PREC_TF32 and exact in NVAE_PREC_TF32X3.  They are rewritten by nvae_spectral_norm's
        last pass whenever the weights change."""
        if self.precision == _lib.NVAE_PREC_FP32 or not self.convs:
            return
        off = 0
        for conv in self.convs:
            n = (conv.kernel.size + 3) // 4 * 4
            if self.pack_exact:  # 3xTF32: the HWIO master kernel itself is the dgrad operand
                conv.rnd_off, conv.tr_off = -1, off
                off += n
            else:
                conv.rnd_off, conv.tr_off = off, off + n
                off += 2 * n
        self.pack = torch.zeros(max(off, 4), device=self.device)
        self._plain_tables: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def _sn_layer_entry(self, L, k, u_off: int, ws_off: int, conv) -> int:
        rows, cout = k.size // k.shape[-1], k.shape[-1]
        nch = -(-rows // _lib.SN_ROWS_PER_CHUNK)
        L.w_off, L.u_off = k.offset, u_off
        L.v_off = ws_off
        ws_off += (rows + 3) // 4 * 4
        L.t_off = ws_off
        ws_off += nch * (cout + 1)
        ws_off = (ws_off + 3) // 4 * 4
        L.rnd_off, L.tr_off = conv.rnd_off, conv.tr_off
        L.rows, L.cout = rows, cout
        L.taps, L.cin = k.shape[0] * k.shape[1], k.shape[2]
        L.cin_pad, L.cout_pad = k.shape[2], cout
        L.n_chunks = nch
        return ws_off

    def pack_plain(self, conv) -> None:
        """Refresh the operand copies of a Conv2D that is not SN-wrapped (pack-only pass, no power iteration)."""
        key = id(conv)
        if key not in self._plain_tables:
            one = (NvaeSnLayer * 1)()
            self._sn_layer_entry(one[0], conv.kernel, 0, 0, conv)
            one[0].chunk0 = 0
            dev = torch.frombuffer(bytearray(bytes(one)), dtype=torch.uint8).to(self.device)
            cl = torch.zeros(one[0].n_chunks, dtype=torch.int32, device=self.device)
            self._plain_tables[key] = (dev, cl)
        dev, cl = self._plain_tables[key]
        self.lib.spectral_norm(self.params.data_ptr(), self.state.data_ptr(), self.pack.data_ptr(), dev.data_ptr(), 1,
                               cl.data_ptr(), cl.numel(), 0, self.pack_exact, None, None, self.stream)

    # ---- spectral normalisation tables -------------------------------------------------
    def _build_sn_tables(self) -> None:
        n = len(self.sn_convs)
        self.sn_n = n
        if n == 0:
            return
        arr = (NvaeSnLayer * n)()
        chunk_layer: List[int] = []
        ws_off = 0
        for i, sn in enumerate(self.sn_convs):
            L = arr[i]
            ws_off = self._sn_layer_entry(L, sn.layer.kernel, sn.u.offset, ws_off, sn.layer)
            L.chunk0 = len(chunk_layer)
            chunk_layer += [i] * L.n_chunks
            sn.index = i
        self.sn_host = arr
        self.sn_layers_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        self.sn_chunk_layer = torch.tensor(chunk_layer, dtype=torch.int32, device=self.device)
        self.sn_ws = torch.zeros(max(ws_off, 4), device=self.device)
        self.sn_sigma = torch.ones(n, device=self.device)
        self._sn_single: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def _pack_ptr(self):
        return self.pack.data_ptr() if self.pack is not None else None

    def spectral_normalize_all(self, power_iter: bool = True) -> None:
        """One power iteration + in-place W/sigma for every SN-wrapped conv (4 launches total) and, in the
        tensor-core modes, the refresh of the TF32 operand copies.  power_iter=False only refreshes the copies
        (inference: SN inactive)."""
        if self.sn_n == 0 or (not power_iter and self.pack is None):
            return
        self.lib.spectral_norm(self.params.data_ptr(), self.state.data_ptr(), self._pack_ptr(),
                               self.sn_layers_dev.data_ptr(), self.sn_n, self.sn_chunk_layer.data_ptr(),
                               self.sn_chunk_layer.numel(), int(power_iter), self.pack_exact, self.sn_sigma.data_ptr(),
                               self.sn_ws.data_ptr(), self.stream)

    @contextmanager
    def weights_ready(self):
        """Inference entry points (NVAE.sample, sample_with_z): refresh the TF32 operand copies once (the
        optimizer may have moved the weights since) and mark them valid for the enclosed layer calls."""
        if self.sn_done or self.pack is None:
            yield
            return
        self.spectral_normalize_all(power_iter=False)
        self.sn_done = True
        try:
            yield
        finally:
            self.sn_done = False

    def spectral_normalize_one(self, index: int, power_iter: bool = True) -> None:
        """Per-layer path for layers called outside NVAE.call (cell micro-benchmarks, tests)."""
        if index not in self._sn_single:
            one = (NvaeSnLayer * 1)()
            C.memmove(one, C.byref(self.sn_host[index]), C.sizeof(NvaeSnLayer))
            one[0].chunk0 = 0
            dev = torch.frombuffer(bytearray(bytes(one)), dtype=torch.uint8).to(self.device)
            cl = torch.zeros(one[0].n_chunks, dtype=torch.int32, device=self.device)
            self._sn_single[index] = (dev, cl)
        dev, cl = self._sn_single[index]
        self.lib.spectral_norm(self.params.data_ptr(), self.state.data_ptr(), self._pack_ptr(), dev.data_ptr(), 1,
                               cl.data_ptr(), cl.numel(), int(power_iter), self.pack_exact,
                               self.sn_sigma[index:].data_ptr(),
                               self.sn_ws.data_ptr(), self.stream)

    # ---- BN-gamma regulariser tables (models.py:252-267) -----------------------------------------
    def _build_bn_loss_tables(self) -> None:
        offs = [bn.gamma.offset for bn in self.bn_loss_layers]
        sizes = [bn.gamma.size for bn in self.bn_loss_layers]
        self.bn_loss_offsets = torch.tensor(offs or [0], dtype=torch.int64, device=self.device)
        self.bn_loss_sizes = torch.tensor(sizes or [0], dtype=torch.int32, device=self.device)
        self.bn_loss_n = len(offs)


# ==========================================================================================
# op wrappers: forward launch + backward closure.  x/y are DeviceTensor, NHWC fp32.
# ==========================================================================================
def _rows(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


def bn_stats(rt: Runtime, x: DeviceTensor, bn, training: bool) -> torch.Tensor:
    """Per-channel {mean, invstd, scale, shift}; updates the moving statistics when training."""
    Cc = x.shape[-1]
    stat = rt.empty(4, Cc)
    ws, wsb = rt.workspace(rt.lib._nvae_bn_ws_bytes(_rows(x.data), Cc))
    rt.lib.bn_stats(x.ptr(), _rows(x.data), Cc, bn.gamma.ptr(), bn.beta.ptr(), bn.moving_mean.ptr(),
                    bn.moving_variance.ptr(), int(training), bn.momentum, bn.epsilon, stat.data_ptr(), ws, wsb,
                    rt.stream)
    return stat


def _bn_backward(rt: Runtime, dout: torch.Tensor, x: DeviceTensor, stat: Optional[torch.Tensor], bn, act: int,
                 up: Tuple[int, int], training: bool) -> None:
    """dx (+)= BN/activation backward of `dout` (gradient w.r.t. the activated output)."""
    Cc = x.shape[-1]
    rows = _rows(x.data)
    want_dx = x.needs_grad
    dx, accum = rt.grad_target(x) if want_dx else (None, 0)
    ws, wsb = rt.workspace(rt.lib._nvae_bn_ws_bytes(rows, Cc))
    rt.lib.bn_act_bwd(dout.data_ptr(), x.ptr(), rows, Cc, stat.data_ptr() if stat is not None else None, act, up[0],
                      up[1], int(training), None, 0.0, accum, dx.data_ptr() if dx is not None else None,
                      bn.gamma.gptr() if bn is not None else None, bn.beta.gptr() if bn is not None else None, ws, wsb,
                      rt.stream)


def bn_act(rt: Runtime, x: DeviceTensor, bn, act: int, training: bool, upsample: bool = False) -> DeviceTensor:
    """act(BN(x)) [-> nearest x2]; bn=None is a bare activation (layers.ELU / activations.swish)."""
    N, H, W, Cc = x.shape
    up = (H, W) if upsample else (0, 0)
    out = rt.empty(N, 2 * H, 2 * W, Cc) if upsample else rt.empty(N, H, W, Cc)
    rnd = int(rt.precision == _lib.NVAE_PREC_TF32)
    if bn is not None:  # statistics + apply + activation: one launch when the tensor fits L2
        stat = rt.empty(4, Cc)
        ws, wsb = rt.workspace(rt.lib._nvae_bn_ws_bytes(_rows(x.data), Cc))
        rt.lib.bn_fwd(x.ptr(), _rows(x.data), Cc, bn.gamma.ptr(), bn.beta.ptr(), bn.moving_mean.ptr(),
                      bn.moving_variance.ptr(), int(training), bn.momentum, bn.epsilon, stat.data_ptr(), act, up[0],
                      up[1], rnd, out.data_ptr(), ws, wsb, rt.stream)
    else:
        stat = None
        rt.lib.bn_act_fwd(x.ptr(), _rows(x.data), Cc, None, act, up[0], up[1], rnd, out.data_ptr(), rt.stream)
    y = DeviceTensor(out, x.needs_grad or bn is not None)
    if rt.tape is not None:
        def bwd():
            if y.grad is None:
                raise RuntimeError("bn_act output has no gradient")
            _bn_backward(rt, y.grad, x, stat, bn, act, up, training and bn is not None)
            y.grad = None
        rt.record(bwd)
    return y


def conv_desc(rt: Runtime, x_shape, cin2: int, kshape, stride: int, shift: Tuple[int, int] = (0, 0),
              y_ld: int = 0, y_off: int = 0, pre: Tuple[float, float] = (0.0, 0.0)) -> NvaeConvDesc:
    N, H, W, Cin = x_shape
    R, S, Ct, Cout = kshape
    if Ct != Cin + cin2:
        raise ValueError(f"kernel expects {Ct} input channels, got {Cin}+{cin2}")
    Ho, pt = same_pad(H - shift[0], R, stride)
    Wo, pl = same_pad(W - shift[1], S, stride)
    d = NvaeConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cin2, d.Cout, d.R, d.S = N, H, W, Cin, cin2, Cout, R, S
    d.stride, d.Ho, d.Wo, d.pad_t, d.pad_l = stride, Ho, Wo, pt - shift[0], pl - shift[1]
    d.precision, d.y_ld, d.y_off, d.pre_scale, d.pre_shift = rt.precision, y_ld, y_off, pre[0], pre[1]
    return d


def conv2d(rt: Runtime, x: DeviceTensor, conv, x2: Optional[DeviceTensor] = None,
           residual: Optional[DeviceTensor] = None, shift: Tuple[int, int] = (0, 0),
           out: Optional[DeviceTensor] = None, y_off: int = 0, pre: Tuple[float, float] = (0.0, 0.0)) -> DeviceTensor:
    """Conv2D(padding='same') on x (++ x2 along channels) + bias (+ residual).

    `shift` convolves the view x[:, shift[0]:, shift[1]:, :] (SkipScaler, preprocess.py:69-71);
    `out`/`y_off` write channels [y_off, y_off+Cout) of an existing tensor (the tf.concat of
    preprocess.py:73)."""
    k = conv.kernel
    d = conv_desc(rt, x.shape, x2.shape[-1] if x2 is not None else 0, k.shape, conv.stride, shift,
                  out.shape[-1] if out is not None else 0, y_off, pre)
    y = out if out is not None else DeviceTensor(rt.empty(d.N, d.Ho, d.Wo, d.Cout))
    ws, wsb = rt.workspace(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), 0))
    bias = conv.bias
    # single-pass TF32 only: operands whose producer did not round them get an explicit RN pass (3xTF32 splits
    # the raw fp32 tiles into high and low parts inside the convolution kernel)
    tc = [bool(rt.lib._nvae_conv2d_uses_tensor_cores(C.byref(d), i)) for i in range(3)] \
        if rt.precision == _lib.NVAE_PREC_TF32 else [False] * 3
    if rt.tf32_round == "all" and (tc[0] or tc[2]):
        for t in (x, x2):
            if t is not None:
                rt.lib.round_tf32(t.ptr(), t.data.numel(), rt.stream)
    rt.lib.conv2d_fwd(C.byref(d), x.ptr(), x2.ptr() if x2 is not None else None, k.ptr(), conv.packed_fwd(),
                      bias.ptr() if bias is not None else None, residual.ptr() if residual is not None else None,
                      y.ptr(), ws, wsb, rt.stream)
    if rt.tape is not None:
        def bwd():
            dy = y.grad
            if dy is None:
                raise RuntimeError(f"conv {k.name}: output has no gradient")
            if rt.tf32_round != "none" and (tc[1] or tc[2]):
                rt.lib.round_tf32(dy.data_ptr(), dy.numel(), rt.stream)
            with rt.side_stream():
                ws, wsb = rt.workspace(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), 2))
                rt.lib.conv2d_wgrad(C.byref(d), x.ptr(), x2.ptr() if x2 is not None else None, dy.data_ptr(),
                                    k.gptr(), bias.gptr() if bias is not None else None, ws, wsb, rt.stream)
            rt.keep_alive(dy)
            need1, need2 = x.needs_grad, x2 is not None and x2.needs_grad
            if need1 or need2:
                if not need1 or (x2 is not None and not need2):
                    raise NotImplementedError("dgrad of a concatenated conv needs both inputs differentiable")
                accum = int(x.grad is not None or (x2 is not None and x2.grad is not None))
                if accum:
                    for t in (x, x2):
                        if t is not None and t.grad is None:
                            t.grad = rt.zeros(*t.shape)
                dx, _ = rt.grad_target(x)
                dx2 = rt.grad_target(x2)[0] if x2 is not None else None
                ws, wsb = rt.workspace(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), 1))
                rt.lib.conv2d_dgrad(C.byref(d), dy.data_ptr(), k.ptr(), conv.packed_dgrad(), dx.data_ptr(),
                                    dx2.data_ptr() if dx2 is not None else None, accum, ws, wsb, rt.stream)
            if residual is not None:
                if d.y_ld not in (0, d.Cout):
                    raise NotImplementedError("residual with a concatenated output")
                # the side-stream wgrad above still reads dy: adopting the buffer as residual.grad would let later
                # main-stream accumulations (se_bwd / bn_act_bwd with accum=1) overwrite it under the reader, so with
                # the side stream on the residual gets its own copy (one axpby of a 2-5 MB tensor)
                rt.add_grad(residual, dy, take=out is None and not rt.use_side_stream)
            if out is None:
                y.grad = None
        rt.record(bwd)
    return y


def bn_conv2d(rt: Runtime, x: DeviceTensor, bn, act: int, conv, training: bool,
              residual: Optional[DeviceTensor] = None) -> DeviceTensor:
    """conv(act(BN(x))) (+ bias, + residual) for the BN -> [Swish] -> 1x1 Conv2D pairs of the cells (decoder.py:125-127,
    143-144; postprocess.py:71-73, 84-96).  Where the kernel takes it (nvae_conv2d_bnact_supported) the BN-apply and the
    activation can run in the operand path of the tensor-core convolution, forward and backward-filter: one statistics
    launch and the convolution, the activated tensor is never written (NVAE_FUSE_BN_CONV=1: BN without activation, 2: all).
    Bit-identical to, and measured SLOWER than, bn_act followed by conv2d (31.1 / 32.6 against 29.8 ms per step: the
    converter warps are these GEMMs' critical path, the activation is redone per N tile and again in backward-filter),
    so the default (0) is the unfused pair."""
    k = conv.kernel
    d = conv_desc(rt, x.shape, 0, k.shape, conv.stride)
    from ._lib import NVAE_ACT_NONE
    if rt.fuse_bn_conv < (1 if act == NVAE_ACT_NONE else 2) or not rt.lib._nvae_conv2d_bnact_supported(C.byref(d)):
        return conv2d(rt, bn_act(rt, x, bn, act, training), conv, residual=residual)
    stat = bn_stats(rt, x, bn, training)
    y = DeviceTensor(rt.empty(d.N, d.Ho, d.Wo, d.Cout))
    ws, wsb = rt.workspace(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), 0))
    bias = conv.bias
    rt.lib.conv2d_fwd_bnact(C.byref(d), x.ptr(), stat.data_ptr(), act, conv.packed_fwd(),
                            bias.ptr() if bias is not None else None, residual.ptr() if residual is not None else None,
                            y.ptr(), ws, wsb, rt.stream)
    if rt.tape is not None:
        def bwd():
            dy = y.grad
            if dy is None:
                raise RuntimeError(f"conv {k.name}: output has no gradient")
            with rt.side_stream():
                ws, wsb = rt.workspace(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), 2))
                rt.lib.conv2d_wgrad_bnact(C.byref(d), x.ptr(), stat.data_ptr(), act, dy.data_ptr(), k.gptr(),
                                          bias.gptr() if bias is not None else None, ws, wsb, rt.stream)
            rt.keep_alive(dy)
            rt.keep_alive(stat)
            da = rt.empty(*x.shape)  # gradient of the activated tensor, consumed by the BN backward right below
            ws, wsb = rt.workspace(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), 1))
            rt.lib.conv2d_dgrad(C.byref(d), dy.data_ptr(), k.ptr(), conv.packed_dgrad(), da.data_ptr(), None, 0, ws, wsb,
                                rt.stream)
            if residual is not None:
                rt.add_grad(residual, dy, take=not rt.use_side_stream)
            _bn_backward(rt, da, x, stat, bn, act, (0, 0), training)
            y.grad = None
        rt.record(bwd)
    return y


def dwconv_bn_act(rt: Runtime, x: DeviceTensor, bn, act: int, dw, training: bool) -> DeviceTensor:
    """DepthwiseConv2D(5x5)(act(BN(x))) with the BN-apply + activation fused into the load (decoder.py:141-142)."""
    N, H, W, Cc = x.shape
    stat = bn_stats(rt, x, bn, training)
    y = DeviceTensor(rt.empty(N, H, W, Cc))
    rt.lib.dwconv5x5_fwd(x.ptr(), stat.data_ptr(), act, N, H, W, Cc, dw.depthwise_kernel.ptr(), dw.bias.ptr(), y.ptr(),
                         rt.stream)
    if rt.tape is not None:
        def bwd():
            dy = y.grad
            da = rt.empty(N, H, W, Cc)
            rt.lib.dwconv5x5_bwd_data(dy.data_ptr(), N, H, W, Cc, dw.depthwise_kernel.ptr(), da.data_ptr(), rt.stream)
            with rt.side_stream():
                ws, wsb = rt.workspace(rt.lib._nvae_dwconv5x5_bwd_filter_ws_bytes(N, H, W, Cc))
                rt.lib.dwconv5x5_bwd_filter(x.ptr(), stat.data_ptr(), act, dy.data_ptr(), N, H, W, Cc,
                                            dw.depthwise_kernel.gptr(), dw.bias.gptr(), ws, wsb, rt.stream)
            rt.keep_alive(dy)
            rt.keep_alive(stat)
            _bn_backward(rt, da, x, stat, bn, act, (0, 0), training)
            y.grad = None
        rt.record(bwd)
    return y


def se_residual(rt: Runtime, t: DeviceTensor, bn, xres: DeviceTensor, se, alpha: float, beta: float,
                training: bool) -> DeviceTensor:
    """y = alpha*xres + beta*SE(BN(t))  (bn may be None): common.py:129-142 fused with the cell tail."""
    N, H, W, Cc = t.shape
    stat = bn_stats(rt, t, bn, training) if bn is not None else None
    hid = se.dense1.units
    pooled, hidden, gate = rt.empty(N, Cc), rt.empty(N, hid), rt.empty(N, Cc)
    y = DeviceTensor(rt.empty(N, H, W, Cc))
    sp = stat.data_ptr() if stat is not None else None
    rt.lib.se_fwd(t.ptr(), sp, xres.ptr(), N, H * W, Cc, hid, se.dense1.kernel.ptr(), se.dense1.bias.ptr(),
                  se.dense2.kernel.ptr(), se.dense2.bias.ptr(), alpha, beta, pooled.data_ptr(), hidden.data_ptr(),
                  gate.data_ptr(), y.ptr(), rt.stream)
    if rt.tape is not None:
        def bwd():
            dy = y.grad
            dt = rt.empty(N, H, W, Cc)
            dxr, accum = rt.grad_target(xres) if xres.needs_grad else (None, 0)
            ws, wsb = rt.workspace(rt.lib._nvae_se_bwd_ws_bytes(N, Cc, hid))
            rt.lib.se_bwd(dy.data_ptr(), t.ptr(), sp, N, H * W, Cc, hid, se.dense1.kernel.ptr(),
                          se.dense2.kernel.ptr(), pooled.data_ptr(), hidden.data_ptr(), gate.data_ptr(), alpha, beta,
                          dt.data_ptr(), dxr.data_ptr() if dxr is not None else None, accum, se.dense1.kernel.gptr(),
                          se.dense1.bias.gptr(), se.dense2.kernel.gptr(), se.dense2.bias.gptr(), ws, wsb, rt.stream)
            if bn is not None:
                _bn_backward(rt, dt, t, stat, bn, NVAE_ACT_NONE, (0, 0), training)
            else:
                rt.add_grad(t, dt)
            y.grad = None
        rt.record(bwd)
    return y


def broadcast_batch(rt: Runtime, var: Variable, batch: int) -> DeviceTensor:
    """tf.tile(tf.expand_dims(h, 0), [B,1,1,1]) (decoder.py:73-74)."""
    out = DeviceTensor(rt.empty(batch, *var.shape))
    rt.lib.broadcast_rows(var.ptr(), var.size, batch, out.ptr(), rt.stream)
    if rt.tape is not None and var.trainable:
        def bwd():
            rt.lib.reduce_rows(out.grad.data_ptr(), var.size, batch, var.gptr(), rt.stream)
            out.grad = None
        rt.record(bwd)
    return out
