"""Drop-in for the reference's models.py: the NVAE model class (models.py:16-267) -- same constructor
kwargs as train.py:110-127 passes, same public methods (`__call__`, `train_step`, `sample`,
`sample_with_z`, `calculate_kl_loss`, `calculate_kl_alphas`, `calculate_recon_loss`,
`calculate_bn_loss`, `on_epoch_begin`, `compile`, `save_weights`/`load_weights`) and the same
`train_step` result keys -- with every arithmetic op a libnvae_b200 launch.

Two things Keras leaves implicit are explicit arguments here:
  * `training`: NVAE.call has no such argument in the reference and train_step passes none
    (models.py:89,117).  `NVAE(..., training=True)` (default, what BASELINE's north star
    measures: batch-statistics BN + spectral-norm power iteration) or `training=False`
    (moving statistics, SN inactive) selects what train_step runs;
  * data parallelism (new): `NVAE(..., process_group=...)` all-reduces the flat
    gradient arena with NCCL before the optimizer; BN statistics and KL-balance coefficients
    stay per replica.
"""
from __future__ import annotations

import os
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from . import runtime as R
from .common import DistributionParams, Rescaler
from .decoder import Decoder, DecoderSampleCombiner
from .encoder import Encoder
from .layers import BatchNormalization
from .postprocess import Postprocess
from .preprocess import Preprocess
from .runtime import DeviceTensor, Runtime


class CosineDecay:
    """tf.keras.experimental.CosineDecay(initial_learning_rate, decay_steps) (train.py:128-130)."""

    def __init__(self, initial_learning_rate: float, decay_steps: int):
        self.initial_learning_rate, self.decay_steps = float(initial_learning_rate), int(decay_steps)


class Adamax:
    """tf.keras.optimizers.Adamax(learning_rate=schedule_or_float) hyper-parameters."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon


class NVAE:
    def __init__(self, n_encoder_channels, n_decoder_channels, res_cells_per_group, n_preprocess_blocks,
                 n_preprocess_cells, n_latent_per_group, n_latent_scales, n_groups_per_scale, n_postprocess_blocks,
                 n_post_process_cells, sr_lambda, scale_factor, total_epochs, n_total_iterations, step_based_warmup,
                 input_shape, *, training: bool = True, precision: Optional[int] = None, seed: int = 1,
                 device: Optional[str] = None, process_group=None, **kwargs):
        precision = _lib.default_precision() if precision is None else precision
        self.rt = rt = Runtime(device=device, precision=precision, seed=seed)
        self.training = training
        self.process_group = process_group
        input_shape = [int(v) for v in input_shape]
        n_groups_per_scale = [int(g) for g in n_groups_per_scale]  # CLI values arrive as str (train.py:201-206)
        with rt:
            self.sr_lambda = sr_lambda
            self.preprocess = Preprocess(n_encoder_channels, n_preprocess_blocks, n_preprocess_cells, scale_factor,
                                         input_shape)
            self.n_latent_per_group = n_latent_per_group
            self.n_latent_scales = n_latent_scales
            self.n_groups_per_scale = n_groups_per_scale
            self.n_total_iterations = n_total_iterations
            self.n_preprocess_blocks = n_preprocess_blocks
            mult = self.preprocess.mult
            self.encoder = Encoder(n_encoder_channels=n_encoder_channels, n_latent_per_group=n_latent_per_group,
                                   res_cells_per_group=res_cells_per_group, n_latent_scales=n_latent_scales,
                                   n_groups_per_scale=n_groups_per_scale, mult=mult, scale_factor=scale_factor,
                                   input_shape=self.preprocess.output_shape_)
            mult = self.encoder.mult
            enc_shape = list(self.encoder.output_shape_)
            enc_shape[3] = n_encoder_channels * mult  # channels z0 is sampled from
            self.decoder = Decoder(n_decoder_channels=n_decoder_channels, n_latent_per_group=n_latent_per_group,
                                   res_cells_per_group=res_cells_per_group, n_latent_scales=n_latent_scales,
                                   n_groups_per_scale=list(reversed(n_groups_per_scale)), mult=mult,
                                   scale_factor=scale_factor, input_shape=enc_shape)
            mult = self.decoder.mult
            self.postprocess = Postprocess(n_postprocess_blocks, n_post_process_cells, scale_factor=scale_factor,
                                           mult=mult, n_channels_decoder=n_decoder_channels,
                                           in_channels=self.decoder.out_channels)
        rt.finalize()
        self.epoch = 0  # updated at the start of each epoch
        self.total_epochs = total_epochs
        self.step_based_warmup = step_based_warmup
        self.steps = 0  # updated for each gradient pass
        self.optimizer: Optional[Adamax] = None
        alphas = self.calculate_kl_alphas(self.n_latent_scales, self.n_groups_per_scale)
        self._alphas = torch.tensor(alphas, dtype=torch.float32, device=rt.device)
        self._counters = torch.zeros(2, dtype=torch.int64, device=rt.device)  # {warm-up metric, optimizer iters}
        rt.counters = self._counters
        self._hyper = torch.zeros(8, device=rt.device)
        self._beta_one = torch.ones(8, device=rt.device)
        self._m = self._v = None
        self._graph = None
        # data parallel: identical weights on every rank (same `seed`) but independent epsilon streams -- the Philox key
        # folds the rank in, so the replicas of a global batch do not draw the same noise
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rt.philox_seed = (seed + 0x9E3779B1 * torch.distributed.get_rank(process_group)) & 0xFFFFFFFFFFFFFFFF

    # ---- Keras-model surface ---------------------------------------------------------------------
    @property
    def trainable_weights(self):
        return self.rt.trainable_variables

    def count_params(self) -> int:
        return self.rt.n_trainable()

    def compile(self, optimizer: Adamax, run_eagerly: bool = True, **kwargs) -> None:
        self.optimizer = optimizer
        self._m = torch.zeros_like(self.rt.params)
        self._v = torch.zeros_like(self.rt.params)

    def save_weights(self, path: str) -> None:
        """Keras `save_weights` (train.py:51): variables under the reference's attribute paths plus what a TF-format
        checkpoint also carries -- the optimizer slots (Adamax m / v per variable, `optimizer/m/<name>`) and its
        iteration counter -- and the warm-up position (`steps`, `epoch`; train.py:133-135 re-derives them from the
        file name, here they travel with the file)."""
        extra = {"__steps": self.steps, "__epoch": self.epoch, "__optimizer_iterations": int(self._counters[1].item())}
        if self._m is not None:
            m, v = self._m.detach().cpu().numpy(), self._v.detach().cpu().numpy()
            for var in self.rt.trainable_variables:
                extra["__optimizer/m/" + var.name] = m[var.offset:var.offset + var.size].reshape(var.shape)
                extra["__optimizer/v/" + var.name] = v[var.offset:var.offset + var.size].reshape(var.shape)
        np.savez(path, **extra, **self.rt.named_values())

    def load_weights(self, path: str) -> None:
        """Restores variables, and -- when the file has them and compile() ran -- the Adamax slots, the optimizer
        iteration count (CosineDecay position, bias correction) and steps / epoch, so a resumed run continues the
        schedules where the saved one stopped."""
        with np.load(path if path.endswith(".npz") else path + ".npz") as f:
            self.rt.load_named({k: f[k] for k in f.files if not k.startswith("__")})
            if "__steps" in f.files:
                self.steps, self.epoch = int(f["__steps"]), int(f["__epoch"])
            if "__optimizer_iterations" in f.files:
                self._counters[1] = int(f["__optimizer_iterations"])
            if self._m is not None:
                for var in self.rt.trainable_variables:
                    for slot, arena in (("m", self._m), ("v", self._v)):
                        key = f"__optimizer/{slot}/{var.name}"
                        if key in f.files:
                            arena[var.offset:var.offset + var.size].copy_(
                                torch.as_tensor(np.asarray(f[key], dtype=np.float32).ravel()))
        self._host_metric = None  # the device warm-up counter is re-synchronised on the next step / replay

    def on_epoch_begin(self, epoch, logs=None):
        self.epoch = epoch

    def _as_device(self, x) -> DeviceTensor:
        if isinstance(x, DeviceTensor):
            return x
        if isinstance(x, torch.Tensor):
            return DeviceTensor(x.to(self.rt.device, torch.float32).contiguous(), needs_grad=False)
        return self.rt.from_host(x)

    # ---- forward (models.py:89-98) ------------------------------------------------------------------
    def __call__(self, inputs, nll=False, training: Optional[bool] = False):
        rt = self.rt
        inputs = self._as_device(inputs)
        training = bool(training)
        did_sn = False
        if not rt.sn_done and (training or rt.pack is not None):
            # all 163 power iterations in 4 launches, before any conv runs; in the tensor-core modes the same
            # pass (or, for inference, its pack-only form) refreshes the TF32 operand copies of the kernels
            rt.spectral_normalize_all(power_iter=training)
            rt.sn_done = did_sn = True
        try:
            x = self.preprocess(inputs, training)
            enc_dec_combiners, final_x = self.encoder(x, training)
            enc_dec_combiners.reverse()  # flip bottom-up to top-down
            reconstruction, z_params, log_p, log_q = self.decoder(final_x, enc_dec_combiners, nll=nll,
                                                                  training=training)
            self._post_mark = len(rt.tape) if rt.tape is not None else None  # first tape entry of the postprocess tower
            reconstruction = self.postprocess(reconstruction, training)
        finally:
            if did_sn:
                rt.sn_done = False
        return reconstruction, z_params, log_p, log_q

    call = __call__

    # ---- train step (models.py:100-135) -----------------------------------------------------------
    def _schedule(self, advance: bool) -> None:
        """beta (models.py:121-122) and the Adamax/CosineDecay step size from DEVICE counters, so the
        whole step can be replayed as a CUDA graph."""
        opt = self.optimizer
        lr = opt.learning_rate if opt is not None else 0.0
        lr0, decay = (lr.initial_learning_rate, float(lr.decay_steps)) if isinstance(lr, CosineDecay) else (float(lr), 0.0)
        adv = 0
        if advance:
            adv = 3 if self.step_based_warmup else 2
        self.rt.lib.schedule_step(self._counters.data_ptr(), self._hyper.data_ptr(), 0.3 * self.n_total_iterations, lr0,
                                  decay, opt.beta_1 if opt is not None else 0.9, adv, self.rt.stream)

    def _sync_counters(self) -> None:
        metric = self.steps if self.step_based_warmup else self.epoch
        cur = getattr(self, "_host_metric", None)
        if cur != metric:
            self._counters[0] = int(metric)
        self._host_metric = metric

    def train_step(self, data, apply_gradients: bool = True, _on_bucket=None) -> Dict[str, torch.Tensor]:
        """One training step.  Returns the reference's dict (models.py:130-135) of DEVICE tensors:
        loss [], reconstruction_loss [B], kl_loss [B], bn_loss []."""
        if isinstance(data, tuple):
            data = data[0]  # labeled data: drop the label
        rt = self.rt
        data = self._as_device(data)
        B = data.shape[0]
        self._sync_counters()
        # Philox stream ids restart every step (the device iteration counter separates the steps), so an eager step and a
        # replay of the captured graph draw the same epsilons
        rt.eps_i = 0
        rt.lib.fill(rt.grads.data_ptr(), rt.grads.numel(), 0.0, rt.stream)
        self._schedule(advance=True)
        with rt.gradient_tape() as tape:
            reconstruction, z_params, *_ = self(data, training=self.training)
            recon_loss = self.calculate_recon_loss(data, reconstruction)
            bn_loss = self.calculate_bn_loss()
            # beta warm-up, KL balancing while beta < 1, mean over the batch and the total in one launch
            samp = self.decoder.sampler
            kl_loss = rt.empty(B)
            scalars = rt.empty(2)
            rt.lib.loss_assemble(samp.kl_all.data_ptr(), recon_loss.data_ptr(), bn_loss.data_ptr(),
                                 self._alphas.data_ptr(), self._hyper.data_ptr(), -1, samp.n_groups, B,
                                 samp.kl_weight.data_ptr(), kl_loss.data_ptr(), scalars.data_ptr(), rt.stream)
        # tape.gradient(total_loss, trainable_weights): seed d(total)/d(logits), replay, add the BN-gamma term
        N, H, W, Cl = reconstruction.shape
        reconstruction.grad = rt.empty(N, H, W, Cl)
        rt.lib.bernoulli_ll_bwd(reconstruction.ptr(), data.ptr(), N, H, W, data.shape[3], Cl, 1.0 / B,
                                reconstruction.grad.data_ptr(), rt.stream)
        rt.backward(tape, split_at=self._post_mark if _on_bucket is not None else None, on_split=_on_bucket)
        if rt.bn_loss_n:
            rt.lib.bn_loss_bwd(rt.params.data_ptr(), rt.grads.data_ptr(), rt.bn_loss_offsets.data_ptr(),
                               rt.bn_loss_sizes.data_ptr(), rt.bn_loss_n, float(self.sr_lambda), rt.stream)
        if apply_gradients:
            self.apply_gradients()
        self.steps += 1
        if self.step_based_warmup:
            self._host_metric = self.steps
        return {"loss": scalars[0], "reconstruction_loss": recon_loss, "kl_loss": kl_loss, "bn_loss": bn_loss[0]}

    def apply_gradients(self) -> None:
        """NCCL all-reduce of the flat gradient arena (data parallel) + one multi-tensor Adamax launch."""
        rt = self.rt
        if self.optimizer is None:
            raise RuntimeError("call compile(optimizer=Adamax(...)) before training")
        world = self._world()
        if world > 1:
            torch.distributed.all_reduce(rt.grads, group=self.process_group)
        self._adamax(world)

    def _world(self) -> int:
        if self.process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return torch.distributed.get_world_size(self.process_group)
        return 1

    def _adamax(self, world: int) -> None:
        rt = self.rt
        opt = self.optimizer
        rt.lib.adamax(rt.params.data_ptr(), rt.grads.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                      rt.params.numel(), self._hyper.data_ptr(), opt.beta_1, opt.beta_2, opt.epsilon, 1.0 / world,
                      rt.stream)

    # ---- CUDA-graph replay of the whole step --------------------------------------------------------
    def capture_train_step(self, batch_shape, warmup: int = 2):
        """Warm the allocator/workspace eagerly, then capture fill+SN+fwd+loss+bwd(+Adamax) as one CUDA graph.
        Returns (static_input, replay): replay() runs one step on whatever static_input holds.  With more
        than one rank the graph ends after backward; the NCCL all-reduce and the Adamax launch follow it."""
        rt = self.rt
        static_in = torch.zeros(tuple(batch_shape), device=rt.device)
        if rt.eps_injected is not None:
            raise RuntimeError("graph capture draws epsilon on device (Philox); clear the injected epsilons")
        world = torch.distributed.get_world_size(self.process_group) if torch.distributed.is_initialized() else 1
        # more than one rank: the graph ends after backward; the all-reduce and the Adamax launch are two eager calls per
        # replay.  (Capturing the NCCL all-reduce into the graph was measured at 2 GPUs: 31.4 vs 31.3 ms/step, no gain, and
        # the process group then hangs at teardown -- not adopted.)
        in_graph = world == 1
        # the main chain (forward, dgrad, BN / SE backward) is the critical path; the weight-gradient side stream only
        # has to finish by the optimizer.  A high-priority capture stream makes the captured kernel nodes win the CTA
        # scheduler whenever both have work (NVAE_STREAM_PRIO=0: equal priorities)
        prio = -1 if os.environ.get("NVAE_STREAM_PRIO", "1") != "0" else 0
        stream = torch.cuda.Stream(device=rt.device, priority=prio)
        stream.wait_stream(torch.cuda.current_stream(rt.device))
        # the warm-up steps size the allocator / workspaces with REAL launches on an all-zero batch; everything they
        # touch (weights, Adamax slots, BN moving statistics, SN u, schedule counters) is put back afterwards, so capture
        # leaves the model exactly as it found it
        torch.cuda.synchronize(rt.device)
        snap = [t.clone() for t in (rt.params, rt.state, self._counters, rt.sn_sigma)] if warmup > 0 else None
        snap_mv = [self._m.clone(), self._v.clone()] if (warmup > 0 and self._m is not None) else None
        steps0 = self.steps
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                out = self.train_step(static_in)
        torch.cuda.current_stream(rt.device).wait_stream(stream)
        torch.cuda.synchronize(rt.device)
        if snap is not None:
            for dst, src in zip((rt.params, rt.state, self._counters, rt.sn_sigma), snap):
                dst.copy_(src)
            if snap_mv is not None:
                self._m.copy_(snap_mv[0])
                self._v.copy_(snap_mv[1])
            self.steps = steps0
            self._host_metric = None
            torch.cuda.synchronize(rt.device)
        self._sync_counters()
        # Kernel nodes record the priority of the stream they were captured on, but a plainly instantiated graph runs
        # every node at the LAUNCH stream's priority: the library instantiates the captured cudaGraph_t with
        # cudaGraphInstantiateFlagUseNodePriority so the main chain really outranks the weight-gradient side stream
        # -- opt-in, NVAE_GRAPH_NODE_PRIO=1: measured 30.6 / 30.8 ms with, 30.7 / 31.1 ms without (run-to-run noise)
        node_prio = prio != 0 and os.environ.get("NVAE_GRAPH_NODE_PRIO", "0") == "1" and \
            hasattr(torch.cuda.CUDAGraph, "raw_cuda_graph")
        graph = torch.cuda.CUDAGraph(keep_graph=True) if node_prio else torch.cuda.CUDAGraph()
        launches0, kernels0 = rt.lib.launches, rt.lib._nvae_launch_count()
        # More than one rank: the step is TWO graphs.  The first ends when backward has left the postprocess tower (the tail
        # of the gradient arena, 35 % of the parameters, produced first); its bucket is all-reduced on NCCL's stream while
        # the second graph -- the rest of backward -- runs, so only the second bucket's exchange is exposed
        # (NVAE_DP_OVERLAP=0: one graph, one all-reduce after it).
        overlap = (not in_graph) and os.environ.get("NVAE_DP_OVERLAP", "1") != "0"
        graph2 = None
        if overlap:
            graph2 = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(rt.device)
            with torch.cuda.stream(stream):
                graph.capture_begin()

                def next_graph():
                    graph.capture_end()
                    graph2.capture_begin(pool=graph.pool())
                try:
                    out = self.train_step(static_in, apply_gradients=False, _on_bucket=next_graph)
                finally:
                    (graph2 if self._post_mark is not None else graph).capture_end()
            torch.cuda.current_stream(rt.device).wait_stream(stream)
            post0 = next(v.offset for v in rt.trainable_variables if v.name.startswith("postprocess/"))
            self._dp_buckets = (rt.grads[post0:], rt.grads[:post0])
        else:
            with torch.cuda.graph(graph, stream=stream):
                out = self.train_step(static_in, apply_gradients=in_graph)
        graph_exec = None
        if node_prio:
            import ctypes
            ex = ctypes.c_void_p()
            rc = rt.lib._nvae_graph_instantiate(ctypes.c_void_p(graph.raw_cuda_graph()), 1, ctypes.byref(ex))
            if rc != 0:
                raise _lib.NvaeError(f"nvae_graph_instantiate failed: cudaError_t {rc}")
            graph_exec = ex
            self._graph_exec = (ex, rt.lib)  # kept alive with the model
        self._graph2 = graph2
        self.graph_kernels = rt.lib._nvae_launch_count() - kernels0 + (0 if in_graph else 1)  # + Adamax outside
        self.steps -= 1  # capture records the launches without running them
        self._host_metric = self.steps if self.step_based_warmup else self.epoch
        self.graph_launches = rt.lib.launches - launches0 + (0 if in_graph else 1)
        self._graph = graph

        def replay():
            # host-side changes of the warm-up metric (on_epoch_begin with epoch-based warm-up, `model.steps = n` after
            # a resume) reach the device counter here: a small H2D write outside the graph, only when it changed
            self._sync_counters()
            if graph_exec is not None:
                rc = rt.lib._nvae_graph_launch(graph_exec, rt.stream)
                if rc != 0:
                    raise _lib.NvaeError(f"nvae_graph_launch failed: cudaError_t {rc}")
            elif graph2 is not None:
                graph.replay()
                first = torch.distributed.all_reduce(self._dp_buckets[0], group=self.process_group, async_op=True)
                graph2.replay()
                torch.distributed.all_reduce(self._dp_buckets[1], group=self.process_group)
                first.wait()
                self._adamax(world)
            else:
                graph.replay()
            if not in_graph and graph2 is None:
                self.apply_gradients()
            self.steps += 1
            if self.step_based_warmup:
                self._host_metric = self.steps
            return out
        return static_in, replay

    def make_train_function(self, batch_shape):
        """Keras `make_train_function` analogue for HOST batches: returns f(host_batch) -> dict of host arrays.
        Each call copies the batch host->device (pinned staging), replays the captured step and reads the
        four losses back device->host -- the end-to-end path bench.py's `e2e` measures.

        A batch of a different size than `batch_shape` (the reference's last batch of an epoch is 96 of 144,
        SURVEY 3.1) gets its own captured graph the first time it is seen (cached per shape); capture leaves the
        model state untouched, so switching shapes mid-epoch is exact."""
        fns = {}

        def build(shape):
            rt = self.rt
            static_in, replay = self.capture_train_step(shape)
            staging = torch.empty(tuple(shape), dtype=torch.float32).pin_memory()
            B = shape[0]
            host_out = torch.empty(2 * B + 2, dtype=torch.float32).pin_memory()
            dev_out = torch.empty(2 * B + 2, device=rt.device)

            def run(batch):
                if isinstance(batch, torch.Tensor) and batch.is_pinned():
                    static_in.copy_(batch, non_blocking=True)
                else:
                    staging.copy_(torch.as_tensor(batch, dtype=torch.float32))
                    static_in.copy_(staging, non_blocking=True)
                out = replay()
                dev_out[0:1].copy_(out["loss"].reshape(1))
                dev_out[1:2].copy_(out["bn_loss"].reshape(1))
                dev_out[2:2 + B].copy_(out["reconstruction_loss"])
                dev_out[2 + B:].copy_(out["kl_loss"])
                host_out.copy_(dev_out, non_blocking=True)
                torch.cuda.current_stream(rt.device).synchronize()
                h = host_out.numpy()
                return {"loss": float(h[0]), "bn_loss": float(h[1]), "reconstruction_loss": h[2:2 + B],
                        "kl_loss": h[2 + B:]}
            run.static_in, run.replay = static_in, replay
            return run

        main_shape = tuple(int(v) for v in batch_shape)
        fns[main_shape] = build(main_shape)

        def train_function(batch):
            shape = tuple(int(v) for v in batch.shape)
            fn = fns.get(shape)
            if fn is None:
                fn = fns[shape] = build(shape)
            return fn(batch)
        train_function.static_in, train_function.replay = fns[main_shape].static_in, fns[main_shape].replay
        train_function.h2d_bytes = int(np.prod(main_shape)) * 4
        train_function.d2h_bytes = (2 * main_shape[0] + 2) * 4
        train_function.captured_shapes = fns
        return train_function

    # ---- sampling (models.py:137-189) ---------------------------------------------------------------------
    def sample(self, n_samples=16, temperature=1.0, greyscale=True):
        with self.rt.weights_ready():
            return self._sample(n_samples, temperature, greyscale)

    def _sample(self, n_samples, temperature, greyscale):
        rt = self.rt
        dec = self.decoder
        s = R.broadcast_batch(rt, dec.h, n_samples)
        z0_shape = (n_samples, *dec.z0_shape)
        # mu = softclamp5(0) = 0, sigma = exp(softclamp5(0)) + 1e-2 = 1.01; only z0 is tempered (models.py:143-144)
        mu = rt.zeros(*z0_shape)
        sigma = rt.empty(*z0_shape)
        rt.lib.fill(sigma.data_ptr(), sigma.numel(), 1.0 + 1e-2, rt.stream)
        z_t = dec.sampler.sample(mu, sigma, float(temperature) if temperature != 1.0 else 1.0)
        if temperature != 1.0:
            rt.lib.axpby(sigma.data_ptr(), float(temperature), sigma.data_ptr(), 0.0, sigma.numel(), rt.stream)
        z = DeviceTensor(z_t, needs_grad=False)
        decoder_index = 0
        last_s = None
        for layer in dec.groups:
            if isinstance(layer, DecoderSampleCombiner):
                if decoder_index > 0:
                    raw = dec.sampler.get_params(dec.sampler.dec_sampler, decoder_index, s)
                    B, h, w, L2 = raw.shape
                    dist = rt.empty(4, B, h, w, L2 // 2)
                    z = DeviceTensor(rt.empty(B, h, w, L2 // 2), needs_grad=False)
                    kl = rt.empty(B)
                    # mu = softclamp5(m), sigma = exp(softclamp5(ls)) + 1e-2, z = mu + eps*sigma: the latent
                    # kernel's z_idx==0 branch applied to the decoder parameters (models.py:154-159)
                    rt.lib.latent_fwd(raw.ptr(), None, rt.next_eps((B, h, w, L2 // 2)).data_ptr(), B, h * w, L2 // 2,
                                      z.ptr(), kl.data_ptr(), None, None, dist.data_ptr(), rt.stream)
                    mu, sigma = dist[0], dist[1]
                last_s = s
                s = layer(s, z)
                decoder_index += 1
            else:
                s = layer(s)
        reconstruction = self.postprocess(s)
        images = self._bernoulli_images(reconstruction, greyscale)
        z1 = dec.sampler.sample(mu, sigma)
        z2 = dec.sampler.sample(mu, sigma)
        # images and the last hierarchical z's (mu, sigma, s) so sample_with_z can re-render (models.py:177-178)
        return images, last_s, z1, z2

    def capture_sample(self, n_samples=16, temperature=1.0, greyscale=True, warmup: int = 1):
        """`sample()` as ONE CUDA graph (BASELINE configs[3]: ~330 launches for 1024 images).  Returns replay() ->
        (images, last_s, z1, z2), the same static device buffers every call.  The epsilons come from Philox keyed by a
        device counter that a node of the graph advances, so every replay draws fresh noise; the operand copies of the
        kernels are refreshed inside the graph, so weights moved by training are picked up."""
        rt = self.rt
        if rt.eps_injected is not None:
            raise RuntimeError("graph capture draws epsilon on device (Philox); clear the injected epsilons")
        if not hasattr(self, "_sample_counters"):
            self._sample_counters = torch.zeros(2, dtype=torch.int64, device=rt.device)
            self._sample_hyper = torch.zeros(8, device=rt.device)
        stream = torch.cuda.Stream(device=rt.device)
        stream.wait_stream(torch.cuda.current_stream(rt.device))

        def body():
            rt.eps_i = 0
            prev, rt.counters = rt.counters, self._sample_counters
            try:
                out = self.sample(n_samples, temperature, greyscale)
                # ++counters[1]: the next replay's Philox step
                rt.lib.schedule_step(self._sample_counters.data_ptr(), self._sample_hyper.data_ptr(), 0.0, 0.0, 0.0, 0.9,
                                     2, rt.stream)
            finally:
                rt.counters = prev
            return out
        with torch.cuda.stream(stream):
            for _ in range(max(warmup, 1)):
                body()
        torch.cuda.current_stream(rt.device).wait_stream(stream)
        torch.cuda.synchronize(rt.device)
        graph = torch.cuda.CUDAGraph()
        k0 = rt.lib._nvae_launch_count()
        with torch.cuda.graph(graph, stream=stream):
            out = body()
        self.sample_graph_kernels = rt.lib._nvae_launch_count() - k0
        self._sample_graph = graph

        def replay():
            graph.replay()
            return out
        return replay

    def _bernoulli_images(self, logits: DeviceTensor, greyscale: bool = True) -> torch.Tensor:
        """distributions.Bernoulli(logits).probs_parameter()/mean() = sigmoid(l); sample() = U < sigmoid(l)."""
        rt = self.rt
        out = rt.empty(*logits.shape)
        rt.lib.bernoulli_image(logits.ptr(), out.numel(), 0 if greyscale else 1, rt.philox_seed, rt.eps_i,
                               out.data_ptr(), rt.stream)
        if not greyscale:
            rt.eps_i += 1
        return out

    def sample_with_z(self, z, s):
        last_gen_layer = self.decoder.groups[-1]
        z = z if isinstance(z, DeviceTensor) else DeviceTensor(z, needs_grad=False)
        with self.rt.weights_ready():
            s = last_gen_layer(s, z)
            reconstruction = self.postprocess(s)
        return self._bernoulli_images(reconstruction, True)

    # ---- losses (models.py:191-267) -----------------------------------------------------------------
    def calculate_kl_loss(self, z_params: List[DistributionParams], balancing) -> torch.Tensor:
        """-> [B].  The per-group row sums already exist (fused into the latent kernel); this runs the
        balancing of models.py:204-218 (or the plain sum :219-222) in one launch."""
        rt = self.rt
        samp = self.decoder.sampler
        if samp.kl_all is not None and len(z_params) == samp.kl_all.shape[0] and \
                z_params[0].kl.data_ptr() == samp.kl_all.data_ptr():
            kl_all = samp.kl_all  # the rows the latent kernel wrote, already stacked [G,B]
        else:
            kl_all = torch.stack([p.kl for p in z_params], 0).contiguous()
        G, B = kl_all.shape
        out = rt.empty(B)
        rt.lib.loss_assemble(kl_all.data_ptr(), None, None, self._alphas.data_ptr(), self._beta_one.data_ptr(),
                             1 if balancing else 0, G, B, None, out.data_ptr(), None, rt.stream)
        return out

    def calculate_kl_alphas(self, num_scales, groups_per_scale) -> np.ndarray:
        """Balancer coefficients with the square decay function (models.py:227-237)."""
        coeffs = []
        for i in range(num_scales):
            n = groups_per_scale[num_scales - i - 1]
            coeffs.append(np.square(2 ** i) / n * np.ones(n, np.float32))
        coeffs = np.concatenate(coeffs, 0)
        return coeffs / coeffs.min()

    def calculate_recon_loss(self, inputs, reconstruction: DeviceTensor, crop_output=False) -> torch.Tensor:
        rt = self.rt
        inputs = self._as_device(inputs)
        N, H, W, Cl = reconstruction.shape
        out = rt.empty(N)
        rt.lib.bernoulli_ll_fwd(reconstruction.ptr(), inputs.ptr(), N, H, W, inputs.shape[3], Cl,
                                2 if crop_output else 0, out.data_ptr(), rt.stream)
        return out

    def calculate_bn_loss(self) -> torch.Tensor:
        """sr_lambda * sum over encoder/decoder-group BN layers of max|gamma| -- one launch over the arena."""
        rt = self.rt
        out = rt.zeros(1)
        if rt.bn_loss_n:
            rt.lib.bn_loss_fwd(rt.params.data_ptr(), rt.bn_loss_offsets.data_ptr(), rt.bn_loss_sizes.data_ptr(),
                               rt.bn_loss_n, float(self.sr_lambda), out.data_ptr(), rt.stream)
        return out
