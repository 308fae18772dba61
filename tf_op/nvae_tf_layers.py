"""TensorFlow side of the drop-in: `tf.custom_gradient` wrappers over the custom ops of tf_op/nvae_ops.cc, shaped so
that the reference's layers (common.py, encoder.py, decoder.py) call them where they call Keras layers today and
`tf.GradientTape` in `NVAE.train_step` (models.py:116-127) works unchanged.

Needs TensorFlow (the reference pins 2.3) and `make -C tf_op` at the user's site; this repository's build image has
neither, so this module is shipped as source (byte-compiled by tests/test_tf_op_source.py) and the kernels it reaches are
tested through the same C ABI with ctypes (tests/test_kernels_gpu.py, tests/test_model_gpu.py).

Activation codes (include/nvae_b200.h): 0 none, 1 swish, 2 ELU.  Precision 2 = NVAE_PREC_TF32X3 (fp32-grade products).
"""
import os

import tensorflow as tf

_HERE = os.path.dirname(os.path.abspath(__file__))
_ops = tf.load_op_library(os.path.join(_HERE, "libnvae_tf_ops.so"))
_EMPTY = tf.zeros([0], tf.float32)


def conv2d(x, kernel, kernel_tr, bias, stride=1, x2=None, residual=None, precision=2):
    """SpectralNormalization(Conv2D(padding="same"))(concat(x, x2)) (+ residual).  `kernel` is the (already normalised)
    HWIO variable, `kernel_tr` its transposed operand copy written by the once-per-step nvae_spectral_norm pass."""
    x2_ = _EMPTY if x2 is None else x2
    res_ = _EMPTY if residual is None else residual

    @tf.custom_gradient
    def f(x, x2_, kernel, bias, res_):
        y = _ops.nvae_conv2d_fwd(x=x, x2=x2_, w=kernel, w_tr=kernel_tr, bias=bias, residual=res_, stride=stride,
                                 precision=precision)

        def grad(dy):
            dx, dx2 = _ops.nvae_conv2d_dgrad(dy=dy, w=kernel, w_rnd=kernel, in_h=x.shape[1], in_w=x.shape[2],
                                             cin=x.shape[3], cin2=0 if x2 is None else x2.shape[3], stride=stride,
                                             precision=precision)
            dw, db = _ops.nvae_conv2d_wgrad(x=x, x2=x2_, dy=dy, r=kernel.shape[0], s=kernel.shape[1], stride=stride,
                                            precision=precision)
            # straight-through to the normalised kernel variable (tfa.SpectralNormalization semantics)
            return dx, (None if x2 is None else dx2), dw, db, (None if residual is None else dy)
        return y, grad
    return f(x, x2_, kernel, bias, res_)


def bn_act(x, bn, act=1, training=True, upsample=False):
    """activation(BatchNormalization(momentum=0.05, epsilon=1e-5)(x)) [-> tf.image.resize nearest x2]: common.py:165-172,
    encoder.py:102-104, decoder.py:139,143.  `bn` is the Keras layer (gamma, beta, moving_mean, moving_variance)."""
    @tf.custom_gradient
    def f(x, gamma, beta):
        out, stat = _ops.nvae_bn_fwd(x=x, gamma=gamma, beta=beta, moving_mean=bn.moving_mean,
                                     moving_var=bn.moving_variance, training=training, momentum=bn.momentum,
                                     epsilon=bn.epsilon, act=act, upsample=upsample)

        def grad(dout):
            return _ops.nvae_bn_act_bwd(dout=dout, x=x, stat=stat, training=training, act=act, upsample=upsample)
        return out, grad
    return f(x, bn.gamma, bn.beta)


def bn_stat(x, bn, training=True):
    """Only the [4, C] statistics block (mean, invstd, scale, shift) for consumers that fuse the BN-apply into their own
    load (depthwise conv, squeeze-excitation).  The gradient flows through those consumers' `da` / `dt` and
    nvae_bn_act_bwd (see dwconv_bn_act)."""
    _, stat = _ops.nvae_bn_fwd(x=x, gamma=bn.gamma, beta=bn.beta, moving_mean=bn.moving_mean, moving_var=bn.moving_variance,
                               training=training, momentum=bn.momentum, epsilon=bn.epsilon, act=0, upsample=False)
    return tf.stop_gradient(stat)


def dwconv_bn_act(x, bn, depthwise_kernel, bias, act=1, training=True):
    """DepthwiseConv2D((5, 5), padding="same")(swish(BN(x))) with the BN-apply + activation fused into the load
    (decoder.py:130,141-142)."""
    @tf.custom_gradient
    def f(x, gamma, beta, w, b):
        stat = bn_stat(x, bn, training)
        y = _ops.nvae_dwconv5x5_fwd(x=x, stat=stat, w=w, bias=b, act=act)

        def grad(dy):
            da, dw, db = _ops.nvae_dwconv5x5_bwd(x=x, stat=stat, w=w, dy=dy, act=act)
            dx, dgamma, dbeta = _ops.nvae_bn_act_bwd(dout=da, x=x, stat=stat, training=training, act=act, upsample=False)
            return dx, dgamma, dbeta, dw, db
        return y, grad
    return f(x, bn.gamma, bn.beta, depthwise_kernel, bias)


def bn_conv2d(x, bn, kernel, kernel_tr, bias, act=1, training=True, residual=None, precision=2):
    """SpectralNormalization(Conv2D(1x1))(act(BN(x))) (+ residual) with the BN-apply + activation inside the convolution's
    operand path, forward and backward-filter: decoder.py:125-127 (act=0), 143-144; postprocess.py:71-73 (act=0), 84-96.
    For the shapes of those call sites; NvaeConv2dFwdBnact raises InvalidArgument for any other -- compose
    conv2d(bn_act(x, ...), ...) there."""
    res_ = _EMPTY if residual is None else residual

    @tf.custom_gradient
    def f(x, gamma, beta, kernel, bias, res_):
        stat = bn_stat(x, bn, training)
        y = _ops.nvae_conv2d_fwd_bnact(x=x, stat=stat, w=kernel, w_tr=kernel_tr, bias=bias, residual=res_, act=act,
                                       precision=precision)

        def grad(dy):
            da, _ = _ops.nvae_conv2d_dgrad(dy=dy, w=kernel, w_rnd=kernel, in_h=x.shape[1], in_w=x.shape[2], cin=x.shape[3],
                                           cin2=0, stride=1, precision=precision)
            dw, db = _ops.nvae_conv2d_wgrad_bnact(x=x, stat=stat, dy=dy, act=act, precision=precision)
            dx, dgamma, dbeta = _ops.nvae_bn_act_bwd(dout=da, x=x, stat=stat, training=training, act=act, upsample=False)
            return dx, dgamma, dbeta, dw, db, (None if residual is None else dy)
        return y, grad
    return f(x, bn.gamma, bn.beta, kernel, bias, res_)


def se_residual(t, xres, se, alpha=0.1, beta=1.0, bn=None, training=True):
    """alpha * xres + beta * SqueezeExcitation(BN(t)) (bn optional): common.py:129-142 with the cell tails encoder.py:107,
    decoder.py:147 (alpha=0.1, beta=1) and preprocess.py:107, postprocess.py:58 (alpha=1, beta=0.1)."""
    w1, b1, w2, b2 = se.dense1.kernel, se.dense1.bias, se.dense2.kernel, se.dense2.bias
    gamma = _EMPTY if bn is None else bn.gamma
    beta_ = _EMPTY if bn is None else bn.beta

    @tf.custom_gradient
    def f(t, xres, w1, b1, w2, b2, gamma, beta_):
        stat = _EMPTY if bn is None else bn_stat(t, bn, training)
        y, pooled, hidden, gate = _ops.nvae_se_fwd(t=t, stat=stat, xres=xres, w1=w1, b1=b1, w2=w2, b2=b2, alpha=alpha,
                                                   beta=beta)

        def grad(dy, *_unused):
            dt, dxres, dw1, db1, dw2, db2 = _ops.nvae_se_bwd(dy=dy, t=t, stat=stat, w1=w1, w2=w2, pooled=pooled,
                                                             hidden=hidden, gate=gate, alpha=alpha, beta=beta)
            if bn is None:
                return dt, dxres, dw1, db1, dw2, db2, None, None
            dx, dgamma, dbeta = _ops.nvae_bn_act_bwd(dout=dt, x=t, stat=stat, training=training, act=0, upsample=False)
            return dx, dxres, dw1, db1, dw2, db2, dgamma, dbeta
        return y, grad
    return f(t, xres, w1, b1, w2, b2, gamma, beta_)


def latent(enc_p, dec_p, eps, kl_weight):
    """Sampler.call (common.py:76-102): returns (z, kl[B], dist[4, B, HW, L]).  `dec_p=None` is the z_idx == 0 branch.
    `kl_weight` (one float: beta * balance coefficient / B, models.py:204-222) scales the KL term's gradient."""
    dec_ = _EMPTY if dec_p is None else dec_p

    @tf.custom_gradient
    def f(enc_p, dec_):
        z, kl, dist = _ops.nvae_latent_fwd(enc_p=enc_p, dec_p=dec_, eps=eps)

        def grad(dz, dkl, ddist):
            d_enc, d_dec = _ops.nvae_latent_bwd(enc_p=enc_p, dec_p=dec_, eps=eps, dz=dz, kl_weight=kl_weight)
            return d_enc, (None if dec_p is None else d_dec)
        return (z, kl, dist), grad
    return f(enc_p, dec_)


# ---------------------------------------------------------------------------------------------------------------------
# Once-per-step launchers around the cells (models.py:116-129, 191-267; train.py:128-131)
# ---------------------------------------------------------------------------------------------------------------------
def spectral_norm_all(arena, training=True):
    """tfa.SpectralNormalization(power_iterations=1) of EVERY wrapped conv in one op (4 launches) + the operand repack the
    tensor-core convolutions read (`kernel_tr`).  `arena` bundles what nvae_tf_b200.runtime.Runtime lays out once at
    model build: flat `params` / `state` / `pack` resource variables (the Keras variables are views of them), the
    NvaeSnLayer table and its chunk index as constant tensors, and a float scratch variable.

    Sequencing under tf.function: the op mutates the arenas, so everything that reads a kernel must run after it.
    `NVAE.call` does

        sigma = spectral_norm_all(self.arena, training)
        with tf.control_dependencies([sigma]):
            x = self.preprocess(inputs) ...

    -- one control edge ahead of the first conv, exactly where the reference's first `SpectralNormalization.call`
    would have normalised its kernel."""
    return _ops.nvae_spectral_norm(params=arena.params, state=arena.state, pack=arena.pack, layers=arena.sn_layers,
                                   chunk_layer=arena.sn_chunk_layer, ws=arena.sn_ws, power_iter=training, pack_exact=True)


def bn_stats(x, bn, training=True):
    """[4, C] statistics block only (moving statistics updated in place), via the dedicated launcher."""
    return tf.stop_gradient(_ops.nvae_bn_stats(x=x, gamma=bn.gamma, beta=bn.beta, moving_mean=bn.moving_mean,
                                               moving_var=bn.moving_variance, training=training, momentum=bn.momentum,
                                               epsilon=bn.epsilon))


def bernoulli_recon_loss(inputs, logits, crop_output=False):
    """NVAE.calculate_recon_loss (models.py:242-250): -sum_{h,w,c} Bernoulli(logits).log_prob(x) -> [B]."""
    @tf.custom_gradient
    def f(logits):
        ll = _ops.nvae_bernoulli_ll_fwd(logits=logits, x=inputs, crop=2 if crop_output else 0)

        def grad(d):  # d(recon[b]) / d(logits) = sigmoid(l) - x, scaled per sample by the upstream gradient
            g = _ops.nvae_bernoulli_ll_bwd(logits=logits, x=inputs, scale=1.0)
            return g * tf.reshape(d, [-1, 1, 1, 1])
        return ll, grad
    return f(logits)


def loss_assemble(kl_all, recon, bn_loss, alphas, hyper, balancing=-1):
    """models.py:121-126, 204-222 in one launch: (kl_weight[G], kl_loss[B], [total, mean(recon + kl_loss)])."""
    return _ops.nvae_loss_assemble(kl_all=kl_all, recon=recon, bn_loss=bn_loss, alphas=alphas, hyper=hyper,
                                   balancing=balancing)


def bn_loss(arena, sr_lambda):
    """NVAE.calculate_bn_loss (models.py:252-267) over the flat parameter arena; its sub-gradient is added to the
    gradient arena by bn_loss_backward (tf.reduce_max splits evenly among ties)."""
    return _ops.nvae_bn_loss_fwd(params=arena.params, offsets=arena.bn_loss_offsets, sizes=arena.bn_loss_sizes,
                                 sr_lambda=sr_lambda)


def bn_loss_backward(arena, sr_lambda):
    return _ops.nvae_bn_loss_bwd(params=arena.params, grads=arena.grads, offsets=arena.bn_loss_offsets,
                                 sizes=arena.bn_loss_sizes, sr_lambda=sr_lambda)


def schedule_step(arena, warmup_iters, lr0, decay_steps, beta_1=0.9, advance=3):
    """beta warm-up (models.py:121-122) and the CosineDecay / Adamax bias-corrected step size (train.py:128-130) from the
    device counters {warm-up metric, optimizer iterations}; returns hyper[8] = {beta, lr_t, lr, t, metric, ...}."""
    return _ops.nvae_schedule_step(counters=arena.counters, warmup_iters=warmup_iters, lr0=lr0, decay_steps=decay_steps,
                                   beta_1=beta_1, advance=advance)


def adamax_apply(arena, hyper, beta_1=0.9, beta_2=0.999, epsilon=1e-7, grad_scale=1.0):
    """optimizer.apply_gradients (models.py:128) as ONE multi-tensor launch over the arenas; grad_scale = 1 / replicas."""
    return _ops.nvae_adamax(params=arena.params, grads=arena.grads, m=arena.adamax_m, v=arena.adamax_v, hyper=hyper,
                            beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, grad_scale=grad_scale)


def philox_normal(arena, n, seed, stream_id=0):
    """tf.random.normal stand-in for Sampler.sample's epsilon (common.py:67), reproducible under graph replay."""
    return _ops.nvae_philox_normal(counters=arena.counters[1:], n=n, seed=seed, stream_id=stream_id)


def reparam(mu, sigma, eps, sigma_scale=1.0):
    """Sampler.sample on materialised parameters: mu + eps * sigma * sigma_scale (common.py:65-68, models.py:140-145)."""
    return _ops.nvae_reparam(mu=mu, sigma=sigma, eps=eps, sigma_scale=sigma_scale)


def iwae_nll(recon, log_q, log_p):
    """evaluate.py:118-121: -mean_b(logsumexp_k(-recon - log_q + log_p) - log K) from [K, B] stacks -> (per_sample, nll)."""
    return _ops.nvae_iwae_nll(recon=recon, log_q=log_q, log_p=log_p)
