#!/usr/bin/env python
"""How well-conditioned is the default-config NVAE step at batch 144?  Runs the oracle forward (training-mode BN, SN, injected
epsilons) in torch-CPU float32 and float64 on the same weights / images / epsilons and prints how far the two drift apart.
usage: python tests/conditioning_probe.py <gamma scale>      (1.0: Keras initial gamma; 0.3: the tests' operating point)
Measured here (8 host cores, ~7 min): gamma 1.0 -> loss 1.8e-4, kl_all 3.2e-4, logits 7.6e-4 (max |z| ~ 300);
gamma 0.3 -> loss 4e-8, kl_all 5e-7, logits 8e-7.  The batch-144 parity tests therefore run at gamma = 0.3."""
import sys, numpy as np, torch, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
from oracle import nvae_oracle as O
torch.set_num_threads(8)
cfg = O.NVAEConfig()
B = 144
params, trainable, bnl, s = O.build_params(cfg, seed=1, jitter=0.05)
params = {k: np.asarray(v, np.float32).astype(np.float64) for k, v in params.items()}
gs=float(sys.argv[1])
for k in params:
    if k.endswith("/gamma"): params[k] = params[k]*gs
rng = np.random.default_rng(2)
x = (rng.random((B, 28, 28, 1)) < 0.13).astype(np.float32)
x = np.pad(x, ((0, 0), (2, 2), (2, 2), (0, 0)))
eps = [np.asarray(e.numpy(), np.float32).astype(np.float64) for e in O.make_eps(s, B, seed=2)]
res={}
for dt in (torch.float32, torch.float64):
    with torch.no_grad():
        ref,_ = O.train_step_loss(cfg, s, O.to_torch(params, [], dtype=dt), bnl, torch.as_tensor(x,dtype=dt), [torch.as_tensor(e,dtype=dt) for e in eps], 20000, training=True)
    res[dt]=ref
    print(dt, float(ref["loss"]), flush=True)
a,b=res[torch.float64],res[torch.float32]
print("gamma scale", gs, "loss rel", abs(float(a["loss"])-float(b["loss"]))/abs(float(a["loss"])))
print("kl_all rel", float((a["kl_all"]-b["kl_all"].double()).abs().max()/a["kl_all"].abs().max()))
print("logits rel", float((a["logits"]-b["logits"].double()).abs().max()/a["logits"].abs().max()), "max logit", float(a["logits"].abs().max()))
