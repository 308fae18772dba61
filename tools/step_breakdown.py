#!/usr/bin/env python
"""Per-kernel time of one eager NVAE train step via torch.profiler (CUPTI), warm caches, real overlap.
usage (GPU box): python tools/step_breakdown.py [batch] > gpurun_out/breakdown.txt
Cheaper than an ncu launch list (seconds instead of minutes); ncu remains the source for profiles/."""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nvae_tf_b200.models import NVAE, Adamax, CosineDecay  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 144
    torch.cuda.set_device(0)
    m = NVAE(**bench.mirror_kwargs(B), training=True, seed=1)
    m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 1000)))
    m.steps = 20000
    x = torch.as_tensor(bench.synthetic_images(B, 1)).cuda()
    for _ in range(2):
        m.train_step(x)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.train_step(x)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t_lo = min(e.time_range.start for e in evs)
    t_hi = max(e.time_range.end for e in evs)
    per_stream = defaultdict(float)
    main_stream = None
    for e in evs:
        sid = getattr(e, "device_resource_id", getattr(e, "device_index", 0))
        if main_stream is None:
            main_stream = sid  # the first kernel of the step (gradient fill) runs on the main stream
        a = agg[("M " if sid == main_stream else "S ") + e.name[:100]]
        a[0] += 1
        a[1] += e.device_time
        a[2] = max(a[2], e.device_time)
        per_stream[getattr(e, "device_resource_id", getattr(e, "device_index", 0))] += e.device_time
    tot = sum(a[1] for a in agg.values())
    print(f"batch {B}: {sum(a[0] for a in agg.values())} kernels, {tot / 1e3:.2f} ms summed device time, "
          f"{(t_hi - t_lo) / 1e3:.2f} ms first-start to last-end (eager, incl. launch gaps); per stream: "
          + ", ".join(f"{k}: {v / 1e3:.2f} ms" for k, v in sorted(per_stream.items())))
    for name, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:5d}  avg {t / n:8.1f} us  max {mx:8.1f} us  {name}")


if __name__ == "__main__":
    main()
