"""Where the end-to-end time of one bench step goes: each stage of the e2e leg of bench.py timed on the host clock
(with a device synchronize after it).  usage: python scripts/gpu_e2e_breakdown.py [config] [spp]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 500
host = gs.HostScene.build(cfg, spp=spp)
host.flatten()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)


def timed(label, fn, acc):
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    acc.setdefault(label, []).append((time.perf_counter() - t) * 1e3)
    return r


acc = {}
for k in range(4):
    sc = timed("gs_scene_create", lambda: gs.DeviceScene(host, device=0), acc)
    n_px = sc.width * sc.height
    a = timed("accumulator alloc+zero", lambda: torch.zeros((n_px, 4), dtype=torch.float32, device=dev), acc)
    timed("gs_render_accumulate", lambda: sc.render_accumulate(a.data_ptr(), seed=1, sample_begin=k * spp, sample_count=spp, stream=stream.cuda_stream), acc)
    timed("gs_resolve (D2H)", lambda: sc.resolve(a.data_ptr(), stream=stream.cuda_stream), acc)
    timed("gs_stats_read", lambda: sc.stats(), acc)
    timed("gs_scene_destroy", lambda: sc.close(), acc)
for label, v in acc.items():
    print(f"{label:28s} first {v[0]:9.2f} ms   later median {sorted(v[1:])[len(v[1:]) // 2]:9.2f} ms")
