"""The reference's eleven scenes at their own sizes and SampleSettings: wall time of the whole gs_render call on the GPU
next to the CPU oracle (all host threads) on the cheap ones; the expensive ones are extrapolated from the oracle's
measured rate on the same scene at a capped sample count.  usage: python scripts/gpu_reference_scenes.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import grayshift_b200 as gs  # noqa: E402
import oracle_ffi as orc  # noqa: E402

cores = os.cpu_count() or 1
print(f"host threads {cores}")
for k in range(12):
    name = f"scene{k}"
    dev = gs.DeviceScene(gs.HostScene.build(name), device=0)
    dev.render(seed=1)
    t = time.perf_counter()
    rgb, counts, st = dev.render(seed=2)
    gpu_s = time.perf_counter() - t
    paths = int(counts.sum())
    ora = orc.OracleScene.build(name)
    if paths <= 12_000_000:
        _, oc, c, sec = ora.render(seed=2, rng=orc.OracleScene.WYRAND, threads=cores)
        cpu_s, how = sec, f"measured ({int(oc.sum()) / 1e6:.1f} M paths)"
    else:  # rate on the same scene with the sampler capped at one batch, scaled to the GPU's path count
        probe = orc.OracleScene.build(name).set_sampling(0.95, 0.0, dev.settings.batch_size, 0)
        _, oc, c, sec = probe.render(seed=2, rng=orc.OracleScene.WYRAND, threads=cores)
        cpu_s, how = paths / (int(oc.sum()) / sec), f"extrapolated from {int(oc.sum()) / sec / 1e6:.2f} M paths/s"
    print(f"{name}: {dev.width}x{dev.height}, {paths / 1e6:8.1f} M paths ({counts.mean():.0f} spp mean) | GPU {gpu_s * 1e3:8.1f} ms | CPU oracle {cpu_s:8.1f} s {how} | x{cpu_s / gpu_s:.0f}",
          flush=True)
    dev.close()
