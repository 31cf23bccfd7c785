"""Which tree wins at which scene size: the config-5 sphere cloud at several sphere counts, binary vs eight-wide tree
(kernel time of gs_render, best of 3).  Justifies kWideMinPrims in grayshift_b200/csrc/gs_host.h."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402

for n in (3000, 10000, 30000, 100000, 300000, 1000000):
    host = gs.HostScene.build("c5", width=1920, spp=8, n_spheres=n)
    dev = gs.DeviceScene(host, device=0)
    line = f"{n:8d} spheres"
    for w in (2, 8):
        dev.set_option("bvh_width", w)
        best = None
        for rep in range(4):
            _, _, st = dev.render(seed=1 + rep, want_counts=False)
            if rep and (best is None or st.kernel_ms < best[0]):
                best = (st.kernel_ms, st.paths)
        line += f" | width {w}: {best[1] / best[0] / 1e3:7.1f} Mpaths/s"
    line += f" | default {dev.info('nodes8')} wide nodes, depth {dev.info('wide_depth')}"
    print(line, flush=True)
    dev.close()
