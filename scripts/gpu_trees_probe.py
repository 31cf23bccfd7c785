import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import grayshift_b200 as gs
for name, kw in (("c4", {"width": 200, "spp": 16}), ("c1", {"width": 200, "spp": 16}), ("scene7", {"width": 120, "spp": 16})):
    host = gs.HostScene.build(name, **kw)
    dev = gs.DeviceScene(host)
    dev.set_option("bvh_width", 2); a, ca, sa = dev.render(seed=5)
    dev.set_option("bvh_width", 8); b, cb, sb = dev.render(seed=5)
    same = np.isclose(a, b, rtol=1e-3, atol=1e-5).all(axis=2).mean()
    print(os.environ.get("GS_VARIANT", "default"), name, int(sa.rays), int(sb.rays), abs(int(sa.rays) - int(sb.rays)) / int(sa.rays), same, flush=True)
