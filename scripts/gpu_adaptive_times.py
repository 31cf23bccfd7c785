"""GPU-only wall times of gs_render on the reference's adaptive scenes (main.rs settings), best of 3.
usage: [GS_TAIL_PIXELS=n] python scripts/gpu_adaptive_times.py [scene ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402

names = sys.argv[1:] or ["scene1", "scene3", "scene5", "scene6", "scene8", "scene10", "scene11"]
line = f"tail_pixels={os.environ.get('GS_TAIL_PIXELS', 'default'):8s}"
for name in names:
    dev = gs.DeviceScene(gs.HostScene.build(name), device=0)
    dev.render(seed=1)
    best = None
    for rep in range(3):
        t = time.perf_counter()
        _, counts, st = dev.render(seed=2)
        dt = time.perf_counter() - t
        best = dt if best is None or dt < best else best
    line += f" | {name} {best * 1e3:7.2f} ms ({int(counts.sum()) / 1e6:.1f} M paths, {st.kernel_launches} launches)"
    dev.close()
print(line, flush=True)
