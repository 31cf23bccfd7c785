"""Renders the eleven scenes of the reference's main.rs (their own cameras and SampleSettings, adaptive sampling) on the
GPU and writes small PNGs + a contact sheet: the "existing scenes render" evidence.  usage: python scripts/gpu_gallery.py <out_dir> [width]"""
import os
import sys
import time

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402

out = sys.argv[1]
width = int(sys.argv[2]) if len(sys.argv) > 2 else 320
os.makedirs(out, exist_ok=True)
tiles = []
for k in range(12):
    name = f"scene{k}"
    dev = gs.DeviceScene(gs.HostScene.build(name, width=width), device=0)
    t = time.perf_counter()
    rgb, counts, st = dev.render(seed=1)
    ms = (time.perf_counter() - t) * 1e3
    img = gs.tonemap(rgb)   # write_color's transfer, color.rs:8-26
    Image.fromarray(img).save(os.path.join(out, f"{name}.png"), optimize=True)
    print(f"{name}: {dev.width}x{dev.height}, {counts.mean():.0f} spp on average ({counts.min()}..{counts.max()}), {st.paths / 1e6:.1f} M paths, "
          f"{st.kernel_ms:.1f} ms kernel, {ms:.1f} ms call", flush=True)
    tiles.append(Image.fromarray(img).resize((240, max(1, 240 * dev.height // dev.width))))
    dev.close()
h = max(t.height for t in tiles)
sheet = Image.new("RGB", (240 * 4, h * 3))
for i, t in enumerate(tiles):
    sheet.paste(t, (240 * (i % 4), h * (i // 4)))
sheet.save(os.path.join(out, "contact_sheet.png"), optimize=True)
