"""Two C4 frames with the fused kernel (tune 0x1000) on a device-resident frame, for
  ncu --set full --clock-control none --import-source on -k regex:k_wf_fused --launch-skip 1 -c 1 -o gpurun_out/prof python tools/profile_fused.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from yahr_b200 import api  # noqa: E402

sc, cam, _ = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c4-terrain")
w, h = api.image_size(cam)
s = api.Scene(sc)
rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
for _ in range(2):
    st = s.render_device(cam, rgb.data_ptr(), None, tune=0x1000, kernel=2)
print("fused frame %.3f ms, %d rays" % (st["gpu_ms"], st["n_primary"] + st["n_shadow"]))
s.close()
