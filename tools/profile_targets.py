"""One pass over every kernel family for `ncu --set full` (run under gpurun; see profiles/).
  1 dense scatter (1e7 events -> 256x256), 1 cnt2event (2 x 2 x 512 x 512, ~1.5e5 events per sample),
  1 pipeline step of cfg2 WITHOUT the CUDA graph (every kernel of the network appears as its own launch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import encodings as enc                     # noqa: E402
from esr_b200.expand import expand                         # noqa: E402
from esr_b200.model import DeepRecurrNet                   # noqa: E402
from esr_b200.pipeline import EventSRPipeline              # noqa: E402
import bench                                                  # noqa: E402  (synthetic weights)

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
n, H = 10**7, 256
xs = torch.randint(0, H, (n,), generator=g, device=dev).float()
ys = torch.randint(0, H, (n,), generator=g, device=dev).float()
ps = (torch.randint(0, 2, (n,), generator=g, device=dev) * 2 - 1).float()
off = torch.tensor([0, n], dtype=torch.int64, device=dev)
for _ in range(2):
    enc.encode_frames(xs, ys, ps, off, hr_size=(H, H), n_max_frame=n)
cnt = torch.poisson(torch.full((2, 2, 512, 512), 0.3, device=dev), generator=g)
for _ in range(2):
    expand(cnt, 0, 0)

B, L, lr, scale = 8, 8, (128, 128), 2
net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
net.load_state_dict(bench.synth_weights(0))
net = net.to(dev).eval()
pipe = EventSRPipeline(net, B, L, lr, scale, dev)
ne = B * L * 2048
exs = torch.randint(0, lr[1], (ne,), generator=g, device=dev).float()
eys = torch.randint(0, lr[0], (ne,), generator=g, device=dev).float()
eps = (torch.randint(0, 2, (ne,), generator=g, device=dev) * 2 - 1).float()
eoff = torch.arange(0, ne + 1, 2048, dtype=torch.int64, device=dev)
pipe.sr_bias = torch.poisson(torch.full(((L - 2) * B, 2, 256, 256), 0.3, device=dev), generator=g)
for _ in range(3):
    pipe.run_device(exs, eys, eps, eoff, 2048)
torch.cuda.synchronize()
print("done")
