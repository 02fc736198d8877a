"""In-kernel cycle counters of the fused Fcomb kernel (CTA 0): build the variant first,
    python tools/build_variant.py fcprof -DFC_PROFILE [-DFC_H2_RING=2]
then run this with PDA_B200_LIB pointing at it.  Prints cycles per (tile, sample) for every role."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from oracle import punet_oracle as po
from probabilistic_domain_adaptation_b200 import _lib, ops

lib = _lib.load()
dev = torch.device('cuda:0')
sd = po.make_state_dict(0, last_layer_gain=1.0)
k = ["fcomb.layers.0", "fcomb.layers.2", "fcomb.last_layer"]
w = [sd[f"{n}.{p}"].to(dev).contiguous() for n in k for p in ("weight", "bias")]
g = torch.Generator().manual_seed(0)
B, H, S = 4, 1024, 16
feat = torch.relu(torch.randn(B, H, H, 64, generator=g)).to(torch.float16).to(dev)
z = torch.randn(S, B, 6, generator=g).to(dev)
fn = ctypes.CDLL(_lib.LIB_PATH).pda_fcomb_profile_read
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
for _ in range(2):
    ops.fcomb_mc_consensus(feat, z, *w, want_mask=True)
torch.cuda.synchronize()
fn(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.fcomb_mc_consensus(feat, z, *w, want_mask=True)
e1.record()
torch.cuda.synchronize()
fn(buf, 0)
tiles_per_cta = (B * H * H // 128) / 296.0
n = tiles_per_cta * S
names = {0: "producer: wait A1 slot", 1: "producer: relu(H1+bz) -> TMEM -> arrive", 2: "epilogue: wait H2", 3: "epilogue: convert (ld, cvt, st, arrive)",
         4: "epilogue: wait D3", 5: "epilogue: finish (ld, sigmoid, stores)", 8: "control: wait A1", 9: "control: wait H2 buffer",
         10: "control: issue MMA2", 11: "control: last-layer MMA (wait A3 + issue)"}
print(f"kernel pair {e0.elapsed_time(e1):.3f} ms; {n:.0f} (tile, sample) steps per CTA")
for i, name in names.items():
    print(f"  {name:48s} {buf[i] / n:8.1f} cycles per step")
