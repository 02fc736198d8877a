import os, sys, torch
sys.path.insert(0, os.getcwd())
from oracle import punet_oracle as po
from probabilistic_domain_adaptation_b200 import ops
dev = torch.device("cuda:0")
sd = po.make_state_dict(0, last_layer_gain=4.0)
g = torch.Generator().manual_seed(23)
b, h, w_ = 1, 16, 8
feat = torch.relu(torch.randn(b, h, w_, 64, generator=g)).to(dev).to(torch.bfloat16)
z = torch.randn(b, 6, generator=g).to(dev)
keys = ["fcomb.layers.0", "fcomb.layers.2", "fcomb.last_layer"]
w = [sd[f"{n}.{p}"].to(dev).contiguous() for n in keys for p in ("weight", "bias")]
go = torch.randn(b, 1, h, w_, generator=g).to(dev)
try:
    out = ops.fcomb_bwd(feat, z, w[0], w[1], w[2], w[3], w[4], go)
    torch.cuda.synchronize()
    print("ok", out[0].float().abs().sum().item())
except Exception as e:
    print("ERR", repr(e))
    try:
        torch.cuda.synchronize()
    except Exception as e2:
        print("SYNC ERR", repr(e2))
