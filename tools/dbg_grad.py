import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import punet_oracle as po
from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, l2_regularisation, ops
dev = torch.device('cuda:0')
g = torch.load('tests/golden/train_bce_64x64.pt', weights_only=False)
def run(simt):
    ops.FORCE_SIMT_CONV = simt
    m = ProbabilisticUnet(1, 1, [64,128,256,512], 6, 3, 1.0, consensus_masking=False, rl_swap=g['rl_swap']).to(dev)
    m.load_state_dict(po.make_state_dict(0, last_layer_gain=4.0)); m.train()
    x, segm = g['x'].to(dev), g['segm'].to(dev)
    m.forward(x, segm, training=True)
    d = m.posterior_latent_space
    z = d.base_dist.loc + d.base_dist.scale * g['eps_post'].to(dev)
    m.posterior_latent_space.rsample = lambda *a, **k: z
    elbo = m.elbo(segm, None)
    reg = l2_regularisation(m.posterior) + l2_regularisation(m.prior) + l2_regularisation(m.fcomb.layers)
    (-elbo + 1e-5*reg).backward()
    return {k: p.grad.clone() for k, p in m.named_parameters()}, z.detach().clone()
a, za = run(False)
b, zb = run(True)
a2, _ = run(False)
print("z diff tc vs simt", (za-zb).abs().max().item())
for k in a:
    rn = g['grad_norms'][k]
    ra, rb = a[k].norm().item()/rn, b[k].norm().item()/rn
    cos = torch.nn.functional.cosine_similarity(a[k].flatten().double(), b[k].flatten().double(), dim=0).item()
    rep = (a[k]-a2[k]).abs().max().item() / (a[k].abs().max().item()+1e-30)
    flag = '***' if abs(ra-1) > 0.1 or abs(rb-1) > 0.1 else ''
    print(f"{k:55s} tc/ref {ra:6.3f} simt/ref {rb:6.3f} cos(tc,simt) {cos:8.5f} rerun-rel {rep:8.1e} {flag}")
