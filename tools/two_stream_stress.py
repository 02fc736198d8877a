"""Stress tool: two models run their MC inference on two CUDA streams at the same time, so that
the CTAs of different launches interleave.  python tools/two_stream_stress.py <image size> <iterations>.  Results must be
bit-identical to the serial run.  (Before the slab-ring protocol fix of the fused up-sampling conv this failed from
2 x 512 x 512 on: a bounded mbarrier wait expired after ~2 s.)"""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import punet_oracle as po
from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, consensus
dev = torch.device('cuda:0')
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ms = []
for seed in (0, 1):
    m = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 1.0).to(dev).eval()
    m.load_state_dict(po.make_state_dict(seed, last_layer_gain=8.0))
    ms.append(m)
x, _, eps, _ = po.synthetic_inputs(2, size, size, s=16)
x, eps = x.to(dev), eps.to(dev)
serial = [consensus.sample_from_teacher(m, x, 16, do_consensus_masking=True, eps=eps) for m in ms]
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
t0 = time.time()
try:
    for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 10):
        outs = []
        for m, st in zip(ms, streams):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(consensus.sample_from_teacher(m, x, 16, do_consensus_masking=True, eps=eps))
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        ok = all(torch.equal(o[0], s[0]) and torch.equal(o[1], s[1]) for o, s in zip(outs, serial))
        if not ok:
            print('iteration', it, 'DIFFERS'); break
    else:
        print(size, 'two streams: ok, bit-identical to serial;', round(time.time() - t0, 2), 's')
except Exception as e:
    print(size, 'FAILED after', round(time.time() - t0, 2), 's:', str(e)[:80])
