import os, sys, torch
sys.path.insert(0, '/root/repo/t-deed_b200')
from model.modules import SGPBlock, SGPMixer
from tdeed_b200 import ops
from tdeed_b200.engine import sgp_up_size
dev = torch.device('cuda')
B, T, C, ks, r = [int(v) for v in sys.argv[1:6]]
up = sgp_up_size(ks, r)
blk = SGPBlock(C, kernel_size=ks, k=r, init_conv_vars=0.1).to(dev).eval()
mixer = SGPMixer(C, kernel_size=ks, k=r, init_conv_vars=0.1, t_size=T).to(dev).eval()
x = torch.randn((B, T, C), device=dev)
xc = torch.randn((B, (T + 1) // 2, C), device=dev)
w, mw = blk.mix_weights(), mixer.mix_weights()
for _ in range(3):
    ops.sgp_mix(x, T, ks, up, w, torch.bfloat16)
    ops.sgp_mixer_mix(xc, x, ks, up, mw, torch.bfloat16)
torch.cuda.synchronize()
