"""bf16 engine vs the reference-generated golden vectors: relative-to-max error of logits / displacement per model case
(the numbers quoted in DESIGN.md section 2; tolerances are in tests/test_gpu_model.py)."""
import os, sys, numpy as np, torch
for _p in ('/root/repo/t-deed_b200', '/root/repo/oracle', '/root/repo', '/root/repo/tests'):
    sys.path.insert(0, _p)
os.chdir('/root/repo')
import importlib.util
spec = importlib.util.spec_from_file_location('tgm', 'tests/test_gpu_model.py'); tgm = importlib.util.module_from_spec(spec); spec.loader.exec_module(tgm)
O = tgm.O
dev = torch.device('cuda')
worst = [0, 0]
for name in sorted(tgm.MODEL_CASES):
    kw, _, wseed, _ = tgm.MODEL_CASES[name]
    g = np.load(os.path.join('tests/golden', 'model_%s.npz' % name))
    cfg = O.Config(**kw)
    m = tgm.build(cfg, O.random_state(cfg, wseed), dev)
    frames = torch.from_numpy(g['frames']).to(dev)
    m._model.eval()
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        pred, _ = m._model(frames, inference=True)
    logits = pred['im_feat'] if isinstance(pred, dict) else pred
    e1 = tgm.rel_err(logits.cpu().numpy(), g['logits'])
    e2 = tgm.rel_err(pred['displ_feat'].cpu().numpy(), g['displ']) if isinstance(pred, dict) else 0.0
    worst = [max(worst[0], e1), max(worst[1], e2)]
    print('%-28s logits %.4f displ %.4f' % (name, e1, e2))
print('worst', worst)
