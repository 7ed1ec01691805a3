#!/usr/bin/env python3
"""One eager batch of the bench workload between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    import contextlib
    import io
    from model.model import TDEEDModel
    from tdeed_b200.pipeline import VideoScores
    from tdeed_b200 import ops
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
    with contextlib.redirect_stdout(io.StringIO()):
        model = TDEEDModel(device='cuda:0', args=bench.model_args())
    bench.randomize_(model._model, 0)
    model._model.eval()
    eng = model._model.engine(precision)
    x = torch.randint(0, 256, (B, 100, 3, bench.FRAME_H, bench.FRAME_W), dtype=torch.uint8, device='cuda')
    starts = bench.clip_starts(bench.VIDEO_FRAMES)[:B]
    K = 5

    def step():
        vs = VideoScores(bench.VIDEO_FRAMES, K, x.device)
        _, _, probs = eng.forward(x)
        vs.add(probs, starts)
        ev = vs.events(0.01)
        ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], K, 1, 0.01, False)
        ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], K, 3, 0.01, True)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.prof = []
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    # launch-ordered family labels of the GEMM launches, so tools/traffic_from_ncu.py can attribute the ncu rows
    import json
    labels = [(lab, nb) for lab, _, nb, _, _ in eng.prof if lab in ('conv1x1', 'conv1x1_ds', 'sgp_gemm')]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(dict(clips_per_batch=B, precision=precision, gemm_launches=labels),
              open(os.path.join(ROOT, 'gpurun_out', 'profile_step_labels.json'), 'w'))


if __name__ == '__main__':
    main()
