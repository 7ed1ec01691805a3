"""TEST INFRASTRUCTURE — CPU oracle (plain Python) for the mAP scorer, restating util/score.py:16-160 of the reference:
    parse_ground_truth       util/score.py:16-32   (events with explicit 'frame' only: the synthetic sets carry them)
    average_precision        util/score.py:45-89   greedy closest-unrecalled matching in descending-score order
    mean_average_precisions  util/score.py:92-128  per class (sorted by name) x tolerance, mean over classes
Pinned by tests/golden/evaluate.npz (`score/*` entries written by oracle/gen_golden_eval.py from the unmodified
reference) in tests/test_oracle_golden.py, and against the live reference in tests/test_oracle_vs_reference.py.
Only tests/ may import this module.
"""


def parse_ground_truth(truth):
    out = {}
    for video in truth:
        for e in video['events']:
            out.setdefault(e['label'], {}).setdefault(video['video'], []).append(e['frame'])
    return out


def ranked_predictions(pred, label):
    flat = [(v['video'], e['frame'], e['score']) for v in pred for e in v['events'] if e['label'] == label]
    return sorted(flat, key=lambda t: -t[2])                 # stable: ties keep list order


def average_precision(ranked, truth, tolerance):
    total = sum(len(v) for v in truth.values())
    recalled = set()
    precisions = []
    for rank, (video, frame, _) in enumerate(ranked, 1):
        best = None
        for gt in truth.get(video, ()):
            if (video, gt) not in recalled and (best is None or abs(frame - gt) < abs(frame - best)):
                best = gt
        if best is not None and abs(frame - best) <= tolerance:
            recalled.add((video, best))
            precisions.append(len(recalled) / rank)
    running = 0.0
    envelope = []
    for p in reversed(precisions):
        running = max(running, p)
        envelope.append(running)
    # the reference integrates with the builtin sum() over the forward-ordered list (util/score.py:76-89); builtin sum()
    # is Neumaier-compensated from Python 3.12 on, so the same call is used here rather than a hand-rolled loop
    return sum(envelope[::-1]) / total


def mean_average_precisions(truth, pred, tolerances):
    by_label = parse_ground_truth(truth)
    table = {(l, t): average_precision(ranked_predictions(pred, l), by_label[l], t) for l in sorted(by_label) for t in tolerances}
    import numpy as np
    means = [float(np.mean([table[(l, t)] for l in sorted(by_label)])) for t in tolerances]      # util/score.py:121
    return means, table
