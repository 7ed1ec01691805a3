"""Drop-in for the reference's util/io.py (wire formats): JSON / text helpers and the SoccerNet
`results_spotting.json` writers, byte-compatible with util/io.py:14-68 (same keys in the same order, same `indent`,
positions in milliseconds, `gameTime` without zero padding — the SoccerNet evaluator parses exactly this)."""
import json
import os

FPS_SN = 25


def load_json(fpath):
    with open(fpath) as fp:
        return json.load(fp)


def store_json(fpath, obj, pretty=False):
    opts = dict(indent=2, sort_keys=True) if pretty else {}
    with open(fpath, 'w') as fp:
        json.dump(obj, fp, **opts)


def load_text(fpath):
    """Non-empty stripped lines of a text file."""
    with open(fpath, 'r') as fp:
        return [line for line in (raw.strip() for raw in fp) if line]


def _spotting_record(event, half, stride):
    """One entry of `predictions` (util/io.py:31-38 / :52-59): frame index at FPS_SN/stride -> milliseconds."""
    position = int(event['frame'] / FPS_SN * 1000 * stride)
    return {'gameTime': '{} - {}:{}'.format(half, position // 60000, int((position % 60000) // 1000)),
            'label': event['label'], 'position': position, 'confidence': event['score'], 'half': half}


def _write_game(pred_path, sub_dir, game_dict):
    path = os.path.join('/'.join(pred_path.split('/')[:-1]) + '/preds', sub_dir)
    os.makedirs(path, exist_ok=True)
    with open(path + '/results_spotting.json', 'w') as fp:
        json.dump(game_dict, fp, indent=4)


def store_json_sn(pred_path, pred, stride=1):
    """SoccerNet v2: consecutive list entries are the two halves of one game (`<game>/1`, `<game>/2`); one file per game,
    written when its second half has been added, `UrlLocal` = the first half's video name (util/io.py:22-46)."""
    game_dict = None
    for i, game in enumerate(pred):
        half = i % 2 + 1
        if half == 1:
            game_dict = {'UrlLocal': game['video'], 'predictions': []}
        game_dict['predictions'].extend(_spotting_record(e, half, stride) for e in game['events'])
        if half == 2:
            _write_game(pred_path, '/'.join(game['video'].split('/')[:-1]), game_dict)


def store_json_snb(pred_path, pred, stride=1):
    """SoccerNet Ball: one single-half game per list entry (util/io.py:48-66)."""
    for game in pred:
        _write_game(pred_path, game['video'],
                    {'UrlLocal': game['video'], 'predictions': [_spotting_record(e, 1, stride) for e in game['events']]})
