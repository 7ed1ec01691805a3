"""Drop-in for the reference's util/dataset.py: class-list and fps helpers (util/dataset.py:6-21)."""
import os

from util.io import load_text

DATASETS = ['tennis', 'fs_perf', 'fs_comp', 'finediving', 'finegym', 'soccernetv2', 'soccernetball']


def load_classes(file_name):
    """class.txt -> {name: 1-based id} in file order (0 is background)."""
    return {name: idx for idx, name in enumerate(load_text(file_name), start=1)}


def read_fps(video_frame_dir):
    with open(os.path.join(video_frame_dir, 'fps.txt')) as fp:
        return float(fp.read())
