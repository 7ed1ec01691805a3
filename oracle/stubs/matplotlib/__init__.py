"""Empty stand-in for matplotlib (util/score.py:9 of the reference imports pyplot)."""
