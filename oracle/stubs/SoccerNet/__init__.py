"""Empty stand-in so the reference's util/eval.py and train_tdeed.py import offline (test infrastructure)."""
