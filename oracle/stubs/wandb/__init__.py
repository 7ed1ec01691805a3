"""Offline stand-in for wandb so the reference's train_tdeed.py can be driven in tests."""


class _Summary(dict):
    pass


summary = _Summary()


def login(*a, **k):
    return True


def init(*a, **k):
    return None


def log(*a, **k):
    return None
