def average_mAP(*a, **k):
    raise NotImplementedError('SoccerNet stub')


def evaluate(*a, **k):
    raise NotImplementedError('SoccerNet stub')
