def LoadJsonFromZip(*a, **k):
    raise NotImplementedError('SoccerNet stub')
