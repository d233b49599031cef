class InfiniteBoundaries(object):
    pass
class PeriodicBoundaries(object):
    pass
