class _Locker(object):
    isServer = True
    def __getattr__(self, name):
        return lambda *a, **k: None
class Repository(object):
    """no-op stand-in for pyrep.Repository"""
    def __init__(self, *a, **k):
        self.locker = _Locker()
    def is_repository_file(self, *a, **k):
        return False, False, False, False
    def __getattr__(self, name):
        return lambda *a, **k: None
