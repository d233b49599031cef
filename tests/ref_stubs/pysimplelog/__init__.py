class SingleLogger(object):
    """no-op stand-in for pysimplelog.SingleLogger"""
    def __init__(self, *args, **kwargs):
        self.custom_init()
    def custom_init(self):
        pass
    def __getattr__(self, name):
        def _noop(*args, **kwargs):
            if name in ("error", "critical") and args:
                return Exception(str(args[-1]))
            return None
        return _noop
    def log(self, *args, **kwargs):
        return None
    def error(self, message, *a, **k):
        return Exception(message)
    def critical(self, message, *a, **k):
        return Exception(message)
Logger = SingleLogger
