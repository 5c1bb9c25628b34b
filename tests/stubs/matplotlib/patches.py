from . import _Noop


def __getattr__(name):
    return _Noop()
