"""See matplotlib/__init__.py: every attribute is a no-op callable."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        raise RuntimeError("matplotlib stub: pyplot.{} was called (visualisation must stay off in the dry run)".format(name))
    return _noop
