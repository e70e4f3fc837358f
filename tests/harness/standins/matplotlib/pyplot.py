"""No-op pyplot: every attribute is a function that accepts anything and returns a harmless object."""


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Anything()
