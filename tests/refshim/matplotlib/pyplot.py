"""Plot calls of the reference (interface_wholebody_qref.py:480-716) become no-ops.  TEST INFRASTRUCTURE."""


class _Nop:
    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Nop()
