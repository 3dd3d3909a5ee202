"""TEST INFRASTRUCTURE: a do-nothing stand-in for matplotlib, so that the reference's example scripts (which plot their
results) can be RUN UNMODIFIED in an image without matplotlib (tests/test_reference_python_overlay.py)."""


class _Anything:
    def __call__(self, *args, **kwargs):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __getitem__(self, key):
        return _Anything()

    def __setitem__(self, key, value):
        pass


def __getattr__(name):
    return _Anything()
