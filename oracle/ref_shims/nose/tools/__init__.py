def eq_(a, b, msg=None):
    assert a == b, msg or f"{a!r} != {b!r}"


def ok_(x, msg=None):
    assert x, msg


def assert_almost_equal(a, b, places=7, msg=None):
    assert round(abs(a - b), places) == 0, msg or f"{a!r} !~ {b!r}"
