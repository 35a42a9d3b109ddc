from . import eq_, ok_, assert_almost_equal  # noqa: F401
