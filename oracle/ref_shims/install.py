"""Make the unmodified reference importable: stub keras/tensorflow modules, py3.9+ semaphore compat.

Test infrastructure only (see README.md).  Usage:  import oracle.ref_shims.install as s; s.install()
"""
import asyncio
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_SRC = "/root/reference/src"


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("keras", "tensorflow")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class _SemCtx:
    def __init__(self, sem):
        self.sem = sem

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        self.sem.release()


def _sem_await(self):
    # `with await self.sem` (agent/player.py:205) was removed in Python 3.9
    yield from self.acquire().__await__()
    return _SemCtx(self)


_installed = False


def available():
    return os.path.isdir(REF_SRC)


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference sources not present (only available in the build container)")
    here = os.path.dirname(os.path.abspath(__file__))
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, here)       # moke_config, nose
    sys.path.insert(0, REF_SRC)    # reversi_zero
    if not hasattr(asyncio.Semaphore, "__await__"):
        asyncio.Semaphore.__await__ = _sem_await
    try:
        asyncio.get_event_loop()
    except RuntimeError:
        asyncio.set_event_loop(asyncio.new_event_loop())
    _installed = True
