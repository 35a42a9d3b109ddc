"""Minimal stand-in for moke-config==0.2.1 (reference requirements.txt:31): ConfigBase + create_config."""


class ConfigBase:
    pass


def _overlay(obj, d):
    for k, v in (d or {}).items():
        cur = getattr(obj, k, None)
        if isinstance(v, dict) and isinstance(cur, ConfigBase):
            _overlay(cur, v)
        else:
            setattr(obj, k, v)
    return obj


def create_config(cls, d=None):
    return _overlay(cls(), d)
