"""Inert stub for matplotlib.pyplot (plots are cosmetic; reference: synthetic_data_gen.py:63-80)."""


class _Nop:
    def __getattr__(self, name):
        return lambda *a, **k: _Nop()


def __getattr__(name):
    return lambda *a, **k: _Nop()
