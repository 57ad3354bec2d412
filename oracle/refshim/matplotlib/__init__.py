"""Inert stub: the reference imports matplotlib for plots only (synthetic_data_gen.py:41,63-80)."""
