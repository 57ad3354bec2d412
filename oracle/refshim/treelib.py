"""Inert stub: treelib is used only by dead code (scone_trajectory_model.py:155-206; call site
commented out at trajectory_experiments.py:508-510)."""


class Tree:
    pass
