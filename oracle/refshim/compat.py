"""Restore names removed from NumPy / NetworkX since the reference was written (TEST INFRA ONLY).

np.float (synthetic_data_gen.py:294), nx.OrderedDiGraph (:117), nx.readwrite.gpickle (:424-425,436,454).
nx.draw_networkx* are made inert (they import the real matplotlib; reference: synthetic_data_gen.py:70-79).
"""
import pickle
import types
import numpy as np
import networkx as nx


def apply():
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(nx, 'OrderedDiGraph'):
        nx.OrderedDiGraph = nx.DiGraph          # dicts are insertion-ordered since Python 3.7
    if not hasattr(nx.readwrite, 'gpickle'):
        g = types.ModuleType('networkx.readwrite.gpickle')

        def write_gpickle(G, path):
            with open(path, 'wb') as f:
                pickle.dump(G, f, protocol=4)

        def read_gpickle(path):
            with open(path, 'rb') as f:
                return pickle.load(f)
        g.write_gpickle, g.read_gpickle = write_gpickle, read_gpickle
        nx.readwrite.gpickle = g
    nx.draw_networkx = lambda *a, **k: None
    nx.draw_networkx_edges = lambda *a, **k: None
