"""Shift operators of the SCCONV / "bunch" model — host-side mirror of trajectory_analysis/bunch_model_matrices.py.

compute_shift_matrices(B1, B2) returns S_00, S_10, S_01, S_11, S_21, S_12, S_22 (bunch_model_matrices.py:118-135).
The reference builds them with dense inv / pinv of DIAGONAL matrices (O(E^3)); every D is diagonal (:44-69,79-85), so
here they are degree vectors and the operators come out as scipy CSR matrices (scalable, same values up to fp rounding):

    d2_1 = max(|B1| 1, 1)   d2_2 = max(|B2| 1, 1)   d1 = 2 |B1| d2_2   d5 = |B2| 1   D3 = I/3   D4 = I
    S_00 = (D2_1 + I - B1 B1^T) (D2_1 + I)^-1                       (A0u_n)
    S_10 = D1^+ B1                 S_01 = D2_2 B1^T D1^+
    S_11 = (D2_2 + I)(D2_2^-1 - D2_2^-1 B2 D3 B2^T D2_2^-1 + I) + (D2_2 - D2_2 B1^T D1^+ B1 D2_2 + I)(D2_2 + I)^-1
    S_21 = B2 D3                   S_12 = B2^T D5^+                 S_22 = 2 (2 I - B2^T D5^+ B2)
"""
import numpy as np
import scipy.sparse as sp


def _pinv_diag(d):
    out = np.zeros_like(d, dtype=np.float64)
    nz = d != 0
    out[nz] = 1.0 / d[nz]
    return out


def compute_shift_matrices(B1, B2):
    """Returns the seven operators as scipy.sparse CSR (float64); accepts dense or sparse B1 [N,E], B2 [E,F]."""
    B1, B2 = sp.csr_matrix(B1, dtype=np.float64), sp.csr_matrix(B2, dtype=np.float64)
    N, E = B1.shape
    F = B2.shape[1]
    d2_2 = np.maximum(np.asarray(abs(B2).sum(axis=1)).ravel(), 1.0)          # compute_D2(B2)
    d2_1 = np.maximum(np.asarray(abs(B1).sum(axis=1)).ravel(), 1.0)          # compute_D2(B1)
    d1 = 2.0 * np.asarray(abs(B1) @ d2_2).ravel()                            # compute_D1(B1, D2_2)
    d5 = np.asarray(abs(B2).sum(axis=1)).ravel()                             # compute_D5(B2)
    D = sp.diags
    d1p, d5p = _pinv_diag(d1), _pinv_diag(d5)
    I_N, I_E, I_F = sp.identity(N), sp.identity(E), sp.identity(F)
    A0u_n = (D(d2_1) + I_N - B1 @ B1.T) @ D(1.0 / (d2_1 + 1.0))
    L1u = D(d2_2) @ B1.T @ D(d1p) @ B1
    A1u_n = (D(d2_2) - L1u @ D(d2_2) + I_E) @ D(1.0 / (d2_2 + 1.0))
    L1d = (B2 / 3.0) @ B2.T @ D(1.0 / d2_2)
    A1d_n = (D(d2_2) + I_E) @ (D(1.0 / d2_2) - D(1.0 / d2_2) @ L1d + I_E)
    L2d = B2.T @ D(d5p) @ B2
    A2d_n = 2.0 * (2.0 * I_F - L2d)
    mats = (A0u_n, D(d1p) @ B1, D(d2_2) @ B1.T @ D(d1p), A1d_n + A1u_n, B2 / 3.0, B2.T @ D(d5p), A2d_n)
    out = []
    for M in mats:
        M = sp.csr_matrix(M)
        M.sum_duplicates()
        M.sort_indices()
        out.append(M)
    return tuple(out)
