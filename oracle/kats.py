"""Known-answer vectors -- TEST INFRASTRUCTURE ONLY.

The reference ships no expected outputs, so these are derived by hand from the
reference formulas on the fixed inputs of its docstrings (integer-valued, so
exact in fp32).  They pin index conventions (pair order, W x vs x W, kernel
axis order), not TF rounding.  Each KAT is (name, inputs dict, expected).
"""
from __future__ import annotations

import numpy as np


def kat1_inner_product():
    """InnerProductNetwork docstring input (2.FM/CustomLayers.py:602-607):
    x = arange(24).reshape(2,3,4); out[b,p] = <x_i,x_j>, pairs (0,1),(0,2),(1,2).
    FM second order on the same tensor == row sums (property FM2 == sum_p IPN)."""
    x = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    ipn = np.array([[38.0, 62.0, 214.0], [950.0, 1166.0, 1510.0]])
    fm2 = np.array([[314.0], [3626.0]])
    return x, ipn, fm2


def kat2_field_aware():
    """FieldAwareInteractionLayer (2.FM/CustomLayers.py:428-462) with
    T[v,f,d] = 100v + 10f + d, V=6, F=3, k=2, X=[[0,2,4],[1,3,5]]."""
    V, F, k = 6, 3, 2
    T = (100.0 * np.arange(V)[:, None, None] + 10.0 * np.arange(F)[None, :, None]
         + np.arange(k)[None, None, :])
    X = np.array([[0, 2, 4], [1, 3, 5]], dtype=np.int64)
    pair_vectors = np.array([
        [[2000.0, 2211.0], [8000.0, 8421.0], [90200.0, 90831.0]],
        [[33000.0, 33411.0], [60000.0, 60621.0], [163200.0, 164031.0]],
    ])
    ffm_term = np.array([201663.0, 514263.0])
    return T, X, pair_vectors, ffm_term


def kat3_cross_vector():
    """CrossLayer (3.DCN/CustomLayers.py:195-203), 2 layers."""
    x0 = np.array([[1.0, 2.0], [3.0, -1.0]])
    w = [np.array([[0.5], [-1.0]]), np.array([[2.0], [0.25]])]
    b = [np.array([[0.1], [0.2]]), np.array([[-0.3], [0.4]])]
    out = np.array([[-1.7, -2.4], [71.425, -23.275]])
    return x0, w, b, out


def kat4_cross_matrix():
    """MatrixCrossLayer (3.DCN/CustomLayers.py:297-305): y = W x (X W^T in
    batch form) -- the values differ under X W."""
    x0 = np.array([[1.0, 2.0], [3.0, -1.0]])
    W = [np.array([[1.0, 2.0], [3.0, 4.0]]), np.array([[0.5, -1.0], [0.0, 2.0]])]
    b = [np.array([[0.1], [0.2]]), np.array([[-0.3], [0.4]])]
    out = np.array([[-15.55, 122.8], [33.45, 5.8]])
    return x0, W, b, out


def kat5_outer_product_mat():
    """OuterProductNetwork('mat') (2.FM/CustomLayers.py:670-680):
    x = arange(12).reshape(2,3,2), K[a,p,c] = 1 + a + 10p + 100c;
    out[b,p] = sum_{a,c} x_i[c] K[a,p,c] x_j[a] = einsum('bpc,apc,bpa->bp')."""
    x = np.arange(12, dtype=np.float64).reshape(2, 3, 2)
    k, P = 2, 3
    K = (1.0 + np.arange(k)[:, None, None] + 10.0 * np.arange(P)[None, :, None]
         + 100.0 * np.arange(k)[None, None, :])
    out = np.array([[508.0, 1004.0, 3670.0], [12238.0, 17846.0, 26584.0]])
    return x, K, out
