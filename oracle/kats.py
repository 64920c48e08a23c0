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


# ---------------------------------------------------------------------------
# Round 2: gradient / optimizer / loss KATs.  Every number below was derived from the reference
# FORMULAS (SURVEY 8a') with scalar arithmetic independent of oracle/reference_layers.py.
def kat6_fm_gradient():
    """FMRankingLayer backward (forward 2.FM/CustomLayers.py:149-155, tape.gradient
    2.FM/ModelManager.py:176).  V=4, k=2, F=2; embed rows v0=[1,2], v1=[3,-1], v2=[1,1], v3=[0,0];
    X=[[0,1],[2,2]]; upstream dL/dz = [2, -1].
      sample 0: z2 = <v0,v1> = 1;   dz/dv0 = S - v0 = v1,  dz/dv1 = v0
      sample 1: duplicate id 2 in one sample: z2 = <v2,v2> = 2;  dz/dv2 = (S-v2) + (S-v2) = 2 v2
    dL/dembed = g * dz/dv, dL/dw = g per occurrence (row 2 occurs twice -> 2g), dL/dbias = sum g."""
    embed = np.array([[1.0, 2.0], [3.0, -1.0], [1.0, 1.0], [0.0, 0.0]])
    w = np.array([[0.5], [-0.5], [0.25], [0.0]])
    bias = np.array([0.125])
    X = np.array([[0, 1], [2, 2]], dtype=np.int64)
    dz = np.array([2.0, -1.0])
    logit = np.array([[0.125 + 0.0 + 1.0], [0.125 + 0.5 + 2.0]])
    g_embed = np.array([[6.0, -2.0], [2.0, 4.0], [-2.0, -2.0], [0.0, 0.0]])
    g_w = np.array([[2.0], [2.0], [-2.0], [0.0]])
    g_bias = np.array([1.0])
    return embed, w, bias, X, dz, logit, g_embed, g_w, g_bias


def kat7_ffm_pair_gradient():
    """FFM pair term backward on KAT-2's table (T[v,f,d] = 100v + 10f + d), sample X=[0,2,4], dL/dz = 1:
    term = sum_{a<c} <T[x_a,c], T[x_c,a]>  =>  dterm/dT[x_a,c,:] = T[x_c,a,:] (c != a), self slot zero."""
    T, X, _, _ = kat2_field_aware()
    X = X[:1]
    g = np.zeros_like(T)
    g[0, 1] = [200.0, 201.0]      # T[2,0]
    g[0, 2] = [400.0, 401.0]      # T[4,0]
    g[2, 0] = [10.0, 11.0]        # T[0,1]
    g[2, 2] = [410.0, 411.0]      # T[4,1]
    g[4, 0] = [20.0, 21.0]        # T[0,2]
    g[4, 1] = [220.0, 221.0]      # T[2,2]
    return T, X, g


def kat8_keras_adam():
    """tf.keras.optimizers.Adam on IndexedSlices (2.FM/ModelManager.py:103-104,178), lr=0.1, Keras
    defaults beta1=.9, beta2=.999, eps=1e-7.  var = [[1],[2]]; step 1: row 0 twice (0.2 and 0.3:
    dedup sums FIRST, so v gets (0.5)^2, not 0.2^2+0.3^2); step 2: row 1, g=-1; step 3: row 0, g=0.25.
    ``keras_dense`` = Keras 2.8 _resource_apply_sparse (m, v of EVERY row decay each step and var moves
    for every row -- the semantics the reference's shipped checkpoints show, tests/golden/deepfm_ckpt.npz);
    ``rowwise`` = only touched rows.  Returns steps and, per mode, (var, m, v) after each step."""
    steps = [(np.array([0, 0]), np.array([[0.2], [0.3]])), (np.array([1]), np.array([[-1.0]])),
             (np.array([0]), np.array([[0.25]]))]
    dense = [([0.9000006324515321, 2.0], [0.05, 0.0], [0.00025, 0.0]),
             ([0.832995231029295, 2.074413447040717], [0.045, -0.1], [0.00024975, 0.001]),
             ([0.7580859681873267, 2.1319352571855443], [0.0655, -0.09], [0.00031200025, 0.000999])]
    rowwise = [([0.9000006324515321, 2.0], [0.05, 0.0], [0.00025, 0.0]),
               ([0.9000006324515321, 2.074413447040717], [0.05, -0.1], [0.00025, 0.001]),
               ([0.8199769537983428, 2.074413447040717], [0.07, -0.1], [0.00031225, 0.001])]
    return 0.1, steps, {"keras_dense": dense, "rowwise": rowwise}


def kat9_keras_bce():
    """tf.keras.losses.BinaryCrossentropy() on probabilities (2.FM/ModelManager.py:99,175):
    p clipped to [1e-7, 1-1e-7], -(y log(p+1e-7) + (1-y) log(1-p+1e-7)), mean.  p=1,y=1 -> log(1.0)=0;
    p=0,y=1 -> -log(2e-7) = 15.42495; mean of [0.10536041, 0.22314343, 0, 15.42494847]."""
    p = np.array([0.9, 0.2, 1.0, 0.0])
    y = np.array([1.0, 0.0, 1.0, 1.0])
    return p, y, 3.9383630753148284
