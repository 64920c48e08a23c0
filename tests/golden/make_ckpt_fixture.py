#!/usr/bin/env python
"""Generates tests/golden/deepfm_ckpt.npz from the TensorFlow checkpoints the reference ships
(/root/reference/2.FM/ranking_model/checkpoint/ckpt-{0,1,2}: a DeepFM trained by the reference's own
train loop, tf.keras.optimizers.Adam, TF 2.8.0) using the TF-free bundle reader
explicit-tf2-recommendation_b200/tf_checkpoint.py.  These are OUTPUTS OF THE REFERENCE ITSELF, so they
pin two things the oracle otherwise restates from memory:

  * the sparse Adam semantics of Keras 2.8 (a16): statistics of the slot variables ``m``/``v`` between
    consecutive epochs (ckpt-1 -> ckpt-2, 5709 steps apart) that distinguish "every row decays every
    step" (Keras ``_resource_apply_sparse``) from a lazy / row-wise Adam;
  * variable names, shapes and order of the DeepFM layer, the optimizer hyper-parameters.

plus a slice of realistic trained weights + optimizer state (rows 0..767 of embed / w and their slots,
the MLP variables) used as a weight fixture by the GPU parity tests.

    python tests/golden/make_ckpt_fixture.py        # needs /root/reference (not present on the GPU box)
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
spec = importlib.util.spec_from_file_location("tfc", os.path.join(ROOT, "explicit-tf2-recommendation_b200",
                                                                  "tf_checkpoint.py"))
tfc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tfc)

CK = "/root/reference/2.FM/ranking_model/checkpoint/ckpt-%d"
ROWS = 768


def main():
    W = [tfc.deepfm_weights(CK % n) for n in (0, 1, 2)]
    # every tensor's checksum is verified once
    for n in (0, 1, 2):
        tfc.load_checkpoint(CK % n, verify_crc=True)
    out = {}
    out["iters"] = np.array([int(w["optimizer/iter"]) for w in W], dtype=np.int64)
    out["beta_1"] = np.float32(W[2]["optimizer/beta_1"])
    out["beta_2"] = np.float32(W[2]["optimizer/beta_2"])
    out["learning_rate"] = np.float32(W[2]["optimizer/learning_rate"])
    out["names"] = np.array(sorted(W[2].keys()))
    out["shapes"] = np.array([str(tuple(W[2][k].shape)) for k in sorted(W[2].keys())])
    for var in ("embed/embeddings", "w/embeddings"):
        tag = var.split("/")[0]
        v1, v2 = W[1][var + "/v"].astype(np.float64), W[2][var + "/v"].astype(np.float64)
        m2 = W[2][var + "/m"].astype(np.float64)
        out[f"{tag}_v_ratio_min"] = np.float64((v2 / v1).min())
        out[f"{tag}_v_ratio_quantiles"] = np.quantile(v2 / v1, [0.0, 0.01, 0.1, 0.5])
        r = np.abs(m2) / np.sqrt(v2)
        out[f"{tag}_m_over_sqrtv_quantiles"] = np.quantile(r, [0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99])
        out[f"{tag}_all_rows_have_state"] = np.bool_((v2 > 0).all())
    w = W[2]
    for key in ("embed/embeddings", "w/embeddings"):
        for slot in ("", "/m", "/v"):
            out["slice/" + key + slot] = w[key + slot][:ROWS]
    for key in sorted(w):
        if key.startswith("MLP_layer") or key.startswith("bias"):
            out["var/" + key] = w[key]
    path = os.path.join(HERE, "deepfm_ckpt.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    print("steps/epoch", out["iters"][2] - out["iters"][1], "beta2^N", 0.999 ** int(out["iters"][2] - out["iters"][1]),
          "min v2/v1 embed", out["embed_v_ratio_min"], "w", out["w_v_ratio_min"])
    print("|m|/sqrt(v) quantiles embed", out["embed_m_over_sqrtv_quantiles"])


if __name__ == "__main__":
    main()
