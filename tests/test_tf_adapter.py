"""The Keras adapter is import-guarded: without TensorFlow the module imports, names the missing dependency when used,
and the torch-side binding it mirrors is importable.  (With TensorFlow the functional-API / GradientTape path of
INTEGRATION.md section 2 applies; TensorFlow cannot be installed in the build image.)"""
import importlib

import pytest

import etr_b200  # noqa: F401


def test_adapter_imports_without_tensorflow_and_says_what_is_missing():
    ad = importlib.import_module("etr_b200.tf_adapter")
    if ad.tf is not None:
        pytest.skip("TensorFlow is installed here: the guard does not apply")
    with pytest.raises(ImportError, match="TensorFlow"):
        ad.require_tf()
    with pytest.raises(ImportError, match="TensorFlow"):
        ad.DeepFMRankingLayer
    with pytest.raises(ImportError, match="TensorFlow"):
        ad.Adam()
    with pytest.raises(AttributeError):
        ad.NoSuchLayer
    assert set(ad._MODEL_LAYERS) >= {"FMRankingLayer", "DeepFMRankingLayer", "FFMLayer", "FwFMLayer", "PNNRankingLayer",
                                     "DeepCrossNetworkLayer"}


def test_torch_binding_is_importable_on_cpu():
    ag = importlib.import_module("etr_b200.autograd")
    assert hasattr(ag, "EtrModule") and hasattr(ag, "EtrAdam")
