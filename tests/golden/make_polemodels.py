"""Generates, from the reference's packaged rider-behaviour models (src/cyclistsocialforce/data/
balancingriderparams/*.yaml, loaded with the reference's own PoleModel.import_from_yaml):

  cyclistsocialforce_b200/data/pole_models.json   the numbers the stochastic pole sampler needs (conditional
      Gaussian mixture over [speed, pole features] and the pre-processing pipeline: log shift, Yeo-Johnson
      lambdas, standard scaler) -- parameters fitted by the reference's authors, re-serialised, no code;
  tests/golden/golden_polemodel.npz              reference answers for every deterministic stage of
      PoleModel.sample_poles (controlbehavior.py:1337-1469) at a few speeds, and quantiles of 40,000
      reference samples per speed for the distribution-level check.

Run in the build container (needs /root/reference):  python tests/golden/make_polemodels.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

rh.install()
import importlib.resources as res  # noqa: E402

from cyclistsocialforce.controlbehavior import PoleModel  # noqa: E402

FILES = ["BR1_ImRe5GivenV_pole-model-params.yaml", "BR0_ImRe5GivenV_pole-model-params.yaml"]
SPEEDS = [2.0, 3.5, 5.0, 6.5]
QLEVELS = np.linspace(0.01, 0.99, 50)


def main():
    models, gold = {}, {"speeds": np.array(SPEEDS), "qlevels": QLEVELS}
    for fn in FILES:
        pm = PoleModel.import_from_yaml(res.files("cyclistsocialforce.data.balancingriderparams").joinpath(fn))
        lt, pt = pm.pp_pipeline.transformers_
        ig = pm.features.index(pm.feature_cond)
        models[fn] = dict(
            features=list(pm.features), feature_cond=pm.feature_cond, index_given=ig,
            log_features=[int(i) for i in pm.pp_pipeline.log_transform_features_],
            log_a=lt.a_.ravel().tolist(), log_sign=lt.sign_.ravel().tolist(),
            lambdas=pt.lambdas_.tolist(), scaler_mean=pt._scaler.mean_.tolist(), scaler_scale=pt._scaler.scale_.tolist(),
            weights=pm.gmm_.weights_.tolist(), means=pm.gmm_.means_.tolist(),
            covariances=pm.gmm_.get_full_covariancematrix().tolist())
        key = fn.split("_")[0]
        rng = np.random.default_rng(5)
        z = rng.normal(size=(64, len(pm.features) - 1)) * 1.5          # points in the transformed feature space
        idx = [i for i, f in enumerate(pm.features) if f != pm.feature_cond]
        gold[f"{key}_z"] = z
        gold[f"{key}_z_inverse"] = pm.pp_pipeline.inverse_transform(z, sparse_column_indices=idx)
        for v in SPEEDS:
            xg = np.zeros((1, len(pm.features)))
            xg[:, ig] = v
            xt = pm.pp_pipeline.transform(xg, sparse_column_indices=[ig])[:, ig]
            gc = pm.gmm_._get_conditional_gmm(xt)
            tag = f"{key}_v{v}"
            gold[tag + "_xt"] = xt
            gold[tag + "_w"] = gc.weights_
            gold[tag + "_mu"] = gc.means_
            gold[tag + "_cov"] = gc.get_full_covariancematrix() if hasattr(gc, "get_full_covariancematrix") else gc.covariances_
            poles = np.concatenate([pm.sample_poles(n_samples=4000, X_given=v)[0] for _ in range(10)])
            feats = np.c_[poles[:, 0].real, poles[:, 1].real, poles[:, 1].imag, poles[:, 3].real, poles[:, 3].imag]
            assert np.all(poles.real <= 0)
            gold[tag + "_quantiles"] = np.quantile(feats, QLEVELS, axis=0)
            gold[tag + "_mean"] = feats.mean(axis=0)
            gold[tag + "_cov_samples"] = np.cov(feats.T)
    with open(os.path.join(ROOT, "cyclistsocialforce_b200", "data", "pole_models.json"), "w") as f:
        json.dump(models, f, indent=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_polemodel.npz"), **gold)
    print("wrote", len(models), "models,", len(gold), "golden arrays")


if __name__ == "__main__":
    main()
