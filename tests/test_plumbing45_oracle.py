"""BASELINE.json configs 4 (reference-attention UNet) and 5 (video UNet) through the reference's own control path
(options -> create_model -> two optimize_parameters() with AdamW + weight decay + EMA, oracle/gen_golden_plumbing45.py):
the oracle's train step reproduces the losses and the parameters / EMA afterwards."""
import os

import pytest
import torch

from oracle import palette_oracle as O
from oracle.vid_oracle import init_params_from_shapes


@pytest.mark.parametrize("which", ["ref", "vid"])
def test_oracle_train_steps_match_reference_plumbing(golden_dir, which):
    from oracle.gen_golden_plumbing45 import batch, draws, oracle_cfg, oracle_forward
    gold = torch.load(os.path.join(golden_dir, "vid_plumbing.pt" if which == "vid" else "refattn_plumbing.pt"))
    cfg = oracle_cfg(which)
    cfg.n_timestep_train, cfg.n_timestep_test = gold["n_timestep_train"], gold["n_timestep_test"]
    state = O.TrainState(params=init_params_from_shapes(gold["shapes"], gold["wseed"]))
    for step in range(2):
        data = batch(which, gold["data_seeds"][step])
        t, u, noise = draws(which, cfg, gold["rng_seeds"][step])
        loss, _, _ = O.train_step(state, cfg, O.OptimCfg(**gold["optim"]), data["B"], data["A"], data["B_label_mask"],
                                  noise, t, u, lambda_G=gold["lambda_G"],
                                  forward=oracle_forward(which, cfg, data, noise, t, u))
        assert abs(float(loss) - gold["losses"][step]) < 1e-5 * gold["losses"][step], (step, float(loss))
    for k, (s, n) in gold["param_stats"].items():
        assert abs(float(state.params[k].double().norm()) - n) <= 1e-4 * n + 1e-9, k
    for k, (s, n) in gold["ema_stats"].items():
        assert abs(float(state.ema[k].double().norm()) - n) <= 1e-4 * n + 1e-9, k
