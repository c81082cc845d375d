"""The driver scripts' shared logic (nav3d/experiment.py, train/*.py) against the reference's constants and file formats."""
import importlib
import math

import numpy as np
import pytest
import torch

from nav3d import experiment as ex


def test_segments_like_grid_train_180_187():
    assert ex.split_segments(20_000_000) == [2_000_000] * 10
    assert ex.split_segments(25) == [2] * 9 + [7]
    assert ex.split_segments(7) == [7]
    assert sum(ex.split_segments(1_234_567)) == 1_234_567 and len(ex.split_segments(1_234_567)) == 10
    assert ex.eval_every_calls(100_000, 8) == 12_500 and ex.eval_every_calls(100_000, 1 << 20) == 1


def test_checkpoint_names_like_grid_train_148_232():
    arch = dict(pi=[256, 256, 128], vf=[256, 256, 128])
    lstm = dict(lstm_hidden_size=256, n_lstm_layers=1)
    a, l = ex.arch_string(arch), ex.lstm_string(lstm)
    assert a == "pi[256, 256, 128]_vf[256, 256, 128]" and l == "h256l1_shared"
    assert ex.lstm_string(dict(lstm, shared_lstm=False)) == "h256l1_separate"
    name = ex.checkpoint_name(1, a, l, 2_000_000, 10)
    assert name == "rppo_hp1_arch_pi[256, 256, 128]_vf[256, 256, 128]_lstm_h256l1_shared_s2000000_view10.zip"
    assert ex.parse_view_suffix(name) == 10 and ex.parse_view_suffix("model.zip") is None
    assert ex.parse_view_suffix("x_view16.zip_i") is None                  # Train_Further.py:119 anchors on ".zip$"
    assert ex.continued_name(name) == name[:-4] + "_P2.zip_i"            # Train_Further.py:176-177


def test_model_name_parser_handles_both_naming_generations():
    got = ex.parse_model_name("rppo_hp1_arch_pi[256, 256, 128]_vf[256, 256, 128]_lstm_h256l1_shared_s2000000_view10")
    assert got == dict(hp_set=1, arch="pi[256, 256, 128]_vf[256, 256, 128]", lstm="h256l1_shared", trained_steps=2_000_000,
                       view_distance=10, crash_penalty=-2.0)
    old = ex.parse_model_name("rppo_hp3_arch128-128_lstm128x1_s250000_view6_crash-1.5")
    assert old == dict(hp_set=3, arch="128-128", lstm="128x1", trained_steps=250_000, view_distance=6, crash_penalty=-1.5)
    cont = ex.parse_model_name("rppo_hp1_arch_pi[64]_vf[64]_lstm_h64l1_shared_s22000000_view10_P2")
    assert cont["trained_steps"] == 22_000_000 and cont["view_distance"] == 10
    # evaluate_grid.py:143-148
    assert [ex.phase_for_steps(s) for s in (1, 1_000_000, 1_000_001, 21_000_000, 21_000_001)] == \
        ["P1_empty", "P1_empty", "P2_small", "P2_small", "P3_large"]


def test_result_tables_have_the_reference_layout(tmp_path):
    txt, csvp = tmp_path / "r.txt", tmp_path / "r.csv"
    ex.write_results_header(txt)
    ex.append_results_row(txt, "modelA", 12.3456, 3.0, 70.0, 801.25, 1234.5)
    lines = txt.read_text().splitlines()
    assert lines[0] == "Evaluation Results" and lines[1] == "=" * 40 and lines[3] == "-" * 120
    assert lines[2] == ("Model Name" + " " * 30 + " |    Avg Score |    Avg Bumps |    Finished (%) |     Avg Discovered |    Avg Steps")
    assert lines[4] == ("modelA" + " " * 34 + " |        12.35 |         3.00 |           70.0% |             801.25 |      1234.50")
    rows = [dict(Model_Name="modelA", hp_set=1, Architecture="a", LSTM_Size="l", Trained_Steps=5, View_Distance=10,
                 Crash_Penalty=-2.0, Episode_Number=1, Score=1.5, Bumps=2, Finished=True, Discovered_Cells=9, Steps_Taken=11)]
    ex.write_episode_csv(csvp, rows)
    out = csvp.read_text().splitlines()
    assert out[0] == ("Model_Name,hp_set,Architecture,LSTM_Size,Trained_Steps,View_Distance,Crash_Penalty,Episode_Number,"
                      "Score,Bumps,Finished,Discovered_Cells,Steps_Taken")              # evaluate_grid.py:224-239
    assert out[1] == "modelA,1,a,l,5,10,-2.0,1,1.5,2,True,9,11"


def test_script_constants_equal_the_reference():
    gt = importlib.import_module("train.Grid_Train")
    assert gt.BASE_SEED == 42 and gt.NUM_ENVS == 8 and gt.LOCAL_MAP_LENGTHS == [10]            # Grid_Train.py:32-36
    assert gt.SAVE_DIR == "./exp3_architectures" and gt.EVAL_FREQ == 100_000                   # :39, :43
    assert gt.PHASES[0] == ("P1_empty", "./rooms/P1_training", "./rooms/P1_evaluate")          # :51
    assert all(p in gt.STEPS_PHASE for p, _, _ in gt.PHASES)                                   # the reference's KeyError, fixed
    assert gt.architectures == [dict(pi=[256, 256, 128], vf=[256, 256, 128])]                  # :73
    assert gt.lstm_sizes == [dict(lstm_hidden_size=256, n_lstm_layers=1)]                      # :77
    assert gt.ppo_hparam_sets == [dict(learning_rate=3e-4, n_steps=2048, batch_size=64, gamma=0.99, gae_lambda=0.95,
                                       ent_coef=0.01, vf_coef=0.5, clip_range=0.2, n_epochs=10)]   # :84-87
    tf = importlib.import_module("train.Train_Further")
    assert tf.LOAD_DIR == "./exp3_architectures/best_P1_empty_r10_cp-2.0" and tf.SAVE_DIR == "./exp3_architectures/P2"
    assert tf.PHASES == [("P2_small", "./rooms/P2_training", "./rooms/P2_evaluate")] and tf.STEPS_PHASE == {"P2_small": 20_000_000}
    ev = importlib.import_module("train.evaluate_grid")
    assert ev.MODELS_DIR == "./exp3_architectures" and ev.EVAL_EPISODES == 10
    assert ev.RESULTS_TXT_FILE == "exp3_viewDistance.txt" and ev.RESULTS_CSV_FILE == "exp3_viewDistance.csv"
    assert ex.EVAL_ROOMS == {"P1_empty": "./rooms/P1_evaluate", "P2_small": "./rooms/P2_evaluate", "P3_large": "./rooms/P3_evaluate"}


def test_evaluate_checkpoint_rows_on_oracle_env():
    from nav3d.ppo import RecurrentPPO
    from test_ppo_host import tiny_rooms
    from train_refs import OracleBatchedEnv, TorchOps
    env = OracleBatchedEnv(tiny_rooms(), 5, seed=3)
    model = RecurrentPPO(None, policy_kwargs=dict(net_arch=dict(pi=[16], vf=[16]), lstm_hidden_size=16), ops=TorchOps(), seed=2)
    agg, rows = ex.evaluate_checkpoint(model, env, "rppo_hp2_arch_pi[16]_vf[16]_lstm_h16l1_shared_s500_view10", 5)
    assert len(rows) == 5 and [r["Episode_Number"] for r in rows] == [1, 2, 3, 4, 5]
    assert all(r["hp_set"] == 2 and r["Trained_Steps"] == 500 and r["View_Distance"] == 10 for r in rows)
    assert all(1 <= r["Steps_Taken"] <= 36 and r["Discovered_Cells"] >= 1 for r in rows)
    assert math.isclose(agg["avg_steps"], np.mean([r["Steps_Taken"] for r in rows]))
    assert math.isclose(agg["avg_score"], np.mean([r["Score"] for r in rows]), rel_tol=1e-6)
    assert agg["finish_percentage"] == 100.0 * sum(r["Finished"] for r in rows) / 5
