"""-m gpu: the three driver scripts end to end on the GPU-resident env (tiny budgets): Grid_Train writes its ten
checkpoints with the reference's names, Train_Further resumes one of them (counter, optimiser) and writes the ``_P2.zip_i``
file, evaluate_grid produces the TXT/CSV tables."""
import csv
import os
import shutil
from pathlib import Path

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_grid_train_then_train_further_then_evaluate(tmp_path, monkeypatch, capsys):
    monkeypatch.chdir(ROOT)                       # the scripts use ./rooms/... like the reference
    from train import Grid_Train, Train_Further, evaluate_grid
    save = tmp_path / "exp3_architectures"
    Grid_Train.main(["--native", "--num-envs", "64", "--steps", str(10 * 128 * 64), "--save-dir", str(save),
                     "--eval-freq", str(128 * 64)])   # one evaluation per segment (the callback is re-created per segment)
    stem = "rppo_hp1_arch_pi[256, 256, 128]_vf[256, 256, 128]_lstm_h256l1_shared"
    names = sorted(p.name for p in save.glob("*.zip"))
    assert names == sorted(f"{stem}_s{(i + 1) * 128 * 64}_view10.zip" for i in range(10))
    assert (save / "best_exp3_P1" / "best_model.zip").exists() and (save / "best_exp3_P1" / "evaluations.npz").exists()

    load = tmp_path / "to_continue"
    load.mkdir()
    last = f"{stem}_s{10 * 128 * 64}_view10.zip"
    shutil.copy(save / last, load / last)
    p2 = tmp_path / "P2"
    assert Train_Further.main(["--load-dir", str(load), "--save-dir", str(p2), "--steps", str(10 * 128 * 32),
                               "--num-envs", "32"]) == 0
    cont = p2 / (last[:-4] + "_P2.zip_i")
    assert cont.exists()
    from nav3d.ppo import RecurrentPPO
    m = RecurrentPPO.load(cont, env=None, device="cuda")
    assert m.num_timesteps == 10 * 128 * 64 + 10 * 128 * 32          # the step counter carried over the reload

    txt, csvp = tmp_path / "res.txt", tmp_path / "res.csv"
    two = tmp_path / "two_models"
    two.mkdir()
    for i in (1, 10):
        n = f"{stem}_s{i * 128 * 64}_view10.zip"
        shutil.copy(save / n, two / n)
    evaluate_grid.main(["--models-dir", str(two), "--episodes", "4", "--txt", str(txt), "--csv", str(csvp)])
    lines = txt.read_text().splitlines()
    assert lines[0] == "Evaluation Results" and len(lines) == 4 + 2
    rows = list(csv.DictReader(open(csvp)))
    assert len(rows) == 8 and {r["Model_Name"] for r in rows} == {f"{stem}_s{i * 128 * 64}_view10" for i in (1, 10)}
    for r in rows:
        assert r["View_Distance"] == "10" and r["Crash_Penalty"] == "-2.0" and r["Finished"] in ("True", "False")
        assert 1 <= int(r["Steps_Taken"]) <= 3240 and 1 <= int(r["Discovered_Cells"]) <= 3240      # P1_evaluate: 3240 free cells
