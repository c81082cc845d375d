"""What the reference's three driver scripts share, as functions: phase/segment schedule, checkpoint
naming and parsing, vec-env construction, the evaluation tables.

File formats are the reference's and are kept byte-compatible where a reader could depend on them:
  * checkpoints ``rppo_hp{i}_arch_{arch}_lstm_{lstm}_s{cumulative_steps}_view{ray_len}.zip`` (train/Grid_Train.py:232),
    continued ones ``{base}_P2.zip_i`` (train/Train_Further.py:176-178), resume regex ``_view(\\d+)\\.zip$`` (:119);
  * ten training segments per phase, remainder added to the last (train/Grid_Train.py:180-187);
  * evaluation ``.txt`` header/row layout (train/evaluate_grid.py:102-107, :259) and ``.csv`` columns (:224-239);
  * the eval-room choice by trained steps (:143-149).
Shipped breakages that are fixed without changing any format: ``STEPS_PHASE`` lacks the key of the only
active phase (Grid_Train.py:53-61, KeyError at :176); the name parser of evaluate_grid.py (:118-142) does not parse the
names Grid_Train.py writes (``lstm`` token, missing ``crash`` token); ``envs.Venv`` does not exist in the reference."""
from __future__ import annotations

import csv
import os
import re
from pathlib import Path
from typing import Dict, List, Optional, Sequence

BASE_SEED = 42                                   # train/Grid_Train.py:32
EVAL_ROOMS = {"P1_empty": "./rooms/P1_evaluate", "P2_small": "./rooms/P2_evaluate", "P3_large": "./rooms/P3_evaluate"}

CSV_COLUMNS = ["Model_Name", "hp_set", "Architecture", "LSTM_Size", "Trained_Steps", "View_Distance", "Crash_Penalty",
               "Episode_Number", "Score", "Bumps", "Finished", "Discovered_Cells", "Steps_Taken"]


# ---- schedule ------------------------------------------------------------------------------------------------------
def split_segments(steps_this_phase: int, n: int = 10) -> List[int]:
    """train/Grid_Train.py:180-187: n equal segments, the remainder goes to the last; fewer than n steps = one segment."""
    seg = steps_this_phase // n
    if seg == 0:
        return [steps_this_phase]
    segments = [seg] * n
    segments[-1] += steps_this_phase - seg * n
    return segments


def eval_every_calls(eval_freq_steps: int, num_envs: int) -> int:
    """train/Grid_Train.py:222: EvalCallback counts vec-env steps, so the step budget is divided by the env count."""
    return max(eval_freq_steps // num_envs, 1)


# ---- names ---------------------------------------------------------------------------------------------------------
def arch_string(arch: Dict[str, Sequence[int]]) -> str:
    return f"pi{list(arch['pi'])}_vf{list(arch['vf'])}"                       # train/Grid_Train.py:148


def lstm_string(lstm_kwargs: dict) -> str:
    shared = "shared" if lstm_kwargs.get("shared_lstm", True) else "separate"   # :152 (the label, not the behaviour)
    return f"h{lstm_kwargs['lstm_hidden_size']}l{lstm_kwargs['n_lstm_layers']}_{shared}"


def checkpoint_name(hp_index: int, arch_str: str, lstm_str: str, cumulative_steps: int, ray_len: int) -> str:
    return f"rppo_hp{hp_index}_arch_{arch_str}_lstm_{lstm_str}_s{cumulative_steps}_view{ray_len}.zip"


def parse_view_suffix(filename: str) -> Optional[int]:
    m = re.search(r"_view(\d+)\.zip$", filename)                              # train/Train_Further.py:119
    return int(m.group(1)) if m else None


def continued_name(model_filename: str) -> str:
    base, ext = os.path.splitext(model_filename)                              # train/Train_Further.py:176-177
    return f"{base}_P2{ext}_i"


def parse_model_name(model_name: str) -> dict:
    """Fields evaluate_grid.py:118-142 pulls out of a checkpoint's stem.  Accepts the names Grid_Train.py writes and the
    older ``..._arch128-128_lstm128x1_s250000_view6_crash-2.0`` style the reference's parser was written for."""
    out = dict(hp_set=1, arch="", lstm="", trained_steps=0, view_distance=None, crash_penalty=-2.0)
    m = re.search(r"hp(\d+)", model_name)
    if m:
        out["hp_set"] = int(m.group(1))
    m = re.search(r"arch_?(.*?)_lstm", model_name)
    if m:
        out["arch"] = m.group(1).strip("_")
    m = re.search(r"lstm_?(.*?)_s\d+(?:_|$)", model_name)
    if m:
        out["lstm"] = m.group(1).strip("_")
    m = re.search(r"_s(\d+)(?:_|$)", model_name)
    if m:
        out["trained_steps"] = int(m.group(1))
    m = re.search(r"view(\d+)", model_name) or re.search(r"_r(\d+)_", model_name)
    if m:
        out["view_distance"] = int(m.group(1))
    m = re.search(r"crash(-?\d+(?:\.\d+)?)", model_name)
    if m:
        out["crash_penalty"] = float(m.group(1))
    return out


def phase_for_steps(trained_steps: int) -> str:
    if trained_steps <= 1_000_000:                                            # train/evaluate_grid.py:143-148
        return "P1_empty"
    if trained_steps <= 21_000_000:
        return "P2_small"
    return "P3_large"


# ---- environments --------------------------------------------------------------------------------------------------
def dist_setup():
    """(rank, world, local_rank); initialises NCCL when launched under torchrun with more than one process."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return rank, world, local_rank


def make_vec_env(room_path, ray_len: int, num_envs: int, seed_offset: int = 0, *, rank: int = 0, device: int = 0,
                 crash_penalty: float = -2.0, auto_reset: bool = True):
    """The GPU-resident replacement for ``SubprocVecEnv([make_env_fn(room_path, ray_len, i + seed_offset) ...])``
    (train/Grid_Train.py:118-126, :170-173, :191-192).  Every env draws its rooms and starts from its own Philox stream
    keyed by (BASE_SEED + seed_offset, global env id); ranks own disjoint global ids.  Like the reference's ``GridEnv``
    (:101-102) the training scripts never forward a crash penalty, so -2.0 applies."""
    from .vec_env import BatchedCubicEnv
    return BatchedCubicEnv(room_path, num_envs=num_envs, local_map_length=ray_len, crash_penalty=crash_penalty,
                           seed=BASE_SEED + seed_offset, device=device, env_id0=rank * num_envs, auto_reset=auto_reset)


# ---- evaluation tables ---------------------------------------------------------------------------------------------
def write_results_header(path) -> None:
    with open(path, "w") as f:                                                # train/evaluate_grid.py:102-107
        f.write("Evaluation Results\n")
        f.write("=" * 40 + "\n")
        f.write(f"{'Model Name':<40} | {'Avg Score':>12} | {'Avg Bumps':>12} | {'Finished (%)':>15} | {'Avg Discovered':>18} | {'Avg Steps':>12}\n")
        f.write("-" * 120 + "\n")


def append_results_row(path, model_name: str, avg_score: float, avg_bumps: float, finish_percentage: float,
                       avg_discovered: float, avg_steps: float) -> None:
    with open(path, "a") as f:                                                # train/evaluate_grid.py:259
        f.write(f"{model_name:<40} | {avg_score:>12.2f} | {avg_bumps:>12.2f} | {finish_percentage:>14.1f}% | {avg_discovered:>18.2f} | {avg_steps:>12.2f}\n")


def write_episode_csv(path, rows: List[dict]) -> None:
    with open(path, "w", newline="") as f:                                    # pandas.DataFrame(rows).to_csv(index=False), :276-278
        w = csv.DictWriter(f, fieldnames=CSV_COLUMNS, lineterminator="\n")
        w.writeheader()
        for r in rows:
            w.writerow({k: r[k] for k in CSV_COLUMNS})


def evaluate_checkpoint(model, env, model_name: str, num_episodes: int) -> (dict, List[dict]):
    """The per-model loop of train/evaluate_grid.py:176-257 on a batched env: ``num_episodes`` deterministic episodes, the
    per-episode rows of the CSV and the aggregates of the TXT line."""
    from .evaluation import evaluate_policy
    info = parse_model_name(model_name)
    st = evaluate_policy(model, env, n_eval_episodes=num_episodes, deterministic=True, return_episode_stats=True)
    rows = []
    for i in range(len(st["r"])):
        rows.append({"Model_Name": model_name, "hp_set": info["hp_set"], "Architecture": info["arch"],
                     "LSTM_Size": info["lstm"], "Trained_Steps": info["trained_steps"],
                     "View_Distance": info["view_distance"], "Crash_Penalty": info["crash_penalty"],
                     "Episode_Number": i + 1, "Score": float(st["r"][i]), "Bumps": int(st["bumps"][i]),
                     "Finished": bool(st["terminated"][i]), "Discovered_Cells": int(st["visited"][i]),
                     "Steps_Taken": int(st["l"][i])})
    n = max(1, len(rows))
    agg = dict(avg_score=float(st["r"].mean()), avg_bumps=float(st["bumps"].mean()),
               finish_percentage=100.0 * float(st["terminated"].sum()) / n, avg_discovered=float(st["visited"].mean()),
               avg_steps=float(st["l"].sum()) / n)
    return agg, rows
