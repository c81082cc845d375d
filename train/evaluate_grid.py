"""Evaluate every saved model of a directory on the evaluation rooms of its training phase.

Drop-in for the reference's ``train/evaluate_grid.py``: for each ``*.zip`` under ``MODELS_DIR`` play ``EVAL_EPISODES``
deterministic episodes, append one aggregated line to ``RESULTS_TXT_FILE`` and write one CSV row per episode to
``RESULTS_CSV_FILE`` (the reference's layouts).  The episodes of one model run side by side in one batched env instead of
one after the other.  The eval rooms follow the reference's rule on the trained-step count parsed from the file name.

    python -m train.evaluate_grid [--models-dir D --episodes N --txt F --csv F]"""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import _nav3d_path  # noqa: E402,F401

from nav3d.experiment import (EVAL_ROOMS, append_results_row, evaluate_checkpoint, make_vec_env, parse_model_name,  # noqa: E402
                              phase_for_steps, write_episode_csv, write_results_header)
from nav3d.ppo import RecurrentPPO  # noqa: E402

MODELS_DIR = "./exp3_architectures"
EVAL_EPISODES = 10
RESULTS_TXT_FILE = "exp3_viewDistance.txt"
RESULTS_CSV_FILE = "exp3_viewDistance.csv"
ROOMS = "./rooms/P1_evaluate"
RENDER_EVAL_EPISODES = 0


def evaluate_models(models_dir: str, num_episodes: int, results_txt_filename: str, results_csv_filename: str,
                    render_episodes: int = 0, device: int = 0):
    print(f"Starting evaluation of models in '{models_dir}'...")
    if not os.path.isdir(models_dir):
        print(f"Error: Directory not found at '{models_dir}'")
        return
    model_files = [f for f in os.listdir(models_dir) if f.endswith(".zip")]
    if not model_files:
        print(f"No .zip models found in '{models_dir}'")
        return
    write_results_header(results_txt_filename)
    all_episode_results = []
    for model_file in model_files:
        model_name = os.path.splitext(model_file)[0]
        print(f"\n--- Evaluating model: {model_name} ---")
        info = parse_model_name(model_name)
        if info["view_distance"] is None:
            print(f"Could not parse the view distance from {model_name}; skipping.")
            continue
        phase = phase_for_steps(info["trained_steps"])
        room_path = EVAL_ROOMS[phase]
        print(f"  Hyperparameter set: {info['hp_set']}\n  Trained steps: {info['trained_steps']}\n"
              f"  Using phase: {phase}\n  Room path: {room_path}")
        env = make_vec_env(room_path, info["view_distance"], num_episodes, 0, device=device,
                           crash_penalty=info["crash_penalty"])
        try:
            model = RecurrentPPO.load(os.path.join(models_dir, model_file), env=None, device=env.device)
        except Exception as e:  # noqa: BLE001
            print(f"Could not load model {model_name}. Error: {e}")
            env.close()
            continue
        agg, rows = evaluate_checkpoint(model, env, model_name, num_episodes)
        env.close()
        all_episode_results.extend(rows)
        for r in rows:
            print(f"  Episode {r['Episode_Number']}/{num_episodes} finished. Score: {r['Score']:.2f}")
        append_results_row(results_txt_filename, model_name, agg["avg_score"], agg["avg_bumps"], agg["finish_percentage"],
                           agg["avg_discovered"], agg["avg_steps"])
        print(f"--- Aggregated Results for {model_name} ---")
        print(f"  Average Score: {agg['avg_score']:.2f}\n  Average Bumps: {agg['avg_bumps']:.2f}\n"
              f"  Times Finished: {agg['finish_percentage']:.1f}%\n  Average Cells Discovered: {agg['avg_discovered']:.2f}\n"
              f"  Average Steps Taken: {agg['avg_steps']:.2f}")
    if all_episode_results:
        write_episode_csv(results_csv_filename, all_episode_results)
        print(f"\nIndividual trial results saved to '{results_csv_filename}'.")
    else:
        print("\nNo models were evaluated, so no individual trial results were saved.")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--models-dir", default=MODELS_DIR)
    ap.add_argument("--episodes", type=int, default=EVAL_EPISODES)
    ap.add_argument("--txt", default=RESULTS_TXT_FILE)
    ap.add_argument("--csv", default=RESULTS_CSV_FILE)
    args = ap.parse_args(argv)
    if not os.path.exists(args.models_dir):
        os.makedirs(args.models_dir)
        print(f"Created directory '{args.models_dir}'. Please place your trained models in this folder.")
        return
    evaluate_models(args.models_dir, args.episodes, args.txt, args.csv, RENDER_EVAL_EPISODES)
    print(f"\nEvaluation complete. Aggregated results saved to '{args.txt}'.")


if __name__ == "__main__":
    main()
