"""Weights & Biases helpers (reference ``training/wandb_utils.py``); a no-op when wandb is absent."""
import os


def _wandb():
    try:
        import wandb

        return wandb
    except Exception:
        return None


def on_save_checkpoint_success(checkpoint_path: str, tracker_filename: str, save_dir: str, iteration: int):
    w = _wandb()
    if w is None or w.run is None:
        return
    art = w.Artifact(f"{os.path.basename(save_dir)}-iter{iteration}", type="model",
                     metadata={"iteration": iteration, "path": checkpoint_path})
    art.add_reference(f"file://{checkpoint_path}", checksum=False)
    art.add_file(tracker_filename)
    w.run.log_artifact(art)


def on_load_checkpoint_success(checkpoint_path: str, load_dir: str):
    w = _wandb()
    if w is None or w.run is None:
        return
    try:
        w.run.use_artifact(f"{os.path.basename(load_dir)}:latest")
    except Exception:
        pass
