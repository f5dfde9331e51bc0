"""Inference-side subset of reference training.py: the checkpoint locator used by
prediction.model_from_checkpoint_path (training.py:41-71) and the model registry the CLI reads
(training.py imports LANDMARKS_MODELS; scripts/cli.py:31,71).  Training itself is out of scope."""
import glob

from .networks.basic_models import LANDMARKS_MODELS  # noqa: F401


def find_latest_checkpoint(checkpoints_path, fail_safe=True):
    """Same contract as reference training.py:41-71: newest "<path>.<epoch>" by numeric suffix (a trailing
    ".index" / ".npz" / ".safetensors" is stripped), None (or ValueError when fail_safe is False) if there is none."""

    def epoch_of(path):
        return path.replace(checkpoints_path, "").strip(".")

    files = glob.glob(checkpoints_path + ".*")
    if len(files) == 0:
        files = glob.glob(checkpoints_path + "*.*")
    files = [f.replace(".index", "") for f in files]
    for ext in (".npz", ".safetensors"):       # the two weight containers Model.load_weights / save_weights handle
        files = [f[:-len(ext)] if f.endswith(ext) else f for f in files]
    files = sorted(set(f for f in files if epoch_of(f).isdigit()))
    if not files:
        if not fail_safe:
            raise ValueError("Checkpoint path {0} invalid".format(checkpoints_path))
        return None
    return max(files, key=lambda f: int(epoch_of(f)))
