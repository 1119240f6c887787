"""Checkpoint / resume of an attack run: the perturbation, the Adam moments and the step counter.

What the reference keeps between runs, and where:
  * i3d_adversarial_main_universal.py:310-348 — `tf.estimator.RunConfig(save_checkpoints_steps=100, keep_checkpoint_max=5)`;
    a run started in a `model_dir` that already holds a checkpoint continues from `tf.train.latest_checkpoint` (global step,
    `eps_rgb` and the optimizer slots), otherwise it warm-starts the network from MODEL.CKPT_PATH_WITH_ZERO_PERT;
  * i3d_adversarial_main_single_class_gen.py:149,192-197,214,373 — `saver.save(sess, ckpt_dst + 'model_step_{:05d}')` at the start
    and after every pass over the training records; a restart restores the latest file and reads the step from its name;
  * r2plus1d_main_universal_attack.py:197-216 — epoch files `{model_name}_{epoch:03d}.npy`; INIT_PERT_FROM_LAST_CKPT re-reads
    `valid/perturbation` of the newest file, CONTINUE_TRAIN continues the epoch numbering (its optimizer restarts).

The network itself is frozen and never part of a checkpoint here.  Files are `model_step_{step:05d}.npz`, written atomically
by rank 0 (every rank of a sharded run holds the same replicated state and restores the same file); a resumed run continues
bit-identically because everything the update reads (delta, m, v, Adam step) is restored exactly."""
import glob
import os
import re

import numpy as np
import torch

from . import dist as fdist

_PAT = re.compile(r"model_step_(\d+)\.npz$")


def _files(prefix):
    out = []
    for f in glob.glob(prefix + "model_step_*.npz"):
        m = _PAT.search(f)
        if m:
            out.append((int(m.group(1)), f))
    return sorted(out)


def latest_checkpoint(prefix):
    """(step, path) of the newest `<prefix>model_step_XXXXX.npz`, or None.  `prefix` is used as the reference uses `ckpt_dst`:
    a directory with a trailing separator, or a directory plus a file-name stem."""
    files = _files(prefix)
    return files[-1] if files else None


def save_checkpoint(prefix, atk, step, keep_max=5, extra=None):
    """Write `<prefix>model_step_{step:05d}.npz` (rank 0 of a sharded run only; a no-op elsewhere) and prune to the newest
    `keep_max` files (tf.estimator's keep_checkpoint_max)."""
    if not fdist.is_writer(getattr(atk, "world", 1)):
        return None
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    sd = atk.state_dict()
    payload = {"delta": sd["delta"].numpy(), "m": sd["m"].numpy(), "v": sd["v"].numpy(),
               "adam_step": np.int64(sd["step"]), "step": np.int64(step)}
    for k, v in (extra or {}).items():
        payload["extra_" + k] = np.asarray(v)
    path = "{}model_step_{:05d}.npz".format(prefix, int(step))
    tmp = path + ".tmp.npz"
    np.savez(tmp, **payload)
    os.replace(tmp, path)
    if keep_max:
        for _, old in _files(prefix)[:-int(keep_max)]:
            try:
                os.remove(old)
            except OSError:
                pass
    return path


def restore_checkpoint(path, atk):
    """Load delta / Adam moments / Adam step into the attack object; returns (driver step, {extra fields})."""
    with np.load(path) as z:
        atk.load_state_dict({"delta": torch.from_numpy(z["delta"]), "m": torch.from_numpy(z["m"]),
                             "v": torch.from_numpy(z["v"]), "step": int(z["adam_step"])})
        extra = {k[6:]: z[k] for k in z.files if k.startswith("extra_")}
        return int(z["step"]), extra


def resume(prefix, atk):
    """Restore the newest checkpoint under `prefix` if there is one; returns the step to continue from (0 for a new run)."""
    if not prefix:
        return 0
    last = latest_checkpoint(prefix)
    if last is None:
        return 0
    step, _ = restore_checkpoint(last[1], atk)
    return step
