"""Reader / writer for the reference's demonstration logs.

``examples/ball_env_imitate.py:253-256`` pickles two lists per trial: the legacy 29-float observations
(``prep_state2``, see ``BallVecEnv.block_counts``) and the human actions as ``[+-10, +-10]`` int64 arrays;
``examples/train_supervise.py:32-59`` reads them back and turns an action into the index of
``move_list`` (``action / 10`` as a tuple; an action that is not in the list gets the last index, 8, because the
search loop falls through).  The shipped files (``examples/State_info_trail_no2`` / ``Trial_no_2`` and
``examples/pathlogs/*``) were written by Python 2, protocol 0: they load with ``encoding='latin1'``.

Host-side utility: nothing here touches the GPU path except the optional ``device=`` of the returned tensors.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

# the order of examples/train_supervise.py:45 (the A2C scripts use the same list, examples/ball_cnn_ac3.py:530)
MOVE_LIST = [(1, 1), (1, -1), (1, 0), (0, 1), (0, -1), (0, 0), (-1, 1), (-1, 0), (-1, -1)]
_INDEX = {m: j for j, m in enumerate(MOVE_LIST)}


def _load(path):
    with open(path, "rb") as f:
        try:
            return pickle.load(f, encoding="latin1")      # Python 2 pickles of numpy arrays
        except TypeError:                                 # pragma: no cover  (a file object without encoding support)
            f.seek(0)
            return pickle.load(f)


def action_labels(actions) -> np.ndarray:
    """``[M, 2]`` logged actions (multiples of 10) -> ``[M]`` indices into MOVE_LIST, as prepare_train_Y does."""
    a = np.asarray(actions)
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError("expected [M, 2] actions, got %r" % (a.shape,))
    unit = np.floor_divide(a, 10) if np.issubdtype(a.dtype, np.integer) else a / 10     # py2 '/' on int64 arrays
    return np.array([_INDEX.get((x, y), len(MOVE_LIST) - 1) for x, y in unit.tolist()], dtype=np.int64)


def load_path_log(state_file, action_file=None, device=None):
    """-> ``(states float32 [M, 29], labels int64 [M] or None)`` as torch tensors (on ``device`` if given)."""
    states = np.asarray(_load(state_file), dtype=np.float64)
    if states.ndim != 2 or states.shape[1] != 29:
        raise ValueError("%s: expected a list of 29-float vectors, got shape %r" % (state_file, states.shape))
    x = torch.from_numpy(states.astype(np.float32))
    y = None
    if action_file is not None:
        labels = action_labels(np.asarray(_load(action_file)))
        if len(labels) != len(states):
            raise ValueError("%d states but %d actions" % (len(states), len(labels)))
        y = torch.from_numpy(labels)
    if device is not None:
        x = x.to(device)
        y = y.to(device) if y is not None else None
    return x, y


def save_path_log(states, action_indices, state_file, action_file):
    """Write the two files in the reference's layout (lists of numpy arrays; actions as 10 * move_list[index])."""
    s = np.asarray(torch.as_tensor(states).cpu(), dtype=np.float64)
    idx = np.asarray(torch.as_tensor(action_indices).cpu(), dtype=np.int64)
    if s.ndim != 2 or s.shape[1] != 29 or idx.shape != (len(s),):
        raise ValueError("expected states [M, 29] and action indices [M]")
    with open(state_file, "wb") as f:
        pickle.dump([row.copy() for row in s], f, protocol=2)
    with open(action_file, "wb") as f:
        pickle.dump([10 * np.array(MOVE_LIST[j], dtype=np.int64) for j in idx.tolist()], f, protocol=2)
