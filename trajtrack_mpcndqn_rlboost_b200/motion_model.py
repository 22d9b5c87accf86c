"""numpy unicycle model, same call signature as the reference's
``pkg_motion_model.motion_model.unicycle_model``
(/root/reference/src/pkg_motion_model/motion_model.py:153-176)."""
import numpy as np


def unicycle_model(state: np.ndarray, action: np.ndarray, ts: float, rk4: bool = True) -> np.ndarray:
    def d_state_f(st, ac):
        return ts * np.array([ac[0] * np.cos(st[2]), ac[0] * np.sin(st[2]), ac[1]])
    if rk4:
        k1 = d_state_f(state, action)
        k2 = d_state_f(state + 0.5 * k1, action)
        k3 = d_state_f(state + 0.5 * k2, action)
        k4 = d_state_f(state + k3, action)
        d_state = (1 / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    else:
        d_state = d_state_f(state, action)
    return state + d_state
