import os, sys
sys.path.insert(0, ".")
import numpy as np
from fedem_solvers_b200 import StressRecovery
from fedem_solvers_b200.model import hex20_block, reduced_history
part = hex20_block(3, 2, 2, ngen=5, seed=6, shuffle_eq=True)
Q = reduced_history(part.sam.ndim, 150, seed=7)
out = {}
for tag, env in (("on", None), ("off", "0"), ("dense", None)):
    if env is not None: os.environ["FSR_HEX20_STEPLANE"] = env
    else: os.environ.pop("FSR_HEX20_STEPLANE", None)
    rec = StressRecovery(part, step_tile=64)
    out[tag] = rec.recover(Q)
    rec.close()
d = np.abs(out["on"] - out["off"]) / np.abs(out["off"]).max()
print("on vs off: max rel", d.max(), "tile0", d[:64].max(), "tile2", d[128:].max(), "equal tile0", np.array_equal(out["on"][:64], out["off"][:64]))
