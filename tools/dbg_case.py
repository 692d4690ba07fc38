import sys, os, warnings
sys.path.insert(0, os.getcwd())
import numpy as np
from tests.helpers import golden, run_b200_case
name = sys.argv[1]
g = golden(name)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    smc, steps, mll, rng = run_b200_case(name, g)
m = min(len(smc.phi_sequence), len(g["phi"]))
print("len", len(smc.phi_sequence), len(g["phi"]))
print("phi rel", np.abs(smc.phi_sequence[:m] - g["phi"][:m]) / np.maximum(g["phi"][:m], 1e-300))
print("mll abs", np.abs(mll[:m] - g["mll"][:m]))
print("mut", [s.attrs["mutation_ratio"] for s in steps][:m], g["mutation_ratio"][:m])
print("p1", np.abs(steps[1].params - g["params_1"]).max(), np.abs(steps[1].log_likes - g["log_likes_1"]).max())
