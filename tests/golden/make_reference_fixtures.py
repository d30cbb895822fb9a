"""Generates tests/golden/reference_fd_coefficients.json by importing the ONE reference module that
runs without TensorFlow (poisson_CNN/dataset/utils/get_fd_coefficients.py: numpy + scipy only).
Run in the build container (where /root/reference exists):  python tests/golden/make_reference_fixtures.py
"""
import importlib.util
import json
import os

REF = "/root/reference/poisson_CNN/dataset/utils/get_fd_coefficients.py"
spec = importlib.util.spec_from_file_location("ref_get_fd_coefficients", REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

cases = []
for positions, order in [([-1, 0, 1], 2), ([-2, -1, 0, 1, 2], 2), ([-3, -2, -1, 0, 1, 2, 3], 2),
                         ([-1, 0, 1], 1), ([-2, -1, 0, 1, 2], 1), ([-2, -1, 0, 1, 2], 4), ([-3, -2, -1, 0, 1], 2)]:
    cases.append({"positions": positions, "order": order,
                  "coefficients": [float(v) for v in mod.get_fd_coefficients(positions, order)]})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fd_coefficients.json")
json.dump({"source": "poisson_CNN/dataset/utils/get_fd_coefficients.py:4-19 (executed, unmodified)", "cases": cases},
          open(out, "w"), indent=1)
print("wrote", out)
