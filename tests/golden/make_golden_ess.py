"""Generate tests/golden/ess_chains.npz from the UNMODIFIED reference's `_ess_from_chain`
(benchmarks/registry.py:348-365 of /root/reference).  Run in the container that mounts /root/reference:

    python tests/golden/make_golden_ess.py

The function is extracted from the reference source text and executed as is (importing benchmarks/registry.py would
pull in the whole benchmark harness); only the input chains and the returned numbers are stored."""

import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/benchmarks/registry.py"


def reference_fn():
    tree = ast.parse(open(SRC).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_ess_from_chain")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    return ns["_ess_from_chain"]


def ar1(n, phi, gen):
    e = torch.randn(n, generator=gen)
    x = torch.empty(n)
    x[0] = e[0]
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    return x


def main():
    ess = reference_fn()
    gen = torch.Generator().manual_seed(77)
    chains = {
        "white_1000": torch.randn(1000, generator=gen),
        "ar1_05_1000": ar1(1000, 0.5, gen),
        "ar1_09_1000": ar1(1000, 0.9, gen),
        "ar1_099_2000": ar1(2000, 0.99, gen),
        "ar1_neg_500": ar1(500, -0.6, gen),
        "trend_300": torch.linspace(0, 1, 300) + 0.01 * torch.randn(300, generator=gen),
        "constant_64": torch.full((64,), 3.5),
        "two_2": torch.tensor([1.0, -2.0]),
        "one_1": torch.tensor([4.0]),
        "short_7": torch.randn(7, generator=gen),
        "energy_like_250": (ar1(250, 0.8, gen) * 0.3 + 12.0),
    }
    out = {}
    for name, c in chains.items():
        out["chain_" + name] = c.numpy()
        out["ess_" + name] = np.float64(ess(c))
        print(name, out["ess_" + name])
    np.savez_compressed(os.path.join(HERE, "ess_chains.npz"), **out)


if __name__ == "__main__":
    main()
