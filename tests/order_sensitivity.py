#!/usr/bin/env python
"""How much does the kinetic energy of the dense pile depend on the ORDER in which pass 2 applies the pair impulses?  (CPU only.)

    python tests/order_sensitivity.py [seed ...]      # output kept in profiles/r02_order_sensitivity.txt

20,000-particle polydisperse pile (BASELINE configs[2] at small size), 1000 substeps, kinetic energy at substeps 100 / 250 / 500 /
1000 relative to the compiled reference as built (oracle/_ref/libref.so), for
  * the reference itself with its pair list reversed / shuffled (oracle/ref_shim.c),
  * the tile Gauss-Seidel order of "resolve"=3 (oracle/ppc_oracle.c:ppc_oracle_tilegs_order: pairs visited in nine phases, run by
    greedy pair-colour), which the device reproduces bit for bit (tests/test_gpu_gs.py),
  * the same pairs run in the visiting order itself (what round 2 first shipped; PPC_ORACLE_GS_VISITING_ORDER=1).
This is test infrastructure: it runs the oracle, not the product.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as om  # noqa: E402
from oracle.oracle import Oracle, State  # noqa: E402
from particlephysicsc_b200 import scenes  # noqa: E402

CHECKPOINTS = (100, 250, 500, 1000)
DT = 1.0 / 1152.0


def main():
    o, R = Oracle(), om.Reference()
    for seed in [int(a) for a in sys.argv[1:]] or [104, 105, 106]:
        sc = scenes.polydisperse_pile(20000, seed=seed, vmax=5.0)
        runs = {}

        def run(label, stepf):
            s = State(sc.px.copy(), sc.py.copy(), sc.vx.copy(), sc.vy.copy(), sc.mass.copy(), sc.radius.copy(), sc.rgba.copy())
            out = []
            for k in range(CHECKPOINTS[-1]):
                stepf(s)
                if k + 1 in CHECKPOINTS:
                    out.append(o.stats(s, sc.W, sc.H)["ke"])
            runs[label] = out

        o.set_tilegs(26, 8, 1024)
        run("reference as built", lambda s: R.substep(s, sc.W, sc.H, DT))
        for label, mode in (("reference, list reversed", 1), ("reference, list shuffled", 2)):
            R.set_pair_order(mode, 7)
            try:
                run(label, lambda s: R.substep(s, sc.W, sc.H, DT))
            finally:
                R.set_pair_order(0)
        os.environ.pop("PPC_ORACLE_GS_VISITING_ORDER", None)
        run("tile Gauss-Seidel, greedy pair-colours (resolve=3)", lambda s: o.substep(s, sc.W, sc.H, DT, 3, False))
        os.environ["PPC_ORACLE_GS_VISITING_ORDER"] = "1"
        run("tile Gauss-Seidel, visiting order", lambda s: o.substep(s, sc.W, sc.H, DT, 3, False))
        os.environ.pop("PPC_ORACLE_GS_VISITING_ORDER", None)
        base = runs["reference as built"]
        print(f"seed {seed}: kinetic energy of the reference as built at substeps {CHECKPOINTS}: {[f'{x:.4g}' for x in base]}")
        for label, v in runs.items():
            if label != "reference as built":
                print(f"  {label:52s}", [f"{100 * (x - y) / y:+.2f}%" for x, y in zip(v, base)])


if __name__ == "__main__":
    main()
