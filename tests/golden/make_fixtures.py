"""Generates tests/golden/nist_spce.npz from the DATA files the reference bundles.

Run in the build container only (needs /root/reference; the GPU box has no copy):
    python tests/golden/make_fixtures.py

Input  (data, not source): /root/reference/Ewald/spce_sample_config_periodic{1..4}.txt
        and Ewald/coord750.txt (CRLF copy of config 4) — NIST SPC/E reference
        configurations, read the way Ewald/initialConfigurations.jl:282-355 (ReadNIST)
        reads them: line 1 = box lengths, line 2 = molecule count, then
        "index x y z element" rows in O,H,H order.
Output: float64 arrays, bit-exact as parsed by Python's float():
        box{c} (scalar), xyz{c} (n_sites x 3) for c in 1..4 and xyz750/box750.
"""
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/Ewald")
OUT = Path(__file__).resolve().parent / "nist_spce.npz"


def read_nist(path):
    box, xyz, el = None, [], []
    for i, line in enumerate(path.read_text().splitlines(), start=1):
        t = line.split()
        if i == 1:
            box = float(t[0])
        if len(t) > 2 and i > 2:
            xyz.append([float(t[1]), float(t[2]), float(t[3])])
            el.append(t[4])
    assert el[0::3] == ["O"] * (len(el) // 3) and set(el[1::3] + el[2::3]) == {"H"}
    return box, np.array(xyz, dtype=np.float64)


def main():
    if not REF.exists():
        sys.exit("reference not mounted; fixtures are already committed")
    out = {}
    for c in (1, 2, 3, 4):
        box, xyz = read_nist(REF / f"spce_sample_config_periodic{c}.txt")
        out[f"box{c}"] = np.float64(box)
        out[f"xyz{c}"] = xyz
    box, xyz = read_nist(REF / "coord750.txt")
    assert box == out["box4"] and np.array_equal(xyz, out["xyz4"]), "coord750 != config 4"
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: getattr(v, "shape", None) for k, v in out.items()})
    # TIP3P model parameters, parsed from the reference's own input files (repository root: water.top, tip3p.pdb —
    # what Ewald/main.jl:156-157 opens) by the product's readers; numbers only, no reference text is copied
    import json
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
    from metropolismontecarlo_b200.systems import model_from_files
    model = model_from_files(REF.parent / "water.top", REF.parent / "tip3p.pdb", "WAT")
    (OUT.parent / "tip3p_model.json").write_text(json.dumps(model, indent=1) + "\n")
    print("wrote tip3p_model.json", model)


if __name__ == "__main__":
    main()
