"""Regenerates tests/golden/journals.json from the oracle (python tests/golden/make_golden.py).
The oracle itself is pinned to the reference's documented known answers by
tests/test_oracle_golden.py and tests/golden/reference_vectors.json; these fixtures freeze its
journals so that (a) an accidental change of the spec shows up as a diff here and (b) the GPU
suite can check the engine without the oracle being built."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_cases as G   # noqa: E402
import oracle_lib as O     # noqa: E402


def main():
    out = {}
    for name, (_, fn) in G.CASES.items():
        o = G.make_oracle(name)
        fn(o, O.body)
        ev, bd = o.journal()
        out[name] = G.digest(ev, bd, o.stats(), o.now, o.round)
    with open(os.path.join(HERE, "journals.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
