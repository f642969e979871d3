"""Writes the attractor fixtures the CPU arm of bench.py needs (the oracle cannot run the product's device search): the exact
attractors of Bittner-28 and the sampled + verified attractors of Bittner-200, exactly as bench.py --config 2 / 4 computes
them.  Run on a GPU:  python tools/make_attractor_fixtures.py gpurun_out/  -> copy the two JSON files into tests/golden/."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))
from gym_PBN.b200 import attractors, compiler, engine  # noqa: E402

out = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
out.mkdir(exist_ok=True)
net = engine.Network(compiler.load_bittner("28_15_median"))
a28 = attractors.exact_attractor_cubes(net)
(out / "b28_exact_attractors.json").write_text(json.dumps([[list(c) for c in a] for a in a28]))
print("b28:", [len(a) for a in a28])
net = engine.Network(compiler.load_bittner("200_5_kmeans"))
a200, info = attractors.verified_attractors(net, resets=512, seed=0)
(out / "b200_verified_attractors.json").write_text(json.dumps([[list(c) for c in a] for a in a200]))
print("b200:", [(i["free"], i["method"]) for i in info])
for name in ("100_5_kmeans", "70_5_kmeans"):
    net = engine.Network(compiler.load_bittner(name))
    atts, info = attractors.verified_attractors(net, resets=512, seed=0)
    print(name, [(i["free"], i["states"]) for i in info])
