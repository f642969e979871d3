"""The reference's example_bittner_inference.py on the GPU path: infer a 200-gene melanoma network from genedata.xls
(shipped fit when upstream ships one for the request, else the GPU fitter), wrap it in PBNTargetEnv and estimate the
steady-state distribution of WNT5A (gene 324901).

    python examples/example_bittner_inference.py [--genes 200] [--method kmeans] [--predictors 5] [--fit]
"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))

from gym_PBN.envs import PBNTargetEnv  # noqa: E402
from gym_PBN.envs.bittner.utils import DATA, spawn  # noqa: E402
from gym_PBN.utils.eval import compute_ssd_hist  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=200)
    ap.add_argument("--method", default="kmeans")
    ap.add_argument("--predictors", type=int, default=5)
    ap.add_argument("--fit", action="store_true", help="always take the reference's route: exact cache file, else fit on the GPU")
    ap.add_argument("--iters", type=int, default=1_200_000)
    ap.add_argument("--resets", type=int, default=300)
    ap.add_argument("--cache-dir", default=str(DATA))
    args = ap.parse_args(argv)

    # Step 1 - Inference
    include_ids = [234237, 324901, 759948, 25485, 266361, 108208, 130057]
    t0 = time.time()
    graph = spawn(file=DATA / "genedata.xls", total_genes=args.genes, include_ids=include_ids, bin_method=args.method,
                  n_predictors=args.predictors, predictor_sets_path=args.cache_dir, predictor_set="fit" if args.fit else None)
    print(f"network: {graph.N} genes in {time.time() - t0:.2f} s")

    goal_config = {"target_nodes": [324901], "intervene_on": [234237], "target_node_values": ((0,),),
                   "undesired_node_values": tuple(), "horizon": 11}
    env = PBNTargetEnv(graph, goal_config, "human", name=f"Bittner-{args.genes}")

    # Step 2 - Evaluation
    t0 = time.time()
    ssd, _ = compute_ssd_hist(env, resets=args.resets, iters=args.iters)
    print(ssd)
    print(f"SSD of WNT5A from {args.iters} iterations in {time.time() - t0:.2f} s")
    return ssd


if __name__ == "__main__":
    main()
