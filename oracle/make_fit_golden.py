"""TEST INFRASTRUCTURE — records golden vectors of the reference's gene-data preparation (runs only in the build
container, where /root/reference exists):  python oracle/make_fit_golden.py  ->  tests/golden/fit_prepare.npz

  * ids / ratios of the 85 spreadsheet rows of the 70 padded ids (tests/test_bittner.py:21-40), read here with OUR xls
    reader (the image has no xlrd, so pandas.read_excel — the reference's reader — cannot run; its expected shapes and the
    padded id list are the constants of the reference's own tests);
  * the output of the UNMODIFIED reference binarise() (gym_PBN/envs/bittner/gen/binarise.py) for "median", "average" and
    "kmeans" on those rows, with the legacy NumPy stream seeded to 0, and the stream position it leaves behind.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gym-pbn-stac_b200"))

spec = importlib.util.spec_from_file_location("ref_binarise", "/root/reference/gym_PBN/envs/bittner/gen/binarise.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

from gym_PBN.envs.bittner import utils  # noqa: E402

gene_data, weight_ids = utils.extract_gene_data(utils.DATA / "genedata.xls")
ids70 = utils.pad_ids([234237, 324901, 759948, 25485, 266361, 108208, 130057], 70, weight_ids)
trimmed = gene_data.loc[ids70]
out = dict(ids=np.asarray(trimmed.index), ratios=trimmed.drop("Name", axis=1).to_numpy(), padded=np.array(ids70))
for method in ("median", "average", "kmeans"):
    np.random.seed(0)
    b = ref.binarise(trimmed, method)
    out[method] = b.drop("Name", axis=1).to_numpy().astype(np.uint8)
    out[method + "_next_draw"] = np.random.rand(1)
np.savez_compressed(ROOT / "tests" / "golden" / "fit_prepare.npz", **out)
print({k: v.shape for k, v in out.items()})
