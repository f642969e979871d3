"""Offline network-inference pipeline of the Bittner envs: spreadsheet -> binarised expression -> best predictor triples.
(reference: gym_PBN/envs/bittner/gen/ and bittner/utils.py).  Host-side NumPy: it runs once per
network size and its output is cached as a predictor-set pickle in the reference's own format."""
from .binarise import binarise  # noqa: F401
from .predictor_sets import generate_predictor_sets  # noqa: F401
from .xls import read_gene_data  # noqa: F401
