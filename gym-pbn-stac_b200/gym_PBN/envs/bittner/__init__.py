"""Bittner melanoma predictor-graph networks (device-backed) and the shipped predictor sets."""
