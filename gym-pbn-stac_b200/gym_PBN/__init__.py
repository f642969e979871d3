"""gym_PBN — B200-native drop-in for jakub-zarzycki2022/gym-PBN-stac (same import name, same env ids)."""
__version__ = "0.1.0"
