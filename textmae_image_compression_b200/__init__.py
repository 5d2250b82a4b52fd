"""B200-native compression forward path of TextMAE.

Public surface:
    MCM                 drop-in module for the reference's `MCM` on the encode + rate path
    PathConfig          geometry derived from the reference constructor arguments
    make_state_dict     seeded synthetic checkpoint with the reference's parameter names
    distributed         image sharding + the scalar rate all-reduce for N GPUs
    scores              GPU patch-score generation (the reference's generate_scores_file.py)
    huffman             HuffmanCoding: the reference's ids_restore side-information coder (utils/huffman.py), host side
"""
from .config import PathConfig, vit_base, vit_large
from .synthetic import make_state_dict

__all__ = ["MCM", "PathConfig", "vit_base", "vit_large", "make_state_dict"]


def __getattr__(name):
    if name == "MCM":
        from .mcm import MCM
        return MCM
    raise AttributeError(name)
