"""Architecture descriptions of NPPNet — data only, semantics identical to the reference's
models/genotypes.py:4-54 (namedtuple field names and the op/index lists are the contract that
`Network` constructors consume; nothing here executes on the GPU)."""
from collections import namedtuple

Genotype = namedtuple("Genotype", "normal normal_concat reduce reduce_concat")
Genotype_up2 = namedtuple("Genotype_up2", "upsample1 upsample_concat1 upsample2 upsample_concat2")
Genotype_inter = namedtuple("Genotype_inter", "task1 task2 task3 task4")
Genotype_fuse = namedtuple("Genotype_fuse", "pose pose_concat par par_concat")

# candidate sets of the encoder/decoder search and of the interaction search (genotypes.py:10-28)
PRIMITIVES_PC = ["std_conv_3x3", "se_connect", "dil_conv_3x3_4", "dil_conv_3x3_2", "std_conv_1x1", "max_pool_3x3",
                 "skip_connect"]
PRIMITIVES_INTER = ["std_conv_3x3", "dil_conv_3x3_4", "se_connect", "max_pool_3x3", "dil_conv_3x3_2", "std_conv_1x1",
                    "poled_conv_x1"]


def _edges(spec):
    """'op@idx op@idx ...' -> [(op, idx), ...]"""
    out = []
    for tok in spec.split():
        op, idx = tok.rsplit("@", 1)
        out.append((op, int(idx)))
    return out


_C3, _C1, _SE, _MP = "std_conv_3x3", "std_conv_1x1", "se_connect", "max_pool_3x3"
_D2, _D4, _PC = "dil_conv_3x3_2", "dil_conv_3x3_4", "poled_conv_x1"

# genotypes.py:30-33
ENCODER = Genotype(
    normal=_edges(f"{_C3}@0 {_SE}@1 {_SE}@1 {_C3}@0 {_MP}@1 {_C3}@2 {_C3}@3 {_C3}@0"),
    normal_concat=range(2, 6),
    reduce=_edges(f"{_C3}@0 {_SE}@1 {_SE}@1 {_C3}@2 {_D4}@3 {_D4}@2 {_MP}@3 {_D2}@0"),
    reduce_concat=range(2, 6))

# genotypes.py:35-38
DECODER = Genotype_up2(
    upsample1=_edges(f"{_C1}@1 {_C1}@0 {_C1}@1 {_C3}@0 {_C1}@0 {_D2}@1 {_C3}@3 {_C1}@1"),
    upsample_concat1=range(2, 6),
    upsample2=_edges(f"{_C3}@1 {_SE}@0 {_D2}@2 {_C1}@1 {_PC}@3 {_C1}@2 {_C3}@1 {_C1}@2"),
    upsample_concat2=range(2, 6))

# genotypes.py:40-49
INTER = Genotype_inter(
    task1=[_edges(f"{_D2}@0"), _edges(f"{_C3}@1"), _edges(f"{_C1}@1 {_C3}@2"), _edges(f"{_C1}@2 {_C3}@3")],
    task2=[_edges(f"{_D2}@0"), _edges(f"{_PC}@1"), _edges(f"{_C1}@2"), _edges(f"{_C3}@1 {_C3}@3")],
    task3=[_edges(f"{_D2}@4 {_D2}@2 {_D2}@1"), _edges(f"{_C3}@1 {_C3}@2 {_D2}@5 {_D2}@0"),
           _edges(f"{_C3}@1 {_D2}@2 {_D4}@5 {_D2}@3")],
    task4=[_edges(f"{_C3}@0"), _edges(f"{_C3}@1"), _edges(f"{_C1}@2 {_C3}@1")])

# genotypes.py:51-54
FUSION = Genotype_fuse(
    pose=_edges(f"{_C3}@1 {_C3}@2 {_C3}@0 {_MP}@2 {_C3}@4 {_C3}@2 {_C3}@4 {_C3}@3"),
    pose_concat=range(3, 7),
    par=_edges(f"{_D2}@2 {_SE}@1 {_D2}@2 {_D2}@3 {_MP}@3 {_C3}@2 {_D2}@5 {_C3}@2"),
    par_concat=range(3, 7))
