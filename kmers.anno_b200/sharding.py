"""Host-side sharding of a CSR batch over ranks / devices: contiguous, residue-balanced ranges
aligned to sequence boundaries (the same rule ka_annotate applies across an engine's devices,
csrc/ka_engine.cu).  No collective is involved: every sequence is independent
(ApplyKmerProcessor.java:122-148) and the table is replicated; results are concatenated."""
import numpy as np


def residue_balanced_cuts(offsets, n_parts):
    """cuts[i]..cuts[i+1] = sequences of part i; every sequence belongs to exactly one part."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = offsets.shape[0] - 1
    cuts = np.zeros(n_parts + 1, dtype=np.int64)
    cuts[n_parts] = n
    if n <= 0:
        return cuts
    total = int(offsets[n] - offsets[0])
    for i in range(1, n_parts):
        target = int(offsets[0]) + total // n_parts * i
        c = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        cuts[i] = min(max(c, cuts[i - 1]), n)
    return cuts


def shard(residues, offsets, rank, world):
    """(residues view, offsets rebased to 0, first sequence index) of this rank's part."""
    cuts = residue_balanced_cuts(offsets, world)
    a, b = int(cuts[rank]), int(cuts[rank + 1])
    offs = np.asarray(offsets[a:b + 1], dtype=np.uint64)
    lo, hi = int(offs[0]), int(offs[-1])
    return residues[lo:hi], offs - np.uint64(lo), a
