"""CPU oracle of the k-mer annotation hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (kmers.anno_b200/, libkmeranno.so) never does.
PARITY UNPINNED: see the header of ka_oracle.c.
"""
from .binding import OracleDb, FastDb, count_probes, kmer_distance_pairs, lib, LIB_PATH  # noqa: F401
