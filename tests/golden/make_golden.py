#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/ (run in the authoring container only:
it reads the reference's fixture /root/reference/src/test/small.gto, which does not exist
on the GPU box).

ORACLE-DERIVED, REFERENCE-UNPINNED: the reference holds no kmerdb.tbl / roles.in.use /
apply-output fixture and cannot be executed here (no JVM), so the expected outputs below
are produced by oracle/ (the Java-shaped restatement) and pin OUR semantics — they catch
regressions and oracle/GPU/CLI disagreement, they do not prove equality with the JVM.

Outputs
  small_proteins.tsv   fid <TAB> function <TAB> protein, the 712 pegs of small.gto in file order
  small.roles.in.use   role ids (column 1) + role names, the `roles.in.use` of the run
  small.kmerdb.tbl     `build` output (BuildKmerProcessor.java:212-216): kmer <TAB> roleId, HashMap order
  small.verify.tsv     `apply --format VERIFY` output (VerifyApplyKmerReporter.java:33-45)
  small.apply.tsv      `apply` default output (DefaultApplyKmerReporter.java:51-55)
  small.expected.npz   per-peg role index / hits / flag arrays for the GPU parity test
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import binding  # noqa: E402

GTO = "/root/reference/src/test/small.gto"
K = 8
MIN_HITS = 5


def roles_of_function(fun):
    """SEEDtk Feature.rolesOfFunction as recalled: strip the comment after '#' / '!', split
    on ' / ', ' @ ' and '; '."""
    fun = re.split(r"\s*[#!]", fun, maxsplit=1)[0]
    return [r.strip() for r in re.split(r"\s+/\s+|\s+@\s+|;\s+", fun) if r.strip()]


def main():
    g = json.load(open(GTO))
    genome_id = g["id"]
    pegs = [(f["id"], f.get("function", ""), f["protein_translation"]) for f in g["features"]
            if f.get("protein_translation") and ".peg." in f["id"]]
    with open(os.path.join(HERE, "small_proteins.tsv"), "w") as fh:
        fh.write(f"#genome_id\t{genome_id}\n")
        for fid, fun, prot in pegs:
            fh.write(f"{fid}\t{fun}\t{prot}\n")

    # hand-made role map: every role name that occurs in >= 2 pegs, 'hypothetical protein' out
    counts = {}
    for _, fun, _ in pegs:
        for r in roles_of_function(fun):
            counts[r] = counts.get(r, 0) + 1
    names = [r for r, c in counts.items() if c >= 2 and r.lower() != "hypothetical protein"]
    role_id = {name: f"Role{idx + 1:04d}" for idx, name in enumerate(names)}
    ids = [role_id[n] for n in names]
    with open(os.path.join(HERE, "small.roles.in.use"), "w") as fh:
        for n in names:
            fh.write(f"{role_id[n]}\t{n}\n")

    # build (BuildKmerProcessor.java:138-223) over this one genome
    seqs = [p.encode("latin-1") for _, _, p in pegs]
    offsets = np.zeros(len(seqs) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    residues = np.frombuffer(b"".join(seqs), np.uint8)
    n_roles = np.zeros(len(pegs), np.int32)
    peg_role = np.full(len(pegs), -1, np.int32)
    for i, (_, fun, _) in enumerate(pegs):
        good = [r for r in roles_of_function(fun) if r in role_id]
        n_roles[i] = len(good)
        if len(good) == 1:
            peg_role[i] = names.index(good[0])
    kmers, roles, stats = binding.build_db(residues, offsets, n_roles, peg_role, K, len(names))
    print("build:", stats, "roles", len(names))
    with open(os.path.join(HERE, "small.kmerdb.tbl"), "w") as fh:
        for j in range(len(roles)):
            fh.write(kmers[j * K:(j + 1) * K].tobytes().decode("latin-1") + "\t" + ids[roles[j]] + "\n")

    # apply (ApplyKmerProcessor.java:114-155)
    db = oracle.OracleDb(kmers, roles, K)
    role, hits, flag = db.apply(residues, offsets, MIN_HITS)
    with open(os.path.join(HERE, "small.verify.tsv"), "w") as fh:
        fh.write("genome_id\tpeg_id\trole\thits\tfunction\n")
        for i, (fid, fun, _) in enumerate(pegs):
            if flag[i] == 1:
                fh.write(f"{genome_id}\t{fid}\t{ids[role[i]]}\t{hits[i]}\t{fun}\n")
    vec = np.zeros(len(names), np.int64)
    for i in range(len(pegs)):
        if flag[i] == 1:
            vec[role[i]] += 1
    with open(os.path.join(HERE, "small.apply.tsv"), "w") as fh:
        fh.write(genome_id + "\t" + "\t".join(str(int(x)) for x in vec) + "\n")
    np.savez_compressed(os.path.join(HERE, "small.expected.npz"), role=role, hits=hits, flag=flag)
    print("apply: flags", np.bincount(flag, minlength=4), "kmers", len(roles))


if __name__ == "__main__":
    main()
