"""Seeded input builders shared by the CPU and GPU tests."""
import numpy as np

ALPHA20 = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def csr(seqs):
    """list of bytes -> (residues u8, offsets u64)"""
    offs = np.zeros(len(seqs) + 1, np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    res = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return res, offs


def random_seq(rng, L, alphabet=ALPHA20):
    return bytes(alphabet[rng.integers(0, len(alphabet), size=L)])


def ragged_case(seed, n_seq=300, K=8, max_len=700, n_roles=12, db_frac=0.3, odd_bytes=True,
                lengths=None):
    """Random sequences + a DB sampled from their own windows.

    Each sequence has a 'home' role; DB k-mers sampled from it get that role except for a
    fraction given another role (ambiguity), DB lines are duplicated with different roles
    (last line wins), some sequences carry tandem repeats (duplicate k-mers), bytes outside
    the DB alphabet ('X', '*', lowercase), and there are empty and shorter-than-K sequences.
    Returns (seqs, kmers_bytes_list, roles)."""
    rng = np.random.default_rng(seed)
    seqs = []
    for i in range(n_seq):
        if lengths is not None:
            L = int(lengths[i])
        else:
            r = rng.random()
            if r < 0.05:
                L = 0
            elif r < 0.12:
                L = int(rng.integers(0, K))          # shorter than K: no windows
            elif r < 0.2:
                L = int(rng.integers(K, K + 3))      # one to three windows
            else:
                L = int(rng.integers(K, max_len))
        s = bytearray(random_seq(rng, L))
        if L > 6 * K and rng.random() < 0.3:         # tandem repeat -> duplicate k-mers
            u = int(rng.integers(K + 1, max(K + 2, min(3 * K, L // 3))))
            st = int(rng.integers(0, L - 2 * u))
            s[st + u: st + 2 * u] = s[st: st + u]
            if rng.random() < 0.5 and st + 3 * u <= L:
                s[st + 2 * u: st + 3 * u] = s[st: st + u]
        if odd_bytes and L > K and rng.random() < 0.25:  # bytes the DB never contains
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(0, L))] = int(rng.choice(list(b"X*alv-")))
        if L > 4 * K and rng.random() < 0.1:         # low complexity run
            st = int(rng.integers(0, L - 2 * K))
            s[st: st + 2 * K] = bytes([s[st]]) * (2 * K)
        seqs.append(bytes(s))
    kmers, roles = [], []
    for i, s in enumerate(seqs):
        L = len(s)
        if L < K:
            continue
        home = int(rng.integers(0, n_roles))
        mode = rng.random()
        if mode < 0.25:
            continue                                  # sequence without hits
        n_take = max(1, int((L - K + 1) * db_frac * rng.random()))
        for p in rng.integers(0, L - K + 1, size=n_take):
            km = s[p: p + K]
            if any(c not in b"ACDEFGHIKLMNPQRSTVWY" for c in km):
                continue
            role = home
            if mode > 0.85 and rng.random() < 0.2:
                role = int(rng.integers(0, n_roles))  # another role: ambiguous peg
            kmers.append(km)
            roles.append(role)
    # duplicate DB lines with a different role: the LAST line must win
    n = len(kmers)
    for j in rng.integers(0, max(n, 1), size=n // 10):
        kmers.append(kmers[int(j)])
        roles.append(int(rng.integers(0, n_roles)))
    if not kmers:
        kmers, roles = [b"A" * K], [0]
    return seqs, kmers, np.asarray(roles, np.int32)


def py_apply(seqs, kmers, roles, K, min_hits, distinct=True, include_last=True):
    """Pure-Python statement of ApplyKmerProcessor.java:122-148 (dict + set), tiny cases only."""
    db = {}
    for k, r in zip(kmers, roles):
        db[bytes(k)] = int(r)                         # put(): last wins
    out_role, out_hits, out_flag = [], [], []
    for s in seqs:
        n = len(s) - K + (1 if include_last else 0)
        wins = [s[i: i + K] for i in range(max(n, 0))]
        if distinct:
            wins = set(wins)
        hit_roles = [db[w] for w in wins if w in db]
        if not hit_roles:
            r, h, f = -1, 0, 0
        elif len(set(hit_roles)) > 1:
            r, h, f = -1, 0, 2
        elif len(hit_roles) >= min_hits:
            r, h, f = hit_roles[0], len(hit_roles), 1
        else:
            r, h, f = -1, len(hit_roles), 3
        out_role.append(r); out_hits.append(h); out_flag.append(f)
    return (np.asarray(out_role, np.int32), np.asarray(out_hits, np.int32), np.asarray(out_flag, np.uint8))
