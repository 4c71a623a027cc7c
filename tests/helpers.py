"""Shared test helpers: small random / adversarial workloads and oracle runners."""
import random

import numpy as np

import synth
from metalign_b200 import codec
from oracle import oracle_py
from oracle.oracle_c import OracleDB, OracleQuery


def oracle_c_run(keys, G, n, K, ks, push, ci_min=2, gate="exact", count_empty_in_den=True):
    """push(q) feeds reads into an OracleQuery; returns (result dict, sorted I as (m,2) uint64)."""
    db = OracleDB(keys, G, n, K, ks)
    q = OracleQuery(db, ci_min, gate, count_empty_in_den)
    push(q)
    r = q.finish()
    I = q.intersection()
    q.close()
    db.close()
    return r, I


def adversarial_case(rng: random.Random, K=None):
    """Tiny workload over one short genome: overlapping sketch k-mers, both orientations, empty slots,
    shared k-mers, reads with N / lower case / short reads."""
    K = K or rng.choice([8, 12, 16, 21])
    nk = rng.choice([1, 2, 3])
    ks = sorted(rng.sample(range(max(2, K - 12), K + 1), nk))
    if rng.random() < 0.6:
        ks[-1] = K
    ks = sorted(set(ks))
    G, n = rng.randint(1, 6), rng.randint(1, 7)
    glen = rng.choice([40, 80, 200])
    genome = "".join(rng.choice("ACGT") for _ in range(max(glen, K + 5)))
    sketches = []
    for _ in range(G):
        sk = []
        for _ in range(n):
            if rng.random() < 0.15:
                sk.append("")
            else:
                p = rng.randint(0, len(genome) - K)
                km = genome[p:p + K]
                if rng.random() < 0.5:
                    km = oracle_py.rc(km)
                sk.append(km)
        sketches.append(sk)
    reads = []
    for _ in range(rng.randint(0, 60)):
        a = rng.randint(0, len(genome) - 1)
        b = rng.randint(a, len(genome))
        r = genome[a:b]
        if rng.random() < 0.5:
            r = oracle_py.rc(r)
        r = list(r)
        for i in range(len(r)):
            x = rng.random()
            if x < 0.02:
                r[i] = "N"
            elif x < 0.04:
                r[i] = r[i].lower()
        reads.append("".join(r))
    return dict(K=K, ks=ks, sketches=sketches, reads=reads)


def synth_small(G=24, n=40, nreads=4000, seed=7, **kw):
    p = synth.params(G=G, n=n, seed=seed, len_min=4000, len_max=12000, n_present=min(8, G), **kw)
    keys = synth.sketch_keys(p)
    return p, keys
