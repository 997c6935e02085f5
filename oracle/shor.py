"""Shor 2x2-minor index enumeration (bit-exact integer work).  TEST INFRASTRUCTURE ONLY.

Restates generate_rank1_matrix_completion_Shor_constraints_indexes (OMC.jl:2545-2612): for each
``num_entries_present`` IN THE ORDER GIVEN, for (i1 < i2) lexicographic, inner loops as coded;
tuples are 1-based (i1, i2, j1, j2) with j1 < j2, exactly as Julia would produce them.
"""
import itertools
import numpy as np


def shor_constraint_indexes(indices, num_entries_present_list):
    ind = np.asarray(indices, dtype=bool)
    n, m = ind.shape
    out = []
    for p in num_entries_present_list:
        for i1, i2 in itertools.combinations(range(n), 2):
            r1, r2 = ind[i1], ind[i2]
            both = np.flatnonzero(r1 & r2); xor = np.flatnonzero(r1 ^ r2); none = np.flatnonzero(~(r1 | r2))
            if p == 4:                                              # OMC.jl:2555-2560
                for j1, j2 in itertools.combinations(both, 2):
                    out.append((i1 + 1, i2 + 1, int(j1) + 1, int(j2) + 1))
            elif p == 3:                                            # OMC.jl:2561-2568
                for j1 in both:
                    for j2 in xor:
                        a, b = sorted((int(j1), int(j2)))
                        out.append((i1 + 1, i2 + 1, a + 1, b + 1))
        if p == 2:
            for i1, i2 in itertools.combinations(range(n), 2):      # (a) OMC.jl:2571-2577
                r1, r2 = ind[i1], ind[i2]
                for j1 in np.flatnonzero(r1 & r2):
                    for j2 in np.flatnonzero(~(r1 | r2)):
                        a, b = sorted((int(j1), int(j2)))
                        out.append((i1 + 1, i2 + 1, a + 1, b + 1))
            for i1, i2 in itertools.combinations(range(n), 2):      # (b) OMC.jl:2579-2583
                r1, r2 = ind[i1], ind[i2]
                for j1, j2 in itertools.combinations(np.flatnonzero(r1 ^ r2), 2):
                    out.append((i1 + 1, i2 + 1, int(j1) + 1, int(j2) + 1))
        elif p == 1:                                                # OMC.jl:2584-2595
            for i1, i2 in itertools.combinations(range(n), 2):
                r1, r2 = ind[i1], ind[i2]
                none = np.flatnonzero(~(r1 | r2))
                for j1 in range(m):
                    if int(r1[j1]) + int(r2[j1]) == 1:
                        for j2 in none:
                            a, b = sorted((j1, int(j2)))
                            out.append((i1 + 1, i2 + 1, a + 1, b + 1))
        elif p == 0:                                                # OMC.jl:2596-2607
            for i1, i2 in itertools.combinations(range(n), 2):
                r1, r2 = ind[i1], ind[i2]
                for j1 in range(m - 1):
                    if not r1[j1] and not r2[j1]:
                        for j2 in np.flatnonzero(~(r1[j1 + 1:] | r2[j1 + 1:])):
                            out.append((i1 + 1, i2 + 1, j1 + 1, j1 + 1 + int(j2) + 1))
    return out


def shor_brute_force(indices, p):
    """All (i1<i2, j1<j2) minors with exactly p observed entries, as a set (order-free cross-check)."""
    ind = np.asarray(indices, dtype=bool)
    n, m = ind.shape
    s = set()
    for i1, i2 in itertools.combinations(range(n), 2):
        for j1, j2 in itertools.combinations(range(m), 2):
            if int(ind[i1, j1]) + int(ind[i1, j2]) + int(ind[i2, j1]) + int(ind[i2, j2]) == p:
                s.add((i1 + 1, i2 + 1, j1 + 1, j2 + 1))
    return s


def violated_minors(X, candidates, existing, n_minors):
    """generate_violated_Shor_minors (OMC.jl:2614-2640) after the candidate list is built: X (k, n, m); candidates / existing are
    0-based (i1, i2, j1, j2) tuples.  Score sum_t |X[t,i1,j1] X[t,i2,j2] - X[t,i1,j2] X[t,i2,j1]| (OMC.jl:2628-2629), candidates in
    `existing` removed (setdiff!, OMC.jl:2625), the n_minors largest by (score, tuple), descending (OMC.jl:2634-2639)."""
    X = np.asarray(X, float)
    ex = {tuple(int(v) for v in t) for t in existing}
    out = []
    for t in candidates:
        i1, i2, j1, j2 = (int(v) for v in t)
        if (i1, i2, j1, j2) in ex:
            continue
        sc = 0.0
        for s_ in range(X.shape[0]):
            sc = sc + abs(X[s_, i1, j1] * X[s_, i2, j2] - X[s_, i1, j2] * X[s_, i2, j1])
        out.append((sc, (i1, i2, j1, j2)))
    out.sort(reverse=True)
    return out[:n_minors]
