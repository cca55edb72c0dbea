"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a plain-Python restatement of the reference's rooted short-cycle
finders, used to check csrc/hgen.cpp.

  cycle4_from_root  <- Matlab/Cycle_Finder_length4_fromroot.m:3-19
  cycle6_from_root  <- Matlab/Cycle_Finder_length6.m:1-76
Vlist / Clist are the script's structures: Vlist[c] = variables of check c, Clist[v] = checks of variable v (0-based
lists here instead of the padded count-first rows).
"""
from __future__ import annotations


def lists_from_csr(indptr, indices, m, n):
    vlist = [list(map(int, indices[indptr[r]:indptr[r + 1]])) for r in range(m)]
    clist = [[] for _ in range(n)]
    for r, row in enumerate(vlist):
        for v in row:
            clist[v].append(r)
    return vlist, clist


def cycle4_from_root(vlist, clist, vroot):
    # :5-8  vnodes_tent = [vroot, every other variable of every check of vroot]
    tent = [vroot]
    for c in clist[vroot]:
        tent += [v for v in vlist[c] if v != vroot]
    # :15-19 a repeated variable closes a 4-cycle
    return 1 if len(set(tent)) != len(tent) else 0


def cycle6_from_root(vlist, clist, vroot):
    if cycle4_from_root(vlist, clist, vroot):       # :72-74 "There is a 4-cycle"
        return 1
    # :33-52 tier-1 variables with the check they were reached through
    tier1 = [(v, c) for c in clist[vroot] for v in vlist[c] if v != vroot]
    # :53-62 tier-2 checks: every check of a tier-1 variable except the one it was reached by
    tier2 = []
    for v, c_from in tier1:
        tier2 += [c for c in clist[v] if c != c_from]
    # :63-69 a duplicate among them closes a 6-cycle
    tier2.sort()
    return 1 if any(a == b for a, b in zip(tier2, tier2[1:])) else 0


def count_short_cycles(vlist, clist):
    n = len(clist)
    return (sum(cycle4_from_root(vlist, clist, v) for v in range(n)),
            sum(cycle6_from_root(vlist, clist, v) for v in range(n)))
