"""The pair order of the sample-batched bilinear backward (csrc/fibinet.cu: rr_shape / rr_pair) restated in Python: a round-robin
tournament (circle method).  The kernel relies on two properties -- every pair of combinations(range(F-1), 2) is visited exactly
once, and the pairs of one round are field-disjoint (one writer per shared-memory element per round) -- checked here for every
field count the C ABI accepts; the device code itself is covered by the GPU parity tests and compute-sanitizer racecheck."""
import itertools

import pytest


def rr_shape(F):
    n = F - 1
    np_ = n + 1 if n % 2 else n
    slot0 = 1 if n % 2 else 0                     # odd n: slot 0 would meet the dummy field
    return n, np_, np_ - 1, np_ // 2 - slot0, slot0


def rr_pair(F, r, slot):
    n, np_, rounds, slots, slot0 = rr_shape(F)
    m1, sl = np_ - 1, slot + slot0
    a, b = (r, np_ - 1) if sl == 0 else ((r + sl) % m1, (r - sl + m1) % m1)
    return (a, b) if a < b else (b, a)


@pytest.mark.parametrize("F", list(range(3, 66)) + [100, 129, 256])
def test_tournament_covers_every_pair_once_with_disjoint_rounds(F):
    n, _, rounds, slots, _ = rr_shape(F)
    seen = []
    for r in range(rounds):
        fields = []
        for s in range(slots):
            i, j = rr_pair(F, r, s)
            assert 0 <= i < j < n                  # never the dummy, never field F-1
            fields += [i, j]
            seen.append((i, j))
        assert len(set(fields)) == len(fields)     # field-disjoint inside a round
    assert sorted(seen) == list(itertools.combinations(range(n), 2))
    assert rounds * slots == n * (n - 1) // 2      # the kernels size the pair table with this identity
