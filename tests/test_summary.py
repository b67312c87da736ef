"""The accuracy report (graphchainer_b200/summary.py = the table of the reference's scripts/summary.py): its edit distance against
the textbook dynamic programme, and the table of the golden `tiny` case (reference GAM and ours give the same rows by construction
of the GAM parity tests; here the reference's GAM is the input)."""
import os
import random

from conftest import GOLDEN
from graphchainer_b200 import summary


def _dp(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def test_edit_distance_matches_dynamic_programme():
    rng = random.Random(5)
    cases = [("", ""), ("A", ""), ("", "ACG"), ("ACGT", "ACGT"), ("A" * 70, "A" * 64), ("ACGT" * 40, "TGCA" * 37)]
    for _ in range(60):
        n, m = rng.randrange(0, 200), rng.randrange(0, 200)
        a = "".join(rng.choice("ACGT") for _ in range(n))
        b = list(a[:m]) + [rng.choice("ACGT") for _ in range(rng.randrange(0, 30))]
        for _ in range(rng.randrange(0, 25)):
            if b:
                b[rng.randrange(len(b))] = rng.choice("ACGT")
        cases.append((a, "".join(b)))
    for a, b in cases:
        assert summary.edit_distance(a, b) == _dp(a, b), (a, b)


def test_summary_table_of_golden_tiny():
    rows = summary.table(os.path.join(GOLDEN, "tiny.gfa"), os.path.join(GOLDEN, "tiny.fa"), os.path.join(GOLDEN, "tiny.gam"))
    assert rows[0][:3] == ["name", "length", "long_pathcnt"] and len(rows) == 9
    acc = summary.accuracy(rows)
    assert acc["reads"] == 8 and acc["aligned"] == 8
    # simulated at 15 % error: the path spells the read's origin, whole first / last nodes included
    assert 0.9 < acc["mean_align_rate"] < 1.2
    assert 0.05 < acc["edit_distance_per_read_base"] < 0.3
    for r in rows[1:]:
        assert int(r[2]) > 0 and int(r[3]) > 0 and r[9] != ""
